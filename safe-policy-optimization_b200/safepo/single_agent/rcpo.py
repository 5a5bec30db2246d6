"""RCPO (reward-constrained policy optimisation) on the libspo kernels.

CLI-compatible with the reference's safepo/single_agent/rcpo.py -- which is natural_pg.py with the Lagrange
multiplier of trpo_lag.py (rcpo.py:214-219,359,368-369): full-length natural-gradient step on the
Lagrange-mixed advantage, Train/LagragianMultiplier logged, no Misc/AcceptanceStep column.  Surrogate gradient, Fisher-vector products, conjugate
gradient and critic regression are the kernels of trpo_lag (safepo/single_agent/_engine.py).
"""
from __future__ import annotations

from safepo.single_agent._engine import run_trust_region

CONJUGATE_GRADIENT_ITERS = 15

default_cfg = {
    "hidden_sizes": [64, 64],
    "gamma": 0.99,
    "target_kl": 0.01,
    "batch_size": 128,
    "learning_iters": 10,
    "max_grad_norm": 40.0,
}


def main(args, cfg_env=None, env=None, max_epochs=None, quiet=False):
    return run_trust_region(args, dict(default_cfg), "rcpo", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
