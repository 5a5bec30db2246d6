"""Natural policy gradient on the libspo kernels.

CLI-compatible with the reference's safepo/single_agent/natural_pg.py -- which is trpo.py without the line
search: the step x * sqrt(2 * target_kl / xHx) is taken at full length (natural_pg.py:379), Train/KL is the KL
after the step, there is no Misc/AcceptanceStep column.  Surrogate gradient, Fisher-vector products, conjugate
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
    return run_trust_region(args, dict(default_cfg), "natural_pg", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
