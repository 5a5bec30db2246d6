"""TRPO-Lagrangian on the libspo kernels (reference: safepo/single_agent/trpo_lag.py).

Same CLI / ``default_cfg`` / log columns.  One natural-gradient step per epoch on the
Lagrange-mixed advantage: ``spo_surrogate_grad`` -> on-device conjugate gradient over the
closed-form Fisher-vector product -> backtracking line search (``spo_linesearch_eval``),
then the critic regression (trpo_lag.py:457-494) as persistent ``spo_pg_update`` launches."""
from __future__ import annotations

from safepo.single_agent._engine import run_trust_region

TRPO_SEARCHING_STEPS = 15
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
    return run_trust_region(args, dict(default_cfg), "trpo_lag", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
