"""CPO on the libspo kernels (reference: safepo/single_agent/cpo.py).

Same CLI / ``default_cfg`` / log columns.  Per epoch: two surrogate gradients (reward g,
cost b) from ``spo_surrogate_grad``, two 15-iteration conjugate-gradient solves entirely on
the device (``spo_conjugate_gradient`` on the closed-form Fisher-vector product
``spo_fvp``), the analytic dual / case analysis on a dozen host scalars (cpo.py:389-463),
a constrained backtracking line search on ``spo_linesearch_eval``, then the critic
regression as persistent ``spo_pg_update`` launches (cpo.py:534-571)."""
from __future__ import annotations

from safepo.single_agent._engine import run_trust_region

STEP_FRACTION = 0.8
CPO_SEARCHING_STEPS = 15
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
    return run_trust_region(args, dict(default_cfg), "cpo", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
