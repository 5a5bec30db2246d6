"""PCPO (projection-based constrained policy optimisation) on the libspo kernels.

CLI-compatible with the reference's safepo/single_agent/pcpo.py -- which is cpo.py with the dual / case
analysis replaced by the projection step (pcpo.py:371,392-401)

    step = sqrt(2 delta / (q + 1e-8)) * F x  -  max(0, (sqrt(2 delta / q) r + c) / s) * p

(optim_case 0) and a line search of up to 200 steps (pcpo.py:44).  Surrogate gradients, both conjugate-gradient
solves, the Fisher-vector product, the line-search evaluations and the critic regression are the kernels of cpo
(safepo/single_agent/_engine.py)."""
from __future__ import annotations

from safepo.single_agent._engine import run_trust_region

STEP_FRACTION = 0.8
PCPO_SEARCHING_STEPS = 200
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
    return run_trust_region(args, dict(default_cfg), "pcpo", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
