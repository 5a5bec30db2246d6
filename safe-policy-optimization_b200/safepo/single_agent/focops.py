"""FOCOPS on the libspo kernels (reference: safepo/single_agent/focops.py).

Same CLI / ``default_cfg`` / log columns as the reference; the loss is the reference's
``[B,1] x [B]`` broadcast form (focops.py:331-337), evaluated in closed form inside
``spo_pg_update`` (mean(kl*m) - mean(ratio*adv)/lam * mean(m)); lambda is capped at
FOCOPS_NU (focops.py:132-137)."""
from __future__ import annotations

from safepo.single_agent._engine import run_policy_gradient

FOCOPS_LAM = 1.5
FOCOPS_NU = 2.0

default_cfg = {
    "hidden_sizes": [64, 64],
    "gamma": 0.99,
    "target_kl": 0.02,
    "batch_size": 64,
    "learning_iters": 40,
    "max_grad_norm": 40.0,
}


def main(args, cfg_env=None, env=None, max_epochs=None, quiet=False):
    return run_policy_gradient(args, dict(default_cfg), "focops", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
