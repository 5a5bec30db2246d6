"""CUP (constrained update projection) on the libspo kernels.

CLI-compatible with the reference's safepo/single_agent/cup.py -- which is ppo_lag.py with (1) a plain PPO first
stage on the reward advantage (cup.py:284), (2) a second, actor-only projection stage (cup.py:355-404): minibatch
steps on (lambda * coef * ratio * adv_c + KL(new || old)).mean(), coef = (1 - gamma * 0.95) / (1 - gamma), with its
own KL early stop (Train/SeconStageStopIter), and (3) the multiplier capped at CUP_NU = 0.2.  Both stages run in the
persistent update kernel (second stage: loss kind SPO_LOSS_CUP_PROJECTION, critic CTAs idle).
"""
from __future__ import annotations

from safepo.single_agent._engine import run_policy_gradient

CUP_LAMBDA = 0.95
CUP_NU = 0.20

default_cfg = {
    "hidden_sizes": [64, 64],
    "gamma": 0.99,
    "target_kl": 0.02,
    "batch_size": 64,
    "learning_iters": 40,
    "max_grad_norm": 40.0,
}


def main(args, cfg_env=None, env=None, max_epochs=None, quiet=False):
    return run_policy_gradient(args, dict(default_cfg), "cup", env=env, max_epochs=max_epochs, quiet=quiet)


if __name__ == "__main__":
    from safepo.single_agent._cli import launch

    launch(main, __file__)
