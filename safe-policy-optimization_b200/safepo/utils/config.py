"""Command-line flags of the single-agent trainers.

Same flag names and defaults as single_agent_args() of the reference
(safepo/utils/config.py:144-162), plus the flags this implementation adds:
``--rng`` (host: noise and minibatch order come from torch's CPU generator exactly as the
reference consumes it -- parity mode; device: in-kernel Philox and device-side shuffles),
``--gae`` (scan | exact), ``--env`` (synthetic | mujoco), ``--episode-len``.
"""
from __future__ import annotations

import argparse


def _bool(x):
    return str(x).lower() in ("1", "true", "yes", "y", "t", "on")


def build_parser():
    p = argparse.ArgumentParser(description="RL Policy")
    p.add_argument("--seed", type=int, default=0, help="Random seed")
    p.add_argument("--use-eval", type=_bool, default=False, help="Use evaluation environment for testing")
    p.add_argument("--task", type=str, default="SafetyPointGoal1-v0", help="The task to run")
    p.add_argument("--num-envs", type=int, default=10, help="The number of parallel game environments")
    p.add_argument("--experiment", type=str, default="single_agent_exp", help="Experiment name")
    p.add_argument("--log-dir", type=str, default="../runs", help="directory to save agent logs")
    p.add_argument("--device", type=str, default="cuda", help="The device to run the model on (cuda only)")
    p.add_argument("--device-id", type=int, default=0, help="The device id to run the model on")
    p.add_argument("--write-terminal", type=_bool, default=True, help="Toggles terminal logging")
    p.add_argument("--headless", type=_bool, default=False, help="Toggles headless mode")
    p.add_argument("--total-steps", type=int, default=10000000, help="Total timesteps of the experiments")
    p.add_argument("--steps-per-epoch", type=int, default=20000, help="Steps per policy rollout (all envs)")
    p.add_argument("--randomize", type=bool, default=False, help="Randomize the environments' initial states")
    p.add_argument("--cost-limit", type=float, default=25.0, help="cost_lim")
    p.add_argument("--lagrangian-multiplier-init", type=float, default=0.001, help="initial lagrangian multiplier")
    p.add_argument("--lagrangian-multiplier-lr", type=float, default=0.035, help="lagrangian multiplier lr")
    # additions
    p.add_argument("--rng", choices=("host", "device"), default="device", help="host = reference-parity RNG streams")
    p.add_argument("--gae", choices=("scan", "exact"), default="scan", help="GAE kernel variant")
    p.add_argument("--env", choices=("synthetic", "mujoco"), default="synthetic", help="environment backend")
    p.add_argument("--episode-len", type=int, default=1000, help="time limit of the synthetic env")
    p.add_argument("--normalize-obs", action="store_true",
                   help="SafeNormalizeObservation on the device (the reference's env factory applies it on the host for MuJoCo tasks)")
    return p


def single_agent_args(argv=None):
    args = build_parser().parse_args(argv)
    return args, {}
