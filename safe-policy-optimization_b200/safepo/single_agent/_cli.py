"""``__main__`` plumbing shared by the algorithm scripts: run-directory naming and the
optional stdout/stderr redirection (reference ppo_lag.py:390-426)."""
from __future__ import annotations

import os
import sys
import time


def launch(main, algo_file, argv=None):
    from safepo.utils.config import single_agent_args

    args, cfg_env = single_agent_args(argv)
    stamp = time.strftime("%Y-%m-%d-%H-%M-%S")
    run = "-".join(["seed", str(args.seed).zfill(3), stamp])
    algo = os.path.basename(algo_file).split(".")[0]
    args.log_dir = os.path.join(args.log_dir, args.experiment, args.task, algo, run)
    if args.write_terminal:
        return main(args, cfg_env)
    os.makedirs(args.log_dir, exist_ok=True)
    out_path = os.path.join(args.log_dir, f"seed{args.seed}_terminal.log")
    err_path = os.path.join(args.log_dir, f"seed{args.seed}_error.log")
    with open(out_path, "w", encoding="utf-8") as f_out, open(err_path, "w", encoding="utf-8") as f_err:
        old = sys.stdout, sys.stderr
        sys.stdout, sys.stderr = f_out, f_err
        try:
            return main(args, cfg_env)
        finally:
            sys.stdout, sys.stderr = old
