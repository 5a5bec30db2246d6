"""Epoch logger: progress.csv (+ TensorBoard when available), config.json, checkpoints.

Keeps the observable behaviour of the reference's safepo/common/logger.py that the
training loop depends on (it is out of the accelerated path, but its ``get_stats``
semantics feed the Lagrange update):

* ``store(**kv)`` appends to per-key lists; ``log_tabular(key)`` logs the mean and clears;
  ``log_tabular(key, val)`` logs ``val``; the first dumped row fixes the CSV header.
* ``get_stats(key)`` is 0.0 until ``key`` has been logged once, afterwards the mean of
  what was stored since the last ``log_tabular`` -- NaN when nothing was (logger.py:369-373).
* artefacts: ``<log_dir>/progress.csv``, ``config.json``, ``torch_save/model{itr}.pt`` (the
  actor ``state_dict``), ``state{itr}.pkl`` (joblib), ``tb/`` (logger.py:194-271).
"""
from __future__ import annotations

import atexit
import csv
import json
import os
import warnings

import numpy as np
import torch

_COLORS = {"green": 32, "yellow": 33, "red": 31, "blue": 34, "magenta": 35, "cyan": 36}


def _jsonable(obj):
    if isinstance(obj, dict):
        return {str(k): _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, (str, int, float, bool)) or obj is None:
        return obj
    if isinstance(obj, (np.integer, np.floating)):
        return obj.item()
    return str(obj)


class EpochLogger:
    def __init__(self, log_dir, seed=None, output_fname="progress.csv", debug=False, level=1,
                 use_tensorboard=True, verbose=True):
        self.log_dir, self.level, self.verbose = log_dir, level, verbose
        os.makedirs(log_dir, exist_ok=True)
        self._file = open(os.path.join(log_dir, output_fname), "w", encoding="utf-8", newline="")
        atexit.register(self._file.close)
        self._csv = csv.writer(self._file)
        parts = log_dir.replace("\\", "/").rstrip("/").split("/")
        self.exp_name = "-".join([parts[-3] if len(parts) >= 3 else "exp", parts[-2] if len(parts) >= 2 else "task",
                                  "seed", str(seed)])
        self.epoch = 0
        self.first_row = True
        self.log_headers, self.log_current_row, self.epoch_dict = [], {}, {}
        self.logged = True                 # the trainers flip this when an episode ends (ppo_lag.py:230)
        self.torch_saver_elements = None
        self.summary_writer = None
        if use_tensorboard:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.summary_writer = SummaryWriter(os.path.join(log_dir, "tb"))
            except Exception:  # tensorboard is optional
                self.summary_writer = None

    # ---- text -----------------------------------------------------------------------------
    def log(self, msg, color="green"):
        if self.verbose and self.level > 0:
            print(f"\x1b[{_COLORS.get(color, 32)}m{msg}\x1b[0m")

    # ---- diagnostics ------------------------------------------------------------------------
    def store(self, **kwargs):
        for k, v in kwargs.items():
            self.epoch_dict.setdefault(k, []).append(v)

    def get_stats(self, key):
        if key not in self.log_headers:
            return 0.0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return np.mean(self.epoch_dict[key])

    def log_tabular(self, key, val=None, min_and_max=False, std=False):
        if val is None:
            vals = self.epoch_dict[key]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self._put(key, np.mean(vals))
                if min_and_max:
                    self._put(key + "/Min", np.min(vals))
                    self._put(key + "/Max", np.max(vals))
                if std:
                    self._put(key + "/Std", np.std(vals))
        else:
            self._put(key, val)
        self.epoch_dict[key] = []

    def _put(self, key, val):
        if self.first_row:
            self.log_headers.append(key)
        elif key not in self.log_headers:
            raise AssertionError(f"new key {key} after the first row")
        if key in self.log_current_row:
            raise AssertionError(f"{key} already set this iteration")
        self.log_current_row[key] = val

    def dump_tabular(self):
        self.logged = True
        self.epoch += 1
        width = max(15, max(len(k) for k in self.log_headers))
        show = self.verbose and self.level > 0
        if show:
            print("-" * (width + 22))
        row = []
        for key in self.log_headers:
            val = self.log_current_row.get(key, "")
            row.append(val)
            if show:
                text = "%8.3g" % val if hasattr(val, "__float__") else val
                print(f"| {key:>{width}s} | {text:>15s} |")
        if show:
            print("-" * (width + 22), flush=True)
        if self.first_row:
            self._csv.writerow(self.log_headers)
        self._csv.writerow(row)
        self._file.flush()
        if self.summary_writer is not None:
            for k, v in self.log_current_row.items():
                self.summary_writer.add_scalar(k, float(v), global_step=self.epoch)
        self.log_current_row.clear()
        self.first_row = False

    # ---- artefacts ------------------------------------------------------------------------
    def save_config(self, config):
        cfg = _jsonable(config)
        cfg["exp_name"] = self.exp_name
        with open(os.path.join(self.log_dir, "config.json"), "w") as f:
            f.write(json.dumps(cfg, separators=(",", ":\t"), indent=4, sort_keys=True))

    def setup_torch_saver(self, what_to_save):
        self.torch_saver_elements = what_to_save

    def torch_save(self, itr=None):
        assert self.torch_saver_elements is not None, "call setup_torch_saver first"
        d = os.path.join(self.log_dir, "torch_save")
        os.makedirs(d, exist_ok=True)
        name = "model" + (f"{itr}" if itr is not None else "") + ".pt"
        sd = {k: v.detach().cpu().clone() for k, v in self.torch_saver_elements.state_dict().items()}
        torch.save(sd, os.path.join(d, name))

    def save_state(self, state_dict, itr=None):
        name = "state.pkl" if itr is None else f"state{itr}.pkl"
        try:
            import joblib
            joblib.dump(state_dict, os.path.join(self.log_dir, name))
        except Exception:
            self.log("Warning: could not pickle state_dict.", color="red")
        if self.torch_saver_elements is not None:
            self.torch_save(itr)

    def close(self):
        self._file.close()
        if self.summary_writer is not None:
            self.summary_writer.close()


Logger = EpochLogger
