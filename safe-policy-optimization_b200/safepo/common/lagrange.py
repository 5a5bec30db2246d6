"""Lagrange multiplier (host-side scalar).

Behaviour of safepo/common/lagrange.py:24-105 of the reference: lambda is moved by one
Adam step (lr 0.035, betas (0.9, 0.999), eps 1e-8) on the loss  -lambda * (Jc - d)  and
then clamped to [0, upper_bound].  It is a single float per epoch, so it stays on the
host; its value enters the advantage-mix kernel (spo_adv_apply) and ``Jc`` is the second
quantity the multi-GPU path all-reduces.

The arithmetic is float32, written out explicitly in numpy scalars in the order torch's
Adam applies it, so the sequence of multipliers is bit-identical to the reference's
(checked against tests/golden/lagrange.pt).
"""
from __future__ import annotations

import math

import numpy as np

_F = np.float32


class Lagrange:
    def __init__(self, cost_limit, lagrangian_multiplier_init, lagrangian_multiplier_lr, lagrangian_upper_bound=None):
        self.cost_limit = cost_limit
        self.lagrangian_multiplier_lr = lagrangian_multiplier_lr
        self.lagrangian_upper_bound = lagrangian_upper_bound
        self._lagrangian_multiplier = _F(max(lagrangian_multiplier_init, 0.0))
        self._m, self._v, self._t = _F(0.0), _F(0.0), 0
        self._beta1, self._beta2, self._eps = 0.9, 0.999, 1e-8

    @property
    def lagrangian_multiplier(self):
        """Projected multiplier as a python float (reference: relu(param).item())."""
        lam = float(self._lagrangian_multiplier)
        return lam if lam > 0.0 or lam != lam else 0.0

    def compute_lambda_loss(self, mean_ep_cost):
        return float(-self._lagrangian_multiplier * _F(mean_ep_cost - self.cost_limit))

    def update_lagrange_multiplier(self, Jc):
        """One Adam step on -lambda*(Jc - cost_limit), then clamp (lagrange.py:89-105)."""
        with np.errstate(all="ignore"):
            g = -_F(Jc - self.cost_limit)                      # d loss / d lambda, rounded to fp32
            self._t += 1
            w1, w2 = _F(1.0 - self._beta1), _F(1.0 - self._beta2)
            # lerp(m, g, 1-beta1) is a fused multiply-add in torch's CPU kernel: the fp32 product is
            # exact in fp64, so one rounding of the fp64 sum reproduces it
            self._m = _F(np.float64(self._m) + np.float64(w1) * np.float64(_F(g - self._m)))
            self._v = _F(_F(self._v * _F(self._beta2)) + _F(_F(w2 * g) * g))
            bc1 = 1.0 - self._beta1 ** self._t
            bc2_sqrt = math.sqrt(1.0 - self._beta2 ** self._t)
            step_size = _F(-(self.lagrangian_multiplier_lr / bc1))
            denom = _F(_F(np.sqrt(self._v) / _F(bc2_sqrt)) + _F(self._eps))
            lam = _F(self._lagrangian_multiplier + _F(_F(step_size * self._m) / denom))
            if lam == lam:  # NaN propagates (reference quirk A3: an epoch without episodes poisons lambda)
                lam = max(lam, _F(0.0))
                if self.lagrangian_upper_bound is not None:
                    lam = min(lam, _F(self.lagrangian_upper_bound))
            self._lagrangian_multiplier = _F(lam)
