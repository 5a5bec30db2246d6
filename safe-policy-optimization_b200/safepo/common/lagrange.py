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


class PIDLagrangian:
    """PID-controlled multiplier of CPPO-PID (Stooke, Achiam, Abbeel 2020), interface and defaults of the
    reference's ``PIDLagrangian`` (safepo/common/lagrange.py:108-200, used by cppo_pid.py:39,128-131).

    State: integral term I (clamped at 0), EMA of the constraint violation (proportional term), EMA of the
    episode cost with a ``pid_d_delay``-step delay line (derivative term, only rises count).  The penalty is
    max(0, Kp * ema(delta) + I + Kd * max(0, ema(cost) - ema(cost)[t - delay])).  Python floats on the host,
    evaluated in the reference's order, so the sequence is bit-identical (tests/golden/siblings.pt)."""

    def __init__(self, cost_limit, lagrangian_multiplier_init=0.005, pid_kp=0.1, pid_ki=0.01, pid_kd=0.01, pid_d_delay=10,
                 pid_delta_p_ema_alpha=0.95, pid_delta_d_ema_alpha=0.95, sum_norm=True, diff_norm=False, penalty_max=100.0):
        from collections import deque
        self.cost_limit = cost_limit
        self._gains = (pid_kp, pid_ki, pid_kd)
        self._ema = (pid_delta_p_ema_alpha, pid_delta_d_ema_alpha)
        self._sum_norm, self._diff_norm, self._penalty_max = sum_norm, diff_norm, penalty_max
        self._integral = lagrangian_multiplier_init
        self._violation_ema = 0.0
        self._cost_ema = 0.0
        self._delay_line = deque([0.0], maxlen=pid_d_delay)
        self._penalty = 0.0

    @property
    def lagrangian_multiplier(self):
        return self._penalty

    def update_lagrange_multiplier(self, ep_cost_avg):
        kp, ki, kd = self._gains
        alpha_p, alpha_d = self._ema
        violation = float(ep_cost_avg - self.cost_limit)
        self._integral = max(0.0, self._integral + violation * ki)
        if self._diff_norm:
            self._integral = max(0.0, min(1.0, self._integral))
        self._violation_ema *= alpha_p
        self._violation_ema += (1 - alpha_p) * violation
        self._cost_ema *= alpha_d
        self._cost_ema += (1 - alpha_d) * float(ep_cost_avg)
        rise = max(0.0, self._cost_ema - self._delay_line[0])
        self._penalty = max(0.0, kp * self._violation_ema + self._integral + kd * rise)
        if self._diff_norm:
            self._penalty = min(1.0, self._penalty)
        if not (self._diff_norm or self._sum_norm):
            self._penalty = min(self._penalty, self._penalty_max)
        self._delay_line.append(self._cost_ema)
