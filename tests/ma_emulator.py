"""TEST INFRASTRUCTURE ONLY: a torch-CPU emulation of the spo_ma_* entry points of libspo (include/spo.h) and of spo_gae_masked,
so that the HOST orchestration of the multi-agent path (safepo/common/ma_model.py: MultiAgentNets / MultiAgentTrainer,
safepo/common/buffer.py: SeparatedReplayBuffer, safepo/multi_agent/mappolag.py: Runner) can be exercised in the `-m "not gpu"`
suite against the oracle: which buffer goes into which argument, the partial-sum layouts it relies on, the order of the calls.
It is not a fallback: nothing in the package imports it, and the kernels themselves are tested on the GPU
(tests/test_gpu_parity.py).  Every function takes the same positional arguments as the C function; "pointers" are the tensors
themselves (tests patch ``safepo._lib.ptr`` to the identity), treated as flat memory like the kernels do."""
import math

import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def _mem(t, *shape):
    """The first prod(shape) floats behind a 'pointer', viewed with that shape (writes go through)."""
    n = 1
    for s in shape:
        n *= s
    return t.reshape(-1)[:n].view(*shape)


def _ln(x, w, b):
    m = x.mean(-1, keepdim=True)
    v = ((x - m) ** 2).mean(-1, keepdim=True)
    return (x - m) * torch.rsqrt(v + 1e-5) * w + b


class EmulatedLib:
    def spo_last_error(self):
        return b""

    # ---- forward ----
    def _layer(self, x, n, K, W, b, lnw, lnb, H, linw, linb, out, pre, xn):
        x = _mem(x, n, K)
        if linw is not None:
            x = _ln(x, _mem(linw, K), _mem(linb, K))
            if xn is not None:
                _mem(xn, n, K).copy_(x)
        e = F.elu(x @ _mem(W, H, K).t() + _mem(b, H))
        if pre is not None:
            _mem(pre, n, H).copy_(e)
        _mem(out, n, H).copy_(_ln(e, _mem(lnw, H), _mem(lnb, H)))
        return 0

    def spo_ma_mlp_layer(self, x, n, K, W, b, lnw, lnb, H, linw, linb, out, stream):
        return self._layer(x, n, K, W, b, lnw, lnb, H, linw, linb, out, None, None)

    def spo_ma_mlp_layer_train(self, x, n, K, W, b, lnw, lnb, H, linw, linb, out, pre, xn, stream):
        return self._layer(x, n, K, W, b, lnw, lnb, H, linw, linb, out, pre, xn)

    def spo_ma_head(self, feat, n, H, W, b, O, log_std, x_coef, y_coef, eps, out, logp, stream):
        mean = _mem(feat, n, H) @ _mem(W, O, H).t() + _mem(b, O)
        if log_std is None:
            _mem(out, n, O).copy_(mean)
            return 0
        std = torch.sigmoid(_mem(log_std, O) / x_coef) * y_coef
        act = mean if eps is None else mean + _mem(eps, n, O) * std
        _mem(out, n, O).copy_(act)
        if logp is not None:
            _mem(logp, n, O).copy_(-((act - mean) ** 2) / (2 * std ** 2) - std.log() - LOG_SQRT_2PI)
        return 0

    # ---- backward pieces ----
    def spo_ma_ln_elu_bwd(self, dy, pre, ln_w, n, H, dz, part, stream):
        dy, e, gam = _mem(dy, n, H), _mem(pre, n, H), _mem(ln_w, H)
        m = e.mean(-1, keepdim=True)
        rstd = torch.rsqrt(((e - m) ** 2).mean(-1, keepdim=True) + 1e-5)
        xh = (e - m) * rstd
        dx = dy * gam
        dpre = rstd * (dx - dx.mean(-1, keepdim=True) - xh * (dx * xh).mean(-1, keepdim=True))
        z = dpre * torch.where(e > 0, torch.ones_like(e), e + 1)
        _mem(dz, n, H).copy_(z)
        nb = (n + 31) // 32
        p = _mem(part, nb, 3, H)
        for blk in range(nb):
            r = slice(32 * blk, min(n, 32 * blk + 32))
            p[blk, 0], p[blk, 1], p[blk, 2] = (dy[r] * xh[r]).sum(0), dy[r].sum(0), z[r].sum(0)
        return 0

    def spo_ma_ln_in_bwd(self, dxn, x, n, K, part, stream):
        d, x = _mem(dxn, n, K), _mem(x, n, K)
        m = x.mean(-1, keepdim=True)
        xh = (x - m) * torch.rsqrt(((x - m) ** 2).mean(-1, keepdim=True) + 1e-5)
        nb = (n + 31) // 32
        p = _mem(part, nb, 2, K)
        for blk in range(nb):
            r = slice(32 * blk, min(n, 32 * blk + 32))
            p[blk, 0], p[blk, 1] = (d[r] * xh[r]).sum(0), d[r].sum(0)
        return 0

    def spo_ma_partial_reduce(self, part, nblk, stride, nseg, length, out0, out1, out2, scale, stream):
        p = _mem(part, nblk, stride)
        for s, out in enumerate((out0, out1, out2)[:nseg]):
            if out is not None:
                _mem(out, length).copy_(p[:, s * length:(s + 1) * length].sum(0) * scale)
        return 0

    def spo_ma_gemm_nn(self, A, B, C, M, N, Kd, stream):
        _mem(C, M, N).copy_(_mem(A, M, Kd) @ _mem(B, Kd, N))
        return 0

    def spo_ma_gemm_tn(self, A, B, part, R, M, N, slices, stream):
        A, B = _mem(A, R, M), _mem(B, R, N)
        rps = ((R + slices - 1) // slices + 31) // 32 * 32
        assert (slices - 1) * rps < R, "empty slice (the C function rejects this)"
        p = _mem(part, slices, M, N)
        for z in range(slices):
            r = slice(z * rps, min(R, (z + 1) * rps))
            p[z] = A[r].t() @ B[r]
        return 0

    # ---- loss heads ----
    def spo_ma_actor_loss(self, feat, n, H, W, b, log_std, A, actions, old_logp, adv, cost_adv, factor, lamda, clip_lo, clip_hi, x_coef, y_coef,
                          dmean, imp, part, stream):
        mu = _mem(feat, n, H) @ _mem(W, A, H).t() + _mem(b, A)
        std = torch.sigmoid(_mem(log_std, A) / x_coef) * y_coef
        diff = _mem(actions, n, A) - mu
        lp = -(diff ** 2) / (2 * std ** 2) - std.log() - LOG_SQRT_2PI
        w = torch.exp(lp - _mem(old_logp, n, A)).prod(-1)
        a_h = _mem(adv, n) - _mem(lamda, 1) * _mem(cost_adv, n)
        fac = _mem(factor, n)
        s1, s2 = w * a_h, w.clamp(clip_lo, clip_hi) * a_h
        crow = torch.where(s1 <= s2, -fac * a_h / n, torch.zeros_like(s1)) * w
        dm = crow[:, None] * diff / std ** 2
        ds = crow[:, None] * (diff ** 2 / std ** 3 - 1 / std)
        _mem(dmean, n, A).copy_(dm)
        _mem(imp, n).copy_(w)
        loss_row = -fac * torch.min(s1, s2)
        nb = (n + 31) // 32
        p = _mem(part, nb, 66)
        p.zero_()
        for blk in range(nb):
            r = slice(32 * blk, min(n, 32 * blk + 32))
            p[blk, 0] = loss_row[r].sum()
            p[blk, 2:2 + A] = dm[r].sum(0)
            p[blk, 34:34 + A] = ds[r].sum(0)
        return 0

    def spo_ma_actor_finalize(self, part, nblk, n, log_std, A, x_coef, y_coef, entropy_coef, g_b, g_log_std, scalars, stream):
        tot = _mem(part, nblk, 66).sum(0)
        sg = torch.sigmoid(_mem(log_std, A) / x_coef)
        std = sg * y_coef
        _mem(g_log_std, A).copy_((tot[34:34 + A] - entropy_coef / (A * std)) * (y_coef * sg * (1 - sg) / x_coef))
        _mem(g_b, A).copy_(tot[2:2 + A])
        sc = _mem(scalars, 2)
        sc[0] = tot[0] / n
        sc[1] = (0.5 + LOG_SQRT_2PI + std.log()).mean()
        return 0

    def spo_ma_value_loss(self, v, vp, rn_c, rn_o, n, clip, delta, scale, dv, part, stream):
        v, vp, rn_c, rn_o = _mem(v, n), _mem(vp, n), _mem(rn_c, n), _mem(rn_o, n)
        dlt = v - vp
        ec, eo = rn_c - (vp + dlt.clamp(-clip, clip)), rn_o - v

        def hub(e):
            return (e.abs() <= delta).float() * e ** 2 / 2 + (e > delta).float() * delta * (e.abs() - delta / 2)

        def hgrad(e):
            return torch.where(e.abs() <= delta, e, torch.where(e > delta, torch.full_like(e, delta), torch.zeros_like(e)))
        hc, ho = hub(ec), hub(eo)
        wo = torch.where(ho > hc, torch.ones_like(ho), torch.where(ho == hc, torch.full_like(ho, 0.5), torch.zeros_like(ho)))
        inr = ((dlt >= -clip) & (dlt <= clip)).float()
        g = scale * (wo * (-hgrad(eo)) + (1 - wo) * inr * (-hgrad(ec)))
        _mem(dv, n).copy_(g)
        L = torch.max(ho, hc)
        nb = (n + 255) // 256
        p = _mem(part, nb, 2)
        for blk in range(nb):
            r = slice(256 * blk, min(n, 256 * blk + 256))
            p[blk, 0], p[blk, 1] = L[r].sum(), g[r].sum()
        return 0

    def spo_ma_popart_normalize(self, x, n, state, beta, epsilon, out, stream):
        x, st = _mem(x, n), _mem(state, 3)
        w, omw = torch.tensor(beta, dtype=torch.float32), torch.tensor(1.0 - beta, dtype=torch.float32)
        st[0] = st[0] * w + x.mean() * omw
        st[1] = st[1] * w + (x ** 2).mean() * omw
        st[2] = st[2] * w + omw
        den = st[2].clamp(min=epsilon)
        mean, msq = st[0] / den, st[1] / den
        _mem(out, n).copy_((x - mean) / torch.sqrt((msq - mean ** 2).clamp(min=1e-2)))
        return 0

    def spo_ma_lagrange_step(self, imp, cost_adv, aver, n, cost_limit, gamma, rate, lamda, stream):
        delta = -((_mem(aver, n).mean() - cost_limit) * (1 - gamma) + (_mem(imp, n) * _mem(cost_adv, n)).mean())
        lam = _mem(lamda, 1)
        lam.copy_(torch.relu(lam - delta * rate))
        return 0

    def spo_ma_clip_adam(self, params, grads, m, v, count, max_norm, lr, b1, b2, eps, wd, step, work, norm_out, stream):
        p, g, m, v = _mem(params, count), _mem(grads, count), _mem(m, count), _mem(v, count)
        norm = torch.sqrt((g ** 2).sum())
        coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        no = _mem(norm_out, 2)
        no[0], no[1] = norm, coef
        gg = g * coef
        if wd != 0.0:
            gg = gg + wd * p
        m.lerp_(gg, 1 - b1)
        v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
        return 0

    # ---- G2 ----
    def spo_gae_masked(self, rew, vp, masks, mean, sqrt_var, gamma, gamma_lambda, out, N, T, stream):
        rew, vp, masks, out = _mem(rew, T, N), _mem(vp, T + 1, N), _mem(masks, T + 1, N), _mem(out, T, N)
        den = vp * sqrt_var + mean
        gae = torch.zeros(N)
        for t in reversed(range(T)):
            delta = rew[t] + gamma * den[t + 1] * masks[t + 1] - den[t]
            gae = delta + gamma_lambda * masks[t + 1] * gae
            out[t] = gae + den[t]
        return 0


def install(monkeypatch):
    """Route the package's library calls to the emulation for one test (pytest's monkeypatch undoes it)."""
    from safepo import _lib as L
    from safepo.common.ma_model import MultiAgentNets
    lib = EmulatedLib()
    monkeypatch.setattr(L, "lib", lambda: lib)
    def ptr(t):
        if t is not None and not t.is_contiguous():
            raise L.SpoError("libspo needs contiguous tensors")       # the real ptr() refuses these too
        return t
    monkeypatch.setattr(L, "ptr", ptr)
    monkeypatch.setattr(L, "stream", lambda: None)
    monkeypatch.setattr(MultiAgentNets, "_require_cuda", staticmethod(lambda device: None))
    return lib
