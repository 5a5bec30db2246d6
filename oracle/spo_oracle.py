"""CPU oracle for the SafePO rollout -> dual-GAE -> update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``safe-policy-optimization_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker / the CPU
arm that is timed beside the CUDA path.

What it is: a restatement, in functional torch-CPU code, of the arithmetic the reference
executes on this path.  The reference is pure Python on top of PyTorch, so the
"third-party arithmetic" it relies on (GEMM, Normal, kl_divergence, Adam,
clip_grad_norm_) is torch itself (torch >= 1.10, un-pinned in the reference's
setup.py:23; 2.11.0 is what this image has) -- the oracle calls the same torch ops in
the same order, which is what makes it bit-identical to the reference on CPU.

Pinning: ``tests/golden/make_golden.py`` imports the real reference from
/root/reference (env packages stubbed) and writes fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against them bit-for-bit (forward,
GAE, buffer layout, Lagrange, PPO-Lag/FOCOPS minibatch chains, FVP, CG, and a full
``main()`` run of ppo_lag/cpo/trpo_lag/focops on a synthetic vector env).

Each function cites the reference file:line it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))

# --------------------------------------------------------------------------------------
# networks  (safepo/common/model.py:30-48, 73-81, 103-108, 131-135)
# --------------------------------------------------------------------------------------

NET_ORDER = ("reward_critic", "cost_critic", "actor")  # registration order, model.py:133-135


def _init_mlp(sizes):
    """Weights of build_mlp_network (model.py:30-48) with the same RNG consumption:
    nn.Linear's own init, then kaiming_uniform_(a=sqrt(5)) over the weight again."""
    out = []
    for j in range(len(sizes) - 1):
        lin = torch.nn.Linear(sizes[j], sizes[j + 1])
        torch.nn.init.kaiming_uniform_(lin.weight, a=np.sqrt(5))
        out.append((lin.weight.detach().clone(), lin.bias.detach().clone()))
    return out


class OraclePolicy:
    """Parameter container mirroring ActorVCritic (model.py:111-170).

    ``nets[name]`` is an ordered dict of leaf tensors named as the reference's
    ``state_dict`` names them (actor: log_std first, model.py:73-76)."""

    def __init__(self, obs_dim, act_dim, hidden_sizes=(64, 64)):
        hidden_sizes = list(hidden_sizes)
        self.obs_dim, self.act_dim, self.hidden_sizes = obs_dim, act_dim, hidden_sizes
        self.nets = {}
        for name in NET_ORDER:  # same construction order => same RNG stream
            out_dim = act_dim if name == "actor" else 1
            layers = _init_mlp([obs_dim] + hidden_sizes + [out_dim])
            prefix = "mean" if name == "actor" else "critic"
            d = {}
            if name == "actor":
                d["log_std"] = torch.zeros(act_dim)
            for li, (w, b) in enumerate(layers):
                d[f"{prefix}.{2 * li}.weight"] = w
                d[f"{prefix}.{2 * li}.bias"] = b
            for v in d.values():
                v.requires_grad_(True)
            self.nets[name] = d

    # parameter lists in the order torch would enumerate them
    def params(self, name):
        return list(self.nets[name].values())

    def all_params(self):
        return [p for n in NET_ORDER for p in self.params(n)]

    def clone(self):
        other = OraclePolicy.__new__(OraclePolicy)
        other.obs_dim, other.act_dim, other.hidden_sizes = self.obs_dim, self.act_dim, list(self.hidden_sizes)
        other.nets = {n: {k: v.detach().clone().requires_grad_(True) for k, v in d.items()}
                      for n, d in self.nets.items()}
        return other

    def load(self, state):
        """state: {net: {name: tensor}}"""
        with torch.no_grad():
            for n, d in state.items():
                for k, v in d.items():
                    self.nets[n][k].copy_(torch.as_tensor(v))

    def state(self):
        return {n: {k: v.detach().clone() for k, v in d.items()} for n, d in self.nets.items()}


def _mlp(d, prefix, obs):
    n_layers = sum(1 for k in d if k.endswith(".weight"))
    x = obs
    for li in range(n_layers):
        x = F.linear(x, d[f"{prefix}.{2 * li}.weight"], d[f"{prefix}.{2 * li}.bias"])
        if li < n_layers - 1:
            x = torch.tanh(x)
    return x


def actor_mean_std(pol, obs):
    """Actor.forward, model.py:78-81: Normal(mean(obs), exp(log_std))."""
    d = pol.nets["actor"]
    mean = _mlp(d, "mean", obs)
    std = torch.exp(d["log_std"])
    # Normal.__init__ broadcasts loc/scale once (broadcast_all); every consumer of the
    # distribution shares that one expanded tensor, which fixes the order in which
    # autograd accumulates d/d(std) -- keep a single expand node here too.
    mean, std = torch.broadcast_tensors(mean, std)
    return mean, std


def critic_value(pol, which, obs):
    """VCritic.forward, model.py:107-108."""
    return torch.squeeze(_mlp(pol.nets[which], "critic", obs), -1)


def normal_log_prob(x, mean, std):
    """torch.distributions.Normal.log_prob (same op order)."""
    std = std.expand_as(mean) if std.shape != mean.shape else std
    var = std ** 2
    return -((x - mean) ** 2) / (2 * var) - std.log() - LOG_SQRT_2PI


def normal_kl(p_mean, p_std, q_mean, q_std):
    """torch.distributions.kl._kl_normal_normal(p, q)."""
    p_std = p_std.expand_as(p_mean) if p_std.shape != p_mean.shape else p_std
    q_std = q_std.expand_as(q_mean) if q_std.shape != q_mean.shape else q_std
    var_ratio = (p_std / q_std).pow(2)
    t1 = ((p_mean - q_mean) / q_std).pow(2)
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def policy_step(pol, obs, eps=None, deterministic=False):
    """ActorVCritic.step, model.py:149-170.  ``eps`` (same shape as the action) replaces
    the draw rsample would take from the global generator (loc + eps*scale)."""
    mean, std = actor_mean_std(pol, obs)
    if deterministic:
        act = mean
    else:
        if eps is None:
            eps = torch.empty(mean.shape, dtype=mean.dtype).normal_()
        act = mean + eps * std
    logp = normal_log_prob(act, mean, std).sum(axis=-1)
    v_r = critic_value(pol, "reward_critic", obs)
    v_c = critic_value(pol, "cost_critic", obs)
    return act, logp, v_r, v_c


# --------------------------------------------------------------------------------------
# GAE  (safepo/common/buffer.py:97-140, 167-188, 191-201)
# --------------------------------------------------------------------------------------

def discount_cumsum(x, discount):
    """buffer.py:167-188: float64 reverse recurrence, python-double discount."""
    x = x.type(torch.float64).clone()
    acc = x[-1]
    for i in reversed(range(x.shape[0] - 1)):
        acc = x[i] + discount * acc
        x[i] = acc
    return x


def gae_path(values, rewards, lam, gamma):
    """calculate_adv_and_value_targets, buffer.py:191-201 (last element = bootstrap)."""
    deltas = rewards[:-1] + gamma * values[1:] - values[:-1]
    adv = discount_cumsum(deltas, gamma * lam)
    return adv, adv + values[:-1]


def gae_dual(rew, cost, v_r, v_c, seg_end, boot_r, boot_c, gamma=0.99, lam=0.95, lam_c=0.95):
    """Batched form of finish_path (buffer.py:97-140) over [N,T] arrays: segment
    boundaries are the slots where seg_end!=0, bootstraps are read at those slots.
    Returns fp32 adv_r, adv_c, tgt_r, tgt_c (the reference stores into fp32 buffers,
    buffer.py:135-138)."""
    N, T = rew.shape
    outs = [torch.zeros(N, T, dtype=torch.float32) for _ in range(4)]
    for n in range(N):
        start = 0
        for t in range(T):
            if seg_end[n, t]:
                sl = slice(start, t + 1)
                for (val, r, boot, l, oa, ot) in (
                    (v_r, rew, boot_r, lam, outs[0], outs[2]),
                    (v_c, cost, boot_c, lam_c, outs[1], outs[3]),
                ):
                    last = boot[n, t].reshape(1)
                    a, tg = gae_path(torch.cat([val[n, sl], last]), torch.cat([r[n, sl], last]), l, gamma)
                    oa[n, sl] = a  # f64 -> f32 rounding on store
                    ot[n, sl] = tg
                start = t + 1
    return outs


def adv_finalize(adv_r, adv_c, lam, standardize_r=True, standardize_c=True):
    """buffer.get (buffer.py:154-160) + the Lagrange mix (ppo_lag.py:280-281)."""
    mean, std, cmean = adv_r.mean(), adv_r.std(), adv_c.mean()
    a = (adv_r - mean) / (std + 1e-8) if standardize_r else adv_r
    c = adv_c - cmean if standardize_c else adv_c
    mixed = a - lam * c
    mixed = mixed / (lam + 1)
    return a, c, mixed


# --------------------------------------------------------------------------------------
# Lagrange multiplier  (safepo/common/lagrange.py:43-105)
# --------------------------------------------------------------------------------------

class OracleLagrange:
    def __init__(self, cost_limit, init, lr, upper_bound=None):
        self.cost_limit, self.upper_bound = cost_limit, upper_bound
        self._lam = torch.nn.Parameter(torch.as_tensor(max(init, 0.0)), requires_grad=True)
        self._opt = torch.optim.Adam([self._lam], lr=lr)

    @property
    def lagrangian_multiplier(self):
        return torch.relu(self._lam).detach().item()

    def update_lagrange_multiplier(self, Jc):
        self._opt.zero_grad()
        (-self._lam * (Jc - self.cost_limit)).backward()
        self._opt.step()
        self._lam.data.clamp_(0.0, self.upper_bound)


# --------------------------------------------------------------------------------------
# minibatch order  (torch DataLoader(shuffle=True) as used at ppo_lag.py:283-294)
# --------------------------------------------------------------------------------------

class OraclePIDLagrange:
    """PIDLagrangian (safepo/common/lagrange.py:108-200) with its default gains: plain Python-float
    arithmetic, restated in the reference's operation order (sum_norm=True, diff_norm=False)."""

    def __init__(self, cost_limit, init=0.005, kp=0.1, ki=0.01, kd=0.01, d_delay=10, p_ema=0.95, d_ema=0.95):
        from collections import deque
        self.kp, self.ki, self.kd, self.p_ema, self.d_ema = kp, ki, kd, p_ema, d_ema
        self.pid_i = init
        self.cost_ds = deque(maxlen=d_delay)
        self.cost_ds.append(0.0)
        self.delta_p = 0.0
        self.cost_d = 0.0
        self.cost_limit = cost_limit
        self.lagrangian_multiplier = 0.0

    def update_lagrange_multiplier(self, ep_cost_avg):
        delta = float(ep_cost_avg - self.cost_limit)
        self.pid_i = max(0.0, self.pid_i + delta * self.ki)
        self.delta_p *= self.p_ema
        self.delta_p += (1 - self.p_ema) * delta
        self.cost_d *= self.d_ema
        self.cost_d += (1 - self.d_ema) * float(ep_cost_avg)
        pid_d = max(0.0, self.cost_d - self.cost_ds[0])
        pid_o = self.kp * self.delta_p + self.pid_i + self.kd * pid_d
        self.lagrangian_multiplier = max(0.0, pid_o)
        self.cost_ds.append(self.cost_d)


def dataloader_perm(S):
    """One ``for ... in dataloader`` consumes two int64 draws from the global generator
    (base seed, then the sampler seed); the order is randperm(S) under the second."""
    torch.empty((), dtype=torch.int64).random_()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(S, generator=g)


# --------------------------------------------------------------------------------------
# PPO-Lag / FOCOPS / critic-only minibatch step (ppo_lag.py:298-336, focops.py:309-357,
# cpo.py:543-571)
# --------------------------------------------------------------------------------------

class OracleOptim:
    """Three Adam optimizers (ppo_lag.py:104-117); actor lr follows LinearLR(1 -> 0)."""

    def __init__(self, pol, lr=3e-4, critic_lr=None, epochs=1):
        self.actor = torch.optim.Adam(pol.params("actor"), lr=lr)
        clr = lr if critic_lr is None else critic_lr
        self.reward_critic = torch.optim.Adam(pol.params("reward_critic"), lr=clr)
        self.cost_critic = torch.optim.Adam(pol.params("cost_critic"), lr=clr)
        self.epochs, self.base_lr, self.epoch = epochs, lr, 0

    def scheduler_step(self):
        """LinearLR(start 1.0, end 0.0, total_iters=epochs).step() (ppo_lag.py:105-111,350).
        The scheduler's closed form: lr_e = base * (1 - e/epochs) for e <= epochs."""
        self.epoch += 1
        e = min(self.epoch, self.epochs)
        # chained form used by torch: lr *= 1 + (end-start)/(total*start + (e-1)*(end-start))
        for g in self.actor.param_groups:
            g["lr"] = g["lr"] * (1.0 + (0.0 - 1.0) / (self.epochs * 1.0 + (e - 1) * (0.0 - 1.0))) \
                if self.epoch <= self.epochs else g["lr"]

    def actor_lr(self):
        return self.actor.param_groups[0]["lr"]


def critic_losses(pol, obs_b, tgt_r_b, tgt_c_b, use_critic_norm=True):
    """ppo_lag.py:306-314."""
    loss_r = F.mse_loss(critic_value(pol, "reward_critic", obs_b), tgt_r_b)
    loss_c = F.mse_loss(critic_value(pol, "cost_critic", obs_b), tgt_c_b)
    if use_critic_norm:
        for p in pol.params("reward_critic"):
            loss_r = loss_r + p.pow(2).sum() * 0.001
        for p in pol.params("cost_critic"):
            loss_c = loss_c + p.pow(2).sum() * 0.001
    return loss_r, loss_c


def ppo_actor_loss(pol, obs_b, act_b, logp_b, adv_b):
    """ppo_lag.py:315-319 (clip range hard-coded 0.8/1.2)."""
    mean, std = actor_mean_std(pol, obs_b)
    logp = normal_log_prob(act_b, mean, std).sum(dim=-1)
    ratio = torch.exp(logp - logp_b)
    ratio_c = torch.clamp(ratio, 0.8, 1.2)
    return -torch.min(ratio * adv_b, ratio_c * adv_b).mean()


def pg_actor_loss(pol, obs_b, act_b, logp_b, adv_b):
    """pg.py:303-309: the importance-weighted surrogate without the clip."""
    mean, std = actor_mean_std(pol, obs_b)
    log_prob = normal_log_prob(act_b, mean, std).sum(-1)
    ratio = torch.exp(log_prob - logp_b)
    return -(ratio * adv_b).mean()


def focops_actor_loss(pol, obs_b, act_b, logp_b, adv_b, old_mean_b, old_std_b, target_kl, focops_lam=1.5):
    """focops.py:323-337, including the [B,1]x[B] broadcast."""
    mean, std = actor_mean_std(pol, obs_b)
    logp = normal_log_prob(act_b, mean, std).sum(dim=-1)
    ratio = torch.exp(logp - logp_b)
    temp_kl = normal_kl(mean, std, old_mean_b, old_std_b).sum(-1, keepdim=True)
    loss = (temp_kl - (1 / focops_lam) * ratio * adv_b) * (temp_kl.detach() <= target_kl).type(torch.float32)
    return loss.mean()


def minibatch_step(pol, opt, batch, kind="ppo", max_grad_norm=40.0, target_kl=0.02):
    """One optimisation step. kind in {"ppo","focops","critic"}.  Returns the three
    logged scalars (loss_r, loss_c, loss_pi)."""
    opt.reward_critic.zero_grad()
    opt.cost_critic.zero_grad()
    loss_r, loss_c = critic_losses(pol, batch["obs"], batch["target_value_r"], batch["target_value_c"])
    if kind == "ppo":
        loss_pi = ppo_actor_loss(pol, batch["obs"], batch["act"], batch["log_prob"], batch["adv"])
    elif kind == "pg":
        loss_pi = pg_actor_loss(pol, batch["obs"], batch["act"], batch["log_prob"], batch["adv"])
    elif kind == "focops":
        loss_pi = focops_actor_loss(pol, batch["obs"], batch["act"], batch["log_prob"], batch["adv"],
                                    batch["old_mean"], batch["old_std"], target_kl)
    else:
        loss_pi = None
    if loss_pi is not None:
        opt.actor.zero_grad()
        total = loss_pi + loss_r + loss_c
    else:
        total = loss_r + loss_c
    total.backward()
    torch.nn.utils.clip_grad_norm_(pol.all_params(), max_grad_norm)
    opt.reward_critic.step()
    opt.cost_critic.step()
    if loss_pi is not None:
        opt.actor.step()
    return loss_r.item(), loss_c.item(), (loss_pi.item() if loss_pi is not None else float("nan"))


def full_batch_kl(pol, obs, old_mean, old_std, reduce="sum_mean"):
    """ppo_lag.py:338-344 (.sum(-1).mean()) or cpo.py:489-491 (.mean() over S*A)."""
    with torch.no_grad():
        mean, std = actor_mean_std(pol, obs)
        kl = normal_kl(old_mean, old_std, mean, std)
        return (kl.sum(-1, keepdim=True).mean() if reduce == "sum_mean" else kl.mean()).item()


def pg_update(pol, opt, data, advantage, kind="ppo", batch_size=64, learning_iters=40,
              target_kl=0.02, max_grad_norm=40.0, perms=None, log=None):
    """The update loop of ppo_lag.py:277-348 / focops.py:280-366.  ``perms`` (list of
    index tensors) overrides the DataLoader-equivalent draws."""
    S = data["obs"].shape[0]
    with torch.no_grad():
        old_mean, old_std = actor_mean_std(pol, data["obs"])
        old_mean, old_std = old_mean.clone(), old_std.clone()
    old_std_full = old_std.expand_as(old_mean).clone()
    losses, update_counts, final_kl = [], 0, None
    for it in range(learning_iters):
        perm = perms[it] if perms is not None else dataloader_perm(S)
        for s in range(0, S, batch_size):
            idx = perm[s:s + batch_size]
            b = {"obs": data["obs"][idx], "act": data["act"][idx], "log_prob": data["log_prob"][idx],
                 "target_value_r": data["target_value_r"][idx], "target_value_c": data["target_value_c"][idx],
                 "adv": advantage[idx]}
            if kind == "focops":
                b["old_mean"], b["old_std"] = old_mean[idx], old_std_full[idx]
            losses.append(minibatch_step(pol, opt, b, kind, max_grad_norm, target_kl))
        final_kl = full_batch_kl(pol, data["obs"], old_mean, old_std)
        update_counts += 1
        if final_kl > target_kl:
            break
    return {"stop_iter": update_counts, "kl": final_kl, "losses": losses}


def cup_second_stage(pol, opt, data, lam, gamma=0.99, cup_lambda=0.95, batch_size=64, learning_iters=40, target_kl=0.02,
                     max_grad_norm=40.0, perms=None):
    """cup.py:355-404: the projection stage.  Actor only (its optimizer, its own grad-norm clip); the loss
    (lam * coef * ratio * adv_c [B] + KL(new || old) [B,1]).mean() broadcasts to [B,B] like FOCOPS."""
    S = data["obs"].shape[0]
    with torch.no_grad():
        old_mean, old_std = actor_mean_std(pol, data["obs"])
        old_mean, old_std = old_mean.clone(), old_std.clone()
    old_std_full = old_std.expand_as(old_mean).clone()
    advantage = data["adv_c"]
    coef = (1 - gamma * cup_lambda) / (1 - gamma)
    update_counts, final_kl = 0, None
    for it in range(learning_iters):
        perm = perms[it] if perms is not None else dataloader_perm(S)
        for s in range(0, S, batch_size):
            idx = perm[s:s + batch_size]
            mean, std = actor_mean_std(pol, data["obs"][idx])
            logp = normal_log_prob(data["act"][idx], mean, std).sum(dim=-1)
            ratio = torch.exp(logp - data["log_prob"][idx])
            temp_kl = normal_kl(mean, std, old_mean[idx], old_std_full[idx]).sum(-1, keepdim=True)
            loss = (lam * coef * ratio * advantage[idx] + temp_kl).mean()
            opt.actor.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(pol.params("actor"), max_grad_norm)
            opt.actor.step()
        final_kl = full_batch_kl(pol, data["obs"], old_mean, old_std)
        update_counts += 1
        if final_kl > target_kl:
            break
    return {"stop_iter": update_counts, "kl": final_kl}


def critic_regression(pol, opt, data, batch_size=128, learning_iters=10, max_grad_norm=40.0, perms=None):
    """cpo.py:534-571 / trpo_lag.py:457-494.  The joint clip runs over *all* parameters,
    so whatever .grad the actor still holds enters the norm (Appendix A6)."""
    S = data["obs"].shape[0]
    losses = []
    for it in range(learning_iters):
        perm = perms[it] if perms is not None else dataloader_perm(S)
        for s in range(0, S, batch_size):
            idx = perm[s:s + batch_size]
            b = {"obs": data["obs"][idx], "target_value_r": data["target_value_r"][idx],
                 "target_value_c": data["target_value_c"][idx]}
            losses.append(minibatch_step(pol, opt, b, "critic", max_grad_norm)[:2])
    return losses


# --------------------------------------------------------------------------------------
# trust-region pieces  (cpo.py:70-157, trpo_lag.py:70-159)
# --------------------------------------------------------------------------------------

def flat_params(pol):
    """get_flat_params_from(policy.actor), cpo.py:70-78 (log_std first)."""
    return torch.cat([p.data.view(-1) for p in pol.params("actor")])


def set_flat_params(pol, vals):
    """set_param_values_to_model, cpo.py:109-121."""
    i = 0
    for p in pol.params("actor"):
        n = p.numel()
        p.data = vals[i:i + n].view(p.size()).clone()
        i += n
    assert i == len(vals)


def flat_grads(pol):
    """get_flat_gradients_from, cpo.py:123-130."""
    return torch.cat([p.grad.view(-1) for p in pol.params("actor")])


def surrogate_loss(pol, obs, act, logp_old, adv):
    """mean(ratio*adv), cpo.py:356-359 / 373-376."""
    mean, std = actor_mean_std(pol, obs)
    logp = normal_log_prob(act, mean, std).sum(dim=-1)
    return (torch.exp(logp - logp_old) * adv).mean()


def fvp_autograd(pol, obs, v):
    """fvp(), cpo.py:132-157: Hessian-vector product of mean_{S*A} KL(old||new) + 0.1*v."""
    for p in pol.params("actor"):
        p.grad = None
    mean, std = actor_mean_std(pol, obs)
    with torch.no_grad():
        omean, ostd = actor_mean_std(pol, obs)
    kl = normal_kl(omean, ostd, mean, std).mean()
    grads = torch.autograd.grad(kl, tuple(pol.params("actor")), create_graph=True)
    flat = torch.cat([g.view(-1) for g in grads])
    kl_p = (flat * v).sum()
    grads = torch.autograd.grad(kl_p, tuple(pol.params("actor")), retain_graph=False)
    return torch.cat([g.contiguous().view(-1) for g in grads]) + v * 0.1


def fvp_analytic(pol, obs, v, damping=0.1):
    """Closed form of the same product: for a Gaussian with
    state-independent log_std,  F v = [2 v_ls / A ; J^T diag(sigma^-2) J v_mean / (S A)] + damping v.
    One JVP and one VJP through the mean MLP; this is the formulation the CUDA kernel uses."""
    d = pol.nets["actor"]
    A = d["log_std"].numel()
    S = obs.shape[0]
    names = [k for k in d if k != "log_std"]
    v_ls = v[:A]
    rest, off = {}, A
    for k in names:
        n = d[k].numel()
        rest[k] = v[off:off + n].view_as(d[k])
        off += n
    with torch.no_grad():
        W1, b1, W2, b2, W3, b3 = (d[k] for k in names)
        V1, c1, V2, c2, V3, c3 = (rest[k] for k in names)
        h1 = torch.tanh(F.linear(obs, W1, b1))
        h2 = torch.tanh(F.linear(h1, W2, b2))
        # JVP
        dz1 = F.linear(obs, V1, c1)
        dh1 = dz1 * (1 - h1 * h1)
        dz2 = F.linear(dh1, W2) + F.linear(h1, V2, c2)
        dh2 = dz2 * (1 - h2 * h2)
        dmu = F.linear(dh2, W3) + F.linear(h2, V3, c3)
        # metric + VJP
        g_mu = dmu * torch.exp(-2 * d["log_std"]) / (S * A)
        gW3, gb3 = g_mu.t() @ h2, g_mu.sum(0)
        gz2 = (g_mu @ W3) * (1 - h2 * h2)
        gW2, gb2 = gz2.t() @ h1, gz2.sum(0)
        gz1 = (gz2 @ W2) * (1 - h1 * h1)
        gW1, gb1 = gz1.t() @ obs, gz1.sum(0)
        out = torch.cat([2.0 * v_ls / A] + [g.reshape(-1) for g in (gW1, gb1, gW2, gb2, gW3, gb3)])
    return out + damping * v


def conjugate_gradients(Avp, b, num_steps=15, residual_tol=1e-10, eps=1e-6):
    """cpo.py:81-106 (including the FVP on zeros at :91)."""
    x = torch.zeros_like(b)
    r = b - Avp(x)
    p = r.clone()
    rdotr = torch.dot(r, r)
    for _ in range(num_steps):
        z = Avp(p)
        alpha = rdotr / (torch.dot(p, z) + eps)
        x += alpha * p
        r -= alpha * z
        new_rdotr = torch.dot(r, r)
        if torch.sqrt(new_rdotr) < residual_tol:
            break
        mu = new_rdotr / (rdotr + eps)
        p = r + mu * p
        rdotr = new_rdotr
    return x


def cpo_step_direction(x, p, xHx, g, b, ep_costs, target_kl):
    """Scalar case analysis of cpo.py:385-463.  All tensor arguments are CPU fp32.
    Returns (step_direction, optim_case, lambda_star, nu_star, alpha)."""
    q, r, s = xHx, g.dot(p), b.dot(p)
    alpha = torch.sqrt(2 * target_kl / (xHx + 1e-8))
    if b.dot(b) <= 1e-6 and ep_costs < 0:
        A, B, case = torch.zeros(1), torch.zeros(1), 4
    else:
        A = q - r ** 2 / (s + 1e-8)
        B = 2 * target_kl - ep_costs ** 2 / (s + 1e-8)
        if ep_costs < 0 and B < 0:
            case = 3
        elif ep_costs < 0 <= B:
            case = 2
        elif ep_costs >= 0 and B >= 0:
            case = 1
        else:
            case = 0
    if case in (3, 4):
        nu_star = torch.zeros(1)
        lambda_star = 1 / (alpha + 1e-8)
        step = alpha * x
    elif case in (1, 2):
        lambda_a = torch.sqrt(A / B)
        lambda_b = torch.sqrt(q / (2 * target_kl))
        r_num = r.item()
        eps_cost = ep_costs + 1e-8
        if ep_costs < 0:
            la = torch.clamp(lambda_a, torch.as_tensor(0.0), r_num / eps_cost)
            lb = torch.clamp(lambda_b, r_num / eps_cost, torch.as_tensor(torch.inf))
        else:
            la = torch.clamp(lambda_a, r_num / eps_cost, torch.as_tensor(torch.inf))
            lb = torch.clamp(lambda_b, torch.as_tensor(0.0), r_num / eps_cost)
        f_a = lambda lam: -0.5 * (A / (lam + 1e-8) + B * lam) - r * ep_costs / (s + 1e-8)
        f_b = lambda lam: -0.5 * (q / (lam + 1e-8) + 2 * target_kl * lam)
        lambda_star = la if f_a(la) >= f_b(lb) else lb
        nu_star = torch.clamp(lambda_star * ep_costs - r, min=0) / (s + 1e-8)
        step = 1.0 / (lambda_star + 1e-8) * (x - nu_star * p)
    else:
        lambda_star = torch.zeros(1)
        nu_star = torch.sqrt(2 * target_kl / (s + 1e-8))
        step = -nu_star * p
    return step, case, lambda_star, nu_star, alpha


def cpo_policy_update(pol, data, ep_costs, target_kl=0.01, cg_iters=15, search_steps=15, step_fraction=0.8, variant="cpo"):
    """cpo.py:351-519 (actor part).  Returns a dict of the logged / decided quantities.
    variant="pcpo": pcpo.py:371,392-401 -- the case analysis is replaced by the projection step
    sqrt(2 delta / (q+1e-8)) * F x  -  max(0, (sqrt(2 delta / q) r + c) / s) * p   (F x, not x: the reference's
    ``H_inv_g = fvp(x)``), optim_case 0, and the line search runs up to 200 steps (pcpo.py:44)."""
    obs, act, logp_old = data["obs"], data["act"], data["log_prob"]
    for p_ in pol.params("actor"):
        p_.grad = None
    loss_pi_r = -surrogate_loss(pol, obs, act, logp_old, data["adv_r"])
    loss_reward_before = loss_pi_r.item()
    with torch.no_grad():
        old_mean, old_std = actor_mean_std(pol, obs)
        old_mean, old_std = old_mean.clone(), old_std.clone()
    loss_pi_r.backward()
    g = -flat_grads(pol)
    Avp = lambda v: fvp_autograd(pol, obs, v)
    x = conjugate_gradients(Avp, g, cg_iters)
    xHx = torch.dot(x, Avp(x))
    for p_ in pol.params("actor"):
        p_.grad = None
    loss_pi_c = surrogate_loss(pol, obs, act, logp_old, data["adv_c"])
    loss_cost_before = loss_pi_c.item()
    loss_pi_c.backward()
    b = flat_grads(pol)
    p = conjugate_gradients(Avp, b, cg_iters)
    if variant == "pcpo":
        Fx = Avp(x)
        q, r, s_ = xHx, g.dot(p), b.dot(p)
        alpha = torch.sqrt(2 * target_kl / (xHx + 1e-8))
        step_dir = (torch.sqrt(2 * target_kl / (q + 1e-8)) * Fx
                    - torch.clamp_min((torch.sqrt(2 * target_kl / q) * r + ep_costs) / s_, torch.tensor(0.0)) * p)
        case, lambda_star, nu_star = 0, None, None
    else:
        step_dir, case, lambda_star, nu_star, alpha = cpo_step_direction(x, p, xHx, g, b, ep_costs, target_kl)
    step_frac = 1.0
    theta_old = flat_params(pol)
    expected = g.dot(step_dir)
    kl = torch.zeros(1)
    acceptance = 0
    accepted = False
    for step in range(search_steps):
        set_flat_params(pol, theta_old + step_frac * step_dir)
        acceptance = step + 1
        with torch.no_grad():
            loss_reward = -surrogate_loss(pol, obs, act, logp_old, data["adv_r"])
            loss_cost = surrogate_loss(pol, obs, act, logp_old, data["adv_c"])
            mean, std = actor_mean_std(pol, obs)
            kl = normal_kl(old_mean, old_std, mean, std).mean()
        improve = loss_reward_before - loss_reward.item()
        cost_diff = loss_cost.item() - loss_cost_before
        if not torch.isfinite(kl):
            continue  # cpo.py:499-501 (no step_frac decay on this branch)
        if (improve < 0) if case > 1 else False:
            pass
        elif cost_diff > max(-ep_costs, 0):
            pass
        elif kl > target_kl:
            pass
        else:
            accepted = True
            break
        step_frac *= step_fraction
    if not accepted:
        step_dir = torch.zeros_like(step_dir)
        acceptance = 0
    set_flat_params(pol, theta_old + step_frac * step_dir)
    return {"g": g, "b": b, "x": x, "p": p, "xHx": xHx, "case": case, "alpha": alpha,
            "step_dir": step_dir, "step_frac": step_frac, "acceptance": acceptance,
            "kl": float(kl), "loss_actor": (loss_pi_r + loss_pi_c).item(), "expected": expected,
            "lambda_star": lambda_star, "nu_star": nu_star}


def trpo_policy_update(pol, data, advantage, target_kl=0.01, cg_iters=15, search_steps=15):
    """trpo_lag.py:363-442."""
    obs, act, logp_old = data["obs"], data["act"], data["log_prob"]
    for p_ in pol.params("actor"):
        p_.grad = None
    theta_old = flat_params(pol)
    loss_pi = -surrogate_loss(pol, obs, act, logp_old, advantage)
    loss_before = loss_pi.item()
    with torch.no_grad():
        old_mean, old_std = actor_mean_std(pol, obs)
        old_mean, old_std = old_mean.clone(), old_std.clone()
    loss_pi.backward()
    g = -flat_grads(pol)
    Avp = lambda v: fvp_autograd(pol, obs, v)
    x = conjugate_gradients(Avp, g, cg_iters)
    xHx = torch.dot(x, Avp(x))
    alpha = torch.sqrt(2 * target_kl / (xHx + 1e-8))
    step_dir = x * alpha
    step_frac, final_kl, acceptance, accepted = 1.0, 0.0, 0, False
    for step in range(search_steps):
        set_flat_params(pol, theta_old + step_frac * step_dir)
        with torch.no_grad():
            loss_pi = -surrogate_loss(pol, obs, act, logp_old, advantage)
            mean, std = actor_mean_std(pol, obs)
            kl = normal_kl(old_mean, old_std, mean, std).mean().item()
        improve = loss_before - loss_pi.item()
        if not torch.isfinite(loss_pi):
            pass
        elif improve < 0:
            pass
        elif kl > target_kl:
            pass
        else:
            acceptance, final_kl, accepted = step + 1, kl, True
            break
        step_frac *= 0.8
    if not accepted:
        step_dir = torch.zeros_like(step_dir)
        acceptance = 0
    set_flat_params(pol, theta_old + step_frac * step_dir)
    return {"g": g, "x": x, "xHx": xHx, "alpha": alpha, "step_dir": step_dir, "step_frac": step_frac,
            "acceptance": acceptance, "kl": final_kl, "loss_actor": loss_pi.item()}


def npg_policy_update(pol, data, advantage, target_kl=0.01, cg_iters=15):
    """natural_pg.py:355-387 / rcpo.py: the TRPO direction x * sqrt(2 delta / xHx) taken at full length, no
    line search; Train/KL is KL(old || new).mean() after the step, Loss/Loss_actor the surrogate before it."""
    obs, act, logp_old = data["obs"], data["act"], data["log_prob"]
    for p_ in pol.params("actor"):
        p_.grad = None
    theta_old = flat_params(pol)
    loss_pi = -surrogate_loss(pol, obs, act, logp_old, advantage)
    with torch.no_grad():
        old_mean, old_std = actor_mean_std(pol, obs)
        old_mean, old_std = old_mean.clone(), old_std.clone()
    loss_pi.backward()
    g = -flat_grads(pol)
    Avp = lambda v: fvp_autograd(pol, obs, v)
    x = conjugate_gradients(Avp, g, cg_iters)
    xHx = torch.dot(x, Avp(x))
    alpha = torch.sqrt(2 * target_kl / (xHx + 1e-8))
    step_dir = x * alpha
    set_flat_params(pol, theta_old + step_dir)
    with torch.no_grad():
        mean, std = actor_mean_std(pol, obs)
        kl = normal_kl(old_mean, old_std, mean, std).mean().item()
    return {"g": g, "x": x, "xHx": xHx, "alpha": alpha, "step_dir": step_dir, "kl": kl, "loss_actor": loss_pi.mean().item()}


# --------------------------------------------------------------------------------------
# vectorised-over-envs restatement of gae_dual for full-size checks (same sequential
# arithmetic per env: separate fp32 mul/add/sub for delta, separate fp64 mul/add for the
# carry -- numpy ufuncs never fuse), validated against gae_dual in tests.
# --------------------------------------------------------------------------------------

def gae_dual_np(rew, cost, v_r, v_c, seg_end, boot_r, boot_c, gamma=0.99, lam=0.95, lam_c=0.95):
    rew, cost, v_r, v_c, boot_r, boot_c = (np.asarray(x, dtype=np.float32) for x in (rew, cost, v_r, v_c, boot_r, boot_c))
    seg = np.asarray(seg_end).astype(bool).copy()
    N, T = rew.shape
    seg[:, -1] = True
    g32 = np.float32(gamma)
    outs = [np.zeros((N, T), dtype=np.float32) for _ in range(4)]
    for (r, v, boot, l, oa, ot) in ((rew, v_r, boot_r, lam, outs[0], outs[2]), (cost, v_c, boot_c, lam_c, outs[1], outs[3])):
        disc = gamma * l
        acc = np.zeros(N, dtype=np.float64)
        for t in range(T - 1, -1, -1):
            end = seg[:, t]
            vnext = np.where(end, boot[:, t], v[:, t + 1] if t + 1 < T else boot[:, t]).astype(np.float32)
            delta = ((r[:, t] + g32 * vnext).astype(np.float32) - v[:, t]).astype(np.float32)
            d64 = delta.astype(np.float64)
            acc = np.where(end, d64, d64 + disc * acc)
            oa[:, t] = acc.astype(np.float32)
            ot[:, t] = (acc + v[:, t].astype(np.float64)).astype(np.float32)
    return [torch.from_numpy(o) for o in outs]
