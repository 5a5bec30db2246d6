"""CPU oracle, part 2: the training loops (rollout control R0 + update) of
safepo/single_agent/{ppo_lag,focops,cpo,trpo_lag}.py restated on top of spo_oracle.

TEST INFRASTRUCTURE ONLY (see spo_oracle.py).  ``train()`` reproduces the reference's
``main()`` bit-for-bit on CPU for a vector env passed in by the caller (checked against
the real reference in tests/golden/make_golden.py -> tests/test_oracle_golden.py), and is
what bench.py times as the CPU arm.
"""
from __future__ import annotations

import random
import time
from collections import deque
from types import SimpleNamespace

import numpy as np
import torch

from . import spo_oracle as O

ALGO_CFG = {
    # ppo_lag.py:45-52, focops.py:47-54, cpo.py:47-54, trpo_lag.py:47-54
    "ppo_lag": dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=0.02, batch_size=64, learning_iters=40, max_grad_norm=40.0),
    "focops": dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=0.02, batch_size=64, learning_iters=40, max_grad_norm=40.0),
    "cpo": dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=0.01, batch_size=128, learning_iters=10, max_grad_norm=40.0),
    "trpo_lag": dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=0.01, batch_size=128, learning_iters=10, max_grad_norm=40.0),
}
# siblings built from the same templates: ppo.py / pg.py / cppo_pid.py are ppo_lag.py with the
# Lagrange lines removed / the clip removed / PIDLagrangian swapped in; trpo.py is trpo_lag.py without Lagrange
for _a, _base in (("ppo", "ppo_lag"), ("pg", "ppo_lag"), ("cppo_pid", "ppo_lag"), ("trpo", "trpo_lag"), ("natural_pg", "trpo_lag"),
                  ("rcpo", "trpo_lag"), ("pcpo", "cpo"), ("cup", "ppo_lag")):
    ALGO_CFG[_a] = dict(ALGO_CFG[_base])
PG_FAMILY = ("ppo_lag", "focops", "ppo", "pg", "cppo_pid", "cup")   # minibatch policy-gradient updates with KL early stop
PG_KIND = {"ppo_lag": "ppo", "ppo": "ppo", "cppo_pid": "ppo", "pg": "pg", "focops": "focops", "cup": "ppo"}
CUP_LAMBDA, CUP_NU = 0.95, 0.20            # cup.py:45-46
NO_LAGRANGE = ("cpo", "pcpo", "ppo", "pg", "trpo", "natural_pg")
NPG_FAMILY = ("natural_pg", "rcpo")        # trpo.py / trpo_lag.py without the line search


def default_args(**kw):
    """Fields of single_agent_args(), safepo/utils/config.py:145-162."""
    a = dict(seed=0, use_eval=False, task="SafetyPointGoal1-v0", num_envs=10, experiment="single_agent_exp",
             log_dir="../runs", device="cpu", device_id=0, write_terminal=True, headless=False,
             total_steps=10000000, steps_per_epoch=20000, randomize=False, cost_limit=25.0,
             lagrangian_multiplier_init=0.001, lagrangian_multiplier_lr=0.035)
    a.update(kw)
    return SimpleNamespace(**a)


class StatLog:
    """The slice of EpochLogger semantics that feeds numbers back into training
    (safepo/common/logger.py:344-373): get_stats is 0.0 until the key was logged once,
    afterwards the mean of what was stored since the last log_tabular (NaN if nothing)."""

    def __init__(self):
        self.epoch_dict, self.log_headers, self.row, self.rows = {}, [], {}, []
        self.first_row, self.logged = True, True

    def store(self, **kw):
        for k, v in kw.items():
            self.epoch_dict.setdefault(k, []).append(v)

    def log_tabular(self, key, val=None):
        if val is None:
            val = np.mean(self.epoch_dict[key])
            self.epoch_dict[key] = []
        if self.first_row:
            self.log_headers.append(key)
        self.row[key] = val

    def get_stats(self, key):
        if key not in self.log_headers:
            return 0.0
        return np.mean(self.epoch_dict[key])

    def dump_tabular(self):
        self.logged = True
        self.rows.append(dict(self.row))
        self.row = {}
        self.first_row = False


class PathBuffer:
    """VectorizedOnPolicyBuffer (safepo/common/buffer.py:24-164) as [N,T] arrays."""

    KEYS = ("reward", "cost", "value_r", "value_c", "log_prob", "adv_r", "adv_c", "target_value_r", "target_value_c")

    def __init__(self, N, T, D, A, gamma, lam=0.95, lam_c=0.95):
        self.N, self.T, self.gamma, self.lam, self.lam_c = N, T, gamma, lam, lam_c
        self.obs = torch.zeros(N, T, D)
        self.act = torch.zeros(N, T, A)
        for k in self.KEYS:
            setattr(self, k, torch.zeros(N, T))
        self.seg_end = torch.zeros(N, T, dtype=torch.uint8)
        self.boot_r, self.boot_c = torch.zeros(N, T), torch.zeros(N, T)
        self.ptr = [0] * N
        self.start = [0] * N

    def store(self, t, obs, act, reward, cost, value_r, value_c, log_prob):
        self.obs[:, t], self.act[:, t] = obs, act
        self.reward[:, t], self.cost[:, t] = reward, cost
        self.value_r[:, t], self.value_c[:, t], self.log_prob[:, t] = value_r, value_c, log_prob
        for i in range(self.N):
            self.ptr[i] += 1

    def finish_path(self, last_r, last_c, idx):
        s, e = self.start[idx], self.ptr[idx]
        sl = slice(s, e)
        self.seg_end[idx, e - 1] = 1
        self.boot_r[idx, e - 1], self.boot_c[idx, e - 1] = last_r, last_c
        a, tg = O.gae_path(torch.cat([self.value_r[idx, sl], last_r.reshape(1)]),
                           torch.cat([self.reward[idx, sl], last_r.reshape(1)]), self.lam, self.gamma)
        self.adv_r[idx, sl], self.target_value_r[idx, sl] = a, tg
        a, tg = O.gae_path(torch.cat([self.value_c[idx, sl], last_c.reshape(1)]),
                           torch.cat([self.cost[idx, sl], last_c.reshape(1)]), self.lam_c, self.gamma)
        self.adv_c[idx, sl], self.target_value_c[idx, sl] = a, tg
        self.start[idx] = e

    def get(self):
        S = self.N * self.T
        data = {"obs": self.obs.reshape(S, -1).clone(), "act": self.act.reshape(S, -1).clone()}
        for k in self.KEYS:
            data[k] = getattr(self, k).reshape(S).clone()
        mean, std, cmean = data["adv_r"].mean(), data["adv_r"].std(), data["adv_c"].mean()
        data["adv_r"] = (data["adv_r"] - mean) / (std + 1e-8)
        data["adv_c"] = data["adv_c"] - cmean
        self.ptr, self.start = [0] * self.N, [0] * self.N
        self.seg_end.zero_()
        return data


def rollout(pol, env, buf, obs, ep, deques, log, T, epoch_T=None):
    """ppo_lag.py:162-235 (identical in cpo/focops/trpo_lag).  ``obs`` is the current
    fp32 observation tensor; returns the next one.  ``epoch_T`` (bench only): length of the
    epoch when just a window of T steps of it is being run."""
    N = buf.N
    ep_ret, ep_cost, ep_len = ep
    rew_dq, cost_dq, len_dq = deques
    for t in range(T):
        with torch.no_grad():
            act, logp, v_r, v_c = O.policy_step(pol, obs)
        next_obs, reward, cost, terminated, truncated, info = env.step(act.detach().squeeze().cpu().numpy())
        ep_ret += reward
        ep_cost += cost
        ep_len += 1
        next_obs, reward, cost, terminated, truncated = (
            torch.as_tensor(x, dtype=torch.float32) for x in (next_obs, reward, cost, terminated, truncated))
        final_obs = None
        if "final_observation" in info:
            final_obs = torch.as_tensor(np.array([a if a is not None else np.zeros(obs.shape[-1])
                                                  for a in info["final_observation"]]), dtype=torch.float32)
        buf.store(t, obs, act, reward, cost, v_r, v_c, logp)
        obs = next_obs
        epoch_end = t >= (epoch_T if epoch_T is not None else T) - 1
        for idx in range(N):
            done, time_out = bool(terminated[idx]), bool(truncated[idx])
            if epoch_end or done or time_out:
                last_r, last_c = torch.zeros(1), torch.zeros(1)
                if not done:
                    if epoch_end:
                        with torch.no_grad():
                            _, _, last_r, last_c = O.policy_step(pol, obs[idx])
                    if time_out:
                        with torch.no_grad():
                            _, _, last_r, last_c = O.policy_step(pol, final_obs[idx])
                    last_r, last_c = last_r.unsqueeze(0), last_c.unsqueeze(0)
                if done or time_out:
                    rew_dq.append(ep_ret[idx]); cost_dq.append(ep_cost[idx]); len_dq.append(ep_len[idx])
                    log.store(**{"Metrics/EpRet": np.mean(rew_dq), "Metrics/EpCost": np.mean(cost_dq),
                                 "Metrics/EpLen": np.mean(len_dq)})
                    ep_ret[idx] = ep_cost[idx] = ep_len[idx] = 0.0
                    log.logged = False
                buf.finish_path(last_r, last_c, idx)
    return obs


def train(algo, args, env, max_epochs=None, hooks=None):
    """main() of the four algorithm scripts.  Returns (policy, StatLog, timing dict)."""
    cfg = ALGO_CFG[algo]
    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    torch.set_num_threads(getattr(args, "torch_threads", 4))  # ppo_lag.py:73
    D, A = env.observation_space.shape[0], env.action_space.shape[0]
    N = args.num_envs
    T = args.steps_per_epoch // N
    epochs = args.total_steps // args.steps_per_epoch
    if max_epochs is not None:
        epochs_run = min(epochs, max_epochs)
    else:
        epochs_run = epochs
    pol = O.OraclePolicy(D, A, cfg["hidden_sizes"])
    trust = algo in ("cpo", "pcpo", "trpo_lag", "trpo") + NPG_FAMILY
    opt = O.OracleOptim(pol, lr=3e-4, critic_lr=1e-3 if trust else 3e-4, epochs=epochs)
    buf = PathBuffer(N, T, D, A, cfg["gamma"])
    lagrange = None
    if algo == "cppo_pid":
        lagrange = O.OraclePIDLagrange(args.cost_limit, args.lagrangian_multiplier_init)      # cppo_pid.py:128-131
    elif algo not in NO_LAGRANGE:
        lagrange = O.OracleLagrange(args.cost_limit, args.lagrangian_multiplier_init, args.lagrangian_multiplier_lr,
                                    upper_bound=2.0 if algo == "focops" else (CUP_NU if algo == "cup" else None))
    log = StatLog()
    deques = (deque(maxlen=50), deque(maxlen=50), deque(maxlen=50))
    obs, _ = env.reset()
    obs = torch.as_tensor(obs, dtype=torch.float32)
    ep = (np.zeros(N), np.zeros(N), np.zeros(N))
    times = {"rollout": [], "update": []}
    for epoch in range(epochs_run):
        t0 = time.time()
        obs = rollout(pol, env, buf, obs, ep, deques, log, T)
        t1 = time.time()
        extra = {}
        ep_costs_stat = log.get_stats("Metrics/EpCost")      # read in every script (ppo_lag.py:272), used or not
        if lagrange is not None:
            lagrange.update_lagrange_multiplier(ep_costs_stat)
        data = buf.get()

        def mixed_advantage():
            if lagrange is None:
                return data["adv_r"]                          # ppo.py:272, trpo.py:361
            lam = lagrange.lagrangian_multiplier
            adv = data["adv_r"] - lam * data["adv_c"]
            adv /= (lam + 1)
            return adv

        if algo in PG_FAMILY:
            advantage = data["adv_r"] if algo == "cup" else mixed_advantage()      # cup.py:284: first stage = plain PPO
            res = O.pg_update(pol, opt, data, advantage, kind=PG_KIND[algo],
                              batch_size=cfg["batch_size"], learning_iters=cfg["learning_iters"],
                              target_kl=cfg["target_kl"], max_grad_norm=cfg["max_grad_norm"])
            for lr_, lc_, lp_ in res["losses"]:
                log.store(**{"Loss/Loss_reward_critic": lr_, "Loss/Loss_cost_critic": lc_, "Loss/Loss_actor": lp_})
            extra = {"Train/StopIter": res["stop_iter"], "Train/KL": res["kl"]}
            if algo == "cup":
                res2 = O.cup_second_stage(pol, opt, data, lagrange.lagrangian_multiplier, gamma=cfg["gamma"], cup_lambda=CUP_LAMBDA,
                                          batch_size=cfg["batch_size"], learning_iters=cfg["learning_iters"],
                                          target_kl=cfg["target_kl"], max_grad_norm=cfg["max_grad_norm"])
                extra.update({"Train/SeconStageStopIter": res2["stop_iter"], "Train/KL": res2["kl"]})
        else:
            if algo in ("cpo", "pcpo"):
                ep_costs = log.get_stats("Metrics/EpCost") - args.cost_limit
                res = O.cpo_policy_update(pol, data, ep_costs, target_kl=cfg["target_kl"], variant=algo,
                                          search_steps=200 if algo == "pcpo" else 15)
            elif algo in NPG_FAMILY:
                res = O.npg_policy_update(pol, data, mixed_advantage(), target_kl=cfg["target_kl"])
            else:
                res = O.trpo_policy_update(pol, data, mixed_advantage(), target_kl=cfg["target_kl"])
            log.store(**{"Misc/Alpha": res["alpha"].item(), "Misc/FinalStepNorm": torch.norm(res["step_dir"]).mean().item(),
                         "Misc/xHx": res["xHx"].item(), "Misc/gradient_norm": torch.norm(res["g"]).mean().item(),
                         "Misc/H_inv_g": res["x"].norm().item(), "Loss/Loss_actor": res["loss_actor"], "Train/KL": res["kl"]})
            if algo not in NPG_FAMILY:
                log.store(**{"Misc/AcceptanceStep": res["acceptance"]})
            for lr_, lc_ in O.critic_regression(pol, opt, data, batch_size=cfg["batch_size"],
                                                learning_iters=cfg["learning_iters"], max_grad_norm=cfg["max_grad_norm"]):
                log.store(**{"Loss/Loss_reward_critic": lr_, "Loss/Loss_cost_critic": lc_})
            extra = {"trust": res}
        t2 = time.time()
        if hooks and "after_update" in hooks:
            hooks["after_update"](epoch, pol, data, extra)
        if algo in PG_FAMILY:
            opt.scheduler_step()
        times["rollout"].append(t1 - t0); times["update"].append(t2 - t1)
        if not log.logged:
            for k in ("Metrics/EpRet", "Metrics/EpCost", "Metrics/EpLen"):
                log.log_tabular(k)
            log.log_tabular("Train/Epoch", epoch + 1)
            log.log_tabular("Train/TotalSteps", (epoch + 1) * args.steps_per_epoch)
            if algo in PG_FAMILY:
                log.log_tabular("Train/StopIter", extra["Train/StopIter"])
                if algo == "cup":
                    log.log_tabular("Train/SeconStageStopIter", extra["Train/SeconStageStopIter"])
                log.log_tabular("Train/KL", extra["Train/KL"])
                if lagrange is not None:
                    log.log_tabular("Train/LagragianMultiplier", lagrange.lagrangian_multiplier)
                log.log_tabular("Train/LR", opt.actor_lr())
            else:
                if algo in ("trpo_lag", "rcpo"):
                    log.log_tabular("Train/LagragianMultiplier", lagrange.lagrangian_multiplier)
                log.log_tabular("Train/KL")
            for k in ("Loss/Loss_reward_critic", "Loss/Loss_cost_critic", "Loss/Loss_actor"):
                log.log_tabular(k)
            log.log_tabular("Time/Rollout", t1 - t0)
            log.log_tabular("Time/Update", t2 - t1)
            log.log_tabular("Time/Total", t2 - t0)
            log.log_tabular("Value/RewardAdv", data["adv_r"].mean().item())
            log.log_tabular("Value/CostAdv", data["adv_c"].mean().item())
            if trust:
                for k in ("Misc/Alpha", "Misc/FinalStepNorm", "Misc/xHx", "Misc/gradient_norm", "Misc/H_inv_g") + \
                        (() if algo in NPG_FAMILY else ("Misc/AcceptanceStep",)):
                    log.log_tabular(k)
            log.dump_tabular()
    return pol, log, times
