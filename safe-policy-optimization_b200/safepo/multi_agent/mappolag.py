"""MAPPO-Lag on the device: the per-iteration part of the reference's ``Runner`` (safepo/multi_agent/mappolag.py:402-504,
583-597: ``collect`` / ``insert`` / ``compute`` / ``train``) around ``MultiAgentNets`` / ``MultiAgentTrainer``
(safepo/common/ma_model.py) and ``SeparatedReplayBuffer`` (safepo/common/buffer.py).  One runner holds all agents of one
GPU's environments; the environments themselves (Isaac Gym / multi-agent MuJoCo in the reference) are the caller's: it
feeds ``insert`` with what ``envs.step`` returned.  Random draws can be injected (``eps`` in ``collect``, ``agent_order`` and
``perms`` in ``train``) so that a run can be replayed against the oracle; by default they come from the device generator.
There is no CPU path."""
from __future__ import annotations

import torch

from safepo.common.buffer import SeparatedReplayBuffer
from safepo.common.ma_model import MultiAgentNets, MultiAgentTrainer


class Runner:
    def __init__(self, nets, config, obs_dim, share_obs_dim, act_dim):
        """``nets``: one MultiAgentNets per agent (all on the same device)."""
        self.config, self.num_agents = dict(config), len(nets)
        self.nets = list(nets)
        self.device = nets[0].device
        self.trainer = [MultiAgentTrainer(n, self.config) for n in self.nets]
        self.buffer = [SeparatedReplayBuffer(self.config, obs_dim, share_obs_dim, act_dim, self.device) for _ in self.nets]
        self.T, self.N = int(config["episode_length"]), int(config["n_rollout_threads"])

    def _dev(self, x):
        return torch.as_tensor(x).to(self.device)

    def warmup(self, obs, share_obs):
        """obs [N, agents, D], share_obs [N, agents, DS] of the reset (mappolag.py:396-405)."""
        obs, share_obs = self._dev(obs), self._dev(share_obs)
        for a, b in enumerate(self.buffer):
            b.share_obs[0].copy_(share_obs[:, a])
            b.obs[0].copy_(obs[:, a])

    @torch.no_grad()
    def collect(self, step, eps=None):
        """get_actions of every agent on its buffer's step-th observations (mappolag.py:408-445): values [N, agents, 1], the
        per-agent lists of actions / per-dimension log-probs, cost predictions [N, agents, 1]."""
        values, actions, logps, cost_preds = [], [], [], []
        for a, nets in enumerate(self.nets):
            b = self.buffer[a]
            e = None if eps is None else self._dev(eps[a]).contiguous()
            v, act, lp, cp = nets.get_actions(b.share_obs[step], b.obs[step], eps=e)
            values.append(v), actions.append(act), logps.append(lp), cost_preds.append(cp)
        return torch.stack(values, dim=1), actions, logps, torch.stack(cost_preds, dim=1)

    @torch.no_grad()
    def insert(self, obs, share_obs, rewards, costs, dones, values, actions, action_log_probs, cost_preds):
        """One environment step into every agent's buffer (mappolag.py:447-487): an environment whose agents are all done
        gets mask 0 (and active mask 1); an agent done alone gets active mask 0."""
        obs, share_obs, rewards, costs = self._dev(obs), self._dev(share_obs), self._dev(rewards), self._dev(costs)
        dones = self._dev(dones).bool()
        dones_env = torch.all(dones, dim=1)
        masks = torch.ones(self.N, self.num_agents, 1, device=self.device)
        masks[dones_env] = 0.0
        active_masks = torch.ones(self.N, self.num_agents, 1, device=self.device)
        active_masks[dones] = 0.0
        active_masks[dones_env] = 1.0
        for a, b in enumerate(self.buffer):
            b.insert(share_obs[:, a], obs[:, a], actions[a], action_log_probs[a], values[:, a], rewards[:, a], masks[:, a],
                     active_masks[:, a], costs=costs[:, a], cost_preds=cost_preds[:, a])

    @torch.no_grad()
    def compute(self):
        """Bootstrap values of the last observations and the masked GAE returns of both critics (mappolag.py:583-597)."""
        for nets, b, tr in zip(self.nets, self.buffer, self.trainer):
            mean, sd = tr.popart_mean_sqrt_var()
            b.compute_returns(nets._value(nets.critic, b.share_obs[-1]), mean, sd)
            b.compute_cost_returns(nets._value(nets.cost_critic, b.share_obs[-1]), mean, sd)

    def train(self, agent_order=None, perms=None):
        """Sequential update of the agents in a random order; every updated agent multiplies the importance ratio of its new
        against its old policy into the factor the following agents see (mappolag.py:489-519)."""
        T, N = self.T, self.N
        factor = torch.ones(T, N, 1, device=self.device)
        order = torch.randperm(self.num_agents).tolist() if agent_order is None else [int(a) for a in agent_order]
        for a in order:
            b, tr, nets = self.buffer[a], self.trainer[a], self.nets[a]
            A = b.actions.shape[-1]
            b.update_factor(factor)
            flat_obs, flat_act = b.obs[:-1].reshape(-1, b.obs.shape[-1]), b.actions.reshape(-1, A)
            with torch.no_grad():
                old_lp = nets.evaluate_actions(flat_obs, flat_act)
            tr.train(b, perms=None if perms is None else perms[a])
            with torch.no_grad():
                new_lp = nets.evaluate_actions(flat_obs, flat_act)
                factor = factor * torch.prod(torch.exp(new_lp - old_lp).reshape(T, N, A), dim=-1, keepdim=True)
            b.after_update()
        return order

    def return_aver_cost(self, aver_episode_costs):
        for b in self.buffer:
            b.return_aver_insert(aver_episode_costs)


__all__ = ["Runner", "MultiAgentNets", "MultiAgentTrainer", "SeparatedReplayBuffer"]
