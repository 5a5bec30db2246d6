"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of row G2 of SURVEY section 8 -- the
multi-agent buffer's masked GAE with PopArt de-normalisation (BASELINE config 5, MAPPO-Lag).

Follows safepo/common/buffer.py:356-384 (SeparatedReplayBuffer.compute_returns / compute_cost_returns) and
safepo/common/popart.py:46-132 (PopArt), same torch ops in the same order, fp32 throughout (unlike the
single-agent path there is no float64 carry here).  Pinned bit for bit against the reference's own classes by
tests/golden/make_golden.py (fixture ma_gae.pt) and tests/test_oracle_golden.py.

Layout: time-major [T+1, N, 1] value predictions / masks, [T, N, 1] rewards -- as the reference stores them."""
import torch


class OraclePopArt:
    """PopArt(1) as MAPPO-Lag uses it (mappolag.py:119): one scalar running mean / mean-square / debias term."""

    def __init__(self, input_shape=1, norm_axes=1, beta=0.99999, per_element_update=False, epsilon=1e-5):
        self.norm_axes, self.beta, self.per_element_update, self.epsilon = norm_axes, beta, per_element_update, epsilon
        self.running_mean = torch.zeros(input_shape, dtype=torch.float32)
        self.running_mean_sq = torch.zeros(input_shape, dtype=torch.float32)
        self.debiasing_term = torch.tensor(0.0, dtype=torch.float32)

    def running_mean_var(self):
        """popart.py:64-74."""
        mean = self.running_mean / self.debiasing_term.clamp(min=self.epsilon)
        mean_sq = self.running_mean_sq / self.debiasing_term.clamp(min=self.epsilon)
        var = (mean_sq - mean ** 2).clamp(min=1e-2)
        return mean, var

    def normalize(self, x, train=True):
        """popart.py:76-112 (forward)."""
        x = x.to(torch.float32)
        if train:
            d = x.detach()
            dims = tuple(range(self.norm_axes))
            batch_mean = d.mean(dim=dims)
            batch_sq_mean = (d ** 2).mean(dim=dims)
            if self.per_element_update:
                n = 1
                for s in d.size()[:self.norm_axes]:
                    n *= s
                weight = self.beta ** n
            else:
                weight = self.beta
            self.running_mean.mul_(weight).add_(batch_mean * (1.0 - weight))
            self.running_mean_sq.mul_(weight).add_(batch_sq_mean * (1.0 - weight))
            self.debiasing_term.mul_(weight).add_(1.0 * (1.0 - weight))
        mean, var = self.running_mean_var()
        return (x - mean[(None,) * self.norm_axes]) / torch.sqrt(var)[(None,) * self.norm_axes]

    def denormalize(self, x):
        """popart.py:114-132."""
        x = x.to(torch.float32)
        mean, var = self.running_mean_var()
        return x * torch.sqrt(var)[(None,) * self.norm_axes] + mean[(None,) * self.norm_axes]


def masked_gae(rewards, value_preds, masks, popart, gamma, gae_lambda):
    """buffer.py:371-376 / :378-384.  value_preds [T+1,N,1] already holds the bootstrap in its last slot.
    Returns ``returns`` [T,N,1] (the reference's buffer keeps an unused zero row T)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    gae = 0
    for step in reversed(range(T)):
        delta = rewards[step] + gamma * popart.denormalize(value_preds[step + 1]) * masks[step + 1] - popart.denormalize(value_preds[step])
        gae = delta + gamma * gae_lambda * masks[step + 1] * gae
        returns[step] = gae + popart.denormalize(value_preds[step])
    return returns


# ---------------------------------------------------------------------------------------------------------------------
# MAPPO-Lag networks and one ppo_update (BASELINE config 5).  Restates
#   safepo/utils/mlp.py:9-61 (MLPBase: LayerNorm -> [Linear, ELU, LayerNorm] x (1 + layer_N)),
#   safepo/utils/distributions.py:21-42 (DiagGaussian: std = sigmoid(log_std / std_x_coef) * std_y_coef),
#   safepo/utils/act.py:24-60 (per-dimension log-probs, entropy mean), safepo/common/model.py:172-363,
#   safepo/multi_agent/mappolag.py:46-113 (MAPPO_L_Policy) and :121-199 (cal_value_loss, ppo_update),
#   safepo/utils/util.py:19-22 (huber_loss -- one-sided: errors below -delta contribute 0, reproduced).
# Parameters are kept under the reference's own state_dict names.  Pinned by tests/golden/make_golden.py (ma_update.pt:
# outputs of the reference's MAPPO_L_Policy / MAPPO_L_Trainer) through tests/test_oracle_golden.py.
# ---------------------------------------------------------------------------------------------------------------------
import torch.nn.functional as F


class OracleMANet:
    """One MultiAgentActor / MultiAgentCritic as a dict of leaf tensors (reference state_dict names)."""

    def __init__(self, state, layer_N=2):
        self.p = {k: v.detach().clone().requires_grad_(True) for k, v in state.items()}
        self.layer_N = layer_N

    def params(self):
        return list(self.p.values())

    def state(self):
        return {k: v.detach().clone() for k, v in self.p.items()}

    def features(self, x):
        """MLPBase.forward (mlp.py:56-61) with use_feature_normalization and ELU (mlp.py:16: both branches are ELU)."""
        p = self.p
        x = F.layer_norm(x, (x.shape[-1],), p["base.feature_norm.weight"], p["base.feature_norm.bias"])
        x = F.layer_norm(F.elu(F.linear(x, p["base.mlp.fc1.0.weight"], p["base.mlp.fc1.0.bias"])),
                         (p["base.mlp.fc1.2.weight"].shape[0],), p["base.mlp.fc1.2.weight"], p["base.mlp.fc1.2.bias"])
        for i in range(self.layer_N):
            w, b = p[f"base.mlp.fc2.{i}.0.weight"], p[f"base.mlp.fc2.{i}.0.bias"]
            x = F.layer_norm(F.elu(F.linear(x, w, b)), (w.shape[0],), p[f"base.mlp.fc2.{i}.2.weight"], p[f"base.mlp.fc2.{i}.2.bias"])
        return x


def ma_actor_dist(actor, obs, std_x_coef=1.0, std_y_coef=0.5):
    """DiagGaussian.forward (distributions.py:38-42)."""
    feat = actor.features(obs)
    mean = F.linear(feat, actor.p["act.action_out.fc_mean.weight"], actor.p["act.action_out.fc_mean.bias"])
    std = torch.sigmoid(actor.p["act.action_out.log_std"] / std_x_coef) * std_y_coef
    return torch.distributions.Normal(mean, std)


def ma_critic_value(critic, cent_obs):
    return F.linear(critic.features(cent_obs), critic.p["v_out.weight"], critic.p["v_out.bias"])


def ma_get_actions(actor, critic, cost_critic, cent_obs, obs, deterministic=False):
    """MAPPO_L_Policy.get_actions (mappolag.py:69-82): values [N,1], actions [N,A], per-dimension log-probs [N,A], cost
    predictions [N,1].  Sampling consumes torch's global generator exactly like Normal.sample()."""
    with torch.no_grad():
        dist = ma_actor_dist(actor, obs)
        actions = dist.mean if deterministic else dist.sample()
        return ma_critic_value(critic, cent_obs), actions, dist.log_prob(actions), ma_critic_value(cost_critic, cent_obs)


def huber_loss(e, d):
    """util.py:19-22 verbatim semantics: quadratic for |e| <= d, linear for e > d, ZERO for e < -d."""
    a = (abs(e) <= d).float()
    b = (e > d).float()
    return a * e ** 2 / 2 + b * d * (abs(e) - d / 2)


class OracleMATrainer:
    """MAPPO_L_Trainer (mappolag.py:115-199) around three OracleMANets: Adam(lr, eps=opti_eps) x 3, one PopArt shared by the
    reward and the cost value loss (mappolag.py:119,125-126: every cal_value_loss call UPDATES it twice), lamda_lagr."""

    def __init__(self, actor, critic, cost_critic, cfg):
        self.actor, self.critic, self.cost_critic, self.cfg = actor, critic, cost_critic, cfg
        self.opt_a = torch.optim.Adam(actor.params(), lr=cfg["actor_lr"], eps=cfg["opti_eps"], weight_decay=cfg["weight_decay"])
        self.opt_c = torch.optim.Adam(critic.params(), lr=cfg["critic_lr"], eps=cfg["opti_eps"], weight_decay=cfg["weight_decay"])
        self.opt_k = torch.optim.Adam(cost_critic.params(), lr=cfg["critic_lr"], eps=cfg["opti_eps"], weight_decay=cfg["weight_decay"])
        self.popart = OraclePopArt(1)
        self.lamda_lagr = cfg["lamda_lagr"]

    def cal_value_loss(self, values, value_preds_batch, return_batch):
        c = self.cfg
        value_pred_clipped = value_preds_batch + (values - value_preds_batch).clamp(-c["clip_param"], c["clip_param"])
        error_clipped = self.popart.normalize(return_batch) - value_pred_clipped       # PopArt.forward: updates, then normalises
        error_original = self.popart.normalize(return_batch) - values
        return torch.max(huber_loss(error_original, c["huber_delta"]), huber_loss(error_clipped, c["huber_delta"])).mean()

    def ppo_update(self, s):
        """s: dict with share_obs, obs, actions, value_preds, returns, old_action_log_probs [N,A], adv_targ, factor,
        cost_preds, cost_returns, cost_adv_targ, aver_episode_costs (use_policy_active_masks False, no recurrence)."""
        c = self.cfg
        dist = ma_actor_dist(self.actor, s["obs"])
        action_log_probs = dist.log_prob(s["actions"])
        dist_entropy = dist.entropy().mean()
        values = ma_critic_value(self.critic, s["share_obs"])
        cost_values = ma_critic_value(self.cost_critic, s["share_obs"])
        adv_targ_hybrid = s["adv_targ"] - self.lamda_lagr * s["cost_adv_targ"]
        imp_weights = torch.prod(torch.exp(action_log_probs - s["old_action_log_probs"]), dim=-1, keepdim=True)
        surr1 = imp_weights * adv_targ_hybrid
        surr2 = torch.clamp(imp_weights, 1.0 - c["clip_param"], 1.0 + c["clip_param"]) * adv_targ_hybrid
        policy_loss = -torch.sum(s["factor"] * torch.min(surr1, surr2), dim=-1, keepdim=True).mean()
        self.opt_a.zero_grad()
        (policy_loss - dist_entropy * c["entropy_coef"]).backward()
        actor_grad_norm = torch.nn.utils.clip_grad_norm_(self.actor.params(), c["max_grad_norm"])
        self.opt_a.step()
        delta = -((s["aver_episode_costs"].mean() - c["cost_limit"]) * (1 - c["gamma"]) + (imp_weights * s["cost_adv_targ"])).mean().detach()
        self.lamda_lagr = torch.nn.ReLU()(self.lamda_lagr - (delta * c["lagrangian_coef_rate"]))
        value_loss = self.cal_value_loss(values, s["value_preds"], s["returns"])
        self.opt_c.zero_grad()
        (value_loss * c["value_loss_coef"]).backward()
        critic_grad_norm = torch.nn.utils.clip_grad_norm_(self.critic.params(), c["max_grad_norm"])
        self.opt_c.step()
        cost_loss = self.cal_value_loss(cost_values, s["cost_preds"], s["cost_returns"])
        self.opt_k.zero_grad()
        (cost_loss * c["value_loss_coef"]).backward()
        cost_grad_norm = torch.nn.utils.clip_grad_norm_(self.cost_critic.params(), c["max_grad_norm"])
        self.opt_k.step()
        return dict(value_loss=value_loss.detach(), critic_grad_norm=critic_grad_norm, policy_loss=policy_loss.detach(),
                    dist_entropy=dist_entropy.detach(), actor_grad_norm=actor_grad_norm, imp_weights=imp_weights.detach(),
                    cost_loss=cost_loss.detach(), cost_grad_norm=cost_grad_norm)


# ---------------------------------------------------------------------------------------------------------------------
# MAPPO-Lag replay buffer, trainer.train and the runner's iteration (the callers of ppo_update).
# Restates safepo/common/buffer.py:209-465 (SeparatedReplayBuffer: insert :281-337, after_update :339-348, compute_returns /
# compute_cost_returns :356-384 = masked_gae above, feed_forward_generator :386-465), safepo/multi_agent/mappolag.py:200-234
# (MAPPO_L_Trainer.train) and :402-504,583-597 (Runner.collect / insert / train / compute), same torch ops in the same order and
# the same draws from torch's global generator (Normal.sample per agent in collect; randperm over the agents and over the batch
# in train).  Pinned by tests/golden/ma_epoch.pt (two iterations of two agents through the reference's own methods).
# Reference quirk kept: train() marks inactive entries with NaN and then takes torch.mean / torch.std (not nanmean), so the
# advantages of a buffer with ANY inactive entry are all NaN; only the all-active case is meaningful (and pinned).
# ---------------------------------------------------------------------------------------------------------------------
class OracleMABuffer:
    def __init__(self, T, N, obs_dim, share_obs_dim, act_dim, gamma, gae_lambda):
        self.T, self.N, self.gamma, self.gae_lambda = T, N, gamma, gae_lambda
        z = lambda *s: torch.zeros(*s, dtype=torch.float32)      # noqa: E731
        self.aver_episode_costs = z(T + 1, N, obs_dim)           # buffer.py:240 (sic: observation-shaped until the first return_aver_insert)
        self.share_obs, self.obs = z(T + 1, N, share_obs_dim), z(T + 1, N, obs_dim)
        self.value_preds, self.returns = z(T + 1, N, 1), z(T + 1, N, 1)
        self.cost_preds, self.cost_returns = z(T + 1, N, 1), z(T + 1, N, 1)
        self.actions, self.action_log_probs = z(T, N, act_dim), z(T, N, act_dim)
        self.rewards, self.costs = z(T, N, 1), z(T, N, 1)
        self.masks, self.active_masks = torch.ones(T + 1, N, 1), torch.ones(T + 1, N, 1)
        self.factor = torch.ones(T, N, 1)
        self.step = 0

    def insert(self, share_obs, obs, actions, action_log_probs, value_preds, rewards, masks, active_masks, costs, cost_preds):
        s = self.step
        self.share_obs[s + 1].copy_(share_obs)
        self.obs[s + 1].copy_(obs)
        self.actions[s].copy_(actions)
        self.action_log_probs[s].copy_(action_log_probs)
        self.value_preds[s].copy_(value_preds)
        self.rewards[s].copy_(rewards)
        self.masks[s + 1].copy_(masks)
        self.active_masks[s + 1].copy_(active_masks)
        self.costs[s].copy_(costs)
        self.cost_preds[s].copy_(cost_preds)
        self.step = (s + 1) % self.T

    def after_update(self):
        for t in (self.share_obs, self.obs, self.masks, self.active_masks):
            t[0].copy_(t[-1])

    def compute_returns(self, next_value, popart):
        self.value_preds[-1] = next_value
        self.returns[:-1] = masked_gae(self.rewards, self.value_preds, self.masks, popart, self.gamma, self.gae_lambda)

    def compute_cost_returns(self, next_cost, popart):
        self.cost_preds[-1] = next_cost
        self.cost_returns[:-1] = masked_gae(self.costs, self.cost_preds, self.masks, popart, self.gamma, self.gae_lambda)

    def whole_batch_sample(self, advantages, cost_adv, perm=None):
        """feed_forward_generator with num_mini_batch = 1: ONE sample holding every row in a random order (randperm from the
        global generator; ``perm`` replaces the draw when a test needs the same order on both sides)."""
        idx = torch.randperm(self.T * self.N) if perm is None else perm
        flat = lambda t: t.reshape(-1, t.shape[-1])[idx]        # noqa: E731
        return dict(share_obs=flat(self.share_obs[:-1]), obs=flat(self.obs[:-1]), actions=flat(self.actions),
                    value_preds=flat(self.value_preds[:-1]), returns=flat(self.returns[:-1]),
                    old_action_log_probs=flat(self.action_log_probs), adv_targ=flat(advantages), factor=flat(self.factor),
                    cost_preds=flat(self.cost_preds[:-1]), cost_returns=flat(self.cost_returns[:-1]), cost_adv_targ=flat(cost_adv),
                    aver_episode_costs=self.aver_episode_costs)


def ma_trainer_train(trainer, buf, learning_iters, perms=None):
    """MAPPO_L_Trainer.train (mappolag.py:200-234): advantages = returns - denormalised predictions, standardised with the
    (NaN-propagating, see the header) mean / unbiased std over the active entries; learning_iters whole-batch updates."""
    def standardise(ret, pred):
        adv = ret[:-1] - trainer.popart.denormalize(pred[:-1])
        copy = adv.clone()
        copy[buf.active_masks[:-1] == 0.0] = float("nan")
        return (adv - torch.mean(copy)) / (torch.std(copy) + 1e-8)
    advantages = standardise(buf.returns, buf.value_preds)
    cost_adv = standardise(buf.cost_returns, buf.cost_preds)
    out = None
    for it in range(learning_iters):
        out = trainer.ppo_update(buf.whole_batch_sample(advantages, cost_adv, None if perms is None else perms[it]))
    return out


class OracleMARunner:
    """The per-iteration part of Runner (mappolag.py:300-345 without environments and logging): agents = list of
    (actor, critic, cost_critic) OracleMANets."""

    def __init__(self, nets, cfg, T, N, obs_dim, share_obs_dim, act_dim):
        self.cfg, self.T, self.N, self.num_agents = cfg, T, N, len(nets)
        self.nets = nets
        self.trainer = [OracleMATrainer(a, c, k, cfg) for a, c, k in nets]
        self.buffer = [OracleMABuffer(T, N, obs_dim, share_obs_dim, act_dim, cfg["gamma"], cfg["gae_lambda"]) for _ in nets]

    def warmup(self, obs, share_obs):
        for a, b in enumerate(self.buffer):
            b.share_obs[0].copy_(share_obs[:, a])
            b.obs[0].copy_(obs[:, a])

    def collect(self, step, eps=None):
        """eps: per-agent standard-normal draws replacing Normal.sample() (tests that need the same draws on both sides)."""
        values, actions, logps, cost_preds = [], [], [], []
        for a, (actor, critic, cost_critic) in enumerate(self.nets):
            b = self.buffer[a]
            if eps is None:
                v, act, lp, cp = ma_get_actions(actor, critic, cost_critic, b.share_obs[step], b.obs[step])
            else:
                with torch.no_grad():
                    dist = ma_actor_dist(actor, b.obs[step])
                    act = dist.mean + dist.stddev * eps[a]
                    v, lp, cp = ma_critic_value(critic, b.share_obs[step]), dist.log_prob(act), ma_critic_value(cost_critic, b.share_obs[step])
            values.append(v), actions.append(act), logps.append(lp), cost_preds.append(cp)
        return torch.transpose(torch.stack(values), 1, 0), actions, logps, torch.transpose(torch.stack(cost_preds), 1, 0)

    def insert(self, obs, share_obs, rewards, costs, dones, values, actions, action_log_probs, cost_preds):
        dones_env = torch.all(dones, dim=1)
        masks = torch.ones(self.N, self.num_agents, 1)
        masks[dones_env] = 0.0
        active_masks = torch.ones(self.N, self.num_agents, 1)
        active_masks[dones] = 0.0
        active_masks[dones_env] = 1.0
        for a, b in enumerate(self.buffer):
            b.insert(share_obs[:, a], obs[:, a], actions[a], action_log_probs[a], values[:, a], rewards[:, a], masks[:, a], active_masks[:, a],
                     costs[:, a], cost_preds[:, a])

    def compute(self):
        with torch.no_grad():
            for (actor, critic, cost_critic), b, tr in zip(self.nets, self.buffer, self.trainer):
                b.compute_returns(ma_critic_value(critic, b.share_obs[-1]), tr.popart)
                b.compute_cost_returns(ma_critic_value(cost_critic, b.share_obs[-1]), tr.popart)

    def train(self, agent_order=None, perms=None):
        factor = torch.ones(self.T, self.N, 1)
        for agent_id in (torch.randperm(self.num_agents) if agent_order is None else agent_order):
            a = int(agent_id)
            b, tr, actor = self.buffer[a], self.trainer[a], self.nets[a][0]
            A = b.actions.shape[-1]
            b.factor.copy_(factor)
            flat_obs, flat_act = b.obs[:-1].reshape(-1, b.obs.shape[-1]), b.actions.reshape(-1, A)
            with torch.no_grad():
                old_lp = ma_actor_dist(actor, flat_obs).log_prob(flat_act)
            ma_trainer_train(tr, b, self.cfg["learning_iters"], None if perms is None else perms[a])
            with torch.no_grad():
                new_lp = ma_actor_dist(actor, flat_obs).log_prob(flat_act)
            factor = factor * torch.prod(torch.exp(new_lp - old_lp).reshape(self.T, self.N, A), dim=-1, keepdim=True)
            b.after_update()
