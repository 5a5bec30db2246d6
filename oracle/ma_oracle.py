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
