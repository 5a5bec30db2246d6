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
