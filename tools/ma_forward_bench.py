"""Time MultiAgentNets.get_actions at BASELINE config 5's shape (obs 398, act 20, hidden 512, layer_N 2; N rows)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))
from safepo.common.ma_model import MultiAgentNets  # noqa: E402

dev = torch.device("cuda:0")
N, D, A, H = int(os.environ.get("N", 8192)), 398, 20, 512
g = torch.Generator().manual_seed(0)


def make(head):
    st = {"base.feature_norm.weight": torch.ones(D), "base.feature_norm.bias": torch.zeros(D)}
    dims = [D, H, H, H]
    for li, name in enumerate(("fc1", "fc2.0", "fc2.1")):
        st[f"base.mlp.{name}.0.weight"] = torch.randn(H, dims[li], generator=g) / dims[li] ** 0.5
        st[f"base.mlp.{name}.0.bias"] = torch.zeros(H)
        st[f"base.mlp.{name}.2.weight"] = torch.ones(H)
        st[f"base.mlp.{name}.2.bias"] = torch.zeros(H)
    if head == "actor":
        st["act.action_out.log_std"] = torch.ones(A)
        st["act.action_out.fc_mean.weight"] = torch.randn(A, H, generator=g) * 0.01
        st["act.action_out.fc_mean.bias"] = torch.zeros(A)
    else:
        st["v_out.weight"] = torch.randn(1, H, generator=g) * 0.1
        st["v_out.bias"] = torch.zeros(1)
    return st


nets = MultiAgentNets(make("actor"), make("critic"), make("critic"), dev)
obs, cent = torch.randn(N, D, device=dev), torch.randn(N, D, device=dev)
for _ in range(3):
    nets.get_actions(cent, obs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(int(os.environ.get('REPS', 10))):
    nets.get_actions(cent, obs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / int(os.environ.get('REPS', 10))
flops = 3 * N * 2 * (D * H + 2 * H * H) + N * 2 * H * (A + 2)
print(f"get_actions at N={N}: {ms:.3f} ms per call (3 nets x 3 layers + heads), {flops / ms / 1e9:.1f} TFLOP/s fp32")

# ---- one MAPPO_L_Trainer.ppo_update on the same shape (config.yaml: whole batch per update, num_mini_batch 1) ----
from safepo.common.ma_model import MultiAgentTrainer  # noqa: E402

cfg = dict(actor_lr=5e-4, critic_lr=5e-4, opti_eps=1e-5, weight_decay=0.0, clip_param=0.2, huber_delta=10.0, entropy_coef=0.01,
           max_grad_norm=10.0, cost_limit=25.0, gamma=0.96, lagrangian_coef_rate=1e-5, value_loss_coef=1.0, lamda_lagr=0.78)
tr = MultiAgentTrainer(nets, cfg)
v, act, logp, cv = nets.get_actions(cent, obs)
sample = dict(share_obs=cent, obs=obs, actions=act, value_preds=v + 0.1 * torch.randn_like(v), returns=torch.randn_like(v) * 4 + 1,
              old_action_log_probs=logp + 0.05 * torch.randn_like(logp), adv_targ=torch.randn_like(v), factor=torch.rand_like(v) + 0.5,
              cost_preds=cv + 0.1 * torch.randn_like(v), cost_returns=torch.randn_like(v).abs() * 30, cost_adv_targ=torch.randn_like(v),
              aver_episode_costs=torch.rand_like(v) * 60)
for _ in range(2):
    out = tr.ppo_update(sample)
torch.cuda.synchronize()
e0.record()
for _ in range(int(os.environ.get('UREPS', 5))):
    out = tr.ppo_update(sample)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / int(os.environ.get('UREPS', 5))
# forward + dW + dx per layer product: 3 x the forward flops (the first layer has no dx into the observations' LayerNorm input... it does: d xn)
uflops = 3 * flops
print(f"ppo_update at N={N}: {ms:.3f} ms per update (3 nets: forward, backward, clip + Adam), {uflops / ms / 1e9:.1f} TFLOP/s fp32; "
      f"value_loss {float(out[0]):.4f} policy_loss {float(out[2]):.4f} actor grad norm {float(out[4]):.4f}")
