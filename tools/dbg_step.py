import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))
from oracle import spo_oracle as O
from safepo import _lib as L
from safepo.common.model import ActorVCritic
from safepo.single_agent._engine import PolicyGradientUpdate
dev = torch.device("cuda:0")
c = torch.load(os.path.join(ROOT, "tests/golden/update.pt"), weights_only=False)["update_chain"]["ppo"]
D, A = c["D"], c["A"]
pol = ActorVCritic(D, A).to(dev)
for n in ("actor", "reward_critic", "cost_critic"):
    getattr(pol, n).load_state_dict(c["init"][n])
opol = O.OraclePolicy(D, A); opol.load(c["init"])
opt = O.OracleOptim(opol)
lam = c["lam"]
adv = (c["data"]["adv_r"] - lam * c["data"]["adv_c"]) / (lam + 1)
cfg = dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=1e9, batch_size=64, learning_iters=1, max_grad_norm=40.0)
upd = PolicyGradientUpdate(pol, cfg, L.LOSS_PPO_CLIP, epochs=10**9, host_rng=False, device=dev)
data = {k: v.to(dev).contiguous() for k, v in c["data"].items()}
data["adv"] = adv.to(dev)
for s_ in range(4):
    idx = c["perms"][0][s_ * 64:(s_ + 1) * 64]
    b = {k: c["data"][k][idx] for k in ("obs", "act", "log_prob", "target_value_r", "target_value_c")}
    b["adv"] = adv[idx]
    # oracle forward pieces at the current (pre-update) weights
    with torch.no_grad():
        m_o, s_o = O.actor_mean_std(opol, b["obs"])
        lp_o = O.normal_log_prob(b["act"], m_o, s_o).sum(-1)
        vr_o = O.critic_value(opol, "reward_critic", b["obs"])
    # my forward (rollout FFMA kernel) at my current weights
    a_m, lp_m, vr_m, vc_m = pol.step(b["obs"].to(dev), deterministic=True)
    lr_, lc_, lp_ = O.minibatch_step(opol, opt, b, "ppo")
    res = upd.run(data, perms=[idx], refresh_old=(s_ == 0))
    torch.cuda.synchronize()
    werr = max(float((getattr(pol, n).state_dict()[k].cpu() - v.detach()).abs().max()) for n in ("actor", "reward_critic", "cost_critic") for k, v in opol.nets[n].items())
    print(f"step {s_}: loss_pi mine {res['loss_pi']:.8f} ref {lp_:.8f} diff {res['loss_pi'] - lp_:+.2e} | loss_r diff {res['loss_r'] - lr_:+.2e} | "
          f"fwd(mean) ffma-vs-oracle {float((a_m.cpu() - m_o).abs().max()):.2e} | max weight err after step {werr:.2e}")
