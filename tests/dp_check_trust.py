"""2+ GPU check of the data-parallel trust-region path (run under torchrun):
   every rank holds 1/G of the samples; the flat gradients, every Fisher-vector product of the CG solve, the line-search
   means and the critic regression's gradients are averaged / summed across the ranks (exchange 3').
   (1) g, the CG solution x, F x and the line-search means equal the single-GPU values on the union of the shards;
   (2) a full CPO update (run_cpo + critic regression) leaves bit-identical weights on all ranks, close to the single-GPU run.
Prints DP_TRUST PASS / FAIL on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))

from safepo.common.dist import DataParallel  # noqa: E402
from safepo.common.model import ActorVCritic  # noqa: E402
from safepo.single_agent._engine import CriticRegression, TrustRegionUpdate  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    dp = DataParallel()
    c = torch.load(os.path.join(ROOT, "tests", "golden", "trust.pt"), weights_only=False)["trust"]
    D, A = c["D"], c["A"]
    S = (c["data"]["obs"].shape[0] // (world * 128)) * world * 128
    full = {k: v[:S] for k, v in c["data"].items()}
    shard = {k: v[rank * (S // world):(rank + 1) * (S // world)].to(dev).contiguous() for k, v in full.items()}
    cfg = dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=0.01, batch_size=128 // world, learning_iters=2, max_grad_norm=40.0)

    def make():
        pol = ActorVCritic(D, A).to(dev)
        for n in ("actor", "reward_critic", "cost_critic"):
            getattr(pol, n).load_state_dict(c["state"][n])
        return pol

    pol = make()
    tr = TrustRegionUpdate(pol, cfg, dev, dp=dp)
    tr._grad(shard, shard["adv_r"], tr.g)
    tr._old_dist(shard)
    tr._cg(shard, tr.g, tr.x)
    tr._fvp(shard, tr.x, tr.Fx)
    g_dp, x_dp, Fx_dp = tr.g.clone(), tr.x.clone(), tr.Fx.clone()
    ev_dp = tr._eval(shard, shard["adv_r"], shard["adv_c"]).clone()
    # full update on the shards
    res = tr.run_cpo(shard, ep_costs=3.0)
    crit = CriticRegression(pol, cfg, host_rng=False, device=dev, dp=dp)
    g = torch.Generator().manual_seed(5)
    perms = [torch.randperm(S // world, generator=g) for _ in range(cfg["learning_iters"])]
    crit.run(shard, perms=perms)
    torch.cuda.synchronize()
    flat = pol.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], t) for t in gathered)
    ok = True
    if rank == 0:
        pol1 = make()
        t1 = TrustRegionUpdate(pol1, cfg, dev)
        data1 = {k: v.to(dev).contiguous() for k, v in full.items()}
        t1._grad(data1, data1["adv_r"], t1.g)
        t1._old_dist(data1)
        t1._cg(data1, t1.g, t1.x)
        t1._fvp(data1, t1.x, t1.Fx)
        ev1 = t1._eval(data1, data1["adv_r"], data1["adv_c"])

        def rel(a, b):
            return float((a - b).norm() / b.norm())
        eg, ex, ef = rel(g_dp, t1.g), rel(x_dp, t1.x), rel(Fx_dp, t1.Fx)
        ee = float((ev_dp - ev1).abs().max())
        res1 = t1.run_cpo(data1, ep_costs=3.0)
        dtheta = float((pol1.flat[: pol1.n_actor] - flat[: pol.n_actor]).abs().max())
        print(f"world={world} identical_replicas={identical} rel err g {eg:.2e} x {ex:.2e} Fx {ef:.2e} evals {ee:.2e}; "
              f"accepted step {res.get('Misc/AcceptanceStep')} vs {res1.get('Misc/AcceptanceStep')}, actor max|dtheta| dp-vs-single {dtheta:.2e}")
        ok = identical and eg < 1e-5 and ef < 1e-5 and ex < 2e-3 and ee < 1e-6 and dtheta < 5e-4 and \
            res.get("Misc/AcceptanceStep") == res1.get("Misc/AcceptanceStep")
        print("DP_TRUST", "PASS" if ok else "FAIL")
    dp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
