"""2+ GPU check of the in-kernel data-parallel gradient sum (run under torchrun):
   (1) G ranks x batch 64/G on disjoint shards == 1 rank x batch 64 on the union (mode B of
       SURVEY 8e: exact single-process semantics up to summation order);
   (2) replicas stay bit-identical across ranks;
   (3) us per minibatch step with the exchange in the loop.
Prints DP_CHECK PASS / FAIL on rank 0."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))

from safepo import _lib as L  # noqa: E402
from safepo.common.dist import DataParallel  # noqa: E402
from safepo.common.model import ActorVCritic  # noqa: E402
from safepo.single_agent._engine import PolicyGradientUpdate  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    dp = DataParallel()
    c = torch.load(os.path.join(ROOT, "tests", "golden", "update.pt"), weights_only=False)["update_chain"]["ppo"]
    D, A, B = c["D"], c["A"], c["batch"]
    S = c["data"]["obs"].shape[0]
    S = (S // (world * B)) * world * B  # equal shards, whole batches
    half = S // world
    bl = B // world
    lam = c["lam"]
    full = {k: v[:S] for k, v in c["data"].items()}
    full["adv"] = (full["adv_r"] - lam * full["adv_c"]) / (lam + 1)

    def make(state):
        pol = ActorVCritic(D, A).to(dev)
        for n in ("actor", "reward_critic", "cost_critic"):
            getattr(pol, n).load_state_dict(state[n])
        return pol

    cfg = dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=1e9, batch_size=bl, learning_iters=1, max_grad_norm=40.0)
    g = torch.Generator().manual_seed(7)
    perms = [torch.randperm(half, generator=g) for _ in range(world)]
    # --- DP run: each rank its shard, batch B/world
    pol = make(c["init"])
    upd = PolicyGradientUpdate(pol, cfg, L.LOSS_PPO_CLIP, epochs=10**9, host_rng=False, device=dev, dp=dp)
    shard = {k: v[rank * half:(rank + 1) * half].to(dev).contiguous() for k, v in full.items()}
    passes = 3
    res_dp = [upd.run(shard, perms=[perms[rank]], refresh_old=(p == 0)) for p in range(passes)]
    torch.cuda.synchronize()
    flat = pol.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    identical = all(torch.equal(gathered[0], x) for x in gathered)
    ok = True
    if rank == 0:
        # --- single-GPU run on the union with the interleaved order
        pol1 = make(c["init"])
        cfg1 = dict(cfg, batch_size=B)
        upd1 = PolicyGradientUpdate(pol1, cfg1, L.LOSS_PPO_CLIP, epochs=10**9, host_rng=False, device=dev)
        steps = half // bl
        order = torch.cat([torch.cat([r * half + perms[r][s * bl:(s + 1) * bl] for r in range(world)]) for s in range(steps)])
        data1 = {k: v.to(dev).contiguous() for k, v in full.items()}
        res_1 = [upd1.run(data1, perms=[order], refresh_old=(p == 0)) for p in range(passes)]
        torch.cuda.synchronize()
        dw = float((pol1.flat - flat).abs().max())
        print(f"world={world} identical_replicas={identical} max|dtheta| dp-vs-single={dw:.3e} "
              f"kl dp={res_dp[-1]['kl']:.6e} single={res_1[-1]['kl']:.6e} stop={res_dp[-1]['stop_iter']}")
        ok = identical and dw < 5e-5 and abs(res_dp[-1]["kl"] - res_1[-1]["kl"]) <= 1e-4 * abs(res_1[-1]["kl"]) + 1e-7
    dist.barrier()
    # --- timing at scale
    S2 = 64 * 4000
    torch.manual_seed(rank)
    big = {"obs": torch.randn(S2, D, device=dev), "act": torch.randn(S2, A, device=dev), "log_prob": torch.full((S2,), -2.5, device=dev),
           "target_value_r": torch.randn(S2, device=dev), "target_value_c": torch.randn(S2, device=dev), "adv": torch.randn(S2, device=dev)}
    cfg2 = dict(cfg, batch_size=64)
    upd.cfg = cfg2
    for _ in range(2):
        upd.run(big)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.time()
    r = upd.run(big)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(f"dp world={world}: {dt * 1e6 / r['steps']:.2f} us per minibatch step (batch 64/rank, {r['steps']} steps), stop flag {r['stop_iter']}")
        print("DP_CHECK", "PASS" if ok else "FAIL")
    dp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
