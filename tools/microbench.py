"""Quick device-side timings of the individual kernels (CUDA events).  Not the bench."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))
from safepo import _lib as L  # noqa: E402
from safepo.common.buffer import VectorizedOnPolicyBuffer  # noqa: E402
from safepo.common.model import ActorVCritic  # noqa: E402
from safepo.single_agent._engine import PolicyGradientUpdate, read_ctrl  # noqa: E402


class Sp:
    def __init__(s, d):
        s.shape = (d,)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    N, T, D, A = 1024, int(os.environ.get("T", 1000)), 60, 2
    torch.manual_seed(0)
    pol = ActorVCritic(D, A).to(dev)
    buf = VectorizedOnPolicyBuffer(Sp(D), Sp(A), size=T, device=dev, num_envs=N)
    obs = torch.randn(N, D, device=dev)
    print("policy_step (N=1024, fused store) us:", timeit(lambda: pol.step(obs, store=(buf.struct, 3)), 200))
    print("critic_values us:", timeit(lambda: pol.values(obs), 200))
    S = N * T
    buf.data["obs"].normal_(); buf.data["reward"].normal_(); buf.data["cost"].uniform_(); buf.data["value_r"].normal_()
    buf.data["value_c"].normal_(); buf.seg_end[:, -1] = 1; buf.seg_end[:, T // 2] = 1
    for mode in ("scan", "exact"):
        buf.gae_mode = {"scan": 0, "exact": 1}[mode]
        us = timeit(buf.compute_gae, 20)
        print(f"gae_dual[{mode}] us: {us:.1f}  -> {S * 33 / us / 1e3:.1f} GB/s algorithmic")
    us = timeit(lambda: buf.finalize(0.3), 20)
    print(f"adv stats+apply us: {us:.1f} -> {S * 20 / us / 1e3:.1f} GB/s algorithmic")
    data = buf.get(0.3)
    data["act"].normal_(); data["log_prob"].fill_(-2.5)
    cfg = dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=1e9, batch_size=64, learning_iters=1, max_grad_norm=40.0)
    upd = PolicyGradientUpdate(pol, cfg, L.LOSS_PPO_CLIP, epochs=100, host_rng=False, device=dev)
    torch.cuda.synchronize()
    t0 = time.time(); res = upd.run(data); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"pg_update one pass: {dt*1e3:.1f} ms, steps={res['steps']}, us/step={dt*1e6/res['steps']:.2f}", res)
    t0 = time.time(); res = upd.run(data); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"pg_update one pass (2nd): {dt*1e3:.1f} ms, us/step={dt*1e6/res['steps']:.2f}")
    us = timeit(lambda: pol.actor_mean(data["obs"]), 10)
    print(f"actor_forward full batch S={S}: {us:.1f} us -> {S*240/us/1e3:.1f} GB/s, {S*16128/us/1e6:.2f} TFLOP/s")


if __name__ == "__main__":
    main()
