"""Per-phase cycle breakdown of one spo_update_kernel pass (needs tools/libspo_timers.so,
a build of libspo with -DSPO_PHASE_TIMERS for spo_update.cu)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))
from safepo import _lib as L  # noqa: E402

L.LIB_PATH = os.path.join(ROOT, "tools", "libspo_timers.so")
from safepo.single_agent._engine import PolicyGradientUpdate  # noqa: E402
from safepo.common.model import ActorVCritic  # noqa: E402

dev = torch.device("cuda:0")
S, D, A = int(os.environ.get("S", 1024000)), 60, 2
steps = min(int(os.environ.get("STEPS", 4000)), S // 64)
torch.manual_seed(0)
pol = ActorVCritic(D, A).to(dev)
data = {"obs": torch.randn(S, D, device=dev), "act": torch.randn(S, A, device=dev), "log_prob": torch.full((S,), -2.5, device=dev),
        "target_value_r": torch.randn(S, device=dev), "target_value_c": torch.randn(S, device=dev), "adv": torch.randn(S, device=dev)}
cfg = dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=1e9, batch_size=64, learning_iters=1, max_grad_norm=40.0)
upd = PolicyGradientUpdate(pol, cfg, L.LOSS_PPO_CLIP, epochs=100, host_rng=False, device=dev)
perm = torch.randperm(S, device=dev)[: steps * 64]
upd.run(data, perms=[perm])
torch.cuda.synchronize()
lib = L.lib()
buf = (C.c_ulonglong * 64)()
lib.spo_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
lib.spo_debug_phase_cycles(buf, 1)
upd.run(data, perms=[perm])
torch.cuda.synchronize()
lib.spo_debug_phase_cycles(buf, 1)
names = ["top (stage-in+barrier)", "fwd L1", "fwd L2", "out+loss", "small grads+dz2", "dW2+dh1", "dW1", "dp exchange", "reg+sumsq", "cluster sync",
         "adam", "(hidden GEMMs x2)", "(dW2 GEMM)", "(db2 sums)", "(dh1 GEMM)"]
for rank, net in enumerate(("actor", "reward critic", "cost critic")):
    row = [buf[rank * 16 + i] / steps for i in range(15)]
    print(f"{net:14s} total {sum(row):8.0f} cyc/step | " + " ".join(f"{n}={v:.0f}" for n, v in zip(names, row)))

# per-warp arrival times at every mark during step 50 (actor CTA): min / max over the 16 warps, relative to the step start
tr = (C.c_longlong * (4 * 16 * 24))()
lib.spo_debug_trace.argtypes = [C.c_void_p]
lib.spo_debug_trace(tr)
order = [22, 15, 0, 11, 1, 2, 17, 18, 3, 19, 4, 12, 13, 14, 5, 6, 7, 20, 8, 16, 9, 21, 10]
label = {22: "top: before sync", 15: "top: after sync", 0: "top done", 16: "next tile requested", 11: "hidden GEMM done (last = L2)", 1: "L1 epilogue+sync", 2: "L2 epilogue+sync",
         17: "out layer done", 18: "loss rows done", 3: "loss sync", 19: "small grads done", 4: "dz2 + sync", 12: "dW2 GEMM", 13: "db2 sums",
         14: "dh1 GEMM", 5: "dh1 epilogue + sync", 6: "dW1 + db1", 7: "dp exchange", 20: "reg+sumsq (before sync)", 8: "block reduce", 9: "cluster sync",
         21: "adam frags", 10: "adam small"}
for rank in (0, 1):
    t = [[tr[(rank * 16 + w) * 24 + m] for m in range(24)] for w in range(16)]
    t0 = min(t[w][22] for w in range(16))
    print(f"--- CTA {rank} step 50: arrival of the 16 warps at each mark (cycles since first warp reached the top)")
    prev = 0
    for m in order:
        col = [t[w][m] - t0 for w in range(16)]
        print(f"  {label[m]:32s} min {min(col):7d} max {max(col):7d}  (+{max(col) - prev:6d})  w0 {col[0]:7d} w1 {col[1]:7d} w8 {col[8]:7d} w15 {col[15]:7d}")
        prev = max(col)
