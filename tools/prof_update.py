"""Short update pass for ncu captures: STEPS minibatch steps of PPO-Lag at S=1,024,000."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "safe-policy-optimization_b200"))
from safepo import _lib as L  # noqa: E402

if os.environ.get("LIB"):      # A/B runs: another build of the library (e.g. tools/libspo_head.so)
    L.LIB_PATH = os.path.join(ROOT, os.environ["LIB"])
from safepo.single_agent._engine import PolicyGradientUpdate  # noqa: E402
from safepo.common.model import ActorVCritic  # noqa: E402

dev = torch.device("cuda:0")
S, D, A = int(os.environ.get("S", 1024000)), 60, 2
steps = int(os.environ.get("STEPS", 300))
torch.manual_seed(0)
pol = ActorVCritic(D, A).to(dev)
data = {"obs": torch.randn(S, D, device=dev), "act": torch.randn(S, A, device=dev), "log_prob": torch.full((S,), -2.5, device=dev),
        "target_value_r": torch.randn(S, device=dev), "target_value_c": torch.randn(S, device=dev), "adv": torch.randn(S, device=dev)}
cfg = dict(hidden_sizes=[64, 64], gamma=0.99, target_kl=1e9, batch_size=64, learning_iters=1, max_grad_norm=40.0)
upd = PolicyGradientUpdate(pol, cfg, L.LOSS_PPO_CLIP, epochs=100, host_rng=False, device=dev)
perm = torch.randperm(S, device=dev)[: steps * 64]
for _ in range(3):
    res = upd.run(data, perms=[perm])
torch.cuda.synchronize()
print(res)
if os.environ.get("TIME"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        upd.run(data, perms=[perm])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"update kernel: {best * 1e3 / steps:.2f} us per minibatch step ({steps} steps, best of 5)")
if os.environ.get("FILLER"):
    # experiment: the same timing while a filler kernel keeps the other SMs busy (tools/filler.cu)
    import ctypes as C
    fl = C.CDLL(os.path.join(ROOT, "tools", "libfiller.so"))
    fl.filler_launch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
    assert fl.filler_prepare() == 0
    side, flagst = torch.cuda.Stream(), torch.cuda.Stream()
    sink = torch.zeros(1, device=dev)
    for mode in (0, 1, 2):
        for ctas in (64, 132):
            for threads in (32, 256):
                stop = torch.zeros(1, dtype=torch.int32, device=dev)
                torch.cuda.synchronize()
                rc = fl.filler_launch(ctas, threads, mode, stop.data_ptr(), int(0.25 * 1.9e9), sink.data_ptr(), side.cuda_stream)
                assert rc == 0, rc
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e9
                for _ in range(3):
                    e0.record()
                    upd.run(data, perms=[perm])
                    e1.record()
                    e1.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                with torch.cuda.stream(flagst):
                    stop.fill_(1)
                torch.cuda.synchronize()
                print(f"filler mode {mode} ctas {ctas} threads {threads}: {best * 1e3 / steps:.2f} us per step")
