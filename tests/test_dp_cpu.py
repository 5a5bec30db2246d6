"""World-size-2 gloo checks of the data-parallel host logic (no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Log:
    def __init__(self, vals, seen):
        self.epoch_dict = {"Metrics/EpCost": list(vals)}
        self.log_headers = ["Metrics/EpCost"] if seen else []


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "safe-policy-optimization_b200"))
    from safepo.common.dist import DataParallel
    dp = DataParallel()
    res = {}
    # Jc: global mean of the per-episode running means stored this epoch on all ranks
    res["jc"] = dp.mean_episode_cost(_Log([10.0, 20.0] if rank == 0 else [60.0], seen=True))
    # first-epoch quirk: 0.0 until the key was logged once anywhere
    res["jc_first"] = dp.mean_episode_cost(_Log([5.0], seen=False))
    # NaN when no episode finished anywhere after the first log (reference quirk A3)
    res["jc_nan"] = dp.mean_episode_cost(_Log([], seen=True))
    # advantage statistics: 4 fp64 sums add up
    stats = torch.tensor([1.0 + rank, 2.0, 3.0, 100.0], dtype=torch.float64)
    dp.all_reduce_sum(stats)
    res["stats"] = stats.tolist()
    w = torch.full((4,), float(rank + 1))
    dp.broadcast(w)
    res["bcast"] = w.tolist()
    # the trust-region exchange (flat gradient / FVP result / line-search means): mean over the ranks, identical everywhere
    v = torch.tensor([1.0, 2.0, 3.0]) * (rank + 1)
    dp.all_reduce_mean(v)
    res["mean"] = v.tolist()
    dp.advance(7)
    res["seq"] = dp.seq
    out[rank] = res
    dist.destroy_process_group()


def test_dp_host_logic_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for r in (0, 1):
        assert out[r]["jc"] == 30.0
        assert out[r]["jc_first"] == 0.0
        assert np.isnan(out[r]["jc_nan"])
        assert out[r]["stats"] == [3.0, 4.0, 6.0, 200.0]
        assert out[r]["bcast"] == [1.0] * 4
        assert out[r]["mean"] == [1.5, 3.0, 4.5]
        assert out[r]["seq"] == 7
