"""The N > 1 host logic on CPU: contiguous env slices per rank and the (optional) episode-return
statistics all-reduce -- the only collective of the design, off the step path -- with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from smart_nanogrid_gym_b200.sharding import ReturnStats, shard_range  # noqa: E402


def test_shard_ranges_partition_the_batch():
    for total in (1048576, 1000, 33, 7):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            # every shard but the last starts and ends on a 32-env state block
            assert all(s[0] % 32 == 0 for s in spans if s[1] > s[0])
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 32 or total < 32 * world


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 1000
    lo, hi = shard_range(total, world, rank)
    rng = np.random.default_rng(0)
    returns = torch.tensor(rng.normal(-400, 90, size=total)[lo:hi])     # this rank's slice of the same global vector
    st = ReturnStats.from_returns(returns).all_reduce()
    q.put((rank, st.count, st.mean, st.std, st.min, st.max))
    dist.destroy_process_group()


def test_return_stats_all_reduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.random.default_rng(0).normal(-400, 90, size=1000)
    for _, count, mean, std, mn, mx in out:
        assert count == 1000
        assert np.isclose(mean, ref.mean()) and np.isclose(std, ref.std()) and np.isclose(mn, ref.min()) and np.isclose(mx, ref.max())


def test_return_stats_single_process():
    x = torch.tensor([-1.0, -2.0, -6.0])
    st = ReturnStats.from_returns(x).all_reduce()      # no process group: a no-op
    assert st.count == 3 and np.isclose(st.mean, -3.0) and st.min == -6.0 and st.max == -1.0
