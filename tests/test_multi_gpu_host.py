"""CPU tests (gloo, world_size 2) of the host side of the sharded path: the rendezvous that
hands every rank the same NCCL id, the shard-layout bookkeeping, result assembly and the
max-over-ranks timing reduction.  The exchange steps themselves run inside the CUDA library
and are covered by tests/test_gpu_sharded.py on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as graft
    import oracle_lib

    pkg = graft.load_package()
    from caps_sa_b200 import multi_gpu

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # 1. every rank ends up with rank 0's id (the real id generator: NCCL loads without a GPU)
        comm_id = multi_gpu.broadcast_comm_id(pkg.comm_unique_id)
        ids = [None] * world
        dist.all_gather_object(ids, comm_id)
        assert len(comm_id) == 128 and all(i == ids[0] for i in ids)
        assert any(b != 0 for b in comm_id)

        # 2. slices tile the text; the last rank takes the remainder
        n = 100_003
        lo, hi = multi_gpu.slice_bounds(n, world, rank)
        spans = [None] * world
        dist.all_gather_object(spans, (lo, hi))
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))

        # 3. layout exchange + assembly: each rank holds one bucket of the oracle's SA/LCP
        text = pkg.synth.random_acgt(5000, 11)
        sa, lcp = oracle_lib.port_sa_lcp(text, subproblems=8)
        cut = [0, 1777, len(text)] if world == 2 else np.linspace(0, len(text), world + 1).astype(int).tolist()
        off, cnt = cut[rank], cut[rank + 1] - cut[rank]
        mine_sa = np.zeros_like(sa)
        mine_lcp = np.zeros_like(lcp)
        mine_sa[off:off + cnt] = sa[off:off + cnt]
        mine_lcp[off:off + cnt] = lcp[off:off + cnt]
        layout = multi_gpu.shard_layout(off, cnt)
        assert layout == [(cut[r], cut[r + 1] - cut[r]) for r in range(world)]
        multi_gpu.check_layout(layout, len(text))
        got_sa, got_lcp = multi_gpu.gather_result(mine_sa, mine_lcp, layout)
        assert np.array_equal(got_sa, sa) and np.array_equal(got_lcp, lcp)
        assert oracle_lib.check_sa_lcp(text, got_sa, got_lcp) == (0, 0)

        # 4. a gap or a short cover is rejected
        for bad in ([(0, 10), (11, 5)], [(0, 10), (10, 4)]):
            with pytest.raises(ValueError):
                multi_gpu.check_layout(bad, 15)
        multi_gpu.check_layout([(0, 15), (15, 0)], 15)  # an empty bucket is fine

        # 5. timings reduce to the slowest rank, counters to the sum
        assert multi_gpu.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        assert multi_gpu.sum_over_ranks(float(rank + 1)) == world * (world + 1) / 2

        # 6. without a CUDA device the engine refuses to start (no CPU fallback)
        if pkg.lib().caps_sa_gpu_device_count() == 0:
            with pytest.raises(pkg.CapsSaError):
                multi_gpu.ShardedEngine(pkg, 0)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_host_side_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_bind_to_gpu_cpus_is_a_noop_without_nvml():
    """bench.py's NUMA binding is strictly an optimisation: without a GPU / NVML it reports None and
    leaves the process affinity alone."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    graft.load_package()
    from caps_sa_b200 import multi_gpu

    before = os.sched_getaffinity(0)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the binding may legitimately change the affinity")
    assert multi_gpu.bind_to_gpu_cpus(0) is None
    assert os.sched_getaffinity(0) == before


def test_slice_bounds_cover_the_text():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    graft.load_package()
    from caps_sa_b200 import multi_gpu

    for n, world in [(0, 3), (7, 8), (100, 1), (3_100_000_000, 8), (12_000_000_001, 7)]:
        at = 0
        for r in range(world):
            lo, hi = multi_gpu.slice_bounds(n, world, r)
            assert lo == at and hi >= lo
            at = hi
        assert at == n
