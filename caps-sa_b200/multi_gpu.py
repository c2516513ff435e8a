"""Host side of the sharded construction with one process per GPU (torchrun).

``torch.distributed`` is plumbing only: rendezvous, the broadcast of the NCCL id that the
library's own communicator is built from, barriers and the max-over-ranks reduction of
timings.  Every data-path exchange (samples, count matrix, the (key, suffix) all-to-all, the
rank exchange of each refinement round, the LCP round trip) happens inside the library over
that communicator (csrc/comm.cu, csrc/sharded_build.cu).

    eng = ShardedEngine(pkg, local_rank)          # collective: joins the communicator
    eng.construct(text, sa_out, lcp_out)          # collective: this rank's shard lands in
                                                  #   sa_out[offset:offset+count]
    offset, count = eng.shard()
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


# ---- pure host logic (covered by the gloo tests on CPU) -----------------------------------
def slice_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Text positions whose suffixes rank `rank` sorts first: n // world each, the last rank
    takes the remainder (reference src/Suffix_Array.cpp:166,172; csrc/sharded_build.cu SliceMap)."""
    s = n // world
    lo = rank * s
    return lo, (n if rank == world - 1 else lo + s)


def _plumbing_device(group=None) -> torch.device:
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")


def broadcast_comm_id(make_id, group=None) -> bytes:
    """Rank 0 calls make_id() (128 bytes); every rank returns the same bytes."""
    dev = _plumbing_device(group)
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if dist.get_rank(group) == 0:
        raw = bytes(make_id())
        assert len(raw) == 128
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    return t.cpu().numpy().tobytes()


def max_over_ranks(value: float, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=_plumbing_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=_plumbing_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def shard_layout(offset: int, count: int, group=None) -> list[tuple[int, int]]:
    """(offset, count) of every rank's shard, in rank order."""
    world = dist.get_world_size(group)
    mine = torch.tensor([offset, count], dtype=torch.int64, device=_plumbing_device(group))
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [(int(t[0]), int(t[1])) for t in out]


def check_layout(layout, n: int) -> None:
    """The non-empty shards tile [0, n) in rank order (rank r owns bucket r)."""
    at = 0
    for off, cnt in layout:
        if cnt:
            if off != at:
                raise ValueError(f"shard layout is not contiguous: {layout}")
            at += cnt
    if at != n:
        raise ValueError(f"shards cover {at} of {n} positions: {layout}")


def gather_result(sa_full: np.ndarray, lcp_full: np.ndarray, layout, group=None):
    """Test/verification helper: every rank holds its shard inside full-length arrays (as
    ShardedEngine.construct leaves them); returns the assembled (SA, LCP) on every rank."""
    rank = dist.get_rank(group)
    off, cnt = layout[rank]
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (sa_full[off:off + cnt].copy(), lcp_full[off:off + cnt].copy()), group=group)
    sa = np.empty_like(sa_full)
    lcp = np.empty_like(lcp_full)
    for (o, c), (s, l) in zip(layout, parts):
        sa[o:o + c] = s
        lcp[o:o + c] = l
    return sa, lcp


# ---- engine wrapper ----------------------------------------------------------------------
class ShardedEngine:
    """One rank of the sharded construction (needs a CUDA device; no CPU fallback)."""

    def __init__(self, pkg, device: int, group=None):
        self.pkg = pkg
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.engine = pkg.Engine(device)
        if self.world > 1:
            comm_id = broadcast_comm_id(pkg.comm_unique_id, group)
        else:
            comm_id = bytes(128)
        self.engine.comm_init(comm_id, self.rank, self.world)

    def construct_device(self, d_text: int, n: int, idx_bytes: int = 4, stream: int = 0) -> None:
        self.engine.construct_sharded_device(d_text, n, idx_bytes, stream)

    def construct(self, text: np.ndarray, sa_out: np.ndarray, lcp_out: np.ndarray) -> None:
        self.engine.construct_sharded(text, sa_out, lcp_out)

    def stats(self) -> dict:
        return self.engine.stats()

    def shard(self) -> tuple[int, int]:
        st = self.engine.stats()
        return st["shard_offset"], st["shard_count"]


def bind_to_gpu_cpus(device_index: int):
    """Runs this rank on the CPUs next to its GPU (NVML's affinity mask) so that the host pages it
    first touches — its shard of the shared SA/LCP arrays, pinned right after — land on that GPU's
    NUMA node and the ranks' device-to-host copies do not all converge on one memory controller.
    Returns the CPU list, or None when NVML or the scheduler call is unavailable (nothing changes)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - strictly an optimisation
        return None


# ---- bench.py, N > 1 -----------------------------------------------------------------------
def bench_main(args, spec, n, pkg, make_text, clock_sampler_cls, load_peaks, kernel_times_cls):
    """bench.py's multi-rank arm: same JSON contract, time = max over ranks, value = n / time.
    The text is fixed (strong scaling): every rank generates the same synthetic text.
    Returns (json line, (text, SA, LCP) host arrays) on rank 0, None on the other ranks."""
    rank = dist.get_rank()
    world = dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpus = bind_to_gpu_cpus(local_rank)

    # Host buffers are shared by the ranks of the node, as the threads of the C++ class share its
    # SA_/LCP_ arrays: one text, one SA and one LCP array in /dev/shm, every
    # rank maps them and pins only what it touches (its piece of the text, its shard of SA/LCP).
    tag = f"capsb_{os.environ.get('MASTER_PORT', '0')}_{n}"
    paths = {k: f"/dev/shm/{tag}_{k}" for k in ("text", "sa", "lcp")}
    if local_rank == 0:
        text_np = make_text(pkg, spec, n)
        shared_text = np.memmap(paths["text"], dtype=np.uint8, mode="w+", shape=(n,))
        shared_text[:] = text_np
        del text_np
        for k in ("sa", "lcp"):
            np.memmap(paths[k], dtype=np.uint32, mode="w+", shape=(n,)).flush()
    dist.barrier()
    text_host = np.memmap(paths["text"], dtype=np.uint8, mode="r+", shape=(n,))
    sa_host = np.memmap(paths["sa"], dtype=np.uint32, mode="r+", shape=(n,))
    lcp_host = np.memmap(paths["lcp"], dtype=np.uint32, mode="r+", shape=(n,))
    dist.barrier()
    if local_rank == 0:  # the mappings keep the memory alive; nothing is left behind if a rank dies
        for path in paths.values():
            os.unlink(path)

    def pin(arr, lo, hi):
        """cudaHostRegister of arr[lo:hi] (page-aligned outwards); False if the driver refuses."""
        item = arr.dtype.itemsize
        base = arr.ctypes.data
        start = (base + lo * item) // 4096 * 4096
        stop = min(base + arr.nbytes, -(-(base + hi * item) // 4096) * 4096)
        if stop <= start:
            return True
        return int(torch.cuda.cudart().cudaHostRegister(start, stop - start, 0)) == 0

    seng = ShardedEngine(pkg, local_rank)
    stream = torch.cuda.current_stream()
    seng.engine.set_stream(stream.cuda_stream)

    # device-resident copy of the text for the `value` leg (before any part of the mapping is
    # pinned: one copy must not span pinned and pageable pages)
    d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_text.copy_(torch.from_numpy(text_host))
    torch.cuda.synchronize()
    piece = (-(-n // world) + 15) // 16 * 16  # csrc/capi.cu stage_text_sharded
    pinned_ok = pin(text_host, min(n, rank * piece), min(n, (rank + 1) * piece))

    def device_step():
        seng.construct_device(d_text.data_ptr(), n, 4, stream.cuda_stream)

    def timed(fn, steps):
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        dist.barrier()
        # the library synchronises with the host inside a step, so the wall clock can only be
        # longer than the event time; report the larger, max over ranks
        return max_over_ranks(max(ev0.elapsed_time(ev1), wall_ms)) / steps

    for _ in range(args.warmup):
        device_step()
    seng.engine.set_kernel_timing(True)
    sampler = clock_sampler_cls(local_rank)
    if rank == 0:
        sampler.start()
    launches = 0.0
    comm_bytes = 0.0
    kernel_times = kernel_times_cls()
    stage_ms = {}

    def counted_step():
        nonlocal launches, comm_bytes
        device_step()
        st = seng.stats()
        launches += st["kernel_launches"]
        comm_bytes += st["comm_bytes"]
        kernel_times.add(st)
        for k in ("ms_pack", "ms_sort", "ms_partition", "ms_merge", "ms_refine", "ms_deep_lcp", "ms_total"):
            stage_ms[k] = stage_ms.get(k, 0.0) + st[k] / args.steps

    dev_ms = timed(counted_step, args.steps)
    seng.engine.set_kernel_timing(False)
    offset, count = seng.shard()
    layout = shard_layout(offset, count)
    check_layout(layout, n)

    # end to end: every rank uploads its piece of the host text (pinned), the pieces are
    # all-gathered over NVLink, construct, D2H of the rank's shard into the shared arrays
    pinned_ok = pin(sa_host, offset, offset + count) and pin(lcp_host, offset, offset + count) and pinned_ok

    def e2e_step():
        seng.construct(text_host, sa_host, lcp_host)

    for _ in range(max(1, args.warmup - 2)):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    st = seng.stats()  # of the last end-to-end step
    e2e_parts = {k: max_over_ranks(st[k]) for k in ("ms_h2d", "ms_total", "ms_d2h")}

    total_launches = sum_over_ranks(launches)
    total_comm = sum_over_ranks(comm_bytes)
    largest = max(c for _, c in layout)
    if rank != 0:
        return None
    peak, peak_kind = load_peaks()
    roofline, kernel_table = kernel_times.report(peak, peak_kind, args.steps, dev_ms)
    if roofline is not None:
        roofline["kernel"] += " (rank 0)"
    line = {
        "metric": "sa_lcp_suffixes_per_sec", "value": n / (dev_ms / 1e3), "unit": "suffixes/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['desc']}", "n": n, "idx_bytes": 4,
                   "parallelism": f"samplesort over {world} ranks: replicated text, slice sort, NCCL all-to-all, bucket merge",
                   "l2": "inputs >> 126 MB L2, no explicit flush", "largest_shard": largest,
                   "shard_imbalance": largest / (n / world)},
        "e2e": {"value": n / (e2e_ms / 1e3), "unit": "suffixes/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": n, "d2h_bytes_per_step": 2 * 4 * n, "host_buffers_pinned": bool(pinned_ok),
                "max_over_ranks_ms": {k: round(v, 3) for k, v in e2e_parts.items()},
                "rank0_cpu_affinity": (f"{len(cpus)} CPUs next to the GPU (NVML)" if cpus else "unchanged")},
        "gpu_launches": int(total_launches),
        "nvlink_bytes_per_step": total_comm / args.steps,
        "stage_ms_rank0": {k: round(v, 3) for k, v in stage_ms.items()},
        "roofline": roofline, "kernels_rank0": kernel_table,
        "cpu_baseline": None, "clocks": clocks,
    }
    # rank 0 verifies the arrays all ranks wrote into the shared host buffers and prints the line
    # once the process group is gone (bench.py), so no rank waits in a collective meanwhile
    return line, (text_host, sa_host, lcp_host)
