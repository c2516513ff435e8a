#!/usr/bin/env python
"""Parity check of the one-process-per-GPU path (NCCL transport).  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/sharded_check.py
Every rank builds the same synthetic texts, runs the sharded construction, the shards are
assembled on every rank and rank 0 compares them with the oracle (test infrastructure)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import oracle_lib  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = graft.load_package()
    from caps_sa_b200 import multi_gpu

    seng = multi_gpu.ShardedEngine(pkg, local_rank)
    s = pkg.synth
    cases = {
        "acgt_1M": (s.random_acgt(1_000_000, 101), 4),
        "genome_like_3M": (s.genome_like(3_000_000, seed=103, scale=0.004), 4),
        "bytes256_300k": (s.random_bytes(300_000, 104), 4),
        "fibonacci_200k": (s.fibonacci(200_000), 4),
        "periodic_unit1000_200k": (s.periodic_random_unit(200_000, 1000, seed=4), 4),
        "allA_50k": (np.full(50_000, ord("A"), dtype=np.uint8), 4),
        "acgt_u64_77777": (s.random_acgt(77_777, 102), 8),
        "tiny": (np.frombuffer(b"mississippi", dtype=np.uint8).copy(), 4),
    }
    failures = 0
    for name, (text, idx_bytes) in cases.items():
        dt = np.uint32 if idx_bytes == 4 else np.uint64
        n = len(text)
        sa = np.zeros(n, dtype=dt)
        lcp = np.zeros(n, dtype=dt)
        seng.construct(text, sa, lcp)
        off, cnt = seng.shard()
        layout = multi_gpu.shard_layout(off, cnt)
        multi_gpu.check_layout(layout, n)
        full_sa, full_lcp = multi_gpu.gather_result(sa, lcp, layout)
        if rank == 0:
            if n >= 16:
                want_sa, want_lcp = oracle_lib.port_sa_lcp(text, subproblems=16, idx_bytes=idx_bytes)
            else:
                a, b = oracle_lib.naive_sa_lcp(text)
                want_sa, want_lcp = a.astype(dt), b.astype(dt)
            ok = np.array_equal(full_sa, want_sa) and np.array_equal(full_lcp, want_lcp)
            failures += 0 if ok else 1
            st = seng.stats()
            print(f"{name}: {'OK' if ok else 'MISMATCH'} world={world} layout={layout} rounds={st['refine_rounds']} "
                  f"comm_bytes={st['comm_bytes']}", flush=True)
    flag = torch.tensor([failures], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_CHECK " + ("PASSED" if failures == 0 else f"FAILED ({failures})"), flush=True)
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
