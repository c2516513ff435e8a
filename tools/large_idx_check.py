#!/usr/bin/env python
"""LARGE_IDX in anger: a text of 2^32 + 1000 symbols needs 64-bit indices (reference src/main.cpp:76-87).
Builds SA + LCP with the sharded construction over the visible GPUs (one host thread per rank) and runs
the independent checker on the full result.   usage: large_idx_check.py [n] [ranks]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import oracle_lib  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else (1 << 32) + 1000
pkg = graft.load_package()
visible = pkg.lib().caps_sa_gpu_device_count()
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else visible
t0 = time.time()
text = pkg.synth.random_acgt_chunked(n, 5)
# a few long repeats so that pair chains, text rounds and rank rounds all see indices beyond 2^32
rng = np.random.default_rng(55)
for length, copies in ((200_000, 3), (5_000, 40), (171 * 300, 2)):
    src = int(rng.integers(0, n - length))
    seg = text[src:src + length].copy()
    for _ in range(copies):
        dst = int(rng.integers(n // 2, n - length))
        text[dst:dst + length] = seg
text[n - 3000:] = text[n - 6000:n - 3000]  # and one at the very end of the text
gen_s = time.time() - t0
t0 = time.time()
obj = pkg.SuffixArray(text, idx_bytes=8, devices=[r % visible for r in range(ranks)])
alloc_s = time.time() - t0
t0 = time.time()
obj.construct()
build_s = time.time() - t0
sa, lcp = obj.SA(), obj.LCP()
assert sa.dtype == np.uint64
t0 = time.time()
code, bad = oracle_lib.check_sa_lcp_mt(text, sa, lcp)
check_s = time.time() - t0
stats = obj._rank_stats
print(json.dumps({
    "n": n, "idx_bytes": 8, "ranks": ranks, "gpus": visible, "text_generation_s": round(gen_s, 1),
    "pinned_alloc_s": round(alloc_s, 1), "construct_wall_s": round(build_s, 3),
    "max_sa": int(sa.max()), "max_lcp": int(lcp.max()),
    "checker": {"code": int(code), "first_bad_position": int(bad) if code else None, "seconds": round(check_s, 1)},
    "rank_stats": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()
                    if k in ("ms_pack", "ms_partition", "ms_sort", "ms_refine", "ms_deep_lcp", "ms_total", "ms_h2d", "ms_d2h",
                             "shard_offset", "shard_count", "tied_after_key_sort", "refine_rounds", "key_bits", "comm_bytes")}
                   for s in stats],
}), flush=True)
sys.exit(0 if code == 0 else 1)
