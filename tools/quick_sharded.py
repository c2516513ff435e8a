#!/usr/bin/env python
"""Timing probe of the sharded path inside one process (thread transport).
usage: quick_sharded.py [n] [kind] [ranks...]   e.g. quick_sharded.py 1e8 genome 1 2"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "genome"
rank_counts = [int(a) for a in sys.argv[3:]] or [1, 2]
text = pkg.synth.genome_like(n, seed=3) if kind == "genome" else pkg.synth.random_acgt_chunked(n, 1)
visible = pkg.lib().caps_sa_gpu_device_count()
keys = ("ms_pack", "ms_sort", "ms_partition", "ms_merge", "ms_refine", "ms_deep_lcp", "ms_total", "ms_h2d", "ms_d2h",
        "shard_count", "comm_bytes", "refine_rounds", "kernel_launches")
one = pkg.SuffixArray(text)
for rep in range(2):
    t0 = time.time()
    one.construct()
    print(f"single rep{rep}: wall {1e3 * (time.time() - t0):.1f} ms", {k: round(v, 2) if isinstance(v, float) else v
                                                                   for k, v in one.stats().items() if k in keys}, flush=True)
for ranks in rank_counts:
    obj = pkg.SuffixArray(text, devices=[r % visible for r in range(ranks)])
    for rep in range(2):
        t0 = time.time()
        obj.construct()
        wall = time.time() - t0
        print(f"ranks={ranks} rep{rep}: wall {wall * 1e3:.1f} ms", flush=True)
        for r, st in enumerate(obj._rank_stats):
            print("   rank", r, {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in keys}, flush=True)
    same = np.array_equal(obj.SA(), one.SA()) and np.array_equal(obj.LCP(), one.LCP())
    print(f"ranks={ranks}: equals single-device result: {same}", flush=True)
