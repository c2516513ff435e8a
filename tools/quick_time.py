#!/usr/bin/env python
"""Quick timing probe (not the benchmark): host-buffer construction of a synthetic text,
printing the engine's per-stage CUDA-event times.  usage: quick_time.py [n] [kind] [repeat]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

pkg = g.load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "acgt"
repeat = int(sys.argv[3]) if len(sys.argv) > 3 else 2
make = {"acgt": lambda: pkg.synth.random_acgt_chunked(n, 1),
        "genome": lambda: pkg.synth.genome_like(n, seed=3),
        "periodic": lambda: pkg.synth.periodic_random_unit(n, 1000, 4),
        "fib": lambda: pkg.synth.fibonacci(n),
        "bytes": lambda: pkg.synth.random_bytes(n, 2)}[kind]
t0 = time.time()
text = make()
print(f"generated {kind} n={n} in {time.time() - t0:.1f}s", flush=True)
eng = pkg.Engine(0)
for r in range(repeat):
    sa = pkg.SuffixArray(text, engine=eng)
    t0 = time.time()
    sa.construct()
    wall = time.time() - t0
    s = sa.stats()
    print(f"run {r}: wall {wall * 1e3:.1f} ms | " + " ".join(
        f"{k}={s[k]:.2f}" if isinstance(s[k], float) else f"{k}={s[k]}" for k in s), flush=True)
    print(f"   device {n / s['ms_total'] / 1e6:.3f} G suffixes/s, e2e {n / wall / 1e9:.3f} G suffixes/s", flush=True)
    del sa
