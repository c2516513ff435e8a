#!/usr/bin/env python
"""Times the key sort stage alone (stage hook, kernel timing on): usage key_sort_time.py [n] [reps] [lsd]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
use_lsd = len(sys.argv) > 3 and sys.argv[3] == "lsd"
pkg = graft.load_package()
text = pkg.synth.genome_like(n, seed=3, scale=n / 3.1e9)
eng = pkg.Engine(0)
eng.set_kernel_timing(True)
for r in range(reps):
    bits, keys, sa = eng.stage_key_sort(text, use_lsd=use_lsd)
    st = eng.stats()
    out = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()
           if k.startswith(("ms_msd", "msd_")) or k in ("ms_sort", "ms_scatter", "scatter_bytes", "key_bits")}
    for cls in ("scatter_a", "scatter_b", "local", "hist"):
        ms, by = st[f"ms_msd_{cls}"], st[f"msd_{cls}_bytes"]
        if ms > 0:
            out[f"{cls}_GBs"] = round(by / ms / 1e6, 1)
    print(("lsd" if use_lsd else "msd"), "n=%d" % n, out, flush=True)
ok = bool((keys[1:] >= keys[:-1]).all())
print("ascending:", ok, "distinct suffixes:", len(np.unique(sa)) == n if n <= 200_000_000 else "skipped")
