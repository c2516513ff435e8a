#!/usr/bin/env python
"""Wall-clock time of the caps_sa CLI (file in -> dump file out) next to the reference CLI built from the
unmodified sources (oracle/_ref/caps_sa_ref), on BASELINE configs 1 and 2; the two dumps must be identical.
usage: cli_time.py [config1|config2|<bases>] [subproblem-count]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "config1"
pkg = graft.load_package()
tmp = tempfile.mkdtemp(prefix="capsb_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
src = os.path.join(tmp, "input")
if what == "config1":
    raw, p = pkg.synth.ecoli_like_fasta(seed=1), "64"
elif what == "config2":  # utils/gen_rand_seq.py's shape: the sequence and a newline
    raw, p = np.concatenate([pkg.synth.random_acgt_chunked(100_000_000, 1), np.array([10], dtype=np.uint8)]), ""
else:
    raw, p = np.concatenate([pkg.synth.random_acgt_chunked(int(float(what)), 1), np.array([10], dtype=np.uint8)]), ""
if len(sys.argv) > 2:
    p = sys.argv[2]
raw.tofile(src)
line = {"config": what, "file_bytes": int(raw.size), "subproblem_count": p or "default (8192)", "host_cores": os.cpu_count()}
outs = {}
for name, exe in (("gpu", os.path.join(ROOT, "bin", "caps_sa")), ("reference", os.path.join(ROOT, "oracle", "_ref", "caps_sa_ref"))):
    if not os.path.exists(exe):
        continue
    out = os.path.join(tmp, name + ".bin")
    times = []
    for rep in range(2 if name == "gpu" else 1):  # the first GPU run pays CUDA context creation
        t0 = time.time()
        proc = subprocess.run([exe, src, out] + ([p] if p else []), capture_output=True, text=True)
        times.append(time.time() - t0)
        if proc.returncode != 0:
            line[name + "_error"] = proc.stderr[-400:]
            break
    line[name + "_wall_s"] = [round(t, 3) for t in times]
    line[name + "_stderr"] = [l for l in proc.stderr.splitlines() if "Time taken" in l][-9:]
    outs[name] = out
if len(outs) == 2:
    line["dumps_identical"] = subprocess.run(["cmp", "-s", outs["gpu"], outs["reference"]]).returncode == 0
print(json.dumps(line), flush=True)
subprocess.run(["rm", "-rf", tmp])
