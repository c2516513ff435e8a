#!/usr/bin/env python
"""BASELINE config 4 at full size on one B200: highly repetitive byte-alphabet text, 1 Gbp.

    python tools/config4_check.py [--n 1e9] [--text periodic|fibonacci] [--unit 1000]

Builds SA + LCP through the host-buffer C-ABI call and validates the result without the CPU
reference (which needs hours on such text): the periodic text with the
multi-threaded closed-form checker (oracle/sa_check.c: caps_check_sa_lcp_periodic), the
Fibonacci text with the generic linear-time checker.  Prints one JSON line.
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=float, default=1e9)
    ap.add_argument("--text", default="periodic", choices=["periodic", "fibonacci"])
    ap.add_argument("--unit", type=int, default=1000)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    n = int(args.n)
    pkg = graft.load_package()
    t0 = time.time()
    text = pkg.synth.periodic_random_unit(n, args.unit, seed=4) if args.text == "periodic" else pkg.synth.fibonacci(n)
    gen_s = time.time() - t0

    import torch

    eng = pkg.Engine(0)
    sa = torch.empty(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    lcp = torch.empty(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    t0 = time.time()
    eng.construct(text, sa, lcp)
    first_s = time.time() - t0
    t0 = time.time()
    eng.construct(text, sa, lcp)
    second_s = time.time() - t0
    st = eng.stats()
    line = {"config": f"BASELINE configs[3]: {args.text} byte text, n={n}" + (f", unit={args.unit}" if args.text == "periodic" else ""),
            "n": n, "bits_per_symbol": st["bits_per_symbol"], "alphabet_size": st["alphabet_size"],
            "text_generation_s": round(gen_s, 2), "construct_s_cold": round(first_s, 3), "construct_s": round(second_s, 3),
            "device_ms": {k: round(st[k], 2) for k in ("ms_pack", "ms_sort", "ms_refine", "ms_deep_lcp", "ms_total", "ms_h2d", "ms_d2h")},
            "refine_rounds": st["refine_rounds"], "tied_after_key_sort": st["tied_after_key_sort"],
            "max_lcp": int(lcp.max()), "suffixes_per_s_e2e": n / second_s}
    if not args.no_check:
        t0 = time.time()
        if args.text == "periodic":
            rc, bad = oracle_lib.check_sa_lcp_periodic(text, args.unit, sa, lcp)
            line["checker"] = "caps_check_sa_lcp_periodic (closed-form LCP, OpenMP)"
        else:
            # Kasai's walk in one range per thread: every range pays one comparison from scratch, and the
            # common prefixes of a Fibonacci word run to hundreds of millions of symbols
            rc, bad = oracle_lib.check_sa_lcp_mt(text, sa, lcp, max_pieces=len(os.sched_getaffinity(0)))
            line["checker"] = "caps_check_sa_lcp_mt (ISA order + Kasai, OpenMP, one range per thread)"
        line["check_s"] = round(time.time() - t0, 1)
        line["check_code"] = rc
        line["check_bad_position"] = bad
        line["valid"] = rc == 0
    print(json.dumps(line), flush=True)
    return 0 if line.get("valid", True) else 1


if __name__ == "__main__":
    sys.exit(main())
