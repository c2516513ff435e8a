#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref
(run in the build container, where /root/reference is mounted):

    python tests/golden/make_golden.py

Each fixture stores the input text, the reference's subproblem count, index width and the
reference's SA and LCP.  The fixtures are small on purpose (they are committed); the larger
known answers are kept as sha256 digests of the reference's dump in golden_hashes.json.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "caps-sa_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)


def cases():
    rng = np.random.default_rng(2024)
    acgt = synth.random_acgt
    yield "simpletest2_cli", synth.map_acgt(np.fromfile("/root/reference/data/simpletest2", dtype=np.uint8)), 0, 4
    yield "acgt_2000_p8", acgt(2000, 11), 8, 4
    yield "acgt_min32_p2", acgt(32, 5), 2, 4
    yield "bytes256_3000_p16", synth.random_bytes(3000, 12), 16, 4
    yield "period3_999_p4", synth.periodic(999, b"ACG"), 4, 4
    yield "fib_2584_p8", synth.fibonacci(2584), 8, 4
    yield "allA_500_p4", np.full(500, ord("A"), dtype=np.uint8), 4, 4
    # planted long repeats: LCPs far beyond one 32-base key window
    t = acgt(4000, 13)
    t[2500:3400] = t[300:1200]
    t[3500:3900] = t[100:500]
    yield "acgt_repeats_4000_p8", t, 8, 4
    # poly-A tail and inner poly-A runs: zero-padded key windows at the end of the text
    t = acgt(1500, 14)
    t[200:330] = ord("A")
    t[900:975] = ord("A")
    t[-47:] = ord("A")
    yield "acgt_polyA_tail_1500_p4", t, 4, 4
    t = acgt(1200, 15)
    t[-70:] = ord("T")
    t[400:520] = ord("T")
    yield "acgt_polyT_tail_1200_p4", t, 4, 4
    yield "sigma2_high_1500_p4", synth.random_bytes(1500, 16, sigma=2, base=0x7F), 4, 4
    yield "sigma3_1500_p4", synth.random_bytes(1500, 17, sigma=3, base=ord("x")), 4, 4
    yield "sigma10_2000_p8", synth.random_bytes(2000, 18, sigma=10, base=0xF8), 8, 4
    yield "sigma17_2000_p8", synth.random_bytes(2000, 19, sigma=17, base=0x20), 8, 4
    yield "period_unit37_bytes_3000_p8", synth.periodic(3000, synth.random_bytes(37, 20)), 8, 4
    yield "acgt_1000_u64_p4", acgt(1000, 21), 4, 8
    yield "fib_987_u64_p4", synth.fibonacci(987), 4, 8
    yield "ecoli_like_small_cli", synth.map_acgt(synth.ecoli_like_fasta(seed=1, bases=6000)), 16, 4
    del rng


def main() -> None:
    if oracle_lib.ref() is None:
        sys.exit("oracle/_ref/libcaps_sa_ref.so not available (needs /root/reference)")
    index = {}
    for name, text, p, w in cases():
        sa, lcp, _ = oracle_lib.ref_sa_lcp(text, subproblems=p, idx_bytes=w)
        rc, bad = oracle_lib.check_sa_lcp(text, sa, lcp)
        assert rc == 0, (name, rc, bad)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), text=text, subproblems=np.int64(p),
                            idx_bytes=np.int64(w), sa=sa, lcp=lcp)
        digest = hashlib.sha256(oracle_lib.dump_bytes(len(text), sa, lcp)).hexdigest()
        index[name] = {"n": int(len(text)), "subproblems": int(p), "idx_bytes": int(w),
                       "dump_sha256": digest, "max_lcp": int(lcp.max())}
        print(f"{name:34s} n={len(text):6d} p={p:3d} w={w} maxLCP={int(lcp.max()):6d} {digest[:16]}")

    # larger known answers: digest of the reference's dump only
    big = {}
    for label, n, seed, p in (("acgt_seed2_16M", 16_000_000, 2, 0), ("acgt_seed7_2M", 2_000_000, 7, 64)):
        text = synth.random_acgt(n, seed)
        sa, lcp, secs = oracle_lib.ref_sa_lcp(text, subproblems=p)
        digest = hashlib.sha256(oracle_lib.dump_bytes(n, sa, lcp)).hexdigest()
        big[label] = {"generator": "random_acgt", "n": n, "seed": seed, "subproblems": p,
                      "dump_sha256": digest, "max_lcp": int(lcp.max()), "ref_construct_s": round(secs, 3)}
        print(label, digest, f"{secs:.2f}s")
    with open(os.path.join(HERE, "golden_hashes.json"), "w") as f:
        json.dump({"small": index, "large": big,
                   "survey_appendix_A": {
                       "simpletest2_cli_dump_sha256": "36c1179e82ddbc8d8c7dc2af9c164ea7e6326b22116fe9f2347d2d4488621713",
                       "acgt_seed2_16M_dump_sha256": "7fb73f82d95dc89e62175b33e2c62e324e2084cbfa23f3aa72ab82e1bd8df754",
                       "acgt_seed3_100M_dump_sha256": "4878e2221dadb4144eeb4984bfe3f3602f64cd60883e73c8147c17bb22af6c1a"}},
                  f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
