#!/usr/bin/env python
"""bench.py — SA+LCP construction throughput (BASELINE.json metric: suffixes/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--n SIZE]
                    [--impl ours|reference]

A step is one full construction (text -> SA + LCP) of the named synthetic text.
  value : suffixes/s with the text already resident in HBM and SA/LCP left in HBM
          (device-resident entry point, CUDA events on the stream the kernels run on).
  e2e   : suffixes/s through the reference-facing host-buffer call
          (caps_sa_gpu_construct_u32: pinned host text -> H2D -> construct -> D2H SA+LCP),
          copies inside the timed region.
  roofline     : the dominant kernel (radix_scatter_kernel), per-launch CUDA-event times
                 gathered inside the library during the timed steps.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, OpenMP stand-in for ParlayLib) timed
                 on the box's host cores on a bounded prefix of the same text (rank 0, N=1).
--impl reference runs only that CPU arm, K+W times, and prints its own JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as graft  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (fits one B200)
    "genome3g": dict(n=3_100_000_000, kind="genome_like", seed=3,
                     desc="3.1 Gbp synthetic genome: uniform ACGT + injected repeats (SURVEY §8d config 3), 32-bit indices"),
    # BASELINE.json configs[1]
    "random100m": dict(n=100_000_001, kind="random_acgt_nl", seed=1,
                       desc="100 Mbp uniform random ACGT + newline mapped to C (config 2)"),
    "genome100m": dict(n=100_000_000, kind="genome_like", seed=3, desc="config 3 scaled to 100 Mbp"),
}
# DRAM traffic of the scatter kernel relative to its algorithmic bytes, from the committed
# `ncu --set full` capture (profiles/r01/scatter_ncu_full_v4.csv: dram__bytes_read.sum 1.310726 GB +
# dram__bytes_write.sum 1.234758 GB per launch over 100 000 001 (u64, u32) pairs = 24 B each)
NCU_TRAFFIC_RATIO = (1.310726e9 + 1.234758e9) / (24.0 * 100_000_001)
CPU_SAMPLE = 310_000_000  # symbols the CPU reference is timed on: 1/10 of the 3.1 Gbp workload (10-15 s per construct())
CPU_SUBPROBLEMS = 256      # the reference's best case in SURVEY.md §6 (default 8192 is ~2x slower)
REF_BUILD = ("oracle/_ref built by oracle/Makefile: g++ -std=c++17 -Ofast -funroll-loops -march=x86-64-v3 -fopenmp -DNDEBUG "
             "(the reference's CMakeLists.txt:49 has -Ofast -mavx2 -funroll-loops -march=native; native is replaced by "
             "x86-64-v3 = AVX2/BMI2 so the binary built in the build container runs on the GPU box's host); "
             "OpenMP stand-in for ParlayLib's three scheduling calls")


def make_text(pkg, spec, n):
    kind = spec["kind"]
    if kind == "genome_like":
        return pkg.synth.genome_like(n, seed=spec["seed"], scale=n / 3.1e9)
    if kind == "random_acgt_nl":
        t = pkg.synth.random_acgt_chunked(n, spec["seed"])
        t[-1] = ord("C")  # the CLI maps the trailing newline of gen_rand_seq.py's output to 'C'
        return t
    raise ValueError(kind)


# Kernel classes timed inside the library (CUDA events around every launch, on the launching stream):
# stats prefix -> (name, what one launch's algorithmic bytes are)
KERNEL_CLASSES = {
    "scatter": ("radix_scatter_kernel", "(u64 key, u32 suffix) pairs read once and written once: 24 B per pair"),
    "msd_scatter_a": ("msd_scatter_kernel, level A", "text window read (~1 B) + one 8-byte record written per suffix"),
    "msd_scatter_b": ("msd_scatter_kernel, level B", "one 8-byte record read and written per suffix: 16 B"),
    "msd_local": ("msd_local_kernel (both instantiations: 512 threads, and 1024 threads for the buckets of 6145..12288 records)",
                  "8-byte record read, 8-byte key + 4-byte suffix written: 20 B per suffix"),
    "msd_hist": ("msd_hist_kernel", "text window (level A) or one 8-byte record (level B) read per suffix"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch over the algorithmic bytes of the launch, from the
# committed `ncu --set full` captures (profiles/): filled in per kernel class as they are captured
_S14 = "profiles/r02/ncu_full_msd_kernels_s14.csv"  # one construction of the 3.1 Gbp workload, 3.1e9 records per level
NCU_TRAFFIC = {
    "scatter": (NCU_TRAFFIC_RATIO, "profiles/r01/scatter_ncu_full_v4.csv"),
    "msd_scatter_a": ((0.798767e9 + 29.127431e9) / (9.0 * 3.1e9), _S14),
    "msd_scatter_b": ((25.027414e9 + 24.795954e9) / (16.0 * 3.1e9), _S14),
    "msd_local": ((23.619910e9 + 35.203110e9 + 1.300618e9 + 1.885808e9) / (20.0 * 3.1e9),
                  "profiles/r02/ncu_full_final_kernels_s24_summary.csv"),
}


class KernelClassTimes:
    """Accumulates the per-kernel-class launch times / bytes / launch counts of engine.stats()."""

    def __init__(self):
        self.ms = {k: 0.0 for k in KERNEL_CLASSES}
        self.bytes = {k: 0 for k in KERNEL_CLASSES}
        self.launches = {k: 0 for k in KERNEL_CLASSES}

    def add(self, st):
        for k in KERNEL_CLASSES:
            ms_key, b_key, l_key = (("ms_scatter", "scatter_bytes", "scatter_launches") if k == "scatter"
                                    else (f"ms_{k}", f"{k}_bytes", f"{k}_launches"))
            self.ms[k] += st.get(ms_key, 0.0)
            self.bytes[k] += st.get(b_key, 0)
            self.launches[k] += st.get(l_key, 0)

    def report(self, peak, peak_kind, steps, step_ms):
        table = {}
        for k, (name, what) in KERNEL_CLASSES.items():
            if self.launches[k] == 0 or self.ms[k] <= 0:
                continue
            gbs = self.bytes[k] / 1e9 / (self.ms[k] / 1e3)
            table[k] = {"kernel": name, "launches": self.launches[k], "ms_per_step": self.ms[k] / steps,
                        "achieved_GBps": gbs, "frac": gbs / peak, "share_of_step": self.ms[k] / steps / step_ms,
                        "algorithmic_bytes": what}
        if not table:
            return None, table
        top = max(table, key=lambda k: table[k]["ms_per_step"])
        t = table[top]
        ratio, source = NCU_TRAFFIC.get(top, (None, None))
        per_launch = self.bytes[top] / self.launches[top]
        roofline = {"kernel": t["kernel"], "bound": "hbm", "achieved": t["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": t["frac"], "peak_source": peak_kind,
                    "traffic": ratio * per_launch if ratio else None,
                    "traffic_source": (f"ncu --set full capture of this kernel ({source}): dram read + write = {ratio:.3f} x "
                                       "algorithmic bytes, scaled to this run's mean launch") if ratio else
                                      "no ncu --set full capture of this kernel committed yet",
                    "launches_timed": self.launches[top], "avg_launch_ms": self.ms[top] / self.launches[top],
                    "share_of_step": t["share_of_step"], "algorithmic_bytes_per_launch": per_launch,
                    "algorithmic_bytes": t["algorithmic_bytes"]}
        return roofline, table


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows = []
        self.proc = None
        self.index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])), smax.append(float(r[2]))
            except ValueError:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_run(text, subproblems, threads):
    """One construct() of the unmodified reference on `text`; returns seconds (the region
    the reference itself times, src/Suffix_Array.cpp:466-494)."""
    import oracle_lib

    os.environ["PARLAY_NUM_THREADS"] = str(threads)
    if oracle_lib.ref() is not None:
        _, _, secs = oracle_lib.ref_sa_lcp(text, subproblems=subproblems)
        return secs, "reference"
    t0 = time.time()
    oracle_lib.port_sa_lcp(text, subproblems=subproblems)
    return time.time() - t0, "port"


def verify_result(text, sa, lcp):
    """Independent check of a finished (SA, LCP) against the text: oracle/sa_check.c
    caps_check_sa_lcp_mt — permutation, suffix order through the inverse permutation, LCP by
    Kasai's walk (the predicate of the reference's is_sorted, src/Suffix_Array.cpp:512-536).
    Runs after the timed regions; a non-zero code fails the run."""
    import oracle_lib

    t0 = time.time()
    code, bad = oracle_lib.check_sa_lcp_mt(text, sa, lcp)
    return {"checker": "oracle/sa_check.c caps_check_sa_lcp_mt (permutation + suffix order via ISA + Kasai LCP, OpenMP)",
            "what": "the SA and LCP arrays the last end-to-end step left in the host buffers, all n entries",
            "code": int(code), "first_bad_position": int(bad) if code else None, "n": int(len(text)),
            "seconds": round(time.time() - t0, 1)}


def sample_text(pkg, spec, n, sample_n):
    """The CPU arm's input: the workload recipe at sample size (same repeat structure, scaled), not a
    prefix of the full text — so `--impl reference` need not generate all n symbols."""
    if sample_n >= n:
        return make_text(pkg, spec, n)
    if spec["kind"] == "genome_like":
        return pkg.synth.genome_like(sample_n, seed=spec["seed"], scale=sample_n / 3.1e9)
    return make_text(pkg, spec, sample_n)


def run_reference_arm(args, spec, n):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = graft.load_package()
    sample_n = min(n, args.cpu_sample)
    text = sample_text(pkg, spec, n, sample_n)
    threads = os.cpu_count() or 1
    times = []
    kind = "reference"
    for _ in range(args.warmup):
        cpu_reference_run(text, args.cpu_subproblems, threads)
    for _ in range(args.steps):
        secs, kind = cpu_reference_run(text, args.cpu_subproblems, threads)
        times.append(secs)
    per_step = float(np.mean(times))
    value = sample_n / per_step
    line = {
        "impl": "reference", "metric": "sa_lcp_suffixes_per_sec", "value": value, "unit": "suffixes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['desc']}", "n": sample_n, "sample_of": n, "idx_bytes": 4,
                   "note": "n is what this arm constructs per step: the workload recipe at sample size "
                           "(the CPU reference needs minutes per construct() at the full n)"},
        "cpu_baseline": {"value": value, "unit": "suffixes/s", "cores": threads, "kind": kind,
                         "seconds_per_step": per_step, "build": REF_BUILD,
                         "sample": f"the workload recipe at {sample_n} symbols (1/{max(1, round(n / sample_n))} of n={n}), "
                                   f"subproblem_count={args.cpu_subproblems}, PARLAY_NUM_THREADS={threads}, construct() only"},
        "e2e": {"value": value, "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="genome3g", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=float, default=None, help="override the text length")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=CPU_SAMPLE)
    ap.add_argument("--cpu-subproblems", type=int, default=CPU_SUBPROBLEMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-verify", action="store_true", help="profiling runs only: skip the post-run check of SA/LCP")
    args = ap.parse_args()
    spec = WORKLOADS[args.workload]
    n = int(args.n) if args.n else spec["n"]

    if args.impl == "reference":
        run_reference_arm(args, spec, n)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        # a rank that dies must not leave its peers waiting in a collective for the default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))

    pkg = graft.load_package()
    if world > 1:
        from caps_sa_b200 import multi_gpu  # sharded path (torch.distributed plumbing + our kernels)
        out = multi_gpu.bench_main(args, spec, n, pkg, make_text, ClockSampler, load_peaks, KernelClassTimes)
        dist.barrier()
        dist.destroy_process_group()
        if out is not None:  # rank 0: check the assembled result of the last end-to-end step, then report
            line, (text_host, sa_host, lcp_host) = out
            line["verified"] = None if args.no_verify else verify_result(text_host, sa_host, lcp_host)
            print(json.dumps(line), flush=True)
            if line["verified"] and line["verified"]["code"] != 0:
                raise SystemExit("bench.py: the benchmarked SA/LCP failed verification (see \"verified\")")
        return

    t_gen = time.time()
    text_np = make_text(pkg, spec, n)
    gen_s = time.time() - t_gen

    eng = pkg.Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    # pinned host buffers (the class shell allocates its SA_/LCP_ the same way)
    text_pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    text_pin.numpy()[:] = text_np
    sa_pin = torch.empty(n, dtype=torch.int32, pin_memory=True)
    lcp_pin = torch.empty(n, dtype=torch.int32, pin_memory=True)
    text_host = text_pin.numpy()
    sa_host = sa_pin.numpy().view(np.uint32)
    lcp_host = lcp_pin.numpy().view(np.uint32)

    d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_text.copy_(text_pin, non_blocking=True)
    d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
    d_lcp = torch.empty(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def device_step():
        eng.construct_device(d_text.data_ptr(), n, d_sa.data_ptr(), d_lcp.data_ptr(), 4, stream.cuda_stream)

    def e2e_step():
        eng.construct(text_host, sa_host, lcp_host)

    # ---- device-resident: `value` ---------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    eng.set_kernel_timing(True)
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    kernel_times = KernelClassTimes()
    stage_ms = {}
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
        st = eng.stats()
        launches += st["kernel_launches"]
        kernel_times.add(st)
        for k in ("ms_pack", "ms_sort", "ms_heads", "ms_refine", "ms_deep_lcp", "ms_total"):
            stage_ms[k] = stage_ms.get(k, 0.0) + st[k] / args.steps
    ev1.record(stream)
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    last_stats = eng.stats()
    eng.set_kernel_timing(False)

    # ---- end to end through the host-buffer C-ABI call: `e2e` ------------------------------
    for _ in range(0 if args.no_e2e else max(1, args.warmup - 2)):
        e2e_step()
    torch.cuda.synchronize()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    t0 = time.perf_counter()
    for _ in range(0 if args.no_e2e else args.steps):
        e2e_step()
    ev3.record(stream)
    torch.cuda.synchronize()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max(ev2.elapsed_time(ev3) / args.steps, e2e_wall_ms)
    clocks = sampler.stop()
    e2e_stats = eng.stats()  # of the last end-to-end step
    e2e_parts = {k: round(e2e_stats[k], 3) for k in ("ms_h2d", "ms_pack", "ms_sort", "ms_refine", "ms_deep_lcp", "ms_total", "ms_d2h")}

    # the device-resident leg left the same arrays in HBM as the end-to-end leg left in the host buffers
    # (all n entries of SA and LCP, compared on the device piece by piece)
    same = None
    if not args.no_e2e:
        same = True
        piece = 1 << 28
        for lo in range(0, n, piece):
            hi = min(n, lo + piece)
            same = same and bool(torch.equal(d_sa[lo:hi], sa_pin[lo:hi].cuda(non_blocking=True)))
            same = same and bool(torch.equal(d_lcp[lo:hi], lcp_pin[lo:hi].cuda(non_blocking=True)))
    # independent check of what was benchmarked (outside the timed regions)
    verified = None
    if not (args.no_e2e or args.no_verify):
        verified = verify_result(text_np, sa_host, lcp_host)

    peak, peak_kind = load_peaks()
    roofline, kernel_table = kernel_times.report(peak, peak_kind, args.steps, dev_ms)
    if roofline is not None:
        roofline["pipeline_frac"] = (n * 9 / 1e9) / (dev_ms / 1e3) / peak  # 9 B/suffix compulsory traffic

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sample_n = min(n, args.cpu_sample)
        threads = os.cpu_count() or 1
        secs, kind = cpu_reference_run(sample_text(pkg, spec, n, sample_n), args.cpu_subproblems, threads)
        cpu_baseline = {"value": sample_n / secs, "unit": "suffixes/s", "cores": threads, "kind": kind,
                        "seconds": secs, "build": REF_BUILD,
                        "sample": f"the workload recipe at {sample_n} symbols (1/{max(1, round(n / sample_n))} of n={n}), "
                                  f"subproblem_count={args.cpu_subproblems}, PARLAY_NUM_THREADS={threads}, construct() only, one run"}

    line = {
        "metric": "sa_lcp_suffixes_per_sec", "value": n / (dev_ms / 1e3), "unit": "suffixes/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['desc']}", "n": n, "idx_bytes": 4,
                   "bits_per_symbol": last_stats["bits_per_symbol"], "l2": "inputs >> 126 MB L2, no explicit flush",
                   "text_generation_s": round(gen_s, 1), "tied_after_key_sort": last_stats["tied_after_key_sort"],
                   "refine_rounds": last_stats["refine_rounds"],
                   "key_sort": (f"packed-record MSD sort, {last_stats['msd_a_bits']} + {last_stats['msd_b_bits']} bits then one CTA per bucket; "
                                f"{last_stats['msd_large_buckets']} oversized buckets ({last_stats['msd_large_records']} records) by the LSD passes"
                                if last_stats["msd_a_bits"] else "LSD passes")},
        "e2e": {"value": n / (e2e_ms / 1e3), "unit": "suffixes/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": n, "d2h_bytes_per_step": 2 * 4 * n, "matches_device_result": same,
                "last_step_ms": e2e_parts,
                "note": "ms_total = compute stream, ms_d2h = first result copy issued .. last byte on the host; the ranges of "
                        "SA/LCP positions are copied out while later ranges are still being refined"},
        "verified": verified,
        "gpu_launches": int(launches),
        "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
        "roofline": roofline, "kernels": kernel_table, "cpu_baseline": cpu_baseline, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if (verified and verified["code"] != 0) or same is False:
        raise SystemExit("bench.py: the benchmarked SA/LCP failed verification (see \"verified\" / \"matches_device_result\")")


if __name__ == "__main__":
    main()
