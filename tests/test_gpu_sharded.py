"""GPU parity tests of the sharded (multi-rank) construction.  The ranks run as host threads
inside one process (caps_sa_gpu_construct_multi_*); listing device 0 several times puts
several ranks on one GPU, so the whole exchange logic — slice sort, pivots, all-to-all, bucket
merge, rank exchange per refinement round, LCP round trip — is exercised on a one-GPU box.
With more GPUs visible the same tests also spread the ranks over the devices."""
import os

import numpy as np
import pytest

import oracle_lib

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

_oracle_cache = {}

CASES = {
    "acgt_1M": lambda s: s.random_acgt(1_000_000, 101),
    "acgt_odd_77777": lambda s: s.random_acgt(77_777, 102),
    "genome_like_2M": lambda s: s.genome_like(2_000_000, seed=103, scale=0.004),
    "bytes256_300k": lambda s: s.random_bytes(300_000, 104),
    "sigma2_300k": lambda s: s.random_bytes(300_000, 106, sigma=2, base=0xFF),
    "periodic_unit1000_200k": lambda s: s.periodic_random_unit(200_000, 1000, seed=4),
    "period3_100k": lambda s: s.periodic(100_000, b"ACG"),
    "fibonacci_200k": lambda s: s.fibonacci(200_000),
    "allA_50k": lambda s: np.full(50_000, ord("A"), dtype=np.uint8),
    "repeat_groups_1500k": lambda s: s.repeat_groups(1_500_000),  # tied groups of 40 .. 5000 suffixes
}


def oracle(case, synth, idx_bytes=4):
    key = (case, idx_bytes)
    if key not in _oracle_cache:
        text = CASES[case](synth)
        # the plain-C restatement: the reference itself writes out of bounds when trailing
        # partitions are empty (src/Suffix_Array.cpp:439-440), which all-equal texts provoke
        sa, lcp = oracle_lib.port_sa_lcp(text, subproblems=16, idx_bytes=idx_bytes)
        assert oracle_lib.check_sa_lcp(text, sa, lcp) == (0, 0)
        _oracle_cache[key] = (text, sa, lcp)
    return _oracle_cache[key]


def device_list(pkg, ranks):
    visible = pkg.lib().caps_sa_gpu_device_count()
    return [r % visible for r in range(ranks)]


def check_shards(stats, n):
    spans = sorted((s["shard_offset"], s["shard_count"]) for s in stats)
    at = 0
    for off, cnt in spans:
        if cnt:
            assert off == at
            at += cnt
    assert at == n


@pytest.mark.parametrize("ranks", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("case", sorted(CASES))
def test_sharded_matches_oracle(pkg, synth, case, ranks):
    text, want_sa, want_lcp = oracle(case, synth)
    obj = pkg.SuffixArray(text, devices=device_list(pkg, ranks))
    obj.construct()
    stats = obj._rank_stats
    check_shards(stats, len(text))
    assert np.array_equal(obj.SA(), want_sa), f"SA differs ({stats})"
    assert np.array_equal(obj.LCP(), want_lcp), f"LCP differs ({stats})"


# The fused partition pass (CAPSB_SHARD_P2P=1: suffix indices stored straight into the owners'
# buckets, csrc/partition.cuh).  Correct on one and two GPUs, thread and NCCL/CUDA-IPC transports
# (profiles/r02 s08), but slower than the partition pass + ncclSend/Recv it would replace: off by default.
@pytest.mark.parametrize("ranks", [1, 2, 3, 8])
@pytest.mark.parametrize("case", ["acgt_1M", "genome_like_2M", "bytes256_300k", "fibonacci_200k", "allA_50k"])
def test_sharded_p2p_partition_matches_oracle(pkg, synth, case, ranks, monkeypatch):
    monkeypatch.setenv("CAPSB_SHARD_P2P", "1")
    text, want_sa, want_lcp = oracle(case, synth)
    obj = pkg.SuffixArray(text, devices=device_list(pkg, ranks))
    obj.construct()
    check_shards(obj._rank_stats, len(text))
    assert np.array_equal(obj.SA(), want_sa)
    assert np.array_equal(obj.LCP(), want_lcp)


@pytest.mark.parametrize("ranks", [2, 3])
def test_sharded_bounded_context_caps_lcp(pkg, synth, ranks):
    """Bounded context through the sharded entry point: exact SA, LCP capped (see
    tests/test_gpu_parity.py: test_bounded_context for the comparison with the reference)."""
    text, want_sa, want_lcp = oracle("genome_like_2M", synth)
    sa = np.empty(len(text), dtype=np.uint32)
    lcp = np.empty(len(text), dtype=np.uint32)
    pkg.construct_multi(text, sa, lcp, device_list(pkg, ranks), subproblem_count=16, max_context=12)
    assert np.array_equal(sa, want_sa)
    assert np.array_equal(lcp, np.minimum(want_lcp, 12))


@pytest.mark.parametrize("ranks", [2, 5])
@pytest.mark.parametrize("case", ["acgt_odd_77777", "fibonacci_200k", "bytes256_300k"])
def test_sharded_u64_indices(pkg, synth, case, ranks):
    text, want_sa, want_lcp = oracle(case, synth, idx_bytes=8)
    obj = pkg.SuffixArray(text, idx_bytes=8, devices=device_list(pkg, ranks))
    obj.construct()
    assert obj.SA().dtype == np.uint64
    assert np.array_equal(obj.SA(), want_sa) and np.array_equal(obj.LCP(), want_lcp)


@pytest.mark.parametrize("ranks", [2, 4])
def test_sharded_tiny_inputs(pkg, ranks):
    # fewer suffixes than ranks, empty slices and empty buckets
    for raw in (b"A", b"AC", b"banana\n", b"mississippi", b"AAAAAAAAAAAAAAA", b"ACGTACGTACGTACGTACGTACGTA"):
        text = np.frombuffer(raw, dtype=np.uint8)
        obj = pkg.SuffixArray(text, devices=device_list(pkg, ranks))
        obj.construct()
        nsa, nlcp = oracle_lib.naive_sa_lcp(text)
        assert np.array_equal(obj.SA(), nsa.astype(np.uint32)), raw
        assert np.array_equal(obj.LCP(), nlcp.astype(np.uint32)), raw


@pytest.mark.parametrize("ranks", [2, 3, 8])
@pytest.mark.parametrize("case", ["acgt_1M", "genome_like_2M", "bytes256_300k", "fibonacci_200k", "allA_50k"])
def test_sharded_merge_mode_matches_oracle(pkg, synth, case, ranks, monkeypatch):
    """CAPSB_SHARD_MODE=merge: the reference's order of stages (slice sort, pivot location in the
    sorted slices, exchange of sorted runs, merge-path bucket merge) instead of partition-first."""
    monkeypatch.setenv("CAPSB_SHARD_MODE", "merge")
    text, want_sa, want_lcp = oracle(case, synth)
    obj = pkg.SuffixArray(text, devices=device_list(pkg, ranks))
    obj.construct()
    assert any(s["ms_merge"] > 0 for s in obj._rank_stats)
    assert np.array_equal(obj.SA(), want_sa) and np.array_equal(obj.LCP(), want_lcp)


def test_sharded_equals_single_device_path(pkg, synth):
    text = synth.genome_like(5_000_000, seed=7, scale=0.01)
    one = pkg.SuffixArray(text)
    one.construct()
    many = pkg.SuffixArray(text, devices=device_list(pkg, 4))
    many.construct()
    assert np.array_equal(one.SA(), many.SA()) and np.array_equal(one.LCP(), many.LCP())
    assert oracle_lib.check_sa_lcp(text, many.SA(), many.LCP()) == (0, 0)


def test_sharded_bad_device_fails_loudly(pkg, synth):
    text = synth.random_acgt(1000, 1)
    obj = pkg.SuffixArray(text, devices=[0, 99])
    with pytest.raises(pkg.CapsSaError):
        obj.construct()


def test_one_process_per_gpu_nccl(pkg):
    """torchrun + NCCL transport (needs at least two GPUs; skipped on a one-GPU box)."""
    import os
    import subprocess
    import sys

    visible = pkg.lib().caps_sa_gpu_device_count()
    if visible < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = min(visible, 4)
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(root, "tools", "sharded_check.py")],
        capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and "SHARDED_CHECK PASSED" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]


def test_one_process_per_gpu_nccl_p2p_partition(pkg):
    """The fused partition pass under torchrun: the bucket buffers cross the process boundary as
    CUDA IPC handles (needs at least two GPUs)."""
    import subprocess
    import sys

    visible = pkg.lib().caps_sa_gpu_device_count()
    if visible < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = min(visible, 4)
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29733", os.path.join(root, "tools", "sharded_check.py")],
        capture_output=True, text=True, timeout=900, env=dict(os.environ, CAPSB_SHARD_P2P="1"))
    assert proc.returncode == 0 and "SHARDED_CHECK PASSED" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]
