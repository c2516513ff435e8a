"""Synthetic text generators for the parity tests and bench.py.

All generators return a 1-D ``numpy.uint8`` array (the text bytes) and are pure functions
of their arguments, so tests, fixtures and benchmarks can regenerate the same input on any
box.  Nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_acgt(n: int, seed: int) -> np.ndarray:
    """Uniform random ACGT, exactly the recipe of SURVEY.md Appendix A2 (no newline)."""
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def random_acgt_chunked(n: int, seed: int, chunk: int = 1 << 28) -> np.ndarray:
    """Uniform random ACGT generated chunk by chunk (bounded temporary memory).

    Chunk ``c`` uses ``default_rng([seed, c])``; used for the multi-Gbp workloads."""
    out = np.empty(n, dtype=np.uint8)
    for c, lo in enumerate(range(0, n, chunk)):
        hi = min(n, lo + chunk)
        rng = np.random.default_rng([seed, c])
        out[lo:hi] = _ACGT[rng.integers(0, 4, size=hi - lo, dtype=np.uint8)]
    return out


def random_bytes(n: int, seed: int, sigma: int = 256, base: int = 0) -> np.ndarray:
    """Uniform random text over ``sigma`` byte values ``base, base+1, ...`` (mod 256).

    With sigma=256 both halves of the signed-char range occur, which exercises the
    reference's signed comparison (bytes >= 0x80 sort first)."""
    rng = np.random.default_rng(seed)
    return ((rng.integers(0, sigma, size=n, dtype=np.int64) + base) % 256).astype(np.uint8)


def periodic(n: int, unit: np.ndarray | bytes) -> np.ndarray:
    """``unit`` repeated (and truncated) to length n."""
    u = np.frombuffer(bytes(unit), dtype=np.uint8) if not isinstance(unit, np.ndarray) else unit
    reps = -(-n // len(u))
    return np.tile(u, reps)[:n].copy()


def periodic_random_unit(n: int, unit_len: int = 1000, seed: int = 4) -> np.ndarray:
    """Config 4(i): a random byte unit over 0..255 repeated to length n."""
    return periodic(n, random_bytes(unit_len, seed))


def fibonacci(n: int, a: int = 0x61, b: int = 0xE1) -> np.ndarray:
    """Config 4(ii): Fibonacci word over two byte values (default 0x61 / 0xE1), length n."""
    s0 = np.array([a], dtype=np.uint8)
    s1 = np.array([a, b], dtype=np.uint8)
    while len(s1) < n:
        s0, s1 = s1, np.concatenate([s1, s0])
    return s1[:n].copy()


def genome_like(n: int, seed: int = 3, scale: float | None = None) -> np.ndarray:
    """Config 3: random ACGT plus injected repeats, scaled with n.

    At n = 3.1e9 (scale 1.0) this injects, as SURVEY.md §8d describes: 1,000,000 x 300 bp
    interspersed repeats from 50 family consensus sequences with 10 % substitutions,
    5,000 x 6,000 bp from 5 families with 2 % substitutions, 200 exact segmental
    duplications of 100 kbp, and 20 exact tandem arrays of a 171-bp unit x 2,000 copies.
    Counts scale linearly with n (lengths do not), so smaller n keeps the same LCP profile.
    """
    if scale is None:
        scale = n / 3.1e9
    text = random_acgt_chunked(n, seed)
    rng = np.random.default_rng([seed, 0xC0FFEE])

    def place(seq: np.ndarray) -> None:
        if len(seq) >= n:
            return
        at = int(rng.integers(0, n - len(seq)))
        text[at:at + len(seq)] = seq

    def scatter_family_copies(copies: int, families: int, length: int, rate: float, batch: int) -> None:
        """`copies` mutated copies of `families` random consensus sequences, vectorised."""
        if length >= n:
            return
        consensus = _ACGT[rng.integers(0, 4, size=(families, length), dtype=np.uint8)]
        offs = np.arange(length, dtype=np.int64)
        for lo in range(0, copies, batch):
            cnt = min(batch, copies - lo)
            fam = rng.integers(0, families, size=cnt)
            at = rng.integers(0, n - length, size=cnt)
            block = consensus[fam]
            hit = rng.random((cnt, length)) < rate
            k = int(hit.sum())
            if k:
                block[hit] = _ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
            text[(at[:, None] + offs[None, :]).reshape(-1)] = block.reshape(-1)

    scatter_family_copies(max(1, int(1_000_000 * scale)), 50, 300, 0.10, 100_000)
    scatter_family_copies(max(1, int(5_000 * scale)), 5, 6000, 0.02, 1_000)
    for _ in range(max(1, int(200 * scale))):
        seg_len = min(100_000, n // 8)
        src = int(rng.integers(0, n - seg_len))
        place(text[src:src + seg_len].copy())
    for _ in range(max(1, int(20 * scale))):
        unit = _ACGT[rng.integers(0, 4, size=171, dtype=np.uint8)]
        copies = min(2000, max(2, n // (171 * 16)))
        place(np.tile(unit, copies))
    return text


def repeat_groups(n: int, sizes=(40, 100, 200, 300, 700, 1500, 3000, 5000), block: int = 64, seed: int = 7,
                  mutated=(2000, 100, 0.05)) -> np.ndarray:
    """Random ACGT with, for every k in `sizes`, k exact copies of a random block of `block` symbols
    dropped at random places, plus `mutated = (copies, length, rate)`: copies of one block with
    substitutions.  Every suffix that starts inside a copy ties with its k - 1 siblings on the rest of
    the block and differs right after it: tied groups of (about) every listed size, which is what
    picks the code path of the refinement's group sorts (a warp up to 256, a CTA up to 4096, the
    global sort above) and of the local sort's ordering loop."""
    rng = np.random.default_rng([seed, 0xB10C])
    text = _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
    for k in sizes:
        unit = _ACGT[rng.integers(0, 4, size=block, dtype=np.uint8)]
        for at in rng.integers(0, n - block, size=k):
            text[at:at + block] = unit
    copies, length, rate = mutated
    if copies and length < n:
        unit = _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
        for at in rng.integers(0, n - length, size=copies):
            seg = unit.copy()
            hit = rng.random(length) < rate
            seg[hit] = _ACGT[rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)]
            text[at:at + length] = seg
    return text


def ecoli_like_fasta(seed: int = 1, bases: int = 4_641_652) -> np.ndarray:
    """Config 1 stand-in (data/ecoli.fa is absent from the reference mount): a FASTA-shaped
    file — one header line, 80-column lines — with a few injected repeats.  Returned as raw
    file bytes; the CLI maps every byte (header and newlines included) to ACGT."""
    rng = np.random.default_rng(seed)
    seq = _ACGT[rng.integers(0, 4, size=bases, dtype=np.uint8)]
    for length, copies in ((5000, 7), (1300, 10)):
        if bases <= 2 * length:
            continue
        src = int(rng.integers(0, bases - length))
        seg = seq[src:src + length].copy()
        for _ in range(copies):
            at = int(rng.integers(0, bases - length))
            seq[at:at + length] = seg
    header = np.frombuffer(f">ecoli_like seed={seed}\n".encode(), dtype=np.uint8)
    full, rem = divmod(bases, 80)
    body = np.empty(bases + full + (1 if rem else 0), dtype=np.uint8)
    lines = seq[:full * 80].reshape(full, 80)
    block = np.concatenate([lines, np.full((full, 1), 10, dtype=np.uint8)], axis=1).reshape(-1)
    body[:len(block)] = block
    if rem:
        body[len(block):len(block) + rem] = seq[full * 80:]
        body[-1] = 10
    return np.concatenate([header, body])


def map_acgt(raw: np.ndarray) -> np.ndarray:
    """numpy statement of the CLI byte mapping (reference src/main.cpp:61-70): every byte
    becomes "ACTG"[(toupper(c) & 6) >> 1]; toupper never changes bits 1-2."""
    table = np.frombuffer(b"ACTG", dtype=np.uint8)
    return table[(raw >> 1) & 3]
