"""ctypes access to the oracle (TEST INFRASTRUCTURE: tests/, smoke() and bench.py's CPU
baseline only — never the product path).

``port``  = oracle/_ref/liboracle_port.so  (plain-C restatement + independent checker)
``ref``   = oracle/_ref/libcaps_sa_ref.so  (the UNMODIFIED reference, compiled in the build
            container from /root/reference by oracle/Makefile; travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_port = None
_ref = None


def build_port() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def build_ref() -> bool:
    """Builds the unmodified reference if /root/reference is mounted; True if available."""
    if os.path.isdir(os.environ.get("CAPS_SA_REFERENCE", "/root/reference")):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)
    return os.path.exists(os.path.join(REF_DIR, "libcaps_sa_ref.so"))


def port():
    global _port
    if _port is None:
        path = os.path.join(REF_DIR, "liboracle_port.so")
        if not os.path.exists(path):
            build_port()
        lib = C.CDLL(path)
        u64, p = C.c_uint64, C.c_void_p
        lib.caps_port_construct.argtypes = [p, u64, u64, u64, p, p]
        lib.caps_port_construct.restype = C.c_int
        lib.caps_port_construct_u32.argtypes = [p, u64, u64, u64, p, p]
        lib.caps_port_construct_u32.restype = C.c_int
        lib.caps_port_map_acgt.argtypes = [p, u64]
        lib.caps_port_map_acgt.restype = None
        lib.caps_check_sa_lcp.argtypes = [p, u64, p, p, C.c_int, C.POINTER(u64)]
        lib.caps_check_sa_lcp.restype = C.c_int
        lib.caps_oracle_set_threads.argtypes = [C.c_int]
        lib.caps_oracle_set_threads.restype = None
        lib.caps_check_sa_lcp_mt.argtypes = [p, u64, p, p, C.c_int, C.POINTER(u64)]
        lib.caps_check_sa_lcp_mt.restype = C.c_int
        lib.caps_check_sa_lcp_mt_pieces.argtypes = [p, u64, p, p, C.c_int, u64, C.POINTER(u64)]
        lib.caps_check_sa_lcp_mt_pieces.restype = C.c_int
        lib.caps_check_sa_lcp_periodic.argtypes = [p, u64, u64, p, p, C.c_int, C.POINTER(u64)]
        lib.caps_check_sa_lcp_periodic.restype = C.c_int
        lib.caps_naive_sa_lcp.argtypes = [p, u64, p, p]
        lib.caps_naive_sa_lcp.restype = C.c_int
        _port = lib
    return _port


def ref():
    """The compiled unmodified reference, or None when it is not available."""
    global _ref
    if _ref is None:
        path = os.path.join(REF_DIR, "libcaps_sa_ref.so")
        if not os.path.exists(path) and not build_ref():
            return None
        lib = C.CDLL(path)
        u64, p = C.c_uint64, C.c_void_p
        for name in ("caps_sa_ref_construct_u32", "caps_sa_ref_construct_u64"):
            fn = getattr(lib, name)
            fn.argtypes = [p, u64, u64, u64, p, p]
            fn.restype = C.c_double
        _ref = lib
    return _ref


def _as_text(text) -> np.ndarray:
    arr = np.frombuffer(bytes(text), dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else text
    return np.ascontiguousarray(arr, dtype=np.uint8)


def port_sa_lcp(text, subproblems: int = 0, max_context: int = 0, idx_bytes: int = 4):
    t = _as_text(text)
    n = len(t)
    sa = np.empty(n, dtype=np.uint64)
    lcp = np.empty(n, dtype=np.uint64)
    rc = port().caps_port_construct(t.ctypes.data, n, subproblems, max_context, sa.ctypes.data,
                                    lcp.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle port failed rc={rc} (n={n})")
    dt = np.uint32 if idx_bytes == 4 else np.uint64
    return sa.astype(dt), lcp.astype(dt)


def ref_sa_lcp(text, subproblems: int = 0, max_context: int = 0, idx_bytes: int = 4):
    """(sa, lcp, construct_seconds) from the unmodified reference."""
    lib = ref()
    if lib is None:
        raise RuntimeError("compiled reference (oracle/_ref/libcaps_sa_ref.so) not available")
    t = _as_text(text)
    n = len(t)
    dt = np.uint32 if idx_bytes == 4 else np.uint64
    sa = np.empty(n, dtype=dt)
    lcp = np.empty(n, dtype=dt)
    fn = lib.caps_sa_ref_construct_u32 if idx_bytes == 4 else lib.caps_sa_ref_construct_u64
    # the reference prints progress to stderr; keep test output clean
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        secs = fn(t.ctypes.data, n, subproblems, max_context, sa.ctypes.data, lcp.ctypes.data)
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(devnull)
    return sa, lcp, secs


def check_sa_lcp(text, sa: np.ndarray, lcp: np.ndarray):
    """(code, first_bad_position); code 0 = valid (see oracle/oracle_port.h)."""
    t = _as_text(text)
    sa = np.ascontiguousarray(sa)
    lcp = np.ascontiguousarray(lcp)
    assert sa.dtype == lcp.dtype and sa.dtype in (np.uint32, np.uint64)
    bad = C.c_uint64(0)
    rc = port().caps_check_sa_lcp(t.ctypes.data, len(t), sa.ctypes.data, lcp.ctypes.data,
                                  sa.dtype.itemsize, C.byref(bad))
    return rc, bad.value


def check_sa_lcp_mt(text, sa: np.ndarray, lcp: np.ndarray, max_pieces: int = 0):
    """check_sa_lcp with OpenMP loops (oracle/sa_check.c: caps_check_sa_lcp_mt): what bench.py applies
    to its full-size results.  No copies are made: pass contiguous arrays."""
    t = _as_text(text)
    assert sa.flags.c_contiguous and lcp.flags.c_contiguous
    assert sa.dtype == lcp.dtype and sa.dtype in (np.uint32, np.uint64)
    bad = C.c_uint64(0)
    # all the cores this process may run on, whatever OMP_NUM_THREADS the launcher exported
    port().caps_oracle_set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    rc = port().caps_check_sa_lcp_mt_pieces(t.ctypes.data, len(t), sa.ctypes.data, lcp.ctypes.data,
                                            sa.dtype.itemsize, max_pieces, C.byref(bad))
    return rc, bad.value


def check_sa_lcp_periodic(text, period: int, sa: np.ndarray, lcp: np.ndarray):
    """check_sa_lcp for a text that is its first `period` bytes repeated: multi-threaded, LCP from
    the closed form for periodic texts (oracle/sa_check.c)."""
    t = _as_text(text)
    sa = np.ascontiguousarray(sa)
    lcp = np.ascontiguousarray(lcp)
    assert sa.dtype == lcp.dtype and sa.dtype in (np.uint32, np.uint64)
    bad = C.c_uint64(0)
    rc = port().caps_check_sa_lcp_periodic(t.ctypes.data, len(t), period, sa.ctypes.data, lcp.ctypes.data,
                                           sa.dtype.itemsize, C.byref(bad))
    return rc, bad.value


def naive_sa_lcp(text):
    t = _as_text(text)
    n = len(t)
    sa = np.empty(n, dtype=np.uint64)
    lcp = np.empty(n, dtype=np.uint64)
    port().caps_naive_sa_lcp(t.ctypes.data, n, sa.ctypes.data, lcp.ctypes.data)
    return sa, lcp


def dump_bytes(n: int, sa: np.ndarray, lcp: np.ndarray) -> bytes:
    """On-disk layout of the reference's dump(): size_t n || SA || LCP
    (src/Suffix_Array.cpp:497-509)."""
    return np.uint64(n).tobytes() + np.ascontiguousarray(sa).tobytes() + np.ascontiguousarray(lcp).tobytes()
