"""caps-sa_b200 — B200-native suffix-array + LCP construction (drop-in for the construction
path of jamshed/CaPS-SA).

Python mirror of the reference's C++ surface, bound over the C-ABI in
include/caps_sa_gpu.h with ctypes:

    sa = SuffixArray(text_bytes, subproblem_count=0, max_context=0)   # ctor   (hpp:155)
    sa.construct()                                                    # :177
    sa.SA(), sa.LCP(), sa.T(), sa.n()                                 # :165-174
    sa.dump(path)                                                     # :180, layout cpp:497-509

The compute path is the CUDA library only; importing works without a GPU (so CPU-side
tests can check symbols), but every compute call raises if the library or a device is
missing — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import synth  # noqa: F401  (generators used by tests and bench.py)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libcaps_sa_gpu.so")

EXPORTED_SYMBOLS = [
    "caps_sa_gpu_device_count", "caps_sa_gpu_last_error", "caps_sa_gpu_engine_create",
    "caps_sa_gpu_engine_destroy", "caps_sa_gpu_engine_stats", "caps_sa_gpu_engine_set_stream",
    "caps_sa_gpu_engine_set_kernel_timing", "caps_sa_gpu_construct_u32",
    "caps_sa_gpu_construct_u64", "caps_sa_gpu_construct_device_u32", "caps_sa_gpu_construct_device_u64",
    "caps_sa_gpu_map_acgt", "caps_sa_gpu_host_alloc", "caps_sa_gpu_host_free", "caps_sa_gpu_stage_pack",
    "caps_sa_gpu_stage_radix_sort_u64_u32", "caps_sa_gpu_stage_scan_u32",
    "caps_sa_gpu_construct_multi_u32", "caps_sa_gpu_construct_multi_u64", "caps_sa_gpu_comm_unique_id",
    "caps_sa_gpu_engine_comm_init", "caps_sa_gpu_construct_sharded_device_u32",
    "caps_sa_gpu_construct_sharded_device_u64", "caps_sa_gpu_construct_sharded_u32",
    "caps_sa_gpu_construct_sharded_u64", "caps_sa_gpu_shard_copy", "caps_sa_gpu_stage_key_sort_u32",
    "caps_sa_gpu_set_cli_byte_mapping",
]
COMM_ID_BYTES = 128


class Stats(C.Structure):
    _fields_ = [("n", C.c_uint64), ("idx_bytes", C.c_uint32), ("bits_per_symbol", C.c_uint32),
                ("alphabet_size", C.c_uint32), ("refine_rounds", C.c_uint32),
                ("tied_after_key_sort", C.c_uint64), ("deep_lcp_direct", C.c_uint64),
                ("deep_lcp_long", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_pack", C.c_float), ("ms_sort", C.c_float), ("ms_heads", C.c_float),
                ("ms_refine", C.c_float), ("ms_deep_lcp", C.c_float), ("ms_total", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("scatter_launches", C.c_uint32),
                ("ms_scatter", C.c_float), ("scatter_bytes", C.c_uint64), ("key_bits", C.c_uint32),
                ("reserved", C.c_uint32), ("ms_partition", C.c_float), ("ms_merge", C.c_float),
                ("comm_bytes", C.c_uint64), ("shard_offset", C.c_uint64), ("shard_count", C.c_uint64),
                ("pairs_chained", C.c_uint64),
                ("msd_a_bits", C.c_uint32), ("msd_b_bits", C.c_uint32), ("msd_large_buckets", C.c_uint32),
                ("msd_reserved", C.c_uint32), ("msd_large_records", C.c_uint64),
                ("ms_msd_scatter_a", C.c_float), ("ms_msd_scatter_b", C.c_float), ("ms_msd_local", C.c_float),
                ("ms_msd_hist", C.c_float),
                ("msd_scatter_a_bytes", C.c_uint64), ("msd_scatter_b_bytes", C.c_uint64),
                ("msd_local_bytes", C.c_uint64), ("msd_hist_bytes", C.c_uint64),
                ("msd_scatter_a_launches", C.c_uint32), ("msd_scatter_b_launches", C.c_uint32),
                ("msd_local_launches", C.c_uint32), ("msd_hist_launches", C.c_uint32)]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class CapsSaError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded C-ABI library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CapsSaError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        u64, p, i32, u32 = C.c_uint64, C.c_void_p, C.c_int, C.c_uint32
        L.caps_sa_gpu_device_count.restype = i32
        L.caps_sa_gpu_last_error.restype = C.c_char_p
        L.caps_sa_gpu_engine_create.argtypes = [i32]
        L.caps_sa_gpu_engine_create.restype = p
        L.caps_sa_gpu_engine_destroy.argtypes = [p]
        L.caps_sa_gpu_engine_destroy.restype = None
        L.caps_sa_gpu_engine_stats.argtypes = [p, C.POINTER(Stats)]
        L.caps_sa_gpu_engine_set_stream.argtypes = [p, p]
        L.caps_sa_gpu_engine_set_kernel_timing.argtypes = [p, i32]
        for name in ("caps_sa_gpu_construct_u32", "caps_sa_gpu_construct_u64"):
            getattr(L, name).argtypes = [p, p, u64, p, p, u64, u64]
        for name in ("caps_sa_gpu_construct_device_u32", "caps_sa_gpu_construct_device_u64"):
            getattr(L, name).argtypes = [p, p, u64, p, p, p]
        L.caps_sa_gpu_map_acgt.argtypes = [p, p, u64]
        L.caps_sa_gpu_host_alloc.argtypes = [C.c_size_t]
        L.caps_sa_gpu_host_alloc.restype = p
        L.caps_sa_gpu_host_free.argtypes = [p]
        L.caps_sa_gpu_host_free.restype = None
        L.caps_sa_gpu_stage_pack.argtypes = [p, p, u64, p, C.POINTER(u64), C.POINTER(u32)]
        L.caps_sa_gpu_stage_radix_sort_u64_u32.argtypes = [p, p, p, u64, C.c_uint, C.c_uint]
        L.caps_sa_gpu_stage_scan_u32.argtypes = [p, p, u64, i32]
        L.caps_sa_gpu_stage_key_sort_u32.argtypes = [p, p, u64, i32, p, p]
        L.caps_sa_gpu_set_cli_byte_mapping.argtypes = [i32]
        L.caps_sa_gpu_set_cli_byte_mapping.restype = i32
        for name in ("caps_sa_gpu_construct_multi_u32", "caps_sa_gpu_construct_multi_u64"):
            getattr(L, name).argtypes = [C.POINTER(i32), i32, p, u64, p, p, u64, u64, C.POINTER(Stats)]
        L.caps_sa_gpu_comm_unique_id.argtypes = [p]
        L.caps_sa_gpu_engine_comm_init.argtypes = [p, p, i32, i32]
        for name in ("caps_sa_gpu_construct_sharded_device_u32", "caps_sa_gpu_construct_sharded_device_u64"):
            getattr(L, name).argtypes = [p, p, u64, p]
        for name in ("caps_sa_gpu_construct_sharded_u32", "caps_sa_gpu_construct_sharded_u64"):
            getattr(L, name).argtypes = [p, p, u64, p, p]
        L.caps_sa_gpu_shard_copy.argtypes = [p, p, p, i32]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise CapsSaError(f"caps_sa_gpu error {rc}: {lib().caps_sa_gpu_last_error().decode()}")


class PinnedArray:
    """numpy array over pinned host memory from caps_sa_gpu_host_alloc.

    The memory belongs to the buffer object the array is built on, not to this wrapper: it is
    returned to the driver when the last array (or view) that refers to it has gone, so
    ``arr = SuffixArray(t).SA()`` stays valid after the SuffixArray is collected."""

    def __init__(self, count: int, dtype, _alloc=None, _free=None):
        import weakref

        self.dtype = np.dtype(dtype)
        self.nbytes = max(1, count * self.dtype.itemsize)
        alloc = _alloc or lib().caps_sa_gpu_host_alloc
        free = _free or lib().caps_sa_gpu_host_free
        self.ptr = alloc(self.nbytes)
        if not self.ptr:
            raise CapsSaError("pinned host allocation failed: " + lib().caps_sa_gpu_last_error().decode())
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        # numpy keeps `buf` alive through .base for as long as any view of the array exists
        self._finalizer = weakref.finalize(buf, free, self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=count)


class Engine:
    """Per-device engine (stream + scratch pools)."""

    def __init__(self, device: int = 0):
        self._h = lib().caps_sa_gpu_engine_create(device)
        if not self._h:
            raise CapsSaError("cannot create GPU engine: " + lib().caps_sa_gpu_last_error().decode())
        self.device = device

    def close(self) -> None:
        if self._h:
            lib().caps_sa_gpu_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def stats(self) -> dict:
        s = Stats()
        _check(lib().caps_sa_gpu_engine_stats(self._h, C.byref(s)))
        return s.as_dict()

    def set_stream(self, stream: int | None) -> None:
        """Issue all work on the given cudaStream_t (e.g. torch's current stream); None = own."""
        _check(lib().caps_sa_gpu_engine_set_stream(self._h, stream or None))

    def set_kernel_timing(self, enabled: bool) -> None:
        _check(lib().caps_sa_gpu_engine_set_kernel_timing(self._h, int(enabled)))

    # -- host-buffer construction ---------------------------------------------------------
    def construct(self, text: np.ndarray, sa_out: np.ndarray, lcp_out: np.ndarray,
                  subproblem_count: int = 0, max_context: int = 0) -> None:
        assert text.dtype == np.uint8 and text.flags.c_contiguous
        assert sa_out.dtype == lcp_out.dtype and sa_out.dtype in (np.uint32, np.uint64)
        n = len(text)
        assert len(sa_out) == n and len(lcp_out) == n
        fn = lib().caps_sa_gpu_construct_u32 if sa_out.dtype == np.uint32 else lib().caps_sa_gpu_construct_u64
        _check(fn(self._h, text.ctypes.data, n, sa_out.ctypes.data, lcp_out.ctypes.data,
                  subproblem_count, max_context))

    # -- device-resident construction (raw device pointers, e.g. torch tensors' data_ptr) --
    def construct_device(self, d_text: int, n: int, d_sa: int, d_lcp: int, idx_bytes: int = 4,
                         stream: int = 0) -> None:
        fn = (lib().caps_sa_gpu_construct_device_u32 if idx_bytes == 4
              else lib().caps_sa_gpu_construct_device_u64)
        _check(fn(self._h, d_text, n, d_sa, d_lcp, stream))

    # -- sharded construction, one process per GPU (see multi_gpu.py for the torchrun plumbing) --
    def comm_init(self, comm_id: bytes, rank: int, world: int) -> None:
        """Collective: joins the NCCL communicator described by the 128-byte id."""
        assert len(comm_id) == COMM_ID_BYTES
        buf = C.create_string_buffer(bytes(comm_id), COMM_ID_BYTES)
        _check(lib().caps_sa_gpu_engine_comm_init(self._h, buf, rank, world))

    def construct_sharded_device(self, d_text: int, n: int, idx_bytes: int = 4, stream: int = 0) -> None:
        fn = (lib().caps_sa_gpu_construct_sharded_device_u32 if idx_bytes == 4
              else lib().caps_sa_gpu_construct_sharded_device_u64)
        _check(fn(self._h, d_text, n, stream))

    def construct_sharded(self, text: np.ndarray, sa_out: np.ndarray, lcp_out: np.ndarray) -> None:
        """Host buffers; this rank's shard lands at sa_out[offset:offset+count] (see stats())."""
        assert text.dtype == np.uint8 and text.flags.c_contiguous
        assert sa_out.dtype == lcp_out.dtype and sa_out.dtype in (np.uint32, np.uint64)
        fn = (lib().caps_sa_gpu_construct_sharded_u32 if sa_out.dtype == np.uint32
              else lib().caps_sa_gpu_construct_sharded_u64)
        _check(fn(self._h, text.ctypes.data, len(text), sa_out.ctypes.data, lcp_out.ctypes.data))

    def shard_copy(self, sa_dst: int, lcp_dst: int, to_host: bool) -> None:
        _check(lib().caps_sa_gpu_shard_copy(self._h, sa_dst, lcp_dst, int(to_host)))

    def map_acgt(self, text: np.ndarray) -> None:
        assert text.dtype == np.uint8 and text.flags.c_contiguous and text.flags.writeable
        _check(lib().caps_sa_gpu_map_acgt(self._h, text.ctypes.data, len(text)))

    # -- stage-level hooks for the parity tests ----------------------------------------------
    def stage_pack(self, text: np.ndarray):
        n = len(text)
        words = np.zeros(-(-n * 8 // 64) + 2, dtype=np.uint64)
        nwords = C.c_uint64(0)
        sigma = C.c_uint32(0)
        bits = lib().caps_sa_gpu_stage_pack(self._h, text.ctypes.data, n, words.ctypes.data,
                                            C.byref(nwords), C.byref(sigma))
        if bits < 0:
            _check(-bits)
        return bits, words[:nwords.value].copy(), sigma.value

    def stage_radix_sort(self, keys: np.ndarray, vals: np.ndarray, begin_bit: int = 0, end_bit: int = 64):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        _check(lib().caps_sa_gpu_stage_radix_sort_u64_u32(self._h, keys.ctypes.data, vals.ctypes.data,
                                                          len(keys), begin_bit, end_bit))
        return keys, vals

    def stage_key_sort(self, text: np.ndarray, use_lsd: bool = False):
        """(key_bits, keys, sa): all suffixes of `text` ordered by their leading key_bits key bits
        (keys left-aligned in 64 bits); equal keys in no particular order."""
        assert text.dtype == np.uint8 and text.flags.c_contiguous
        n = len(text)
        keys = np.empty(n, dtype=np.uint64)
        sa = np.empty(n, dtype=np.uint32)
        bits = lib().caps_sa_gpu_stage_key_sort_u32(self._h, text.ctypes.data, n, int(use_lsd), keys.ctypes.data,
                                                    sa.ctypes.data)
        if bits < 0:
            _check(-bits)
        return bits, keys, sa

    def stage_scan(self, data: np.ndarray, inclusive_max: bool) -> np.ndarray:
        data = np.ascontiguousarray(data, dtype=np.uint32).copy()
        _check(lib().caps_sa_gpu_stage_scan_u32(self._h, data.ctypes.data, len(data), int(inclusive_max)))
        return data


def comm_unique_id() -> bytes:
    """A fresh NCCL unique id (rank 0 creates it, the host side broadcasts it)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib().caps_sa_gpu_comm_unique_id(buf))
    return buf.raw


def construct_multi(text: np.ndarray, sa_out: np.ndarray, lcp_out: np.ndarray, devices,
                    subproblem_count: int = 0, max_context: int = 0) -> list:
    """Sharded construction inside one process: rank r runs on CUDA device devices[r] (a device
    may be listed several times).  Returns the per-rank statistics."""
    assert text.dtype == np.uint8 and text.flags.c_contiguous
    assert sa_out.dtype == lcp_out.dtype and sa_out.dtype in (np.uint32, np.uint64)
    n = len(text)
    assert len(sa_out) == n and len(lcp_out) == n
    devs = (C.c_int * len(devices))(*devices)
    stats = (Stats * len(devices))()
    fn = lib().caps_sa_gpu_construct_multi_u32 if sa_out.dtype == np.uint32 else lib().caps_sa_gpu_construct_multi_u64
    _check(fn(devs, len(devices), text.ctypes.data, n, sa_out.ctypes.data, lcp_out.ctypes.data,
              subproblem_count, max_context, stats))
    return [s.as_dict() for s in stats]


_default_engines: dict[int, Engine] = {}


def default_engine(device: int = 0) -> Engine:
    if device not in _default_engines:
        _default_engines[device] = Engine(device)
    return _default_engines[device]


class SuffixArray:
    """Mirror of CaPS_SA::Suffix_Array<idx_t> (reference include/Suffix_Array.hpp:22-181).

    idx width follows the reference CLI's rule (src/main.cpp:76): 32-bit iff n <= 2^32-1,
    unless idx_bytes is given.  The text is borrowed (kept alive by this object)."""

    def __init__(self, text, subproblem_count: int = 0, max_context: int = 0, idx_bytes: int | None = None,
                 engine: Engine | None = None, devices=None):
        if isinstance(text, (bytes, bytearray, memoryview)):
            text = np.frombuffer(bytes(text), dtype=np.uint8)
        self._text = np.ascontiguousarray(text, dtype=np.uint8)
        self._n = len(self._text)
        if idx_bytes is None:
            idx_bytes = 4 if self._n <= 0xFFFFFFFF else 8
        self._dtype = np.uint32 if idx_bytes == 4 else np.uint64
        self._p = subproblem_count
        self._ctx = max_context
        self._engine = engine
        self._devices = list(devices) if devices is not None else None  # several ranks -> sharded path
        # result arrays are allocated here, as the reference allocates SA_/LCP_ in its
        # constructor (src/Suffix_Array.cpp:20-21); pinned so the D2H copy runs at PCIe speed
        self._sa_mem = PinnedArray(self._n, self._dtype)
        self._lcp_mem = PinnedArray(self._n, self._dtype)
        self._built = False

    def T(self) -> np.ndarray:
        return self._text

    def n(self) -> int:
        return self._n

    def construct(self) -> None:
        if self._devices is not None:
            self._rank_stats = construct_multi(self._text, self._sa_mem.array, self._lcp_mem.array, self._devices,
                                               self._p, self._ctx)
            self._stats = self._rank_stats[0]
        else:
            eng = self._engine or default_engine()
            eng.construct(self._text, self._sa_mem.array, self._lcp_mem.array, self._p, self._ctx)
            self._stats = eng.stats()
        self._built = True

    def SA(self) -> np.ndarray:
        if not self._built:
            raise CapsSaError("construct() has not been called")
        return self._sa_mem.array

    def LCP(self) -> np.ndarray:
        if not self._built:
            raise CapsSaError("construct() has not been called")
        return self._lcp_mem.array

    def stats(self) -> dict:
        return dict(self._stats)

    def dump(self, path_or_file) -> None:
        """size_t n || SA[n] || LCP[n]  (reference src/Suffix_Array.cpp:497-509)."""
        close = False
        f = path_or_file
        if isinstance(path_or_file, (str, os.PathLike)):
            f = open(path_or_file, "wb")
            close = True
        try:
            f.write(np.uint64(self._n).tobytes())
            f.write(self.SA().tobytes())
            f.write(self.LCP().tobytes())
        finally:
            if close:
                f.close()
