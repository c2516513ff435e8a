"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/caps_sa_gpu.h declares, and fails loudly (no CPU fallback) without a device."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "caps_sa_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(caps_sa_gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/caps_sa_gpu.h but not exported"
    assert set(pkg.EXPORTED_SYMBOLS) == set(names)


def test_library_has_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.join(ROOT, "caps-sa_b200", "lib", "libcaps_sa_gpu.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device(pkg):
    lib = pkg.lib()
    if lib.caps_sa_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.CapsSaError):
        pkg.Engine(0)
    with pytest.raises(pkg.CapsSaError):
        # the constructor already needs the CUDA runtime (pinned result arrays)
        sa = pkg.SuffixArray(np.frombuffer(b"ACGTACGTACGTACGTACGT", dtype=np.uint8))
        sa.construct()


def test_cli_fails_loudly_without_device(pkg, tmp_path):
    if pkg.lib().caps_sa_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    src = tmp_path / "in.txt"
    src.write_bytes(b"ACGT" * 16)
    proc = subprocess.run([os.path.join(ROOT, "bin", "caps_sa"), str(src), str(tmp_path / "out")],
                          capture_output=True, text=True)
    assert proc.returncode != 0
    assert "CUDA" in proc.stderr


def test_cli_usage_message():
    proc = subprocess.run([os.path.join(ROOT, "bin", "caps_sa")], capture_output=True, text=True)
    assert proc.returncode != 0
    assert proc.stderr.startswith("Usage: CaPS_SA <input_path> <output_path>")


def test_class_header_keeps_reference_surface():
    hdr = open(os.path.join(ROOT, "include", "Suffix_Array.hpp")).read()
    for needle in ("namespace CaPS_SA", "class Suffix_Array",
                   "Suffix_Array(const char* T, idx_t n, idx_t subproblem_count = 0, idx_t max_context = 0);",
                   "Suffix_Array(const Suffix_Array&) = delete;", "Suffix_Array(Suffix_Array&&) = delete;",
                   "const char* T() const", "idx_t n() const", "const idx_t* SA() const",
                   "const idx_t* LCP() const", "void construct();", "void dump(std::ofstream& output);"):
        assert needle in hdr, needle
    impl = open(os.path.join(ROOT, "src", "Suffix_Array.cpp")).read()
    assert "template class CaPS_SA::Suffix_Array<uint32_t>;" in impl
    assert "template class CaPS_SA::Suffix_Array<uint64_t>;" in impl


def test_pinned_array_memory_outlives_its_wrapper(pkg):
    """The numpy array handed out by SuffixArray.SA()/LCP() owns the pinned block through its
    buffer: the block is freed when the last view goes, not when the wrapper does."""
    import ctypes as C
    import gc

    import numpy as np

    freed, keep = [], []

    def alloc(nbytes):
        keep.append(C.create_string_buffer(nbytes))
        return C.addressof(keep[-1])

    pa = pkg.PinnedArray(10, np.uint32, _alloc=alloc, _free=freed.append)
    arr = pa.array
    view = arr[2:5]
    del pa
    gc.collect()
    assert freed == []
    del arr
    gc.collect()
    assert freed == []
    del view
    gc.collect()
    assert freed == [C.addressof(keep[0])]
