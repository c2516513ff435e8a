"""In-tree build of the native code (no pip, no JIT cache): nvcc -> lib/libcaps_sa_gpu.so,
g++ -> lib/libcore.a (class shell) and bin/caps_sa (CLI).  sm_100a only."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(PKG, "build")
BIN_DIR = os.path.join(ROOT, "bin")
LIB_PATH = os.path.join(LIB_DIR, "libcaps_sa_gpu.so")

CU_SOURCES = ["text_pack.cu", "sa_build.cu", "sharded_build.cu", "comm.cu", "capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-diag-suppress", "186"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    # the image's $CXX wrapper lacks parts of the toolchain; prefer the distro compiler
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(BIN_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "caps_sa_gpu.h"))

    jobs = []
    objs = []
    for src in CU_SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not _newer(obj, [src_path] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, "-c", src_path, "-o", obj])
    with cf.ThreadPoolExecutor(max_workers=4) as pool:
        list(pool.map(_run, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs])
        if verbose:
            print("built", LIB_PATH)

    # host side: class shell (static lib `core`, as in the reference's src/CMakeLists.txt:11)
    # and the caps_sa CLI
    cxx = _host_cxx()
    inc = os.path.join(ROOT, "include")
    shell_src = os.path.join(ROOT, "src", "Suffix_Array.cpp")
    main_src = os.path.join(ROOT, "src", "main.cpp")
    if os.path.exists(shell_src) and os.path.exists(main_src):
        shell_obj = os.path.join(OBJ_DIR, "Suffix_Array.o")
        core_a = os.path.join(LIB_DIR, "libcore.a")
        cli = os.path.join(BIN_DIR, "caps_sa")
        hdrs = [os.path.join(inc, "Suffix_Array.hpp"), os.path.join(inc, "caps_sa_gpu.h")]
        if force or not _newer(core_a, [shell_src] + hdrs):
            _run([cxx, "-std=c++17", "-O2", "-fPIC", "-Wall", "-I", inc, "-c", shell_src, "-o", shell_obj])
            if os.path.exists(core_a):
                os.remove(core_a)
            _run(["ar", "rcs", core_a, shell_obj])
        if force or not _newer(cli, [main_src, core_a, LIB_PATH] + hdrs):
            _run([cxx, "-std=c++17", "-O2", "-Wall", "-I", inc, main_src, core_a, "-L", LIB_DIR,
                  "-lcaps_sa_gpu", "-Wl,-rpath," + LIB_DIR, "-Wl,-rpath,$ORIGIN/../caps-sa_b200/lib",
                  "-o", cli])
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
