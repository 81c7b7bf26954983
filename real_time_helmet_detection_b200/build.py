"""In-tree build of libhd_b200.so (hand-written sm_100a kernels + C ABI).

nvcc cross-compiles for sm_100a without a GPU, so this runs on the CPU-only build container;
the resulting .so sits next to this file and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(PKG_DIR, "build")
LIB = os.path.join(PKG_DIR, "libhd_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libhd_b200.so cannot be built")


def _sources() -> list[str]:
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "hd_b200.h"))
    return max((os.path.getmtime(h) for h in hs if os.path.exists(h)), default=0.0)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    lib_m = os.path.getmtime(LIB)
    if _headers_mtime() > lib_m:
        return True
    return any(os.path.getmtime(os.path.join(CSRC, s)) > lib_m for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    hdr_m = _headers_mtime()
    inc = ["-I", CSRC, "-I", os.path.join(os.path.dirname(PKG_DIR), "include")]

    def compile_one(src: str) -> str:
        obj = os.path.join(BUILD, src[:-3] + ".o")
        src_path = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src_path)
                and os.path.getmtime(obj) > hdr_m):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *inc, "-c", src_path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", LIB + ".tmp", *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(LIB + ".tmp", LIB)
    return LIB


RUNNER_SRC = os.path.join(os.path.dirname(PKG_DIR), "runner", "hd_infer.cpp")
RUNNER_BIN = os.path.join(os.path.dirname(PKG_DIR), "runner", "hd_infer")


def build_runner(force: bool = False) -> str:
    """Native inference runner (runner/hd_infer.cpp): plain C++ over the C ABI, linked against the in-tree .so."""
    lib = build()
    if (not force and os.path.exists(RUNNER_BIN) and os.path.getmtime(RUNNER_BIN) > os.path.getmtime(RUNNER_SRC)
            and os.path.getmtime(RUNNER_BIN) > os.path.getmtime(lib)):
        return RUNNER_BIN
    cmd = [_nvcc(), "-O2", "-std=c++17", RUNNER_SRC, "-o", RUNNER_BIN + ".tmp",
           "-I", os.path.join(os.path.dirname(PKG_DIR), "include"), "-L", PKG_DIR, "-lhd_b200",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../real_time_helmet_detection_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"runner build failed:\n{r.stdout}\n{r.stderr}")
    os.replace(RUNNER_BIN + ".tmp", RUNNER_BIN)
    return RUNNER_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_runner(force="--force" in sys.argv))
