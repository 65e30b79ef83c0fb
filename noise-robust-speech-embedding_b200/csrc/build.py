"""Build libnrse_b200.so (the C-ABI library of include/nrse_b200.h) in-tree with plain nvcc for sm_100a.

No torch headers are involved: the library takes raw device pointers and a cudaStream_t, so it builds in
seconds and has no ABI coupling to PyTorch.  Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libnrse_b200.so")
SOURCES = ["runtime.cu", "check.cu", "mix.cu", "ema.cu", "optim.cu", "pool.cu", "loss.cu", "allreduce.cu", "posconv.cu", "frontend.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "epilogue_math.cuh", os.path.join(ROOT, "include", "nrse_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the nrse_b200 CUDA library cannot be built (there is no CPU fallback)")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """Product library: csrc/libnrse_b200.so.  ``experiments=True`` builds the scripts-only timing variant
    csrc/build_exp/libnrse_b200_exp.so with -DNRSE_EXPERIMENTS (its NRSE_EXPERIMENT environment hooks skip stores /
    statistics and produce WRONG results; select it with NRSE_B200_LIB from scripts/, never from tests or bench.py --
    both refuse a library whose nrse_experiments_build() returns 1)."""
    build_dir = os.path.join(HERE, "build_exp") if experiments else BUILD
    lib = os.path.join(build_dir, "libnrse_b200_exp.so") if experiments else LIB
    flags = NVCC_FLAGS + (["-DNRSE_EXPERIMENTS"] if experiments else [])
    os.makedirs(build_dir, exist_ok=True)
    nvcc = _nvcc()
    headers = [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]

    def compile_one(src: str):
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        if not force and not _newer(obj, [os.path.join(HERE, src)] + headers):
            return obj, ""
        cmd = [nvcc, *flags, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _newer(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, experiments="--experiments" in sys.argv))
