"""Build libfasn.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python flash-attention-softmax-n_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The library lands next to the Python package
(`flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn.so`) so it travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "flash_attention_softmax_n", "libfasn.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SOURCES = ["fasn_api.cu", "fasn_fwd.cu", "fasn_bwd.cu", "fasn_bwd_aux.cu", "fasn_aux.cu", "fasn_softmax.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", INCLUDE]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_dep() -> float:
    t = os.path.getmtime(os.path.join(INCLUDE, "fasn.h"))
    for f in os.listdir(CSRC):
        t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def build_variant(name: str, extra_flags) -> str:
    """Experimental variant libfasn_<name>.so with extra nvcc flags (tuning sweeps; select it with FASN_LIBRARY)."""
    global OBJ, LIB
    old = (OBJ, LIB, list(NVCC_FLAGS))
    try:
        OBJ, LIB = os.path.join(HERE, "build_" + name), os.path.join(HERE, "flash_attention_softmax_n", f"libfasn_{name}.so")
        NVCC_FLAGS.extend(extra_flags)
        return build(force=True)
    finally:
        OBJ, LIB = old[0], old[1]
        NVCC_FLAGS[:] = old[2]


def build(force: bool = False, verbose: bool = False, timeline: bool = False) -> str:
    """Compile every translation unit (in parallel) and link the shared library; returns its path.
    `timeline=True` builds the instrumented variant libfasn_timeline.so (-DFASN_TIMELINE, scripts/timeline.py)."""
    global OBJ, LIB
    flags = list(NVCC_FLAGS)
    if timeline:
        OBJ, LIB = os.path.join(HERE, "build_timeline"), os.path.join(HERE, "flash_attention_softmax_n", "libfasn_timeline.so")
        flags.append("-DFASN_TIMELINE")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_dep():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, timeline="--timeline" in sys.argv))
