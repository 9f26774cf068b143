"""Builds libqst.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared object is git-ignored but travels to the GPU box with the repo snapshot.
nvcc cross-compiles without a GPU, so this runs in the authoring container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libqst.so")
OBJ_DIR = os.path.join(HERE, "build")

SOURCES = ["common.cu", "quad_loss.cu", "quad_loss_f32.cu", "quad_loss_f16.cu", "quad_loss_bf16.cu", "quad_eval.cu", "prep.cu", "score_select.cu", "finalize.cu", "metrics.cu",
           "comm.cu"]
HEADERS = ["qst_common.cuh", "sm100_ptx.cuh", "select_common.cuh", "quad_loss_kernels.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libqst.so cannot be built (there is no CPU fallback)")


def _newest_input_mtime() -> float:
    paths = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(INCLUDE, "qst.h"), __file__]
    return max(os.path.getmtime(p) for p in paths)


def is_stale() -> bool:
    return (not os.path.isfile(LIB_PATH)) or os.path.getmtime(LIB_PATH) < _newest_input_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libqst.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_hdr = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    newest_hdr = max(newest_hdr, os.path.getmtime(os.path.join(INCLUDE, "qst.h")), os.path.getmtime(__file__))

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        src_path = os.path.join(CSRC, src)
        if (not force and os.path.isfile(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src_path), newest_hdr)):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src_path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
