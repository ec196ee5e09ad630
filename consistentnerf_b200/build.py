"""Builds libcnerf.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m consistentnerf_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU
box with the working tree.  Nothing here depends on torch.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "libcnerf.so")
SOURCES = ["api.cu", "sampling.cu", "composite.cu", "crossview.cu", "linear_simt.cu", "mlp_tc.cu", "mlp_bwd_tc.cu", "mlp_fwd3.cu", "mlp_fwd5.cu"]
# opt-in experiments (a measured negative result): linked only by `--experiments` / CNERF_BUILD_EXPERIMENTS=1
EXPERIMENT_SOURCES = ["experiments/mlp_fwd4.cu", "experiments/pair_selftest.cu", "experiments/mlp_fwd6.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, flags) -> str:
    h = hashlib.sha256(" ".join(flags).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src: str, verbose: bool, flags) -> str:
    obj = os.path.join(OBJ, os.path.splitext(os.path.basename(src))[0] + ".o")
    cmd = [_nvcc(), *flags, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]      # -I: experiments/ include the shared headers
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build_library(force: bool = False, verbose: bool = False, experiments: bool = None) -> str:
    """Compile if sources changed (content hash), return the path of libcnerf.so."""
    if experiments is None:
        experiments = os.environ.get("CNERF_BUILD_EXPERIMENTS", "0") == "1"
    os.makedirs(OBJ, exist_ok=True)
    sources = SOURCES + (EXPERIMENT_SOURCES if experiments else [])
    flags = NVCC_FLAGS + (["-DCNERF_EXPERIMENTS"] if experiments else [])
    deps = [os.path.join(CSRC, s) for s in sources] + [os.path.join(CSRC, h) for h in ("common.cuh", "umma.cuh", "mlp_layout.cuh", "mlp_blocks.cuh", "soft_weight.h")] + [
        os.path.join(PKG, "..", "include", "cnerf.h"), os.path.join(PKG, "..", "include", "cnerf_debug.h")]
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest(deps, flags)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    with cf.ThreadPoolExecutor(max_workers=len(sources)) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose, flags), sources))
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv,
                        experiments=True if "--experiments" in sys.argv else None))
