"""Build libevavos_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m evavos_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libevavos_sm100.so")
SOURCES = ["api.cu", "bank.cu", "select_simt.cu", "score_tc.cu", "readout.cu", "aggregate.cu", "merge.cu", "argmax.cu", "attention.cu", "peer.cu", "metrics.cu", "select_dense.cu", "decoder_ops.cu"]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + \
          [os.path.join(os.path.dirname(HERE), "include", "evavos.h")]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC or add /usr/local/cuda/bin to PATH")


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True builds libevavos_sm100_trace.so with -DEVAVOS_TRACE (kernel timeline, debugging only)."""
    out = OUT.replace(".so", "_trace.so") if trace else OUT
    if not force and not trace and not is_stale():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DEVAVOS_TRACE"] if trace else []) + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libevavos_sm100.so")
    return out


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv)
    print(path)
