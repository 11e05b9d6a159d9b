"""Build timing / tracing variants of libevavos_sm100.so (objects of unchanged sources are reused).

    python scripts/build_variants.py name1=-DFLAG1,-DFLAG2 name2=... -> evavos_b200/libevavos_sm100_<name>.so
Only score_tc.cu, select_simt.cu and api.cu see the variant flags; use with EVAVOS_LIB=... scripts/filter_time.py.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evavos_b200.build import CSRC, NVCC_FLAGS, SOURCES, _nvcc  # noqa: E402

VARIANT_SOURCES = {"score_tc.cu", "select_simt.cu", "api.cu", "readout.cu", "select_dense.cu"}
OBJ = os.path.join(ROOT, "build", "obj")
os.makedirs(OBJ, exist_ok=True)
flags = [f for f in NVCC_FLAGS if f != "-shared"]


def compile_obj(src, extra, tag):
    out = os.path.join(OBJ, f"{os.path.splitext(src)[0]}_{tag}.o")
    srcp = os.path.join(CSRC, src)
    deps = [srcp] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
           [os.path.join(ROOT, "include", "evavos.h")]
    if os.path.exists(out) and all(os.path.getmtime(d) < os.path.getmtime(out) for d in deps):
        return out
    subprocess.run([_nvcc()] + flags + extra + ["-c", "-o", out, srcp], check=True)
    return out


procs = []
for spec in sys.argv[1:]:
    name, _, fl = spec.partition("=")
    extra = [f for f in fl.split(",") if f]
    objs = [compile_obj(s, extra if s in VARIANT_SOURCES else [], name if s in VARIANT_SOURCES else "base") for s in SOURCES]
    out = os.path.join(ROOT, "evavos_b200", f"libevavos_sm100_{name}.so")
    subprocess.run([_nvcc(), "-shared", "-o", out] + objs, check=True)
    print(out)
