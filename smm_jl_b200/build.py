"""In-tree build of libsmm_b200.so (hand-written sm_100a CUDA + the C ABI) with nvcc.

    python -m smm_jl_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SO = os.environ.get("SMM_B200_SO") or os.path.join(HERE, "libsmm_b200.so")   # (override: A/B runs of two builds)
# (source, extra flags, object name): smm_kernels.cu is compiled twice, the second time with -DSMM_LL_TU (exchange_mode 3)
UNITS = [("smm_kernels.cu", [], "smm_kernels.o"), ("smm_kernels.cu", ["-DSMM_LL_TU"], "smm_kernels_ll.o"),
         ("smm_api.cu", [], "smm_api.o"), ("smm_stats.cu", [], "smm_stats.o")]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = [os.path.join(CSRC, "smm_device.cuh"), os.path.join(CSRC, "smm_panel.cuh"), os.path.join(INCLUDE, "smm_b200.h"),
           os.path.join(INCLUDE, "smm_stream.h"), os.path.join(INCLUDE, "smm_stream_tables.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",              # products and sums stay separately rounded; fma only where written
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-ffp-contract=off",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    if not force and not needs_build():
        return SO
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src, flags, obj in UNITS:     # the translation units compile side by side
        cmd = [nvcc()] + NVCC_FLAGS + flags + (extra or []) + ["-c", "-o", os.path.join(objdir, obj), os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
        if verbose and out:
            print(out, file=sys.stderr)
    cmd = [nvcc(), "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + \
          [os.path.join(objdir, u[2]) for u in UNITS] + ["-lnccl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else None)
    print(SO)
