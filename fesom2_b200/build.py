"""Build the CUDA library in-tree (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "adv_capi.cu")
DEPS = [SRC, os.path.join(_HERE, "csrc", "adv_kernels.cuh"),
        os.path.join(_HERE, "..", "include", "fesom_adv_b200.h")]
OUT = os.path.join(_HERE, "libfesom_adv_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              # parity with the non-contracted CPU restatement: no FMA contraction on the device,
              # none in the host-side precomputation either
              "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def build_library(force: bool = False, verbose: bool = False) -> str:
    newest = max(os.path.getmtime(p) for p in DEPS)
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < newest:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC, "-ldl"]
        subprocess.check_call(cmd)
    return OUT
