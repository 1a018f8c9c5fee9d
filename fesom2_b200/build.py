"""Build the CUDA library in-tree (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "adv_capi.cu")
DEPS = [SRC, os.path.join(_HERE, "csrc", "adv_kernels.cuh"), os.path.join(_HERE, "csrc", "adv_pipe.cuh"),
        os.path.join(_HERE, "..", "include", "fesom_adv_b200.h")]
OUT = os.path.join(_HERE, "libfesom_adv_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              # parity with the non-contracted CPU restatement: no FMA contraction on the device,
              # none in the host-side precomputation either
              "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def build_library(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """``out``/``defines`` build tuning variants next to the product library (experiments only)."""
    newest = max(os.path.getmtime(p) for p in DEPS)
    if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC, "-ldl"]
        subprocess.check_call(cmd)
    return out
