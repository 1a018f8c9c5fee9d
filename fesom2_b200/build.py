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


HOST_DIR = os.path.join(_HERE, "host")
HOST_BIN = os.path.join(_HERE, "dwarf_tracer_b200")
HOST_SRC = [os.path.join(HOST_DIR, "fesom_host.cpp"), os.path.join(HOST_DIR, "fesom_restart.cpp"), os.path.join(HOST_DIR, "dwarf_tracer.cpp")]


def build_host(force: bool = False) -> str:
    """The compiled host side above the C ABI (fesom2_b200/host: C++ mirror of the reference interface + the tracer dwarf),
    linked against the in-tree library (rpath $ORIGIN, so the binary travels with the snapshot)."""
    deps = HOST_SRC + [os.path.join(HOST_DIR, "fesom_host.hpp"), os.path.join(HOST_DIR, "fesom_restart.hpp"), DEPS[-1], OUT]
    newest = max(os.path.getmtime(p) for p in deps)
    if force or not os.path.exists(HOST_BIN) or os.path.getmtime(HOST_BIN) < newest:
        cxx = os.environ.get("CXX", "g++")
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-ffp-contract=off", "-I", os.path.join(_HERE, "..", "include"), "-I", HOST_DIR]
                              + HOST_SRC + ["-L", _HERE, "-lfesom_adv_b200", "-pthread", "-Wl,-rpath,$ORIGIN", "-o", HOST_BIN])
    return HOST_BIN
