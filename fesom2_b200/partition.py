"""Mesh partitioning for the multi-GPU path ("the repo's own partitioning": METIS recursive
bisection with the options of the reference's do_partit, src/fort_part.c:46-241)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .mesh import Mesh, node_graph, simple_partition

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "partit.c")
OUT = os.path.join(_HERE, "libfesom_partit.so")
METIS_A = "/usr/local/cuda/lib64/libmetis_static.a"


def build_library(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-o", OUT, SRC, METIS_A, "-lm"])
    return OUT


def metis_partition(g: Mesh, npes: int) -> np.ndarray:
    """part(n) in 0..npes-1 for every node of the global mesh ``g``."""
    if npes == 1:
        return np.zeros(g.Nh, np.int32)
    lib = C.CDLL(build_library())
    lib.fesom_partit.restype = C.c_longlong
    ptr, adj = node_graph(g)
    wgt = np.ascontiguousarray(g.nlevels_nod2D, dtype=np.int32)
    part = np.zeros(g.Nh, np.int32)
    ip = C.POINTER(C.c_int32)
    ec = lib.fesom_partit(int(g.Nh), ptr.ctypes.data_as(ip), adj.ctypes.data_as(ip), wgt.ctypes.data_as(ip),
                          int(npes), part.ctypes.data_as(ip))
    if ec < 0:
        raise RuntimeError("METIS_PartGraphRecursive failed")
    return part


def partition(g: Mesh, npes: int, method: str = "metis") -> np.ndarray:
    if method == "metis":
        return metis_partition(g, npes)
    return simple_partition(g, npes)
