"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED: see oracle/adv_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

HOR = {"UPW1": 0, "MUSCL": 1, "MFCT": 2}
VER = {"UPW1": 0, "QR4C": 1, "PPM": 2, "CDIFF": 3}
LIM = {"NON": 0, "NONE": 0, "FCT": 1}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class OraMesh(C.Structure):
    _fields_ = [("nl", C.c_int), ("myDim_nod2D", C.c_int), ("eDim_nod2D", C.c_int),
                ("myDim_elem2D", C.c_int), ("eDim_elem2D", C.c_int), ("myDim_edge2D", C.c_int),
                ("nod_in_elem_ld", C.c_int),
                ("edges", c_ip), ("edge_tri", c_ip), ("elem2D_nodes", c_ip), ("nod_in_elem2D", c_ip),
                ("nod_in_elem2D_num", c_ip), ("nlevels", c_ip), ("ulevels", c_ip),
                ("nlevels_nod2D", c_ip), ("ulevels_nod2D", c_ip),
                ("edge_cross_dxdy", c_dp), ("edge_dxdy", c_dp), ("elem_cos", c_dp),
                ("area", c_dp), ("areasvol", c_dp),
                ("helem", c_dp), ("hnode", c_dp), ("hnode_new", c_dp), ("zbar_3d_n", c_dp),
                ("Z_3d_n", c_dp), ("zbar_n_bot", c_dp)]


class OraWork(C.Structure):
    _fields_ = [("fct_LO", c_dp), ("adv_flux_hor", c_dp), ("adv_flux_ver", c_dp),
                ("fct_ttf_min", c_dp), ("fct_ttf_max", c_dp), ("fct_plus", c_dp), ("fct_minus", c_dp),
                ("tvert_max", c_dp), ("tvert_min", c_dp), ("AUX", c_dp), ("nboundary_lay", c_ip),
                ("tra_advhoriz", c_dp), ("tra_advvert", c_dp), ("dvd_trflx_hor", c_dp), ("dvd_trflx_ver", c_dp)]


class OraRank(C.Structure):
    _fields_ = [("mesh", OraMesh), ("work", OraWork),
                ("vel", c_dp), ("w", c_dp), ("wi", c_dp), ("we", c_dp),
                ("use_wsplit", C.c_int), ("ntr", C.c_int),
                ("values", C.POINTER(c_dp)), ("valuesAB", C.POINTER(c_dp)),
                ("edge_up_dn_grad", C.POINTER(c_dp)), ("dttf_h", C.POINTER(c_dp)), ("dttf_v", C.POINTER(c_dp)),
                ("hor", c_ip), ("ver", c_ip), ("lim", c_ip), ("opth", c_dp), ("optv", c_dp),
                ("halo_owner", c_ip), ("halo_owner_idx", c_ip),
                ("tra_advhoriz", C.POINTER(c_dp)), ("tra_advvert", C.POINTER(c_dp)),
                ("dvd_trflx_hor", C.POINTER(c_dp)), ("dvd_trflx_ver", C.POINTER(c_dp))]


def build(fast: bool = False, force: bool = False) -> str:
    """Compile the restatement with gcc (Makefile recipe).  ``fast`` = -O3 -march=native flavour
    for timing on the machine it is built on."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "adv_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_libs = {}


def lib(fast: bool = False):
    if fast not in _libs:
        L = C.CDLL(build(fast))
        L.ora_run_ranks.restype = C.c_double
        L.ora_run_ranks.argtypes = [C.c_int, C.POINTER(OraRank), C.c_double, C.c_int, C.c_int]
        L.ora_do_oce_adv_tra.restype = C.c_int
        _libs[fast] = L
    return _libs[fast]


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


def _np(t) -> np.ndarray:
    if isinstance(t, np.ndarray):
        return np.ascontiguousarray(t)
    return np.ascontiguousarray(t.detach().cpu().numpy())


class OracleRank:
    """All arrays of one rank, kept alive on the Python side, plus the C structs pointing at them."""

    def __init__(self, mesh, state, tracers, nboundary_lay, alias_aux: bool = False, tra_diag: bool = False, dvd: bool = False):
        """tra_diag: tracers%data(:)%ltra_diag = .true. (tra_advhoriz / tra_advvert per tracer); dvd: ldiag_DVD (the first
        two tracers get dvd_trflx_hor / dvd_trflx_ver, oce_adv_tra_driver.F90:263,:395)"""
        m = mesh
        self.mesh_py = m
        L, nl, Nh, N, T, E = m.L, m.nl, m.Nh, m.N, m.T, m.E
        k = self.keep = {}
        for name in ("edges", "edge_tri", "elem2D_nodes", "nod_in_elem2D", "nod_in_elem2D_num",
                     "nlevels", "ulevels", "nlevels_nod2D", "ulevels_nod2D"):
            k[name] = np.ascontiguousarray(getattr(m, name), dtype=np.int32)
        for name in ("edge_cross_dxdy", "edge_dxdy", "elem_cos", "area", "areasvol"):
            k[name] = np.ascontiguousarray(getattr(m, name), dtype=np.float64)
        for name in ("helem", "hnode", "hnode_new", "zbar_3d_n", "Z_3d_n", "zbar_n_bot", "uv", "w", "w_e", "w_i"):
            k[name] = _np(getattr(state, name)).astype(np.float64)
        k["nboundary_lay"] = np.ascontiguousarray(nboundary_lay, dtype=np.int32)
        om = OraMesh(nl=nl, myDim_nod2D=N, eDim_nod2D=m.eDim_nod2D, myDim_elem2D=T, eDim_elem2D=m.eDim_elem2D,
                     myDim_edge2D=E, nod_in_elem_ld=k["nod_in_elem2D"].shape[1])
        for name in ("edges", "edge_tri", "elem2D_nodes", "nod_in_elem2D", "nod_in_elem2D_num",
                     "nlevels", "ulevels", "nlevels_nod2D", "ulevels_nod2D"):
            setattr(om, name, _ip(k[name]))
        for name in ("edge_cross_dxdy", "edge_dxdy", "elem_cos", "area", "areasvol",
                     "helem", "hnode", "hnode_new", "zbar_3d_n", "Z_3d_n", "zbar_n_bot"):
            setattr(om, name, _dp(k[name]))
        self.cmesh = om
        # work arrays (oce_adv_tra_fct_init, oce_adv_tra_fct.F90:35-67)
        for name, shape in (("fct_LO", (Nh, L)), ("adv_flux_hor", (E, L)), ("adv_flux_ver", (N, nl)),
                            ("fct_ttf_min", (Nh, L)), ("fct_ttf_max", (Nh, L)), ("fct_plus", (Nh, L)),
                            ("fct_minus", (Nh, L)), ("tvert_max", (Nh, L)), ("tvert_min", (Nh, L))):
            k[name] = np.zeros(shape)
        self.ntr = len(tracers)
        self.values = [_np(t.values).copy() for t in tracers]
        self.valuesAB = [_np(t.valuesAB).copy() for t in tracers]
        self.grad = [_np(t.edge_up_dn_grad).copy() for t in tracers]
        self.dttf_h = [np.zeros((Nh, L)) for _ in tracers]
        self.dttf_v = [np.zeros((Nh, L)) for _ in tracers]
        self.tra_advhoriz = [np.zeros((Nh, L)) for _ in tracers] if tra_diag else None
        self.tra_advvert = [np.zeros((Nh, L)) for _ in tracers] if tra_diag else None
        self.dvd_trflx_hor = [np.zeros((E, L)) if i < 2 else None for i in range(len(tracers))] if dvd else None
        self.dvd_trflx_ver = [np.zeros((N, nl)) if i < 2 else None for i in range(len(tracers))] if dvd else None
        self.hor = np.array([HOR[t.tra_adv_hor] for t in tracers], np.int32)
        self.ver = np.array([VER[t.tra_adv_ver] for t in tracers], np.int32)
        self.lim = np.array([LIM[t.tra_adv_lim] for t in tracers], np.int32)
        self.opth = np.array([t.tra_adv_ph for t in tracers], np.float64)
        self.optv = np.array([t.tra_adv_pv for t in tracers], np.float64)
        self.alias_aux = alias_aux
        k["AUX"] = np.zeros((E, L, 4))
        wk = OraWork()
        for name in ("fct_LO", "adv_flux_hor", "adv_flux_ver", "fct_ttf_min", "fct_ttf_max", "fct_plus",
                     "fct_minus", "tvert_max", "tvert_min", "AUX"):
            setattr(wk, name, _dp(k[name]))
        wk.nboundary_lay = _ip(k["nboundary_lay"])
        self.cwork = wk
        self.use_wsplit = int(bool(state.use_wsplit))
        self.halo_owner = np.zeros(max(m.eDim_nod2D, 1), np.int32)
        self.halo_owner_idx = np.zeros(max(m.eDim_nod2D, 1), np.int32)

    def fill(self, rk: OraRank):
        k = self.keep
        rk.mesh = self.cmesh
        rk.work = self.cwork
        rk.vel, rk.w, rk.wi, rk.we = _dp(k["uv"]), _dp(k["w"]), _dp(k["w_i"]), _dp(k["w_e"])
        rk.use_wsplit = self.use_wsplit
        rk.ntr = self.ntr
        PA = c_dp * self.ntr
        self._pa = [PA(*[_dp(a) for a in lst]) for lst in (self.values, self.valuesAB, self.grad, self.dttf_h, self.dttf_v)]
        rk.values, rk.valuesAB, rk.edge_up_dn_grad, rk.dttf_h, rk.dttf_v = [C.cast(p, C.POINTER(c_dp)) for p in self._pa]
        rk.hor, rk.ver, rk.lim = _ip(self.hor), _ip(self.ver), _ip(self.lim)
        rk.opth, rk.optv = _dp(self.opth), _dp(self.optv)
        rk.halo_owner, rk.halo_owner_idx = _ip(self.halo_owner), _ip(self.halo_owner_idx)
        self._pd = []
        for name in ("tra_advhoriz", "tra_advvert", "dvd_trflx_hor", "dvd_trflx_ver"):
            lst = getattr(self, name)
            if lst is not None:
                pa = PA(*[(_dp(a) if a is not None else c_dp()) for a in lst])
                self._pd.append(pa)
                setattr(rk, name, C.cast(pa, C.POINTER(c_dp)))


def link_halos(ranks: Sequence[OracleRank]):
    """Pair every rank's recv segment from p with p's send segment to it (same order: ascending
    global id), i.e. what the MPI datatypes of init_mpi_types encode
    (gen_modules_partitioning.F90:416-514)."""
    for r, rk in enumerate(ranks):
        com = rk.mesh_py.com_nod2D
        N = rk.mesh_py.N
        for i, p in enumerate(com.rPE):
            seg = com.rlist[com.rptr[i] - 1:com.rptr[i + 1] - 1]
            pc = ranks[int(p)].mesh_py.com_nod2D
            j = int(np.flatnonzero(pc.sPE == r)[0])
            sseg = pc.slist[pc.sptr[j] - 1:pc.sptr[j + 1] - 1]
            assert seg.size == sseg.size
            rk.halo_owner[seg - N - 1] = int(p)
            rk.halo_owner_idx[seg - N - 1] = sseg


def run(ranks: Sequence[OracleRank], dt: float, nsteps: int = 1, mode: int = 0, fast: bool = False) -> float:
    """Run ``nsteps`` of the path on all ranks (one thread each). mode 0: do_oce_adv_tra only
    (del_ttf_adv* accumulate); mode 1: dwarf iteration with value update + exchange."""
    n = len(ranks)
    if n > 1:
        link_halos(ranks)
    arr = (OraRank * n)()
    for r, rk in enumerate(ranks):
        rk.fill(arr[r])
    return lib(fast).ora_run_ranks(n, arr, float(dt), int(nsteps), int(mode))


# ---- producer of edge_up_dn_grad ---------------------------------------
def _ipk(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


def tracer_gradient_elements(mesh, ttf: np.ndarray) -> np.ndarray:
    """ora_tracer_gradient_elements: ttf (Nh, L) -> tr_xy (T, L, 2), zero where the reference does not write."""
    L_ = lib()
    L_.ora_tracer_gradient_elements.argtypes = [C.c_int, C.c_int, c_ip, c_ip, c_ip, c_dp, c_dp, c_dp]
    L_.ora_tracer_gradient_elements.restype = None
    en, enp = _ipk(mesh.elem2D_nodes)
    nlv, nlvp = _ipk(mesh.nlevels)
    ulv, ulvp = _ipk(mesh.ulevels)
    g = np.ascontiguousarray(mesh.gradient_sca, dtype=np.float64)
    t = np.ascontiguousarray(ttf, dtype=np.float64)
    out = np.zeros((mesh.T + mesh.eDim_elem2D, mesh.L, 2), np.float64)
    L_.ora_tracer_gradient_elements(mesh.nl, mesh.T, enp, nlvp, ulvp, g.ctypes.data_as(c_dp), t.ctypes.data_as(c_dp),
                                    out.ctypes.data_as(c_dp))
    return out


def fill_up_dn_grad(mesh, tr_xy: np.ndarray, edge_up_dn_tri: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
    """ora_fill_up_dn_grad: tr_xy (>=T, L, 2) -> edge_up_dn_grad (E, L, 4); entries the reference leaves alone
    keep the contents of ``out`` (zeros by default, like the freshly allocated array of the reference)."""
    L_ = lib()
    L_.ora_fill_up_dn_grad.argtypes = [C.c_int, C.c_int, c_ip, c_ip, c_ip, C.c_int, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip,
                                       c_dp, c_dp, c_dp]
    L_.ora_fill_up_dn_grad.restype = None
    keep = [_ipk(a) for a in (mesh.edges, edge_up_dn_tri, mesh.nod_in_elem2D, mesh.nod_in_elem2D_num, mesh.nlevels, mesh.ulevels,
                             mesh.nlevels_nod2D, mesh.ulevels_nod2D, mesh.nlevels_nod2D_min, mesh.ulevels_nod2D_max)]
    p = [k[1] for k in keep]
    ea = np.ascontiguousarray(mesh.elem_area, dtype=np.float64)
    tx = np.ascontiguousarray(tr_xy, dtype=np.float64)
    if out is None:
        out = np.zeros((mesh.E, mesh.L, 4), np.float64)
    assert out.flags.c_contiguous and out.dtype == np.float64
    L_.ora_fill_up_dn_grad(mesh.nl, mesh.E, p[0], p[1], p[2], keep[2][0].shape[1], p[3], p[4], p[5], p[6], p[7], p[8], p[9],
                           ea.ctypes.data_as(c_dp), tx.ctypes.data_as(c_dp), out.ctypes.data_as(c_dp))
    return out


def find_up_downwind_triangles(mesh) -> np.ndarray:
    """ora_find_up_downwind_triangles on a single-rank mesh: edge_up_dn_tri (E, 2), 1-based, 0 = none."""
    L_ = lib()
    L_.ora_find_up_downwind_triangles.argtypes = [C.c_int, c_ip, c_ip, c_ip, C.c_int, c_ip, c_dp, C.c_double, c_ip]
    L_.ora_find_up_downwind_triangles.restype = None
    ed, edp = _ipk(mesh.edges)
    en, enp = _ipk(mesh.elem2D_nodes)
    nie, niep = _ipk(mesh.nod_in_elem2D)
    num, nump = _ipk(mesh.nod_in_elem2D_num)
    co = np.ascontiguousarray(mesh.coord_nod2D, dtype=np.float64)
    out = np.zeros((mesh.E, 2), np.int32)
    L_.ora_find_up_downwind_triangles(mesh.E, edp, enp, niep, nie.shape[1], nump, co.ctypes.data_as(c_dp),
                                      float(mesh.cyclic_length), out.ctypes.data_as(c_ip))
    return out


def vert_vel_ale_core(rank: "OracleRank") -> np.ndarray:
    """ora_vert_vel_ale_core on a rank's mesh and uv: Wvel (Nh, nl), completed on the owned nodes."""
    L_ = lib()
    L_.ora_vert_vel_ale_core.argtypes = [C.POINTER(OraMesh), c_dp, c_dp]
    L_.ora_vert_vel_ale_core.restype = None
    m = rank.mesh_py
    out = np.zeros((m.Nh, m.nl), np.float64)
    L_.ora_vert_vel_ale_core(C.byref(rank.cmesh), _dp(rank.keep["uv"]), _dp(out))
    return out


# ---- memory-lean single-rank driver with an exchange callback (bench.py parity leg) ---------------
XCHG = C.CFUNCTYPE(None, C.c_void_p, c_dp, C.c_int)


class LeanOracle:
    """ora_do_oce_adv_tra for one rank, one tracer at a time: the mesh, the state and ONE set of work arrays are
    held (AUX aliases the tracer's edge_up_dn_grad, as the reference does, oce_adv_tra_driver.F90:383-384), so a
    3.0M-node x 70-layer rank needs ~40 words per (node, layer) instead of ~70.  ``exchange(field, nlev)``: the
    rank's exchange_nod3D (None on one rank); ``field`` is the (Nh, nlev) NumPy view of the C array."""

    def __init__(self, mesh, state_np: dict, nboundary_lay, use_wsplit: bool = False, fast: bool = False):
        m = mesh
        self.m = m
        L, nl, Nh, N, E = m.L, m.nl, m.Nh, m.N, m.E
        k = self.keep = {}
        ints = ("edges", "edge_tri", "elem2D_nodes", "nod_in_elem2D", "nod_in_elem2D_num", "nlevels", "ulevels",
                "nlevels_nod2D", "ulevels_nod2D")
        for name in ints:
            k[name] = np.ascontiguousarray(getattr(m, name), dtype=np.int32)
        for name in ("edge_cross_dxdy", "edge_dxdy", "elem_cos", "area", "areasvol"):
            k[name] = np.ascontiguousarray(getattr(m, name), dtype=np.float64)
        for name in ("helem", "hnode", "hnode_new", "zbar_3d_n", "Z_3d_n", "zbar_n_bot", "uv", "w", "w_e", "w_i"):
            k[name] = np.ascontiguousarray(state_np[name], dtype=np.float64)
        k["nboundary_lay"] = np.ascontiguousarray(nboundary_lay, dtype=np.int32)
        om = OraMesh(nl=nl, myDim_nod2D=N, eDim_nod2D=m.eDim_nod2D, myDim_elem2D=m.T, eDim_elem2D=m.eDim_elem2D,
                     myDim_edge2D=E, nod_in_elem_ld=k["nod_in_elem2D"].shape[1])
        for name in ints:
            setattr(om, name, _ip(k[name]))
        for name in ("edge_cross_dxdy", "edge_dxdy", "elem_cos", "area", "areasvol", "helem", "hnode", "hnode_new",
                     "zbar_3d_n", "Z_3d_n", "zbar_n_bot"):
            setattr(om, name, _dp(k[name]))
        self.cmesh = om
        wk = OraWork()
        for name, shape in (("fct_LO", (Nh, L)), ("adv_flux_hor", (E, L)), ("adv_flux_ver", (N, nl)),
                            ("fct_ttf_min", (Nh, L)), ("fct_ttf_max", (Nh, L)), ("fct_plus", (Nh, L)),
                            ("fct_minus", (Nh, L)), ("tvert_max", (Nh, L)), ("tvert_min", (Nh, L))):
            k[name] = np.zeros(shape)
            setattr(wk, name, _dp(k[name]))
        wk.nboundary_lay = _ip(k["nboundary_lay"])
        self.cwork = wk
        self.use_wsplit = int(bool(use_wsplit))
        self.L_ = lib(fast)
        self.L_.ora_do_oce_adv_tra.argtypes = [C.POINTER(OraMesh), C.POINTER(OraWork), C.c_double, c_dp, c_dp, c_dp, c_dp,
                                               C.c_int, c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                               c_dp, c_dp, XCHG, C.c_void_p]

    def run_tracer(self, values, valuesAB, grad, hor: str, ver: str, lim: str, ph: float, pv: float, dt: float,
                   dttf_h=None, dttf_v=None, exchange=None):
        """One do_oce_adv_tra.  ``grad`` (E, L, 4) is CLOBBERED (it doubles as the FCT scratch AUX); may be None for
        UPW1.  Returns (dttf_h, dttf_v), accumulated into the arrays passed in (zeros by default)."""
        m, k = self.m, self.keep
        v = np.ascontiguousarray(values, dtype=np.float64)
        vab = np.ascontiguousarray(valuesAB, dtype=np.float64)
        if grad is None:
            grad = np.zeros((m.E, m.L, 4))
        assert grad.dtype == np.float64 and grad.flags.c_contiguous and grad.shape == (m.E, m.L, 4)
        self.cwork.AUX = _dp(grad)
        dh = np.zeros((m.Nh, m.L)) if dttf_h is None else dttf_h
        dv = np.zeros((m.Nh, m.L)) if dttf_v is None else dttf_v
        Nh = m.Nh

        def _cb(user, p, nlev):
            exchange(np.ctypeslib.as_array(p, shape=(Nh, nlev)), nlev)
        cb = XCHG(_cb) if exchange is not None else C.cast(None, XCHG)
        self.L_.ora_do_oce_adv_tra(C.byref(self.cmesh), C.byref(self.cwork), float(dt), _dp(k["uv"]), _dp(k["w"]),
                                   _dp(k["w_i"]), _dp(k["w_e"]), self.use_wsplit, _dp(v), _dp(vab), _dp(grad),
                                   HOR[hor], VER[ver], LIM.get(lim, 0), float(ph), float(pv), _dp(dh), _dp(dv), cb, None)
        return dh, dv


def compute_cflz_and_split(rank: "OracleRank", dt: float, Wvel: np.ndarray, use_wsplit: bool, wsplit_maxcfl: float):
    """ora_compute_cflz + ora_compute_wvel_split on a rank's mesh: returns (CFL_z, Wvel_e, Wvel_i), each (Nh, nl);
    entries the reference does not write are zero (CFL_z) / keep the zero of the allocation (Wvel_e, Wvel_i)."""
    L_ = lib()
    L_.ora_compute_cflz.argtypes = [C.POINTER(OraMesh), C.c_double, c_dp, c_dp]
    L_.ora_compute_cflz.restype = None
    L_.ora_compute_wvel_split.argtypes = [C.POINTER(OraMesh), C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp]
    L_.ora_compute_wvel_split.restype = None
    m = rank.mesh_py
    W = np.ascontiguousarray(Wvel, dtype=np.float64)
    cfl = np.zeros((m.Nh, m.nl)); we = np.zeros((m.Nh, m.nl)); wi = np.zeros((m.Nh, m.nl))
    L_.ora_compute_cflz(C.byref(rank.cmesh), float(dt), _dp(W), _dp(cfl))
    L_.ora_compute_wvel_split(C.byref(rank.cmesh), int(bool(use_wsplit)), float(wsplit_maxcfl), _dp(W), _dp(cfl), _dp(we), _dp(wi))
    return cfl, we, wi


def vert_vel_ale_zstar(rank: "OracleRank", dt: float, Wvel: np.ndarray, hbar, hbar_old, water_flux):
    """ora_vert_vel_ale_zstar: returns (Wvel, hnode_new) after the zstar correction; the rank's hnode_new is the start."""
    L_ = lib()
    L_.ora_vert_vel_ale_zstar.argtypes = [C.POINTER(OraMesh), C.c_double, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp]
    L_.ora_vert_vel_ale_zstar.restype = None
    m = rank.mesh_py
    W = np.ascontiguousarray(Wvel, dtype=np.float64).copy()
    hn = rank.keep["hnode_new"].copy()
    nmin = np.ascontiguousarray(m.nlevels_nod2D_min, dtype=np.int32)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (hbar, hbar_old, water_flux)]
    L_.ora_vert_vel_ale_zstar(C.byref(rank.cmesh), float(dt), _ip(nmin), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(W), _dp(hn))
    return W, hn


def vert_vel_ale_zlevel(rank: "OracleRank", dt: float, Wvel: np.ndarray, hbar, hbar_old, water_flux, zbar, cfl_z, min_hnode: float, lzstar_lev: int):
    """ora_vert_vel_ale_zlevel: returns (Wvel, hnode_new) after the zlevel correction; the rank's hnode_new is the start."""
    L_ = lib()
    L_.ora_vert_vel_ale_zlevel.argtypes = [C.POINTER(OraMesh), C.c_double, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_int, c_dp, c_dp]
    L_.ora_vert_vel_ale_zlevel.restype = None
    m = rank.mesh_py
    W = np.ascontiguousarray(Wvel, dtype=np.float64).copy()
    hn = rank.keep["hnode_new"].copy()
    nmin = np.ascontiguousarray(m.nlevels_nod2D_min, dtype=np.int32)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (hbar, hbar_old, water_flux, zbar, cfl_z)]
    L_.ora_vert_vel_ale_zlevel(C.byref(rank.cmesh), float(dt), _ip(nmin), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]),
                               float(min_hnode), int(lzstar_lev), _dp(W), _dp(hn))
    return W, hn
