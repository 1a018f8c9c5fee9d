"""Host-side mirror of the reference's interface for the tracer-advection path, on top of the
C ABI (include/fesom_adv_b200.h, libfesom_adv_b200.so).

The reference's seam is the Fortran procedure ``do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics,
tracers, partit, mesh)`` (src/oce_adv_tra_driver.F90:46).  ``AdvB200`` keeps that shape: a mesh
(+partit) object creates the context once (``oce_adv_tra_fct_init``), the per-step state carries
``uv,w,w_i,w_e`` and the ALE thicknesses, and ``do_oce_adv_tra`` takes the tracer batch and
accumulates into ``del_ttf_advhoriz/del_ttf_advvert``.  torch is used only to own device memory
and streams; all arithmetic happens in the CUDA library.  There is no CPU fallback: if the
library is missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FESOM_ADV_LIB") or os.path.join(_HERE, "libfesom_adv_b200.so")   # override: tuning builds only

ADV_HOST, ADV_DEVICE = 0, 1
ADV_OK, ADV_EINVAL, ADV_ECUDA, ADV_ESCHEME, ADV_ENCCL, ADV_ESTATE = 0, -1, -2, -3, -4, -5   # include/fesom_adv_b200.h:32-37

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class AdvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fesom_adv_b200 error {code}: {msg}")
        self.code = code


class MeshDesc(C.Structure):
    _fields_ = [("nl", C.c_int32), ("myDim_nod2D", C.c_int32), ("eDim_nod2D", C.c_int32),
                ("myDim_elem2D", C.c_int32), ("eDim_elem2D", C.c_int32), ("myDim_edge2D", C.c_int32),
                ("nod_in_elem2D_ld", C.c_int32),
                ("edges", c_ip), ("edge_tri", c_ip), ("elem2D_nodes", c_ip), ("nod_in_elem2D", c_ip),
                ("nod_in_elem2D_num", c_ip), ("nlevels", c_ip), ("ulevels", c_ip),
                ("nlevels_nod2D", c_ip), ("ulevels_nod2D", c_ip),
                ("edge_cross_dxdy", c_dp), ("edge_dxdy", c_dp), ("elem_cos", c_dp),
                ("area", c_dp), ("areasvol", c_dp), ("nboundary_lay", c_ip),
                ("mype", C.c_int32), ("npes", C.c_int32),
                ("rPEnum", C.c_int32), ("rPE", c_ip), ("rptr", c_ip), ("rlist", c_ip),
                ("sPEnum", C.c_int32), ("sPE", c_ip), ("sptr", c_ip), ("slist", c_ip)]


class StateDesc(C.Structure):
    _fields_ = [("uv", c_dp), ("w", c_dp), ("w_e", c_dp), ("w_i", c_dp), ("helem", c_dp),
                ("hnode", c_dp), ("hnode_new", c_dp), ("zbar_3d_n", c_dp), ("Z_3d_n", c_dp),
                ("zbar_n_bot", c_dp), ("use_wsplit", C.c_int32)]


class TracerDesc(C.Structure):
    _fields_ = [("values", c_dp), ("valuesAB", c_dp), ("edge_up_dn_grad", c_dp),
                ("del_ttf_advhoriz", c_dp), ("del_ttf_advvert", c_dp),
                ("tra_adv_hor", C.c_char_p), ("tra_adv_ver", C.c_char_p), ("tra_adv_lim", C.c_char_p),
                ("tra_adv_ph", C.c_double), ("tra_adv_pv", C.c_double),
                ("tra_advhoriz", c_dp), ("tra_advvert", c_dp), ("dvd_trflx_hor", c_dp), ("dvd_trflx_ver", c_dp)]


class ZstarDesc(C.Structure):
    _fields_ = [("hbar", c_dp), ("hbar_old", c_dp), ("water_flux", c_dp), ("nlevels_nod2D_min", c_ip), ("hnode_new", c_dp)]


class ZlevelDesc(C.Structure):
    _fields_ = [("hbar", c_dp), ("hbar_old", c_dp), ("water_flux", c_dp), ("nlevels_nod2D_min", c_ip), ("hnode_new", c_dp),
                ("zbar", c_dp), ("min_hnode", C.c_double), ("lzstar_lev", C.c_int32)]


class GradientMeshDesc(C.Structure):
    _fields_ = [("n_elem", C.c_int32), ("n_nod_in_elem", C.c_int32), ("nod_in_elem2D_ld", C.c_int32),
                ("nod_in_elem2D", c_ip), ("nod_in_elem2D_num", c_ip), ("nlevels", c_ip), ("ulevels", c_ip),
                ("edge_up_dn_tri", c_ip), ("nlevels_nod2D_min", c_ip), ("ulevels_nod2D_max", c_ip),
                ("gradient_sca", c_dp), ("elem_area", c_dp),
                ("rPEnum", C.c_int32), ("rPE", c_ip), ("rptr", c_ip), ("rlist", c_ip),
                ("sPEnum", C.c_int32), ("sPE", c_ip), ("sptr", c_ip), ("slist", c_ip)]


EXPORTS = ["adv_ctx_create", "adv_ctx_destroy", "adv_last_error", "adv_comm_unique_id",
           "adv_ctx_comm_init", "adv_ctx_comm_init_local", "adv_exchange_elem", "adv_ctx_halo_stats",
           "adv_ctx_wait_for", "adv_ctx_signal", "adv_vert_vel_ale", "adv_vert_vel_ale_zstar", "adv_vert_vel_ale_zlevel", "adv_ctx_set_state_step", "adv_ctx_set_host_register", "adv_ctx_set_state", "adv_do_oce_adv_tra", "adv_do_oce_adv_tra_async",
           "adv_ctx_synchronize", "adv_exchange_nod", "adv_update_values", "adv_init_tracers_AB",
           "adv_ctx_set_gradient_mesh", "adv_tracer_gradient_elements", "adv_fill_up_dn_grad", "adv_ctx_get_work",
           "adv_ctx_launch_count", "adv_ctx_stream", "adv_ctx_last_elapsed_ms",
           "adv_ctx_set_profiling", "adv_ctx_phase_ms", "adv_selftest_div"]

_lib = None


def load_library():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                                    "There is no CPU fallback for this path.")
        L = C.CDLL(LIB_PATH)
        L.adv_last_error.restype = C.c_char_p
        L.adv_ctx_launch_count.restype = C.c_int64
        L.adv_ctx_launch_count.argtypes = [C.c_void_p]
        L.adv_ctx_stream.restype = C.c_void_p
        L.adv_ctx_stream.argtypes = [C.c_void_p]
        L.adv_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(MeshDesc), C.c_int, C.c_int]
        L.adv_ctx_destroy.argtypes = [C.c_void_p]
        L.adv_comm_unique_id.argtypes = [C.c_char_p]
        L.adv_ctx_comm_init.argtypes = [C.c_void_p, C.c_char_p]
        L.adv_ctx_set_state.argtypes = [C.c_void_p, C.POINTER(StateDesc), C.c_int]
        L.adv_vert_vel_ale.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp]
        L.adv_vert_vel_ale_zstar.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.POINTER(ZstarDesc), c_dp, c_dp, c_dp, c_dp]
        L.adv_vert_vel_ale_zlevel.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.POINTER(ZlevelDesc), c_dp, c_dp, c_dp, c_dp]
        L.adv_ctx_set_state_step.argtypes = [C.c_void_p, C.POINTER(StateDesc), C.c_int, C.c_int64]
        L.adv_ctx_set_host_register.argtypes = [C.c_void_p, C.c_int]
        L.adv_do_oce_adv_tra.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(TracerDesc), C.c_int]
        L.adv_do_oce_adv_tra_async.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(TracerDesc)]
        L.adv_ctx_synchronize.argtypes = [C.c_void_p]
        L.adv_ctx_wait_for.argtypes = [C.c_void_p, C.c_void_p]
        L.adv_ctx_signal.argtypes = [C.c_void_p, C.c_void_p]
        L.adv_exchange_nod.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_dp), C.c_int]
        L.adv_exchange_elem.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_dp), C.c_int]
        L.adv_ctx_comm_init_local.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.adv_ctx_halo_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.adv_update_values.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_dp), C.POINTER(c_dp), C.POINTER(c_dp)]
        L.adv_init_tracers_AB.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double] + [C.POINTER(c_dp)] * 6
        L.adv_ctx_set_gradient_mesh.argtypes = [C.c_void_p, C.POINTER(GradientMeshDesc)]
        L.adv_tracer_gradient_elements.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_dp), C.POINTER(c_dp)]
        L.adv_fill_up_dn_grad.argtypes = [C.c_void_p, C.c_int, C.POINTER(c_dp), C.POINTER(c_dp)]
        L.adv_ctx_get_work.argtypes = [C.c_void_p, C.c_char_p, C.c_int, c_dp]
        L.adv_ctx_last_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.adv_ctx_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.adv_ctx_phase_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise AdvError(rc, load_library().adv_last_error().decode())


def selftest_div(count: int, seed: int, mode: int) -> int:
    """number of mismatches between the library's reciprocal-based division and IEEE `/`"""
    L = load_library()
    bad = C.c_uint64(0)
    L.adv_selftest_div.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    _check(L.adv_selftest_div(int(count), int(seed), int(mode), C.byref(bad)))
    return int(bad.value)


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load_library().adv_comm_unique_id(buf))
    return buf.raw


def _ptr(t) -> c_dp:
    """Raw pointer of a float64 contiguous torch tensor (cpu or cuda) or numpy array."""
    if t is None:
        return c_dp()
    if isinstance(t, np.ndarray):
        assert t.dtype == np.float64 and t.flags.c_contiguous
        return t.ctypes.data_as(c_dp)
    assert t.dtype.is_floating_point and t.element_size() == 8 and t.is_contiguous()
    return C.cast(t.data_ptr(), c_dp)


def _where(t) -> int:
    if isinstance(t, np.ndarray):
        return ADV_HOST
    return ADV_DEVICE if t.is_cuda else ADV_HOST


def comm_init_local(ctxs: Sequence["AdvB200"]):
    """adv_ctx_comm_init_local: link the contexts of this process (ctxs[r] built from rank r's local mesh) into an
    in-process communicator; afterwards every context must be driven by its own host thread."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    _check(load_library().adv_ctx_comm_init_local(arr, len(ctxs)))


class AdvB200:
    """One rank's advection context (mesh + partit resident on one GPU)."""

    def __init__(self, mesh, nboundary_lay: Optional[np.ndarray] = None, device: int = 0, max_tracers: int = 2):
        self.lib = load_library()
        self.mesh = mesh
        m = mesh
        k = self._keep = {}
        d = MeshDesc(nl=m.nl, myDim_nod2D=m.N, eDim_nod2D=m.eDim_nod2D, myDim_elem2D=m.T,
                     eDim_elem2D=m.eDim_elem2D, myDim_edge2D=m.E)
        for name in ("edges", "edge_tri", "elem2D_nodes", "nod_in_elem2D", "nod_in_elem2D_num",
                     "nlevels", "ulevels", "nlevels_nod2D", "ulevels_nod2D"):
            k[name] = np.ascontiguousarray(getattr(m, name), dtype=np.int32)
            setattr(d, name, k[name].ctypes.data_as(c_ip))
        d.nod_in_elem2D_ld = k["nod_in_elem2D"].shape[1]
        for name in ("edge_cross_dxdy", "edge_dxdy", "elem_cos", "area", "areasvol"):
            k[name] = np.ascontiguousarray(getattr(m, name), dtype=np.float64)
            setattr(d, name, k[name].ctypes.data_as(c_dp))
        if nboundary_lay is not None:
            k["nb"] = np.ascontiguousarray(nboundary_lay, dtype=np.int32)
            d.nboundary_lay = k["nb"].ctypes.data_as(c_ip)
        com = m.com_nod2D
        d.mype, d.npes = m.mype, m.npes
        for name in ("rPE", "rptr", "rlist", "sPE", "sptr", "slist"):
            k[name] = np.ascontiguousarray(getattr(com, name), dtype=np.int32)
            setattr(d, name, k[name].ctypes.data_as(c_ip))
        d.rPEnum, d.sPEnum = com.rPEnum, com.sPEnum
        self.max_tracers = max_tracers
        self.device = device
        h = C.c_void_p()
        _check(self.lib.adv_ctx_create(C.byref(h), C.byref(d), device, max_tracers))
        self.h = h
        self._state = None

    # -- life cycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.adv_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, uid: bytes):
        _check(self.lib.adv_ctx_comm_init(self.h, uid))

    # -- stream ordering against torch ---------------------------------------------------------
    def _torch_stream(self):
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream if torch.cuda.is_available() else None

    def _after_torch(self):
        """device tensors handed to the library may still be in flight on torch's current stream"""
        ts = self._torch_stream()
        if ts is not None:
            _check(self.lib.adv_ctx_wait_for(self.h, C.c_void_p(ts)))

    def _before_torch(self):
        """torch work submitted after an asynchronous library call sees its results"""
        ts = self._torch_stream()
        if ts is not None:
            _check(self.lib.adv_ctx_signal(self.h, C.c_void_p(ts)))

    # -- per step -----------------------------------------------------------------------------
    def set_host_register(self, on: bool):
        _check(self.lib.adv_ctx_set_host_register(self.h, int(on)))

    def set_state(self, st, step: Optional[int] = None):
        """``st``: fields.OceanState with torch tensors (all cpu or all cuda).  ``step``: the model's step counter --
        repeated calls with the same step and the same arrays are no-ops (adv_ctx_set_state_step)."""
        self._state = st   # keep alive: device pointers are used in place
        if _where(st.uv) == ADV_DEVICE:
            self._after_torch()
        sd = StateDesc(uv=_ptr(st.uv), w=_ptr(st.w), w_e=_ptr(st.w_e), w_i=_ptr(st.w_i), helem=_ptr(st.helem),
                       hnode=_ptr(st.hnode), hnode_new=_ptr(st.hnode_new), zbar_3d_n=_ptr(st.zbar_3d_n),
                       Z_3d_n=_ptr(st.Z_3d_n), zbar_n_bot=_ptr(st.zbar_n_bot), use_wsplit=int(bool(st.use_wsplit)))
        if step is None:
            _check(self.lib.adv_ctx_set_state(self.h, C.byref(sd), _where(st.uv)))
        else:
            _check(self.lib.adv_ctx_set_state_step(self.h, C.byref(sd), _where(st.uv), int(step)))

    def _descs(self, tracers, dttf_h, dttf_v, tra_advhoriz=None, tra_advvert=None, dvd_trflx_hor=None, dvd_trflx_ver=None):
        n = len(tracers)
        arr = (TracerDesc * n)()
        keep = []
        for i, t in enumerate(tracers):
            hs, vs, ls = t.tra_adv_hor.encode(), t.tra_adv_ver.encode(), t.tra_adv_lim.encode()
            keep += [hs, vs, ls]
            arr[i] = TracerDesc(values=_ptr(t.values), valuesAB=_ptr(t.valuesAB), edge_up_dn_grad=_ptr(t.edge_up_dn_grad),
                                del_ttf_advhoriz=_ptr(dttf_h[i]), del_ttf_advvert=_ptr(dttf_v[i]),
                                tra_adv_hor=hs, tra_adv_ver=vs, tra_adv_lim=ls,
                                tra_adv_ph=float(t.tra_adv_ph), tra_adv_pv=float(t.tra_adv_pv),
                                tra_advhoriz=_ptr(tra_advhoriz[i]) if tra_advhoriz else c_dp(),
                                tra_advvert=_ptr(tra_advvert[i]) if tra_advvert else c_dp(),
                                dvd_trflx_hor=_ptr(dvd_trflx_hor[i]) if dvd_trflx_hor else c_dp(),
                                dvd_trflx_ver=_ptr(dvd_trflx_ver[i]) if dvd_trflx_ver else c_dp())
        return arr, keep

    def do_oce_adv_tra(self, dt: float, tracers: Sequence, dttf_h: Sequence, dttf_v: Sequence, sync: bool = True,
                       tra_advhoriz: Optional[Sequence] = None, tra_advvert: Optional[Sequence] = None,
                       dvd_trflx_hor: Optional[Sequence] = None, dvd_trflx_ver: Optional[Sequence] = None):
        """Batched ``do_oce_adv_tra``.  Tensors on cuda are used in place; cpu tensors / numpy arrays
        are copied to the device and the tendencies copied back (blocking).  ``tra_advhoriz`` / ``tra_advvert``: per-tracer
        (Nh, L) arrays (entries may be None) that receive the ltra_diag diagnostics of the reference
        (src/oce_adv_tra_driver.F90:221-229, :307-318, :464-488); ``dvd_trflx_hor`` (E, L) / ``dvd_trflx_ver`` (N, nl) the
        ldiag_DVD fluxes (:263-296, :395-458)."""
        arr, keep = self._descs(tracers, dttf_h, dttf_v, tra_advhoriz, tra_advvert, dvd_trflx_hor, dvd_trflx_ver)
        where = _where(tracers[0].values)
        if where == ADV_DEVICE:
            self._after_torch()
        if where == ADV_DEVICE and not sync:
            _check(self.lib.adv_do_oce_adv_tra_async(self.h, float(dt), len(tracers), arr))
            self._before_torch()
        else:
            _check(self.lib.adv_do_oce_adv_tra(self.h, float(dt), len(tracers), arr, where))

    def synchronize(self):
        _check(self.lib.adv_ctx_synchronize(self.h))

    def exchange_nod(self, fields: Sequence, nlev: int):
        PA = c_dp * len(fields)
        self._after_torch()
        _check(self.lib.adv_exchange_nod(self.h, len(fields), PA(*[_ptr(f) for f in fields]), int(nlev)))
        self._before_torch()

    def vert_vel_ale(self, dt: float, use_wsplit: bool, wsplit_maxcfl: float, w, w_e, w_i, cfl_z=None):
        """``vert_vel_ale`` continuity part (linfs) + exchange_nod(Wvel) + ``compute_CFLz`` + ``compute_Wvel_split``
        (src/oce_ale.F90:2164-2310, :2654, :2906-3049) from the state's uv / helem / hnode_new; (Nh, nl) device tensors."""
        self._after_torch()
        _check(self.lib.adv_vert_vel_ale(self.h, float(dt), int(bool(use_wsplit)), float(wsplit_maxcfl), _ptr(w), _ptr(w_e),
                                         _ptr(w_i), _ptr(cfl_z)))
        self._before_torch()

    def vert_vel_ale_zstar(self, dt: float, use_wsplit: bool, wsplit_maxcfl: float, hbar, hbar_old, water_flux, nlevels_nod2D_min,
                           hnode_new, w, w_e, w_i, cfl_z=None):
        """``vert_vel_ale`` for which_ALE = 'zstar' (src/oce_ale.F90:2164-2310, :2539-2603, :2654-2666): device tensors;
        ``nlevels_nod2D_min`` an int32 device tensor, ``hnode_new`` updated in place."""
        self._after_torch()
        z = ZstarDesc(hbar=_ptr(hbar), hbar_old=_ptr(hbar_old), water_flux=_ptr(water_flux),
                      nlevels_nod2D_min=C.cast(nlevels_nod2D_min.data_ptr(), c_ip), hnode_new=_ptr(hnode_new))
        _check(self.lib.adv_vert_vel_ale_zstar(self.h, float(dt), int(bool(use_wsplit)), float(wsplit_maxcfl), C.byref(z),
                                               _ptr(w), _ptr(w_e), _ptr(w_i), _ptr(cfl_z)))
        self._before_torch()

    def vert_vel_ale_zlevel(self, dt: float, use_wsplit: bool, wsplit_maxcfl: float, hbar, hbar_old, water_flux, nlevels_nod2D_min,
                            hnode_new, zbar, min_hnode: float, lzstar_lev: int, w, w_e, w_i, cfl_z):
        """``vert_vel_ale`` for which_ALE = 'zlevel' (src/oce_ale.F90:2164-2310, :2336-2538, :2654-2666): device tensors;
        ``zbar`` the (nl) rest interfaces, ``cfl_z`` in/out (the previous step's CFL_z on entry), ``hnode_new`` updated in place."""
        self._after_torch()
        z = ZlevelDesc(hbar=_ptr(hbar), hbar_old=_ptr(hbar_old), water_flux=_ptr(water_flux),
                       nlevels_nod2D_min=C.cast(nlevels_nod2D_min.data_ptr(), c_ip), hnode_new=_ptr(hnode_new),
                       zbar=_ptr(zbar), min_hnode=float(min_hnode), lzstar_lev=int(lzstar_lev))
        _check(self.lib.adv_vert_vel_ale_zlevel(self.h, float(dt), int(bool(use_wsplit)), float(wsplit_maxcfl), C.byref(z),
                                                _ptr(w), _ptr(w_e), _ptr(w_i), _ptr(cfl_z)))
        self._before_torch()

    def update_values(self, values: Sequence, dttf_h: Sequence, dttf_v: Sequence):
        n = len(values)
        PA = c_dp * n
        self._after_torch()
        _check(self.lib.adv_update_values(self.h, n, PA(*[_ptr(v) for v in values]),
                                          PA(*[_ptr(v) for v in dttf_h]), PA(*[_ptr(v) for v in dttf_v])))
        self._before_torch()

    def init_tracers_AB(self, values: Sequence, valuesold: Sequence, valuesAB: Sequence, ab_order: int = 2,
                        epsilon: float = 0.1, del_ttf: Optional[Sequence] = None, dttf_h: Optional[Sequence] = None,
                        dttf_v: Optional[Sequence] = None):
        """``init_tracers_AB`` (src/oce_tracer_mod.F90:13-123) for a batch: AB2/AB3 extrapolation into
        valuesAB, rotation of valuesold, zeroing of the tendencies.  Device tensors, asynchronous."""
        n = len(values)
        PA = c_dp * n

        def arr(lst):
            return PA(*[_ptr(v) for v in lst]) if lst is not None else None
        self._after_torch()
        _check(self.lib.adv_init_tracers_AB(self.h, n, int(ab_order), float(epsilon), arr(values), arr(valuesold),
                                            arr(valuesAB), arr(del_ttf), arr(dttf_h), arr(dttf_v)))
        self._before_torch()

    # -- producer of edge_up_dn_grad ---------------------------------------------
    def set_gradient_mesh(self, edge_up_dn_tri: Optional[np.ndarray] = None, gmesh=None):
        """Static inputs of tracer_gradient_elements / fill_up_dn_grad (t_tracer_work%edge_up_dn_tri and the
        t_mesh fields gradient_sca, elem_area, nlevels_nod2D_min, ulevels_nod2D_max, nod_in_elem2D).  One rank:
        ``edge_up_dn_tri`` is enough; a rank of a partitioned mesh passes ``gmesh`` = mesh.gradient_mesh(...),
        which carries the element halo (eDim + eXDim) and com_elem2D_full."""
        m, k = self.mesh, self._keep
        if gmesh is None:
            from .mesh import gradient_mesh
            gmesh = gradient_mesh(m, edge_up_dn_tri)
        self.gmesh = gmesh
        d = GradientMeshDesc(n_elem=int(gmesh.n_elem), n_nod_in_elem=int(gmesh.nod_in_elem2D.shape[0]),
                             nod_in_elem2D_ld=int(gmesh.nod_in_elem2D.shape[1]))
        for name, src in (("nod_in_elem2D", gmesh.nod_in_elem2D), ("nod_in_elem2D_num", gmesh.nod_in_elem2D_num),
                          ("nlevels", gmesh.nlevels), ("ulevels", gmesh.ulevels), ("edge_up_dn_tri", gmesh.edge_up_dn_tri),
                          ("nlevels_nod2D_min", m.nlevels_nod2D_min), ("ulevels_nod2D_max", m.ulevels_nod2D_max)):
            k["g_" + name] = np.ascontiguousarray(src, dtype=np.int32)
            setattr(d, name, k["g_" + name].ctypes.data_as(c_ip))
        for name, src in (("gradient_sca", m.gradient_sca), ("elem_area", gmesh.elem_area)):
            k["g_" + name] = np.ascontiguousarray(src, dtype=np.float64)
            setattr(d, name, k["g_" + name].ctypes.data_as(c_dp))
        com = gmesh.com_elem2D_full
        for name in ("rPE", "rptr", "rlist", "sPE", "sptr", "slist"):
            k["ge_" + name] = np.ascontiguousarray(getattr(com, name), dtype=np.int32)
            setattr(d, name, k["ge_" + name].ctypes.data_as(c_ip))
        d.rPEnum, d.sPEnum = com.rPEnum, com.sPEnum
        _check(self.lib.adv_ctx_set_gradient_mesh(self.h, C.byref(d)))

    def exchange_elem(self, fields: Sequence, nwords: int):
        """``exchange_elem`` over com_elem2D_full for element fields with ``nwords`` doubles per element column."""
        PA = c_dp * len(fields)
        self._after_torch()
        _check(self.lib.adv_exchange_elem(self.h, len(fields), PA(*[_ptr(f) for f in fields]), int(nwords)))
        self._before_torch()

    def halo_stats(self):
        """(bytes sent, comm_ms[2], exposed_ms[2]) of the last multi-rank FCT call."""
        b = C.c_int64(0)
        cm, ex = (C.c_float * 2)(), (C.c_float * 2)()
        _check(self.lib.adv_ctx_halo_stats(self.h, C.byref(b), cm, ex))
        return int(b.value), [float(x) for x in cm], [float(x) for x in ex]

    def tracer_gradient_elements(self, ttf: Sequence, tr_xy: Sequence):
        """``tracer_gradient_elements`` (src/oce_tracer_mod.F90:146-188): ttf[i] (Nh, L) -> tr_xy[i] (n_elem, L, 2)."""
        PA = c_dp * len(ttf)
        self._after_torch()
        _check(self.lib.adv_tracer_gradient_elements(self.h, len(ttf), PA(*[_ptr(t) for t in ttf]), PA(*[_ptr(t) for t in tr_xy])))
        self._before_torch()

    def fill_up_dn_grad(self, tr_xy: Sequence, edge_up_dn_grad: Sequence):
        """``fill_up_dn_grad`` (src/oce_muscl_adv.F90:356-525): tr_xy[i] -> edge_up_dn_grad[i] (E, L, 4)."""
        PA = c_dp * len(tr_xy)
        self._after_torch()
        _check(self.lib.adv_fill_up_dn_grad(self.h, len(tr_xy), PA(*[_ptr(t) for t in tr_xy]), PA(*[_ptr(t) for t in edge_up_dn_grad])))
        self._before_torch()

    # -- introspection ------------------------------------------------------------------------
    def get_work(self, name: str, slot: int = 0) -> np.ndarray:
        m = self.mesh
        shape = {"fct_LO": (m.Nh, m.L), "fct_plus": (m.Nh, m.L), "fct_minus": (m.Nh, m.L),
                 "adv_flux_hor": (m.E, m.L), "adv_flux_ver": (m.N, m.nl), "edge_volflux": (m.E, m.L)}[name]
        out = np.empty(shape, np.float64)
        _check(self.lib.adv_ctx_get_work(self.h, name.encode(), slot, out.ctypes.data_as(c_dp)))
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.adv_ctx_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.adv_ctx_stream(self.h) or 0)

    def last_elapsed_ms(self) -> float:
        ms = C.c_float()
        _check(self.lib.adv_ctx_last_elapsed_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def set_profiling(self, on: bool):
        _check(self.lib.adv_ctx_set_profiling(self.h, int(on)))

    def phase_ms(self) -> List[float]:
        arr = (C.c_float * 8)()
        _check(self.lib.adv_ctx_phase_ms(self.h, arr))
        return [float(x) for x in arr]
