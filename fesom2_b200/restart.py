"""Derived-type binary restarts of the reference: the files an unmodified FESOM2 writes
with ``write_all_bin_restarts`` and the dwarf reads with ``read_all_bin_restarts``
(src/io_restart_derivedtype.F90:29-234) -- ``t_mesh.<rank>``, ``t_partit.<rank>``, ``t_tracer.<rank>``,
``t_dynamics.<rank>`` -- so that the B200 harness can consume real dwarf inputs and produce inputs a dwarf can read.

Format.  Each file is ONE Fortran ``write(unit) <derived type>`` on a sequential unformatted unit.  The types carry
user-defined derived-type I/O (``generic :: write(unformatted) => write_t_mesh`` ..., src/MOD_MESH.F90:173,
src/MOD_PARTIT.F90:118-119, src/MOD_TRACER.F90:107-108, src/MOD_DYN.F90:199-202): the child ``write(unit)``
statements of those procedures do not start records of their own, their items are appended to the parent's record.
The payload is therefore the plain concatenation of the items in the order of the WRITE_T_* routines, wrapped in
the compiler's record markers (4-byte length before and after; gfortran splits a record longer than 2 GiB - 9 into
sub-records and flags the continuation with a negative length).  ``Stream`` strips the markers -- which also
tolerates a runtime that emits one record per child statement -- and hands out the items.

Arrays written by ``write_bin_array`` (src/MOD_WRITE_BINARY_ARRAYS.F90) are preceded by their extents
(default integers), with all extents 0 for an unallocated array; ``write1d_int_static`` is the same for
fixed-size components.  Defaults: integer and logical 4 bytes, WP = real64 (src/oce_modules.F90:17), all
little-endian.  Arrays come back in C order with the Fortran index order reversed, e.g. ``values(nl-1, Nh)`` as
``(Nh, nl-1)`` -- the layout the rest of this package uses.

The field lists below follow the WRITE_T_* procedures line by line (file:line at each schema); they are data,
checked by tests/test_restart_io.py against the reference sources when /root/reference is present.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

# item kinds: "i" default integer, "r" real(WP), "l" default logical, "c20" character(20),
#             "aiN"/"arN" write_bin_array of rank N, "si" write1d_int_static
Schema = List[Tuple[str, str]]

# src/MOD_MESH.F90:179-279 (write_t_mesh)
T_MESH: Schema = [
    ("i", "nod2D"), ("r", "ocean_area"), ("r", "ocean_areawithcav"), ("i", "edge2D"), ("i", "edge2D_in"), ("i", "elem2D"),
    ("ai2", "elem2D_nodes"), ("ai2", "edges"), ("ai2", "edge_tri"), ("ai2", "elem_edges"), ("ar1", "elem_area"),
    ("ar2", "edge_dxdy"), ("ar2", "edge_cross_dxdy"), ("ar1", "elem_cos"), ("ar1", "metric_factor"),
    ("ai2", "elem_neighbors"), ("ai2", "nod_in_elem2D"), ("ar2", "x_corners"), ("ar2", "y_corners"),
    ("ai1", "nod_in_elem2D_num"), ("ar1", "depth"), ("ar2", "gradient_vec"), ("ar2", "gradient_sca"),
    ("ai1", "bc_index_nod2D"), ("i", "nl"), ("ar1", "zbar"), ("ar1", "Z"), ("ar1", "elem_depth"), ("ai1", "ulevels"),
    ("ai1", "ulevels_nod2D"), ("ai1", "ulevels_nod2D_max"), ("ai1", "nlevels"), ("ai1", "nlevels_nod2D"),
    ("ai1", "nlevels_nod2D_min"), ("ar2", "area"), ("ar2", "area_inv"), ("ar2", "areasvol"), ("ar2", "areasvol_inv"),
    ("ar1", "mesh_resolution"), ("ai1", "cavity_flag_n"), ("ai1", "cavity_flag_e"), ("ar1", "cavity_depth"),
    ("ar2", "cavity_nrst_cavlpnt_xyz"), ("i", "ssh_stiff%dim"), ("i", "ssh_stiff%nza"), ("ai1", "ssh_stiff%rowptr"),
    ("ai1", "ssh_stiff%colind"), ("ar1", "ssh_stiff%values"), ("ai1", "ssh_stiff%colind_loc"),
    ("ai1", "ssh_stiff%rowptr_loc"), ("ar1", "lump2d_south"), ("ar1", "lump2d_north"), ("ai1", "ind_south"),
    ("ai1", "ind_north"), ("i", "nn_size"), ("ai1", "nn_num"), ("ai2", "nn_pos"), ("ar2", "hnode"), ("ar2", "hnode_new"),
    ("ar2", "zbar_3d_n"), ("ar2", "Z_3d_n"), ("ar2", "Z_3d_n_ib"), ("ar2", "helem"), ("ar1", "bottom_elem_thickness"),
    ("ar1", "bottom_node_thickness"), ("ar1", "dhe"), ("ar1", "hbar"), ("ar1", "hbar_old"), ("ar1", "zbar_n_bot"),
    ("ar1", "zbar_e_bot"), ("ar1", "zbar_n_srf"), ("ar1", "zbar_e_srf"), ("ar1", "coriolis"), ("ar1", "coriolis_node"),
]

# src/MOD_PARTIT.F90:123-143 (WRITE_T_COM_STRUCT); rPE(32), rptr(33), sPE(32), sptr(32) are fixed-size
# (MAX_NEIGHBOR_PARTITIONS = 32, :15,:20-25)
T_COM_STRUCT: Schema = [("i", "rPEnum"), ("si", "rPE"), ("si", "rptr"), ("ai1", "rlist"), ("i", "sPEnum"), ("si", "sPE"),
                        ("si", "sptr"), ("ai1", "slist"), ("i", "nreq")]
# src/MOD_PARTIT.F90:163-193 (WRITE_T_PARTIT), after the three communicators
T_PARTIT_TAIL: Schema = [("i", "npes"), ("i", "mype"), ("i", "maxPEnum"), ("ai1", "part"), ("i", "myDim_nod2D"),
                         ("i", "eDim_nod2D"), ("ai1", "myList_nod2D"), ("i", "myDim_elem2D"), ("i", "eDim_elem2D"),
                         ("i", "eXDim_elem2D"), ("ai1", "myList_elem2D"), ("i", "myDim_edge2D"), ("i", "eDim_edge2D"),
                         ("ai1", "myList_edge2D"), ("i", "pe_status")]
# src/MOD_TRACER.F90:112-132 (WRITE_T_TRACER_DATA)
T_TRACER_DATA: Schema = [("ar2", "values"), ("ar3", "valuesold"), ("ar2", "valuesAB"), ("l", "smooth_bh_tra"),
                         ("r", "gamma0_tra"), ("r", "gamma1_tra"), ("r", "gamma2_tra"), ("l", "i_vert_diff"),
                         ("c20", "tra_adv_hor"), ("c20", "tra_adv_ver"), ("c20", "tra_adv_lim"), ("r", "tra_adv_ph"),
                         ("r", "tra_adv_pv"), ("i", "ID")]
# src/MOD_TRACER.F90:156-177 (WRITE_T_TRACER_WORK)
T_TRACER_WORK: Schema = [("ar2", "del_ttf"), ("ar2", "del_ttf_advhoriz"), ("ar2", "del_ttf_advvert"), ("ar3", "dvd_trflx_hor"),
                         ("ar3", "dvd_trflx_ver"), ("ar2", "fct_LO"), ("ar2", "adv_flux_hor"), ("ar2", "adv_flux_ver"),
                         ("ar2", "fct_ttf_max"), ("ar2", "fct_ttf_min"), ("ar2", "fct_plus"), ("ar2", "fct_minus"),
                         ("ai1", "nboundary_lay"), ("ai2", "edge_up_dn_tri"), ("ar3", "edge_up_dn_grad")]
# src/MOD_DYN.F90:212-230 (WRITE_T_SOLVERINFO), :260-271 (WRITE_T_DYN_WORK), :290-340 (WRITE_T_DYN)
T_SOLVERINFO: Schema = [("i", "ident"), ("i", "maxiter"), ("i", "restart"), ("i", "fillin"), ("i", "lutype"), ("r", "droptol"),
                        ("r", "soltol"), ("ar1", "rr"), ("ar1", "zz"), ("ar1", "pp"), ("ar1", "App")]
T_DYN_WORK: Schema = [("ar3", "uvnode_rhs"), ("ar2", "u_c"), ("ar2", "v_c"), ("ar2", "u_b"), ("ar2", "v_b")]
T_DYN_HEAD: Schema = [("i", "opt_visc"), ("r", "visc_gamma0"), ("r", "visc_gamma1"), ("r", "visc_gamma2"),
                      ("r", "visc_easybsreturn"), ("l", "use_ivertvisc"), ("i", "momadv_opt"), ("l", "use_freeslip"),
                      ("l", "use_wsplit"), ("r", "wsplit_maxcfl"), ("l", "use_ssh_se_subcycl")]
T_DYN_ARRAYS: Schema = [("ar3", "uv"), ("ar3", "uv_rhs"), ("ar4", "uv_rhsAB"), ("ar3", "uvnode"), ("ar2", "w"), ("ar2", "w_e"),
                        ("ar2", "w_i"), ("ar2", "cfl_z")]
T_DYN_FER: Schema = [("ar2", "fer_w"), ("ar3", "fer_uv")]                         # only if Fer_GM
T_DYN_SE: Schema = [("ar3", "se_uvh"), ("ar2", "se_uvBT_rhs"), ("ar2", "se_uvBT_4AB"), ("ar2", "se_uvBT"), ("ar2", "se_uvBT_theta"),
                    ("ar2", "se_uvBT_mean"), ("ar2", "se_uvBT_12"), ("ar2", "se_uvBT_stab_hvisc"),
                    ("ar1", "se_uvBT_stab_bdrag")]                                # only if use_ssh_se_subcycl

STATIC_LEN = {"rPE": 32, "rptr": 33, "sPE": 32, "sptr": 32}
_SUBREC = (1 << 31) - 9      # gfortran's maximum sub-record payload


class Stream:
    """Payload of a Fortran sequential unformatted file with the record markers stripped."""

    def __init__(self, path: str):
        raw = np.fromfile(path, dtype=np.uint8)
        parts, pos, n = [], 0, raw.size
        while pos < n:
            if pos + 4 > n:
                raise ValueError(f"{path}: truncated record marker at byte {pos}")
            ln = int(raw[pos:pos + 4].view("<i4")[0])
            size = abs(ln)
            end = pos + 4 + size
            if end + 4 > n:
                raise ValueError(f"{path}: record of {size} bytes at byte {pos} runs past the end of the file")
            tail = int(raw[end:end + 4].view("<i4")[0])
            if abs(tail) != size:
                raise ValueError(f"{path}: record markers disagree at byte {pos} ({ln} / {tail}): not a sequential unformatted file")
            parts.append(raw[pos + 4:end])
            pos = end + 4
        self.buf = np.concatenate(parts) if len(parts) != 1 else parts[0]
        self.pos = 0
        self.path = path

    def take(self, nbytes: int) -> np.ndarray:
        if self.pos + nbytes > self.buf.size:
            raise ValueError(f"{self.path}: payload ends after {self.buf.size} bytes, item needs bytes {self.pos}..{self.pos + nbytes}")
        out = self.buf[self.pos:self.pos + nbytes]
        self.pos += nbytes
        return out

    def ints(self, n: int) -> np.ndarray:
        return self.take(4 * n).view("<i4")

    def done(self) -> bool:
        return self.pos == self.buf.size


def _read_items(s: Stream, schema: Schema, out: Dict[str, object], prefix: str = ""):
    for kind, name in schema:
        key = prefix + name
        if kind == "i":
            out[key] = int(s.ints(1)[0])
        elif kind == "l":
            out[key] = bool(s.ints(1)[0] != 0)
        elif kind == "r":
            out[key] = float(s.take(8).view("<f8")[0])
        elif kind.startswith("c"):
            out[key] = bytes(s.take(int(kind[1:]))).decode("ascii", "replace").rstrip()
        elif kind == "si":
            n = int(s.ints(1)[0])
            if n != STATIC_LEN.get(name, n):
                raise ValueError(f"{s.path}: {key} has {n} entries, the reference declares {STATIC_LEN[name]}")
            out[key] = s.ints(n).copy()
        elif kind[0] == "a":
            rank = int(kind[2:])
            dims = [int(x) for x in s.ints(rank)]
            if any(d < 0 for d in dims):
                raise ValueError(f"{s.path}: negative extent {dims} for {key}")
            cnt = int(np.prod(dims, dtype=np.int64))
            if cnt == 0:
                out[key] = None if all(d == 0 for d in dims) else np.zeros(dims[::-1], "<i4" if kind[1] == "i" else "<f8")
                continue
            dt = "<i4" if kind[1] == "i" else "<f8"
            out[key] = s.take(cnt * (4 if kind[1] == "i" else 8)).view(dt).reshape(dims[::-1]).copy()
        else:
            raise ValueError(kind)


def read_t_mesh(path: str) -> Dict[str, object]:
    s, out = Stream(path), {}
    _read_items(s, T_MESH, out)
    if not s.done():
        raise ValueError(f"{path}: {s.buf.size - s.pos} bytes left after t_mesh (different FESOM version?)")
    return out


def read_t_partit(path: str) -> Dict[str, object]:
    s, out = Stream(path), {}
    for com in ("com_nod2D", "com_elem2D", "com_elem2D_full"):                   # src/MOD_PARTIT.F90:171-173
        _read_items(s, T_COM_STRUCT, out, com + "%")
    _read_items(s, T_PARTIT_TAIL, out)
    if not s.done():
        raise ValueError(f"{path}: {s.buf.size - s.pos} bytes left after t_partit")
    return out


def read_t_tracer(path: str) -> Dict[str, object]:
    s, out = Stream(path), {}
    out["num_tracers"] = int(s.ints(1)[0])                                      # src/MOD_TRACER.F90:209
    out["data"] = []
    for _ in range(out["num_tracers"]):
        d: Dict[str, object] = {}
        _read_items(s, T_TRACER_DATA, d)
        out["data"].append(d)
    out["work"] = {}
    _read_items(s, T_TRACER_WORK, out["work"])
    if not s.done():
        raise ValueError(f"{path}: {s.buf.size - s.pos} bytes left after t_tracer")
    return out


def read_t_dynamics(path: str, fer_gm: bool = False) -> Dict[str, object]:
    """``fer_gm``: the run's Fer_GM switch (o_PARAM); it decides whether fer_w / fer_uv follow cfl_z (MOD_DYN.F90:327-330)."""
    s, out = Stream(path), {}
    _read_items(s, T_DYN_HEAD, out)
    _read_items(s, T_SOLVERINFO, out, "solverinfo%")
    _read_items(s, T_DYN_WORK, out, "work%")
    _read_items(s, T_DYN_ARRAYS, out)
    if fer_gm:
        _read_items(s, T_DYN_FER, out)
    if out["use_ssh_se_subcycl"]:
        _read_items(s, T_DYN_SE, out)
    if not s.done():
        raise ValueError(f"{path}: {s.buf.size - s.pos} bytes left after t_dynamics (Fer_GM = {fer_gm}?)")
    return out


# ------------------------------------------------------------------------------------------------ writing
def _write_items(parts: List[bytes], schema: Schema, src: Dict[str, object], prefix: str = ""):
    for kind, name in schema:
        v = src.get(prefix + name)
        if kind == "i":
            parts.append(struct.pack("<i", int(v or 0)))
        elif kind == "l":
            parts.append(struct.pack("<i", 1 if v else 0))
        elif kind == "r":
            parts.append(struct.pack("<d", float(v or 0.0)))
        elif kind.startswith("c"):
            n = int(kind[1:])
            parts.append(str(v or "").encode("ascii")[:n].ljust(n))
        elif kind == "si":
            n = STATIC_LEN[name]
            a = np.zeros(n, "<i4")
            if v is not None:
                a[:len(v)] = np.asarray(v, "<i4")[:n]
            parts.append(struct.pack("<i", n)); parts.append(a.tobytes())
        else:
            rank = int(kind[2:])
            if v is None:
                parts.append(struct.pack("<" + "i" * rank, *([0] * rank)))
                continue
            a = np.ascontiguousarray(v, "<i4" if kind[1] == "i" else "<f8")
            if a.ndim != rank:
                raise ValueError(f"{prefix + name}: rank {a.ndim}, the reference declares {rank}")
            parts.append(struct.pack("<" + "i" * rank, *a.shape[::-1])); parts.append(a.tobytes())


def _write_record(path: str, parts: List[bytes]):
    payload = b"".join(parts)
    with open(path, "wb") as f:
        pos, n = 0, len(payload)
        while True:                     # gfortran sub-records: every piece but the last carries a negative length in front
            size = min(_SUBREC, n - pos)
            last = pos + size >= n
            first = pos == 0
            f.write(struct.pack("<i", size if last else -size))
            f.write(payload[pos:pos + size])
            f.write(struct.pack("<i", size if first else -size))
            pos += size
            if last:
                break


def write_t_mesh(path: str, d: Dict[str, object]):
    parts: List[bytes] = []
    _write_items(parts, T_MESH, d)
    _write_record(path, parts)


def write_t_partit(path: str, d: Dict[str, object]):
    parts: List[bytes] = []
    for com in ("com_nod2D", "com_elem2D", "com_elem2D_full"):
        _write_items(parts, T_COM_STRUCT, d, com + "%")
    _write_items(parts, T_PARTIT_TAIL, d)
    _write_record(path, parts)


def write_t_tracer(path: str, d: Dict[str, object]):
    parts: List[bytes] = [struct.pack("<i", int(d["num_tracers"]))]
    for t in d["data"]:
        _write_items(parts, T_TRACER_DATA, t)
    _write_items(parts, T_TRACER_WORK, d["work"])
    _write_record(path, parts)


def write_t_dynamics(path: str, d: Dict[str, object], fer_gm: bool = False):
    parts: List[bytes] = []
    _write_items(parts, T_DYN_HEAD, d)
    _write_items(parts, T_SOLVERINFO, d, "solverinfo%")
    _write_items(parts, T_DYN_WORK, d, "work%")
    _write_items(parts, T_DYN_ARRAYS, d)
    if fer_gm:
        _write_items(parts, T_DYN_FER, d)
    if d.get("use_ssh_se_subcycl"):
        _write_items(parts, T_DYN_SE, d)
    _write_record(path, parts)


# ------------------------------------------------------------------------------------------------ harness glue
def rank_suffix(mype: int, npes: int) -> str:
    """``mpirank_to_txt`` (src/fortran_utils.F90:45-56): the rank padded to the width of npes."""
    return str(int(mype)).zfill(int(np.log10(float(npes))) + 1)


def _com(d: Dict[str, object], name: str):
    from .mesh import ComStruct
    nr, ns = int(d[name + "%rPEnum"]), int(d[name + "%sPEnum"])
    return ComStruct(rPE=d[name + "%rPE"][:nr].astype(np.int32), rptr=d[name + "%rptr"][:nr + 1].astype(np.int32),
                     rlist=(d[name + "%rlist"] if d[name + "%rlist"] is not None else np.zeros(0, np.int32)).astype(np.int32),
                     sPE=d[name + "%sPE"][:ns].astype(np.int32), sptr=d[name + "%sptr"][:ns + 1].astype(np.int32),
                     slist=(d[name + "%slist"] if d[name + "%slist"] is not None else np.zeros(0, np.int32)).astype(np.int32))


def load_dwarf(path: str, mype: int = 0, npes: int = 1, fer_gm: bool = False):
    """Everything the path needs from a dwarf input directory (``read_all_bin_restarts``): returns
    ``(Mesh, OceanState, [TracerFields], extras)`` ready for driver.AdvB200 -- tensors on the CPU."""
    import torch
    from .fields import OceanState, TracerFields
    from .mesh import Mesh
    sfx = rank_suffix(mype, npes)
    tm = read_t_mesh(os.path.join(path, "t_mesh." + sfx))
    tp = read_t_partit(os.path.join(path, "t_partit." + sfx))
    tt = read_t_tracer(os.path.join(path, "t_tracer." + sfx))
    td = read_t_dynamics(os.path.join(path, "t_dynamics." + sfx), fer_gm=fer_gm)
    N, eN = int(tp["myDim_nod2D"]), int(tp["eDim_nod2D"])
    T, eT, E = int(tp["myDim_elem2D"]), int(tp["eDim_elem2D"]), int(tp["myDim_edge2D"])
    Nh = N + eN

    def i32(a, n=None):
        a = np.ascontiguousarray(a, np.int32)
        return a if n is None else np.ascontiguousarray(a[:n])

    m = Mesh(nl=int(tm["nl"]), myDim_nod2D=N, eDim_nod2D=eN, myDim_elem2D=T, eDim_elem2D=eT, myDim_edge2D=E,
             cyclic_length=2.0 * np.pi, cartesian=False, coord_nod2D=None,
             elem2D_nodes=i32(tm["elem2D_nodes"], T), edges=i32(tm["edges"], E), edge_tri=i32(tm["edge_tri"], E),
             nlevels=i32(tm["nlevels"]), ulevels=i32(tm["ulevels"]), nlevels_nod2D=i32(tm["nlevels_nod2D"]),
             ulevels_nod2D=i32(tm["ulevels_nod2D"]), zbar=np.asarray(tm["zbar"], np.float64))
    m.nod_in_elem2D = i32(tm["nod_in_elem2D"])
    m.nod_in_elem2D_num = i32(tm["nod_in_elem2D_num"])
    for name in ("elem_area", "elem_cos", "edge_dxdy", "edge_cross_dxdy", "gradient_sca", "area", "areasvol"):
        setattr(m, name, np.ascontiguousarray(tm[name], np.float64))
    m.edge_dxdy, m.edge_cross_dxdy = m.edge_dxdy[:E], m.edge_cross_dxdy[:E]
    m.nlevels_nod2D_min, m.ulevels_nod2D_max = i32(tm["nlevels_nod2D_min"]), i32(tm["ulevels_nod2D_max"])
    m.mype, m.npes = int(tp["mype"]), int(tp["npes"])
    m.myList_nod2D, m.myList_elem2D, m.myList_edge2D = i32(tp["myList_nod2D"]), i32(tp["myList_elem2D"]), i32(tp["myList_edge2D"])
    m.com_nod2D = _com(tp, "com_nod2D")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float64))          # noqa: E731
    st = OceanState(uv=t(td["uv"]), w=t(td["w"]), w_e=t(td["w_e"]), w_i=t(td["w_i"]), helem=t(tm["helem"]), hnode=t(tm["hnode"]),
                    hnode_new=t(tm["hnode_new"]), zbar_3d_n=t(tm["zbar_3d_n"]), Z_3d_n=t(tm["Z_3d_n"]),
                    zbar_n_bot=t(tm["zbar_n_bot"]), use_wsplit=bool(td["use_wsplit"]))
    wk = tt["work"]
    trs = [TracerFields(values=t(d["values"]), valuesAB=t(d["valuesAB"]),
                        edge_up_dn_grad=t(wk["edge_up_dn_grad"]) if wk["edge_up_dn_grad"] is not None else None,
                        tra_adv_hor=d["tra_adv_hor"], tra_adv_ver=d["tra_adv_ver"], tra_adv_lim=d["tra_adv_lim"],
                        tra_adv_ph=d["tra_adv_ph"], tra_adv_pv=d["tra_adv_pv"]) for d in tt["data"]]
    extras = dict(nboundary_lay=i32(wk["nboundary_lay"]) if wk["nboundary_lay"] is not None else None,
                  edge_up_dn_tri=i32(wk["edge_up_dn_tri"]) if wk["edge_up_dn_tri"] is not None else None,
                  wsplit_maxcfl=float(td["wsplit_maxcfl"]), com_elem2D_full=_com(tp, "com_elem2D_full"),
                  eXDim_elem2D=int(tp["eXDim_elem2D"]), t_mesh=tm, t_partit=tp, t_tracer=tt, t_dynamics=td)
    return m, st, trs, extras


def dump_dwarf(path: str, mesh, state, tracers, nboundary_lay, edge_up_dn_tri=None, wsplit_maxcfl: float = 1.0):
    """The inverse: write the four files of rank ``mesh.mype`` from harness objects (fields the path does not use are
    written as unallocated), so that a dwarf built from the reference can read what the synthetic harness generates."""
    os.makedirs(path, exist_ok=True)
    m = mesh
    sfx = rank_suffix(m.mype, m.npes)
    npa = lambda x: None if x is None else (x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x))   # noqa: E731
    tm = {k: npa(getattr(m, k)) for k in ("elem2D_nodes", "edges", "edge_tri", "elem_area", "edge_dxdy", "edge_cross_dxdy", "elem_cos",
                                          "nod_in_elem2D", "nod_in_elem2D_num", "gradient_sca", "zbar", "ulevels", "ulevels_nod2D",
                                          "ulevels_nod2D_max", "nlevels", "nlevels_nod2D", "nlevels_nod2D_min", "area", "areasvol")}
    tm.update(nod2D=int(m.Nh), edge2D=int(m.E), elem2D=int(m.T), nl=int(m.nl), Z=np.asarray(m.Z),
              hnode=npa(state.hnode), hnode_new=npa(state.hnode_new), zbar_3d_n=npa(state.zbar_3d_n), Z_3d_n=npa(state.Z_3d_n),
              helem=npa(state.helem), zbar_n_bot=npa(state.zbar_n_bot))
    write_t_mesh(os.path.join(path, "t_mesh." + sfx), tm)
    tp: Dict[str, object] = dict(npes=int(m.npes), mype=int(m.mype), myDim_nod2D=int(m.N), eDim_nod2D=int(m.eDim_nod2D),
                                 myDim_elem2D=int(m.T), eDim_elem2D=int(m.eDim_elem2D), myDim_edge2D=int(m.E),
                                 myList_nod2D=npa(m.myList_nod2D), myList_elem2D=npa(m.myList_elem2D), myList_edge2D=npa(m.myList_edge2D))
    c = m.com_nod2D
    tp.update({"com_nod2D%rPEnum": c.rPEnum, "com_nod2D%rPE": c.rPE, "com_nod2D%rptr": c.rptr, "com_nod2D%rlist": c.rlist if c.rlist.size else None,
               "com_nod2D%sPEnum": c.sPEnum, "com_nod2D%sPE": c.sPE, "com_nod2D%sptr": c.sptr, "com_nod2D%slist": c.slist if c.slist.size else None})
    write_t_partit(os.path.join(path, "t_partit." + sfx), tp)
    data = [dict(values=npa(t.values), valuesAB=npa(t.valuesAB), tra_adv_hor=t.tra_adv_hor, tra_adv_ver=t.tra_adv_ver,
                 tra_adv_lim=t.tra_adv_lim, tra_adv_ph=float(t.tra_adv_ph), tra_adv_pv=float(t.tra_adv_pv), ID=k + 1)
            for k, t in enumerate(tracers)]
    work = dict(nboundary_lay=npa(nboundary_lay), edge_up_dn_tri=npa(edge_up_dn_tri),
                edge_up_dn_grad=npa(tracers[-1].edge_up_dn_grad) if tracers and tracers[-1].edge_up_dn_grad is not None else None)
    write_t_tracer(os.path.join(path, "t_tracer." + sfx), dict(num_tracers=len(data), data=data, work=work))
    td = dict(uv=npa(state.uv), w=npa(state.w), w_e=npa(state.w_e), w_i=npa(state.w_i), use_wsplit=bool(state.use_wsplit),
              wsplit_maxcfl=float(wsplit_maxcfl))
    write_t_dynamics(os.path.join(path, "t_dynamics." + sfx), td)
