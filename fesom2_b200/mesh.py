"""Host-side mesh model for the tracer-advection path.

Mirrors the slice of the reference's ``t_mesh`` / ``t_partit`` that
``do_oce_adv_tra`` reads (reference: src/MOD_MESH.F90:22-175,
src/MOD_PARTIT.F90:18-120, src/associate_mesh_ass.h:9-79).  This module is host
logic only (numpy): it reads the reference's ASCII mesh fixtures, generates the
synthetic lon-lat meshes of BASELINE.json configs 3-5, derives the geometric
arrays exactly as Appendix B of SURVEY.md describes (reference:
src/oce_mesh.F90:2140-2349, :2445-2670) and cuts per-rank local meshes with halo
lists the way ``communication_nodn`` / ``mymesh`` do (reference:
src/gen_comm.F90:8-220, :529-658).

Conventions
-----------
* Index *values* stored in the arrays are 1-based, exactly what the Fortran host
  hands to the C ABI (``edge_tri[:,1] <= 0`` marks a boundary edge).
* numpy shapes are the transposes of the Fortran shapes, C-contiguous, i.e. a
  Fortran ``values(nz, n)`` is a numpy ``values[n, nz]``: identical bytes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

R_EARTH = 6367500.0            # src/oce_modules.F90:29
RAD = np.pi / 180.0            # src/oce_modules.F90:25

# Standard 48-interface vertical grid of the `pi` test mesh (depths in metres).
ZBAR_48 = -np.array([
    0, 5, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 115, 135, 160, 190, 230, 280, 340, 410, 490,
    580, 680, 790, 910, 1040, 1180, 1330, 1500, 1700, 1920, 2150, 2400, 2650, 2900, 3150, 3400,
    3650, 3900, 4150, 4400, 4650, 4900, 5150, 5400, 5650, 6000, 6250], dtype=np.float64)


def zbar_levels(nl: int) -> np.ndarray:
    """nl interface depths (<= 0).  nl == 48 is the pi grid; other counts are obtained by
    monotone interpolation of that grid in index space (config 4: nl = 71)."""
    if nl == 48:
        return ZBAR_48.copy()
    s = np.linspace(0.0, 47.0, nl)
    z = np.interp(s, np.arange(48.0), ZBAR_48)
    z[0] = 0.0
    return z


def _trim_cyclic(d: np.ndarray, cyc: float) -> np.ndarray:
    """src/gen_modules_rotate_grid.F90:204-211."""
    d = np.where(d > cyc / 2.0, d - cyc, d)
    d = np.where(d < -cyc / 2.0, d + cyc, d)
    return d


@dataclass
class ComStruct:
    """``com_struct`` of src/MOD_PARTIT.F90:18-33 (lists hold 1-based local ids, ptr are 1-based)."""
    rPE: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rptr: np.ndarray = field(default_factory=lambda: np.ones(1, np.int32))
    rlist: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    sPE: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    sptr: np.ndarray = field(default_factory=lambda: np.ones(1, np.int32))
    slist: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))

    @property
    def rPEnum(self) -> int:
        return int(self.rPE.size)

    @property
    def sPEnum(self) -> int:
        return int(self.sPE.size)


@dataclass
class Mesh:
    """The arrays of ``t_mesh``/``t_partit`` the advection path reads, for one rank."""
    nl: int
    myDim_nod2D: int
    eDim_nod2D: int
    myDim_elem2D: int
    eDim_elem2D: int          # vel/helem carry this many extra element columns (unused by the path)
    myDim_edge2D: int
    cyclic_length: float
    cartesian: bool
    coord_nod2D: np.ndarray   # (Nh, 2) radians
    elem2D_nodes: np.ndarray  # (T, 3) int32, 1-based
    edges: np.ndarray         # (E, 2) int32, 1-based
    edge_tri: np.ndarray      # (E, 2) int32, 1-based, <=0 for none
    nlevels: np.ndarray       # (T,)  interfaces per element
    ulevels: np.ndarray       # (T,)
    nlevels_nod2D: np.ndarray  # (Nh,)
    ulevels_nod2D: np.ndarray  # (Nh,)
    zbar: np.ndarray          # (nl,)
    # derived
    nod_in_elem2D: np.ndarray = None      # (N or Nh, maxdeg) int32 1-based, 0 padded
    nod_in_elem2D_num: np.ndarray = None  # (N or Nh,)
    elem_area: np.ndarray = None          # (T,) m^2
    elem_cos: np.ndarray = None           # (T,)
    edge_dxdy: np.ndarray = None          # (E, 2) radians
    edge_cross_dxdy: np.ndarray = None    # (E, 4) metres
    gradient_sca: np.ndarray = None       # (T, 6) 1/m
    area: np.ndarray = None               # (Nh, nl)
    areasvol: np.ndarray = None           # (Nh, nl)
    nlevels_nod2D_min: np.ndarray = None  # (Nh,)
    ulevels_nod2D_max: np.ndarray = None  # (Nh,)
    # partition bookkeeping (global ids, 1-based) -- identity on a 1-rank mesh
    mype: int = 0
    npes: int = 1
    myList_nod2D: np.ndarray = None
    myList_elem2D: np.ndarray = None
    myList_edge2D: np.ndarray = None
    com_nod2D: ComStruct = field(default_factory=ComStruct)

    # ---- sizes ---------------------------------------------------------------------------
    @property
    def L(self) -> int:
        return self.nl - 1

    @property
    def N(self) -> int:
        return self.myDim_nod2D

    @property
    def Nh(self) -> int:
        return self.myDim_nod2D + self.eDim_nod2D

    @property
    def T(self) -> int:
        return self.myDim_elem2D

    @property
    def E(self) -> int:
        return self.myDim_edge2D

    @property
    def Z(self) -> np.ndarray:
        return 0.5 * (self.zbar[:-1] + self.zbar[1:])


# =============================================================================================
# topology helpers
# =============================================================================================
def build_nod_in_elem(elem2D_nodes: np.ndarray, n_nodes: int):
    """nod_in_elem2D(:,n): elements around node n in ascending element order
    (reference: src/oce_mesh.F90:2056-2064).  Returns (table 1-based 0-padded, counts)."""
    T = elem2D_nodes.shape[0]
    nodes = (elem2D_nodes.astype(np.int64) - 1).ravel()          # element-major: ascending elem
    elems = np.repeat(np.arange(T, dtype=np.int64), 3)
    order = np.argsort(nodes, kind="stable")                     # stable keeps ascending elem
    nodes_s, elems_s = nodes[order], elems[order]
    num = np.bincount(nodes_s, minlength=n_nodes).astype(np.int32)
    maxdeg = int(num.max()) if num.size else 0
    start = np.concatenate(([0], np.cumsum(num)[:-1]))
    pos = np.arange(nodes_s.size) - start[nodes_s]
    table = np.zeros((n_nodes, maxdeg), dtype=np.int32)
    table[nodes_s, pos] = elems_s + 1
    return table, num


def build_edges(elem2D_nodes: np.ndarray, coord: np.ndarray, cyc: float):
    """Unique edges with the reference orientation: ``edge_tri(1,e)`` lies to the LEFT of the
    direction edges(1,e)->edges(2,e) (src/MOD_MESH.F90:33-35); boundary edges carry their only
    element first and 0 second.  Interior edges first, boundary edges last, as the reference's
    mesh files are laid out (writer: src/fvom_init.F90:476-690)."""
    en = elem2D_nodes.astype(np.int64) - 1
    T = en.shape[0]
    a = np.concatenate([en[:, 0], en[:, 1], en[:, 2]])
    b = np.concatenate([en[:, 1], en[:, 2], en[:, 0]])
    el = np.tile(np.arange(T, dtype=np.int64), 3)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    nn = int(max(hi.max(), lo.max())) + 1
    key = lo * nn + hi
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    first = np.ones(key_s.size, bool)
    first[1:] = key_s[1:] != key_s[:-1]
    idx = np.flatnonzero(first)
    cnt = np.diff(np.append(idx, key_s.size))
    assert cnt.max() <= 2, "non-manifold mesh"
    e_lo, e_hi = lo[order][idx], hi[order][idx]
    el_a = el[order][idx]
    el_b = np.where(cnt == 2, el[order][np.minimum(idx + 1, key_s.size - 1)], -1)

    # which side of lo->hi is el_a on?  use its centroid relative to lo
    x, y = coord[:, 0], coord[:, 1]

    def side(elm):
        cx = sum(_trim_cyclic(x[en[elm, k]] - x[e_lo], cyc) for k in range(3)) / 3.0
        cy = sum(y[en[elm, k]] - y[e_lo] for k in range(3)) / 3.0
        dx = _trim_cyclic(x[e_hi] - x[e_lo], cyc)
        dy = y[e_hi] - y[e_lo]
        return dx * cy - dy * cx                     # > 0: element left of lo->hi

    a_left = side(el_a) > 0
    interior = el_b >= 0
    n1 = e_lo.copy(); n2 = e_hi.copy()
    t1 = np.where(a_left, el_a, el_b)
    t2 = np.where(a_left, el_b, el_a)
    # boundary edge whose only element sits on the right of lo->hi: flip the direction
    flip = (~interior) & (~a_left)
    n1[flip], n2[flip] = e_hi[flip], e_lo[flip]
    t1[flip], t2[flip] = el_a[flip], -1
    t1[(~interior) & a_left] = el_a[(~interior) & a_left]
    t2[~interior] = -1
    perm = np.concatenate([np.flatnonzero(interior), np.flatnonzero(~interior)])
    edges = np.stack([n1[perm], n2[perm]], 1).astype(np.int32) + 1
    edge_tri = np.stack([t1[perm], t2[perm]], 1).astype(np.int32) + 1      # -1 -> 0
    return edges, edge_tri


# =============================================================================================
# geometry (Appendix B of SURVEY.md)
# =============================================================================================
def elem_centers(m: Mesh):
    """src/oce_mesh.F90:2160-2181."""
    en = m.elem2D_nodes.astype(np.int64) - 1
    ax = m.coord_nod2D[en, 0].copy()
    amin = ax.min(axis=1, keepdims=True)
    ax = np.where(ax - amin >= m.cyclic_length / 2.0, ax - m.cyclic_length, ax)
    ax = np.where(ax - amin < -m.cyclic_length / 2.0, ax + m.cyclic_length, ax)
    cx = ax.sum(axis=1) / 3.0
    cy = m.coord_nod2D[en, 1].sum(axis=1) / 3.0
    return cx, cy


def derive_geometry(m: Mesh, need_nod_in_elem_for_all: bool = True) -> Mesh:
    """Fill every derived array of ``Mesh`` from coordinates, connectivity and level counts."""
    cyc = m.cyclic_length
    en = m.elem2D_nodes.astype(np.int64) - 1
    x, y = m.coord_nod2D[:, 0], m.coord_nod2D[:, 1]
    Nh, T, E, nl = m.Nh, m.T, m.E, m.nl

    if m.nod_in_elem2D is None:
        m.nod_in_elem2D, m.nod_in_elem2D_num = build_nod_in_elem(m.elem2D_nodes, Nh)

    # --- element centres, cos, areas: oce_mesh.F90:2226-2239, :2533-2553 ----------------------
    cx, cy = elem_centers(m)
    m.elem_cos = np.ones(T) if m.cartesian else np.cos(cy)
    ay = np.cos(y[en].sum(axis=1) / 3.0)
    if m.cartesian:
        ay = np.ones(T)
    a1 = _trim_cyclic(x[en[:, 1]] - x[en[:, 0]], cyc) * ay
    a2 = y[en[:, 1]] - y[en[:, 0]]
    b1 = _trim_cyclic(x[en[:, 2]] - x[en[:, 0]], cyc) * ay
    b2 = y[en[:, 2]] - y[en[:, 0]]
    area_rad = 0.5 * np.abs(a1 * b2 - b1 * a2)
    m.elem_area = area_rad * R_EARTH * R_EARTH

    # --- scalar-cell areas: oce_mesh.F90:2277-2343 (no cavity => areasvol = area) -------------
    area = np.zeros((Nh, nl))
    third = area_rad / 3.0
    flat_nodes = en.ravel()                                   # element-major = nod_in_elem order
    if (m.ulevels == 1).all():
        # no cavity: bin every element at its deepest layer, then sum bottom-up (additions only)
        deepest = np.repeat(m.nlevels.astype(np.int64) - 2, 3)              # 0-based layer index
        c = np.bincount(flat_nodes * nl + deepest, weights=np.repeat(third, 3), minlength=Nh * nl)
        area = np.cumsum(c.reshape(Nh, nl)[:, ::-1], axis=1)[:, ::-1].copy()
    else:
        for nz in range(1, nl):                               # 1-based layer index
            act = (m.ulevels <= nz) & (nz <= m.nlevels - 1)
            w = np.repeat(np.where(act, third, 0.0), 3)
            area[:, nz - 1] = np.bincount(flat_nodes, weights=w, minlength=Nh)
    area *= R_EARTH * R_EARTH
    m.area = area
    m.areasvol = np.zeros_like(area)
    lev = np.arange(1, nl + 1)[None, :]
    valid = (lev >= m.ulevels_nod2D[:, None]) & (lev <= m.nlevels_nod2D[:, None] - 1)
    m.areasvol[valid] = area[valid]
    if not (m.ulevels == 1).all():
        # use_cavity (oce_mesh.F90:2299-2321): where a cavity element sits on the upper face of the scalar cell
        # (some element around the node starts below layer nz) the "mid" area is the area of the LOWER face
        contrib = np.zeros((Nh, nl), np.int64)
        for nz in range(1, nl):
            contrib[:, nz - 1] = np.bincount(flat_nodes, weights=np.repeat((m.ulevels - 1 >= nz).astype(np.float64), 3), minlength=Nh)
        nzmax = (m.nlevels_nod2D.astype(np.int64) - 1)[:, None]
        below = np.take_along_axis(area, np.clip(np.minimum(lev + 1, nzmax) - 1, 0, nl - 1) * np.ones((Nh, 1), np.int64), axis=1)
        m.areasvol = np.where(valid & (contrib > 0), below, m.areasvol)

    # --- edge vectors: oce_mesh.F90:2559-2598 -------------------------------------------------
    ed = m.edges.astype(np.int64) - 1
    et = m.edge_tri.astype(np.int64) - 1
    dxy = m.coord_nod2D[ed[:, 1]] - m.coord_nod2D[ed[:, 0]]
    dxy[:, 0] = _trim_cyclic(dxy[:, 0], cyc)
    m.edge_dxdy = dxy
    # edge_center: oce_mesh.F90:2140-2158
    ax_ = x[ed[:, 0]].copy(); bx_ = x[ed[:, 1]].copy()
    d = ax_ - bx_
    ax_ = np.where(d > cyc / 2.0, ax_ - cyc, ax_)
    bx_ = np.where(d < -cyc / 2.0, bx_ - cyc, bx_)
    mx = 0.5 * (ax_ + bx_)
    my = 0.5 * (y[ed[:, 0]] + y[ed[:, 1]])
    cross = np.zeros((E, 4))
    b_x = _trim_cyclic(cx[et[:, 0]] - mx, cyc) * m.elem_cos[et[:, 0]]
    b_y = cy[et[:, 0]] - my
    cross[:, 0] = b_x * R_EARTH
    cross[:, 1] = b_y * R_EARTH
    has2 = et[:, 1] >= 0
    e2 = np.where(has2, et[:, 1], 0)
    b_x = _trim_cyclic(cx[e2] - mx, cyc) * m.elem_cos[e2]
    b_y = cy[e2] - my
    cross[:, 2] = np.where(has2, b_x * R_EARTH, 0.0)
    cross[:, 3] = np.where(has2, b_y * R_EARTH, 0.0)
    m.edge_cross_dxdy = cross

    # --- P1 gradients: oce_mesh.F90:2644-2666 -------------------------------------------------
    dX31 = m.elem_cos * _trim_cyclic(x[en[:, 2]] - x[en[:, 0]], cyc)
    dX21 = m.elem_cos * _trim_cyclic(x[en[:, 1]] - x[en[:, 0]], cyc)
    dY31 = y[en[:, 2]] - y[en[:, 0]]
    dY21 = y[en[:, 1]] - y[en[:, 0]]
    dfac = -0.5 * R_EARTH / m.elem_area
    g = np.empty((T, 6))
    g[:, 0] = (-dY31 + dY21) * dfac
    g[:, 1] = dY31 * dfac
    g[:, 2] = -dY21 * dfac
    g[:, 3] = (dX31 - dX21) * dfac
    g[:, 4] = -dX31 * dfac
    g[:, 5] = dX21 * dfac
    m.gradient_sca = g

    # --- min/max levels around a node: oce_mesh.F90:1656-1686 ---------------------------------
    big = np.iinfo(np.int32).max
    nmin = np.full(Nh, big, np.int64)
    umax = np.zeros(Nh, np.int64)
    for k in range(3):
        np.minimum.at(nmin, en[:, k], m.nlevels)
        np.maximum.at(umax, en[:, k], m.ulevels)
    m.nlevels_nod2D_min = nmin.astype(np.int32)
    m.ulevels_nod2D_max = umax.astype(np.int32)
    return m


def nboundary_lay(m: Mesh) -> np.ndarray:
    """``nboundary_lay`` of muscl_adv_init (src/oce_muscl_adv.F90:92-155)."""
    nb = np.full(m.Nh, m.nl - 1, np.int64)
    ed = m.edges.astype(np.int64) - 1
    et = m.edge_tri.astype(np.int64) - 1
    bnd = (et <= -1).any(axis=1)
    mn = np.minimum(m.nlevels[np.maximum(et[:, 0], 0)], m.nlevels[np.maximum(et[:, 1], 0)]) - 1
    # sequential semantics: a boundary edge sets 0, an interior edge takes the min; 0 is absorbing
    # for min with non-negative values, so order does not matter.
    for k in range(2):
        np.minimum.at(nb, ed[~bnd, k], mn[~bnd])
    nb[ed[bnd].ravel()] = 0
    return nb.astype(np.int32)


# =============================================================================================
# readers for the reference's ASCII fixtures (Appendix A of SURVEY.md)
# =============================================================================================
def read_fesom_mesh(path: str, cyclic_length_deg: float = 360.0, cartesian: bool = False) -> Mesh:
    """Read nod2d/elem2d/aux3d/nlvls/elvls/edges/edge_tri (reference readers:
    src/oce_mesh.F90:353-600, :977-1100, :1751-1983)."""
    def p(f):
        return os.path.join(path, f)

    nod = np.loadtxt(p("nod2d.out"), skiprows=1)
    coord = np.ascontiguousarray(nod[:, 1:3]) * RAD
    elem = np.loadtxt(p("elem2d.out"), skiprows=1, dtype=np.int64).astype(np.int32)
    with open(p("aux3d.out")) as f:
        tok = f.read().split()
    nl = int(tok[0])
    zbar = np.array(tok[1:1 + nl], dtype=np.float64)
    if zbar[1] > 0:
        zbar = -zbar                                        # oce_mesh.F90:598
    nlv_n = np.loadtxt(p("nlvls.out"), dtype=np.int64).astype(np.int32)
    nlv_e = np.loadtxt(p("elvls.out"), dtype=np.int64).astype(np.int32)
    edges = np.loadtxt(p("edges.out"), dtype=np.int64).astype(np.int32)
    etri = np.loadtxt(p("edge_tri.out"), dtype=np.int64).astype(np.int32)
    etri[etri < 0] = 0                                      # oce_mesh.F90:1900
    N, T, E = coord.shape[0], elem.shape[0], edges.shape[0]
    ulv_e, ulv_n = np.ones(T, np.int32), np.ones(N, np.int32)
    if os.path.exists(p("cavity_elvls.out")):               # use_cavity: oce_mesh.F90:1139-1350
        ulv_e = np.loadtxt(p("cavity_elvls.out"), dtype=np.int64).astype(np.int32)
        ulv_n = np.loadtxt(p("cavity_nlvls.out"), dtype=np.int64).astype(np.int32)
    m = Mesh(nl=nl, myDim_nod2D=N, eDim_nod2D=0, myDim_elem2D=T, eDim_elem2D=0, myDim_edge2D=E,
             cyclic_length=cyclic_length_deg * RAD, cartesian=cartesian, coord_nod2D=coord,
             elem2D_nodes=elem, edges=edges, edge_tri=etri, nlevels=nlv_e,
             ulevels=ulv_e, nlevels_nod2D=nlv_n, ulevels_nod2D=ulv_n,
             zbar=zbar)
    m.myList_nod2D = np.arange(1, N + 1, dtype=np.int32)
    m.myList_elem2D = np.arange(1, T + 1, dtype=np.int32)
    m.myList_edge2D = np.arange(1, E + 1, dtype=np.int32)
    return derive_geometry(m)


def read_dist(path: str, npes: int) -> Dict[str, object]:
    """Read the checked-in ``dist_N`` partition (rpart.out, my_listRRRRR.out,
    com_infoRRRRR.out; reference reader src/oce_mesh.F90:279-340, :820-940)."""
    d = os.path.join(path, f"dist_{npes}")
    with open(os.path.join(d, "rpart.out")) as f:
        tok = f.read().split()
    assert int(tok[0]) == npes
    counts = np.array(tok[1:1 + npes], dtype=np.int64)
    out = {"counts": counts, "ranks": []}
    for r in range(npes):
        with open(os.path.join(d, f"my_list{r:05d}.out")) as f:
            t = np.array(f.read().split(), dtype=np.int64)
        i = 0
        assert t[i] == r; i += 1
        myN, eN = int(t[i]), int(t[i + 1]); i += 2
        nodes = t[i:i + myN + eN]; i += myN + eN
        myT, eT, eXT = int(t[i]), int(t[i + 1]), int(t[i + 2]); i += 3
        elems = t[i:i + myT + eT + eXT]; i += myT + eT + eXT
        myE, eE = int(t[i]), int(t[i + 1]); i += 2
        edges = t[i:i + myE + eE]
        with open(os.path.join(d, f"com_info{r:05d}.out")) as f:
            c = np.array(f.read().split(), dtype=np.int64)
        j = 1
        coms = []
        for _ in range(3):
            rn = int(c[j]); j += 1
            rPE = c[j:j + rn]; j += rn
            rptr = c[j:j + rn + 1]; j += rn + 1
            rlist = c[j:j + int(rptr[-1]) - 1]; j += int(rptr[-1]) - 1
            sn = int(c[j]); j += 1
            sPE = c[j:j + sn]; j += sn
            sptr = c[j:j + sn + 1]; j += sn + 1
            slist = c[j:j + int(sptr[-1]) - 1]; j += int(sptr[-1]) - 1
            coms.append(ComStruct(rPE.astype(np.int32), rptr.astype(np.int32), rlist.astype(np.int32),
                                  sPE.astype(np.int32), sptr.astype(np.int32), slist.astype(np.int32)))
        out["ranks"].append(dict(myDim_nod2D=myN, eDim_nod2D=eN, myList_nod2D=nodes,
                                 myDim_elem2D=myT, eDim_elem2D=eT, eXDim_elem2D=eXT,
                                 myList_elem2D=elems, myDim_edge2D=myE, eDim_edge2D=eE,
                                 myList_edge2D=edges, com_nod2D=coms[0], com_elem2D=coms[1], com_elem2D_full=coms[2]))
    # node -> owning rank
    part = np.empty(int(counts.sum()), np.int32)
    for r, info in enumerate(out["ranks"]):
        part[info["myList_nod2D"][:info["myDim_nod2D"]] - 1] = r
    out["part"] = part
    return out


def element_halo(g: Mesh, part: np.ndarray, mype: int) -> Dict[str, object]:
    """Element halo of rank ``mype``: the reference's ``communication_elemn`` (src/gen_comm.F90:222-527)
    followed by ``com_global2local`` / ``save_dist_mesh`` (src/oce_local.F90:76-197), vectorised.

    An element is local if any of its nodes is owned (myDim_elem2D, ascending global id).  It is
    *received* if none of its nodes is owned and it shares an edge (``com_elem2D`` -> eDim_elem2D) or only
    a node (``com_elem2D_full`` -> eXDim_elem2D) with a local element; its sender is the owner of its first
    node.  A local element whose first node is owned is *sent* to every rank that owns a node of a
    neighbouring element and does not itself own a node of the element.  Lists are grouped by rank
    ascending, ascending global id inside a group, and renumbered locally: own elements 1..myDim, then the
    eDim elements in ``com_elem2D%rlist`` order, then the eXDim elements in ``com_elem2D_full%rlist`` order.

    Needed by the multi-rank form of tracer_gradient_elements / fill_up_dn_grad (exchange_elem(tr_xy),
    src/oce_tracer_mod.F90:140); the advection path itself only needs the node halo."""
    part = np.asarray(part).astype(np.int64)
    en = g.elem2D_nodes.astype(np.int64) - 1
    T = en.shape[0]
    pe_n = part[en]                                              # (T, 3) owner of each element node
    local = (pe_n == mype).any(axis=1)
    my_elem = np.flatnonzero(local)
    main_mine = local & (pe_n[:, 0] == mype)

    # element pairs (el, elem): edge neighbours from edge_tri, node neighbours from nod_in_elem2D
    et = g.edge_tri.astype(np.int64) - 1
    inner = et[:, 1] >= 0
    pair_edge = np.concatenate([et[inner], et[inner][:, ::-1]], axis=0)          # both directions
    nie, num = build_nod_in_elem(g.elem2D_nodes, g.Nh)
    nie = nie.astype(np.int64) - 1

    def node_pairs(els):
        if els.size == 0:
            return np.zeros((0, 2), np.int64)
        nb = nie[en[els]]                                        # (n, 3, ld) elements around the element's nodes
        src = np.broadcast_to(els[:, None, None], nb.shape)
        ok = nb >= 0
        return np.stack([src[ok], nb[ok]], axis=1)

    def recv_send(pairs, recv_prev, send_prev):
        el, elem = pairs[:, 0], pairs[:, 1]
        # receive: elem has no owned node, el is local
        r = np.unique(elem[local[el] & ~local[elem]])
        recv = np.union1d(recv_prev, r)
        # send: el's main owner is mype, elem has a foreign node; to every owner of elem that owns no node of el
        m = main_mine[el] & (pe_n[elem] != mype).any(axis=1)
        el_m, elem_m = el[m], elem[m]
        out = [send_prev]
        for i in range(3):
            ep = pe_n[elem_m, i]
            need = ~(pe_n[el_m] == ep[:, None]).any(axis=1)
            out.append(np.stack([ep[need], el_m[need]], axis=1))
        send = np.unique(np.concatenate(out, axis=0), axis=0) if out else send_prev
        return recv, send

    empty_send = np.zeros((0, 2), np.int64)
    recv1, send1 = recv_send(pair_edge, np.zeros(0, np.int64), empty_send)
    # the full communicator continues from the edge one (the reference does not reset its tables);
    # only elements near the partition boundary can receive or send: restrict the node-neighbour pairs
    near = np.zeros(g.Nh, bool)
    near[en[~local].ravel()] = True                              # nodes of non-local elements ...
    cand = my_elem[near[en[my_elem]].any(axis=1)]                # ... touched by local elements
    foreign_node = np.zeros(g.Nh, bool)
    foreign_node[part != mype] = True
    ring = np.zeros(g.Nh, bool)
    ring[en[(foreign_node[en]).any(axis=1)].ravel()] = True      # nodes of elements that have a foreign node
    cand2 = my_elem[ring[en[my_elem]].any(axis=1)]
    recv2, send2 = recv_send(node_pairs(np.union1d(cand, cand2)), recv1, send1)

    def order_recv(recv):
        owner = pe_n[recv, 0]
        o = np.lexsort((recv, owner))
        return recv[o], owner[o]

    r1, o1 = order_recv(recv1)
    r2, o2 = order_recv(recv2)
    ex = r2[~np.isin(r2, r1)]                                    # full-list order restricted to the new ones
    loc_of = np.full(T, -1, np.int64)
    loc_of[my_elem] = np.arange(my_elem.size)
    loc_of[r1] = my_elem.size + np.arange(r1.size)
    loc_of[ex] = my_elem.size + r1.size + np.arange(ex.size)

    def com(recv, owner, send):
        rPE, rc = np.unique(owner, return_counts=True)
        s = send[np.lexsort((send[:, 1], send[:, 0]))] if send.size else send
        sPE, sc = np.unique(s[:, 0], return_counts=True) if s.size else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        return ComStruct(rPE.astype(np.int32), np.concatenate(([1], 1 + np.cumsum(rc))).astype(np.int32),
                         (loc_of[recv] + 1).astype(np.int32), sPE.astype(np.int32),
                         np.concatenate(([1], 1 + np.cumsum(sc))).astype(np.int32),
                         (loc_of[s[:, 1]] + 1).astype(np.int32) if s.size else np.zeros(0, np.int32))

    return dict(myDim_elem2D=int(my_elem.size), eDim_elem2D=int(r1.size), eXDim_elem2D=int(ex.size),
                myList_elem2D=(np.concatenate([my_elem, r1, ex]) + 1).astype(np.int32),
                com_elem2D=com(r1, o1, send1), com_elem2D_full=com(r2, o2, send2))


@dataclass
class GradientMesh:
    """Static inputs of tracer_gradient_elements / fill_up_dn_grad on one rank (adv_gradient_mesh_desc_t):
    the element halo eDim + eXDim appended to the own elements, element neighbourhoods of ALL local nodes."""
    n_elem: int
    nod_in_elem2D: np.ndarray        # (Nh, ld) local element numbers, 1-based, 0 padded
    nod_in_elem2D_num: np.ndarray    # (Nh,)
    nlevels: np.ndarray              # (n_elem,)
    ulevels: np.ndarray
    edge_up_dn_tri: np.ndarray       # (E, 2) local element numbers, 0 = none
    elem_area: np.ndarray            # (n_elem,)
    myList_elem2D: np.ndarray        # (n_elem,) global ids, 1-based
    com_elem2D_full: ComStruct = field(default_factory=ComStruct)


def gradient_mesh(g: Mesh, tri_global: np.ndarray, part: Optional[np.ndarray] = None, loc: Optional[Mesh] = None) -> GradientMesh:
    """Gradient mesh of the 1-rank mesh ``g`` (part is None) or of rank ``loc.mype``'s local mesh ``loc`` cut from
    ``g``: own elements, then the eDim and eXDim halo elements of ``element_halo`` (communication_elemn,
    src/gen_comm.F90:222-527); nod_in_elem2D of halo nodes is the owner's list (src/oce_mesh.F90:2064-2088);
    edge_up_dn_tri is the global table (find_up_downwind_triangles is evaluated on the whole mesh by the
    reference as well: its coord_elem / e_nodes arrays are halo-exchanged, src/oce_muscl_adv.F90:192-216)."""
    if part is None:
        nie = np.asarray(g.nod_in_elem2D)
        return GradientMesh(n_elem=int(g.T), nod_in_elem2D=nie.astype(np.int32), nod_in_elem2D_num=np.asarray(g.nod_in_elem2D_num, np.int32),
                            nlevels=np.asarray(g.nlevels, np.int32), ulevels=np.asarray(g.ulevels, np.int32),
                            edge_up_dn_tri=np.asarray(tri_global, np.int32), elem_area=np.asarray(g.elem_area, np.float64),
                            myList_elem2D=np.arange(1, g.T + 1, dtype=np.int32))
    h = element_halo(g, part, loc.mype)
    elist = h["myList_elem2D"].astype(np.int64) - 1
    assert np.array_equal(elist[:loc.T], loc.myList_elem2D.astype(np.int64) - 1)
    e_g2l = np.full(g.T, -1, np.int64)
    e_g2l[elist] = np.arange(elist.size)
    nodes = loc.myList_nod2D.astype(np.int64) - 1
    edges = loc.myList_edge2D.astype(np.int64) - 1
    nie_g = np.asarray(g.nod_in_elem2D)[nodes].astype(np.int64) - 1
    if (e_g2l[nie_g[nie_g >= 0]] < 0).any():
        raise ValueError("an element around a local node is outside the element halo")
    nie_l = np.where(nie_g >= 0, e_g2l[np.maximum(nie_g, 0)] + 1, 0).astype(np.int32)
    tri_g = np.asarray(tri_global)[edges].astype(np.int64) - 1
    if (e_g2l[tri_g[tri_g >= 0]] < 0).any():
        raise ValueError("an up/down-wind triangle is outside the element halo")
    tri_l = np.where(tri_g >= 0, e_g2l[np.maximum(tri_g, 0)] + 1, 0).astype(np.int32)
    return GradientMesh(n_elem=int(elist.size), nod_in_elem2D=nie_l, nod_in_elem2D_num=np.asarray(g.nod_in_elem2D_num)[nodes].astype(np.int32),
                        nlevels=np.asarray(g.nlevels)[elist].astype(np.int32), ulevels=np.asarray(g.ulevels)[elist].astype(np.int32),
                        edge_up_dn_tri=tri_l, elem_area=np.asarray(g.elem_area)[elist].astype(np.float64),
                        myList_elem2D=(elist + 1).astype(np.int32), com_elem2D_full=h["com_elem2D_full"])


# =============================================================================================
# synthetic meshes (configs 3-5 of BASELINE.json)
# =============================================================================================
def _hash01(i: np.ndarray, salt: int) -> np.ndarray:
    """Deterministic U(0,1) from integer ids (splitmix-style), identical on every rank."""
    z = (i.astype(np.uint64) + np.uint64(salt) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def synth_mesh(nx: int, ny: int, nl: int = 48, lon0: float = -30.0, lon1: float = 30.0,
               lat0: float = -30.0, lat1: float = 30.0, staircase: bool = True,
               min_layers: int = 6, derive: bool = True) -> Mesh:
    """Structured lon-lat patch, every quad cut into two clockwise triangles (the reference's
    element orientation), optional bottom staircase: a smooth basin plus per-element noise
    removing up to ~30 bottom layers.  Columns keep >= ``min_layers`` layers."""
    with np.errstate(over="ignore"):
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")     # (ny, nx)
        lon = lon0 + (lon1 - lon0) * ii / (nx - 1)
        lat = lat0 + (lat1 - lat0) * jj / (ny - 1)
        coord = np.stack([lon.ravel(), lat.ravel()], 1) * RAD
        nid = (jj * nx + ii)                                                   # node id 0-based
        n00 = nid[:-1, :-1].ravel(); n10 = nid[:-1, 1:].ravel()
        n01 = nid[1:, :-1].ravel(); n11 = nid[1:, 1:].ravel()
        # clockwise (x east, y north): (n00, n11, n10) and (n00, n01, n11)
        t_a = np.stack([n00, n11, n10], 1)
        t_b = np.stack([n00, n01, n11], 1)
        elem = np.empty((2 * t_a.shape[0], 3), np.int64)
        elem[0::2] = t_a; elem[1::2] = t_b
        elem2D_nodes = (elem + 1).astype(np.int32)
        T = elem.shape[0]
        N = nx * ny
        zbar = zbar_levels(nl)
        if staircase:
            cxe = coord[elem, 0].mean(1); cye = coord[elem, 1].mean(1)
            sx = (cxe - lon0 * RAD) / ((lon1 - lon0) * RAD); sy = (cye - lat0 * RAD) / ((lat1 - lat0) * RAD)
            basin = 0.55 + 0.45 * np.sin(np.pi * sx) * np.sin(np.pi * sy) \
                + 0.12 * np.sin(7 * np.pi * sx) * np.cos(5 * np.pi * sy)
            noise = _hash01(np.arange(T), 1)
            frac = np.clip(basin - 0.10 * noise, 0.0, 1.0)
            nlev_e = np.clip(np.rint(min_layers + 1 + frac * (nl - min_layers - 1)), min_layers + 1, nl)
        else:
            nlev_e = np.full(T, nl)
        nlev_e = nlev_e.astype(np.int32)
        nlev_n = np.zeros(N, np.int64)
        for k in range(3):
            np.maximum.at(nlev_n, elem[:, k], nlev_e)
    edges, edge_tri = build_edges(elem2D_nodes, coord, 360.0 * RAD)
    m = Mesh(nl=nl, myDim_nod2D=N, eDim_nod2D=0, myDim_elem2D=T, eDim_elem2D=0,
             myDim_edge2D=edges.shape[0], cyclic_length=360.0 * RAD, cartesian=False,
             coord_nod2D=coord, elem2D_nodes=elem2D_nodes, edges=edges, edge_tri=edge_tri,
             nlevels=nlev_e, ulevels=np.ones(T, np.int32), nlevels_nod2D=nlev_n.astype(np.int32),
             ulevels_nod2D=np.ones(N, np.int32), zbar=zbar)
    m.myList_nod2D = np.arange(1, N + 1, dtype=np.int32)
    m.myList_elem2D = np.arange(1, T + 1, dtype=np.int32)
    m.myList_edge2D = np.arange(1, m.E + 1, dtype=np.int32)
    return derive_geometry(m) if derive else m


# =============================================================================================
# partition -> local meshes (gen_comm.F90)
# =============================================================================================
def node_graph(g: Mesh):
    """CSR node adjacency (1-based, as ``do_partit`` expects; built by stiff_mat_ini,
    reference src/fvom_init.F90:1609-1677)."""
    ed = g.edges.astype(np.int64) - 1
    a = np.concatenate([ed[:, 0], ed[:, 1]])
    b = np.concatenate([ed[:, 1], ed[:, 0]])
    order = np.lexsort((b, a))
    a, b = a[order], b[order]
    cnt = np.bincount(a, minlength=g.Nh)
    ptr = np.concatenate(([0], np.cumsum(cnt)))
    return (ptr + 1).astype(np.int32), (b + 1).astype(np.int32)


def localize(g: Mesh, part: np.ndarray, mype: int) -> Mesh:
    """Cut rank ``mype``'s local mesh out of the global mesh ``g``.

    Numbering rules (reference: src/gen_comm.F90:25-47,:172-196,:616-639, src/oce_local.F90:36-47):
    owned nodes ascending global id, then halo nodes grouped by owner rank ascending, ascending
    global id inside a group; elements = those with >= 1 owned node, ascending; edges likewise.
    Halo = every non-owned node of an element that touches an owned node.  ``slist`` for rank p =
    owned nodes that share an element with a node owned by p."""
    part = np.asarray(part)
    npes = int(part.max()) + 1
    en = g.elem2D_nodes.astype(np.int64) - 1
    ed = g.edges.astype(np.int64) - 1
    own_n = part == mype
    my_elem = np.flatnonzero(own_n[en].any(axis=1))                     # ascending
    my_edge = np.flatnonzero(own_n[ed].any(axis=1))
    owned = np.flatnonzero(own_n)
    touched = np.unique(en[my_elem].ravel())
    halo = touched[~own_n[touched]]
    halo = halo[np.lexsort((halo, part[halo]))]                         # by owner, then id
    loc_nodes = np.concatenate([owned, halo])
    N, eN = owned.size, halo.size
    g2l = np.full(g.Nh, -1, np.int64)
    g2l[loc_nodes] = np.arange(N + eN)
    e2l = np.full(g.T, -1, np.int64)
    e2l[my_elem] = np.arange(my_elem.size)

    # recv lists: halo nodes are already grouped by owner
    rPE, rcount = np.unique(part[halo], return_counts=True)
    rptr = np.concatenate(([1], 1 + np.cumsum(rcount)))
    rlist = N + 1 + np.arange(eN)
    # send lists: owned node n goes to rank p if an element containing n has a node owned by p
    pe_of = part[en[my_elem]]                                            # (Tm, 3)
    pairs = []
    for k in range(3):
        for j in range(3):
            if j == k:
                continue
            sel = (pe_of[:, k] == mype) & (pe_of[:, j] != mype)
            pairs.append(np.stack([pe_of[sel, j], en[my_elem][sel, k]], 1))
    pairs = np.unique(np.concatenate(pairs, 0), axis=0) if pairs else np.zeros((0, 2), np.int64)
    sPE, scount = np.unique(pairs[:, 0], return_counts=True)
    sptr = np.concatenate(([1], 1 + np.cumsum(scount)))
    slist = g2l[pairs[:, 1]] + 1                                         # sorted by (pe, global id)
    com = ComStruct(rPE.astype(np.int32), rptr.astype(np.int32), rlist.astype(np.int32),
                    sPE.astype(np.int32), sptr.astype(np.int32), slist.astype(np.int32))

    et = g.edge_tri.astype(np.int64) - 1
    l_et = np.where(et[my_edge] >= 0, e2l[np.maximum(et[my_edge], 0)], -1)
    assert (l_et[:, 0] >= 0).all()
    m = Mesh(nl=g.nl, myDim_nod2D=N, eDim_nod2D=eN, myDim_elem2D=my_elem.size, eDim_elem2D=0,
             myDim_edge2D=my_edge.size, cyclic_length=g.cyclic_length, cartesian=g.cartesian,
             coord_nod2D=g.coord_nod2D[loc_nodes],
             elem2D_nodes=(g2l[en[my_elem]] + 1).astype(np.int32),
             edges=(g2l[ed[my_edge]] + 1).astype(np.int32),
             edge_tri=(l_et + 1).astype(np.int32),
             nlevels=g.nlevels[my_elem], ulevels=g.ulevels[my_elem],
             nlevels_nod2D=g.nlevels_nod2D[loc_nodes], ulevels_nod2D=g.ulevels_nod2D[loc_nodes],
             zbar=g.zbar)
    assert (m.elem2D_nodes > 0).all() and (m.edges > 0).all()
    m.mype, m.npes = mype, npes
    m.myList_nod2D = (loc_nodes + 1).astype(np.int32)
    m.myList_elem2D = (my_elem + 1).astype(np.int32)
    m.myList_edge2D = (my_edge + 1).astype(np.int32)
    m.com_nod2D = com
    # per-entity geometry is copied from the global derivation (the reference fills halo values
    # by exchange_nod / exchange_elem, which gives the owner's = the global value)
    m.nod_in_elem2D, m.nod_in_elem2D_num = build_nod_in_elem(m.elem2D_nodes, N + eN)
    m.elem_area = g.elem_area[my_elem]
    m.elem_cos = g.elem_cos[my_elem]
    m.edge_dxdy = g.edge_dxdy[my_edge]
    m.edge_cross_dxdy = g.edge_cross_dxdy[my_edge]
    m.gradient_sca = g.gradient_sca[my_elem]
    m.area = g.area[loc_nodes]
    m.areasvol = g.areasvol[loc_nodes]
    m.nlevels_nod2D_min = g.nlevels_nod2D_min[loc_nodes]
    m.ulevels_nod2D_max = g.ulevels_nod2D_max[loc_nodes]
    return m


def simple_partition(g: Mesh, npes: int) -> np.ndarray:
    """Fallback partitioner (recursive coordinate bisection on node coordinates, balanced by node
    count).  Used only when the METIS wrapper is unavailable; the product partitioner is
    ``fesom2_b200.partition.metis_partition`` which mirrors ``do_partit``."""
    part = np.zeros(g.Nh, np.int32)

    def rec(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        left = np_ // 2
        c = g.coord_nod2D[idx]
        ax = 0 if (c[:, 0].max() - c[:, 0].min()) * np.cos(c[:, 1].mean()) >= (c[:, 1].max() - c[:, 1].min()) else 1
        order = np.argsort(c[:, ax], kind="stable")
        cut = idx.size * left // np_
        rec(idx[order[:cut]], p0, left)
        rec(idx[order[cut:]], p0 + left, np_ - left)

    rec(np.arange(g.Nh), 0, npes)
    return part


def load_npz_mesh(path: str) -> Mesh:
    """Load a mesh fixture written by tests/golden/make_mesh_fixtures.py (raw mesh data of the
    reference's test/meshes/*); returns the derived global mesh.  Partition vectors stored in the
    file are attached as ``mesh.parts = {npes: part}``."""
    z = np.load(path)
    coord = np.ascontiguousarray(z["coord_deg"], dtype=np.float64) * RAD
    N, T, E = coord.shape[0], z["elem2D_nodes"].shape[0], z["edges"].shape[0]
    m = Mesh(nl=int(z["nl"]), myDim_nod2D=N, eDim_nod2D=0, myDim_elem2D=T, eDim_elem2D=0, myDim_edge2D=E,
             cyclic_length=float(z["cyclic_length_deg"]) * RAD, cartesian=False, coord_nod2D=coord,
             elem2D_nodes=z["elem2D_nodes"].astype(np.int32), edges=z["edges"].astype(np.int32),
             edge_tri=z["edge_tri"].astype(np.int32), nlevels=z["nlevels"].astype(np.int32),
             ulevels=z["ulevels"].astype(np.int32) if "ulevels" in z.files else np.ones(T, np.int32),
             nlevels_nod2D=z["nlevels_nod2D"].astype(np.int32),
             ulevels_nod2D=z["ulevels_nod2D"].astype(np.int32) if "ulevels_nod2D" in z.files else np.ones(N, np.int32),
             zbar=z["zbar"].astype(np.float64))
    m.myList_nod2D = np.arange(1, N + 1, dtype=np.int32)
    m.myList_elem2D = np.arange(1, T + 1, dtype=np.int32)
    m.myList_edge2D = np.arange(1, E + 1, dtype=np.int32)
    derive_geometry(m)
    m.parts = {int(k[4:]): z[k].astype(np.int32) for k in z.files if k.startswith("part")}
    return m
