"""Synthetic ocean state and tracer fields for the advection path.

Everything here is INPUT SYNTHESIS for tests and the benchmark -- it is not on the hot path.  It
is written with torch so that the same code fills a 3k-node test mesh on the CPU and a 3M-node
mesh on the GPU.  The formulas restate the reference routines that produce the path's inputs:

* ``helem/hnode/zbar_3d_n/Z_3d_n``  src/oce_ale.F90:349-374, :818-864 (linfs, full cells)
* ``w`` from continuity              src/oce_ale.F90:2164-2310 (vert_vel_ale)
* ``w_e/w_i`` split                  src/oce_ale.F90:3001-3049 (compute_Wvel_split)
* ``valuesAB``                       src/oce_tracer_mod.F90:45-54 (AB2, epsilon = 0.1)
* ``tr_xy``                          src/oce_tracer_mod.F90:147-188 (tracer_gradient_elements)
* ``edge_up_dn_tri``                 src/oce_muscl_adv.F90:162-352
* ``edge_up_dn_grad``                src/oce_muscl_adv.F90:356-525 (fill_up_dn_grad)

numpy/torch shapes are the transposes of the Fortran shapes (Fortran ``w(nz,n)`` = ``w[n,nz]``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from .mesh import Mesh, R_EARTH, nboundary_lay

F64 = torch.float64
EPSILON_AB = 0.1          # src/oce_modules.F90:105


def _t(a, device, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a), device=device)
    return t.to(dtype) if dtype is not None else t


def _hash01(ids: torch.Tensor, salt: int) -> torch.Tensor:
    """Deterministic U(0,1) from int64 ids (xorshift-multiply), rank-independent."""
    z = ids.to(torch.int64) + salt * 0x9E3779B97F4A7C1
    z = (z ^ (z >> 30)) * 0x3F58476D1CE4E5B9
    z = (z ^ (z >> 27)) * 0x14D049BB133111EB
    z = z ^ (z >> 31)
    return (z & ((1 << 52) - 1)).to(F64) / float(1 << 52)


@dataclass
class OceanState:
    """Per-step ALE/dynamics state the path reads (t_dyn: uv,w,w_e,w_i; t_mesh: helem,hnode,...)."""
    uv: torch.Tensor          # (T, L, 2)
    w: torch.Tensor           # (Nh, nl)
    w_e: torch.Tensor         # (Nh, nl)
    w_i: torch.Tensor         # (Nh, nl)
    helem: torch.Tensor       # (T, L)
    hnode: torch.Tensor       # (Nh, L)
    hnode_new: torch.Tensor   # (Nh, L)
    zbar_3d_n: torch.Tensor   # (Nh, nl)
    Z_3d_n: torch.Tensor      # (Nh, L)
    zbar_n_bot: torch.Tensor  # (Nh,)
    use_wsplit: bool = False


@dataclass
class TracerFields:
    """One entry of t_tracer%data plus its slice of t_tracer%work."""
    values: torch.Tensor           # (Nh, L)
    valuesAB: torch.Tensor         # (Nh, L)
    edge_up_dn_grad: torch.Tensor  # (E, L, 4)
    tra_adv_hor: str = "MFCT"
    tra_adv_ver: str = "QR4C"
    tra_adv_lim: str = "FCT"
    tra_adv_ph: float = 0.0
    tra_adv_pv: float = 1.0


# ---------------------------------------------------------------------------------------------
def layer_masks(m: Mesh, device):
    L = m.L
    lev = torch.arange(1, m.nl + 1, device=device)
    nle = _t(m.nlevels, device, torch.int64); ule = _t(m.ulevels, device, torch.int64)
    nln = _t(m.nlevels_nod2D, device, torch.int64); uln = _t(m.ulevels_nod2D, device, torch.int64)
    emask = (lev[None, :L] >= ule[:, None]) & (lev[None, :L] <= nle[:, None] - 1)      # (T, L)
    nmask = (lev[None, :L] >= uln[:, None]) & (lev[None, :L] <= nln[:, None] - 1)      # (Nh, L)
    return emask, nmask


def make_state(m: Mesh, device="cpu", vel_amp: float = 0.5, ale_amp: float = 0.0,
               use_wsplit: bool = False, dt: float = 1800.0, w_max_cfl: float = 1.0) -> OceanState:
    """Streamfunction velocities (x level-wise perturbation), layer thicknesses, continuity ``w``."""
    L, nl, Nh, T, E = m.L, m.nl, m.Nh, m.T, m.E
    emask, nmask = layer_masks(m, device)
    zbar = _t(m.zbar, device); Zmid = _t(m.Z, device)
    nln = _t(m.nlevels_nod2D, device, torch.int64)
    lev = torch.arange(1, nl + 1, device=device)

    helem = torch.where(emask, (zbar[:-1] - zbar[1:])[None, :].expand(T, L), torch.zeros((), dtype=F64, device=device))
    zbar_3d = torch.where(lev[None, :] <= nln[:, None], zbar[None, :].expand(Nh, nl), torch.zeros((), dtype=F64, device=device))
    Z_3d = torch.where(nmask, Zmid[None, :].expand(Nh, L), torch.zeros((), dtype=F64, device=device))
    hnode = torch.where(nmask, zbar_3d[:, :-1] - zbar_3d[:, 1:], torch.zeros((), dtype=F64, device=device))
    hnode_new = hnode.clone()
    coord = _t(m.coord_nod2D, device)
    if ale_amp != 0.0:                      # zlevel-like: only the surface layer breathes
        eta = ale_amp * torch.sin(5.0 * coord[:, 0]) * torch.cos(3.0 * coord[:, 1])
        hnode_new[:, 0] = hnode[:, 0] + eta
    zbar_n_bot = zbar[(nln - 1).clamp(min=0)]

    # --- velocities: u = -dpsi/dy, v = dpsi/dx of psi = A sin(kx x) sin(ky y), metres via r_earth
    en = _t(m.elem2D_nodes, device, torch.int64) - 1
    cx = coord[en, 0].mean(1); cy = coord[en, 1].mean(1)
    gid = _t(m.myList_elem2D, device, torch.int64)
    kx, ky = 6.0, 4.0
    u0 = -vel_amp * torch.sin(kx * cx) * torch.cos(ky * cy)
    v0 = vel_amp * torch.cos(kx * cx) * torch.sin(ky * cy) * (kx / ky)
    pert = 1.0 + 0.3 * (2.0 * _hash01(gid[:, None] * 131 + torch.arange(L, device=device)[None, :], 7) - 1.0)
    decay = torch.exp(Zmid / 1500.0)[None, :]
    uv = torch.stack([u0[:, None] * pert * decay, v0[:, None] * pert * decay], dim=2)
    uv = torch.where(emask[:, :, None], uv, torch.zeros((), dtype=F64, device=device)).contiguous()

    w = continuity_w(m, uv, helem, device)
    w_e, w_i = w.clone(), torch.zeros_like(w)
    if use_wsplit:
        # compute_Wvel_split (oce_ale.F90:3001-3049): the part of w above a CFL threshold goes
        # implicit.  c1 = 1 where cfl_z <= w_max_cfl, else w_max_cfl/cfl_z
        hn = torch.cat([hnode_new, hnode_new[:, -1:]], dim=1).clamp(min=1e-30)
        cfl = w.abs() * dt / hn
        c = torch.where(cfl > w_max_cfl, w_max_cfl / cfl.clamp(min=1e-300), torch.ones_like(cfl))
        w_e = w * c
        w_i = w * (1.0 - c)
    return OceanState(uv=uv, w=w, w_e=w_e, w_i=w_i, helem=helem.contiguous(), hnode=hnode.contiguous(),
                      hnode_new=hnode_new.contiguous(), zbar_3d_n=zbar_3d.contiguous(),
                      Z_3d_n=Z_3d.contiguous(), zbar_n_bot=zbar_n_bot.contiguous(), use_wsplit=use_wsplit)


def continuity_w(m: Mesh, uv: torch.Tensor, helem: torch.Tensor, device) -> torch.Tensor:
    """vert_vel_ale core (oce_ale.F90:2164-2310): scatter edge transports, cumulative sum
    bottom-up, divide by area.  Complete for owned nodes of a local mesh."""
    L, nl, Nh = m.L, m.nl, m.Nh
    emask, nmask = layer_masks(m, device)
    ed = _t(m.edges, device, torch.int64) - 1
    et = _t(m.edge_tri, device, torch.int64) - 1
    cross = _t(m.edge_cross_dxdy, device)
    div = torch.zeros((Nh, nl), dtype=F64, device=device)
    e1 = et[:, 0]
    c1 = (uv[e1, :, 1] * cross[:, 0:1] - uv[e1, :, 0] * cross[:, 1:2]) * helem[e1]
    c1 = torch.where(emask[e1], c1, torch.zeros((), dtype=F64, device=device))
    has2 = et[:, 1] >= 0
    e2 = et[:, 1].clamp(min=0)
    c2 = -(uv[e2, :, 1] * cross[:, 2:3] - uv[e2, :, 0] * cross[:, 3:4]) * helem[e2]
    c2 = torch.where(emask[e2] & has2[:, None], c2, torch.zeros((), dtype=F64, device=device))
    c = c1 + c2
    div[:, :L].index_add_(0, ed[:, 0], c)
    div[:, :L].index_add_(0, ed[:, 1], -c)
    div[:, :L] = torch.where(nmask, div[:, :L], torch.zeros((), dtype=F64, device=device))
    wflux = torch.flip(torch.cumsum(torch.flip(div, dims=[1]), dim=1), dims=[1])
    area = _t(m.area, device)
    w = torch.zeros_like(wflux)
    w[:, :L] = torch.where(nmask, wflux[:, :L] / area[:, :L].clamp(min=1e-30), torch.zeros((), dtype=F64, device=device))
    return w.contiguous()


def make_tracer_values(m: Mesh, device="cpu", kind: int = 0, noise: float = 0.1):
    """(values, valuesold): kind 0 = temperature-like, 1 = salinity-like, k>=2 = T*(1+0.01k)+k."""
    emask, nmask = layer_masks(m, device)
    coord = _t(m.coord_nod2D, device)
    Zmid = _t(m.Z, device)
    gid = _t(m.myList_nod2D, device, torch.int64)
    x, y = coord[:, 0:1], coord[:, 1:2]
    lev = torch.arange(m.L, device=device)[None, :]

    def temp(shift):
        base = 10.0 + 10.0 * torch.exp(Zmid / 1000.0)[None, :] + 2.0 * torch.sin(3.0 * x + shift) * torch.cos(2.0 * y)
        return base + noise * _hash01(gid[:, None] * 257 + lev, 11)

    def salt(shift):
        base = 35.0 - 1.5 * torch.exp(Zmid / 400.0)[None, :] + 0.5 * torch.cos(2.0 * x - shift) * torch.sin(3.0 * y)
        return base + 0.2 * noise * _hash01(gid[:, None] * 263 + lev, 13)

    if kind == 0:
        v, vo = temp(0.0), temp(0.01)
    elif kind == 1:
        v, vo = salt(0.0), salt(0.01)
    else:
        v, vo = temp(0.0) * (1.0 + 0.01 * kind) + kind, temp(0.01) * (1.0 + 0.01 * kind) + kind
    zero = torch.zeros((), dtype=F64, device=device)
    return torch.where(nmask, v, zero).contiguous(), torch.where(nmask, vo, zero).contiguous()


def ab2(values, valuesold):
    """oce_tracer_mod.F90:52-54."""
    return -(0.5 + EPSILON_AB) * valuesold + (1.5 + EPSILON_AB) * values


def tracer_gradient_elements(m: Mesh, ttf: torch.Tensor, device) -> torch.Tensor:
    """tr_xy(1:2,nz,elem) (oce_tracer_mod.F90:173-182) -> (T, L, 2)."""
    emask, _ = layer_masks(m, device)
    en = _t(m.elem2D_nodes, device, torch.int64) - 1
    g = _t(m.gradient_sca, device)
    tx = sum(g[:, k:k + 1] * ttf[en[:, k]] for k in range(3))
    ty = sum(g[:, k + 3:k + 4] * ttf[en[:, k]] for k in range(3))
    z = torch.zeros((), dtype=F64, device=device)
    return torch.stack([torch.where(emask, tx, z), torch.where(emask, ty, z)], dim=2)


def _find_up_downwind_triangles_torch(m: Mesh, device) -> np.ndarray:
    """The same search with torch ops on ``device`` (large synthetic meshes: seconds instead of a minute).  atan2
    of the device may differ from libm in the last bit, which only matters where the edge direction lies exactly on
    an element edge (the reference's `ab == ax` tie, see tests/test_oracle_pinning.py): a synthetic-input generator,
    not the pinned restatement."""
    cyc = float(m.cyclic_length)
    coord = _t(m.coord_nod2D, device)
    ed = _t(m.edges, device, torch.int64) - 1
    en = _t(m.elem2D_nodes, device, torch.int64) - 1
    nie = _t(m.nod_in_elem2D, device, torch.int64)
    num = _t(m.nod_in_elem2D_num, device, torch.int64)
    out = torch.zeros((m.E, 2), dtype=torch.int32, device=device)

    def trim(v):
        v = torch.where(v > cyc / 2, v - cyc, v)
        return torch.where(v < -cyc / 2, v + cyc, v)

    xvec = coord[ed[:, 1]] - coord[ed[:, 0]]
    xvec = torch.stack([trim(xvec[:, 0]), xvec[:, 1]], dim=1)
    rows = torch.arange(m.E, device=device)
    for side in (0, 1):
        node = ed[:, side]
        x = -xvec if side == 0 else xvec
        res = torch.zeros(m.E, dtype=torch.int32, device=device)
        for k in range(nie.shape[1]):
            have = k < num[node]
            elem = torch.where(have, nie[node, k] - 1, torch.zeros_like(node))
            nodes = en[elem]
            pos = torch.where(nodes[:, 0] == node, 0, torch.where(nodes[:, 1] == node, 1, 2))
            bi = torch.where(pos == 0, 1, 0)
            ci = torch.where(pos == 2, 1, 2)
            p0 = coord[nodes[rows, pos]]
            b = coord[nodes[rows, bi]] - p0
            c = coord[nodes[rows, ci]] - p0
            b0, c0 = trim(b[:, 0]), trim(c[:, 0])
            b1, c1 = b[:, 1], c[:, 1]
            cr = c0 * c0 + c1 * c1
            ab = torch.atan2((-b0 * c1 + b1 * c0) / cr, (b0 * c0 + b1 * c1) / cr)
            ax = torch.atan2((-x[:, 0] * c1 + x[:, 1] * c0) / cr, (x[:, 0] * c0 + x[:, 1] * c1) / cr)
            hit = (((ab > 0) & (ax > 0) & (ax < ab)) | ((ab < 0) & (ax < 0) & (ax > ab)) | (ab == ax) | (ax == 0)) & have
            res = torch.where(hit, (elem + 1).to(torch.int32), res)
        out[:, side] = res
    return out.cpu().numpy()


def find_up_downwind_triangles(m: Mesh, device=None) -> np.ndarray:
    """edge_up_dn_tri(2,E), 1-based, 0 = none (oce_muscl_adv.F90:240-333).  The reference's loop
    keeps the LAST element of nod_in_elem2D that satisfies the test.  ``device``: evaluate with torch there
    (input generation for the large bench meshes); default: NumPy, the version pinned against the C restatement."""
    if device is not None:
        return _find_up_downwind_triangles_torch(m, device)
    cyc = m.cyclic_length
    coord = m.coord_nod2D
    ed = m.edges.astype(np.int64) - 1
    en = m.elem2D_nodes.astype(np.int64) - 1
    out = np.zeros((m.E, 2), np.int32)
    xvec = coord[ed[:, 1]] - coord[ed[:, 0]]
    xvec[:, 0] = np.where(xvec[:, 0] > cyc / 2, xvec[:, 0] - cyc, xvec[:, 0])
    xvec[:, 0] = np.where(xvec[:, 0] < -cyc / 2, xvec[:, 0] + cyc, xvec[:, 0])
    maxdeg = m.nod_in_elem2D.shape[1]

    def trim(v):
        v = np.where(v > cyc / 2, v - cyc, v)
        return np.where(v < -cyc / 2, v + cyc, v)

    for side in (0, 1):
        node = ed[:, side]
        x = -xvec if side == 0 else xvec
        for k in range(maxdeg):
            have = k < m.nod_in_elem2D_num[node]
            elem = np.where(have, m.nod_in_elem2D[node, k] - 1, 0).astype(np.int64)
            nodes = en[elem]                                             # (E, 3)
            pos = np.where(nodes[:, 0] == node, 0, np.where(nodes[:, 1] == node, 1, 2))
            # (b, c) node choice per position: pos0 -> (2,3)-(1); pos1 -> (1,3)-(2); pos2 -> (1,2)-(3)
            bi = np.where(pos == 0, 1, 0)
            ci = np.where(pos == 2, 1, 2)
            rows = np.arange(m.E)
            p0 = coord[nodes[rows, pos]]
            b = coord[nodes[rows, bi]] - p0
            c = coord[nodes[rows, ci]] - p0
            b[:, 0] = trim(b[:, 0]); c[:, 0] = trim(c[:, 0])
            cr = (c * c).sum(1)
            bx = (b * c).sum(1) / cr
            by = (-b[:, 0] * c[:, 1] + b[:, 1] * c[:, 0]) / cr
            xx = (x * c).sum(1) / cr
            xy = (-x[:, 0] * c[:, 1] + x[:, 1] * c[:, 0]) / cr
            ab = np.arctan2(by, bx)
            ax = np.arctan2(xy, xx)
            hit = ((ab > 0) & (ax > 0) & (ax < ab)) | ((ab < 0) & (ax < 0) & (ax > ab)) | (ab == ax) | (ax == 0)
            hit &= have
            out[hit, side] = (elem[hit] + 1).astype(np.int32)
    return out


def fill_up_dn_grad(m: Mesh, tr_xy: torch.Tensor, up_dn_tri: np.ndarray, device) -> torch.Tensor:
    """edge_up_dn_grad(4,L,E) -> (E, L, 4) (oce_muscl_adv.F90:378-522)."""
    L, Nh, E = m.L, m.Nh, m.E
    emask, nmask = layer_masks(m, device)
    z = torch.zeros((), dtype=F64, device=device)
    # area-weighted node-mean gradient over the valid elements around a node
    en = _t(m.elem2D_nodes, device, torch.int64) - 1
    ea = _t(m.elem_area, device)
    wgt = torch.where(emask, ea[:, None].expand(-1, L), z)
    tvol = torch.zeros((Nh, L), dtype=F64, device=device)
    tx = torch.zeros((Nh, L), dtype=F64, device=device)
    ty = torch.zeros((Nh, L), dtype=F64, device=device)
    for k in range(3):
        tvol.index_add_(0, en[:, k], wgt)
        tx.index_add_(0, en[:, k], tr_xy[:, :, 0] * wgt)
        ty.index_add_(0, en[:, k], tr_xy[:, :, 1] * wgt)
    tv = tvol.clamp(min=1e-300)
    gx, gy = tx / tv, ty / tv

    ed = _t(m.edges, device, torch.int64) - 1
    tri = _t(up_dn_tri, device, torch.int64) - 1
    n1, n2 = ed[:, 0], ed[:, 1]
    both = (tri[:, 0] >= 0) & (tri[:, 1] >= 0)
    nmin = _t(m.nlevels_nod2D_min, device, torch.int64); umax = _t(m.ulevels_nod2D_max, device, torch.int64)
    nln = _t(m.nlevels_nod2D, device, torch.int64); uln = _t(m.ulevels_nod2D, device, torch.int64)
    lev = torch.arange(1, L + 1, device=device)[None, :]
    nzmin = torch.maximum(umax[n1], umax[n2])[:, None]
    nzmax = torch.minimum(nmin[n1], nmin[n2])[:, None]
    shared = both[:, None] & (lev >= nzmin) & (lev <= nzmax - 1)
    valid1 = (lev >= uln[n1][:, None]) & (lev <= nln[n1][:, None] - 1)
    valid2 = (lev >= uln[n2][:, None]) & (lev <= nln[n2][:, None] - 1)
    own1 = valid1 & ~shared       # levels filled from node-mean gradients at node 1
    own2 = valid2 & ~shared
    up = tri[:, 0].clamp(min=0); dn = tri[:, 1].clamp(min=0)
    g = torch.zeros((E, L, 4), dtype=F64, device=device)
    g[:, :, 0] = torch.where(shared, tr_xy[up, :, 0], torch.where(own1, gx[n1], z))
    g[:, :, 1] = torch.where(shared, tr_xy[dn, :, 0], torch.where(own2, gx[n2], z))
    g[:, :, 2] = torch.where(shared, tr_xy[up, :, 1], torch.where(own1, gy[n1], z))
    g[:, :, 3] = torch.where(shared, tr_xy[dn, :, 1], torch.where(own2, gy[n2], z))
    return g.contiguous()


def make_tracers(m: Mesh, ntr: int = 2, device="cpu", hor="MFCT", ver="QR4C", lim="FCT",
                 ph: float = 0.0, pv: float = 1.0, up_dn_tri: Optional[np.ndarray] = None) -> List[TracerFields]:
    if up_dn_tri is None:
        up_dn_tri = find_up_downwind_triangles(m)
    out = []
    for k in range(ntr):
        v, vo = make_tracer_values(m, device, kind=k)
        vab = ab2(v, vo).contiguous()
        tr_xy = tracer_gradient_elements(m, v, device)
        g = fill_up_dn_grad(m, tr_xy, up_dn_tri, device)
        out.append(TracerFields(values=v, valuesAB=vab, edge_up_dn_grad=g, tra_adv_hor=hor,
                                tra_adv_ver=ver, tra_adv_lim=lim, tra_adv_ph=ph, tra_adv_pv=pv))
    return out


def cfl_dt(m: Mesh, st: OceanState, cfl: float = 0.3) -> float:
    """A time step with horizontal+vertical Courant number <= ``cfl`` (config 3-5: 'dt from CFL')."""
    emask, nmask = layer_masks(m, st.uv.device)
    speed = torch.sqrt(st.uv[:, :, 0] ** 2 + st.uv[:, :, 1] ** 2).max(dim=1).values
    dx = torch.sqrt(torch.as_tensor(m.elem_area, device=st.uv.device))
    dt_h = (dx / speed.clamp(min=1e-12)).min().item()
    hn = torch.where(nmask, st.hnode_new, torch.full((), 1e30, dtype=F64, device=st.uv.device))
    wmax = torch.maximum(st.w[:, :-1].abs(), st.w[:, 1:].abs()).clamp(min=1e-12)
    dt_v = (hn / wmax).min().item()
    return cfl * min(dt_h, dt_v)


def scatter_to_local(g: Mesh, loc: Mesh, st: OceanState, trs: List[TracerFields]):
    """Restrict globally generated fields to a local (partitioned) mesh via the myList_* maps.
    This is what the reference gets from reading restarts + exchange_nod/exchange_elem."""
    dev = st.uv.device
    ni = torch.as_tensor(loc.myList_nod2D.astype(np.int64) - 1, device=dev)
    ei = torch.as_tensor(loc.myList_elem2D.astype(np.int64) - 1, device=dev)
    di = torch.as_tensor(loc.myList_edge2D.astype(np.int64) - 1, device=dev)
    lst = OceanState(uv=st.uv[ei].contiguous(), w=st.w[ni].contiguous(), w_e=st.w_e[ni].contiguous(),
                     w_i=st.w_i[ni].contiguous(), helem=st.helem[ei].contiguous(), hnode=st.hnode[ni].contiguous(),
                     hnode_new=st.hnode_new[ni].contiguous(), zbar_3d_n=st.zbar_3d_n[ni].contiguous(),
                     Z_3d_n=st.Z_3d_n[ni].contiguous(), zbar_n_bot=st.zbar_n_bot[ni].contiguous(),
                     use_wsplit=st.use_wsplit)
    ltr = [TracerFields(values=t.values[ni].contiguous(), valuesAB=t.valuesAB[ni].contiguous(),
                        edge_up_dn_grad=t.edge_up_dn_grad[di].contiguous(), tra_adv_hor=t.tra_adv_hor,
                        tra_adv_ver=t.tra_adv_ver, tra_adv_lim=t.tra_adv_lim, tra_adv_ph=t.tra_adv_ph,
                        tra_adv_pv=t.tra_adv_pv) for t in trs]
    return lst, ltr


def make_tracers_kind(m: Mesh, kind: int, device, up_dn_tri: np.ndarray, hor="MFCT", ver="QR4C", lim="FCT",
                      ph: float = 0.0, pv: float = 1.0) -> List[TracerFields]:
    """One tracer of the given kind (0 = T, 1 = S, k>=2 = T*(1+0.01k)+k) as a one-element list."""
    v, vo = make_tracer_values(m, device, kind=kind)
    vab = ab2(v, vo).contiguous()
    tr_xy = tracer_gradient_elements(m, v, device)
    g = fill_up_dn_grad(m, tr_xy, up_dn_tri, device)
    del tr_xy, vo
    return [TracerFields(values=v, valuesAB=vab, edge_up_dn_grad=g, tra_adv_hor=hor, tra_adv_ver=ver,
                         tra_adv_lim=lim, tra_adv_ph=ph, tra_adv_pv=pv)]
