"""Shared helpers of the parity tests: build inputs once, run the oracle, run the CUDA path."""
from __future__ import annotations

import numpy as np
import torch

from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

TOL_STEP = 1e-12      # north_star: 1e-12 relative per step
TOL_100 = 1e-10       # and 1e-10 after 100 steps


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    """Appendix C of SURVEY.md: max|x - x_ref| / max|x_ref|."""
    den = max(float(np.abs(ref).max()), 1e-300)
    return float(np.abs(got - ref).max()) / den


def make_case(mesh, ntr=2, hor="MFCT", ver="QR4C", lim="FCT", ph=0.0, pv=1.0, use_wsplit=False,
              ale_amp=0.0, cfl=0.3, dt=None, kinds=None):
    st = F.make_state(mesh, "cpu", use_wsplit=use_wsplit, ale_amp=ale_amp, dt=dt or 1800.0)
    if dt is None:
        dt = F.cfl_dt(mesh, st, cfl)
        if use_wsplit:
            st = F.make_state(mesh, "cpu", use_wsplit=True, ale_amp=ale_amp, dt=dt, w_max_cfl=0.1)
    trs = F.make_tracers(mesh, ntr, "cpu", hor=hor, ver=ver, lim=lim, ph=ph, pv=pv)
    nb = M.nboundary_lay(mesh)
    return st, trs, nb, dt


def to_device(st, trs, dev):
    st_d = F.OceanState(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in st.__dict__.items()})
    trs_d = [F.TracerFields(values=t.values.to(dev), valuesAB=t.valuesAB.to(dev),
                            edge_up_dn_grad=t.edge_up_dn_grad.to(dev), tra_adv_hor=t.tra_adv_hor,
                            tra_adv_ver=t.tra_adv_ver, tra_adv_lim=t.tra_adv_lim,
                            tra_adv_ph=t.tra_adv_ph, tra_adv_pv=t.tra_adv_pv) for t in trs]
    return st_d, trs_d


def run_oracle(mesh, st, trs, nb, dt, nsteps=1, mode=0):
    from oracle import oracle_py as O
    rk = O.OracleRank(mesh, st, trs, nb)
    O.run([rk], dt, nsteps, mode)
    return rk


def run_cuda(mesh, st, trs, nb, dt, device=0, host_ptrs=False):
    from fesom2_b200.driver import AdvB200
    ctx = AdvB200(mesh, nb, device=device, max_tracers=len(trs))
    if host_ptrs:
        st_d, trs_d = st, trs
        dev = "cpu"
    else:
        dev = torch.device(f"cuda:{device}")
        st_d, trs_d = to_device(st, trs, dev)
    dh = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx.set_state(st_d)
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
    return ctx, [x.cpu().numpy() for x in dh], [x.cpu().numpy() for x in dv]


def zlevel_case(mesh, st, lz=4, min_hnode=0.5):
    """Inputs that drive every branch of the reference's which_ALE = 'zlevel' correction (src/oce_ale.F90:2336-2538):
    modifies st.hnode in place (squeezed subsurface layers in a quarter of the columns) and returns
    (hbar, hbar_old, water_flux, cfl_z_old).  Column classes by node id mod 4: 0 small change (plain zlevel),
    1 a drop that empties the surface layer (local zstar over several layers, some of them already at their
    minimum or CFL-limited), 2 a rise over squeezed subsurface layers (refill), 3 a rise / small drop over rest layers."""
    Nh = mesh.Nh
    ids = np.arange(Nh, dtype=np.float64)
    cls = np.arange(Nh) % 4
    zb = np.asarray(mesh.zbar, dtype=np.float64)
    rest = zb[:lz] - zb[1:lz + 1]
    h = st.hnode.numpy()
    sq = 0.55 + 0.4 * np.abs(np.sin(0.3 * ids))                   # 0.55 .. 0.95 of the rest thickness
    for k in range(1, lz):
        col = (cls == 2) | ((cls == 1) & (np.arange(Nh) % 3 == k % 3))
        h[col, k] = (rest[k] * sq * (1.0 - 0.05 * k))[col]
    hbar_old = 0.05 * np.sin(0.37 * ids)
    dh = 0.01 * np.cos(0.11 * ids)
    dh = np.where(cls == 1, -rest[0] * (0.55 + 1.2 * np.abs(np.sin(0.23 * ids))), dh)
    dh = np.where(cls == 2, rest[0] * (0.02 + 0.5 * np.abs(np.sin(0.19 * ids))), dh)
    dh = np.where(cls == 3, 0.3 * np.sin(0.41 * ids), dh)
    wflux = 1.0e-6 * np.sin(0.05 * ids)
    cfl_old = 1.1 * np.abs(np.sin(0.7 * ids[:, None] + np.arange(mesh.nl)[None, :]))
    return hbar_old + dh, hbar_old, wflux, cfl_old
