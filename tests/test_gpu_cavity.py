"""-m gpu: the reference's cavity mesh test/meshes/pi_cavity (use_cavity: 170 elements and 95 nodes start below layer 1,
areasvol of the cells under the ice is the LOWER face, oce_mesh.F90:2299-2321) through every entry point of the C ABI,
bit for bit against the C restatement.  Cavities exercise what no other fixture does: nzmin > 1 in every vertical
stencil, edge ranges A / B of the flux routines, padded FCT clusters above the shallowest element, boundary edges whose
volume flux starts at layer 1, columns vert_vel_ale's zstar / zlevel corrections must skip."""
import numpy as np
import pytest
import torch

from common import make_case, run_cuda, run_oracle, to_device, zlevel_case
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True),
                                                ("MFCT", "QR4C", "NON", False), ("UPW1", "UPW1", "FCT", False),
                                                ("MUSCL", "CDIFF", "FCT", False), ("MUSCL", "PPM", "NON", False)])
def test_cavity_mesh_matches_the_oracle(cav_mesh, hor, ver, lim, wsplit):
    g = cav_mesh
    assert (g.ulevels > 1).sum() == 170 and (g.ulevels_nod2D > 1).sum() == 95
    st, trs, nb, dt = make_case(g, 3, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    ora = run_oracle(g, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(g, st, trs, nb, dt)
    cav = g.ulevels_nod2D > 1
    for k in range(3):
        assert np.isfinite(dh[k]).all() and np.isfinite(dv[k]).all()
        assert np.array_equal(dh[k], ora.dttf_h[k]), (k, np.abs(dh[k] - ora.dttf_h[k]).max())
        assert np.array_equal(dv[k], ora.dttf_v[k]), (k, np.abs(dv[k] - ora.dttf_v[k]).max())
        assert np.abs(dh[k][cav]).max() > 0
    if lim == "FCT":
        N = g.N
        for name in ("fct_LO", "fct_plus", "fct_minus"):
            assert np.array_equal(ctx.get_work(name, 2)[:N], ora.keep[name][:N]), name
    ctx.close()


def test_cavity_device_gradients_and_diagnostics(cav_mesh):
    """edge_up_dn_grad = NULL (tracer_gradient_elements + fill_up_dn_grad fused into the edge kernel) and ltra_diag on
    the cavity mesh"""
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200
    g = cav_mesh
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    tri = F.find_up_downwind_triangles(g)
    for t in trs:
        t.edge_up_dn_grad = torch.as_tensor(O.fill_up_dn_grad(g, O.tracer_gradient_elements(g, t.values.numpy()), tri))
    rk = O.OracleRank(g, st, trs, nb, tra_diag=True)
    O.run([rk], dt)
    dev = torch.device("cuda:0")
    st_d, trs_d = to_device(st, trs, dev)
    for t in trs_d:
        t.edge_up_dn_grad = None
    ctx = AdvB200(g, nb, max_tracers=2)
    ctx.set_gradient_mesh(tri)
    ctx.set_state(st_d)
    z = lambda v=0.0: [torch.full((g.Nh, g.L), v, dtype=torch.float64, device=dev) for _ in trs]   # noqa: E731
    dh, dv, tah, tav = z(), z(), z(7.0), z(7.0)
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv, tra_advhoriz=tah, tra_advvert=tav)
    lev = np.arange(1, g.L + 1)[None, :]
    wet = (lev >= np.asarray(g.ulevels_nod2D)[:, None]) & (lev <= np.asarray(g.nlevels_nod2D)[:, None] - 1)
    for k in range(2):
        assert np.array_equal(dh[k].cpu().numpy(), rk.dttf_h[k]) and np.array_equal(dv[k].cpu().numpy(), rk.dttf_v[k])
        for got, ref in ((tah[k], rk.tra_advhoriz[k]), (tav[k], rk.tra_advvert[k])):
            got = got.cpu().numpy()
            assert np.array_equal(got[wet], ref[wet]) and (got[~wet] == 7.0).all()
    ctx.close()


def test_cavity_gradient_producers(cav_mesh):
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200
    g = cav_mesh
    st, trs, nb, dt = make_case(g, 2)
    tri = F.find_up_downwind_triangles(g)
    dev = torch.device("cuda:0")
    ctx = AdvB200(g, nb, max_tracers=2)
    ctx.set_gradient_mesh(tri)
    ttf = [t.values.to(dev) for t in trs]
    tr_xy = [torch.zeros((g.T, g.L, 2), dtype=torch.float64, device=dev) for _ in trs]
    grad = [torch.full((g.E, g.L, 4), -777.0, dtype=torch.float64, device=dev) for _ in trs]
    ctx.tracer_gradient_elements(ttf, tr_xy)
    ctx.fill_up_dn_grad(tr_xy, grad)
    ctx.synchronize()
    for k in range(2):
        ref_xy = O.tracer_gradient_elements(g, trs[k].values.numpy())
        ref_g = O.fill_up_dn_grad(g, ref_xy, tri, out=np.full((g.E, g.L, 4), -777.0))
        assert np.array_equal(tr_xy[k].cpu().numpy(), ref_xy)
        got = grad[k].cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref_g)) and np.array_equal(np.nan_to_num(got), np.nan_to_num(ref_g))
    ctx.close()


@pytest.mark.parametrize("ale", ["linfs", "zstar", "zlevel"])
def test_cavity_vert_vel_ale(cav_mesh, ale):
    """the continuity part under the ice (nzmin > 1) and the free-surface corrections, which skip cavity columns
    (src/oce_ale.F90:2354, :2550)"""
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200
    g = cav_mesh
    st, trs, nb, dt = make_case(g, 1)
    dtc = 40.0 * dt
    lz, mh = 4, 0.5
    hbar, hbar_old, wflux, cfl_old = zlevel_case(g, st, lz, mh)
    rk = O.OracleRank(g, st, trs, nb)
    W = O.vert_vel_ale_core(rk)
    hn = st.hnode_new.numpy().copy()
    if ale == "zstar":
        W, hn = O.vert_vel_ale_zstar(rk, dtc, W, hbar, hbar_old, wflux)
    elif ale == "zlevel":
        W, hn = O.vert_vel_ale_zlevel(rk, dtc, W, hbar, hbar_old, wflux, g.zbar, cfl_old, mh, lz)
    rk.keep["hnode_new"][...] = hn
    cfl, we, wi = O.compute_cflz_and_split(rk, dtc, W, True, 0.5)
    dev = torch.device("cuda:0")
    st_d, _ = to_device(st, [], dev)
    ctx = AdvB200(g, nb, max_tracers=1)
    ctx.set_state(st_d)
    t = lambda a, dt_=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt_, device=dev)   # noqa: E731
    out = [torch.zeros((g.Nh, g.nl), dtype=torch.float64, device=dev) for _ in range(3)] + [t(cfl_old)]
    nmin = t(g.nlevels_nod2D_min, torch.int32)
    if ale == "linfs":
        ctx.vert_vel_ale(dtc, True, 0.5, *out)
    elif ale == "zstar":
        ctx.vert_vel_ale_zstar(dtc, True, 0.5, t(hbar), t(hbar_old), t(wflux), nmin, st_d.hnode_new, *out)
    else:
        ctx.vert_vel_ale_zlevel(dtc, True, 0.5, t(hbar), t(hbar_old), t(wflux), nmin, st_d.hnode_new, t(g.zbar), mh, lz, *out)
    ctx.synchronize()
    for got, ref, name in zip(out, (W, we, wi, cfl), ("w", "w_e", "w_i", "cfl_z")):
        assert np.array_equal(got.cpu().numpy(), ref), (ale, name)
    assert np.array_equal(st_d.hnode_new.cpu().numpy(), hn)
    cav = np.asarray(g.ulevels_nod2D)[:g.N] > 1
    assert np.array_equal(hn[:g.N][cav], st.hnode_new.numpy()[:g.N][cav])          # cavity columns keep their thickness
    ctx.close()


def test_cavity_two_local_ranks(cav_mesh):
    """the reference's own dist_2 partition of pi_cavity: owned nodes of both ranks = the one-rank oracle"""
    from local_ranks import run_local_ranks
    g = cav_mesh
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    one = run_oracle(g, st, trs, nb, dt)
    res = run_local_ranks(g, g.parts[2], st, trs, dt, exchange_inputs=True)
    for r in res:
        own = r["owned"].astype(np.int64) - 1
        n = r["N"]
        for k in range(2):
            assert np.array_equal(r["dv"][k][:n], one.dttf_v[k][own])
            assert np.array_equal(r["dh"][k][:n], one.dttf_h[k][own])
