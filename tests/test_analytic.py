"""Analytic pins of the restated formulas -- independent of any restatement: what the schemes MUST give on fields for
which the continuous answer is known.

* Horizontal: the edge-based median-dual discretisation with the tracer evaluated at the edge midpoint is exact for a
  horizontally LINEAR tracer in a constant velocity field on any triangulation (Green-Gauss on the median dual), and the
  MUSCL / MFCT reconstruction (src/oce_adv_tra_hor.F90:446-461, :736-751) reduces to exactly that midpoint value when the
  upwind / downwind gradients are the true gradient: Tmean = T(n) + [2 (T2 - T1) + dx . g] / 6 = T(n) + (T2 - T1) / 2.  So
  del_ttf_advhoriz = -dt * h * (u . grad T) on every interior node, for every blending order num_ord -- this pins the signs,
  the 2 / 6 weights, the metric factors of edge_dxdy, the volume flux and the division by areasvol at once.
* Vertical: for a tracer LINEAR in depth on uniform layers and constant w, QR4C (src/oce_adv_tra_ver.F90:417-424), the
  centred scheme and PPM all reconstruct the exact interface value, so del_ttf_advvert = -dt * h * w * dT/dz on the layers
  whose two interfaces use the interior stencil.

The C restatement and the NumPy restatement are both checked here; the CUDA path is bit-identical to the former
(tests/test_gpu_parity.py::test_analytic_linear_fields)."""
import numpy as np
import pytest
import torch

from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

R_EARTH = 6367500.0
U0, V0, W0 = 0.31, -0.17, 2.0e-5
GA, GB, GC = 3.0e-6, -2.0e-6, 4.0e-3          # dT/dx, dT/dy per metre; dT/dz per metre


def linear_case(hor="MFCT", ver="QR4C", ph=0.0, pv=1.0, vertical=False, nx=17, ny=13, nl=14, uniform_z=True):
    """flat-bottom cartesian patch with uniform (or stretched) layers; returns (mesh, state, [tracer], nboundary_lay, dt)"""
    g = M.synth_mesh(nx, ny, nl=nl, lon0=-2.0, lon1=2.0, lat0=-1.5, lat1=1.5, staircase=False, derive=False)
    g.cartesian = True
    if uniform_z:
        g.zbar = -np.linspace(0.0, 50.0 * (nl - 1), nl)        # uniform 50 m layers
    else:
        g.zbar = -np.cumsum(np.concatenate([[0.0], 10.0 * 1.35 ** np.arange(nl - 1)]))   # 10 m at the top, growing by 35 % per layer
    if not vertical:
        # an IRREGULAR triangulation: every interior node moved by up to 30 % of the grid spacing (on a uniform grid even
        # first-order upwind is exact for a linear field, by symmetry, and the test could not fail)
        rng = np.random.default_rng(20261017)
        ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        inner = ((ii > 0) & (ii < nx - 1) & (jj > 0) & (jj < ny - 1)).ravel()
        dx, dy = 4.0 / (nx - 1) * M.RAD, 3.0 / (ny - 1) * M.RAD
        g.coord_nod2D[inner, 0] += 0.3 * dx * rng.uniform(-1, 1, inner.sum())
        g.coord_nod2D[inner, 1] += 0.3 * dy * rng.uniform(-1, 1, inner.sum())
    M.derive_geometry(g)
    st = F.make_state(g, "cpu")
    L = g.L
    h = (g.zbar[:-1] - g.zbar[1:])[None, :]
    st.helem[...] = torch.as_tensor(np.broadcast_to(h, (g.T, L)).copy())
    st.hnode[...] = torch.as_tensor(np.broadcast_to(h, (g.Nh, L)).copy())
    st.hnode_new[...] = st.hnode
    st.zbar_3d_n[...] = torch.as_tensor(np.broadcast_to(g.zbar[None, :], (g.Nh, g.nl)).copy())
    st.Z_3d_n[...] = torch.as_tensor(np.broadcast_to(g.Z[None, :], (g.Nh, L)).copy())
    st.uv[...] = 0.0
    for name in ("w", "w_e", "w_i"):
        getattr(st, name)[...] = 0.0
    x, y = g.coord_nod2D[:, 0] * R_EARTH, g.coord_nod2D[:, 1] * R_EARTH
    if vertical:
        st.w[...] = W0
        st.w_e[...] = W0
        T = GC * np.broadcast_to(g.Z[None, :], (g.Nh, L))
        grad = np.zeros((g.E, L, 4))
    else:
        st.uv[..., 0] = U0
        st.uv[..., 1] = V0
        T = np.broadcast_to((GA * x + GB * y)[:, None], (g.Nh, L))
        grad = np.zeros((g.E, L, 4))
        grad[..., 0] = grad[..., 1] = GA                        # (gx_up, gx_dn, gy_up, gy_dn), oce_adv_tra_hor.F90:431-434
        grad[..., 2] = grad[..., 3] = GB
    T = 10.0 + T
    tr = F.TracerFields(values=torch.as_tensor(T.copy()), valuesAB=torch.as_tensor(T.copy()), edge_up_dn_grad=torch.as_tensor(grad),
                        tra_adv_hor=hor, tra_adv_ver=ver, tra_adv_lim="NON", tra_adv_ph=ph, tra_adv_pv=pv)
    return g, st, [tr], M.nboundary_lay(g), 600.0


def interior_nodes(g, rings=1):
    """nodes at least `rings` edges away from the lateral boundary"""
    bnd = g.edge_tri[:, 1] <= 0
    out = np.zeros(g.Nh, bool)
    out[g.edges[bnd].ravel() - 1] = True
    for _ in range(rings - 1):
        grow = out.copy()
        a, b = g.edges[:, 0] - 1, g.edges[:, 1] - 1
        grow[a[out[b]]] = True
        grow[b[out[a]]] = True
        out = grow
    return ~out


def _both(g, st, trs, nb, dt):
    from oracle import numpy_ref as R, oracle_py as O
    rk = O.OracleRank(g, st, trs, nb)
    O.run([rk], dt)
    dh, dv = np.zeros((g.Nh, g.L)), np.zeros((g.Nh, g.L))
    R.NumpyAdv(g, st, nb).do_oce_adv_tra(dt, trs[0], dh, dv)
    return (rk.dttf_h[0], rk.dttf_v[0]), (dh, dv)


@pytest.mark.parametrize("hor", ["MFCT", "MUSCL"])
@pytest.mark.parametrize("ph", [0.0, 0.5, 1.0])
def test_horizontal_schemes_are_exact_for_a_linear_tracer(hor, ph):
    g, st, trs, nb, dt = linear_case(hor=hor, ver="UPW1", ph=ph)
    # MUSCL drops to low order at an edge end whose node touches the boundary (c_lo = 0, oce_adv_tra_hor.F90:411-412:
    # nboundary_lay = 0 there), so its exact region starts one ring further in
    inner = interior_nodes(g, rings=2 if hor == "MUSCL" else 1)
    assert inner.sum() > 80
    h = st.hnode.numpy()
    exact = -dt * h * (U0 * GA + V0 * GB)
    for dh, dv in _both(g, st, trs, nb, dt):
        ok = np.broadcast_to(inner[:, None], dh.shape)
        err = np.abs(dh - exact)[ok].max() / np.abs(exact).max()
        assert err <= 1e-11, (hor, ph, err)
        assert np.abs(dv).max() == 0.0                          # w = 0


def test_upwind_is_not_exact_so_the_test_can_fail():
    g, st, trs, nb, dt = linear_case(hor="UPW1", ver="UPW1")
    inner = interior_nodes(g)
    exact = -dt * st.hnode.numpy() * (U0 * GA + V0 * GB)
    (dh, _), _ = _both(g, st, trs, nb, dt)
    assert np.abs(dh - exact)[inner].max() / np.abs(exact).max() > 1e-3


@pytest.mark.parametrize("ver,pv", [("QR4C", 1.0), ("QR4C", 0.0), ("QR4C", 0.5), ("CDIFF", 1.0), ("PPM", 1.0)])
def test_vertical_schemes_are_exact_for_a_linear_profile(ver, pv):
    g, st, trs, nb, dt = linear_case(hor="UPW1", ver=ver, pv=pv, vertical=True)
    h = st.hnode.numpy()
    exact = -dt * h * W0 * GC
    lev = np.arange(1, g.L + 1)
    ok = (lev >= 3) & (lev <= g.L - 3)                           # both interfaces of the layer use the interior stencil
    for dh, dv in _both(g, st, trs, nb, dt):
        err = np.abs(dv - exact)[:, ok].max() / np.abs(exact).max()
        assert err <= 1e-10, (ver, pv, err)
        assert np.abs(dh).max() == 0.0                          # uv = 0


def test_gradient_producers_are_exact_for_a_linear_tracer():
    """tracer_gradient_elements (P1 gradients through gradient_sca, src/oce_tracer_mod.F90:146-188) of a linear field is the
    true gradient on every triangle of an irregular mesh, and fill_up_dn_grad (src/oce_muscl_adv.F90:356-525) -- whether
    it picks the upwind / downwind triangle or an area-weighted node mean -- hands exactly that to the flux routines"""
    from oracle import numpy_ref as R, oracle_py as O
    g, st, trs, nb, dt = linear_case()
    v = trs[0].values.numpy()
    tri = F.find_up_downwind_triangles(g)
    for mod in (O, R):
        xy = mod.tracer_gradient_elements(g, v)[:g.T]
        assert np.abs(xy[..., 0] - GA).max() <= 1e-10 * abs(GA) and np.abs(xy[..., 1] - GB).max() <= 1e-10 * abs(GB)
        gr = mod.fill_up_dn_grad(g, xy, tri)
        both = (tri[:, 0] != 0) & (tri[:, 1] != 0)
        assert both.sum() > 0.5 * g.E
        want = np.array([GA, GA, GB, GB])
        assert np.abs(gr[both] - want).max() <= 1e-10 * abs(GB)
        # and feeding them to the scheme reproduces the exact tendency
    trs[0].edge_up_dn_grad = torch.as_tensor(O.fill_up_dn_grad(g, O.tracer_gradient_elements(g, v), tri))
    inner = interior_nodes(g, rings=2)
    exact = -dt * st.hnode.numpy() * (U0 * GA + V0 * GB)
    (dh, _), _ = _both(g, st, trs, nb, dt)
    assert np.abs(dh - exact)[inner].max() / np.abs(exact).max() <= 1e-10


@pytest.mark.parametrize("hor,ver", [("MFCT", "QR4C"), ("MUSCL", "PPM")])
def test_fct_does_not_touch_a_linear_field(hor, ver):
    """low-order solution + limited antidiffusive fluxes (src/oce_adv_tra_fct.F90, src/oce_adv_tra_driver.F90:529-633): on a
    linear tracer the high-order update creates no new extremum, every limiter factor is 1, and the SUM of the two tendencies
    must again be the exact -dt h (u . grad T) -- which pins how the driver splits the update between del_ttf_advvert
    (-ttf hnode + LO hnode_new + vertical part) and del_ttf_advhoriz"""
    g, st, trs, nb, dt = linear_case(hor=hor, ver=ver)
    trs[0].tra_adv_lim = "FCT"
    inner = interior_nodes(g, rings=3)
    assert inner.sum() > 40
    exact = -dt * st.hnode.numpy() * (U0 * GA + V0 * GB)
    for dh, dv in _both(g, st, trs, nb, dt):
        err = np.abs(dh + dv - exact)[inner].max() / np.abs(exact).max()
        assert err <= 1e-9, (hor, ver, err)
        assert np.abs(dv).max() > 0                               # with FCT the low-order part sits in del_ttf_advvert


def test_continuity_of_a_uniform_flow_and_zstar_conservation():
    """vert_vel_ale (src/oce_ale.F90:2164-2310): the transports of a constant velocity through the closed boundary of a
    median-dual cell sum to zero, so W vanishes at every interior node of an irregular mesh (to round-off of the
    individual transports) -- a sign or metric slip in edge_cross_dxdy would leave O(u h / dx) there.  The zstar correction
    (:2539-2603) must hand the whole elevation change to the layers: sum_k (hnode_new - hnode) = hbar - hbar_old."""
    from oracle import numpy_ref as R, oracle_py as O
    g, st, trs, nb, dt = linear_case()
    rk = O.OracleRank(g, st, trs, nb)
    h = st.hnode.numpy()[0, 0]
    scale = np.hypot(U0, V0) * h / (4.0 / 16 * M.RAD * R_EARTH)            # u h / dx: what an unbalanced cell would show
    inner = interior_nodes(g)
    for W in (O.vert_vel_ale_core(rk), R.vert_vel_ale_core(g, st.uv.numpy(), st.helem.numpy())):
        assert np.abs(W[inner]).max() <= 1e-12 * scale * g.L
        assert np.abs(W[~inner]).max() > 1e-3 * scale                      # open cells at the wall are NOT balanced
    ids = np.arange(g.Nh, dtype=np.float64)
    hbar_old = 0.05 * np.sin(0.37 * ids)
    hbar = hbar_old + 0.2 * np.cos(0.11 * ids)
    W0_ = np.zeros((g.Nh, g.nl))
    for Wz, hn in (O.vert_vel_ale_zstar(rk, dt, W0_, hbar, hbar_old, np.zeros(g.Nh)),
                   R.vert_vel_ale_zstar(g, st.zbar_3d_n.numpy(), st.hnode.numpy(), st.hnode_new.numpy(), dt, W0_, hbar, hbar_old, np.zeros(g.Nh))):
        dsum = (hn - st.hnode.numpy())[:g.N].sum(axis=1)
        assert np.abs(dsum - (hbar - hbar_old)[:g.N]).max() <= 1e-12
        # and the surface velocity is the rate of that change: W(1) = -(hbar - hbar_old) / dt
        assert np.abs(Wz[:g.N, 0] + (hbar - hbar_old)[:g.N] / dt).max() <= 1e-12 * np.abs(hbar - hbar_old).max() / dt * 10


@pytest.mark.parametrize("ver,pv", [("QR4C", 1.0), ("QR4C", 0.0), ("PPM", 1.0)])
def test_vertical_schemes_on_stretched_layers(ver, pv):
    """the same on layers that grow by 35 % each: QR4C's slopes over the layer-centre distances Z(k-1) - Z(k) and its
    extrapolation to zbar(k), and PPM's non-uniform interface weights, are exact for a profile linear in depth"""
    g, st, trs, nb, dt = linear_case(hor="UPW1", ver=ver, pv=pv, vertical=True, uniform_z=False)
    h = st.hnode.numpy()
    assert h[0, 5] > 3.0 * h[0, 1]
    exact = -dt * h * W0 * GC
    lev = np.arange(1, g.L + 1)
    ok = (lev >= 4) & (lev <= g.L - 4)
    for dh, dv in _both(g, st, trs, nb, dt):
        err = (np.abs(dv - exact)[:, ok] / np.abs(exact)[:, ok]).max()
        assert err <= 1e-9, (ver, pv, err)
