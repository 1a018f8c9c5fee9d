"""Pins for the CPU oracle (oracle/adv_oracle.c).  The reference (Fortran + MPI) cannot be built in
this image and ships no golden vector for the path -> PARITY UNPINNED against the
reference itself.  What pins the restatement instead:

 (v)   an independently written, vectorised NumPy restatement (oracle/numpy_ref.py) agrees with the
       C oracle on the reference's own pi / soufflet meshes for every scheme it covers;
 (ii)  a constant tracer stays constant when w satisfies continuity and hnode_new = hnode;
 (iii) global conservation: sum(areasvol * (dttf_h + dttf_v)) = 0 on a closed basin;
 (iv)  FCT monotonicity: the updated value stays within the cluster bounds ("no new extrema");
 (vi)  1-rank vs N-rank identity (tests/test_multirank_cpu.py).
"""
import itertools

import numpy as np
import pytest
import torch

from common import make_case, rel_err, run_oracle
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M
from oracle.numpy_ref import NumpyAdv

TOL = 1e-12


def _numpy_run(mesh, st, trs, nb, dt):
    ref = NumpyAdv(mesh, st, nb)
    out = []
    for t in trs:
        dh, dv = np.zeros((mesh.Nh, mesh.L)), np.zeros((mesh.Nh, mesh.L))
        ref.do_oce_adv_tra(dt, t, dh, dv)
        out.append((dh, dv, dict(getattr(ref, "keep", {}))))
    return out


@pytest.mark.parametrize("hor,ver,lim", [("UPW1", "UPW1", "NON"), ("MFCT", "QR4C", "FCT"), ("MUSCL", "QR4C", "FCT"),
                                         ("MFCT", "CDIFF", "NON"), ("MUSCL", "UPW1", "FCT"), ("UPW1", "QR4C", "FCT")])
def test_numpy_restatement_agrees_on_pi(pi_mesh, hor, ver, lim):
    st, trs, nb, dt = make_case(pi_mesh, 2, hor, ver, lim, ph=0.25, pv=0.75)
    ora = run_oracle(pi_mesh, st, trs, nb, dt)
    res = _numpy_run(pi_mesh, st, trs, nb, dt)
    N = pi_mesh.N
    for k, (dh, dv, keep) in enumerate(res):
        assert rel_err(dh[:N], ora.dttf_h[k][:N]) <= TOL, ("dttf_h", k)
        assert rel_err(dv[:N], ora.dttf_v[k][:N]) <= TOL, ("dttf_v", k)
    if lim == "FCT":       # the oracle's shared work arrays hold the last tracer
        keep = res[-1][2]
        for name in ("fct_LO", "fct_plus", "fct_minus"):
            assert rel_err(keep[name][:N], ora.keep[name][:N]) <= TOL, name


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "PPM", "FCT", False), ("MUSCL", "PPM", "NON", False),
                                                ("MFCT", "QR4C", "FCT", True), ("MUSCL", "PPM", "FCT", True),
                                                ("UPW1", "CDIFF", "FCT", True)])
@pytest.mark.parametrize("which", ["pi", "small"])
def test_numpy_restatement_ppm_and_wsplit(pi_mesh, small_mesh, which, hor, ver, lim, wsplit):
    """PPM (src/oce_adv_tra_ver.F90:438-631), adv_tra_vert_impl (:90-240) and the use_wsplit branch of the driver
    (src/oce_adv_tra_driver.F90:320-334) in the second, vectorised restatement: every scheme now has two
    independently written CPU restatements that must agree."""
    mesh = pi_mesh if which == "pi" else small_mesh
    st, trs, nb, dt = make_case(mesh, 2, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    ora = run_oracle(mesh, st, trs, nb, dt)
    res = _numpy_run(mesh, st, trs, nb, dt)
    N = mesh.N
    for k, (dh, dv, keep) in enumerate(res):
        assert rel_err(dh[:N], ora.dttf_h[k][:N]) <= TOL, ("dttf_h", k)
        assert rel_err(dv[:N], ora.dttf_v[k][:N]) <= TOL, ("dttf_v", k)
    if lim == "FCT":
        keep = res[-1][2]
        for name in ("fct_LO", "fct_plus", "fct_minus"):
            assert rel_err(keep[name][:N], ora.keep[name][:N]) <= TOL, name


def test_numpy_restatement_agrees_on_soufflet(souf_mesh):
    st, trs, nb, dt = make_case(souf_mesh, 2, "MFCT", "QR4C", "FCT")
    ora = run_oracle(souf_mesh, st, trs, nb, dt)
    res = _numpy_run(souf_mesh, st, trs, nb, dt)
    N = souf_mesh.N
    for k, (dh, dv, _) in enumerate(res):
        assert rel_err(dh[:N], ora.dttf_h[k][:N]) <= TOL
        assert rel_err(dv[:N], ora.dttf_v[k][:N]) <= TOL


@pytest.mark.parametrize("hor,ver,lim", list(itertools.product(("UPW1", "MUSCL", "MFCT"), ("UPW1", "QR4C", "PPM", "CDIFF"), ("FCT", "NON"))))
def test_constant_tracer_stays_constant(small_mesh, hor, ver, lim):
    """(ii): w comes from the continuity restatement (vert_vel_ale, src/oce_ale.F90:2164-2310) and
    hnode_new = hnode, so advecting T = const must give a zero tendency to round-off."""
    g = small_mesh
    st, trs, nb, dt = make_case(g, 1, hor, ver, lim, ph=0.25, pv=0.75)
    c0 = 7.25
    _, nmask = F.layer_masks(g, "cpu")
    trs[0].values = torch.where(nmask, torch.full_like(trs[0].values, c0), torch.zeros_like(trs[0].values))
    trs[0].valuesAB = trs[0].values.clone()
    trs[0].edge_up_dn_grad = torch.zeros_like(trs[0].edge_up_dn_grad)
    ora = run_oracle(g, st, trs, nb, dt)
    N = g.N
    hn = st.hnode_new.numpy()[:N]
    mask = nmask.numpy()[:N]
    dval = np.where(mask, (ora.dttf_h[0][:N] + ora.dttf_v[0][:N]) / np.where(mask, hn, 1.0), 0.0)
    assert np.abs(dval).max() <= 1e-11 * c0


@pytest.mark.parametrize("lim", ["NON", "FCT"])
def test_global_conservation(small_mesh, lim):
    """(iii): horizontal fluxes cancel pairwise and vertical fluxes telescope, so the area-weighted sum
    of the tendencies equals the flux through the surface interface (linfs: w(surface) /= 0):
    dt * sum_n [LO surface flux (FCT only) + (limited) high-order surface flux]"""
    g = small_mesh
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", lim)
    ora = run_oracle(g, st, trs, nb, dt)
    _, nmask = F.layer_masks(g, "cpu")
    N, L = g.N, g.L
    mask = nmask.numpy()[:N]
    av = g.areasvol[:N, :L]
    k = 1                                   # the oracle's work arrays hold the last tracer
    tend = np.where(mask, (ora.dttf_h[k][:N] + ora.dttf_v[k][:N]) * av, 0.0)
    top = g.ulevels_nod2D[:N] - 1
    rows = np.arange(N)
    surf = ora.keep["adv_flux_ver"][rows, top].copy()                 # (limited) antidiffusive / HO flux at nzmin
    if lim == "FCT":                                                  # + LO flux -we*ttf*area (oce_adv_tra_ver.F90:301)
        surf += -st.w_e.numpy()[rows, top] * trs[k].values.numpy()[rows, top] * g.area[rows, top]
    scale = np.abs(tend).sum()
    assert abs(tend.sum() - dt * surf.sum()) <= 1e-10 * scale


@pytest.mark.parametrize("which,hor,ver", [("pi", "MFCT", "QR4C"), ("cavity", "MFCT", "QR4C"), ("nw2", "MUSCL", "PPM"), ("souf", "MFCT", "CDIFF")])
def test_fct_no_new_extrema(pi_mesh, cav_mesh, nw2_mesh, souf_mesh, which, hor, ver):
    """(iv) / Appendix C: values_new within the 3-D cluster bounds of oce_adv_tra_fct.F90:124-248
    recomputed here from the oracle's ttf and fct_LO -- on all four meshes of the reference"""
    g = {"pi": pi_mesh, "cavity": cav_mesh, "nw2": nw2_mesh, "souf": souf_mesh}[which]
    st, trs, nb, dt = make_case(g, 1, hor, ver, "FCT")
    ora = run_oracle(g, st, trs, nb, dt)
    N, L = g.N, g.L
    ttf = trs[0].values.numpy()
    lo = ora.keep["fct_LO"]
    _, nmask = F.layer_masks(g, "cpu")
    mask = nmask.numpy()
    hi_n = np.where(mask, np.maximum(ttf, lo), -np.inf)
    lo_n = np.where(mask, np.minimum(ttf, lo), np.inf)
    # cluster = node + edge neighbours (all nodes of its elements), then the level above/below
    n1, n2 = g.edges[:, 0] - 1, g.edges[:, 1] - 1
    cmax, cmin = hi_n.copy(), lo_n.copy()
    np.maximum.at(cmax, n1, hi_n[n2]); np.maximum.at(cmax, n2, hi_n[n1])
    np.minimum.at(cmin, n1, lo_n[n2]); np.minimum.at(cmin, n2, lo_n[n1])
    pad = lambda a, f: np.concatenate([np.full((a.shape[0], 1), f), a, np.full((a.shape[0], 1), f)], 1)
    vmax = np.maximum.reduce([pad(cmax, -np.inf)[:, :-2], cmax, pad(cmax, -np.inf)[:, 2:]])
    vmin = np.minimum.reduce([pad(cmin, np.inf)[:, :-2], cmin, pad(cmin, np.inf)[:, 2:]])
    new = ttf[:N] + np.where(mask[:N], (ora.dttf_h[0][:N] + ora.dttf_v[0][:N]) / np.where(mask[:N], st.hnode_new.numpy()[:N], 1.0), 0.0)
    eps = 1e-12 * np.abs(ttf).max()
    m = mask[:N]
    assert (new[m] <= vmax[:N][m] + eps).all()
    assert (new[m] >= vmin[:N][m] - eps).all()


@pytest.mark.parametrize("which", ["pi", "soufflet"])
def test_gradient_producer_restatements_agree(which, pi_mesh, souf_mesh):
    """tracer_gradient_elements / fill_up_dn_grad: the C restatement, the NumPy restatement (bit for bit)
    and the torch generator of the synthetic inputs (to round-off: it sums with index_add)."""
    import torch
    from fesom2_b200 import fields as F
    from oracle import numpy_ref as R, oracle_py as O
    g = {"pi": pi_mesh, "soufflet": souf_mesh}[which]
    v, _ = F.make_tracer_values(g, "cpu", kind=0)
    tri = F.find_up_downwind_triangles(g)
    a, b = O.tracer_gradient_elements(g, v.numpy()), R.tracer_gradient_elements(g, v.numpy())
    assert np.array_equal(a[:g.T], b)
    ga, gb = O.fill_up_dn_grad(g, a, tri), R.fill_up_dn_grad(g, b, tri)
    assert np.isfinite(ga).all() and np.array_equal(ga, gb)
    gt = F.fill_up_dn_grad(g, F.tracer_gradient_elements(g, v, "cpu"), tri, "cpu").numpy()
    assert np.abs(ga - gt).max() <= 1e-13 * np.abs(ga).max()


def test_vert_vel_ale_core_restatements_agree(pi_mesh):
    """SURVEY 8f row 3, oracle first: the continuity part of vert_vel_ale -- C restatement vs NumPy restatement
    bit for bit, and vs the torch generator of the synthetic w (which scatters with index_add) to round-off"""
    from oracle import numpy_ref as R, oracle_py as O
    g = pi_mesh
    st = F.make_state(g, "cpu")
    trs = F.make_tracers(g, 1, "cpu")
    rk = O.OracleRank(g, st, trs, M.nboundary_lay(g))
    a = O.vert_vel_ale_core(rk)
    b = R.vert_vel_ale_core(g, st.uv.numpy(), st.helem.numpy())
    assert np.isfinite(a).all() and np.array_equal(a, b)
    w = st.w.numpy()
    assert np.abs(a - w).max() <= 1e-12 * np.abs(w).max()


@pytest.mark.parametrize("which", ["pi", "souf", "small"])
def test_find_up_downwind_triangles_two_restatements(pi_mesh, souf_mesh, small_mesh, which):
    """find_up_downwind_triangles (src/oce_muscl_adv.F90:162-352): the whole-array NumPy version that feeds the
    harness (fesom2_b200/fields.py) against the loop-for-loop C restatement (oracle/adv_oracle.c, libm atan2 as
    in a gfortran build).  On the reference's unstructured pi mesh they agree exactly.  On structured meshes the
    edge direction can lie EXACTLY on an element edge around the far node; the reference's test `ab == ax` then
    depends on the last bit of atan2, and the two restatements may pick different -- adjacent -- triangles: every
    disagreement must be such a tie (both candidates contain the end node and share the edge the direction lies on)."""
    from oracle import oracle_py as O
    mesh = {"pi": pi_mesh, "souf": souf_mesh, "small": small_mesh}[which]
    a = F.find_up_downwind_triangles(mesh)
    b = O.find_up_downwind_triangles(mesh)
    assert a.shape == b.shape == (mesh.E, 2)
    bad = np.argwhere(a != b)
    if which == "pi":
        assert bad.size == 0
    en = mesh.elem2D_nodes.astype(np.int64)
    for e, side in bad:
        ea, eb = int(a[e, side]), int(b[e, side])
        assert ea > 0 and eb > 0, (e, side, ea, eb)
        node = int(mesh.edges[e, side])
        na, nb_ = set(en[ea - 1].tolist()), set(en[eb - 1].tolist())
        assert node in na and node in nb_
        assert len(na & nb_) == 2, "candidates are not adjacent triangles"
    assert bad.shape[0] <= 0.02 * mesh.E


@pytest.mark.parametrize("use_wsplit", [False, True])
def test_cflz_and_wvel_split_two_restatements(pi_mesh, use_wsplit):
    """compute_CFLz + compute_Wvel_split (src/oce_ale.F90:2906-3049): loop-for-loop C restatement against the
    whole-array NumPy one, on a w large enough that part of it goes implicit."""
    from oracle import numpy_ref as R, oracle_py as O
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 1)
    rk = O.OracleRank(g, st, trs, nb)
    W = O.vert_vel_ale_core(rk)
    dtc = 40.0 * dt                                  # CFL_z well above the threshold somewhere
    a = O.compute_cflz_and_split(rk, dtc, W, use_wsplit, 0.5)
    b = R.compute_cflz_and_split(g, st.hnode_new.numpy(), dtc, W, use_wsplit, 0.5)
    for x, y in zip(a, b):
        assert np.isfinite(x).all() and np.array_equal(x, y)
    if use_wsplit:
        assert (a[2] != 0.0).any(), "nothing went implicit: the test does not exercise the split"
        assert np.abs(a[1] + a[2] - np.where(a[1] != 0, W, 0.0)).max() <= 1e-15 * np.abs(W).max()


def test_vert_vel_ale_zstar_two_restatements(pi_mesh):
    """the zstar free-surface correction of vert_vel_ale (src/oce_ale.F90:2539-2603): C loop vs whole-array NumPy"""
    from oracle import numpy_ref as R, oracle_py as O
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 1)
    rk = O.OracleRank(g, st, trs, nb)
    W = O.vert_vel_ale_core(rk)
    ids = np.arange(g.Nh, dtype=np.float64)
    hbar_old = 0.05 * np.sin(0.37 * ids)
    hbar = hbar_old + 0.01 * np.cos(0.11 * ids)
    wflux = 1.0e-6 * np.sin(0.05 * ids)
    a = O.vert_vel_ale_zstar(rk, 1800.0, W, hbar, hbar_old, wflux)
    b = R.vert_vel_ale_zstar(g, st.zbar_3d_n.numpy(), st.hnode.numpy(), st.hnode_new.numpy(), 1800.0, W, hbar, hbar_old, wflux)
    for x, y in zip(a, b):
        assert np.isfinite(x).all() and np.array_equal(x, y)
    assert not np.array_equal(a[0], W) and not np.array_equal(a[1], st.hnode_new.numpy())


def test_vert_vel_ale_zlevel_two_restatements(pi_mesh):
    """the zlevel free-surface correction of vert_vel_ale (src/oce_ale.F90:2336-2538: plain zlevel, local zstar over the
    first lzstar_lev layers, refill on the way back): C loop vs masked whole-array NumPy, every branch exercised"""
    from common import zlevel_case
    from oracle import numpy_ref as R, oracle_py as O
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 1)
    for lz, mh in ((4, 0.5), (3, 0.7)):
        hbar, hbar_old, wflux, cfl_old = zlevel_case(g, st, lz, mh)
        rk = O.OracleRank(g, st, trs, nb)
        W = O.vert_vel_ale_core(rk)
        a = O.vert_vel_ale_zlevel(rk, 1800.0, W, hbar, hbar_old, wflux, g.zbar, cfl_old, mh, lz)
        b = R.vert_vel_ale_zlevel(g, g.zbar, st.hnode.numpy(), st.hnode_new.numpy(), cfl_old, 1800.0, W, hbar, hbar_old, wflux, mh, lz)
        for x, y in zip(a, b):
            assert np.isfinite(x).all() and np.array_equal(x, y)
        # every branch is taken: layers below the surface change in some columns (local zstar / refill), only the
        # surface layer in others, and the elevation change is conserved by the distribution wherever it fits
        ch = a[1][:g.N] != st.hnode_new.numpy()[:g.N]
        deep = ch[:, 1:lz].any(axis=1)
        assert deep.any() and (ch[:, 0] & ~deep).any() and not ch[:, lz:].any()
        top = np.asarray(g.ulevels_nod2D)[:g.N] == 1
        dsum = (a[1][:g.N, :lz] - st.hnode.numpy()[:g.N, :lz]).sum(axis=1)
        ok = top & (np.abs(dsum - (hbar - hbar_old)[:g.N]) < 1e-10)
        assert ok.sum() > 0.5 * top.sum()


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True),
                                                ("MFCT", "QR4C", "NON", False), ("UPW1", "UPW1", "FCT", False)])
def test_diagnostics_two_restatements(pi_mesh, hor, ver, lim, wsplit):
    """ltra_diag (tra_advhoriz / tra_advvert, src/oce_adv_tra_driver.F90:221-229, :307-318, :464-488) and ldiag_DVD
    (dvd_trflx_hor / dvd_trflx_ver, :263-296, :395-458): the C loops and the whole-array NumPy restatement agree bit
    for bit, and switching the diagnostics on does not change the tendencies"""
    from oracle import numpy_ref as R, oracle_py as O
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 2, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    a = O.OracleRank(g, st, trs, nb)
    O.run([a], dt)
    b = O.OracleRank(g, st, trs, nb, tra_diag=True, dvd=True)
    O.run([b], dt)
    na = R.NumpyAdv(g, st, nb)
    for k in range(2):
        assert np.array_equal(a.dttf_h[k], b.dttf_h[k]) and np.array_equal(a.dttf_v[k], b.dttf_v[k])
        d = {}
        na.do_oce_adv_tra(dt, trs[k], np.zeros((g.Nh, g.L)), np.zeros((g.Nh, g.L)), diag=d)
        for name in ("tra_advhoriz", "tra_advvert", "dvd_trflx_hor", "dvd_trflx_ver"):
            x = getattr(b, name)[k]
            assert np.isfinite(x).all() and np.abs(x).max() > 0 and np.array_equal(x, d[name]), (name, k)


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True),
                                                ("MFCT", "QR4C", "NON", False), ("UPW1", "UPW1", "FCT", False),
                                                ("MUSCL", "CDIFF", "FCT", False)])
def test_cavity_mesh_two_restatements(cav_mesh, hor, ver, lim, wsplit):
    """the reference's cavity mesh (test/meshes/pi_cavity, use_cavity): nzmin > 1 in the vertical stencils, one-sided edge
    ranges A / B, padded FCT clusters, areasvol = lower face under the ice -- C loops vs whole-array NumPy, bit for bit,
    tendencies and all four diagnostics"""
    from oracle import numpy_ref as R, oracle_py as O
    g = cav_mesh
    assert (g.ulevels > 1).sum() == 170 and (g.ulevels_nod2D > 1).sum() == 95
    under = (np.arange(1, g.nl + 1)[None, :] < np.asarray(g.ulevels_nod2D_max)[:, None]) & (g.areasvol > 0)
    assert under.any() and (g.areasvol[under] != g.area[under]).any()      # the lower-face rule is in effect
    st, trs, nb, dt = make_case(g, 2, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    b = O.OracleRank(g, st, trs, nb, tra_diag=True, dvd=True)
    O.run([b], dt)
    na = R.NumpyAdv(g, st, nb)
    cav = np.asarray(g.ulevels_nod2D) > 1
    for k in range(2):
        d = {}
        dh, dv = np.zeros((g.Nh, g.L)), np.zeros((g.Nh, g.L))
        with np.errstate(invalid="ignore"):
            na.do_oce_adv_tra(dt, trs[k], dh, dv, diag=d)
        assert np.isfinite(dh).all() and np.isfinite(dv).all() and np.abs(dh[cav]).max() > 0
        assert np.array_equal(dh, b.dttf_h[k]) and np.array_equal(dv, b.dttf_v[k])
        for name in ("tra_advhoriz", "tra_advvert", "dvd_trflx_hor", "dvd_trflx_ver"):
            assert np.array_equal(getattr(b, name)[k], d[name]), (name, k)


def test_cavity_mesh_1rank_vs_2ranks(cav_mesh):
    """the reference's dist_2 partition of pi_cavity: owned nodes identical on 1 and 2 ranks (three dwarf iterations)"""
    from oracle import oracle_py as O
    g = cav_mesh
    st, trs, nbg, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    one = O.OracleRank(g, st, trs, nbg)
    O.run([one], dt, 3, 1)
    ranks = []
    for r in range(2):
        loc = M.localize(g, g.parts[2], r)
        lst, ltr = F.scatter_to_local(g, loc, st, trs)
        ranks.append(O.OracleRank(loc, lst, ltr, nbg[loc.myList_nod2D - 1]))
    O.run(ranks, dt, 3, 1)
    for rk in ranks:
        loc = rk.mesh_py
        own, alln = loc.myList_nod2D[:loc.N] - 1, loc.myList_nod2D - 1
        for k in range(2):
            assert np.array_equal(rk.dttf_v[k][:loc.N], one.dttf_v[k][own])
            assert np.array_equal(rk.dttf_h[k][:loc.N], one.dttf_h[k][own])
            assert np.array_equal(rk.values[k], one.values[k][alln])


def test_cavity_mesh_side_rows_two_restatements(cav_mesh):
    """rows f-1 and f-3 on the cavity mesh: tracer_gradient_elements, fill_up_dn_grad, the continuity part of
    vert_vel_ale, compute_CFLz / compute_Wvel_split and the zstar / zlevel corrections (which skip the cavity columns,
    src/oce_ale.F90:2354, :2550) -- C loops vs whole-array NumPy, bit for bit"""
    from common import zlevel_case
    from oracle import numpy_ref as R, oracle_py as O
    g = cav_mesh
    st, trs, nb, dt = make_case(g, 1)
    v = trs[0].values.numpy()
    tri = F.find_up_downwind_triangles(g)
    assert np.array_equal(tri, O.find_up_downwind_triangles(g))
    a, b = O.tracer_gradient_elements(g, v), R.tracer_gradient_elements(g, v)
    assert np.array_equal(a[:g.T], b)
    ga, gb = O.fill_up_dn_grad(g, a, tri), R.fill_up_dn_grad(g, b, tri)
    # under the ice the reference's last two loops (src/oce_muscl_adv.F90:445-485) start at nzmax, above the node's own top:
    # 0/0 there, a NaN nobody reads -- both restatements keep it
    assert np.isnan(ga).sum() == 4 and np.array_equal(np.isnan(ga), np.isnan(gb)) and np.array_equal(np.nan_to_num(ga), np.nan_to_num(gb))
    hbar, hbar_old, wflux, cfl_old = zlevel_case(g, st, 4, 0.5)
    rk = O.OracleRank(g, st, trs, nb)
    W = O.vert_vel_ale_core(rk)
    assert np.isfinite(W).all() and np.array_equal(W, R.vert_vel_ale_core(g, st.uv.numpy(), st.helem.numpy()))
    dtc = 40.0 * dt
    for x, y in zip(O.compute_cflz_and_split(rk, dtc, W, True, 0.5), R.compute_cflz_and_split(g, st.hnode_new.numpy(), dtc, W, True, 0.5)):
        assert np.isfinite(x).all() and np.array_equal(x, y)
    za = O.vert_vel_ale_zstar(rk, dtc, W, hbar, hbar_old, wflux)
    zb = R.vert_vel_ale_zstar(g, st.zbar_3d_n.numpy(), st.hnode.numpy(), st.hnode_new.numpy(), dtc, W, hbar, hbar_old, wflux)
    la = O.vert_vel_ale_zlevel(rk, dtc, W, hbar, hbar_old, wflux, g.zbar, cfl_old, 0.5, 4)
    lb = R.vert_vel_ale_zlevel(g, g.zbar, st.hnode.numpy(), st.hnode_new.numpy(), cfl_old, dtc, W, hbar, hbar_old, wflux, 0.5, 4)
    cav = np.asarray(g.ulevels_nod2D)[:g.N] > 1
    for (x, y) in (za, zb), (la, lb):
        for p, q in zip(x, y):
            assert np.isfinite(p).all() and np.array_equal(p, q)
        assert np.array_equal(x[0][:g.N][cav], W[:g.N][cav]) and np.array_equal(x[1][:g.N][cav], st.hnode_new.numpy()[:g.N][cav])
        assert not np.array_equal(x[0], W)


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True), ("MFCT", "QR4C", "NON", False)])
def test_neverworld2_mesh_two_restatements(nw2_mesh, hor, ver, lim, wsplit):
    """the reference's test/meshes/neverworld2 (periodic 60-degree sector, columns of 4 to 15 layers)"""
    from oracle import numpy_ref as R, oracle_py as O
    g = nw2_mesh
    st, trs, nb, dt = make_case(g, 2, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    b = O.OracleRank(g, st, trs, nb, tra_diag=True)
    O.run([b], dt)
    na = R.NumpyAdv(g, st, nb)
    for k in range(2):
        d = {}
        dh, dv = np.zeros((g.Nh, g.L)), np.zeros((g.Nh, g.L))
        with np.errstate(invalid="ignore"):
            na.do_oce_adv_tra(dt, trs[k], dh, dv, diag=d)
        assert np.isfinite(dh).all() and np.abs(dh).max() > 0
        assert np.array_equal(dh, b.dttf_h[k]) and np.array_equal(dv, b.dttf_v[k])
        assert np.array_equal(b.tra_advhoriz[k], d["tra_advhoriz"]) and np.array_equal(b.tra_advvert[k], d["tra_advvert"])


@pytest.mark.parametrize("which", ["cavity", "nw2"])
@pytest.mark.parametrize("hor,ver,lim", [("MFCT", "QR4C", "FCT"), ("MUSCL", "PPM", "NON"), ("UPW1", "CDIFF", "FCT")])
def test_invariants_on_the_cavity_and_neverworld2_meshes(cav_mesh, nw2_mesh, which, hor, ver, lim):
    """(ii) and (iii) of SURVEY 8c on the reference's other two meshes: a constant tracer keeps a zero tendency under the
    ice (nzmin > 1, areasvol = lower face) and in 4-layer columns, and the area-weighted sum of the tendencies equals the
    flux through the top interface of every column (which is the cavity base where there is one)"""
    g = {"cavity": cav_mesh, "nw2": nw2_mesh}[which]
    _, nmask = F.layer_masks(g, "cpu")
    N, L = g.N, g.L
    mask = nmask.numpy()[:N]
    st, trs, nb, dt = make_case(g, 1, hor, ver, lim, ph=0.25, pv=0.75)
    real = run_oracle(g, st, trs, nb, dt)                                    # a real field first: conservation
    tend = np.where(mask, (real.dttf_h[0][:N] + real.dttf_v[0][:N]) * g.areasvol[:N, :L], 0.0)
    top, rows = g.ulevels_nod2D[:N] - 1, np.arange(N)
    surf = real.keep["adv_flux_ver"][rows, top].copy()
    if lim == "FCT":
        surf += -st.w_e.numpy()[rows, top] * trs[0].values.numpy()[rows, top] * g.area[rows, top]
    assert abs(tend.sum() - dt * surf.sum()) <= 1e-10 * np.abs(tend).sum()
    c0 = 7.25                                                                # then T = const
    trs[0].values = torch.where(nmask, torch.full_like(trs[0].values, c0), torch.zeros_like(trs[0].values))
    trs[0].valuesAB = trs[0].values.clone()
    trs[0].edge_up_dn_grad = torch.zeros_like(trs[0].edge_up_dn_grad)
    ora = run_oracle(g, st, trs, nb, dt)
    hn = st.hnode_new.numpy()[:N]
    dval = np.where(mask, (ora.dttf_h[0][:N] + ora.dttf_v[0][:N]) / np.where(mask, hn, 1.0), 0.0)
    assert np.abs(dval).max() <= 1e-11 * c0


@pytest.mark.parametrize("hor,ver", [("MFCT", "QR4C"), ("MUSCL", "PPM")])
def test_constant_tracer_with_wsplit(small_mesh, hor, ver):
    """(ii) with dynamics%use_wsplit: the explicit upwind flux on w_e plus the implicit sweep adv_tra_vert_impl on w_i
    (src/oce_adv_tra_ver.F90:90-240, src/oce_adv_tra_driver.F90:320-334) together still advect nothing when T = const --
    which pins the tridiagonal coefficients of the implicit part against the explicit one"""
    g = small_mesh
    st, trs, nb, dt = make_case(g, 1, hor, ver, "FCT", ph=0.25, pv=0.75, use_wsplit=True)
    assert (st.w_i.numpy() != 0).sum() > 10
    c0 = 7.25
    _, nmask = F.layer_masks(g, "cpu")
    trs[0].values = torch.where(nmask, torch.full_like(trs[0].values, c0), torch.zeros_like(trs[0].values))
    trs[0].valuesAB = trs[0].values.clone()
    trs[0].edge_up_dn_grad = torch.zeros_like(trs[0].edge_up_dn_grad)
    ora = run_oracle(g, st, trs, nb, dt)
    N = g.N
    mask = nmask.numpy()[:N]
    dval = np.where(mask, (ora.dttf_h[0][:N] + ora.dttf_v[0][:N]) / np.where(mask, st.hnode_new.numpy()[:N], 1.0), 0.0)
    assert np.abs(dval).max() <= 1e-11 * c0
