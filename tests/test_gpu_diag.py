"""-m gpu: the optional diagnostics of do_oce_adv_tra -- tracers%data(tr_num)%ltra_diag (the reference's default,
src/MOD_TRACER.F90:25): tra_advhoriz / tra_advvert (src/oce_adv_tra_driver.F90:221-229, :307-318, :464-488) -- through
the C ABI against the C restatement, bit for bit on the wet layers of the owned nodes; everything else untouched.
And ldiag_DVD: dvd_trflx_hor / dvd_trflx_ver (:263-296, :395-458), every entry."""
import numpy as np
import pytest
import torch

from common import make_case, to_device
from fesom2_b200 import mesh as M

pytestmark = pytest.mark.gpu
FILL = 7.0


def _oracle(g, st, trs, nb, dt, init_h, init_v):
    from oracle import oracle_py as O
    rk = O.OracleRank(g, st, trs, nb, tra_diag=True)
    for k in range(len(trs)):
        rk.dttf_h[k][...] = init_h[k]
        rk.dttf_v[k][...] = init_v[k]
    O.run([rk], dt)
    return rk


def _wet(g):
    lev = np.arange(1, g.L + 1)[None, :]
    w = (lev >= np.asarray(g.ulevels_nod2D)[:, None]) & (lev <= np.asarray(g.nlevels_nod2D)[:, None] - 1)
    w[g.N:] = False
    return w


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True),
                                                ("MFCT", "QR4C", "NON", False), ("UPW1", "CDIFF", "FCT", False),
                                                ("MUSCL", "UPW1", "NON", False)])
@pytest.mark.parametrize("host", [False, True])
def test_ltra_diag_matches_the_oracle(small_mesh, hor, ver, lim, wsplit, host):
    """three tracers (one chunk of two + one single), tendencies that are non-zero on entry (the diagnostics use the
    accumulated arrays), the middle tracer with ltra_diag = .false."""
    from fesom2_b200.driver import AdvB200
    g = small_mesh
    st, trs, nb, dt = make_case(g, 3, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    rng = np.random.default_rng(5)
    init_h = [1e-3 * rng.standard_normal((g.Nh, g.L)) for _ in trs]
    init_v = [1e-3 * rng.standard_normal((g.Nh, g.L)) for _ in trs]
    ora = _oracle(g, st, trs, nb, dt, init_h, init_v)
    dev = "cpu" if host else torch.device("cuda:0")
    ctx = AdvB200(g, nb, device=0, max_tracers=3)
    st_d, trs_d = (st, trs) if host else to_device(st, trs, dev)
    t = lambda a: torch.as_tensor(a.copy(), dtype=torch.float64, device=dev)   # noqa: E731
    dh, dv = [t(a) for a in init_h], [t(a) for a in init_v]
    tah = [torch.full((g.Nh, g.L), FILL, dtype=torch.float64, device=dev) for _ in trs]
    tav = [torch.full((g.Nh, g.L), FILL, dtype=torch.float64, device=dev) for _ in trs]
    tah[1] = None
    tav[1] = None
    ctx.set_state(st_d)
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv, tra_advhoriz=tah, tra_advvert=tav)
    wet = _wet(g)
    for k in (0, 2):
        assert np.array_equal(dh[k].cpu().numpy(), ora.dttf_h[k]) and np.array_equal(dv[k].cpu().numpy(), ora.dttf_v[k])
        for got, ref, name in ((tah[k], ora.tra_advhoriz[k], "tra_advhoriz"), (tav[k], ora.tra_advvert[k], "tra_advvert")):
            got = got.cpu().numpy()
            assert np.isfinite(got).all()
            assert np.array_equal(got[wet], ref[wet]), (k, name, np.abs(got[wet] - ref[wet]).max())
            assert (got[~wet] == FILL).all(), (k, name, "written outside the wet layers of the owned nodes")
            assert np.abs(ref[wet]).max() > 0
    ctx.close()


@pytest.mark.parametrize("lim", ["FCT", "NON"])
def test_ltra_diag_on_two_local_ranks(pi_mesh, lim):
    """owned nodes of every rank = the one-rank oracle (the halo entries, where the reference leaves partial edge sums,
    are not written)"""
    from local_ranks import run_local_ranks
    from oracle import oracle_py as O
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", lim)
    rk = O.OracleRank(g, st, trs, nb, tra_diag=True)
    O.run([rk], dt)
    res = run_local_ranks(g, g.parts[2], st, trs, dt, tra_diag=True)
    for r in res:
        own = r["owned"].astype(np.int64) - 1
        n = r["N"]
        lev = np.arange(1, g.L + 1)[None, :]
        wet = (lev >= np.asarray(g.ulevels_nod2D)[own, None]) & (lev <= np.asarray(g.nlevels_nod2D)[own, None] - 1)
        for k in range(2):
            for got, ref in ((r["tah"][k], rk.tra_advhoriz[k]), (r["tav"][k], rk.tra_advvert[k])):
                assert np.array_equal(got[:n][wet], ref[own][wet])
                assert (got[:n][~wet] == FILL).all() and (got[n:] == FILL).all()


def _scatter_range(g):
    """(E, L) mask of the layers an edge's flux is scattered on (oce_adv_tra_driver.F90:154-156)"""
    lev = np.arange(1, g.L + 1)[None, :]
    el1 = g.edge_tri[:, 0].astype(np.int64) - 1
    has2 = g.edge_tri[:, 1] > 0
    el2 = np.where(has2, g.edge_tri[:, 1].astype(np.int64) - 1, 0)
    nu1, nl1 = g.ulevels[el1][:, None], (g.nlevels[el1] - 1)[:, None]
    nu2 = np.where(has2, g.ulevels[el2], 0)[:, None]
    nl2 = np.where(has2, g.nlevels[el2] - 1, 0)[:, None]
    lo = np.where(nu2 > 0, np.minimum(nu1, nu2), nu1)
    return (lev >= lo) & (lev <= np.maximum(nl1, nl2))


@pytest.mark.parametrize("which", ["pi", "cavity"])
@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True), ("MFCT", "QR4C", "NON", False)])
@pytest.mark.parametrize("host", [False, True])
def test_ldiag_dvd_matches_the_oracle(pi_mesh, cav_mesh, which, hor, ver, lim, wsplit, host):
    """three tracers, the last one without DVD arrays (the reference does it for temperature and salinity only)"""
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200
    g = {"pi": pi_mesh, "cavity": cav_mesh}[which]
    st, trs, nb, dt = make_case(g, 3, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    rk = O.OracleRank(g, st, trs, nb, dvd=True)
    O.run([rk], dt)
    dev = "cpu" if host else torch.device("cuda:0")
    ctx = AdvB200(g, nb, device=0, max_tracers=3)
    st_d, trs_d = (st, trs) if host else to_device(st, trs, dev)
    z = lambda shape, v=0.0: [torch.full(shape, v, dtype=torch.float64, device=dev) for _ in trs]   # noqa: E731
    dh, dv = z((g.Nh, g.L)), z((g.Nh, g.L))
    fh, fv = z((g.E, g.L), FILL), z((g.N, g.nl), FILL)
    fh[2] = None
    fv[2] = None
    ctx.set_state(st_d)
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv, dvd_trflx_hor=fh, dvd_trflx_ver=fv)
    scat = _scatter_range(g)
    for k in range(2):
        assert np.array_equal(dh[k].cpu().numpy(), rk.dttf_h[k]) and np.array_equal(dv[k].cpu().numpy(), rk.dttf_v[k])
        gh, gv = fh[k].cpu().numpy(), fv[k].cpu().numpy()
        assert np.isfinite(gh).all() and np.isfinite(gv).all()
        assert np.array_equal(gh[scat], rk.dvd_trflx_hor[k][scat]), np.abs(gh[scat] - rk.dvd_trflx_hor[k][scat]).max()
        assert (gh[~scat] == 0.0).all()            # incl. the layers of boundary edges above a cavity top (header)
        if which == "pi":
            assert np.array_equal(gh, rk.dvd_trflx_hor[k])
        assert np.array_equal(gv, rk.dvd_trflx_ver[k]), np.abs(gv - rk.dvd_trflx_ver[k]).max()
        assert np.abs(gh).max() > 0 and np.abs(gv).max() > 0
    ctx.close()
