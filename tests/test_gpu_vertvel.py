"""-m gpu: SURVEY.md section 8f row 3 on the device -- vert_vel_ale (continuity part, linfs), compute_CFLz,
compute_Wvel_split (src/oce_ale.F90:2164-2310, :2906-3049) and the coalesced adv_tra_vert_impl -- bit for bit
against the C restatement, on one rank and on in-process ranks (exchange_nod(Wvel))."""
import numpy as np
import pytest
import torch

from common import make_case, run_oracle, to_device
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

pytestmark = pytest.mark.gpu


def _oracle_w(g, st, trs, nb, dt, use_wsplit, maxcfl):
    from oracle import oracle_py as O
    rk = O.OracleRank(g, st, trs, nb)
    W = O.vert_vel_ale_core(rk)
    cfl, we, wi = O.compute_cflz_and_split(rk, dt, W, use_wsplit, maxcfl)
    return W, we, wi, cfl


@pytest.mark.parametrize("which", ["pi", "soufflet", "small"])
@pytest.mark.parametrize("use_wsplit", [False, True])
def test_vert_vel_ale_matches_the_oracle(which, use_wsplit, pi_mesh, souf_mesh, small_mesh):
    from fesom2_b200.driver import AdvB200
    g = {"pi": pi_mesh, "soufflet": souf_mesh, "small": small_mesh}[which]
    st, trs, nb, dt = make_case(g, 1)
    dtc = 40.0 * dt
    W, we, wi, cfl = _oracle_w(g, st, trs, nb, dtc, use_wsplit, 0.5)
    dev = torch.device("cuda:0")
    st_d, _ = to_device(st, [], dev)
    ctx = AdvB200(g, nb, max_tracers=1)
    ctx.set_state(st_d)
    out = [torch.full((g.Nh, g.nl), -7.0, dtype=torch.float64, device=dev) for _ in range(4)]
    out[1].zero_(); out[2].zero_()                    # w_e / w_i: entries the reference leaves alone stay as they are
    ctx.vert_vel_ale(dtc, use_wsplit, 0.5, *out)
    ctx.synchronize()
    for got, ref, name in zip(out, (W, we, wi, cfl), ("w", "w_e", "w_i", "cfl_z")):
        assert np.array_equal(got.cpu().numpy(), ref), name
    if use_wsplit:
        assert (wi != 0).any()
    ctx.close()


def test_wsplit_step_entirely_on_the_device(pi_mesh):
    """test_pi runs with use_wsplit (setups/test_pi/setup.yml:23): uv -> w, w_e, w_i on the device -> FCT step with
    adv_tra_vert_impl, against the oracle chain (vert_vel_ale_core + split + do_oce_adv_tra)"""
    from fesom2_b200.driver import AdvB200
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    dtc = 30.0 * dt
    W, we, wi, _ = _oracle_w(g, st, trs, nb, dtc, True, 0.5)
    st.w, st.w_e, st.w_i, st.use_wsplit = torch.as_tensor(W), torch.as_tensor(we), torch.as_tensor(wi), True
    assert (wi != 0).any()
    ora = run_oracle(g, st, trs, nb, dtc)
    dev = torch.device("cuda:0")
    st_d, trs_d = to_device(st, trs, dev)
    for x in (st_d.w, st_d.w_e, st_d.w_i):
        x.fill_(float("nan"))                          # must all come from the device computation
    ctx = AdvB200(g, nb, max_tracers=2)
    ctx.set_state(st_d)
    st_d.w_e.zero_(); st_d.w_i.zero_()
    ctx.vert_vel_ale(dtc, True, 0.5, st_d.w, st_d.w_e, st_d.w_i)
    ctx.set_state(st_d)
    dh = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx.do_oce_adv_tra(dtc, trs_d, dh, dv)
    for k in range(2):
        assert np.array_equal(dh[k].cpu().numpy(), ora.dttf_h[k])
        assert np.array_equal(dv[k].cpu().numpy(), ora.dttf_v[k])
    ctx.close()


@pytest.mark.parametrize("world", [2, 8])
def test_local_ranks_wsplit_and_vert_vel(world, pi_mesh):
    """N in-process ranks: vert_vel_ale with its exchange_nod(Wvel), then the use_wsplit FCT step with the
    boundary-first / interior overlap kept on (adv_tra_vert_impl runs per node range)"""
    import threading
    from fesom2_b200.driver import AdvB200, comm_init_local
    g = pi_mesh
    part = g.parts[world]
    st, trs, nb_g, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    dtc = 30.0 * dt
    W, we, wi, _ = _oracle_w(g, st, trs, nb_g, dtc, True, 0.5)
    st.w, st.w_e, st.w_i, st.use_wsplit = torch.as_tensor(W), torch.as_tensor(we), torch.as_tensor(wi), True
    # 1-rank reference: W of halo-exchanged columns is complete everywhere
    ora = run_oracle(g, st, trs, nb_g, dtc)
    ndev = torch.cuda.device_count()
    ranks = []
    for r in range(world):
        loc = M.localize(g, part, r)
        lst, ltr = F.scatter_to_local(g, loc, st, trs)
        dev = torch.device(f"cuda:{r % ndev}")
        st_d, trs_d = to_device(lst, ltr, dev)
        for x in (st_d.w, st_d.w_e, st_d.w_i):
            x.zero_()
        ctx = AdvB200(loc, nb_g[loc.myList_nod2D - 1], device=r % ndev, max_tracers=2)
        dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in trs]
        dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in trs]
        ranks.append(dict(loc=loc, ctx=ctx, st=st_d, trs=trs_d, dh=dh, dv=dv, err=None))
    comm_init_local([rk["ctx"] for rk in ranks])
    torch.cuda.synchronize()

    def work(rk):
        try:
            ctx, s = rk["ctx"], rk["st"]
            torch.cuda.set_device(ctx.device)
            ctx.set_state(s)
            ctx.vert_vel_ale(dtc, True, 0.5, s.w, s.w_e, s.w_i)
            ctx.set_state(s)
            ctx.do_oce_adv_tra(dtc, rk["trs"], rk["dh"], rk["dv"])
            ctx.synchronize()
        except Exception as ex:
            rk["err"] = ex
    th = [threading.Thread(target=work, args=(rk,)) for rk in ranks]
    [t.start() for t in th]
    [t.join() for t in th]
    for rk in ranks:
        if rk["err"] is not None:
            raise rk["err"]
        loc = rk["loc"]
        alln = loc.myList_nod2D.astype(np.int64) - 1
        assert np.array_equal(rk["st"].w.cpu().numpy(), W[alln])
        assert np.array_equal(rk["st"].w_i.cpu().numpy(), wi[alln])
        own = alln[:loc.N]
        for k in range(2):
            assert np.array_equal(rk["dh"][k][:loc.N].cpu().numpy(), ora.dttf_h[k][own])
            assert np.array_equal(rk["dv"][k][:loc.N].cpu().numpy(), ora.dttf_v[k][own])
        rk["ctx"].close()


@pytest.mark.parametrize("world", [1, 2])
def test_vert_vel_ale_zstar(world, pi_mesh):
    """which_ALE = 'zstar' (the reference's default): continuity part + elevation change distributed over the layers
    (Wvel, hnode_new) + fresh-water flux + both exchanges + CFLz/split with the NEW hnode_new, against the oracle chain"""
    import threading
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200, comm_init_local
    g = pi_mesh
    st, trs, nb_g, dt = make_case(g, 1)
    dtc = 40.0 * dt
    ids = np.arange(g.Nh, dtype=np.float64)
    hbar_old = 0.05 * np.sin(0.37 * ids)
    hbar = hbar_old + 0.01 * np.cos(0.11 * ids)
    wflux = 1.0e-6 * np.sin(0.05 * ids)
    rk = O.OracleRank(g, st, trs, nb_g)
    W0 = O.vert_vel_ale_core(rk)
    W, hn = O.vert_vel_ale_zstar(rk, dtc, W0, hbar, hbar_old, wflux)
    rk.keep["hnode_new"][...] = hn                     # compute_CFLz sees the new thickness
    cfl, we, wi = O.compute_cflz_and_split(rk, dtc, W, True, 0.5)
    ndev = torch.cuda.device_count()
    part = g.parts[world] if world > 1 else np.zeros(g.Nh, np.int32)
    ranks = []
    for r in range(world):
        loc = M.localize(g, part, r) if world > 1 else g
        lst = F.scatter_to_local(g, loc, st, trs)[0] if world > 1 else st
        alln = (loc.myList_nod2D.astype(np.int64) - 1) if world > 1 else np.arange(g.Nh)
        dev = torch.device(f"cuda:{r % ndev}")
        st_d, _ = to_device(lst, [], dev)
        ctx = AdvB200(loc, nb_g[alln], device=r % ndev, max_tracers=1)
        t = lambda a, dt_=torch.float64: torch.as_tensor(np.ascontiguousarray(a[alln]), dtype=dt_, device=dev)   # noqa: E731
        ranks.append(dict(loc=loc, ctx=ctx, st=st_d, alln=alln, err=None, hbar=t(hbar), hbar_old=t(hbar_old), wflux=t(wflux),
                          nmin=torch.as_tensor(np.ascontiguousarray(g.nlevels_nod2D_min[alln]), dtype=torch.int32, device=dev),
                          out=[torch.zeros((loc.Nh, loc.nl), dtype=torch.float64, device=dev) for _ in range(4)]))
    if world > 1:
        comm_init_local([rk_["ctx"] for rk_ in ranks])
    torch.cuda.synchronize()

    def work(q):
        try:
            ctx, s = q["ctx"], q["st"]
            torch.cuda.set_device(ctx.device)
            ctx.set_state(s)
            ctx.vert_vel_ale_zstar(dtc, True, 0.5, q["hbar"], q["hbar_old"], q["wflux"], q["nmin"], s.hnode_new, *q["out"])
            ctx.synchronize()
        except Exception as ex:
            q["err"] = ex
    th = [threading.Thread(target=work, args=(q,)) for q in ranks]
    [x.start() for x in th]
    [x.join() for x in th]
    for q in ranks:
        if q["err"] is not None:
            raise q["err"]
        a = q["alln"]
        for got, ref, name in zip(q["out"], (W, we, wi, cfl), ("w", "w_e", "w_i", "cfl_z")):
            assert np.array_equal(got.cpu().numpy(), ref[a]), (name, world)
        assert np.array_equal(q["st"].hnode_new.cpu().numpy(), hn[a])
        q["ctx"].close()
    assert (wi != 0).any() and not np.array_equal(hn, st.hnode_new.numpy())


@pytest.mark.parametrize("world", [1, 2])
def test_vert_vel_ale_zlevel(world, pi_mesh):
    """which_ALE = 'zlevel' (src/oce_ale.F90:2336-2538): continuity part + the surface-layer / local-zstar / refill
    correction of Wvel and hnode_new (all three branches occur in the inputs) + fresh-water flux + both exchanges +
    CFLz/split with the NEW hnode_new, against the oracle chain; cfl_z carries the previous step's CFL_z in"""
    import threading
    from common import zlevel_case
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200, comm_init_local
    g = pi_mesh
    st, trs, nb_g, dt = make_case(g, 1)
    dtc = 40.0 * dt
    lz, mh = 4, 0.5
    hbar, hbar_old, wflux, cfl_old = zlevel_case(g, st, lz, mh)
    rk = O.OracleRank(g, st, trs, nb_g)
    W0 = O.vert_vel_ale_core(rk)
    W, hn = O.vert_vel_ale_zlevel(rk, dtc, W0, hbar, hbar_old, wflux, g.zbar, cfl_old, mh, lz)
    assert (hn[:g.N, 1:lz] != st.hnode_new.numpy()[:g.N, 1:lz]).any()
    rk.keep["hnode_new"][...] = hn                     # compute_CFLz sees the new thickness
    cfl, we, wi = O.compute_cflz_and_split(rk, dtc, W, True, 0.5)
    ndev = torch.cuda.device_count()
    part = g.parts[world] if world > 1 else np.zeros(g.Nh, np.int32)
    ranks = []
    for r in range(world):
        loc = M.localize(g, part, r) if world > 1 else g
        lst = F.scatter_to_local(g, loc, st, trs)[0] if world > 1 else st
        alln = (loc.myList_nod2D.astype(np.int64) - 1) if world > 1 else np.arange(g.Nh)
        dev = torch.device(f"cuda:{r % ndev}")
        st_d, _ = to_device(lst, [], dev)
        ctx = AdvB200(loc, nb_g[alln], device=r % ndev, max_tracers=1)
        t = lambda a, dt_=torch.float64: torch.as_tensor(np.ascontiguousarray(a[alln]), dtype=dt_, device=dev)   # noqa: E731
        out = [torch.zeros((loc.Nh, loc.nl), dtype=torch.float64, device=dev) for _ in range(3)] + [t(cfl_old)]
        ranks.append(dict(loc=loc, ctx=ctx, st=st_d, alln=alln, err=None, hbar=t(hbar), hbar_old=t(hbar_old), wflux=t(wflux),
                          zbar=torch.as_tensor(np.ascontiguousarray(g.zbar), dtype=torch.float64, device=dev),
                          nmin=torch.as_tensor(np.ascontiguousarray(g.nlevels_nod2D_min[alln]), dtype=torch.int32, device=dev), out=out))
    if world > 1:
        comm_init_local([rk_["ctx"] for rk_ in ranks])
    torch.cuda.synchronize()

    def work(q):
        try:
            ctx, s = q["ctx"], q["st"]
            torch.cuda.set_device(ctx.device)
            ctx.set_state(s)
            ctx.vert_vel_ale_zlevel(dtc, True, 0.5, q["hbar"], q["hbar_old"], q["wflux"], q["nmin"], s.hnode_new, q["zbar"], mh, lz, *q["out"])
            ctx.synchronize()
        except Exception as ex:
            q["err"] = ex
    th = [threading.Thread(target=work, args=(q,)) for q in ranks]
    [x.start() for x in th]
    [x.join() for x in th]
    for q in ranks:
        if q["err"] is not None:
            raise q["err"]
        a = q["alln"]
        for got, ref, name in zip(q["out"], (W, we, wi, cfl), ("w", "w_e", "w_i", "cfl_z")):
            assert np.array_equal(got.cpu().numpy(), ref[a]), (name, world)
        assert np.array_equal(q["st"].hnode_new.cpu().numpy(), hn[a])
        q["ctx"].close()


def test_vert_vel_ale_zlevel_rejects_bad_arguments(pi_mesh):
    """cfl_z = NULL (nothing to read the previous CFL_z from) and lzstar_lev outside 1..16 are refused"""
    from fesom2_b200.driver import AdvB200, AdvError
    g = pi_mesh
    st, trs, nb_g, dt = make_case(g, 1)
    dev = torch.device("cuda:0")
    st_d, _ = to_device(st, [], dev)
    ctx = AdvB200(g, nb_g, device=0, max_tracers=1)
    ctx.set_state(st_d)
    z = lambda n, dt_=torch.float64: torch.zeros(n, dtype=dt_, device=dev)   # noqa: E731
    out = [z((g.Nh, g.nl)) for _ in range(4)]
    args = (z(g.Nh), z(g.Nh), z(g.Nh), torch.ones(g.Nh, dtype=torch.int32, device=dev), st_d.hnode_new, z(g.nl))
    with pytest.raises(AdvError):
        ctx.vert_vel_ale_zlevel(1.0, False, 1.0, *args, 0.5, 4, out[0], out[1], out[2], None)
    with pytest.raises(AdvError):
        ctx.vert_vel_ale_zlevel(1.0, False, 1.0, *args, 0.5, 17, *out)
    ctx.close()
