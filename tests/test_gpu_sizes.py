"""-m gpu: the remaining BASELINE configurations and the long run.

* config 3 (CORE2-sized synthetic mesh, 127k nodes x 47 layers) against the C oracle, bit for bit;
* config 5 (30-tracer RECOM-style batch on that mesh): every tracer of the batch against the oracle;
* 100 dwarf iterations on the pi mesh: north_star's 1e-10 after 100 steps (in practice bit-identical);
* the bench workload (config 4 share, 376k nodes x 70 layers), where the serial oracle is too slow for a
  test: size-independent properties -- a constant tracer stays constant, the FCT step introduces no new
  extrema (bounds recomputed here from ttf and fct_LO), and the step is linear in a
  tracer offset for the vertical upwind/horizontal upwind combination.
"""
import numpy as np
import pytest
import torch

from common import TOL_100, TOL_STEP, make_case, rel_err, run_cuda, run_oracle, to_device
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core2_mesh():
    return M.synth_mesh(357, 356, nl=48)


def test_config3_core2_sized_mfct_qr4c(core2_mesh):
    g = core2_mesh
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    ora = run_oracle(g, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(g, st, trs, nb, dt)
    for k in range(2):
        assert np.array_equal(dh[k], ora.dttf_h[k]) and np.array_equal(dv[k], ora.dttf_v[k])
    ctx.close()


def test_config5_thirty_tracer_batch():
    """one batched call for 30 tracers (RECOM-style); the oracle runs them one call at a time like the
    reference.  A 13k-node mesh keeps the 30 gradient fields small; bench.py --workload recom30 measures
    the CORE2-sized case."""
    g = M.synth_mesh(120, 110, nl=48)
    ntr = 30
    st, trs, nb, dt = make_case(g, ntr, "MFCT", "QR4C", "FCT")
    ctx, dh, dv = run_cuda(g, st, trs, nb, dt)
    assert ctx.launch_count == 4 * (ntr // 2)         # 15 chunks of two tracers, four kernels each
    for k in (0, 1, 14, 29):                             # the oracle is serial: spot-check four tracers
        ora = run_oracle(g, st, [trs[k]], nb, dt)
        assert np.array_equal(dh[k], ora.dttf_h[0]), k
        assert np.array_equal(dv[k], ora.dttf_v[0]), k
    ctx.close()


def test_hundred_dwarf_iterations(pi_mesh):
    """do_oce_adv_tra + values += del_ttf / hnode_new, 100 times (dwarf_ini/fesom.F90:85-128)"""
    from fesom2_b200.driver import AdvB200
    from oracle import oracle_py as O
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 2, "MUSCL", "QR4C", "FCT")
    rk = O.OracleRank(g, st, trs, nb)
    O.run([rk], dt, 100, 1)
    dev = torch.device("cuda:0")
    st_d, trs_d = to_device(st, trs, dev)
    ctx = AdvB200(g, nb, max_tracers=2)
    ctx.set_state(st_d)
    dh = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    for _ in range(100):
        for x in dh + dv:
            x.zero_()
        ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
        ctx.update_values([t.values for t in trs_d], dh, dv)
    ctx.synchronize()
    for k in range(2):
        got = trs_d[k].values.cpu().numpy()
        assert np.isfinite(got).all()
        assert rel_err(got, rk.values[k]) <= TOL_100
    ctx.close()


@pytest.fixture(scope="module")
def bench_case():
    g = M.synth_mesh(613, 613, nl=71)
    dev = torch.device("cuda:0")
    st = F.make_state(g, dev)
    dt = F.cfl_dt(g, st, 0.3)
    return g, st, dt, M.nboundary_lay(g), F.find_up_downwind_triangles(g)


def _step(g, st, nb, dt, trs):
    from fesom2_b200.driver import AdvB200
    dev = trs[0].values.device
    ctx = AdvB200(g, nb, max_tracers=len(trs))
    ctx.set_state(st)
    dh = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx.do_oce_adv_tra(dt, trs, dh, dv)
    return ctx, dh, dv


def _wet(g, dev):
    nlev = torch.as_tensor(np.asarray(g.nlevels_nod2D), device=dev)
    ulev = torch.as_tensor(np.asarray(g.ulevels_nod2D), device=dev)
    k = torch.arange(1, g.L + 1, device=dev)[None, :]
    return (k >= ulev[:, None]) & (k <= nlev[:, None] - 1)


def test_bench_size_constant_tracer_and_no_new_extrema(bench_case):
    g, st, dt, nb, tri = bench_case
    dev = torch.device("cuda:0")
    # tracer 0: the bench's temperature; tracer 1: a constant (w is derived from uv by continuity)
    t0 = F.make_tracers_kind(g, 0, dev, tri)[0]
    c = torch.full_like(t0.values, 3.25)
    t1 = F.TracerFields(values=c, valuesAB=c.clone(), edge_up_dn_grad=torch.zeros_like(t0.edge_up_dn_grad),
                        tra_adv_hor="MFCT", tra_adv_ver="QR4C", tra_adv_lim="FCT", tra_adv_ph=0.0, tra_adv_pv=1.0)
    ctx, dh, dv = _step(g, st, nb, dt, [t0, t1])
    wet = _wet(g, dev)
    new = [t.values + (dh[k] + dv[k]) / st.hnode_new for k, t in enumerate((t0, t1))]
    # (ii) constant stays constant to round-off
    assert float((new[1] - 3.25).abs()[wet].max()) <= 1e-11
    # (iv) no new extrema: within the range of {ttf, fct_LO} over the node's 3-D FCT cluster
    lo = torch.as_tensor(ctx.get_work("fct_LO", 0), device=dev)
    hi_n = torch.maximum(lo, t0.values)
    lo_n = torch.minimum(lo, t0.values)
    big = 1.0e30
    hi_n = torch.where(wet, hi_n, torch.full_like(hi_n, -big))
    lo_n = torch.where(wet, lo_n, torch.full_like(lo_n, big))
    e = torch.as_tensor(np.asarray(g.edges, dtype=np.int64) - 1, device=dev)     # (E, 2)
    cmax, cmin = hi_n.clone(), lo_n.clone()
    for a, b in ((0, 1), (1, 0)):
        cmax.index_reduce_(0, e[:, a], hi_n[e[:, b]], "amax")
        cmin.index_reduce_(0, e[:, a], lo_n[e[:, b]], "amin")
    vmax, vmin = cmax.clone(), cmin.clone()
    vmax[:, 1:] = torch.maximum(vmax[:, 1:], cmax[:, :-1]); vmax[:, :-1] = torch.maximum(vmax[:, :-1], cmax[:, 1:])
    vmin[:, 1:] = torch.minimum(vmin[:, 1:], cmin[:, :-1]); vmin[:, :-1] = torch.minimum(vmin[:, :-1], cmin[:, 1:])
    eps = 1e-12 * float(t0.values.abs().max())
    own = wet[: g.N]
    assert bool((new[0][: g.N][own] <= vmax[: g.N][own] + eps).all())
    assert bool((new[0][: g.N][own] >= vmin[: g.N][own] - eps).all())
    ctx.close()


def test_bench_size_batch_position_is_irrelevant(bench_case):
    """the same tracer gives the same bits as slot 0 or slot 1 of a chunk, alone or paired (idempotence of
    the batching: geometry amortisation must not change a result)"""
    g, st, dt, nb, tri = bench_case
    dev = torch.device("cuda:0")
    t0 = F.make_tracers_kind(g, 0, dev, tri)[0]
    t1 = F.make_tracers_kind(g, 1, dev, tri)[0]
    ctx, dh_a, dv_a = _step(g, st, nb, dt, [t0, t1]); ctx.close()
    ctx, dh_b, dv_b = _step(g, st, nb, dt, [t1, t0]); ctx.close()
    ctx, dh_c, dv_c = _step(g, st, nb, dt, [t0]); ctx.close()
    assert torch.equal(dh_a[0], dh_b[1]) and torch.equal(dv_a[0], dv_b[1])
    assert torch.equal(dh_a[1], dh_b[0]) and torch.equal(dv_a[1], dv_b[0])
    assert torch.equal(dh_a[0], dh_c[0]) and torch.equal(dv_a[0], dv_c[0])


def test_bench_size_against_the_oracle(bench_case):
    """The measured workload itself (config 4 per-GPU share, 376k nodes x 70 layers, T+S MFCT + QR4C + FCT): one step
    against the C oracle (serial, memory-lean driver: ~10 s), bit for bit on every node."""
    from oracle import oracle_py as O
    g, st, dt, nb, tri = bench_case
    dev = torch.device("cuda:0")
    trs = [F.make_tracers_kind(g, k, dev, tri)[0] for k in range(2)]
    ctx, dh, dv = _step(g, st, nb, dt, trs)
    lo = O.LeanOracle(g, {k: v.cpu().numpy() for k, v in st.__dict__.items() if torch.is_tensor(v)}, nb)
    for k, t in enumerate(trs):
        rh, rv = lo.run_tracer(t.values.cpu().numpy(), t.valuesAB.cpu().numpy(), t.edge_up_dn_grad.cpu().numpy(),
                               t.tra_adv_hor, t.tra_adv_ver, t.tra_adv_lim, t.tra_adv_ph, t.tra_adv_pv, dt)
        assert np.array_equal(dh[k].cpu().numpy(), rh)
        assert np.array_equal(dv[k].cpu().numpy(), rv)
    ctx.close()
