"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance: north_star demands 1e-12 relative per step (SURVEY Appendix C metric); the gather
formulation reproduces the serial summation order, so in practice the results are bit-identical
and the strictest tests assert exactly that."""
import itertools

import numpy as np
import pytest
import torch

from common import TOL_STEP, make_case, rel_err, run_cuda, run_oracle

pytestmark = pytest.mark.gpu


def _compare(mesh, ctx, dh, dv, ora, ntr, fct=True, exact=False):
    worst = 0.0
    for k in range(ntr):
        for name, got, ref in (("del_ttf_advhoriz", dh[k], ora.dttf_h[k]), ("del_ttf_advvert", dv[k], ora.dttf_v[k])):
            assert np.isfinite(got).all(), name
            e = rel_err(got, ref)
            worst = max(worst, e)
            assert e <= TOL_STEP, (k, name, e)
            if exact:
                assert np.array_equal(got, ref), (k, name, "not bit-identical", e)
    if fct:
        # the last tracer of the batch is what the oracle's shared work arrays hold
        k = ntr - 1
        N = mesh.N
        lo = ctx.get_work("fct_LO", k)
        assert rel_err(lo[:N], ora.keep["fct_LO"][:N]) <= TOL_STEP
        assert rel_err(ctx.get_work("fct_plus", k)[:N], ora.keep["fct_plus"][:N]) <= TOL_STEP
        assert rel_err(ctx.get_work("fct_minus", k)[:N], ora.keep["fct_minus"][:N]) <= TOL_STEP
    return worst


def test_config1_pi_muscl_qr4c_fct(pi_mesh):
    """BASELINE config 1: dwarf_tracer on the pi mesh, T+S MUSCL + QR4C + FCT (ph=0, pv=1)."""
    st, trs, nb, dt = make_case(pi_mesh, 2, "MUSCL", "QR4C", "FCT")
    ora = run_oracle(pi_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(pi_mesh, st, trs, nb, dt)
    _compare(pi_mesh, ctx, dh, dv, ora, 2, exact=True)
    ctx.close()


def test_config1_pi_dwarf_dt(pi_mesh):
    """same with the dwarf's literal dt = 1.e-3 (dwarf_ini/fesom.F90:97)."""
    st, trs, nb, _ = make_case(pi_mesh, 1, "MUSCL", "QR4C", "FCT")
    ora = run_oracle(pi_mesh, st, trs, nb, 1.0e-3)
    ctx, dh, dv = run_cuda(pi_mesh, st, trs, nb, 1.0e-3)
    _compare(pi_mesh, ctx, dh, dv, ora, 1, exact=True)
    ctx.close()


def test_config2_soufflet_mfct_qr4c_fct(souf_mesh):
    """BASELINE config 2: soufflet channel mesh (cyclic_length 4.5 deg), MFCT + QR4C + FCT."""
    st, trs, nb, dt = make_case(souf_mesh, 2, "MFCT", "QR4C", "FCT")
    ora = run_oracle(souf_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(souf_mesh, st, trs, nb, dt)
    _compare(souf_mesh, ctx, dh, dv, ora, 2, exact=True)
    ctx.close()


@pytest.mark.parametrize("hor,ver,lim,wsplit", [("MFCT", "QR4C", "FCT", False), ("MUSCL", "PPM", "FCT", True), ("MFCT", "CDIFF", "NON", False)])
def test_neverworld2_mesh(nw2_mesh, hor, ver, lim, wsplit):
    """the reference's fourth test mesh (test/meshes/neverworld2: 60-degree periodic sector, 15 layers, columns of 4 to 15
    layers): up to kMaxCols short columns share a CTA of the compacted node kernels"""
    g = nw2_mesh
    st, trs, nb, dt = make_case(g, 3, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    ora = run_oracle(g, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(g, st, trs, nb, dt)
    _compare(g, ctx, dh, dv, ora, 3, fct=(lim == "FCT"), exact=True)
    ctx.close()


@pytest.mark.parametrize("hor,ver,lim", list(itertools.product(("UPW1", "MUSCL", "MFCT"),
                                                               ("UPW1", "QR4C", "PPM", "CDIFF"), ("FCT", "NON"))))
def test_all_scheme_combinations(small_mesh, hor, ver, lim):
    """every tra_adv_hor x tra_adv_ver x tra_adv_lim the driver dispatches on
    (oce_adv_tra_driver.F90:343-379), with fractional num_ord"""
    st, trs, nb, dt = make_case(small_mesh, 2, hor, ver, lim, ph=0.25, pv=0.75)
    ora = run_oracle(small_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(small_mesh, st, trs, nb, dt)
    _compare(small_mesh, ctx, dh, dv, ora, 2, fct=(lim == "FCT"))
    ctx.close()


def test_odd_tracer_count_and_mixed_schemes(small_mesh):
    st, trs, nb, dt = make_case(small_mesh, 5, "MFCT", "QR4C", "FCT")
    trs[1].tra_adv_hor = "MUSCL"
    trs[3].tra_adv_ver = "PPM"
    trs[4].tra_adv_lim = "NON"
    ora = run_oracle(small_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(small_mesh, st, trs, nb, dt)
    _compare(small_mesh, ctx, dh, dv, ora, 5, fct=False)
    ctx.close()


def test_use_wsplit(small_mesh):
    """dynamics%use_wsplit: implicit vertical part adv_tra_vert_impl (oce_adv_tra_driver.F90:323-334)"""
    st, trs, nb, dt = make_case(small_mesh, 2, "MFCT", "QR4C", "FCT", use_wsplit=True)
    assert st.w_i.abs().max() > 0
    ora = run_oracle(small_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(small_mesh, st, trs, nb, dt)
    _compare(small_mesh, ctx, dh, dv, ora, 2)
    ctx.close()


def test_ale_thickness_change(small_mesh):
    """hnode_new /= hnode (zlevel / zstar ALE)"""
    st, trs, nb, dt = make_case(small_mesh, 2, "MFCT", "QR4C", "FCT", ale_amp=0.3)
    ora = run_oracle(small_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(small_mesh, st, trs, nb, dt)
    _compare(small_mesh, ctx, dh, dv, ora, 2)
    ctx.close()


def test_host_pointer_call(pi_mesh):
    """the HOST-pointer flavour of the C ABI (H2D / D2H inside the call)"""
    st, trs, nb, dt = make_case(pi_mesh, 2, "MFCT", "QR4C", "FCT")
    ora = run_oracle(pi_mesh, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(pi_mesh, st, trs, nb, dt, host_ptrs=True)
    _compare(pi_mesh, ctx, dh, dv, ora, 2, exact=True)
    ctx.close()


@pytest.mark.parametrize("host", [True, False])
def test_reference_call_order_one_tracer_per_call(souf_mesh, host):
    """The reference's own sequence (src/oce_ale_tracer.F90:260-312): state refreshed once per step, then
    do_oce_adv_tra once PER TRACER.  The drop-in wrapper calls adv_ctx_set_state_step with the model's step counter
    before every tracer: only the first call of a step uploads the state and computes the volume flux Q; a new step
    number takes the new state.  HOST arrays are pageable here (page-locked by the library on first use)."""
    from fesom2_b200.driver import AdvB200
    from fesom2_b200 import fields as F
    mesh = souf_mesh
    st, trs, nb, dt = make_case(mesh, 3, "MFCT", "QR4C", "FCT")
    ora = run_oracle(mesh, st, trs, nb, dt)
    st2 = F.OceanState(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.__dict__.items()})
    st2.uv *= 0.5
    st2.w *= 0.5
    st2.w_e *= 0.5
    ora2 = run_oracle(mesh, st2, trs, nb, dt)
    dev = "cpu" if host else torch.device("cuda:0")
    if host:
        st_d, st2_d, trs_d = st, st2, trs
    else:
        from common import to_device
        st_d, trs_d = to_device(st, trs, dev)
        st2_d, _ = to_device(st2, [], dev)
    ctx = AdvB200(mesh, nb, max_tracers=1)
    for step, (s_d, ref) in enumerate(((st_d, ora), (st2_d, ora2))):
        n0 = ctx.launch_count
        for k, t in enumerate(trs_d):
            dh = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev)]
            dv = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev)]
            ctx.set_state(s_d, step=step + 1)
            ctx.do_oce_adv_tra(dt, [t], dh, dv)
            assert np.array_equal(dh[0].cpu().numpy(), ref.dttf_h[k]), (step, k)
            assert np.array_equal(dv[0].cpu().numpy(), ref.dttf_v[k]), (step, k)
        assert ctx.launch_count - n0 == 3 * 4      # four launches per tracer call: Q is fused into the first edge kernel
    ctx.close()


def test_accumulates_into_del_ttf(small_mesh):
    """do_oce_adv_tra ACCUMULATES into del_ttf_advhoriz/advvert (driver :535,:556,:607)"""
    from fesom2_b200.driver import AdvB200
    from common import to_device
    st, trs, nb, dt = make_case(small_mesh, 1)
    dev = torch.device("cuda:0")
    st_d, trs_d = to_device(st, trs, dev)
    ctx = AdvB200(small_mesh, nb, max_tracers=1)
    ctx.set_state(st_d)
    dh = [torch.full((small_mesh.Nh, small_mesh.L), 1.5, dtype=torch.float64, device=dev)]
    dv = [torch.full((small_mesh.Nh, small_mesh.L), -2.5, dtype=torch.float64, device=dev)]
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
    from oracle import oracle_py as O
    rk = O.OracleRank(small_mesh, st, trs, nb)
    rk.dttf_h[0][:] = 1.5
    rk.dttf_v[0][:] = -2.5
    O.run([rk], dt, 1, 0)
    assert rel_err(dh[0].cpu().numpy(), rk.dttf_h[0]) <= TOL_STEP
    assert rel_err(dv[0].cpu().numpy(), rk.dttf_v[0]) <= TOL_STEP
    ctx.close()


def test_unknown_scheme_is_an_error(small_mesh):
    """unknown scheme string -> error code (the reference calls par_ex, driver :351-353,:373-375)"""
    from fesom2_b200.driver import AdvError, ADV_ESCHEME
    st, trs, nb, dt = make_case(small_mesh, 1)
    trs[0].tra_adv_hor = "WENO"
    with pytest.raises(AdvError) as ei:
        run_cuda(small_mesh, st, trs, nb, dt)
    assert ei.value.code == ADV_ESCHEME


@pytest.mark.parametrize("knobs", [{"ADV_BULK": "0"}, {"ADV_E1_D": "3", "ADV_E1_NG": "5"}, {"ADV_E1_D": "4", "ADV_E1_NG": "1"},
                                   {"ADV_G_LO": "6", "ADV_G_K2": "6", "ADV_G_K3": "6"}, {"ADV_G_LO": "2", "ADV_G_K2": "1", "ADV_G_K3": "1"}])
def test_kernel_variants_are_bit_identical(souf_mesh, monkeypatch, knobs):
    """every launch configuration of the library (register-gather or bulk-copy edge kernel, pipeline
    depth, groups per CTA, gather batch sizes) gives the same bits as the oracle-checked default"""
    st, trs, nb, dt = make_case(souf_mesh, 3, "MFCT", "QR4C", "FCT")     # 3 tracers: one TB=2 and one TB=1 chunk
    ora = run_oracle(souf_mesh, st, trs, nb, dt)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    ctx, dh, dv = run_cuda(souf_mesh, st, trs, nb, dt)
    _compare(souf_mesh, ctx, dh, dv, ora, 3, exact=True)
    ctx.close()


def test_unaligned_device_pointers_are_staged(small_mesh):
    """edge_up_dn_grad / uv that are only 8-byte aligned cannot be read as 16-byte words or bulk-copied:
    the library stages an aligned device copy, same result"""
    from fesom2_b200.driver import AdvB200
    from common import to_device
    st, trs, nb, dt = make_case(small_mesh, 2, "MUSCL", "QR4C", "FCT")
    ora = run_oracle(small_mesh, st, trs, nb, dt)
    dev = torch.device("cuda:0")
    st_d, trs_d = to_device(st, trs, dev)
    for t in trs_d:                                                   # shift the gradients by one double
        buf = torch.empty(t.edge_up_dn_grad.numel() + 1, dtype=torch.float64, device=dev)
        view = buf[1:].view_as(t.edge_up_dn_grad)
        view.copy_(t.edge_up_dn_grad)
        assert view.data_ptr() % 16 == 8
        t.edge_up_dn_grad = view
    buf = torch.empty(st_d.uv.numel() + 1, dtype=torch.float64, device=dev)
    uv = buf[1:].view_as(st_d.uv)
    uv.copy_(st_d.uv)
    st_d.uv = uv
    ctx = AdvB200(small_mesh, nb, max_tracers=2)
    ctx.set_state(st_d)
    dh = [torch.zeros((small_mesh.Nh, small_mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((small_mesh.Nh, small_mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
    _compare(small_mesh, ctx, [x.cpu().numpy() for x in dh], [x.cpu().numpy() for x in dv], ora, 2)
    ctx.close()


@pytest.mark.parametrize("order", [2, 3])
def test_init_tracers_AB(small_mesh, order):
    """the prologue of the tracer step (src/oce_tracer_mod.F90:13-123), bit for bit against the NumPy restatement"""
    from fesom2_b200.driver import AdvB200, AdvError, ADV_EINVAL
    from oracle import numpy_ref as R
    g = small_mesh
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(20261017)
    v = rng.normal(10.0, 3.0, (g.Nh, g.L))
    o = rng.normal(10.0, 3.0, (g.Nh, g.L, order - 1))
    ref_ab, ref_old = R.init_tracers_AB(v, o, order, 0.1)
    ctx = AdvB200(g, M_nb(g), max_tracers=1)
    tv, to = torch.as_tensor(v, device=dev), torch.as_tensor(o, device=dev).contiguous()
    tab = torch.empty_like(tv)
    d = [torch.full_like(tv, 7.0) for _ in range(3)]
    ctx.init_tracers_AB([tv], [to], [tab], order, 0.1, [d[0]], [d[1]], [d[2]])
    ctx.synchronize()
    assert np.array_equal(tab.cpu().numpy(), ref_ab)
    assert np.array_equal(to.cpu().numpy(), ref_old)
    assert all(float(x.abs().max()) == 0.0 for x in d)
    with pytest.raises(AdvError) as ei:
        ctx.init_tracers_AB([tv], [to], [tab], 4, 0.1)
    assert ei.value.code == ADV_EINVAL
    ctx.close()


def M_nb(g):
    from fesom2_b200 import mesh as M
    return M.nboundary_lay(g)


@pytest.mark.parametrize("hor,ver,lim", [("MFCT", "QR4C", "NON"), ("MUSCL", "QR4C", "FCT")])
def test_analytic_linear_fields(hor, ver, lim):
    """tests/test_analytic.py on the device: a linear tracer in a constant flow on an irregular triangulation -- the CUDA
    path gives the exact -dt h (u . grad T) in the interior (and the C restatement's bits everywhere)"""
    import test_analytic as A
    g, st, trs, nb, dt = A.linear_case(hor=hor, ver=ver, ph=0.5)
    trs[0].tra_adv_lim = lim
    ora = run_oracle(g, st, trs, nb, dt)
    ctx, dh, dv = run_cuda(g, st, trs, nb, dt)
    assert np.array_equal(dh[0], ora.dttf_h[0]) and np.array_equal(dv[0], ora.dttf_v[0])
    inner = A.interior_nodes(g, rings=3)
    exact = -dt * st.hnode.numpy() * (A.U0 * A.GA + A.V0 * A.GB)
    assert np.abs(dh[0] + dv[0] - exact)[inner].max() / np.abs(exact).max() <= 1e-9
    ctx.close()
