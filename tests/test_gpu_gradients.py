"""-m gpu: the producer of edge_up_dn_grad -- tracer_gradient_elements
(src/oce_tracer_mod.F90:146-188) and fill_up_dn_grad (src/oce_muscl_adv.F90:356-525) on the device,
bit for bit against the C restatement (oracle/adv_oracle.c), and the advection step fed by them."""
import numpy as np
import pytest
import torch

from common import make_case, run_oracle, to_device
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

pytestmark = pytest.mark.gpu


def _case(mesh, ntr=2):
    from fesom2_b200.driver import AdvB200
    from oracle import oracle_py as O
    dev = torch.device("cuda:0")
    tri = F.find_up_downwind_triangles(mesh)
    vals = [F.make_tracer_values(mesh, "cpu", kind=k)[0] for k in range(ntr)]
    ref_xy = [O.tracer_gradient_elements(mesh, v.numpy()) for v in vals]
    ref_g = [O.fill_up_dn_grad(mesh, x, tri) for x in ref_xy]
    ctx = AdvB200(mesh, M.nboundary_lay(mesh), max_tracers=ntr)
    ctx.set_gradient_mesh(tri)
    ttf = [v.to(dev) for v in vals]
    n_elem = ref_xy[0].shape[0]
    tr_xy = [torch.zeros((n_elem, mesh.L, 2), dtype=torch.float64, device=dev) for _ in vals]
    grad = [torch.zeros((mesh.E, mesh.L, 4), dtype=torch.float64, device=dev) for _ in vals]
    ctx.tracer_gradient_elements(ttf, tr_xy)
    ctx.fill_up_dn_grad(tr_xy, grad)
    ctx.synchronize()
    return ctx, ref_xy, ref_g, tr_xy, grad


@pytest.mark.parametrize("which", ["pi", "soufflet", "small"])
def test_gradients_match_the_oracle(which, pi_mesh, souf_mesh, small_mesh):
    mesh = {"pi": pi_mesh, "soufflet": souf_mesh, "small": small_mesh}[which]
    ctx, ref_xy, ref_g, tr_xy, grad = _case(mesh)
    for k in range(2):
        assert np.array_equal(tr_xy[k].cpu().numpy(), ref_xy[k])
        got = grad[k].cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref_g[k]))
        assert np.array_equal(np.nan_to_num(got), np.nan_to_num(ref_g[k]))
    ctx.close()


def test_untouched_entries_stay_untouched(small_mesh):
    """the reference writes edge_up_dn_grad only inside the loop bounds of :388-524"""
    from oracle import oracle_py as O
    mesh = small_mesh
    ctx, ref_xy, _, tr_xy, _ = _case(mesh, 1)
    dev = tr_xy[0].device
    tri = F.find_up_downwind_triangles(mesh)
    sentinel = np.full((mesh.E, mesh.L, 4), -777.0)
    ref = O.fill_up_dn_grad(mesh, ref_xy[0], tri, out=sentinel.copy())
    g = torch.as_tensor(sentinel.copy(), device=dev)
    ctx.fill_up_dn_grad(tr_xy, [g])
    ctx.synchronize()
    assert np.array_equal(g.cpu().numpy(), ref)
    assert (ref == -777.0).any()
    ctx.close()


def test_advection_fed_by_device_gradients(souf_mesh):
    """values -> tr_xy -> edge_up_dn_grad -> do_oce_adv_tra entirely on the device == oracle chain"""
    from oracle import oracle_py as O
    mesh = souf_mesh
    st, trs, nb, dt = make_case(mesh, 2, "MFCT", "QR4C", "FCT")
    tri = F.find_up_downwind_triangles(mesh)
    for t in trs:                                                   # the oracle chain's gradients
        t.edge_up_dn_grad = torch.as_tensor(O.fill_up_dn_grad(mesh, O.tracer_gradient_elements(mesh, t.values.numpy()), tri))
    ora = run_oracle(mesh, st, trs, nb, dt)
    from fesom2_b200.driver import AdvB200
    dev = torch.device("cuda:0")
    st_d, trs_d = to_device(st, trs, dev)
    ctx = AdvB200(mesh, nb, max_tracers=2)
    ctx.set_gradient_mesh(tri)
    tr_xy = [torch.zeros((mesh.elem_area.shape[0], mesh.L, 2), dtype=torch.float64, device=dev) for _ in trs]
    for t in trs_d:
        t.edge_up_dn_grad = torch.zeros_like(t.edge_up_dn_grad)
    ctx.tracer_gradient_elements([t.values for t in trs_d], tr_xy)
    ctx.fill_up_dn_grad(tr_xy, [t.edge_up_dn_grad for t in trs_d])
    ctx.set_state(st_d)
    dh = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
    for k in range(2):
        assert np.array_equal(dh[k].cpu().numpy(), ora.dttf_h[k])
        assert np.array_equal(dv[k].cpu().numpy(), ora.dttf_v[k])
    ctx.close()


def test_call_order_is_checked(small_mesh):
    from fesom2_b200.driver import AdvB200, AdvError, ADV_ESTATE
    ctx = AdvB200(small_mesh, M.nboundary_lay(small_mesh), max_tracers=1)
    dev = torch.device("cuda:0")
    t = torch.zeros((small_mesh.Nh, small_mesh.L), dtype=torch.float64, device=dev)
    x = torch.zeros((small_mesh.T, small_mesh.L, 2), dtype=torch.float64, device=dev)
    with pytest.raises(AdvError) as ei:
        ctx.tracer_gradient_elements([t], [x])
    assert ei.value.code == ADV_ESTATE
    ctx.close()


@pytest.mark.parametrize("host", [False, True])
def test_null_gradient_pointer_lets_the_library_compute_them(souf_mesh, host):
    """edge_up_dn_grad = NULL in the tracer descriptor: the library runs tracer_gradient_elements +
    fill_up_dn_grad itself (saves the 4 E L words of H2D per tracer on the host-pointer path)"""
    from oracle import oracle_py as O
    from fesom2_b200.driver import AdvB200
    mesh = souf_mesh
    st, trs, nb, dt = make_case(mesh, 3, "MFCT", "QR4C", "FCT")
    trs[2].tra_adv_hor = "MUSCL"
    tri = F.find_up_downwind_triangles(mesh)
    for t in trs:
        t.edge_up_dn_grad = torch.as_tensor(O.fill_up_dn_grad(mesh, O.tracer_gradient_elements(mesh, t.values.numpy()), tri))
    ora = run_oracle(mesh, st, trs, nb, dt)
    dev = "cpu" if host else torch.device("cuda:0")
    st_d, trs_d = (st, trs) if host else to_device(st, trs, dev)
    for t in trs_d:
        t.edge_up_dn_grad = None
    ctx = AdvB200(mesh, nb, max_tracers=3)
    ctx.set_gradient_mesh(tri)
    ctx.set_state(st_d)
    dh = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((mesh.Nh, mesh.L), dtype=torch.float64, device=dev) for _ in trs]
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
    for k in range(3):
        assert np.array_equal(dh[k].cpu().numpy(), ora.dttf_h[k])
        assert np.array_equal(dv[k].cpu().numpy(), ora.dttf_v[k])
    ctx.close()
