"""-m gpu: the partitioned mesh on N ranks against the 1-rank CPU oracle.  Owned-node results must be
bit-identical to the single-rank run (the reference guarantees it, src/gen_comm.F90:172-196).

Two transports for the same launch sequence (boundary-first ordering, interior overlap, K3 over S + halo):
  * in-process communicator (adv_ctx_comm_init_local, tests/local_ranks.py): N contexts in this process, one host
    thread each, on however many GPUs the box has -- these tests run on a ONE-GPU box;
  * NCCL (adv_ctx_comm_init, tests/mgpu_worker.py): one process per GPU, needs >= N GPUs."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _oracle_1rank(g, st, trs, dt, nsteps):
    from oracle import oracle_py as O
    one = O.OracleRank(g, st, trs, M.nboundary_lay(g))
    O.run([one], dt, nsteps, 0 if nsteps == 1 else 1)
    return one


def _check_owned(res, one, ntr=2):
    for r in res:
        own = r["owned"] - 1
        n = len(own)
        for k in range(ntr):
            assert np.array_equal(r["dv"][k][:n], one.dttf_v[k][own])
            # del_ttf_advhoriz: owned nodes complete; halo nodes hold the partial sums of the local
            # edges, exactly like the reference's scatter -- compare owned only
            assert np.array_equal(r["dh"][k][:n], one.dttf_h[k][own])


# ------------------------------------------------------------------------------- in-process ranks (any GPU count)
@pytest.mark.parametrize("mesh_name,world", [("pi", 2), ("synth", 2), ("pi", 8), ("soufflet", 8), ("synth", 5), ("neverworld2", 8), ("pi_cavity", 2)])
def test_local_ranks_match_single_rank_oracle(mesh_name, world):
    import mgpu_worker
    from local_ranks import run_local_ranks
    if mesh_name == "soufflet":
        g = M.load_npz_mesh(os.path.join(os.path.dirname(__file__), "golden", "mesh_soufflet.npz"))
        part = g.parts[world]
        st = F.make_state(g, "cpu")
        dt = F.cfl_dt(g, st, 0.3)
        trs = F.make_tracers(g, 2, "cpu", hor="MFCT", ver="QR4C", lim="FCT")
    else:
        g, part, st, trs, dt = mgpu_worker.build_global(mesh_name, world)
    res = run_local_ranks(g, part, st, trs, dt, exchange_inputs=True)
    one = _oracle_1rank(g, st, trs, dt, 1)
    _check_owned(res, one)
    bytes_sent, comm_ms, exposed_ms = res[0]["halo_stats"]
    assert bytes_sent > 0 and min(comm_ms) > 0.0 and min(exposed_ms) >= 0.0


@pytest.mark.parametrize("hor,ver", [("MUSCL", "PPM"), ("UPW1", "UPW1"), ("MFCT", "CDIFF")])
def test_local_ranks_other_schemes(hor, ver):
    import mgpu_worker
    from local_ranks import run_local_ranks
    g, part, st, trs, dt = mgpu_worker.build_global("pi", 2, hor=hor, ver=ver)
    res = run_local_ranks(g, part, st, trs, dt)
    _check_owned(res, _oracle_1rank(g, st, trs, dt, 1))


def test_local_ranks_three_dwarf_iterations():
    """do_oce_adv_tra + value update + exchange_nod(values), three times (dwarf loop)"""
    import mgpu_worker
    from local_ranks import run_local_ranks
    g, part, st, trs, dt = mgpu_worker.build_global("pi", 2)
    res = run_local_ranks(g, part, st, trs, dt, nsteps=3)
    one = _oracle_1rank(g, st, trs, dt, 3)
    for r in res:
        alln = r["all_nodes"] - 1
        for k in range(2):
            assert np.array_equal(r["values"][k], one.values[k][alln])


@pytest.mark.parametrize("mesh_name,world", [("pi", 2), ("pi", 8), ("synth", 3)])
def test_local_ranks_device_gradients_with_exchange_elem(mesh_name, world):
    """edge_up_dn_grad = NULL on N ranks: tracer_gradient_elements on the own elements, exchange_elem(tr_xy) over
    com_elem2D_full, fill_up_dn_grad, then the step -- bit-identical to the 1-rank oracle chain"""
    import mgpu_worker
    from local_ranks import run_local_ranks
    from oracle import oracle_py as O
    g, part, st, trs, dt = mgpu_worker.build_global(mesh_name, world)
    tri = F.find_up_downwind_triangles(g)
    for t in trs:
        t.edge_up_dn_grad = torch.as_tensor(O.fill_up_dn_grad(g, O.tracer_gradient_elements(g, t.values.numpy()), tri))
    res = run_local_ranks(g, part, st, trs, dt, null_grad=True, tri=tri)
    _check_owned(res, _oracle_1rank(g, st, trs, dt, 1))


def test_local_comm_rejects_wrong_rank_layout(small_mesh):
    from fesom2_b200.driver import AdvB200, AdvError, comm_init_local, ADV_EINVAL
    a = AdvB200(small_mesh, M.nboundary_lay(small_mesh))
    with pytest.raises(AdvError) as ei:
        comm_init_local([a, a])
    assert ei.value.code == ADV_EINVAL
    a.close()


# ------------------------------------------------------------------------------- NCCL, one process per GPU
def _run_nccl(world, mesh_name, nsteps=1):
    import mgpu_worker
    out = tempfile.mkdtemp()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(mgpu_worker.worker, args=(world, "nccl", mesh_name, port, out, nsteps), nprocs=world, join=True)
    return [np.load(os.path.join(out, f"rank{r}.npy"), allow_pickle=True)[0] for r in range(world)]


@pytest.mark.parametrize("mesh_name,world", [("pi", 2), ("synth", 2), ("pi", 8)])
def test_nccl_ranks_match_single_rank_oracle(mesh_name, world):
    if _ngpu() < world:
        pytest.skip(f"NCCL needs one GPU per rank ({world}); the same launch sequence runs in test_local_ranks_* on this box")
    import mgpu_worker
    res = _run_nccl(world, mesh_name)
    g, part, st, trs, dt = mgpu_worker.build_global(mesh_name, world)
    _check_owned(res, _oracle_1rank(g, st, trs, dt, 1))


def test_nccl_three_dwarf_iterations(world=2):
    if _ngpu() < world:
        pytest.skip(f"NCCL needs one GPU per rank ({world}); see test_local_ranks_three_dwarf_iterations")
    import mgpu_worker
    res = _run_nccl(world, "pi", nsteps=3)
    g, part, st, trs, dt = mgpu_worker.build_global("pi", world)
    one = _oracle_1rank(g, st, trs, dt, 3)
    for r in res:
        alln = r["all_nodes"] - 1
        for k in range(2):
            assert np.array_equal(r["values"][k], one.values[k][alln])
