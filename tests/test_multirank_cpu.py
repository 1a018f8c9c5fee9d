"""CPU tests of the N>1 path: partition lists against the reference's checked-in dist_N files,
world_size-2 gloo exchange over those lists, and 1-rank vs N-rank identity of the oracle."""
import os
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M

REF_MESHES = "/root/reference/test/meshes"


def _spawn(world, backend, mesh_name, nsteps=1):
    import mgpu_worker
    out = tempfile.mkdtemp()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(mgpu_worker.worker, args=(world, backend, mesh_name, port, out, nsteps), nprocs=world, join=True)
    return [np.load(os.path.join(out, f"rank{r}.npy"), allow_pickle=True)[0] for r in range(world)]


@pytest.mark.parametrize("mesh_name", ["pi", "synth"])
def test_gloo_world2_halo_exchange(mesh_name):
    """two processes, gloo: the send/recv lists deliver every owner's column into the halo tail"""
    res = _spawn(2, "gloo", mesh_name)
    assert all(r["halo_ok"] for r in res)
    owned = np.concatenate([r["owned"] for r in res])
    assert len(np.unique(owned)) == len(owned)          # every node has exactly one owner


@pytest.mark.skipif(not os.path.isdir(REF_MESHES), reason="reference fixtures not mounted")
@pytest.mark.parametrize("name,npes", [("pi", 2), ("pi", 8), ("soufflet", 2), ("soufflet", 8), ("pi_cavity", 2), ("neverworld2", 2), ("neverworld2", 8)])
def test_localize_reproduces_reference_dist_files(name, npes):
    """mesh.localize == the reference's own partition bookkeeping (test/meshes/*/dist_N)"""
    g = M.read_fesom_mesh(os.path.join(REF_MESHES, name), cyclic_length_deg={"soufflet": 4.5, "neverworld2": 60.0}.get(name, 360.0))
    d = M.read_dist(os.path.join(REF_MESHES, name), npes)
    for r in range(npes):
        m = M.localize(g, d["part"], r)
        info = d["ranks"][r]
        c = info["com_nod2D"]
        assert m.N == info["myDim_nod2D"] and m.eDim_nod2D == info["eDim_nod2D"]
        assert np.array_equal(m.myList_nod2D, info["myList_nod2D"])
        assert np.array_equal(m.myList_elem2D, info["myList_elem2D"][:m.T])
        assert np.array_equal(m.myList_edge2D, info["myList_edge2D"][:m.E])
        for a in ("rPE", "rptr", "rlist", "sPE", "sptr", "slist"):
            assert np.array_equal(getattr(m.com_nod2D, a), getattr(c, a)), a


def test_npz_fixture_matches_reference_mesh():
    if not os.path.isdir(REF_MESHES):
        pytest.skip("reference fixtures not mounted")
    here = os.path.dirname(os.path.abspath(__file__))
    a = M.load_npz_mesh(os.path.join(here, "golden", "mesh_pi.npz"))
    b = M.read_fesom_mesh(os.path.join(REF_MESHES, "pi"))
    assert np.array_equal(a.edges, b.edges) and np.array_equal(a.elem2D_nodes, b.elem2D_nodes)
    assert np.allclose(a.area, b.area, rtol=1e-9) and np.array_equal(a.nlevels, b.nlevels)
    assert np.array_equal(a.parts[2], M.read_dist(os.path.join(REF_MESHES, "pi"), 2)["part"])


@pytest.mark.parametrize("npes", [2, 8])
def test_oracle_1rank_vs_nrank_identical_on_owned_nodes(pi_mesh, npes):
    """with the reference's numbering every owned node sees the same sequence of
    addends on 1 or N ranks, so the results are bit-identical"""
    from oracle import oracle_py as O
    g = pi_mesh
    st = F.make_state(g, "cpu")
    dt = F.cfl_dt(g, st, 0.3)
    trs = F.make_tracers(g, 2, "cpu", hor="MUSCL", ver="QR4C", lim="FCT")
    nbg = M.nboundary_lay(g)
    one = O.OracleRank(g, st, trs, nbg)
    O.run([one], dt, 3, 1)
    ranks = []
    for r in range(npes):
        loc = M.localize(g, g.parts[npes], r)
        lst, ltr = F.scatter_to_local(g, loc, st, trs)
        ranks.append(O.OracleRank(loc, lst, ltr, nbg[loc.myList_nod2D - 1]))
    O.run(ranks, dt, 3, 1)
    for rk in ranks:
        loc = rk.mesh_py
        own = loc.myList_nod2D[:loc.N] - 1
        alln = loc.myList_nod2D - 1
        for k in range(2):
            assert np.array_equal(rk.dttf_v[k][:loc.N], one.dttf_v[k][own])
            assert np.array_equal(rk.dttf_h[k][:loc.N], one.dttf_h[k][own])
            assert np.array_equal(rk.values[k], one.values[k][alln])     # halo refreshed by the exchange


@pytest.mark.skipif(not os.path.isdir(REF_MESHES), reason="reference fixtures not mounted")
@pytest.mark.parametrize("name,npes", [("pi", 2), ("pi", 8), ("soufflet", 2), ("soufflet", 8), ("pi_cavity", 2), ("neverworld2", 2), ("neverworld2", 8)])
def test_element_halo_reproduces_reference_dist_files(name, npes):
    """mesh.element_halo == communication_elemn + com_global2local of the reference: myList_elem2D with its
    eDim / eXDim tails and both element communicators, against test/meshes/*/dist_N"""
    g = M.read_fesom_mesh(os.path.join(REF_MESHES, name), cyclic_length_deg={"soufflet": 4.5, "neverworld2": 60.0}.get(name, 360.0))
    d = M.read_dist(os.path.join(REF_MESHES, name), npes)
    for r in range(npes):
        h, info = M.element_halo(g, d["part"], r), d["ranks"][r]
        for k in ("myDim_elem2D", "eDim_elem2D", "eXDim_elem2D"):
            assert h[k] == info[k], k
        assert np.array_equal(h["myList_elem2D"], info["myList_elem2D"])
        for cname in ("com_elem2D", "com_elem2D_full"):
            for a in ("rPE", "rptr", "rlist", "sPE", "sptr", "slist"):
                assert np.array_equal(getattr(h[cname], a), getattr(info[cname], a)), (cname, a)


@pytest.mark.parametrize("npes", [2, 4])
def test_element_halo_send_and_recv_lists_pair_up(npes):
    """exchange_elem over these lists is consistent on a mesh the reference ships no partition for: what rank
    r sends to p is, element by element, what p expects from r (global ids, same order)"""
    from fesom2_b200 import partition as P
    g = M.synth_mesh(30, 26, nl=12)
    part = P.partition(g, npes, "rcb")
    halos = [M.element_halo(g, part, r) for r in range(npes)]
    seen = 0
    for r in range(npes):
        hr = halos[r]
        for ci in ("com_elem2D", "com_elem2D_full"):
            cr = hr[ci]
            for k, p in enumerate(cr.sPE):
                sent = hr["myList_elem2D"][cr.slist[cr.sptr[k] - 1:cr.sptr[k + 1] - 1] - 1]
                cp = halos[int(p)][ci]
                j = int(np.flatnonzero(cp.rPE == r)[0])
                want = halos[int(p)]["myList_elem2D"][cp.rlist[cp.rptr[j] - 1:cp.rptr[j + 1] - 1] - 1]
                assert np.array_equal(sent, want), (r, int(p), ci)
                seen += sent.size
    assert seen > 0


@pytest.mark.parametrize("mesh_name,npes", [("pi", 2), ("pi", 8), ("soufflet", 8)])
def test_gradient_producer_is_rank_local_with_the_element_halo(mesh_name, npes, pi_mesh, souf_mesh):
    """SURVEY 8f row 1 on N ranks (oracle level): with tr_xy exchanged over the eDim + eXDim element halo
    (exchange_elem, src/oce_tracer_mod.F90:140) and nod_in_elem2D of the halo nodes taken from their owners
    (src/oce_mesh.F90:2064-2088), fill_up_dn_grad on a rank's local mesh gives, bit for bit, the rows of
    the 1-rank result for all its edges -- i.e. the element halo of mesh.element_halo is sufficient."""
    from types import SimpleNamespace
    from oracle import oracle_py as O
    g = {"pi": pi_mesh, "soufflet": souf_mesh}[mesh_name]
    part = g.parts[npes]
    v, _ = F.make_tracer_values(g, "cpu", kind=0)
    v = v.numpy()
    tri = F.find_up_downwind_triangles(g)
    xy_glob = O.tracer_gradient_elements(g, v)
    grad_glob = O.fill_up_dn_grad(g, xy_glob, tri)
    for r in range(npes):
        loc = M.localize(g, part, r)
        h = M.element_halo(g, part, r)
        elist = h["myList_elem2D"].astype(np.int64) - 1               # global ids in local order
        e_g2l = np.full(g.T, -1, np.int64)
        e_g2l[elist] = np.arange(elist.size)
        nodes = loc.myList_nod2D.astype(np.int64) - 1
        edges = loc.myList_edge2D.astype(np.int64) - 1
        # element neighbourhoods of ALL local nodes (halo nodes: the owner's list), local element numbers
        nie_g = g.nod_in_elem2D[nodes].astype(np.int64) - 1
        assert (e_g2l[nie_g[nie_g >= 0]] >= 0).all(), "an element around a local node is outside the element halo"
        nie_l = np.where(nie_g >= 0, e_g2l[np.maximum(nie_g, 0)] + 1, 0).astype(np.int32)
        tri_g = tri[edges].astype(np.int64) - 1
        assert (e_g2l[tri_g[tri_g >= 0]] >= 0).all(), "an up/down-wind triangle is outside the element halo"
        tri_l = np.where(tri_g >= 0, e_g2l[np.maximum(tri_g, 0)] + 1, 0).astype(np.int32)
        # tr_xy: own elements computed locally from the local (owned + halo) nodal values, halo part "exchanged"
        own = SimpleNamespace(elem2D_nodes=loc.elem2D_nodes, nlevels=loc.nlevels, ulevels=loc.ulevels,
                              gradient_sca=loc.gradient_sca, nl=g.nl, L=g.L, T=loc.T, eDim_elem2D=elist.size - loc.T)
        xy = O.tracer_gradient_elements(own, v[nodes])
        assert np.array_equal(xy[:loc.T], xy_glob[elist[:loc.T]])
        xy[loc.T:] = xy_glob[elist[loc.T:]]
        lm = SimpleNamespace(edges=loc.edges, nod_in_elem2D=nie_l, nod_in_elem2D_num=g.nod_in_elem2D_num[nodes],
                             nlevels=g.nlevels[elist], ulevels=g.ulevels[elist],
                             nlevels_nod2D=loc.nlevels_nod2D, ulevels_nod2D=loc.ulevels_nod2D,
                             nlevels_nod2D_min=loc.nlevels_nod2D_min, ulevels_nod2D_max=loc.ulevels_nod2D_max,
                             elem_area=g.elem_area[elist], nl=g.nl, L=g.L, E=loc.E)
        got = O.fill_up_dn_grad(lm, xy, tri_l)
        assert np.array_equal(got, grad_glob[edges]), (mesh_name, npes, r)


@pytest.mark.parametrize("lim", ["FCT", "NON"])
def test_oracle_diagnostics_1rank_vs_nrank(pi_mesh, lim):
    """ltra_diag / ldiag_DVD on the reference's dist_2 partition: tra_advhoriz / tra_advvert of the owned nodes and
    dvd_trflx_ver of the owned nodes are identical on 1 and 2 ranks; dvd_trflx_hor of every local edge equals the global
    edge's value (an edge's flux needs its two end nodes only, and both are local)"""
    from oracle import oracle_py as O
    g = pi_mesh
    st = F.make_state(g, "cpu")
    dt = F.cfl_dt(g, st, 0.3)
    trs = F.make_tracers(g, 2, "cpu", hor="MFCT", ver="QR4C", lim=lim)
    nbg = M.nboundary_lay(g)
    one = O.OracleRank(g, st, trs, nbg, tra_diag=True, dvd=True)
    O.run([one], dt)
    ranks = []
    for r in range(2):
        loc = M.localize(g, g.parts[2], r)
        lst, ltr = F.scatter_to_local(g, loc, st, trs)
        ranks.append(O.OracleRank(loc, lst, ltr, nbg[loc.myList_nod2D - 1], tra_diag=True, dvd=True))
    O.run(ranks, dt)
    for rk in ranks:
        loc = rk.mesh_py
        own = loc.myList_nod2D[:loc.N] - 1
        edges = loc.myList_edge2D[:loc.E] - 1
        for k in range(2):
            assert np.array_equal(rk.tra_advhoriz[k][:loc.N], one.tra_advhoriz[k][own])
            assert np.array_equal(rk.tra_advvert[k][:loc.N], one.tra_advvert[k][own])
            assert np.array_equal(rk.dvd_trflx_ver[k], one.dvd_trflx_ver[k][own])
            assert np.array_equal(rk.dvd_trflx_hor[k], one.dvd_trflx_hor[k][edges])
