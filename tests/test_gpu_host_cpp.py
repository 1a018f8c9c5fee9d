"""-m gpu: the COMPILED host side above the C ABI (fesom2_b200/host: a C++ mirror of the reference's derived types and of
`do_oce_adv_tra(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh)`, plus the tracer dwarf of
dwarf/dwarf_tracer/dwarf_ini/fesom.F90:85-128 written against it).  The binary gets pageable host arrays from a file, drives
the model's call sequence -- state once per step, one call per tracer, the reference's single work set -- and its results
must equal the C restatement's dwarf iteration bit for bit.  No Python and no torch sit between the host program and the
library."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from common import make_case
from fesom2_b200 import build as B

pytestmark = pytest.mark.gpu


def write_case(path, g, st, trs, nb, dt, nsteps, ltra_diag=True, dvd=False, partitioned=False):
    """partitioned: 'FADW' = a rank's local mesh with mype / npes and the com_nod2D lists after the header"""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32).tobytes()       # noqa: E731
    f64 = lambda a: np.ascontiguousarray(np.asarray(a), dtype=np.float64).tobytes()   # noqa: E731
    with open(path, "wb") as f:
        f.write(b"FADW" if partitioned else b"FADV")
        f.write(struct.pack("<11i", g.nl, g.N, g.eDim_nod2D, g.T, g.eDim_elem2D, g.E, g.nod_in_elem2D.shape[1], len(trs), nsteps,
                            int(bool(st.use_wsplit)), int(dvd)))
        f.write(struct.pack("<d", dt))
        if partitioned:
            c = g.com_nod2D
            f.write(struct.pack("<6i", g.mype, g.npes, c.rPEnum, c.sPEnum, len(c.rlist), len(c.slist)))
            for a in (c.rPE, c.rptr, c.rlist, c.sPE, c.sptr, c.slist):
                f.write(i32(a))
        for name in ("edges", "edge_tri", "elem2D_nodes", "nod_in_elem2D", "nod_in_elem2D_num", "nlevels", "ulevels",
                     "nlevels_nod2D", "ulevels_nod2D"):
            f.write(i32(getattr(g, name)))
        f.write(i32(nb))
        for name in ("edge_cross_dxdy", "edge_dxdy", "elem_cos", "area", "areasvol"):
            f.write(f64(getattr(g, name)))
        for name in ("uv", "w", "w_e", "w_i", "helem", "hnode", "hnode_new", "zbar_3d_n", "Z_3d_n", "zbar_n_bot"):
            f.write(f64(getattr(st, name).numpy()))
        for t in trs:
            f.write(("%-8s%-8s%-8s" % (t.tra_adv_hor, t.tra_adv_ver, t.tra_adv_lim)).encode())
            f.write(struct.pack("<2di", t.tra_adv_ph, t.tra_adv_pv, int(ltra_diag)))
            f.write(f64(t.values.numpy())); f.write(f64(t.valuesAB.numpy())); f.write(f64(t.edge_up_dn_grad.numpy()))


def run_dwarf(g, st, trs, nb, dt, nsteps, **kw):
    exe = B.build_host()
    d = tempfile.mkdtemp()
    case, res = os.path.join(d, "case.bin"), os.path.join(d, "result.bin")
    write_case(case, g, st, trs, nb, dt, nsteps, **kw)
    p = subprocess.run([exe, case, res], capture_output=True, text=True, timeout=600)
    return p, res


@pytest.mark.parametrize("which,hor,ver,lim,wsplit", [("pi", "MUSCL", "QR4C", "FCT", False), ("cavity", "MFCT", "PPM", "FCT", True),
                                                      ("nw2", "MFCT", "QR4C", "NON", False)])
def test_cpp_dwarf_matches_the_oracle(pi_mesh, cav_mesh, nw2_mesh, which, hor, ver, lim, wsplit):
    from oracle import oracle_py as O
    g = {"pi": pi_mesh, "cavity": cav_mesh, "nw2": nw2_mesh}[which]
    nsteps = 3
    st, trs, nb, dt = make_case(g, 3, hor, ver, lim, ph=0.25, pv=0.75, use_wsplit=wsplit)
    rk = O.OracleRank(g, st, trs, nb, tra_diag=True, dvd=True)
    O.run([rk], dt, nsteps, 1)                           # dwarf iteration: zero del_ttf_adv*, advect, update values
    p, res = run_dwarf(g, st, trs, nb, dt, nsteps, dvd=True)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.strip().splitlines()
    assert len(lines) == nsteps * 3                      # the dwarf's min / max / sum line per call (fesom.F90:99)
    n = g.Nh * g.L
    out = np.fromfile(res, dtype=np.float64)
    pos = 0

    def take(shape):
        nonlocal pos
        k = int(np.prod(shape))
        a = out[pos:pos + k].reshape(shape)
        pos += k
        return a
    for k in range(3):
        assert np.array_equal(take((g.Nh, g.L)), rk.values[k]), ("values", k)
        assert np.array_equal(take((g.Nh, g.L)), rk.dttf_h[k]), ("del_ttf_advhoriz", k)
        assert np.array_equal(take((g.Nh, g.L)), rk.dttf_v[k]), ("del_ttf_advvert", k)
    lev = np.arange(1, g.L + 1)[None, :]
    wet = (lev >= np.asarray(g.ulevels_nod2D)[:, None]) & (lev <= np.asarray(g.nlevels_nod2D)[:, None] - 1)
    tah, tav = take((3, g.Nh, g.L)), take((3, g.Nh, g.L))
    for k in range(3):
        assert np.array_equal(tah[k][wet], rk.tra_advhoriz[k][wet]) and np.array_equal(tav[k][wet], rk.tra_advvert[k][wet])
        assert (tah[k][~wet] == 0).all()
    fh, fv = take((2, g.E, g.L)), take((2, g.N, g.nl))
    for k in range(2):
        assert np.array_equal(fv[k], rk.dvd_trflx_ver[k])
        if which != "cavity":
            assert np.array_equal(fh[k], rk.dvd_trflx_hor[k])
    assert pos == out.size and n > 0
    last = lines[-1].split()
    assert int(last[0]) == nsteps and int(last[1]) == 3


def test_cpp_dwarf_unknown_scheme_ends_in_par_ex(pi_mesh):
    """the reference prints 'Unknown ... advection type' and calls par_ex -> MPI_ABORT (src/oce_adv_tra_driver.F90:351-353)"""
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 1, "MFCT", "QR4C", "FCT")
    trs[0].tra_adv_ver = "QR5C"
    p, _ = run_dwarf(g, st, trs, nb, dt, 1)
    assert p.returncode == 1
    assert "Unknown vertical advection type QR5C" in p.stderr and "par_ex" in p.stderr


def test_cpp_dwarf_from_reference_format_restarts(pi_mesh, tmp_path):
    """`dwarf_tracer_b200 --restart <npepath>`: the dwarf as the reference ships it -- derived-type binary restarts in
    (read_all_bin_restarts), tracer 1, dt = 1.e-3, ten iterations with del_ttf carried over (fesom.F90:85-128, literally) --
    against the same loop around the C restatement"""
    from fesom2_b200 import restart as R
    from oracle import oracle_py as O
    g = pi_mesh
    nsteps, dt = 10, 1.0e-3
    st, trs, nb, _ = make_case(g, 1, "MUSCL", "QR4C", "FCT")
    npepath = str(tmp_path / "np1")
    R.dump_dwarf(npepath, g, st, trs, nb)
    res = str(tmp_path / "result.bin")
    p = subprocess.run([B.build_host(), "--restart", npepath, res], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert len(p.stdout.strip().splitlines()) == nsteps
    rk = O.OracleRank(g, st, trs, nb)
    lev = np.arange(1, g.L + 1)[None, :]
    wet = (lev >= np.asarray(g.ulevels_nod2D)[:, None]) & (lev <= np.asarray(g.nlevels_nod2D)[:, None] - 1)
    del_ttf = np.zeros((g.Nh, g.L))
    hnn = st.hnode_new.numpy()
    for _ in range(nsteps):
        rk.dttf_h[0][...] = 0.0
        rk.dttf_v[0][...] = 0.0
        O.run([rk], dt, 1, 0)
        del_ttf[wet] = (del_ttf + rk.dttf_h[0] + rk.dttf_v[0])[wet]
        own = wet.copy()
        own[g.N:] = False
        rk.values[0][own] = (rk.values[0] + del_ttf / np.where(wet, hnn, 1.0))[own]
    out = np.fromfile(res, dtype=np.float64).reshape(4, g.Nh, g.L)
    assert np.array_equal(out[0], rk.values[0]) and np.array_equal(out[1], del_ttf)
    assert np.array_equal(out[2], rk.dttf_h[0]) and np.array_equal(out[3], rk.dttf_v[0])
    vals = [float(x) for x in p.stdout.strip().splitlines()[-1].split()]
    assert len(vals) == 3 and vals[0] <= vals[1]


@pytest.mark.parametrize("which,world", [("pi", 2), ("pi", 8), ("cavity", 2)])
def test_cpp_dwarf_ranks_as_threads(pi_mesh, cav_mesh, which, world, tmp_path):
    """`dwarf_tracer_b200 --ranks N`: the reference's dist_N partition, one context and one host thread per rank, the halo
    exchanges of do_oce_adv_tra inside the library (in-process communicator), exchange_nod(values) between the ranks' host
    arrays -- three dwarf iterations of two tracers; every rank's nodes (owned and halo) equal the one-rank C restatement"""
    from fesom2_b200 import fields as F, mesh as M
    from oracle import oracle_py as O
    g = {"pi": pi_mesh, "cavity": cav_mesh}[which]
    nsteps = 3
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", "FCT")
    one = O.OracleRank(g, st, trs, nb)
    O.run([one], dt, nsteps, 1)
    part = g.parts[world]
    locs = []
    for r in range(world):
        loc = M.localize(g, part, r)
        lst, ltr = F.scatter_to_local(g, loc, st, trs)
        write_case(str(tmp_path / f"case.{r}"), loc, lst, ltr, nb[loc.myList_nod2D - 1], dt, nsteps, partitioned=True)
        locs.append(loc)
    p = subprocess.run([B.build_host(), "--ranks", str(world), str(tmp_path / "case"), str(tmp_path / "result")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr
    assert len(p.stdout.strip().splitlines()) == nsteps * 2          # rank 0 prints
    for r, loc in enumerate(locs):
        out = np.fromfile(str(tmp_path / f"result.{r}"), dtype=np.float64)
        n = loc.Nh * loc.L
        alln = loc.myList_nod2D.astype(np.int64) - 1
        own = alln[:loc.N]
        for k in range(2):
            vals, dh, dv = (out[(3 * k + j) * n:(3 * k + j + 1) * n].reshape(loc.Nh, loc.L) for j in range(3))
            assert np.array_equal(vals, one.values[k][alln]), (r, k, "values incl. halo")
            assert np.array_equal(dh[:loc.N], one.dttf_h[k][own]) and np.array_equal(dv[:loc.N], one.dttf_v[k][own]), (r, k)
