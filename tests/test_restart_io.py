"""the reference's derived-type binary restarts (t_mesh / t_partit / t_tracer / t_dynamics,
src/io_restart_derivedtype.F90:29-234).  No real dwarf input ships with the reference and no Fortran compiler exists
here, so the reader is pinned three ways: (i) its field lists against the WRITE_T_* procedures of the reference
SOURCES (names and order; kinds and ranks against the type declarations) whenever /root/reference is present,
(ii) a write -> read round trip through the same record format including gfortran's sub-record splitting,
(iii) byte-level checks of the record framing."""
import os
import re
import struct

import numpy as np
import pytest
import torch

from common import make_case
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M
from fesom2_b200 import restart as R

REF = "/root/reference/src"


def _routine(src, name):
    m = re.search(r"subroutine\s+" + name + r"\b(.*?)end subroutine\s+" + name, src, flags=re.S | re.I)
    assert m, name
    return m.group(1)


def _written(body, var):
    """item names in write order: plain write(unit) var%x, write_bin_array(var%x), write1d_int_static(var%x)"""
    out = []
    for line in body.splitlines():
        line = line.split("!")[0]
        m = re.search(r"(write\(unit[^)]*\)|call\s+write_bin_array\(|call\s+write1d_int_static\()\s*" + var + r"%([\w%]+)", line, flags=re.I)
        if m:
            out.append(m.group(2).lower())
    return out


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not on this machine")
def test_schemas_follow_the_reference_write_routines():
    mesh = open(os.path.join(REF, "MOD_MESH.F90")).read()
    assert _written(_routine(mesh, "write_t_mesh"), "mesh") == [n.lower() for _, n in R.T_MESH]
    part = open(os.path.join(REF, "MOD_PARTIT.F90")).read()
    assert _written(_routine(part, "WRITE_T_COM_STRUCT"), "tstruct") == [n.lower() for _, n in R.T_COM_STRUCT]
    assert _written(_routine(part, "WRITE_T_PARTIT"), "partit") == [n.lower() for _, n in R.T_PARTIT_TAIL]
    body = _routine(part, "WRITE_T_PARTIT")
    assert [x.lower() for x in re.findall(r"call partit%(\w+)%WRITE_T_COM_STRUCT", body)] == ["com_nod2d", "com_elem2d", "com_elem2d_full"]
    tra = open(os.path.join(REF, "MOD_TRACER.F90")).read()
    assert _written(_routine(tra, "WRITE_T_TRACER_DATA"), "tdata") == [n.lower() for _, n in R.T_TRACER_DATA]
    assert _written(_routine(tra, "WRITE_T_TRACER_WORK"), "twork") == [n.lower() for _, n in R.T_TRACER_WORK]
    dyn = open(os.path.join(REF, "MOD_DYN.F90")).read()
    assert _written(_routine(dyn, "WRITE_T_SOLVERINFO"), "tsolverinfo") == [n.lower() for _, n in R.T_SOLVERINFO]
    assert _written(_routine(dyn, "WRITE_T_DYN_WORK"), "twork") == [n.lower() for _, n in R.T_DYN_WORK]
    want = [n.lower() for _, n in R.T_DYN_HEAD + R.T_DYN_ARRAYS + R.T_DYN_FER + R.T_DYN_SE]
    assert _written(_routine(dyn, "WRITE_T_DYN"), "dynamics") == want


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not on this machine")
def test_schema_kinds_and_ranks_follow_the_type_declarations():
    def decls(text):
        info = {}
        for line in text.splitlines():
            line = line.split("!")[0]
            if "::" not in line:
                continue
            left, right = line.split("::", 1)
            m = re.match(r"\s*(integer(\(int32\))?|real\(kind=WP\)|logical|character\((\d+)\))\s*(,.*)?$", left.strip(), flags=re.I)
            if not m:
                continue
            base = m.group(1).lower()
            kind = "i" if base.startswith("integer") else "r" if base.startswith("real") else "l" if base == "logical" else "c" + m.group(3)
            attrs = (m.group(4) or "").lower()
            dim = re.search(r"dimension\(([^)]*)\)", attrs)
            for name, dims in re.findall(r"(\w+)\s*(\([^)]*\))?(?:\s*=\s*[^,]+)?", right):
                if not name or name[0].isdigit() or name.startswith("_"):
                    continue
                d = dims or (("(" + dim.group(1) + ")") if dim else "")
                rank = d.count(":") if ":" in d else (-1 if d else 0)           # -1: fixed size
                info.setdefault(name.lower(), (kind, rank))
        return info
    files = {n: open(os.path.join(REF, n)).read() for n in ("MOD_MESH.F90", "MOD_PARTIT.F90", "MOD_TRACER.F90", "MOD_DYN.F90")}
    for fname, schemas in (("MOD_MESH.F90", [R.T_MESH]), ("MOD_PARTIT.F90", [R.T_COM_STRUCT, R.T_PARTIT_TAIL]),
                           ("MOD_TRACER.F90", [R.T_TRACER_DATA, R.T_TRACER_WORK]),
                           ("MOD_DYN.F90", [R.T_SOLVERINFO, R.T_DYN_WORK, R.T_DYN_HEAD, R.T_DYN_ARRAYS, R.T_DYN_FER, R.T_DYN_SE])):
        info = decls(files[fname][:files[fname].lower().index("\ncontains")] if fname == "MOD_MESH.F90" else files[fname])
        for schema in schemas:
            for kind, name in schema:
                key = name.split("%")[-1].lower()
                assert key in info, (fname, name)
                k, rank = info[key]
                if kind in ("i", "r", "l") or kind.startswith("c"):
                    assert (k, rank) == (kind, 0), (fname, name, k, rank)
                elif kind == "si":
                    assert (k, rank) == ("i", -1), (fname, name, k, rank)
                else:
                    assert (k, rank) == (kind[1], int(kind[2:])), (fname, name, k, rank)


def _case(g):
    st, trs, nb, dt = make_case(g, 2, "MFCT", "QR4C", "FCT", use_wsplit=True)
    return st, trs, nb, F.find_up_downwind_triangles(g)


@pytest.mark.parametrize("split", [False, True])
def test_dump_and_load_round_trip(tmp_path, pi_mesh, monkeypatch, split):
    g = pi_mesh
    st, trs, nb, tri = _case(g)
    if split:                                           # force gfortran's sub-record path on a small file
        monkeypatch.setattr(R, "_SUBREC", 100003)
    R.dump_dwarf(str(tmp_path), g, st, trs, nb, tri, wsplit_maxcfl=0.95)
    raw = open(tmp_path / "t_mesh.0", "rb").read()
    first = struct.unpack("<i", raw[:4])[0]
    assert (first < 0) == split                        # continued records announce themselves with a negative length
    m, st2, trs2, ex = R.load_dwarf(str(tmp_path))
    for name in ("elem2D_nodes", "edges", "edge_tri", "nlevels", "ulevels", "nlevels_nod2D", "ulevels_nod2D", "nod_in_elem2D",
                 "nod_in_elem2D_num", "elem_area", "elem_cos", "edge_dxdy", "edge_cross_dxdy", "gradient_sca", "area", "areasvol",
                 "nlevels_nod2D_min", "ulevels_nod2D_max"):
        assert np.array_equal(np.asarray(getattr(m, name)), np.asarray(getattr(g, name))), name
    assert (m.nl, m.N, m.T, m.E, m.eDim_nod2D) == (g.nl, g.N, g.T, g.E, g.eDim_nod2D)
    for k in ("uv", "w", "w_e", "w_i", "helem", "hnode", "hnode_new", "zbar_3d_n", "Z_3d_n", "zbar_n_bot"):
        assert torch.equal(getattr(st2, k), getattr(st, k)), k
    assert st2.use_wsplit is True and ex["wsplit_maxcfl"] == 0.95
    assert np.array_equal(ex["nboundary_lay"], nb) and np.array_equal(ex["edge_up_dn_tri"], tri)
    for a, b in zip(trs2, trs):
        assert torch.equal(a.values, b.values) and torch.equal(a.valuesAB, b.valuesAB)
        assert (a.tra_adv_hor, a.tra_adv_ver, a.tra_adv_lim, a.tra_adv_ph, a.tra_adv_pv) == ("MFCT", "QR4C", "FCT", 0.0, 1.0)
    # the reference keeps ONE tracers%work%edge_up_dn_grad: the file holds the last tracer's
    assert torch.equal(trs2[0].edge_up_dn_grad, trs[-1].edge_up_dn_grad)


def test_partitioned_rank_round_trip_and_file_names(tmp_path, pi_mesh):
    g = pi_mesh
    part = g.parts[8]
    st, trs, nb, tri = _case(g)
    loc = M.localize(g, part, 3)
    lst, ltr = F.scatter_to_local(g, loc, st, trs)
    R.dump_dwarf(str(tmp_path), loc, lst, ltr, nb[loc.myList_nod2D - 1])
    assert os.path.exists(tmp_path / "t_partit.3")      # mpirank_to_txt: padded to the width of npes (8 -> 1 digit)
    assert R.rank_suffix(3, 128) == "003" and R.rank_suffix(0, 1) == "0" and R.rank_suffix(7, 10) == "07"
    m, st2, trs2, ex = R.load_dwarf(str(tmp_path), 3, 8)
    c0, c1 = loc.com_nod2D, m.com_nod2D
    for a in ("rPE", "rptr", "rlist", "sPE", "sptr", "slist"):
        assert np.array_equal(getattr(c0, a), getattr(c1, a)), a
    assert np.array_equal(m.myList_nod2D, loc.myList_nod2D) and (m.mype, m.npes) == (3, 8)
    tp = ex["t_partit"]
    assert tp["com_nod2D%rPE"].size == 32 and tp["com_nod2D%rptr"].size == 33 and tp["com_nod2D%sptr"].size == 32


def test_framing_errors_are_reported(tmp_path):
    p = tmp_path / "t_mesh.0"
    p.write_bytes(struct.pack("<i", 8) + b"\0" * 8 + struct.pack("<i", 12))
    with pytest.raises(ValueError, match="record markers disagree"):
        R.read_t_mesh(str(p))
    p.write_bytes(struct.pack("<i", 4) + struct.pack("<i", 5) + struct.pack("<i", 4))
    with pytest.raises(ValueError, match="payload ends"):
        R.read_t_mesh(str(p))


@pytest.mark.gpu
def test_step_from_dwarf_files_equals_step_from_memory(tmp_path, pi_mesh):
    from common import run_cuda
    g = pi_mesh
    st, trs, nb, dt = make_case(g, 1, "MFCT", "QR4C", "FCT")
    R.dump_dwarf(str(tmp_path), g, st, trs, nb)
    m, st2, trs2, ex = R.load_dwarf(str(tmp_path))
    ctx, dh, dv = run_cuda(g, st, trs, nb, dt)
    ctx2, dh2, dv2 = run_cuda(m, st2, trs2, ex["nboundary_lay"], dt)
    assert np.array_equal(dh[0], dh2[0]) and np.array_equal(dv[0], dv2[0])
    ctx.close(); ctx2.close()


def test_cpp_reader_agrees_with_the_python_reader(tmp_path):
    """fesom2_b200/host/fesom_restart.cpp (read_all_bin_restarts of the compiled host side) on files written in the
    reference's format: every dimension, scheme string and array sum equals what the Python reader delivers"""
    import subprocess
    import numpy as np
    from fesom2_b200 import build as B, fields as F, mesh as M, restart as R
    g = M.synth_mesh(23, 19, nl=14, min_layers=4)
    st = F.make_state(g, "cpu", use_wsplit=True)
    trs = F.make_tracers(g, 3, "cpu", hor="MUSCL", ver="PPM", lim="FCT", ph=0.25, pv=0.75)
    nb = M.nboundary_lay(g)
    d = str(tmp_path / "np1")
    R.dump_dwarf(d, g, st, trs, nb, wsplit_maxcfl=0.9)
    B.build_library()
    p = subprocess.run([B.build_host(), "--dump-restart", d], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    out = {ln.split()[0] + (ln.split()[1] if ln.startswith("tracer") else ""): ln.split()[1:] for ln in p.stdout.strip().splitlines()}
    m2, st2, trs2, ex = R.load_dwarf(d)
    assert [int(x) for x in out["dims"]] == [g.nl, g.N, g.eDim_nod2D, g.T, g.eDim_elem2D, g.E, g.nod_in_elem2D.shape[1], 3, 1]
    ints = [m2.edges, m2.edge_tri, m2.elem2D_nodes, m2.nod_in_elem2D, m2.nod_in_elem2D_num, m2.nlevels, m2.ulevels, m2.nlevels_nod2D,
            m2.ulevels_nod2D, ex["nboundary_lay"]]
    assert [int(x) for x in out["ints"]] == [int(np.asarray(a, np.int64).sum()) for a in ints]
    reals = [m2.edge_cross_dxdy, m2.edge_dxdy, m2.elem_cos, m2.area, m2.areasvol, st2.helem, st2.hnode, st2.hnode_new, st2.zbar_3d_n,
             st2.Z_3d_n, st2.zbar_n_bot, st2.uv, st2.w, st2.w_e, st2.w_i]
    for got, a in zip(out["reals"], reals):
        a = np.asarray(a, np.float64).ravel()
        ref = 0.0
        for x in a:                                   # the C++ side sums left to right
            ref += x
        assert float(got) == ref
    for k, t in enumerate(trs2):
        f = out[f"tracer{k + 1}"]
        assert f[1:4] == [t.tra_adv_hor, t.tra_adv_ver, t.tra_adv_lim] and float(f[4]) == t.tra_adv_ph and float(f[5]) == t.tra_adv_pv
        assert abs(float(f[6]) - float(t.values.sum())) <= 1e-9 * abs(float(t.values.sum()))


def test_cpp_reader_rejects_a_truncated_file(tmp_path):
    import subprocess
    from fesom2_b200 import build as B, fields as F, mesh as M, restart as R
    g = M.synth_mesh(11, 9, nl=8, min_layers=4)
    st = F.make_state(g, "cpu")
    trs = F.make_tracers(g, 1, "cpu")
    d = str(tmp_path / "np1")
    R.dump_dwarf(d, g, st, trs, M.nboundary_lay(g))
    path = os.path.join(d, "t_mesh.0")
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[:len(raw) // 2])
    p = subprocess.run([B.build_host(), "--dump-restart", d], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "t_mesh.0" in p.stderr
