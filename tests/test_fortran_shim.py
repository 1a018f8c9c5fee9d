"""The ISO_C_BINDING shim (fesom2_b200/fortran/oce_adv_tra_b200.F90) against the C header, field by field.

No Fortran compiler exists in this image or on the GPU box (gpurun_out/r4a_probe.log), so the shim cannot be
compiled here.  What CAN be checked mechanically is the part a compiler would not check either -- that every
bind(C) derived type has the header struct's fields in the same order with interoperable kinds, and that every
interface block passes the arguments the C prototype expects (by value / by reference, in order)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "fesom_adv_b200.h")
F90 = os.path.join(ROOT, "fesom2_b200", "fortran", "oce_adv_tra_b200.F90")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def c_structs():
    src = _strip_c_comments(open(HDR).read())
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(int32_t|double|char)\s+(.*)$", decl, flags=re.S)
            assert m, decl
            base = m.group(2)
            for item in m.group(3).split(","):
                item = item.strip()
                ptr = item.startswith("*")
                nm = item.lstrip("* ").strip()
                kind = "ptr" if ptr else {"int32_t": "i32", "double": "f64"}[base]
                fields.append((nm.lower(), kind))
        out[name] = fields
    return out


def f_types():
    src = open(F90).read()
    out = {}
    for name, body in re.findall(r"type,\s*bind\(C\)\s*::\s*(\w+)(.*?)end type", src, flags=re.S | re.I):
        fields = []
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if not line:
                continue
            m = re.match(r"(integer\(c_int32_t\)|type\(c_ptr\)|real\(c_double\))\s*::\s*(.*)$", line, flags=re.I)
            assert m, line
            kind = {"integer(c_int32_t)": "i32", "type(c_ptr)": "ptr", "real(c_double)": "f64"}[m.group(1).lower()]
            fields += [(x.strip().lower(), kind) for x in m.group(2).split(",")]
        out[name] = fields
    return out


def test_bind_c_types_match_the_header_structs():
    cs, fs = c_structs(), f_types()
    assert set(fs) == {"adv_mesh_desc_t", "adv_state_desc_t", "adv_tracer_desc_t", "adv_gradient_mesh_desc_t", "adv_zstar_desc_t", "adv_zlevel_desc_t"}
    for name, ff in fs.items():
        assert name in cs, name
        assert ff == cs[name], (name, [x for x in zip(ff, cs[name]) if x[0] != x[1]][:3], len(ff), len(cs[name]))


def c_prototypes():
    src = _strip_c_comments(open(HDR).read())
    src = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"\b(int|const char \*|int64_t|void \*)\s*(adv_\w+)\s*\(([^)]*)\)\s*;", src):
        kinds = []
        for a in args.split(","):
            a = " ".join(a.split())
            if a in ("void", ""):
                continue
            if re.match(r"adv_ctx_t \*\*\w+$", a):
                kinds.append("ptr&")                                  # pointer returned through the argument
            elif re.match(r"(const )?adv_ctx_t \*\w+$", a) or re.match(r"void \*\w+$", a) or re.match(r"(const )?double \*\w+$", a):
                kinds.append("ptr")
            elif re.match(r"const (adv_\w+_desc_t) \*\w+$", a):
                kinds.append("ref:" + re.match(r"const (adv_\w+_desc_t)", a).group(1))
            elif re.match(r"(const )?char \w+\[128\]$", a):
                kinds.append("chars")
            elif re.match(r"(const )?char \*\w+$", a):
                kinds.append("cstr")
            elif re.match(r"int \w+$", a):
                kinds.append("int")
            elif re.match(r"int64_t \w+$", a) or re.match(r"uint64_t \w+$", a):
                kinds.append("i64")
            elif re.match(r"double \w+$", a):
                kinds.append("f64")
            elif re.match(r"(const )?double \*const \*\w+$", a) or re.match(r"adv_ctx_t \*const \*\w+$", a):
                kinds.append("ptr[]")
            else:
                kinds.append("other:" + a)
        out[name] = kinds
    return out


def f_interfaces():
    src = open(F90).read()
    blk = re.search(r"\n\s*interface\n(.*?)\n\s*end interface", src, flags=re.S | re.I).group(1)
    blk = re.sub(r"&\s*\n\s*", " ", blk)                              # continuation lines
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)(.*?)end function", blk, flags=re.S | re.I):
        fname, arglist, cname, body = m.groups()
        assert fname == cname
        args = [a.strip().lower() for a in arglist.split(",") if a.strip()]
        kind = {}
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if not line or line.lower().startswith("import"):
                continue
            d = re.match(r"(.*?)::\s*(.*)$", line)
            assert d, line
            spec = d.group(1).lower().replace(" ", "")
            for item in re.findall(r"(\w+)(\([^)]*\))?", d.group(2)):
                nm, dims = item[0].lower(), item[1]
                if spec.startswith("type(c_ptr)"):
                    k = "ptr" if ",value" in spec else ("ptr[]" if dims == "(*)" else "ptr&")
                elif spec.startswith("type(adv_"):
                    k = "ref:" + re.match(r"type\((adv_\w+)\)", spec).group(1)
                    assert ",value" not in spec
                elif spec.startswith("integer(c_int)"):
                    k = "int"; assert ",value" in spec
                elif spec.startswith("integer(c_int64_t)"):
                    k = "i64"; assert ",value" in spec
                elif spec.startswith("real(c_double)"):
                    k = "f64"; assert ",value" in spec
                elif spec.startswith("character(kind=c_char)"):
                    k = "chars"; assert dims == "(128)"
                else:
                    raise AssertionError(line)
                kind[nm] = k
        out[cname] = [kind[a] for a in args]
    return out


def test_interface_blocks_match_the_c_prototypes():
    cp, fi = c_prototypes(), f_interfaces()
    must = {"adv_ctx_create", "adv_ctx_destroy", "adv_last_error", "adv_comm_unique_id", "adv_ctx_comm_init", "adv_ctx_set_state",
            "adv_ctx_set_state_step", "adv_do_oce_adv_tra", "adv_ctx_set_gradient_mesh", "adv_tracer_gradient_elements",
            "adv_fill_up_dn_grad", "adv_exchange_elem", "adv_init_tracers_AB", "adv_update_values", "adv_exchange_nod",
            "adv_ctx_wait_for", "adv_ctx_signal", "adv_ctx_synchronize", "adv_vert_vel_ale", "adv_vert_vel_ale_zstar"}
    assert must <= set(fi), must - set(fi)
    for name, fk in fi.items():
        assert name in cp, name
        ck = ["ref:" + k[4:] if k.startswith("ref:") else k for k in cp[name]]
        assert fk == ck, (name, fk, ck)


def test_wrapper_keeps_the_reference_signature_and_guards_the_batch():
    src = open(F90).read()
    # the reference's external procedure, argument for argument (src/oce_adv_tra_driver.F90:46)
    assert re.search(r"^subroutine do_oce_adv_tra\(dt, vel, w, wi, we, tr_num, dynamics, tracers, partit, mesh\)", src, flags=re.M)
    # a batch of several tracers must not alias the single tracers%work set
    assert "n > 1 .and. .not. (present(grad_b) .and. present(dh_b) .and. present(dv_b))" in src
    assert src.count("tracers%work%edge_up_dn_grad") == 1 and "grad_b(:,:,:,k)" in src
    # device addresses are taken with acc_deviceptr (scope-independent), never via HOST_DATA in a caller
    assert "HOST_DATA" not in src.upper().replace("`!$ACC HOST_DATA USE_DEVICE`", "")
    assert "#define ADV_ADDR(x) acc_deviceptr(x)" in src and "#define ADV_ADDR(x) c_loc(x)" in src
    # optional: gradients computed by the library (NULL pointer in the descriptor), element halo handed over at init
    assert "if (adv_b200_device_gradients) td(k)%edge_up_dn_grad = c_null_ptr" in src
    assert "gd%n_elem = partit%myDim_elem2D + partit%eDim_elem2D + partit%eXDim_elem2D" in src and "partit%com_elem2D_full%slist" in src
    # state handed over once per model step
    assert "adv_ctx_set_state_step(adv_b200_ctx, st, ADV_WHERE, int(mstep, c_int64_t))" in src
