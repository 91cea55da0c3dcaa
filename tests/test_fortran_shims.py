"""The ISO_C_BINDING shims under fortran/ cannot be compiled in this image (no Fortran compiler), so their interface
blocks are checked statically against include/rbc3d.h: every bind(C) name is a declared entry point, the argument counts
agree, and an argument is passed by VALUE in Fortran exactly where the C prototype takes a scalar (type(c_ptr) handles by value
where it takes a pointer)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_prototypes():
    txt = open(os.path.join(ROOT, "include", "rbc3d.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(rbc3d_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        args = [] if args in ([""], ["void"]) else args
        out[m.group(1)] = ["*" in a or "[" in a for a in args]            # True = pointer / array parameter
    return out


def fortran_interfaces():
    out = {}
    for fn in sorted(glob.glob(os.path.join(ROOT, "fortran", "*.F90"))):
        src = open(fn).read()
        src = re.sub(r"&\s*\n\s*&?", " ", src)                           # join continuation lines
        for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)\s*result\((\w+)\)(.*?)end function",
                             src, flags=re.S | re.I):
            args = [a.strip().lower() for a in m.group(2).split(",") if a.strip()]
            body = m.group(5)
            byval, cptr = set(), set()
            for line in body.splitlines():
                line = line.split("!")[0]
                if "::" not in line:
                    continue
                decl, names = line.split("::")[0], line.split("::")[1]
                names = [re.sub(r"\(.*", "", n).strip().lower() for n in re.split(r",(?![^()]*\))", names)]
                if re.search(r",\s*value\b", decl, flags=re.I):
                    byval.update(names)
                if re.search(r"type\(c_ptr\)", decl, flags=re.I):
                    cptr.update(names)
            out[m.group(1).lower()] = (os.path.basename(fn), args, byval, cptr, m.group(3))
    return out


def test_interfaces_match_the_header():
    protos, ifaces = c_prototypes(), fortran_interfaces()
    assert len(protos) >= 55 and len(ifaces) >= 30
    for fname, (fn, args, byval, cptr, cname) in ifaces.items():
        assert cname in protos, f"{fn}: {cname} is not declared in include/rbc3d.h"
        ptr = protos[cname]
        assert len(args) == len(ptr), f"{fn}: {cname} has {len(args)} arguments, the header {len(ptr)}"
        for a, is_ptr in zip(args, ptr):
            # a C pointer is either a Fortran array / scalar passed by reference, or a type(c_ptr) passed by value
            # (an opaque handle or a host address); a C scalar is a Fortran scalar with the VALUE attribute
            f_ptr = (a not in byval) or (a in cptr)
            assert f_ptr == is_ptr, f"{fn}: {cname}({a}): by value = {a in byval}, c_ptr = {a in cptr}, C pointer = {is_ptr}"
    # the operator entry points the replaced modules forward to must all be bound
    for need in ("rbc3d_add_int_on_rbcs", "rbc3d_add_int_on_walls", "rbc3d_pme_distrib_source", "rbc3d_pme_transform",
                 "rbc3d_pme_add_interp_vel", "rbc3d_cells_set_geometry", "rbc3d_cells_set_density", "rbc3d_walls_set",
                 "rbc3d_wall_prepare_sing", "rbc3d_sing_int_on_wall", "rbc3d_ctx_create", "rbc3d_ctx_destroy"):
        assert need in {v[4] for v in ifaces.values()}, need


def test_ctypes_signatures_match_the_header():
    """rbc3d_b200/capi.py SIGNATURES (the ctypes stub INTEGRATION.md points other host languages at): same set of
    entry points as the header, same argument counts, pointer types exactly where the prototype has a pointer, double
    vs integer scalars as declared."""
    import ctypes as C
    from rbc3d_b200 import capi
    txt = open(os.path.join(ROOT, "include", "rbc3d.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(rbc3d_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if args in ([""], ["void"]) else args
    assert set(protos) == set(capi.SIGNATURES)
    for name, (_, argtypes) in capi.SIGNATURES.items():
        cargs = protos[name]
        assert len(argtypes) == len(cargs), name
        for t, a in zip(argtypes, cargs):
            is_ptr = "*" in a or "[" in a
            t_ptr = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer)
            assert t_ptr == is_ptr, (name, a, t)
            if not is_ptr:
                want = C.c_double if a.startswith("double") else (C.c_size_t if a.startswith("size_t") else C.c_int)
                assert t is want, (name, a, t)


def test_shim_calls_have_the_interface_arity():
    """every call of an rbc3d_* function in the shim procedures passes as many arguments as its interface declares"""
    ifaces = fortran_interfaces()
    ncalls = 0
    for fn in sorted(glob.glob(os.path.join(ROOT, "fortran", "*.F90"))):
        src = open(fn).read()
        src = re.sub(r"&\s*\n\s*&?", " ", src)
        src = "\n".join(line.split("!")[0] if "bind(C" not in line else "" for line in src.splitlines())
        for m in re.finditer(r"=\s*(rbc3d_\w+)\s*\(", src):
            name = m.group(1).lower()
            depth, i, nargs, any_char = 1, m.end(), 1, False
            while depth:
                ch = src[i]
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "," and depth == 1:
                    nargs += 1
                elif not ch.isspace():
                    any_char = True
                i += 1
            nargs = nargs if any_char else 0
            assert name in ifaces, (os.path.basename(fn), name)
            assert nargs == len(ifaces[name][1]), (os.path.basename(fn), name, nargs, len(ifaces[name][1]))
            ncalls += 1
    assert ncalls >= 25


def _procedures():
    """name -> body (comments stripped, continuation lines joined, lower case) of every shim procedure"""
    out = {}
    for fn in sorted(glob.glob(os.path.join(ROOT, "fortran", "*.F90"))):
        src = open(fn).read()
        src = re.sub(r"&\s*\n\s*&?", " ", src)
        src = "\n".join(line.split("!")[0] for line in src.splitlines()).lower()
        src = re.sub(r"interface.*?end interface", "", src, flags=re.S)
        for m in re.finditer(r"^\s*(?:subroutine|(?:[\w()]+\s+)?function)\s+(\w+).*?^\s*end (?:subroutine|function)", src,
                             flags=re.S | re.M):
            out[m.group(1)] = m.group(0)
    return out


def test_every_shim_creates_the_context_before_using_it():
    """The reference's TimeInt_Init calls PrepareSingIntOnWall (ModTimeInt.F90:87-91) before PME_Init (:96), and the
    post-processing programs use the lists without either: context creation is lazy (B200_EnsureInit, guarded by
    c_associated) and every procedure reaches it before the first use of b200_ctx."""
    procs = _procedures()
    ensure = procs["b200_ensureinit"]
    assert "c_associated(b200_ctx)" in ensure.split("rbc3d_ctx_create")[0]
    assert ensure.index("rbc3d_ctx_attach_comm") < ensure.index("rbc3d_cells_set_mesh")
    assert "b200_check(ierr, 'rbc3d_comm_unique_id')" in ensure
    creators = ("b200_ensureinit", "b200_init", "b200_syncwalls", "b200_synccells", "b200_syncdensity",
                "b200_syncwalltraction")
    for c in creators[1:]:
        first = procs[c].index("b200_ctx") if "b200_ctx" in procs[c] else len(procs[c])
        assert "call b200_ensureinit" in procs[c][:first] or "call b200_ensureinit" in procs[c], c
    used = 0
    for name, body in procs.items():
        if name in ("b200_ensureinit", "pme_finalize") or "b200_ctx" not in body:
            continue
        used += 1
        first = body.index("b200_ctx")
        assert any("call " + c in body[:first] for c in creators), f"{name} uses b200_ctx before creating it"
    assert used >= 15
    # replay of TimeInt_Init (walls present): PrepareSingIntOnWall first, then PME_Init, which must not create a second context
    prep = procs["preparesingintonwall"]
    assert prep.index("call b200_syncwalls") < prep.index("rbc3d_wall_prepare_sing")
    assert "call b200_init" in procs["pme_init"] and "call b200_ensureinit" in procs["b200_init"]
    # Fourier-only ranks and ModPostProcess: tractions are sent by PME_Distrib_Source itself
    ds = procs["pme_distrib_source"]
    assert ds.index("call b200_syncwalltraction") < ds.index("rbc3d_pme_distrib_source")


def test_noslip_shim_hands_the_whole_solve_to_the_library():
    """fortran/ModNoSlip_b200.F90 keeps ModNoSlip's public procedures (ModNoSlip.F90:32-34); NoSlipWall packs wall%f and
    indxVertGlb of all walls back to back, makes ONE library call with the reference's GMRES settings (rtol = eps_Ewd,
    60 iterations, ModNoSlip.F90:83-87) and writes wall%f back; Compute_Wall_Residual_Vel is operator #3 + vBkg."""
    src = open(os.path.join(ROOT, "fortran", "ModNoSlip_b200.F90")).read().lower()
    assert re.search(r"public\s*::\s*noslipwall,\s*compute_wall_residual_vel,\s*wallbuildmat", src)
    procs = _procedures()
    ns = procs["noslipwall"]
    assert ns.count("rbc3d_noslip_solve") == 1 and "rbc3d_apply" not in ns and "rbc3d_add_int" not in ns
    assert re.search(r"rbc3d_noslip_solve\(b200_ctx, indx, nindep, vbkg, .*?, eps_ewd, maxit, f, niter, history, slip\)", ns)
    assert "maxit = 60" in ns and "wall%indxvertglb" in ns
    assert ns.index("= wall%f") < ns.index("rbc3d_noslip_solve") < ns.index("wall%f = f(")
    rv = procs["compute_wall_residual_vel"]
    assert rv.index("call b200_syncwalltraction") < rv.index("rbc3d_apply(b200_ctx, c1, c1,")
    assert "tl_walls" in rv and "rbc3d_collect_array" in rv and "vbkg(ii)" in rv
