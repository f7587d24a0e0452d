"""CPU: static consistency of the Julia `ccall` shim (baorec.jl_b200/julia/BAOrecB200.jl) with the C ABI.
Julia is not in the image, so the shim cannot be executed; this test parses every `ccall` in it and holds its
return type and argument-type tuple against the prototype in include/baorec_b200.h (a `ccall` with the wrong
arity or an Int64 where the ABI takes an int corrupts the call silently), and the `Params` / `CosmologyParams`
struct layouts against the C structs."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SHIM = (ROOT / "baorec.jl_b200" / "julia" / "BAOrecB200.jl").read_text()
HEADER = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "baorec_b200.h").read_text(), flags=re.S)

# C parameter type -> the set of Julia ccall types that are ABI-compatible with it
POINTER = {"Ptr{Cvoid}", "Ptr{Cfloat}", "Ptr{Cdouble}", "Ptr{Params}", "Ref{Params}", "Ref{BaorecParams}", "Ptr{Ptr{Cvoid}}",
           "Ref{CosmologyParams}", "Ptr{CosmologyParams}", "Ptr{Cint}", "Ptr{Int64}", "Cstring", "Ptr{UInt8}"}
SCALAR = {"int": {"Cint"}, "int32_t": {"Cint", "Int32"}, "int64_t": {"Int64", "Clonglong"}, "float": {"Cfloat", "Float32"},
          "double": {"Cdouble", "Float64"}}


def c_prototypes():
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(baorec_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", HEADER):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or "[" in a or a.startswith("baorec_stream"):
                    params.append("ptr")
                else:
                    params.append(re.sub(r"\s+\w+$", "", re.sub(r"\bconst\b", "", a)).strip())
        protos[name] = (ret, params)
    return protos


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def shim_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(\s*(:[a-z0-9_]+|sym|\$\(QuoteNode\(sym\)\))\s*,\s*libbaorec\)\s*,\s*(\w+)\s*,\s*\(", SHIM):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(SHIM[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), split_top(SHIM[m.end():i - 1]), SHIM.count("\n", 0, m.start()) + 1))
    return calls


# two call sites name their symbol through a variable: the symbols they are instantiated with, read from the shim itself
GENERIC = {"$(QuoteNode(sym))": re.findall(r"\(:\w+!?, :(baorec_\w+)\)", SHIM),        # the `for (fn, sym) in (...)` @eval loop
           "sym": re.findall(r"_read\(:(baorec_\w+),", SHIM)}                            # read_shifts / reconstructed_positions


def test_every_ccall_matches_its_prototype():
    protos, calls = c_prototypes(), shim_ccalls()
    assert len(protos) == 60 and len(calls) >= 25
    seen = set()
    for sym, ret, types, line in calls:
        if sym.startswith(":"):
            names = [sym[1:]]
        else:
            names = GENERIC[sym]
            assert len(names) == 2, f"generic ccall at line {line}: expected two instantiations, found {names}"
        for name in names:
            assert name in protos, f"line {line}: {name} is not declared in the header"
            seen.add(name)
            cret, cparams = protos[name]
            if "char" in cret:
                assert ret == "Cstring", (name, ret)
            else:
                assert ret == "Cint" and cret == "int", (name, ret, cret)
            assert len(types) == len(cparams), f"line {line}: {name} takes {len(cparams)} arguments, the ccall passes {len(types)}"
            for k, (jt, ct) in enumerate(zip(types, cparams)):
                if ct == "ptr":
                    assert jt in POINTER, f"line {line}: {name} argument {k} is a pointer in C, {jt} in the shim"
                else:
                    assert jt in SCALAR[ct], f"line {line}: {name} argument {k} is {ct} in C, {jt} in the shim"
    # everything the reference's CuArray methods need is bound
    for must in ("baorec_create", "baorec_plan", "baorec_cic_scatter_f32", "baorec_gather_f32", "baorec_smooth_f32",
                 "baorec_setup_overdensity_f32", "baorec_iterate_f32", "baorec_reconstructed_overdensity_f32",
                 "baorec_reconstructed_potential_f32", "baorec_compute_displacements_f32", "baorec_read_shifts_f32",
                 "baorec_reconstructed_positions_f32", "baorec_setup_box_f32", "baorec_mg_jacobi_f32", "baorec_mg_residual_f32",
                 "baorec_mg_restrict_f32", "baorec_mg_prolong_f32", "baorec_mg_vcycle_f32", "baorec_mg_fmg_f32",
                 "baorec_cosmo_set", "baorec_sky_to_cartesian_f32", "baorec_cartesian_to_sky_f32", "baorec_fkp_weights_f32",
                 "baorec_wrap_positions_f32", "baorec_power_multipoles_f32"):
        assert must in seen, must


def c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, rest = decl.split(None, 1)
        for field in rest.split(","):
            m = re.match(r"\s*(\w+)(\[(\d+)\])?", field)
            out.append((m.group(1), ctype, int(m.group(3)) if m.group(3) else 1))
    return out


def julia_struct_fields(name):
    body = re.search(r"struct %s\b[^\n]*\n(.*?)\nend" % name, SHIM, flags=re.S).group(1)
    out = []
    for line in body.splitlines():
        for part in line.split(";"):
            m = re.match(r"\s*(\w+)::([\w{},]+)", part)
            if m:
                out.append((m.group(1), m.group(2)))
    return out


JTYPE = {"float": "Cfloat", "int32_t": "Int32", "double": "Cdouble", "int64_t": "Int64"}


@pytest.mark.parametrize("cname,jname", [("baorec_params", "Params"), ("baorec_cosmology", "CosmologyParams")])
def test_struct_layouts_match(cname, jname):
    cf, jf = c_struct_fields(cname), julia_struct_fields(jname)
    assert [f[0] for f in cf] == [f[0] for f in jf]
    for (_, ctype, count), (_, jt) in zip(cf, jf):
        want = JTYPE[ctype] if count == 1 else "NTuple{%d,%s}" % (count, JTYPE[ctype])
        assert jt == want, (cname, ctype, count, jt)
