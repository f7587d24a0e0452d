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
           "Ref{CosmologyParams}", "Ptr{CosmologyParams}", "Ptr{Cint}", "Ptr{Int64}", "Cstring", "Ptr{UInt8}", "Ptr{Cstring}"}
SCALAR = {"int": {"Cint"}, "int32_t": {"Cint", "Int32"}, "int64_t": {"Int64", "Clonglong"}, "float": {"Cfloat", "Float32"},
          "double": {"Cdouble", "Float64"}, "char": {"Cchar"}}


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


CCALL = re.compile(r"ccall\(\(\s*([^,]+?)\s*,\s*libbaorec\)\s*,\s*(\w+)\s*,\s*\(")


def shim_ccalls():
    """(symbol expression, return type, argument types, argument expressions, line) of every ccall into the library"""
    calls = []
    for m in CCALL.finditer(SHIM):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(SHIM[i], 0)
            i += 1
        types = split_top(SHIM[m.end():i - 1])
        j, depth = i, 1                         # the rest of the ccall( ... ): the argument expressions
        while depth:
            depth += {"(": 1, ")": -1}.get(SHIM[j], 0)
            j += 1
        args = split_top(SHIM[i:j - 1].lstrip(", \n"))
        calls.append((m.group(1), m.group(2), types, args, SHIM.count("\n", 0, m.start()) + 1))
    return calls


def eval_loops():
    """[(start, end, {loop variable: [values]})] of the `for (...) in (...) @eval function ... end end` blocks"""
    out = []
    for m in re.finditer(r"^for (\(?[\w, ]+\)?) in \((.*?)\)\n\s+@eval", SHIM, flags=re.S | re.M):
        names = [n.strip() for n in m.group(1).strip("()").split(",")]
        rows = re.findall(r"\(([^()]*)\)", m.group(2)) or [v for v in split_top(m.group(2))]
        values = {n: [] for n in names}
        for row in rows:
            parts = [q.strip() for q in row.split(",")]
            for n, q in zip(names, parts):
                values[n].append(q)
        end = SHIM.index("\nend\n", m.end()) + 5
        out.append((m.start(), end, values))
    return out


def test_ccall_symbols_are_literals():
    """Julia lowers ccall((name, lib), ...) only when the tuple is a constant expression: `:literal`, or a symbol
    interpolated into an @eval'd definition.  A function argument or any other local is a syntax error
    ("ccall function name and library expression cannot reference local variables") that fails `include`."""
    loops = eval_loops()
    assert len(loops) >= 3
    for m in CCALL.finditer(SHIM):
        sym, line = m.group(1), SHIM.count("\n", 0, m.start()) + 1
        if re.fullmatch(r":[a-z0-9_]+", sym):
            continue
        im = re.fullmatch(r"\$\(QuoteNode\((\w+)\)\)", sym)
        assert im, f"line {line}: ccall symbol {sym!r} is neither a literal nor an @eval interpolation"
        inside = [v for a, b, v in loops if a < m.start() < b]
        assert inside and im.group(1) in inside[0], f"line {line}: {sym} is not a loop variable of an enclosing @eval loop"
    assert re.search(r"^const libbaorec\b", SHIM, flags=re.M), "the library name must be a const global"


def test_no_splat_and_matching_arity_in_ccalls():
    """ccall checks the argument count against the type tuple when the call is lowered; a splatted argument cannot
    be counted.  Every call passes exactly as many expressions as it declares types."""
    for sym, ret, types, args, line in shim_ccalls():
        assert not any(a.endswith("...") for a in args), f"line {line}: splatted ccall argument"
        assert len(args) == len(types), f"line {line}: {len(types)} argument types, {len(args)} arguments"


def test_overrides_are_at_least_as_typed_as_the_reference():
    """An override with an untyped argument where the reference's CuArray method types it (k⃗ / x_vec tuples,
    SVector boxes, the recon type) is ambiguous with that method instead of more specific."""
    for name, must in (("iterate!", ["k⃗::Vec3", "β::Float32"]),
                       ("jacobi!", ["x_vec::Vec3", "box_size::V3", "box_min::V3"]),
                       ("residual!", ["x_vec::Vec3", "box_size::V3", "box_min::V3"]),
                       ("vcycle!", ["box_size::V3", "box_min::V3"]),
                       ("fmg", ["box_size::V3", "box_min::V3"]),
                       ("smooth!", ["box_size::SVector{3,Float32}"]),
                       ("cic!", ["box_size::SVector{3,Float32}", "box_min::SVector{3,Float32}"])):
        mine = [q for q in re.findall(r"function %s\((.*?)\)\n" % re.escape(name), SHIM, flags=re.S)]
        assert mine, name
        for q in mine:
            for need in must:
                assert need in q, (name, need)
    # one compute_displacements per recon type; no Vararg reconstructed_* methods
    assert "compute_displacements(mesh::CuArray{Float32,3}, x::CuVector{Float32}, y::CuVector{Float32},\n" in SHIM
    assert "recon::$R)" in SHIM and "recon::AbstractRecon)\n    ctx = plan!(mesh" not in SHIM
    assert "CuArray{Float32}...)" not in SHIM
    # out-parameters / 3-vectors travel as Vector{Float32} (a Ref of a tuple does not convert to Ptr{Cfloat})
    assert "Ref((" not in SHIM


# loop-generated call sites: the symbols they are instantiated with, read from the loop headers
def generic_names(pos, var):
    for a, b, values in eval_loops():
        if a < pos < b:
            return [v.lstrip(":") for v in values[var]]
    raise AssertionError("no enclosing @eval loop")


def test_every_ccall_matches_its_prototype():
    protos, calls = c_prototypes(), shim_ccalls()
    assert len(protos) == 75 and len(calls) >= 32
    seen = set()
    starts = [m.start() for m in CCALL.finditer(SHIM)]
    for (sym, ret, types, args, line), pos in zip(calls, starts):
        if sym.startswith(":"):
            names = [sym[1:]]
        else:
            names = generic_names(pos, re.fullmatch(r"\$\(QuoteNode\((\w+)\)\)", sym).group(1))
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
                 "baorec_wrap_positions_f32", "baorec_power_multipoles_f32", "baorec_batch_host_f32", "baorec_batch_files_f32",
                 "baorec_text_catalog_read_f32", "baorec_npy_read_columns_f32", "baorec_npy_write_columns_f32",
                 "baorec_catalog_select_f32"):
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
