"""The Julia binding (fourierflows_jl_b200/julia/FourierFlowsB200.jl) cannot be executed in this image (no Julia toolchain);
these checks keep it honest statically: every `ccall` names a symbol the header declares and the library exports, with the
declared number of arguments, and the `struct`s that cross the boundary list the same fields, in the same order, as the C
header's ctypes mirror."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = open(os.path.join(ROOT, "fourierflows_jl_b200", "julia", "FourierFlowsB200.jl")).read()
HDR = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fourierflows_b200.h")).read(), flags=re.S)


def _ccalls():
    out = []
    for m in re.finditer(r"ccall\(\(:(ffb_[a-z0-9_]+),\s*lib\),\s*(\w+),\s*\(", JL):
        i, depth = m.end(), 1
        while depth:                      # the argument-type tuple, balanced parentheses
            depth += {"(": 1, ")": -1}.get(JL[i], 0)
            i += 1
        types = JL[m.end():i - 1].strip().rstrip(",")
        # split on top-level commas (Ptr{...} contains none, NTuple{3,Int32} would -- not used in signatures)
        tl = [] if not types else [t.strip() for t in re.split(r",(?![^{]*\})", types) if t.strip()]
        out.append((m.group(1), m.group(2), len(tl), tl))
    return out


def _header_args(name):
    m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", HDR, flags=re.S)
    args = m.group(1).strip()
    return [] if args in ("", "void") else [a.strip() for a in args.split(",")]


def _compatible(jl, c):
    if "*" in c:
        return jl.startswith("Ptr{") or jl == "Cstring"
    base = c.replace("const", "").split()
    ctype = " ".join(base[:-1]) if len(base) > 1 else base[0]
    return {"int": jl in ("Cint",), "int64_t": jl == "Int64", "double": jl == "Cdouble", "size_t": jl == "Csize_t"}.get(ctype, False)


def _header_arity(name):
    m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", HDR, flags=re.S)
    assert m, f"{name} is not declared in the header"
    args = m.group(1).strip()
    return 0 if args in ("", "void") else len(args.split(","))


def test_every_ccall_targets_a_declared_symbol_with_the_right_arity():
    import fourierflows_jl_b200 as ff
    lib = ff._lib.load()
    calls = _ccalls()
    assert len(calls) >= 35
    for name, ret, nargs, types in calls:
        assert hasattr(lib, name), f"{name} is not exported"
        assert _header_arity(name) == nargs, f"{name}: header takes {_header_arity(name)} arguments, the ccall passes {nargs}"
        assert ret in ("Cint", "Cstring"), (name, ret)
        for k, (jl, c) in enumerate(zip(types, _header_args(name))):
            assert _compatible(jl, c), f"{name} argument {k + 1}: Julia {jl} vs C `{c}`"


@pytest.mark.parametrize("jl_struct,c_struct", [("FFBDesc", "ffb_desc"), ("FFBCoef", "ffb_coef"), ("FFBFuse", "ffb_fuse"),
                                                 ("FFBProblemConfig", "ffb_problem_config")])
def test_julia_structs_mirror_the_c_structs(jl_struct, c_struct):
    import fourierflows_jl_b200 as ff
    body = re.search(r"struct " + jl_struct + r"\b(.*?)\bend\b", JL, flags=re.S).group(1)
    fields = re.findall(r"(\w+)\s*::", body)
    assert fields == [f for f, _ in getattr(ff._lib, c_struct)._fields_]
