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


# ------------------------------------------------------------------------------------------------------------------------
# Static walk of the path INTEGRATION.md documents: `TwoDGrid(B200(); ...)` -> `Problem(eqn, stepper, dt, grid)` ->
# `stepforward!`.  Every operation the REFERENCE performs on a device array along that path (cited by file:line under
# /root/reference/src) needs a B200 method in the wrapper, because generic broadcasting has no CPU fallback.
REQUIRED_METHODS = [
    # (what the reference does, file:line, regex that must match the wrapper)
    ("device_array(dev){T}(reshape(fftfreq(...)))  upload of a lazy host array", "domains.jl:77-78,193-195,333-336",
     r"B200Array\{T,N\}\(h::AbstractArray\{S,N\}\)"),
    ("device_array(dev){T,N}(undef, nx, ny)", "domains.jl:86-87,207-208,348-349", r"B200Array\{T,N\}\(u::UndefInitializer, dims::Vararg\{Integer,N\}\)"),
    ("OneDGrid(dev; ...) without generic broadcast (`@. 1 / k^2`, `invksq[1] = 0`)", "domains.jl:61-101", r"function FourierFlows\.OneDGrid\(dev::B200;"),
    ("TwoDGrid(dev; ...) (`@. k^2 + l^2`, `invKsq[1, 1] = 0`)", "domains.jl:175-223", r"function FourierFlows\.TwoDGrid\(dev::B200;"),
    ("ThreeDGrid(dev; ...)", "domains.jl:311-366", r"function FourierFlows\.ThreeDGrid\(dev::B200;"),
    ("CUDA.@allowscalar invKsq[1, 1] = 0", "domains.jl:199,203,340,344", r"function Base\.setindex!\(a::B200Array\{T\}, v, I::Integer\.\.\.\)"),
    ("zeros(dev, T, dims) / @devzeros", "utils.jl:79-94; problem.jl:108", r"function zeros\(::B200, ::Type\{T\}, dims\)"),
    ("supersize(L)", "problem.jl:30; utils.jl:57", r"supersize\(a::B200Array\) = size\(a\)"),
    ("ForwardEulerTimeStepper(N) = new(0N)", "timesteppers.jl:103", r"\*\(s::Number, a::B200Array\{T\}\) where T<:Real"),
    ("getexpLs: @. exp(dt * equation.L)", "timesteppers.jl:673-678", r"function getexpLs\(dt, eq::Equation\{TT,<:B200Array\{S\}\}\)"),
    ("getetdcoeffs(dt, L)", "timesteppers.jl:689-721", r"function getetdcoeffs\(dt, L::B200Array\{S\}"),
    ("makefilter(grid, T, sz) on the device", "domains.jl:545-546", r"function makefilter\(g::AbstractGrid\{Tg,<:B200Array\}, T, sz;"),
    ("dealias!(fh, grid)", "domains.jl:428-476", r"function dealias!\(fh::B200Array, g::AbstractGrid\{T,A,<:UnitRange\}\)"),
    ("@. N = 0", "diffusion.jl:131", r"copyto!\(dest::B200Array, bc::Broadcasted\{<:DefaultArrayStyle\{0\}\}\)"),
    ("@. L = -kappa * kr^2 ; @. cxh = im * kr * sol ; @. cx *= kappa ; @. ch = sol", "diffusion.jl:84,136,138,151-152",
     r"function copyto!\(dest::B200Array\{T\}, bc::Broadcasted\{B200Style\}\)"),
    ("broadcast style of device arrays", "—", r"BroadcastStyle\(::Type\{<:B200Array\}\) = B200Style\(\)"),
    ("deepcopy(vars.ch) before ldiv!", "diffusion.jl:154-155", r"Base\.deepcopy\(a::B200Array\) = copy\(a\)"),
    ("A(c) in set_c!", "diffusion.jl:169-171", r"B200Array\{T,N\}\(a::B200Array\{T,N\}\) where \{T,N\} = a"),
    ("Array(devarray) download", "output.jl:79", r"function Array\(a::B200Array\{T,N\}\)"),
    ("mul!(out, plan, in) / ldiv!(out, plan, in)", "diffusion.jl:137,139,154-155,171", r"mul!\(out::B200Array, p::B200Plan, a::B200Array\)"),
]


@pytest.mark.parametrize("what,where,pattern", REQUIRED_METHODS, ids=[r[0][:40] for r in REQUIRED_METHODS])
def test_wrapper_has_a_method_for_every_device_array_operation_of_the_reference(what, where, pattern):
    assert re.search(pattern, JL), f"no B200 method for `{what}` ({where})"


def test_grid_constructors_fill_every_field_of_the_reference_structs():
    """field lists of OneDGrid / TwoDGrid / ThreeDGrid (src/domains.jl:14-47, 111-160, 232-297): the B200 constructors call the
    default struct constructor positionally, so argument order and count must match"""
    fields = {
        "OneDGrid": "dev, nx, nk, nkr, dx, Lx, x, k, kr, invksq, invkrsq, fftplan, rfftplan, aliased_fraction, kalias, kralias",
        "TwoDGrid": "dev, nx, ny, nk, nl, nkr, dx, dy, Lx, Ly, x, y, k, l, kr, Ksq, invKsq, Krsq, invKrsq, fftplan, rfftplan, aliased_fraction, kalias, kralias, lalias",
        "ThreeDGrid": "dev, nx, ny, nz, nk, nl, nm, nkr, dx, dy, dz, Lx, Ly, Lz, x, y, z, k, l, m, kr, Ksq, invKsq, Krsq, invKrsq, fftplan, rfftplan, aliased_fraction, kalias, kralias, lalias, malias",
    }
    for name, want in fields.items():
        m = re.search(r"return " + name + r"\{T, typeof\(k\), typeof\(x\), typeof\(fftplan\), typeof\(rfftplan\), typeof\(kalias\), typeof\(dev\)\}\((.*?)\)\nend", JL, flags=re.S)
        assert m, f"{name}: B200 constructor does not end in the 7-parameter struct constructor"
        got = [a.strip() for a in m.group(1).replace("\n", " ").split(",")]
        assert got == [a.strip() for a in want.split(",")], (name, got)


def test_one_stepforward_method_per_stepper_type_and_no_ambiguous_union():
    """the reference defines `stepforward!(sol, clock, ts::XTimeStepper, ...)` for ten types (src/timesteppers.jl:111-667); a B200
    method with `ts::Union{...}` would be ambiguous with them (more specific in `sol`, less specific in `ts`)"""
    assert "ts::Union{" not in JL
    names = set()
    for m in re.finditer(r"for TS in \(([^)]*)\)\n  @eval function stepforward!\(sol::B200Sol\{T\}, clock, ts::FourierFlows\.\$TS,", JL):
        names |= {n.strip().lstrip(":") for n in m.group(1).split(",")}
    want = {f"{f}{s}TimeStepper" for f in ("", "Filtered") for s in ("ForwardEuler", "RK4", "LSRK54", "ETDRK4", "AB3")}
    assert names == want, names ^ want
