# FourierFlowsB200.jl -- thin Julia binding of libfourierflows_b200.so for FourierFlows.jl v0.10.7.
#
# NOT EXECUTABLE IN THE BUILD IMAGE (no Julia toolchain): this file is the reference-side binding a maintainer
# would add; every `ccall` targets a symbol declared in include/fourierflows_b200.h (checked by tests/test_julia_wrapper.py,
# which also walks every operation the reference's grid / stepper / Problem constructors perform on a device array and
# requires a B200 method for it).  It hooks the three seams of SURVEY.md section 8b without touching user code:
#   B1  `zeros(::B200, T, dims)` / `device_array(::B200)` / upload / download / scalar writes   (src/utils.jl:79-80, 329-332)
#       `OneDGrid / TwoDGrid / ThreeDGrid(::B200; ...)` built on ffb_wavenumbers / ffb_ksq        (src/domains.jl:61-101,175-223,311-366)
#   B2  `plan_flows_rfft` / `plan_flows_fft`, `mul!`, `ldiv!`                                     (src/domains.jl:2-5; AbstractFFTs protocol)
#   B3  one `stepforward!` method per stepper type on B200 arrays                                (src/timesteppers.jl:111-667)
# `dealias!`, `makefilter`, `getexpLs`, `getetdcoeffs`, `parsevalsum(2)` get B200 methods as well.  Broadcasts on B200Arrays
# are lowered to a CLOSED vocabulary of library kernels (linear combinations, products / squares of same-shape arrays,
# `c * kx^a * l^b * m^c * w * field`, zero fill); any other broadcast raises -- there is no CPU fallback.  The fused
# transforms, the C-driven problem and the multi-GPU plan are bound at the end of the file.  No CUDA.jl, no KernelAbstractions.
module FourierFlowsB200

using FourierFlows
using FourierFlows: AbstractGrid, OneDGrid, TwoDGrid, ThreeDGrid, Device, Equation, fltype, cxtype, getaliasedwavenumbers
import FourierFlows: device_array, plan_flows_fft, plan_flows_rfft, dealias!, makefilter, getetdcoeffs, getexpLs,
                     stepforward!, supersize
import LinearAlgebra: mul!, ldiv!
import Base: size, zeros, copyto!, Array, \, *, +, -, /
using Base.Broadcast: Broadcasted, BroadcastStyle, AbstractArrayStyle, DefaultArrayStyle

const lib = get(ENV, "FFB200_LIB", "libfourierflows_b200.so")

struct B200 <: Device end          # `TwoDGrid(B200(); nx, Lx)` selects this backend

# ---------------------------------------------------------------- status handling
function check(rc::Cint)
  rc == 0 && return nothing
  msg = unsafe_string(ccall((:ffb_last_error, lib), Cstring, ()))
  rc == -2 && throw(DomainError(msg))      # FFB_EDOMAIN  <-> src/domains.jl:66,179,316
  rc == -3 && throw(OutOfMemoryError())
  error("libfourierflows_b200 ($rc): $msg")
end

ffbtype(::Type{Float32}) = Cint(0); ffbtype(::Type{Float64}) = Cint(1)
ffbtype(::Type{Complex{T}}) where T = ffbtype(T)

# ---------------------------------------------------------------- B1: device arrays
mutable struct B200Array{T,N} <: AbstractArray{T,N}
  ptr  :: Ptr{Cvoid}
  dims :: NTuple{N,Int}
  function B200Array{T,N}(::UndefInitializer, dims::NTuple{N,Int}) where {T,N}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ffb_malloc, lib), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), p, max(prod(dims), 1) * sizeof(T)))
    a = new{T,N}(p[], dims)
    finalizer(x -> ccall((:ffb_free, lib), Cint, (Ptr{Cvoid},), x.ptr), a)   # ffb_free is thread-safe
    return a
  end
end
# `device_array(dev){T,N}(undef, nx, ny)` (src/domains.jl:86-87,207-208,348-349) and `device_array(dev){T}(undef, dims...)`
B200Array{T,N}(u::UndefInitializer, dims::Vararg{Integer,N}) where {T,N} = B200Array{T,N}(u, map(Int, dims))
B200Array{T}(u::UndefInitializer, dims::Integer...) where T = B200Array{T,length(dims)}(u, map(Int, dims))
B200Array{T}(u::UndefInitializer, dims::NTuple{N,Integer}) where {T,N} = B200Array{T,N}(u, map(Int, dims))
size(a::B200Array) = a.dims
supersize(a::B200Array) = size(a)                                              # src/utils.jl:57
Base.IndexStyle(::Type{<:B200Array}) = IndexLinear()
Base.similar(a::B200Array{T,N}) where {T,N} = B200Array{T,N}(undef, size(a))
Base.similar(a::B200Array, ::Type{S}, dims::Dims{M}) where {S,M} = B200Array{S,M}(undef, dims)
Base.unsafe_convert(::Type{Ptr{T}}, a::B200Array{T}) where T = Ptr{T}(a.ptr)
Base.show(io::IO, a::B200Array{T,N}) where {T,N} = print(io, join(size(a), "x"), " B200Array{", T, ",", N, "} (device memory)")
Base.show(io::IO, ::MIME"text/plain", a::B200Array) = show(io, a)

# upload: `device_array(dev){T}(host)` (src/domains.jl:77-78,193-195,333-336 pass `reshape(fftfreq(...), (nk, 1))`, a lazy
# AbstractArray of Float64: it is collected and converted to T on the host, like `CuArray{T}(x)` does)
function B200Array{T,N}(h::AbstractArray{S,N}) where {T,S,N}
  hc = convert(Array{T,N}, collect(h))
  a = B200Array{T,N}(undef, size(hc))
  GC.@preserve hc check(ccall((:ffb_h2d, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), a.ptr, pointer(hc), sizeof(hc)))
  check(ccall((:ffb_sync, lib), Cint, ()))
  return a
end
B200Array{T}(h::AbstractArray{S,N}) where {T,S,N} = B200Array{T,N}(h)
B200Array(h::AbstractArray{T,N}) where {T,N} = B200Array{T,N}(h)
B200Array{T,N}(a::B200Array{T,N}) where {T,N} = a                               # `A(c)` in set_c! (src/diffusion.jl:169-171)
function Array(a::B200Array{T,N}) where {T,N}                                   # download: src/output.jl:79
  h = Array{T,N}(undef, size(a))
  GC.@preserve h check(ccall((:ffb_d2h, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), pointer(h), a.ptr, sizeof(h)))
  return h
end
Base.collect(a::B200Array) = Array(a)
function zeros(::B200, ::Type{T}, dims) where T                                 # src/utils.jl:80
  d = dims isa Integer ? (Int(dims),) : map(Int, Tuple(dims))
  a = B200Array{T,length(d)}(undef, d)
  check(ccall((:ffb_memset_zero, lib), Cint, (Ptr{Cvoid}, Csize_t), a.ptr, prod(d) * sizeof(T)))
  return a
end
device_array(::B200) = B200Array
device_array(::B200, T, dim) = B200Array{T,dim}
copyto!(d::B200Array{T}, s::B200Array{T}) where T =
  (check(ccall((:ffb_d2d, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, s.ptr, prod(size(d)) * sizeof(T))); d)
copyto!(d::B200Array{T,N}, s::Array{T,N}) where {T,N} =
  (GC.@preserve s check(ccall((:ffb_h2d, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), d.ptr, pointer(s), sizeof(s))); check(ccall((:ffb_sync, lib), Cint, ())); d)
Base.copy(a::B200Array) = copyto!(similar(a), a)
Base.deepcopy(a::B200Array) = copy(a)                                           # updatevars! (src/diffusion.jl:154-155)
function Base.fill!(a::B200Array{T}, v) where T
  iszero(v) || error("fill!(::B200Array, v) is implemented for v = 0 only (no CPU fallback)")
  check(ccall((:ffb_memset_zero, lib), Cint, (Ptr{Cvoid}, Csize_t), a.ptr, length(a) * sizeof(T)))
  return a
end
# scalar access moves ONE element over PCIe: what `CUDA.@allowscalar invKsq[1, 1] = 0` does in the grid constructors
# (src/domains.jl:82-83,199,203,340,344) and what tests / `show` of a single value need; never used on a hot path
function Base.setindex!(a::B200Array{T}, v, I::Integer...) where T
  i = LinearIndices(size(a))[I...]
  r = Ref{T}(convert(T, v))
  check(ccall((:ffb_h2d, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), a.ptr + (i - 1) * sizeof(T), r, sizeof(T)))
  check(ccall((:ffb_sync, lib), Cint, ()))
  return a
end
function Base.getindex(a::B200Array{T}, I::Integer...) where T
  i = LinearIndices(size(a))[I...]
  r = Ref{T}()
  check(ccall((:ffb_d2h, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), r, a.ptr + (i - 1) * sizeof(T), sizeof(T)))
  return r[]
end

# ---------------------------------------------------------------- B2: plans (AbstractFFTs protocol as used by the reference)
mutable struct B200Plan{T,K}    # K = :r2c | :c2c
  handle :: Ptr{Cvoid}
  sz     :: Tuple
end
function makeplan(::Type{T}, sz::Tuple, kind::Symbol) where T
  n = Int64[sz...]
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_plan_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Cint, Cint, Cint),
              h, length(sz), n, ffbtype(T), kind == :r2c ? 0 : 1, 1, 0))
  p = B200Plan{T,kind}(h[], sz)
  finalizer(x -> ccall((:ffb_plan_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), p)
  return p
end
# src/domains.jl:4-5 analogues (the `flags=effort` keyword is dropped exactly like for CuArrays)
plan_flows_fft(a::B200Array{Complex{T}}, args...; flags=nothing, kw...) where T = makeplan(T, size(a), :c2c)
plan_flows_rfft(a::B200Array{T}, args...; flags=nothing, kw...) where T<:AbstractFloat = makeplan(T, size(a), :r2c)

mul!(out::B200Array, p::B200Plan, a::B200Array) =                               # src/diffusion.jl:139,171
  (check(ccall((:ffb_fft_forward, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.handle, a.ptr, out.ptr)); out)
ldiv!(out::B200Array, p::B200Plan, ah::B200Array) =                             # src/diffusion.jl:137,154-155
  (check(ccall((:ffb_fft_inverse, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), p.handle, ah.ptr, out.ptr)); out)
*(p::B200Plan{T,:r2c}, a::B200Array{T}) where T = mul!(B200Array{Complex{T}}(undef, (p.sz[1] ÷ 2 + 1, p.sz[2:end]...)...), p, a)
\(p::B200Plan{T,:r2c}, ah::B200Array{Complex{T}}) where T = ldiv!(B200Array{T}(undef, p.sz...), p, ah)
*(p::B200Plan{T,:c2c}, a::B200Array{Complex{T}}) where T = mul!(similar(a), p, a)
\(p::B200Plan{T,:c2c}, a::B200Array{Complex{T}}) where T = ldiv!(similar(a), p, a)
# `fft / ifft / rfft / irfft` (FFTW names re-exported by the reference, src/FourierFlows.jl:72; `jacobianh` of user code calls them)
FourierFlows.rfft(a::B200Array{T}) where T<:AbstractFloat = plan_flows_rfft(a) * a
FourierFlows.irfft(ah::B200Array{Complex{T}}, nx::Integer) where T = makeplan(T, (Int(nx), size(ah)[2:end]...), :r2c) \ ah
FourierFlows.fft(a::B200Array{Complex{T}}) where T = plan_flows_fft(a) * a
FourierFlows.ifft(a::B200Array{Complex{T}}) where T = plan_flows_fft(a) \ a

# ---------------------------------------------------------------- descriptors shared by the grid-side kernels
struct FFBDesc
  ndim :: Cint; dims :: NTuple{4,Int64}; dtype :: Cint; alias_lo :: NTuple{3,Int32}; alias_hi :: NTuple{3,Int32}
end
rng(r::Nothing) = (Int32(0), Int32(0)); rng(r::UnitRange) = (Int32(first(r)), Int32(last(r)))
function desc(fh::B200Array{T}, g::AbstractGrid, kal) where T
  nd = g isa OneDGrid ? 1 : g isa TwoDGrid ? 2 : 3
  d = ntuple(i -> i <= nd ? Int64(size(fh, i)) : Int64(1), 3)
  nf = Int64(prod(size(fh)[nd+1:end]))
  al = (rng(kal), nd >= 2 ? rng(g.lalias) : rng(nothing), nd >= 3 ? rng(g.malias) : rng(nothing))
  FFBDesc(nd, (d..., nf), ffbtype(T), map(first, al), map(last, al))
end

# `dealias!(fh, grid)` src/domains.jl:428-476 (kralias vs kalias from size(fh,1) == grid.nkr, :437,450,464)
function dealias!(fh::B200Array, g::AbstractGrid{T,A,<:UnitRange}) where {T,A}
  kal = size(fh, 1) == g.nkr ? g.kralias : g.kalias
  d = Ref(desc(fh, g, kal))
  check(ccall((:ffb_dealias, lib), Cint, (Ptr{Cvoid}, Ptr{FFBDesc}), fh.ptr, d))
  return nothing
end

# ---------------------------------------------------------------- grids on the device (src/domains.jl:61-101, 175-223, 311-366)
# The reference builds k, l, m, kr by uploading `fftfreq` ranges and Ksq / invKsq by broadcasts plus a scalar write; here the
# same arrays come from two library kernels (bit-identical: tests/test_gpu_grid.py), so no generic broadcast is needed.
function wavenumbers(::Type{T}, n::Integer, L, half::Bool, shape::NTuple{N,Int}) where {T,N}   # fftfreq / rfftfreq(n, 2π/L*n), Float64 -> T
  a = B200Array{T,N}(undef, shape)
  check(ccall((:ffb_wavenumbers, lib), Cint, (Ptr{Cvoid}, Int64, Cdouble, Cint, Cint), a.ptr, n, L, ffbtype(T), half))
  return a
end
function ksq(::Type{T}, kx::B200Array, l, m, dims::NTuple{N,Int}) where {T,N}                  # (Ksq, invKsq) with invKsq[1,1,1] = 0
  K = B200Array{T,N}(undef, dims); invK = B200Array{T,N}(undef, dims)
  d = Ref(FFBDesc(N, (ntuple(i -> Int64(dims[i]), N)..., ntuple(_ -> Int64(1), 4 - N)...), ffbtype(T), (Int32(0), Int32(0), Int32(0)), (Int32(0), Int32(0), Int32(0))))
  check(ccall((:ffb_ksq, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBDesc}),
              K.ptr, invK.ptr, kx.ptr, l === nothing ? C_NULL : l.ptr, m === nothing ? C_NULL : m.ptr, d))
  return K, invK
end

function FourierFlows.OneDGrid(dev::B200; nx, Lx, x0 = -Lx/2, nthreads = Sys.CPU_THREADS, effort = nothing, T = Float64, aliased_fraction = 1/3)
  mod(nx, 2) != 0 && throw(DomainError("nx must be even"))                       # :66
  dx = Lx/nx; nk = nx; nkr = Int(nx/2 + 1)
  x = range(T(x0), step=T(dx), length=nx)
  k  = wavenumbers(T, nx, Lx, false, (nk,))
  kr = wavenumbers(T, nx, Lx, true, (nkr,))
  _, invksq  = ksq(T, k, nothing, nothing, (nk,))                                # `@. 1 / k^2`, `invksq[1] = 0` (:80-83)
  _, invkrsq = ksq(T, kr, nothing, nothing, (nkr,))
  fftplan  = plan_flows_fft(B200Array{Complex{T},1}(undef, nx))
  rfftplan = plan_flows_rfft(B200Array{T,1}(undef, nx))
  kalias, kralias = getaliasedwavenumbers(nk, nkr, aliased_fraction)
  return OneDGrid{T, typeof(k), typeof(x), typeof(fftplan), typeof(rfftplan), typeof(kalias), typeof(dev)}(
    dev, nx, nk, nkr, dx, Lx, x, k, kr, invksq, invkrsq, fftplan, rfftplan, aliased_fraction, kalias, kralias)
end

function FourierFlows.TwoDGrid(dev::B200; nx, Lx, ny=nx, Ly=Lx, x0=-Lx/2, y0=-Ly/2, nthreads=Sys.CPU_THREADS, effort=nothing,
                               T=Float64, aliased_fraction=1/3)
  (mod(nx, 2) != 0 || mod(ny, 2) != 0) && throw(DomainError("nx and ny must be even"))   # :179
  dx = Lx/nx; dy = Ly/ny; nk = nx; nl = ny; nkr = Int(nx/2 + 1)
  x = range(T(x0), step=T(dx), length=nx); y = range(T(y0), step=T(dy), length=ny)
  k  = wavenumbers(T, nx, Lx, false, (nk, 1))
  l  = wavenumbers(T, ny, Ly, false, (1, nl))
  kr = wavenumbers(T, nx, Lx, true, (nkr, 1))
  Ksq, invKsq   = ksq(T, k, l, nothing, (nk, nl))                                # :197-199
  Krsq, invKrsq = ksq(T, kr, l, nothing, (nkr, nl))                              # :201-203
  fftplan  = plan_flows_fft(B200Array{Complex{T},2}(undef, nx, ny))
  rfftplan = plan_flows_rfft(B200Array{T,2}(undef, nx, ny))
  kalias, kralias = getaliasedwavenumbers(nk, nkr, aliased_fraction)
  lalias, _       = getaliasedwavenumbers(nl, nl, aliased_fraction)
  return TwoDGrid{T, typeof(k), typeof(x), typeof(fftplan), typeof(rfftplan), typeof(kalias), typeof(dev)}(
    dev, nx, ny, nk, nl, nkr, dx, dy, Lx, Ly, x, y, k, l, kr, Ksq, invKsq, Krsq, invKrsq, fftplan, rfftplan,
    aliased_fraction, kalias, kralias, lalias)
end

function FourierFlows.ThreeDGrid(dev::B200; nx, Lx, ny=nx, Ly=Lx, nz=nx, Lz=Lx, x0=-Lx/2, y0=-Ly/2, z0=-Lz/2,
                                 nthreads=Sys.CPU_THREADS, effort=nothing, T=Float64, aliased_fraction=1/3)
  (mod(nx, 2) != 0 || mod(ny, 2) != 0 || mod(nz, 2) != 0) && throw(DomainError("nx, ny, and nz must be even"))   # :316
  dx = Lx/nx; dy = Ly/ny; dz = Lz/nz; nk = nx; nl = ny; nm = nz; nkr = Int(nx/2 + 1)
  x = range(T(x0), step=T(dx), length=nx); y = range(T(y0), step=T(dy), length=ny); z = range(T(z0), step=T(dz), length=nz)
  k  = wavenumbers(T, nx, Lx, false, (nk, 1, 1))
  l  = wavenumbers(T, ny, Ly, false, (1, nl, 1))
  m  = wavenumbers(T, nz, Lz, false, (1, 1, nm))
  kr = wavenumbers(T, nx, Lx, true, (nkr, 1, 1))
  Ksq, invKsq   = ksq(T, k, l, m, (nk, nl, nm))                                  # :338-340
  Krsq, invKrsq = ksq(T, kr, l, m, (nkr, nl, nm))                                # :342-344
  fftplan  = plan_flows_fft(B200Array{Complex{T},3}(undef, nx, ny, nz))
  rfftplan = plan_flows_rfft(B200Array{T,3}(undef, nx, ny, nz))
  kalias, kralias = getaliasedwavenumbers(nk, nkr, aliased_fraction)
  lalias, _       = getaliasedwavenumbers(nl, nl, aliased_fraction)
  malias, _       = getaliasedwavenumbers(nm, nm, aliased_fraction)
  return ThreeDGrid{T, typeof(k), typeof(x), typeof(fftplan), typeof(rfftplan), typeof(kalias), typeof(dev)}(
    dev, nx, ny, nz, nk, nl, nm, nkr, dx, dy, dz, Lx, Ly, Lz, x, y, z, k, l, m, kr, Ksq, invKsq, Krsq, invKrsq,
    fftplan, rfftplan, aliased_fraction, kalias, kralias, lalias, malias)
end

# ---------------------------------------------------------------- coefficients (`equation.L`, ETD coefficients)
struct FFBCoef; ptr :: Ptr{Cvoid}; kind :: Cint; dtype :: Cint; re :: Cdouble; im :: Cdouble; end
coef(L::Number, T) = FFBCoef(C_NULL, 0, ffbtype(fltype(T)), real(L), imag(L))
coef(L::B200Array{S}, T) where S<:Real = FFBCoef(L.ptr, 1, ffbtype(S), 0.0, 0.0)
coef(L::B200Array{Complex{S}}, T) where S = FFBCoef(L.ptr, 2, ffbtype(S), 0.0, 0.0)
fptr(ts) = hasproperty(ts, :filter) ? ts.filter.ptr : C_NULL

# ---------------------------------------------------------------- B3: one stepforward! method per stepper TYPE (reference control flow kept)
# (a union type in the `ts` argument would be ambiguous with the reference's `stepforward!(sol, clock, ts::XTimeStepper, ...)` methods:
#  each concrete stepper type gets its own method, strictly more specific in `sol`)
const B200Sol = B200Array

for TS in (:ETDRK4TimeStepper, :FilteredETDRK4TimeStepper)
  @eval function stepforward!(sol::B200Sol{T}, clock, ts::FourierFlows.$TS, eq, vars, params, grid) where T
    n = Int64(length(sol)); dt = ffbtype(T)
    cE, cE2, cz, ca, cb, cg = (Ref(coef(c, T)) for c in (ts.expLdt, ts.exp½Ldt, ts.ζ, ts.α, ts.β, ts.Γ))
    eq.calcN!(ts.N₁, sol, clock.t, clock, vars, params, grid)                                    # :520
    check(ccall((:ffb_stage_etdrk4_substep12, lib), Cint, (Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{Cvoid}, Cint, Int64),
                ts.sol₁.ptr, cE2, sol.ptr, cz, ts.N₁.ptr, dt, n))                                  # :521
    t2 = clock.t + clock.dt/2
    eq.calcN!(ts.N₂, ts.sol₁, t2, clock, vars, params, grid)                                     # :525
    check(ccall((:ffb_stage_etdrk4_substep12, lib), Cint, (Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{Cvoid}, Cint, Int64),
                ts.sol₂.ptr, cE2, sol.ptr, cz, ts.N₂.ptr, dt, n))                                  # :526
    eq.calcN!(ts.N₃, ts.sol₂, t2, clock, vars, params, grid)                                     # :529
    check(ccall((:ffb_stage_etdrk4_substep3, lib), Cint, (Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64),
                ts.sol₂.ptr, cE2, ts.sol₁.ptr, cz, ts.N₁.ptr, ts.N₃.ptr, dt, n))                   # :530
    eq.calcN!(ts.N₄, ts.sol₂, clock.t + clock.dt, clock, vars, params, grid)                     # :534
    check(ccall((:ffb_stage_etdrk4_update, lib), Cint,
                (Ptr{Cvoid}, Ptr{FFBCoef}, Ptr{FFBCoef}, Ptr{FFBCoef}, Ptr{FFBCoef}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64),
                sol.ptr, cE, ca, cb, cg, ts.N₁.ptr, ts.N₂.ptr, ts.N₃.ptr, ts.N₄.ptr, fptr(ts), dt, n))   # :541 (+ :552)
    clock.t += clock.dt; clock.step += 1
    return nothing
  end
end

for TS in (:RK4TimeStepper, :FilteredRK4TimeStepper)
  @eval function stepforward!(sol::B200Sol{T}, clock, ts::FourierFlows.$TS, eq, vars, params, grid) where T
    n = Int64(length(sol)); dty = ffbtype(T); L = Ref(coef(eq.L, T)); t = clock.t; dt = clock.dt
    sub(rhs, u, c) = check(ccall((:ffb_stage_rk4_substep, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBCoef}, Cdouble, Cint, Int64),
                                 ts.sol₁.ptr, rhs.ptr, u.ptr, sol.ptr, L, c, dty, n))
    eq.calcN!(ts.RHS₁, sol, t, clock, vars, params, grid);            sub(ts.RHS₁, sol, dt/2)      # :239-243
    eq.calcN!(ts.RHS₂, ts.sol₁, t+dt/2, clock, vars, params, grid);   sub(ts.RHS₂, ts.sol₁, dt/2)  # :244-248
    eq.calcN!(ts.RHS₃, ts.sol₁, t+dt/2, clock, vars, params, grid);   sub(ts.RHS₃, ts.sol₁, dt)    # :249-253
    eq.calcN!(ts.RHS₄, ts.sol₁, t+dt, clock, vars, params, grid)                                   # :254
    check(ccall((:ffb_stage_rk4_final, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBCoef}, Cdouble, Ptr{Cvoid}, Cint, Cint, Int64),
                sol.ptr, ts.RHS₁.ptr, ts.RHS₂.ptr, ts.RHS₃.ptr, ts.RHS₄.ptr, ts.sol₁.ptr, L, dt, fptr(ts), 1, dty, n))  # :255,261,279
    clock.t += clock.dt; clock.step += 1
    return nothing
  end
end

for TS in (:LSRK54TimeStepper, :FilteredLSRK54TimeStepper)
  @eval function stepforward!(sol::B200Sol{T}, clock, ts::FourierFlows.$TS, eq, vars, params, grid) where T
    n = Int64(length(sol)); dty = ffbtype(T); L = Ref(coef(eq.L, T)); t = clock.t; dt = clock.dt
    for i = 1:5                                                                                      # :386-392 (`S² = 0` folded into i = 1)
      eq.calcN!(ts.RHS, sol, t + ts.C[i] * dt, clock, vars, params, grid)
      check(ccall((:ffb_stage_lsrk54, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBCoef}, Cdouble, Cdouble, Cdouble, Cint, Ptr{Cvoid}, Cint, Int64),
                  sol.ptr, ts.S².ptr, ts.RHS.ptr, L, real(ts.A[i]), real(ts.B[i]), dt, i == 1, i == 5 ? fptr(ts) : C_NULL, dty, n))
    end
    clock.t += clock.dt; clock.step += 1
    return nothing
  end
end

for TS in (:ForwardEulerTimeStepper, :FilteredForwardEulerTimeStepper)
  @eval function stepforward!(sol::B200Sol{T}, clock, ts::FourierFlows.$TS, eq, vars, params, grid) where T
    eq.calcN!(ts.N, sol, clock.t, clock, vars, params, grid)                                        # :112,143
    L = Ref(coef(eq.L, T))
    check(ccall((:ffb_stage_fe, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBCoef}, Cdouble, Ptr{Cvoid}, Cint, Int64),
                sol.ptr, ts.N.ptr, L, clock.dt, fptr(ts), ffbtype(T), Int64(length(sol))))          # :113,144
    clock.t += clock.dt; clock.step += 1
    return nothing
  end
end

# AB3 keeps the reference's field names; the history "copies" (:647-648) become pointer swaps inside the arrays
for TS in (:AB3TimeStepper, :FilteredAB3TimeStepper)
  @eval function stepforward!(sol::B200Sol{T}, clock, ts::FourierFlows.$TS, eq, vars, params, grid) where T
    eq.calcN!(ts.RHS, sol, clock.t, clock, vars, params, grid)                                      # :639,654
    L = Ref(coef(eq.L, T))
    check(ccall((:ffb_stage_ab3, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBCoef}, Cdouble, Int64, Ptr{Cvoid}, Cint, Int64),
                sol.ptr, ts.RHS.ptr, ts.RHS₋₁.ptr, ts.RHS₋₂.ptr, L, clock.dt, clock.step, fptr(ts), ffbtype(T), Int64(length(sol))))
    clock.t += clock.dt; clock.step += 1
    ts.RHS.ptr, ts.RHS₋₁.ptr, ts.RHS₋₂.ptr = ts.RHS₋₂.ptr, ts.RHS.ptr, ts.RHS₋₁.ptr                 # RHS₋₂ ← RHS₋₁ ← RHS
    return nothing
  end
end

# ---------------------------------------------------------------- ETD coefficients and filter on the device
function getetdcoeffs(dt, L::B200Array{S}; ncirc=32, rcirc=1) where S             # src/timesteppers.jl:689-721
  (ncirc == 32 && rcirc == 1) || error("libfourierflows_b200 implements the reference's default contour (ncirc=32, rcirc=1)")
  CT = S <: Real ? Float64 : Complex{Float64}
  E, E2, ζ, α, β, Γ = (B200Array{CT}(undef, size(L)...) for _ in 1:6)
  c = Ref(coef(L, S))
  check(ccall((:ffb_etd_coeffs, lib), Cint, (Cdouble, Ptr{FFBCoef}, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}),
              dt, c, ffbtype(S), 1, length(L), E.ptr, E2.ptr, ζ.ptr, α.ptr, β.ptr, Γ.ptr, C_NULL))
  return ζ, α, β, Γ
end

# `getexpLs(dt, equation)` (src/timesteppers.jl:673-678): exp(dt L), exp(dt L / 2) evaluated in T on the device by the same
# kernel that forms the contour means (it rounds E, E2 to T before storing, as `@. exp(dt * L)` would)
function getexpLs(dt, eq::Equation{TT,<:B200Array{S}}) where {TT,S}
  L = eq.L
  CT = S <: Real ? Float64 : Complex{Float64}
  E, E2, ζ, α, β, Γ = (B200Array{CT}(undef, size(L)...) for _ in 1:6)
  c = Ref(coef(L, S))
  check(ccall((:ffb_etd_coeffs, lib), Cint, (Cdouble, Ptr{FFBCoef}, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}),
              dt, c, ffbtype(S), 1, length(L), E.ptr, E2.ptr, ζ.ptr, α.ptr, β.ptr, Γ.ptr, C_NULL))
  return E, E2      # Float64 storage of T-rounded values: the stage kernels take (state T, coefficients Float64)
end

function makefilter(g::AbstractGrid{Tg,<:B200Array}, T, sz; order=4, innerK=2/3, outerK=1, tol=1e-15) where Tg   # src/domains.jl:545-546
  f = B200Array{T}(undef, sz...)
  kx = sz[1] == g.nkr ? g.kr : g.k
  d = Ref(desc(f, g, nothing))
  lp = g isa OneDGrid ? C_NULL : g.l.ptr; mp = g isa ThreeDGrid ? g.m.ptr : C_NULL
  dy = g isa OneDGrid ? 0.0 : g.dy;       dz = g isa ThreeDGrid ? g.dz : 0.0
  check(ccall((:ffb_make_filter, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{FFBDesc}),
              f.ptr, kx.ptr, lp, mp, g.dx, dy, dz, order, innerK, outerK, tol, d))
  return f
end

# ---------------------------------------------------------------- elementwise vocabulary for user calcN! (closed set; anything else errors)
"`spectral_mul!(out, in, grid; coef=im, px=1, ...)` lowers `@. out = coef * kx^px * l^py * m^pz * w * in` [+ dealias!]"
function spectral_mul!(out::B200Array, inp::B200Array, g; coef=1, px=0, py=0, pz=0, w=nothing, accumulate=false, dealias=false)
  half = size(inp, 1) == g.nkr
  kal = dealias ? (half ? g.kralias : g.kalias) : nothing
  d = Ref(desc(inp, g, kal))
  lp = g isa OneDGrid ? C_NULL : g.l.ptr; mp = g isa ThreeDGrid ? g.m.ptr : C_NULL
  check(ccall((:ffb_ew_spectral_mul, lib), Cint,
              (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{FFBDesc}),
              out.ptr, inp.ptr, real(coef), imag(coef), (half ? g.kr : g.k).ptr, px, lp, py, mp, pz, w === nothing ? C_NULL : w.ptr, accumulate, dealias && kal !== nothing, d))
  return out
end
"`axpby!(out, a, x, b, y)` lowers `@. out = a*x + b*y` (y may be `nothing`): `L = @. -ν * grid.Krsq`, history copies, sums of fields"
function axpby!(out::B200Array{T}, a::Real, x::B200Array{T}, b::Real=0.0, y::Union{B200Array{T},Nothing}=nothing) where T
  check(ccall((:ffb_ew_axpby, lib), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Cint, Int64),
              out.ptr, a, x.ptr, b, y === nothing ? C_NULL : y.ptr, T <: Complex, ffbtype(T), length(out)))
  return out
end
"`mul_real!(out, x, y)` lowers the physical-space products `@. cx *= κ`, `@. u *= ζ` (src/diffusion.jl:138)"
mul_real!(out::B200Array{T}, x::B200Array{T}, y::B200Array{T}) where T<:AbstractFloat =
  (check(ccall((:ffb_ew_mul_real, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Int64), out.ptr, x.ptr, y.ptr, ffbtype(T), length(out))); out)

# ---------------------------------------------------------------- arithmetic and broadcast on B200Arrays: a closed vocabulary
# The reference's own code performs these on device arrays: `0N` (ForwardEulerTimeStepper, src/timesteppers.jl:103), `@. N = 0`
# (src/diffusion.jl:131), `@. L = -κ * kr^2` (:84), `@. cxh = im * kr * sol` (:136,152), `@. cx *= κ` (:138), `@. ch = sol` (:151),
# `vars.c .= A(c)` (:171).  They are lowered onto ffb_ew_axpby / ffb_ew_mul_real / ffb_ew_spectral_mul / memset; anything
# outside the vocabulary raises (no CPU fallback).
*(s::Number, a::B200Array{T}) where T<:Real = axpby!(similar(a), s, a)
*(a::B200Array{T}, s::Number) where T<:Real = s * a
*(s::Real, a::B200Array{T}) where T<:Complex = axpby!(similar(a), s, a)
*(a::B200Array{T}, s::Real) where T<:Complex = s * a
/(a::B200Array, s::Real) = inv(s) * a
-(a::B200Array) = -1 * a
+(a::B200Array{T}, b::B200Array{T}) where T = axpby!(similar(a), 1, a, 1, b)
-(a::B200Array{T}, b::B200Array{T}) where T = axpby!(similar(a), 1, a, -1, b)

struct B200Style <: AbstractArrayStyle{Any} end
B200Style(::Val) = B200Style()
BroadcastStyle(::Type{<:B200Array}) = B200Style()
BroadcastStyle(s::B200Style, ::DefaultArrayStyle) = s
Base.similar(bc::Broadcasted{B200Style}, ::Type{T}) where T = B200Array{T}(undef, map(length, axes(bc)))

unsupported(bc) = error("broadcast `", bc.f, "` over B200Arrays is outside the library's elementwise vocabulary ",
                        "(linear combinations, same-shape products / squares, c * kx^a * l^b * m^c * w * field, zero fill): ",
                        "write it with spectral_mul!, axpby!, mul_real! -- there is no CPU fallback")
hasdev(x) = x isa B200Array
hasdev(bc::Broadcasted) = any(hasdev, bc.args)
scalarof(x::Number) = x
scalarof(bc::Broadcasted) = hasdev(bc) ? nothing : Base.Broadcast.materialize(bc)     # e.g. `-κ`: a broadcast over scalars only
scalarof(x) = nothing
isvec(a::B200Array, dest) = size(a) != size(dest) && count(>(1), size(a)) <= 1        # kr (nkr,1), l (1,nl), m (1,1,nm)

# a product  c * v1^p1 * v2^p2 * ... * [w] * [field]  flattened into (scalar, vectors-with-powers, dense real factor, field)
function factors!(acc, x, dest)
  if (c = scalarof(x)) !== nothing
    acc.c *= c
  elseif x isa B200Array
    if isvec(x, dest); push!(acc.vecs, (x, 1))
    elseif eltype(x) <: Real && acc.w === nothing && eltype(dest) <: Complex; acc.w = x
    elseif acc.field === nothing; acc.field = x
    elseif eltype(x) <: Real && acc.w === nothing; acc.w = x
    else return false end
  elseif x isa Broadcasted && x.f === (*)
    for a in x.args; factors!(acc, a, dest) || return false; end
  elseif x isa Broadcasted && x.f === Base.literal_pow && x.args[2] isa B200Array      # `kr^2` lowers to literal_pow(^, kr, Val(2))
    pw = typeof(x.args[3]).parameters[1]
    (pw isa Integer && pw >= 1) || return false
    if isvec(x.args[2], dest); push!(acc.vecs, (x.args[2], pw))
    elseif pw == 2 && acc.field === nothing && acc.w === nothing; acc.field = x.args[2]; acc.w = x.args[2]   # A^2 = A * A
    else return false end
  elseif x isa Broadcasted && x.f === (-) && length(x.args) == 1
    acc.c *= -1
    factors!(acc, x.args[1], dest) || return false
  else
    return false
  end
  return true
end
mutable struct Factors; c :: Any; vecs :: Vector{Any}; w :: Any; field :: Any; end

axisof(v::B200Array) = something(findfirst(>(1), size(v)), 1)
function lower_product!(dest::B200Array{T,N}, bc, accumulate::Bool=false) where {T,N}
  acc = Factors(1, Any[], nothing, nothing)
  factors!(acc, bc, dest) || unsupported(bc)
  vec = Any[nothing, nothing, nothing]; pw = [0, 0, 0]
  for (v, q) in acc.vecs
    ax = axisof(v); vec[ax] === nothing || vec[ax] === v || unsupported(bc)
    vec[ax] = v; pw[ax] += q
  end
  if T <: Complex && acc.field !== nothing && eltype(acc.field) <: Complex
    d = Ref(FFBDesc(min(N, 3), (ntuple(i -> Int64(size(dest, i)), min(N, 3))..., ntuple(_ -> Int64(1), 3 - min(N, 3))..., Int64(prod(size(dest)[4:end]))),
                    ffbtype(T), (Int32(0), Int32(0), Int32(0)), (Int32(0), Int32(0), Int32(0))))
    vp(i) = vec[i] === nothing ? C_NULL : vec[i].ptr
    check(ccall((:ffb_ew_spectral_mul, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{FFBDesc}),
                dest.ptr, acc.field.ptr, real(acc.c), imag(acc.c), vp(1), pw[1], vp(2), pw[2], vp(3), pw[3],
                acc.w === nothing ? C_NULL : acc.w.ptr, accumulate, false, d))
  elseif T <: Real && isempty(acc.vecs) && acc.field !== nothing && !accumulate       # real same-shape: c * A [* W]
    imag(acc.c) == 0 || unsupported(bc)
    if acc.w === nothing
      axpby!(dest, real(acc.c), acc.field)
    else
      mul_real!(dest, acc.field, acc.w)
      real(acc.c) == 1 || axpby!(dest, real(acc.c), dest)
    end
  elseif T <: Real && length(acc.vecs) == 1 && acc.field === nothing && acc.w === nothing && N == 1 && !accumulate   # `-κ * kr^2` on a 1-D grid
    v, q = acc.vecs[1]; size(v) == size(dest) || unsupported(bc)
    q == 1 ? axpby!(dest, real(acc.c), v) : q == 2 ? (mul_real!(dest, v, v); axpby!(dest, real(acc.c), dest)) : unsupported(bc)
  else
    unsupported(bc)
  end
  return dest
end

# dest .= bc
function copyto!(dest::B200Array{T}, bc::Broadcasted{B200Style}) where T
  f, args = bc.f, bc.args
  if f === identity && length(args) == 1 && args[1] isa B200Array                      # `@. ch = sol`, `c .= A(c)`
    return copyto!(dest, args[1])
  elseif (f === (+) || f === (-)) && length(args) == 2                                 # a*x ± b*y with x, y same-shape arrays
    t1, t2 = linterm(args[1], dest), linterm(args[2], dest)
    (t1 === nothing || t2 === nothing) && unsupported(bc)
    return axpby!(dest, t1[1], t1[2], f === (+) ? t2[1] : -t2[1], t2[2])
  else
    return lower_product!(dest, bc)
  end
end
function copyto!(dest::B200Array, bc::Broadcasted{<:DefaultArrayStyle{0}})              # `@. N = 0`
  v = Base.Broadcast.materialize(bc)
  return fill!(dest, v)
end
Base.copy(bc::Broadcasted{B200Style}) = copyto!(similar(bc, Base.Broadcast.combine_eltypes(bc.f, bc.args)), bc)
# one addend of a linear combination: (real scalar, same-shape array) or nothing
function linterm(x, dest)
  x isa B200Array && size(x) == size(dest) && return (1, x)
  if x isa Broadcasted && x.f === (*) && length(x.args) == 2
    c1, c2 = scalarof(x.args[1]), scalarof(x.args[2])
    c1 !== nothing && x.args[2] isa B200Array && size(x.args[2]) == size(dest) && c1 isa Real && return (c1, x.args[2])
    c2 !== nothing && x.args[1] isa B200Array && size(x.args[1]) == size(dest) && c2 isa Real && return (c2, x.args[1])
  end
  return nothing
end

# `parsevalsum2(uh, grid)` / `parsevalsum(uh, grid)` (src/utils.jl:113-183): device reduction, L/n^2 normalisation on the host
function parsevalpartial(uh::B200Array, g, abs2::Bool)
  r = Ref{Cdouble}(0.0)
  d = Ref(desc(uh, g, nothing))
  check(ccall((:ffb_parseval_sum, lib), Cint, (Ptr{Cdouble}, Ptr{Cvoid}, Cint, Cint, Ptr{FFBDesc}), r, uh.ptr, abs2, size(uh, 1) == g.nkr, d))
  return r[]
end
FourierFlows.parsevalsum2(uh::B200Array, g::OneDGrid) = parsevalpartial(uh, g, true) * g.Lx / g.nx^2
FourierFlows.parsevalsum2(uh::B200Array, g::TwoDGrid) = parsevalpartial(uh, g, true) * g.Lx * g.Ly / (g.nx^2 * g.ny^2)
FourierFlows.parsevalsum(uh::B200Array, g::OneDGrid) = parsevalpartial(uh, g, false) * g.Lx / g.nx^2
FourierFlows.parsevalsum(uh::B200Array, g::TwoDGrid) = parsevalpartial(uh, g, false) * g.Lx * g.Ly / (g.nx^2 * g.ny^2)

# ---------------------------------------------------------------- fused transforms (north star: multiplies / dealias folded into the passes)
struct FFBFuse
  cr :: Cdouble; ci :: Cdouble
  kx :: Ptr{Cvoid}; l :: Ptr{Cvoid}; m :: Ptr{Cvoid}; w :: Ptr{Cvoid}; acc :: Ptr{Cvoid}
  ar :: Cdouble; ai :: Cdouble
  akx :: Ptr{Cvoid}; al :: Ptr{Cvoid}; am :: Ptr{Cvoid}
  dealias :: Cint; alias_lo :: NTuple{3,Int32}; alias_hi :: NTuple{3,Int32}
  mul :: Ptr{Cvoid}
  square_input :: Cint
  galias_lo :: NTuple{3,Int32}; galias_hi :: NTuple{3,Int32}
end
"`ldiv!(out, plan, ah, fuse)`: out = irfft(factor .* ah) [.* mul];  `mul!(outh, plan, a, fuse)`: outh = factor .* rfft(a) [+ g .* acc] [dealiased]"
ldiv!(out::B200Array, p::B200Plan, ah::B200Array, f::FFBFuse) =
  (check(ccall((:ffb_fft_inverse_ex, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBFuse}), p.handle, ah.ptr, out.ptr, Ref(f))); out)
mul!(out::B200Array, p::B200Plan, a::B200Array, f::FFBFuse) =
  (check(ccall((:ffb_fft_forward_ex, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{FFBFuse}), p.handle, a.ptr, out.ptr, Ref(f))); out)
"`ldiv!(outs, plan, ah, fuses)`: outs[v] = irfft(factor_v .* ah) [.* mul_v] for every v, in order; `ah` (and a dense factor the variants share)
is read once by the first four-step sub-pass (zeta, u, v of a vorticity calcN! from one `sol`)"
function ldiv!(outs::Vector{<:B200Array}, p::B200Plan, ah::B200Array, fs::Vector{FFBFuse})
  length(outs) == length(fs) || throw(ArgumentError("one FFBFuse per output"))
  ptrs = Ptr{Cvoid}[o.ptr for o in outs]
  GC.@preserve outs check(ccall((:ffb_fft_inverse_multi, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{FFBFuse}),
                                p.handle, ah.ptr, length(outs), ptrs, fs))
  return outs
end

# ---------------------------------------------------------------- `jacobianh` (src/utils.jl:190-197) and the asynchronous output path
function FourierFlows.jacobianh(a::B200Array{T,2}, b::B200Array{T,2}, g::TwoDGrid) where T<:AbstractFloat
  out = B200Array{Complex{T}}(undef, g.nkr, g.nl); sh = similar(out); p1 = similar(a); p2 = similar(a)
  check(ccall((:ffb_jacobianh, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
              g.rfftplan.handle, out.ptr, a.ptr, b.ptr, g.kr.ptr, g.l.ptr, sh.ptr, p1.ptr, p2.ptr))
  return out
end
FourierFlows.jacobian(a::B200Array{T,2}, b::B200Array{T,2}, g::TwoDGrid) where T<:AbstractFloat = g.rfftplan \ FourierFlows.jacobianh(a, b, g)
"`mul!(out, x, y)`: out = x .* y for same-shape real / complex device arrays"
function mul!(out::B200Array{To}, x::B200Array{Tx}, y::B200Array{Ty}) where {To,Tx,Ty}
  check(ccall((:ffb_ew_mul, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint, Int64),
              out.ptr, x.ptr, Tx <: Complex, y.ptr, Ty <: Complex, ffbtype(To), length(out)))
  return out
end

# `saveoutput(out)` (src/output.jl:61-79) downloads with a blocking `Array(data)`; the snapshot ring stages the field on the device in
# stream order and copies it to pinned host memory on its own stream: `slot = snapshot!(ring, a)` inside the step loop (non-blocking),
# `h = wait(ring, slot, T, dims)` on the writer's side, `release!(ring, slot)` afterwards.
mutable struct SnapshotRing; handle :: Ptr{Cvoid}; end
function SnapshotRing(bytes::Integer, nbuf::Integer=2)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_snapshot_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Csize_t, Cint), h, bytes, nbuf))
  r = SnapshotRing(h[])
  finalizer(x -> ccall((:ffb_snapshot_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), r)
  return r
end
function snapshot!(r::SnapshotRing, a::B200Array{T}) where T
  slot = Ref{Cint}(-1)
  check(ccall((:ffb_snapshot_begin, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Ptr{Cint}), r.handle, a.ptr, length(a) * sizeof(T), slot))
  return slot[]
end
function Base.wait(r::SnapshotRing, slot::Integer, ::Type{T}, dims::Dims) where T
  p = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_snapshot_wait, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), r.handle, slot, p))
  return unsafe_wrap(Array, Ptr{T}(p[]), dims)      # view of the pinned buffer, valid until release!
end
release!(r::SnapshotRing, slot::Integer) = check(ccall((:ffb_snapshot_release, lib), Cint, (Ptr{Cvoid}, Cint), r.handle, slot))

# ---------------------------------------------------------------- C-driven loop (`ffb_step`): benchmarks, or user calcN! passed as @cfunction
# typedef int (*ffb_calcN_fn)(void* N, const void* sol, double t, void* user)
struct FFBProblemConfig
  ndim :: Cint; n :: NTuple{3,Int64}; L :: NTuple{3,Cdouble}; dtype :: Cint; aliased_fraction :: Cdouble
  stepper :: Cint; filtered :: Cint; filter_order :: Cdouble; filter_innerK :: Cdouble; filter_outerK :: Cdouble; filter_tol :: Cdouble
  dt :: Cdouble; calcN :: Cint; callback :: Ptr{Cvoid}; user :: Ptr{Cvoid}; nu :: Cdouble; scalar_zero_L :: Cint
  kappa :: Ptr{Cvoid}; coef_dtype :: Cint; fused :: Cint; dist :: Ptr{Cvoid}
end
mutable struct B200Problem; handle :: Ptr{Cvoid}; end
function B200Problem(cfg::FFBProblemConfig)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_problem_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Ptr{FFBProblemConfig}), h, Ref(cfg)))
  p = B200Problem(h[])
  finalizer(x -> ccall((:ffb_problem_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), p)
  return p
end
stepforward!(p::B200Problem, nsteps::Integer=1) = check(ccall((:ffb_step, lib), Cint, (Ptr{Cvoid}, Int64), p.handle, nsteps))   # src/timesteppers.jl:14-20
FourierFlows.step_until!(p::B200Problem, stop_time) = check(ccall((:ffb_step_until, lib), Cint, (Ptr{Cvoid}, Cdouble), p.handle, stop_time))  # :734-760

# Host-buffer pipeline: independent spectral states in pinned host memory (`pinned(T, dims)`), uploaded / stepped / downloaded with the
# copies of neighbouring submissions beside the steps -- instead of `sol .= A(h); stepforward!(prob); Array(sol)`, which idles the GPU
# during both copies.  `t = submit!(pipe, hin, hout, nsteps)`; `wait(pipe, t)`: `hout` holds the stepped state.
function pinned(::Type{T}, dims::Dims) where T
  p = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_host_alloc_pinned, lib), Cint, (Ptr{Ptr{Cvoid}}, Csize_t), p, prod(dims) * sizeof(T)))
  a = unsafe_wrap(Array, Ptr{T}(p[]), dims)
  finalizer(x -> ccall((:ffb_host_free_pinned, lib), Cint, (Ptr{Cvoid},), pointer(x)), a)
  return a
end
mutable struct HostPipeline; handle :: Ptr{Cvoid}; depth :: Int; end
function HostPipeline(prob::B200Problem, depth::Integer=3)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_pipeline_create, lib), Cint, (Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Cint), h, prob.handle, depth))
  q = HostPipeline(h[], depth)
  finalizer(x -> ccall((:ffb_pipeline_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), q)
  return q
end
function submit!(q::HostPipeline, hin::Array, hout::Array, nsteps::Integer=1)
  t = Ref{Cint}(-1)
  check(ccall((:ffb_pipeline_submit, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cint}), q.handle, hin, hout, nsteps, t))
  return t[]
end
Base.wait(q::HostPipeline, ticket::Integer) = check(ccall((:ffb_pipeline_wait, lib), Cint, (Ptr{Cvoid}, Cint), q.handle, ticket))

# ---------------------------------------------------------------- multi-GPU: one Julia process per GPU (MPI.jl moves the id and the IPC handles)
# comm = MPI.COMM_WORLD; id = Vector{UInt8}(undef, 128); rank == 0 && ffb_dist_unique_id(id); MPI.Bcast!(id, 0, comm)
# dist = ffb_dist_init(rank, nranks, id); plan = ffb_plan_create_dist(3, n, dtype, dist, 0)
# peer memory: ffb_plan_dist_recv_buffers -> ffb_dist_ipc_export (handle + offset) -> MPI.Allgather -> ffb_dist_ipc_open ->
#              ffb_plan_dist_set_peers -> ffb_sync + MPI.Barrier -> ffb_plan_dist_set_exchange(plan, 1 #= FFB_EXCHANGE_PEER_STORE =#)
function distinit(rank::Integer, nranks::Integer, id::Vector{UInt8})
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_dist_init, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{UInt8}), h, rank, nranks, id))
  return h[]
end
function makedistplan(::Type{T}, sz::NTuple{3,Int}, dist::Ptr{Cvoid}; nchunks=0) where T
  h = Ref{Ptr{Cvoid}}(C_NULL)
  check(ccall((:ffb_plan_create_dist, lib), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Cint, Ptr{Cvoid}, Cint), h, 3, Int64[sz...], ffbtype(T), dist, nchunks))
  p = B200Plan{T,:r2c}(h[], sz)     # mul!/ldiv! act on the local slabs: (nx, ny, nz/P) <-> (nx/2+1, ny/P, nz)
  finalizer(x -> ccall((:ffb_plan_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), p)
  return p
end

end # module
