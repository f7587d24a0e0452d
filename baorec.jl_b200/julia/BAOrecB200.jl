# BAOrecB200.jl -- drop-in device back end for BAOrec.jl on NVIDIA B200 (sm_100a).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither Julia nor CUDA.jl is installed in the build
# image.  The identical C ABI is exercised end to end by the Python ctypes mirror
# (baorec.jl_b200/host.py) and the GPU parity tests; this file is the binding a BAOrec.jl
# maintainer adds.  It re-defines the CuArray methods of the reference (same names, same
# positional signatures, same in-place semantics) as thin `ccall`s into libbaorec_b200.so:
#
#   reference method (file:line)                         -> C entry point
#   setup_fft!(recon, ::CuArray)        src/recon.jl:35   -> baorec_plan
#   cic!(ρ::CuArray, ...)               src/mas.jl:100    -> baorec_cic_scatter_f32
#   read_cic!(out::CuArray, ...)        src/mas.jl:320    -> baorec_gather_f32
#   smooth!(::CuArray, ...)             src/utils.jl:85   -> baorec_smooth_f32
#   setup_overdensity!(::CuArray, ...)  src/recon.jl:42,60-> baorec_setup_overdensity_f32
#   iterate!(::CuArray, ...)            src/iterative.jl:151 -> baorec_iterate_f32
#   reconstructed_overdensity!          src/recon.jl:93,111  -> baorec_reconstructed_overdensity_f32
#   jacobi!/residual!/reduce!/prolong!/vcycle!/fmg  src/multigrid.jl:193,378,641,511,689,722
#                                                        -> baorec_mg_*_f32
#   reconstructed_potential!            src/recon.jl:184,198 -> baorec_reconstructed_potential_f32
#   compute_displacements(::CuArray..)  src/iterative.jl:229, src/multigrid.jl:781
#                                                        -> baorec_compute_displacements_f32
#   read_shifts(...::CuVector...)       src/recon.jl:333  -> baorec_read_shifts_f32
#   reconstructed_positions             src/recon.jl:366  -> baorec_reconstructed_positions_f32
#   setup_box                           src/utils.jl:100  -> baorec_setup_box_f32
#
# Usage inside BAOrec.jl:  include("BAOrecB200.jl") after the other includes; every method
# below is more specific than the generic AbstractArray ones, so CuArray inputs dispatch here.

module BAOrecB200

using CUDA
using StaticArrays
import ..BAOrec: AbstractRecon, IterativeRecon, MultigridRecon, setup_fft!, cic!, read_cic!, smooth!,
                 setup_overdensity!, iterate!, reconstructed_overdensity!, reconstructed_potential!,
                 jacobi!, residual!, reduce!, prolong!, vcycle!, fmg, compute_displacements,
                 read_shifts, reconstructed_positions, setup_box

const libbaorec = get(ENV, "BAOREC_B200_LIB", "libbaorec_b200.so")

# struct baorec_params (include/baorec_b200.h) -- field order and types must match exactly
struct Params
    bias::Cfloat
    f::Cfloat
    smoothing_radius::Cfloat
    beta::Cfloat
    n_iter::Int32
    has_los::Int32
    los::NTuple{3,Cfloat}
    jacobi_damping_factor::Cfloat
    jacobi_niterations::Int32
    vcycle_niterations::Int32
    mas::Int32
    ran_min::Cfloat
    box_pad::Cfloat
end

const ITERATIVE = Cint(0)
const MULTIGRID = Cint(1)
const FIELD = Dict(:disp => Cint(0), :rsd => Cint(1), :sum => Cint(2))

algorithm(::IterativeRecon) = ITERATIVE
algorithm(::MultigridRecon) = MULTIGRID

function Params(r::AbstractRecon)
    los = r.los === nothing ? (0f0, 0f0, 0f0) : (Float32(r.los[1]), Float32(r.los[2]), Float32(r.los[3]))
    Params(Float32(r.bias), Float32(r.f), Float32(r.smoothing_radius), Float32(r.β),
           Int32(r isa IterativeRecon ? r.n_iter : 0), Int32(r.los === nothing ? 0 : 1), los,
           Float32(r isa MultigridRecon ? r.jacobi_damping_factor : 0.4f0),
           Int32(r isa MultigridRecon ? r.jacobi_niterations : 5),
           Int32(r isa MultigridRecon ? r.vcycle_niterations : 6),
           Int32(0), 0.01f0, 500f0)
end

struct BaorecError <: Exception
    code::Cint
    msg::String
end

function check(code::Cint)
    code == 0 && return nothing
    msg = unsafe_string(ccall((:baorec_last_error, libbaorec), Cstring, ()))
    throw(BaorecError(code, msg))     # -5 = particle outside the mesh (reference: BoundsError)
end

# One context per device, created lazily; it owns plans and all scratch memory.
const CTX = Dict{Int,Ptr{Cvoid}}()
function context()
    dev = CUDA.deviceid(CUDA.device())
    get!(CTX, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:baorec_create, libbaorec), Cint, (Cint, Ptr{Ptr{Cvoid}}), dev, h))
        h[]
    end
end

stream() = CUDA.stream().handle            # the caller's task-local stream
# a Vector{Float32} converts to Ptr{Cfloat} and is rooted by ccall for the duration of the call (Base defines no
# unsafe_convert(Ptr{Cfloat}, ::Ref{NTuple{3,Float32}}), so a Ref of a tuple would be a MethodError)
f3(v) = Float32[v[1], v[2], v[3]]
# Every method below must be MORE SPECIFIC than the reference's in EVERY argument (an untyped argument where the
# reference's CuArray method has a typed one makes the call ambiguous): the same argument types with T = Float32.
const V3 = SVector{3,Float32}
const Vec3 = Tuple{AbstractVector{Float32},AbstractVector{Float32},AbstractVector{Float32}}   # k_vec / x_vec results
ptr(a::CuArray{Float32}) = reinterpret(Ptr{Cvoid}, pointer(a))
ptr(::Nothing) = C_NULL

function plan!(mesh::CuArray{Float32,3}, box_size, box_min)
    ctx = context()
    nx, ny, nz = size(mesh)
    check(ccall((:baorec_plan, libbaorec), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}),
                ctx, nx, ny, nz, f3(box_size), f3(box_min)))
    ctx
end

# setup_fft! -- the plan object the reference stores in recon.fft_plan is the context handle here
function setup_fft!(recon::AbstractRecon, field::CuArray{Float32,3})
    recon.fft_plan = context()
end

function cic!(ρ::CuArray{Float32,3}, x::CuArray{Float32}, y::CuArray{Float32}, z::CuArray{Float32},
              w::CuArray{Float32}, box_size::SVector{3,Float32}, box_min::SVector{3,Float32}; wrap::Bool = true)
    ctx = plan!(ρ, box_size, box_min)
    check(ccall((:baorec_cic_scatter_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Cvoid}),
                ctx, ptr(ρ), ptr(x), ptr(y), ptr(z), ptr(w), length(x), wrap, 0, stream()))
    ρ
end

function read_cic!(out::CuArray{Float32}, field::CuArray{Float32,3}, x::CuArray{Float32}, y::CuArray{Float32},
                   z::CuArray{Float32}, box_size::SVector{3,Float32}, box_min::SVector{3,Float32}; wrap = true)
    ctx = plan!(field, box_size, box_min)
    check(ccall((:baorec_gather_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                ctx, ptr(field), ptr(x), ptr(y), ptr(z), length(x), ptr(out), 0, stream()))
    out
end

function smooth!(field::CuArray{Float32,3}, smoothing_radius::Float32, box_size::SVector{3,Float32}, fft_plan)
    ctx = plan!(field, box_size, SVector(0f0, 0f0, 0f0))
    check(ccall((:baorec_smooth_f32, libbaorec), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cfloat, Ptr{Cvoid}),
                ctx, ptr(field), smoothing_radius, stream()))
    field
end

function _catalog_call(mesh, recon, r)
    ctx = plan!(mesh, recon.box_size, recon.box_min)
    p = Ref(Params(recon))
    nr = r === nothing ? 0 : length(r[1])
    rp = r === nothing ? (C_NULL, C_NULL, C_NULL, C_NULL) : map(ptr, r)
    ctx, p, nr, rp
end

function setup_overdensity!(δ::CuArray{Float32,3}, recon::AbstractRecon, x::CuArray{Float32}, y::CuArray{Float32},
                            z::CuArray{Float32}, w::CuArray{Float32}, wrap = true)
    ctx, p, _, rp = _catalog_call(δ, recon, nothing)
    check(ccall((:baorec_setup_overdensity_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                ctx, p, ptr(δ), ptr(x), ptr(y), ptr(z), ptr(w), length(x), rp[1], rp[2], rp[3], rp[4], 0, wrap, stream()))
    δ
end

function setup_overdensity!(δ::CuArray{Float32,3}, recon::AbstractRecon, x::CuArray{Float32}, y::CuArray{Float32},
                            z::CuArray{Float32}, w::CuArray{Float32}, rx::CuArray{Float32}, ry::CuArray{Float32},
                            rz::CuArray{Float32}, rw::CuArray{Float32}, ran_min = 0.01)
    ctx, p, nr, rp = _catalog_call(δ, recon, (rx, ry, rz, rw))
    check(ccall((:baorec_setup_overdensity_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Ptr{Cvoid}),
                ctx, p, ptr(δ), ptr(x), ptr(y), ptr(z), ptr(w), length(x), rp[1], rp[2], rp[3], rp[4], nr, false, stream()))
    δ
end

function iterate!(δ_r::CuArray{Float32,3}, δ_s::CuArray{Float32,3}, k⃗::Vec3, iter::Int, β::Float32, fft_plan;
                  r̂ = nothing, x⃗ = nothing)
    # k⃗ / x⃗ are accepted for signature parity; the plan holds the same tables on the device
    los = r̂ === nothing ? C_NULL : f3(r̂)
    check(ccall((:baorec_iterate_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cfloat, Ptr{Cfloat}, Ptr{Cvoid}),
                context(), ptr(δ_r), ptr(δ_s), iter, β, los, stream()))
    δ_r
end

# reconstructed_overdensity! (src/recon.jl:93, 111) / reconstructed_potential! (src/recon.jl:184, 197): the two
# arities of the reference, each strictly more specific than its AbstractArray method (no Vararg: a Vararg method is
# not a subtype of either fixed-arity signature and would leave the dispatch to the specificity heuristics).
for (fn, sym, R) in ((:reconstructed_overdensity!, :baorec_reconstructed_overdensity_f32, :IterativeRecon),
                     (:reconstructed_potential!, :baorec_reconstructed_potential_f32, :MultigridRecon))
    @eval function $fn(mesh::CuArray{Float32,3}, recon::$R, x::CuVector{Float32}, y::CuVector{Float32},
                       z::CuVector{Float32}, w::CuVector{Float32})
        ctx, p, _, rp = _catalog_call(mesh, recon, nothing)
        check(ccall(($(QuoteNode(sym)), libbaorec), Cint,
                    (Ptr{Cvoid}, Ptr{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                    ctx, p, ptr(mesh), ptr(x), ptr(y), ptr(z), ptr(w), length(x), rp[1], rp[2], rp[3], rp[4], 0, stream()))
        mesh
    end
    @eval function $fn(mesh::CuArray{Float32,3}, recon::$R, x::CuVector{Float32}, y::CuVector{Float32},
                       z::CuVector{Float32}, w::CuVector{Float32}, rx::CuVector{Float32}, ry::CuVector{Float32},
                       rz::CuVector{Float32}, rw::CuVector{Float32})
        ctx, p, nr, rp = _catalog_call(mesh, recon, (rx, ry, rz, rw))
        check(ccall(($(QuoteNode(sym)), libbaorec), Cint,
                    (Ptr{Cvoid}, Ptr{Params}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                    ctx, p, ptr(mesh), ptr(x), ptr(y), ptr(z), ptr(w), length(x), rp[1], rp[2], rp[3], rp[4], nr, stream()))
        mesh
    end
end

# compute_displacements (src/iterative.jl:229, src/multigrid.jl:781): one method per recon type, as in the reference
for R in (:IterativeRecon, :MultigridRecon)
    @eval function compute_displacements(mesh::CuArray{Float32,3}, x::CuVector{Float32}, y::CuVector{Float32},
                                         z::CuVector{Float32}, recon::$R)
        ctx = plan!(mesh, recon.box_size, recon.box_min)
        out = (similar(x), similar(x), similar(x))
        check(ccall((:baorec_compute_displacements_f32, libbaorec), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid},
                     Ptr{Cvoid}, Cint, Ptr{Cvoid}),
                    ctx, ptr(mesh), algorithm(recon), ptr(x), ptr(y), ptr(z), length(x), ptr(out[1]), ptr(out[2]),
                    ptr(out[3]), 0, stream()))
        out
    end
end

# read_shifts / reconstructed_positions (src/recon.jl:333-364, 366-388).  The C symbol of a ccall must be a literal
# (the (name, library) tuple may not reference local variables), so both methods are generated with the symbol
# interpolated at definition time, like reconstructed_overdensity! / reconstructed_potential! above.
for (fn, sym, positions) in ((:read_shifts, :baorec_read_shifts_f32, 0),
                             (:reconstructed_positions, :baorec_reconstructed_positions_f32, 1))
    @eval function $fn(recon::AbstractRecon, x::CuVector{Float32}, y::CuVector{Float32}, z::CuVector{Float32},
                       mesh::CuArray{Float32,3}; field = :disp)
        ctx = plan!(mesh, recon.box_size, recon.box_min)
        out = (similar(x), similar(x), similar(x))
        p = Ref(Params(recon))
        if mesh === recon.result_cache
            # the mesh run! produced: the library kept its delta_k, the forward transform is skipped
            check(ccall((:baorec_read_result_cache_f32, libbaorec), Cint,
                        (Ptr{Cvoid}, Ptr{Params}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint,
                         Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                        ctx, p, algorithm(recon), ptr(mesh), ptr(x), ptr(y), ptr(z), length(x), FIELD[field],
                        $positions, ptr(out[1]), ptr(out[2]), ptr(out[3]), stream()))
            return out
        end
        check(ccall(($(QuoteNode(sym)), libbaorec), Cint,
                    (Ptr{Cvoid}, Ptr{Params}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint,
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    ctx, p, algorithm(recon), ptr(mesh), ptr(x), ptr(y), ptr(z), length(x), FIELD[field],
                    ptr(out[1]), ptr(out[2]), ptr(out[3]), stream()))
        out
    end
end

# reconstructed_positions(recon, x, y, z; field) without a mesh needs no override: the reference's method
# (src/recon.jl:382-388) forwards recon.result_cache to the method above.

function setup_box(x::CuVector{Float32}, y::CuVector{Float32}, z::CuVector{Float32}, box_pad)
    bs = Vector{Float32}(undef, 3); bm = Vector{Float32}(undef, 3)
    check(ccall((:baorec_setup_box_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cfloat, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                context(), ptr(x), ptr(y), ptr(z), length(x), Float32(box_pad), bs, bm, stream()))
    SVector{3,Float32}(bs), SVector{3,Float32}(bm)
end

# ---- multigrid primitives (src/multigrid.jl) -------------------------------------------------
_los(los) = los === nothing ? C_NULL : f3(los)

function jacobi!(v::CuArray{Float32,3}, f::CuArray{Float32,3}, x_vec::Vec3, box_size::V3, box_min::V3, β::Float32,
                 damping_factor::Float32, niterations::Int; los = nothing)
    ctx = plan!(v, box_size, box_min)
    check(ccall((:baorec_mg_jacobi_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cfloat, Cfloat, Cint, Ptr{Cfloat}, Ptr{Cvoid}),
                ctx, ptr(v), ptr(f), size(v, 1), size(v, 2), size(v, 3), β, damping_factor, niterations, _los(los), stream()))
    v
end

function residual!(r::CuArray{Float32,3}, v::CuArray{Float32,3}, f::CuArray{Float32,3}, x_vec::Vec3, box_size::V3, box_min::V3,
                   β::Float32, damping_factor::Float32, niterations::Int; los = nothing)
    ctx = plan!(v, box_size, box_min)
    check(ccall((:baorec_mg_residual_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cfloat, Ptr{Cfloat}, Ptr{Cvoid}),
                ctx, ptr(r), ptr(v), ptr(f), size(v, 1), size(v, 2), size(v, 3), β, _los(los), stream()))
    r
end

function reduce!(v2h::CuArray{Float32,3}, v1h::CuArray{Float32,3})
    check(ccall((:baorec_mg_restrict_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}),
                context(), ptr(v2h), ptr(v1h), size(v1h, 1), size(v1h, 2), size(v1h, 3), stream()))
    v2h
end

function prolong!(v1h::CuArray{Float32,3}, v2h::CuArray{Float32,3})
    check(ccall((:baorec_mg_prolong_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}),
                context(), ptr(v1h), ptr(v2h), size(v1h, 1), size(v1h, 2), size(v1h, 3), stream()))
    v1h
end

function vcycle!(v::CuArray{Float32,3}, f::CuArray{Float32,3}, box_size::V3, box_min::V3, β::Float32,
                 damping_factor::Float32, niterations::Int; los = nothing)
    ctx = plan!(v, box_size, box_min)
    check(ccall((:baorec_mg_vcycle_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cfloat, Cfloat, Cint, Ptr{Cfloat}, Ptr{Cvoid}),
                ctx, ptr(v), ptr(f), β, damping_factor, niterations, _los(los), stream()))
    v
end

function fmg(f1h::CuArray{Float32,3}, v1h::Union{CuArray{Float32,3},Nothing}, box_size::V3, box_min::V3, β::Float32,
             jacobi_damping_factor::Float32, jacobi_niterations::Int, vcycle_niterations::Int; los = nothing)
    v1h === nothing && (v1h = CUDA.zeros(Float32, size(f1h)...))
    ctx = plan!(v1h, box_size, box_min)
    check(ccall((:baorec_mg_fmg_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cfloat, Cfloat, Cint, Cint, Ptr{Cfloat}, Ptr{Cvoid}),
                ctx, ptr(f1h), ptr(v1h), β, jacobi_damping_factor, jacobi_niterations, vcycle_niterations,
                _los(los), stream()))
    v1h
end

# ---- catalog pre/post-processing on the device (SURVEY.md 8f N1) ------------------------------------
# The lightcone examples define sky_to_cartesian / cartesian_to_sky / fkp_weights themselves and run
# them on CPU threads (examples/lightcone.jl:30-82); with CuVector columns these methods do the same
# work on the device.  The comoving-distance table replaces the `cache` of
# comoving_distance_interp / redshift_interp (src/cosmo.jl:86-103).
struct CosmologyParams            # struct baorec_cosmology
    h::Cdouble
    Omega_b0::Cdouble; Omega_c0::Cdouble; Omega_nu0::Cdouble; Omega_g0::Cdouble; Omega_k0::Cdouble; Omega_L0::Cdouble
    w0::Cdouble; wa::Cdouble
    z_tab_min::Cdouble; z_tab_max::Cdouble
    z_tab_num::Int64
end
CosmologyParams(c) = CosmologyParams(c.h, c.Ω_b₀, c.Ω_c₀, c.Ω_ν₀, c.Ω_γ₀, c.Ω_k₀, c.Ω_Λ₀, c.w0, c.wa,
                                     c.z_tab_min, c.z_tab_max, c.z_tab_num)

const BOUND_COSMO = Dict{Ptr{Cvoid},Any}()      # context -> the Cosmology whose table it holds
function bind_cosmology!(cosmo)
    ctx = context()
    if get(BOUND_COSMO, ctx, nothing) !== cosmo
        check(ccall((:baorec_cosmo_set, libbaorec), Cint, (Ptr{Cvoid}, Ref{CosmologyParams}), ctx, CosmologyParams(cosmo)))
        BOUND_COSMO[ctx] = cosmo
    end
    ctx
end

function sky_to_cartesian(ra::CuVector{Float32}, dec::CuVector{Float32}, red::CuVector{Float32}, cosmo)
    ctx = bind_cosmology!(cosmo)
    x, y, z = similar(ra), similar(ra), similar(ra)
    check(ccall((:baorec_sky_to_cartesian_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx, ptr(ra), ptr(dec), ptr(red), length(ra), Float32(cosmo.H₀ / 100), ptr(x), ptr(y), ptr(z), stream()))
    x, y, z                        # the examples' 3 x N matrix, as three columns
end

function cartesian_to_sky(x::CuVector{Float32}, y::CuVector{Float32}, z::CuVector{Float32}, cosmo)
    ctx = bind_cosmology!(cosmo)
    ra, dec, red = similar(x), similar(x), similar(x)
    check(ccall((:baorec_cartesian_to_sky_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx, ptr(x), ptr(y), ptr(z), length(x), Float32(cosmo.h), ptr(ra), ptr(dec), ptr(red), stream()))
    ra, dec, red
end

function fkp_weights(nz::CuVector{Float32}, P0)
    w = similar(nz)
    check(ccall((:baorec_fkp_weights_f32, libbaorec), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cfloat, Ptr{Cvoid}, Ptr{Cvoid}),
                context(), ptr(nz), length(nz), Float32(P0), ptr(w), stream()))
    w
end

# (pos + L) % L of test_helpers/simulation.py:38, in place, for a box starting at box_min
function wrap_positions!(x::CuVector{Float32}, y::CuVector{Float32}, z::CuVector{Float32}, box_size, box_min = (0f0, 0f0, 0f0))
    check(ccall((:baorec_wrap_positions_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                context(), ptr(x), ptr(y), ptr(z), length(x), f3(box_size), f3(box_min), stream()))
    x, y, z
end

# Engine options (include/baorec_b200.h: "fuse_kspace", "unified_sort", "deterministic_scatter", "scatter_pairs",
# "gather_stage", ...): BAOrecB200.set_option("deterministic_scatter", 1)
set_option(name::AbstractString, value::Integer) =
    check(ccall((:baorec_set_option, libbaorec), Cint, (Ptr{Cvoid}, Cstring, Int64), context(), name, value))

# P_0, P_2, P_4 of a density mesh (periodic box): the before / after check of test_helpers/simulation.py:56-75
# (pypowspec compute_auto_box; with `randoms`, the mesh of a shifted random catalog: compute_auto_box_rand) on the device.
# mas_power: 2 = CIC window, 3 = TSC, 0 = none.
function power_multipoles(rho::CuArray{Float32,3}, los = (0f0, 0f0, 1f0); kmin = 0.0, dk, nbins::Integer, mas_power::Integer = 2, shot = 0.0,
                          randoms::Union{CuArray{Float32,3},Nothing} = nothing)
    k, nmodes, p0, p2, p4 = (Vector{Float64}(undef, nbins) for _ in 1:5)
    check(ccall((:baorec_power_multipoles_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Cdouble, Cdouble, Cint, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                context(), ptr(rho), randoms === nothing ? C_NULL : ptr(randoms), f3(los), kmin, dk, nbins, mas_power, shot, k, nmodes, p0, p2, p4, stream()))
    (k = k, nmodes = nmodes, p0 = p0, p2 = p2, p4 = p4)
end

# The helpers' own configuration in ONE call (test_helpers/simulation.py:36-70 with test_helpers/powspec_auto.conf: TSC +
# interlaced grids): compute_auto_box(x, y, z, w, ...) / compute_auto_box_rand(x, y, z, w, rx, ry, rz, rw, ...) on device
# catalogs, on the grid and box of the last plan! (mas: 0 = CIC, 1 = TSC, 2 = PCS).  The catalogs are not modified.
function compute_auto_box(x::CuVector{Float32}, y::CuVector{Float32}, z::CuVector{Float32}, w::CuVector{Float32}, los = (0f0, 0f0, 1f0);
                          kmin = 0.0, dk, nbins::Integer, mas::Integer = 1, interlace::Bool = true, shot = 0.0,
                          randoms::Union{NTuple{4,CuVector{Float32}},Nothing} = nothing)
    k, nmodes, p0, p2, p4 = (Vector{Float64}(undef, nbins) for _ in 1:5)
    rx, ry, rz, rw = randoms === nothing ? (nothing, nothing, nothing, nothing) : randoms
    nr = randoms === nothing ? 0 : length(rx)
    check(ccall((:baorec_compute_auto_box_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cint,
                 Ptr{Cfloat}, Cdouble, Cdouble, Cint, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                context(), ptr(x), ptr(y), ptr(z), ptr(w), length(x), ptr(rx), ptr(ry), ptr(rz), ptr(rw), nr, mas, interlace ? 1 : 0,
                f3(los), kmin, dk, nbins, shot, k, nmodes, p0, p2, p4, stream()))
    (k = k, nmodes = nmodes, p0 = p0, p2 = p2, p4 = p4)
end

# Positions of the interlaced mesh (half a cell further along every axis, periodic) on the grid of the last plan!.
function interlace_positions(x::CuVector{Float32}, y::CuVector{Float32}, z::CuVector{Float32})
    ox, oy, oz = similar(x), similar(y), similar(z)
    check(ccall((:baorec_interlace_positions_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                context(), ptr(x), ptr(y), ptr(z), length(x), ptr(ox), ptr(oy), ptr(oz), stream()))
    (ox, oy, oz)
end

# The interlaced estimate from two pairs of meshes (rho / randoms painted from the catalog and from interlace_positions).
function power_multipoles_interlaced(rho::CuArray{Float32,3}, rho_shifted::CuArray{Float32,3}, los = (0f0, 0f0, 1f0); kmin = 0.0, dk,
                                     nbins::Integer, mas_power::Integer = 3, shot = 0.0,
                                     randoms::Union{CuArray{Float32,3},Nothing} = nothing,
                                     randoms_shifted::Union{CuArray{Float32,3},Nothing} = nothing)
    k, nmodes, p0, p2, p4 = (Vector{Float64}(undef, nbins) for _ in 1:5)
    check(ccall((:baorec_power_multipoles_interlaced_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Cdouble, Cdouble, Cint, Cint, Cdouble, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                context(), ptr(rho), ptr(randoms), ptr(rho_shifted), ptr(randoms_shifted), f3(los), kmin, dk, nbins, mas_power, shot,
                k, nmodes, p0, p2, p4, stream()))
    (k = k, nmodes = nmodes, p0 = p0, p2 = p2, p4 = p4)
end

# Many mocks per process (the reference's README: "one process, many reconstructions"): run! + reconstructed_positions
# (or read_shifts with positions = false) for every HOST catalog (x, y, z, w) of a periodic box, the PCIe transfers of
# neighbouring catalogs overlapping the solve.  Returns one (x, y, z) tuple of Vectors per catalog.
function run_batch(recon::AbstractRecon, grid_size::NTuple{3,Int}, catalogs::Vector{<:NTuple{4,Vector{Float32}}};
                   field = :disp, positions::Bool = true)
    ctx = context()
    check(ccall((:baorec_plan, libbaorec), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}),
                ctx, grid_size[1], grid_size[2], grid_size[3], f3(recon.box_size), f3(recon.box_min)))
    n = Int64[length(c[1]) for c in catalogs]
    out = [Tuple(Vector{Float32}(undef, k) for _ in 1:3) for k in n]
    col(arrs) = Ptr{Cvoid}[Ptr{Cvoid}(pointer(a)) for a in arrs]
    p = Ref(Params(recon))
    GC.@preserve catalogs out begin
        check(ccall((:baorec_batch_host_f32, libbaorec), Cint,
                    (Ptr{Cvoid}, Ptr{Params}, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64},
                     Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                    ctx, p, algorithm(recon), length(catalogs), col(c[1] for c in catalogs), col(c[2] for c in catalogs),
                    col(c[3] for c in catalogs), col(c[4] for c in catalogs), n, FIELD[field], positions ? 0 : 1,
                    col(o[1] for o in out), col(o[2] for o in out), col(o[3] for o in out)))
    end
    out
end

# ---- catalog files either side of the path (what the examples do with CSV.jl and NPZ.jl) --------------------------
# CSV.File(path, delim = ' ', ignorerepeated = true, header = [...], types = [Float32 ...]) followed by picking columns
# (examples/simulation.jl:12-15): the 1-based file columns `columns` as Vector{Float32}s, parsed by the library's threads.
function read_text_catalog(path::AbstractString, columns::Vector{Int}; delim::Char = ' ', nthreads::Integer = 0)
    n = Ref{Int64}(0)
    k = Ref{Cint}(0)
    check(ccall((:baorec_text_catalog_scan, libbaorec), Cint, (Cstring, Cchar, Ptr{Int64}, Ptr{Cint}, Cint),
                path, Cchar(delim), n, k, nthreads))
    out = [Vector{Float32}(undef, n[]) for _ in columns]
    ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(a)) for a in out]
    got = Ref{Int64}(0)
    GC.@preserve out begin
        check(ccall((:baorec_text_catalog_read_f32, libbaorec), Cint,
                    (Cstring, Cchar, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Int64}, Cint),
                    path, Cchar(delim), length(columns), Cint[c - 1 for c in columns], ptrs, n[], got, nthreads))
    end
    Tuple(out)
end

# npzread(path) of an (N, K) matrix, split into Float32 columns (`columns` 1-based; all of them by default).
function read_npy_catalog(path::AbstractString, columns::Union{Vector{Int},Nothing} = nothing; nthreads::Integer = 0)
    dtype = Ref{Cint}(0)
    forder = Ref{Cint}(0)
    n = Ref{Int64}(0)
    k = Ref{Int64}(0)
    check(ccall((:baorec_npy_info, libbaorec), Cint, (Cstring, Ptr{Cint}, Ptr{Cint}, Ptr{Int64}, Ptr{Int64}),
                path, dtype, forder, n, k))
    cols = columns === nothing ? collect(1:Int(k[])) : columns
    out = [Vector{Float32}(undef, n[]) for _ in cols]
    ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(a)) for a in out]
    got = Ref{Int64}(0)
    GC.@preserve out begin
        check(ccall((:baorec_npy_read_columns_f32, libbaorec), Cint,
                    (Cstring, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Int64}, Cint),
                    path, length(cols), Cint[c - 1 for c in cols], ptrs, n[], got, nthreads))
    end
    Tuple(out)
end

# npzwrite(path, hcat(cols...)) (examples/simulation.jl:38-40) without building the matrix.
function write_npy(path::AbstractString, cols::Vector{Float32}...)
    v = collect(cols)
    ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(a)) for a in v]
    GC.@preserve v begin
        check(ccall((:baorec_npy_write_columns_f32, libbaorec), Cint, (Cstring, Cint, Ptr{Ptr{Cvoid}}, Int64),
                    path, length(v), ptrs, length(v[1])))
    end
    nothing
end

# cat[map(z -> ((z > lo) & (z < hi)), cat.z), :] (examples/lightcone.jl:25-26) on SoA columns: the kept rows, in order.
function select_rows(cols::Vector{Vector{Float32}}, key::Integer, lo::Real, hi::Real; nthreads::Integer = 0)
    ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(a)) for a in cols]
    kept = Ref{Int64}(0)
    GC.@preserve cols begin
        check(ccall((:baorec_catalog_select_f32, libbaorec), Cint,
                    (Cint, Ptr{Ptr{Cvoid}}, Int64, Cint, Cfloat, Cfloat, Ptr{Int64}, Cint),
                    length(cols), ptrs, length(cols[1]), key - 1, lo, hi, kept, nthreads))
    end
    [resize!(a, kept[]) for a in cols]
end

# run_batch fed from catalog files and writing NPY files (examples/simulation.jl:12-40 looped over mocks): a reader thread
# parses the next catalogs into pinned memory and a writer thread stores the previous results while the device
# reconstructs the current one.  `columns` = 1-based file columns of x, y, z and the weights (0: weights of one).
function run_batch_files(recon::AbstractRecon, grid_size::NTuple{3,Int}, in_paths::Vector{String},
                         out_paths::Union{Vector{String},Nothing} = nothing; columns::NTuple{4,Int} = (1, 2, 3, 0),
                         delim::Char = ' ', field = :disp, positions::Bool = true, nthreads::Integer = 0)
    ctx = context()
    check(ccall((:baorec_plan, libbaorec), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}),
                ctx, grid_size[1], grid_size[2], grid_size[3], f3(recon.box_size), f3(recon.box_min)))
    p = Ref(Params(recon))
    rows = Vector{Int64}(undef, length(in_paths))
    seconds = Vector{Float64}(undef, 4)
    outs = out_paths === nothing ? C_NULL : out_paths
    check(ccall((:baorec_batch_files_f32, libbaorec), Cint,
                (Ptr{Cvoid}, Ptr{Params}, Cint, Cint, Ptr{Cstring}, Cchar, Ptr{Cint}, Cint, Cint, Ptr{Cstring}, Cint, Ptr{Int64},
                 Ptr{Cdouble}),
                ctx, p, algorithm(recon), length(in_paths), in_paths, Cchar(delim), Cint[c - 1 for c in columns], FIELD[field],
                positions ? 0 : 1, outs, nthreads, rows, seconds))
    (rows = rows, read_s = seconds[1], write_s = seconds[2], wait_s = seconds[3], total_s = seconds[4])
end

# run! needs no override: the reference's run! (src/recon.jl:134-261) allocates a CuArray mesh when
# data_x isa CuArray and calls setup_fft!, setup_box and reconstructed_*! -- all of which dispatch
# to the methods above.

end # module
