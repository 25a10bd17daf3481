# TemporalGPsB200.jl — the Julia side of the drop-in: a storage tag that routes the LGSSM hot path of
# TemporalGPs.jl (logpdf / _filter / posterior / marginals / marginals_diag on to_sde-wrapped GPs) to libtgpb200.so
# through `ccall`. Pure marshalling: models are still built by TemporalGPs' own `lgssm_components` on the host.
#
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia there). The same C ABI is exercised (a) from Python by
# temporalgps.jl_b200/_lib.py, which mirrors this file call for call, and (b) by tests/abi_c/abi_layout.c, a C program
# that hands the library buffers laid out exactly as `reinterpret(Float64, ::Vector{SMatrix})` / `Fill` produce them.
# See INTEGRATION.md.
#
#   using TemporalGPs, TemporalGPsB200
#   f  = to_sde(GP(Matern52Kernel()), B200Storage(Float64))
#   fx = f(RegularSpacing(0.0, 0.01, 10_000_000), 0.1)
#   logpdf(fx, y); marginals(posterior(fx, y)(x, 1e-2)); logpdf(posterior(fx, y)(x_pr, 0.1), y_pr)
module TemporalGPsB200

using TemporalGPs, AbstractGPs, StaticArrays, FillArrays, StructArrays, LinearAlgebra, Random
import TemporalGPs: StorageType, SArrayStorage, ArrayStorage, LGSSM, GaussMarkovModel, Forward, Reverse, Gaussian,
    ScalarOutputLGC, SmallOutputLGC, build_lgssm, LTISDE, ordering, x0, transitions, emissions,
    replace_observation_noise_cov, _filter, transform_model_and_obs, _logpdf_volume_compensation, marginals_diag
import AbstractGPs: logpdf, marginals, posterior

export B200Storage

const LIB = get(ENV, "TGP_B200_LIB", "libtgpb200.so")

"Storage tag (plug-in point of src/util/storage_types.jl:1, src/gp/lti_sde.jl:12-14)."
struct B200Storage{T<:Real} <: StorageType{T}
    device::Int
end
B200Storage(::Type{T}=Float64; device::Int=0) where {T} = B200Storage{T}(device)

# ---- handle -----------------------------------------------------------------------------------------
const HANDLES = Dict{Int,Ptr{Cvoid}}()
function handle(dev::Int)
    get!(HANDLES, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:tgp_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, dev)
        rc == 0 || error("tgp_create: ", unsafe_string(ccall((:tgp_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        h[]
    end
end
# Arithmetic of the large-state / vector-observation path follows the storage tag's element type (tgp_b200.h: TGP_OPT_DENSE_MATH):
# Float32 -> FP32 storage on the tcgen05 tensor cores (3xTF32), otherwise FP64. The small-state scan kernels are FP64 either way.
set_dense_math(h, f32::Bool) = ccall((:tgp_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), h, 6, f32 ? 1 : 0)

# status -> exception. TGP_ENOTPD carries the failing time index in its message ("... at time index N (0-based)"): the reference's
# `cholesky` throws PosDefException(info) with the index of the failing pivot; here `info` is the 1-based TIME index.
function check(h, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:tgp_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    if rc == 2
        m = match(r"time index (\d+)", msg)
        throw(PosDefException(m === nothing ? 0 : parse(Int, m.captures[1]) + 1))
    end
    error("libtgpb200 ($rc): $msg")                       # ErrorException, as lgssm.jl:202-208
end

# ---- descriptor (include/tgp_b200.h: tgp_lgssm) ------------------------------------------------------
struct Desc
    D::Int32; M::Int32; T::Int64; ordering::Int32; R_kind::Int32
    A::Ptr{Float64}; sA::Int64; a::Ptr{Float64}; sa::Int64; Q::Ptr{Float64}; sQ::Int64
    H::Ptr{Float64}; sH::Int64; h::Ptr{Float64}; sh::Int64; R::Ptr{Float64}; sR::Int64
    m0::Ptr{Float64}; P0::Ptr{Float64}
end

# One per-step ELEMENT as a flat column-major Float64 vector. Adjoint row vectors (the `h'` of ScalarOutputLGC, lti_sde.jl:88-101)
# and Diagonal noise matrices are unwrapped here, so every container method below is generic in the element.
elem(x::Real) = Float64[x]
elem(x::Adjoint{<:Any,<:AbstractVector}) = collect(Float64, parent(x))
elem(x::Diagonal) = collect(Float64, diag(x))
elem(x::AbstractArray) = collect(Float64, vec(x))          # column-major, as the ABI wants it

# A per-step array as (flat Float64 buffer, stride in elements): `Fill` -> ONE element, stride 0 (lti_sde.jl:148-160); any other
# vector -> its elements back to back. `Fill` is matched first, so there is no overlap with the AbstractVector method.
flat(x::Fill) = (elem(FillArrays.getindex_value(x)), 0)
function flat(x::AbstractVector)
    isempty(x) && return (Float64[], 0)
    n = length(elem(first(x)))
    buf = Vector{Float64}(undef, n * length(x))
    for (t, v) in enumerate(x)
        buf[(t-1)*n+1:t*n] = elem(v)
    end
    (buf, n)
end

r_kind(::Type{<:Real}) = 0                                  # TGP_R_SCALAR
r_kind(::Type{<:Diagonal}) = 1                              # TGP_R_DIAG
r_kind(::Type{<:AbstractMatrix}) = 2                        # TGP_R_DENSE

struct Marshalled
    bufs::Vector{Vector{Float64}}
    desc::Desc
end
function Marshalled(model::LGSSM)
    tr, em = transitions(model), emissions(model)
    (A, sA), (a, sa), (Q, sQ) = flat(tr.As), flat(tr.as), flat(tr.Qs)
    (H, sH), (h, sh), (R, sR) = flat(em.A), flat(em.a), flat(em.Q)   # StructArray fields .A, .a, .Q of the emission LGCs
    m0 = collect(Float64, x0(model).m); P0 = collect(Float64, vec(x0(model).P))
    D = length(m0)
    scalar = eltype(em) <: ScalarOutputLGC
    M = scalar ? 1 : length(elem(first(em.a)))
    kind = scalar ? 0 : r_kind(eltype(em.Q))
    ord = ordering(model) isa Forward ? 0 : 1
    d = Desc(D, M, length(model), ord, kind, pointer(A), sA, pointer(a), sa, pointer(Q), sQ,
             pointer(H), sH, pointer(h), sh, pointer(R), sR, pointer(m0), pointer(P0))
    Marshalled([A, a, Q, H, h, R, m0, P0], d)
end
obs_dim(mm::Marshalled) = Int(mm.desc.M)
# observations as the ABI takes them: T x M doubles, M fastest
yflat(y::AbstractVector{<:Real}) = convert(Vector{Float64}, y)
yflat(y::AbstractVector{<:AbstractVector}) = reduce(vcat, map(v -> convert(Vector{Float64}, v), y))

# ---- wrapper model: what build_lgssm returns for B200Storage -----------------------------------------
struct B200LGSSM{Tm<:LGSSM}
    model::Tm
    device::Int
    f32::Bool          # storage tag was B200Storage{Float32}: large-state products on the tensor cores (TGP_DENSE_TF32X3)
end
B200LGSSM(model, device::Int) = B200LGSSM(model, device, false)
function handle(m::B200LGSSM)       # handle of the model's device with the dense-path arithmetic of its storage tag selected
    h = handle(m.device)
    set_dense_math(h, m.f32)
    h
end
Base.length(m::B200LGSSM) = length(m.model)
ordering(m::B200LGSSM) = ordering(m.model)
emissions(m::B200LGSSM) = emissions(m.model)

# The ONLY method added to the reference's construction chain: build the components with the reference's own static-array code
# (scalar GPs) or dense-array code (space-time grids) by swapping the storage tag, then wrap. No `lgssm_components` method is
# defined here, so nothing can become ambiguous with the reference's (::Separable, ::SpaceTimeGrid, ::StorageType) method.
host_storage(f::LTISDE, x) = x isa TemporalGPs.RectilinearGrid ? ArrayStorage(Float64) : SArrayStorage(Float64)
function build_lgssm(f::LTISDE{<:GP,<:B200Storage}, x::AbstractVector, Σys::AbstractVector)
    # the ABI takes Float64 arrays; with a Float32 tag the library converts to FP32 storage on the device
    inner = build_lgssm(LTISDE(f.f, host_storage(f, x)), x, Σys)
    B200LGSSM(inner, f.storage.device, eltype(f.storage) === Float32)
end

check_lengths(m, y) = length(m) == length(y) ||
    error("Dimension mismatch. length(prior) is $(length(m)), but length(y) is $(length(y))")

# logpdf(model, y) — src/models/lgssm.jl:147-151
function logpdf(m::B200LGSSM, y::AbstractVector{<:Union{AbstractVector{<:Real},Real}})
    check_lengths(m, y)
    h = handle(m); mm = Marshalled(m.model); yy = yflat(y); out = Ref(0.0)
    GC.@preserve mm yy begin
        check(h, ccall((:tgp_logpdf, LIB), Cint, (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                       h, mm.desc, yy, out, C_NULL))
    end
    out[]
end
# missing data: host transform, kernels see plain Σ_t (src/models/missings.jl:8-13, 25-53)
function logpdf(m::B200LGSSM, y::AbstractVector{Union{Missing,T}}) where {T}
    model2, y2 = transform_model_and_obs(m.model, y)
    logpdf(B200LGSSM(model2, m.device, m.f32), y2) + _logpdf_volume_compensation(y, m.model)
end

# _filter(model, y) — src/models/lgssm.jl:171-173: the records (m, P) of Vector{Gaussian{SVector{D},SMatrix{D,D}}} written in place
function _filter(m::B200LGSSM, y::AbstractVector{<:Union{AbstractVector{<:Real},Real}})
    check_lengths(m, y)
    h = handle(m); mm = Marshalled(m.model); yy = yflat(y)
    D = Int(mm.desc.D); T = length(m); rec = D + D * D
    buf = Vector{Float64}(undef, rec * T)
    GC.@preserve mm yy buf begin
        check(h, ccall((:tgp_filter, LIB), Cint,
                       (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                       h, mm.desc, yy, pointer(buf), rec, pointer(buf) + 8D, rec, C_NULL))
    end
    [Gaussian(SVector{D}(buf[(t-1)*rec+1:(t-1)*rec+D]), SMatrix{D,D}(buf[(t-1)*rec+D+1:t*rec])) for t in 1:T]
end
function _filter(m::B200LGSSM, y::AbstractVector{Union{Missing,T}}) where {T}
    model2, y2 = transform_model_and_obs(m.model, y)
    _filter(B200LGSSM(model2, m.device, m.f32), y2)
end

# ---- posterior ---------------------------------------------------------------------------------------
# posterior(model, y) is LAZY: marginals(replace_observation_noise_cov(posterior(model, y), Σ)) — the chain of
# src/gp/posterior_lti_sde.jl:27-36 — is ONE library call (tgp_posterior_marginals) that never materialises (G, g, Σ);
# everything else (logpdf, rand, marginals with full covariances) goes through `materialise`, which runs tgp_posterior and
# returns the reference's own Reverse-ordered LGSSM (lgssm.jl:193-200), wrapped again so its calls come back to the library.
struct B200Posterior{Tm<:B200LGSSM,Ty,TΣ}
    prior::Tm
    y::Ty
    Σs_new::TΣ
end
posterior(m::B200LGSSM, y::AbstractVector) = (check_lengths(m, y); B200Posterior(m, y, nothing))
replace_observation_noise_cov(p::B200Posterior, Σs) = B200Posterior(p.prior, p.y, Σs)
replace_observation_noise_cov(m::B200LGSSM, Σs) = B200LGSSM(replace_observation_noise_cov(m.model, Σs), m.device, m.f32)
Base.length(p::B200Posterior) = length(p.prior)

observed(p::B200Posterior) = eltype(p.y) >: Missing ? transform_model_and_obs(p.prior.model, p.y) : (p.prior.model, p.y)

"The posterior as the reference's Reverse-ordered LGSSM (tgp_posterior: G, g, Σ per step and x0 = the last filtering distribution)."
function materialise(p::B200Posterior)
    m = p.prior; h = handle(m)
    model, y = observed(p)
    mm = Marshalled(model); yy = yflat(y)
    D = Int(mm.desc.D); T = length(m)
    G = Vector{Float64}(undef, D * D * T); g = Vector{Float64}(undef, D * T); S = Vector{Float64}(undef, D * D * T)
    mT = Vector{Float64}(undef, D); PT = Vector{Float64}(undef, D * D)
    GC.@preserve mm yy begin
        check(h, ccall((:tgp_posterior, LIB), Cint,
                       (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       h, mm.desc, yy, G, g, S, mT, PT))
    end
    As = [SMatrix{D,D}(G[(t-1)*D*D+1:t*D*D]) for t in 1:T]
    as = [SVector{D}(g[(t-1)*D+1:t*D]) for t in 1:T]
    Qs = [SMatrix{D,D}(S[(t-1)*D*D+1:t*D*D]) for t in 1:T]
    new_order = ordering(model) isa Forward ? Reverse() : Forward()
    ems = p.Σs_new === nothing ? emissions(model) : emissions(replace_observation_noise_cov(model, p.Σs_new))
    inner = LGSSM(GaussMarkovModel(new_order, As, as, Qs, Gaussian(SVector{D}(mT), SMatrix{D,D}(PT))), ems)
    B200LGSSM(inner, m.device, m.f32)
end

# logpdf(::FinitePosteriorLTISDE, y) and rand reach here through replace_observation_noise_cov(posterior(model, ys), Σ)
# (src/gp/posterior_lti_sde.jl:49-78)
logpdf(p::B200Posterior, y::AbstractVector) = logpdf(materialise(p), y)
_filter(p::B200Posterior, y::AbstractVector) = _filter(materialise(p), y)
Base.rand(rng::AbstractRNG, p::B200Posterior) = rand(rng, materialise(p).model)      # sampling stays on the host (lgssm.jl:65-91)

function marginals_diag(p::B200Posterior)
    m = p.prior; h = handle(m)
    model, y = observed(p)
    mm = Marshalled(model); yy = yflat(y); M = obs_dim(mm)
    (Rn, sRn) = flat(p.Σs_new === nothing ? emissions(model).Q : p.Σs_new)
    T = length(m); mu = Vector{Float64}(undef, T * M); v = Vector{Float64}(undef, T * M)
    GC.@preserve mm yy Rn begin
        check(h, ccall((:tgp_posterior_marginals, LIB), Cint,
                       (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       h, mm.desc, yy, Rn, sRn, mu, v, C_NULL))
    end
    M == 1 ? [Gaussian(mu[t], v[t]) for t in 1:T] :
             [Gaussian(mu[(t-1)*M+1:t*M], Diagonal(v[(t-1)*M+1:t*M])) for t in 1:T]
end
# scalar emissions: the marginal IS its diagonal (what posterior_lti_sde.jl:27-36 asks for); vector emissions: full covariances
marginals(p::B200Posterior) = obs_dim(Marshalled(observed(p)[1])) == 1 ? marginals_diag(p) : marginals(materialise(p))

# marginals(model) / marginals_diag(model) — src/models/lgssm.jl:99-141 (data-free)
function emit_marginals(m::B200LGSSM, diag::Bool)
    h = handle(m); mm = Marshalled(m.model); T = length(m); M = obs_dim(mm)
    nc = diag ? M : M * M
    mu = Vector{Float64}(undef, T * M); c = Vector{Float64}(undef, T * nc)
    GC.@preserve mm begin
        check(h, ccall((diag ? :tgp_marginals_diag : :tgp_marginals, LIB), Cint, (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}),
                       h, mm.desc, mu, c))
    end
    M == 1 && return [Gaussian(mu[t], c[t]) for t in 1:T]
    diag ? [Gaussian(mu[(t-1)*M+1:t*M], Diagonal(c[(t-1)*M+1:t*M])) for t in 1:T] :
           [Gaussian(mu[(t-1)*M+1:t*M], reshape(c[(t-1)*nc+1:t*nc], M, M)) for t in 1:T]
end
marginals(m::B200LGSSM) = emit_marginals(m, false)
marginals_diag(m::B200LGSSM) = emit_marginals(m, true)

# broadcast_components((F, q, H), x0, t::AbstractVector, storage) — src/gp/lti_sde.jl:136-147 — on the device: A[i] = exp(F * dt[i]),
# Q[i] = P - A[i] P A[i]' for an irregular grid, without T host matrix exponentials. F0: drift of the first transition (dt = 1) when
# the kernel is stretched in time (lti_sde.jl:361-373), else F. Returns what the reference returns (Vectors of SMatrix); A_out / Q_out
# may instead be device pointers (CuPtr from CUDA.jl) to keep the arrays resident and pass them on in a Desc.
function lti_components(F::AbstractMatrix{<:Real}, P::AbstractMatrix{<:Real}, t::AbstractVector{<:Real}; F0 = F, device::Int = 0)
    D = size(F, 1); T = length(t); h = handle(device)
    Fc = Matrix{Float64}(F); F0c = Matrix{Float64}(F0); Pc = Matrix{Float64}(P); tc = Vector{Float64}(t)
    A = Vector{Float64}(undef, T * D * D); Q = Vector{Float64}(undef, T * D * D)
    GC.@preserve Fc F0c Pc tc A Q begin
        check(h, ccall((:tgp_lti_components, LIB), Cint,
                       (Ptr{Cvoid}, Int32, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       h, D, T, Fc, F0c, Pc, tc, A, Q))
    end
    As = collect(reinterpret(SMatrix{D,D,Float64,D * D}, A)); Qs = collect(reinterpret(SMatrix{D,D,Float64,D * D}, Q))
    return As, Qs
end

end # module
