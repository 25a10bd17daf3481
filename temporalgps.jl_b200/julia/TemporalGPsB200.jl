# TemporalGPsB200.jl — the Julia side of the drop-in: a storage tag that routes the LGSSM hot path of
# TemporalGPs.jl (logpdf / _filter / posterior / marginals on to_sde-wrapped GPs) to libtgpb200.so through
# `ccall`. Pure marshalling: models are still built by TemporalGPs' own `lgssm_components` on the host.
#
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia there); the same C ABI is exercised from Python by
# temporalgps.jl_b200/_lib.py, which mirrors this file call for call. See INTEGRATION.md.
#
#   using TemporalGPs, TemporalGPsB200
#   f  = to_sde(GP(Matern52Kernel()), B200Storage(Float64))
#   fx = f(RegularSpacing(0.0, 0.01, 10_000_000), 0.1)
#   logpdf(fx, y); marginals(posterior(fx, y)(x, 1e-2))
module TemporalGPsB200

using TemporalGPs, AbstractGPs, StaticArrays, FillArrays, StructArrays, LinearAlgebra
import TemporalGPs: StorageType, SArrayStorage, LGSSM, GaussMarkovModel, Forward, Reverse, Gaussian,
    lgssm_components, build_lgssm, LTISDE, ordering, x0, transitions, emissions,
    replace_observation_noise_cov, _filter, transform_model_and_obs, _logpdf_volume_compensation
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
set_dense_math(h, ::Type{T}) where {T} =
    ccall((:tgp_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), h, 6, T === Float32 ? 1 : 0)
function check(h, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:tgp_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    rc == 2 && throw(PosDefException(0))                 # TGP_ENOTPD  <- cholesky failure in the reference
    error("libtgpb200 ($rc): $msg")                       # ErrorException, as lgssm.jl:202-208
end

# ---- descriptor (include/tgp_b200.h: tgp_lgssm) ------------------------------------------------------
struct Desc
    D::Int32; M::Int32; T::Int64; ordering::Int32; R_kind::Int32
    A::Ptr{Float64}; sA::Int64; a::Ptr{Float64}; sa::Int64; Q::Ptr{Float64}; sQ::Int64
    H::Ptr{Float64}; sH::Int64; h::Ptr{Float64}; sh::Int64; R::Ptr{Float64}; sR::Int64
    m0::Ptr{Float64}; P0::Ptr{Float64}
end

# A per-step array as (flat Float64 buffer, stride): `Fill` -> one element, stride 0 (lti_sde.jl:148-160);
# Vector{SMatrix}/Vector{SVector}/Vector{Float64} -> reinterpret, stride = length of one element.
flat(x::Fill) = (collect(Float64, vec(collect(FillArrays.getindex_value(x)))), 0)
flat(x::AbstractVector{<:Real}) = (convert(Vector{Float64}, x), 1)
flat(x::AbstractVector{<:StaticArray}) = (collect(reinterpret(Float64, x)), length(first(x)))
flat(x::AbstractVector{<:AbstractArray}) = (reduce(vcat, vec.(x)), length(first(x)))
flat(x::AbstractVector{<:Adjoint}) = flat(map(parent, x))       # emissions store H as adjoint vectors

struct Marshalled
    bufs::Vector{Vector{Float64}}
    desc::Desc
end
function Marshalled(model::LGSSM)
    tr, em = transitions(model), emissions(model)
    (A, sA), (a, sa), (Q, sQ) = flat(tr.As), flat(tr.as), flat(tr.Qs)
    (H, sH), (h, sh), (R, sR) = flat(em.A), flat(em.a), flat(em.Q)  # ScalarOutputLGC fields (lti_sde.jl:88-101)
    m0 = collect(Float64, x0(model).m); P0 = collect(Float64, vec(x0(model).P))
    D = length(m0)
    ord = ordering(model) isa Forward ? 0 : 1
    d = Desc(D, 1, length(model), ord, 0, pointer(A), sA, pointer(a), sa, pointer(Q), sQ,
             pointer(H), sH, pointer(h), sh, pointer(R), sR, pointer(m0), pointer(P0))
    Marshalled([A, a, Q, H, h, R, m0, P0], d)
end

# ---- wrapper model: what build_lgssm returns for B200Storage -----------------------------------------
struct B200LGSSM{Tm<:LGSSM}
    model::Tm
    device::Int
    f32::Bool          # storage tag was B200Storage{Float32}: large-state products on the tensor cores (TGP_DENSE_TF32X3)
end
B200LGSSM(model, device::Int) = B200LGSSM(model, device, false)
# handle of the model's device with the dense-path arithmetic of its storage tag selected
function handle(m::B200LGSSM)
    h = handle(m.device)
    set_dense_math(h, m.f32 ? Float32 : Float64)
    h
end
Base.length(m::B200LGSSM) = length(m.model)

lgssm_components(k, t::AbstractVector, s::B200Storage{T}) where {T} = lgssm_components(k, t, SArrayStorage(T))
function build_lgssm(f::LTISDE{<:GP,<:B200Storage}, x::AbstractVector, Σys::AbstractVector)
    # the ABI takes Float64 arrays; with a Float32 tag the library converts to FP32 storage on the device
    inner = build_lgssm(LTISDE(f.f, SArrayStorage(Float64)), x, Σys)
    B200LGSSM(inner, f.storage.device, eltype(f.storage) === Float32)
end

# logpdf(model, y) — src/models/lgssm.jl:147-151
function logpdf(m::B200LGSSM, y::AbstractVector{<:Real})
    length(m) == length(y) || error("Dimension mismatch. length(prior) is $(length(m)), but length(y) is $(length(y))")
    h = handle(m); mm = Marshalled(m.model); yy = convert(Vector{Float64}, y); out = Ref(0.0)
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

# _filter(model, y) — src/models/lgssm.jl:171-173: Vector{Gaussian{SVector{D},SMatrix{D,D}}} written in place
function _filter(m::B200LGSSM, y::AbstractVector{<:Real})
    h = handle(m); mm = Marshalled(m.model); yy = convert(Vector{Float64}, y)
    D = Int(mm.desc.D); T = length(m); rec = D + D * D
    buf = Vector{Float64}(undef, rec * T)
    GC.@preserve mm yy buf begin
        check(h, ccall((:tgp_filter, LIB), Cint,
                       (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                       h, mm.desc, yy, pointer(buf), rec, pointer(buf) + 8D, rec, C_NULL))
    end
    [Gaussian(SVector{D}(buf[(t-1)*rec+1:(t-1)*rec+D]), SMatrix{D,D}(buf[(t-1)*rec+D+1:t*rec])) for t in 1:T]
end

# posterior(model, y) — lazy: marginals(replace_observation_noise_cov(posterior(model, y), Σ)) is ONE library
# call (tgp_posterior_marginals; src/gp/posterior_lti_sde.jl:27-36); the materialised Reverse LGSSM of
# lgssm.jl:193-200 is available through `materialise` (tgp_posterior).
struct B200Posterior{Tm<:B200LGSSM,Ty,TΣ}
    prior::Tm
    y::Ty
    Σs_new::TΣ
end
posterior(m::B200LGSSM, y::AbstractVector) = B200Posterior(m, y, nothing)
replace_observation_noise_cov(p::B200Posterior, Σs) = B200Posterior(p.prior, p.y, Σs)

function marginals(p::B200Posterior)
    m = p.prior; h = handle(m.device)
    model, y = eltype(p.y) >: Missing ? transform_model_and_obs(m.model, p.y) : (m.model, p.y)
    mm = Marshalled(model); yy = convert(Vector{Float64}, y)
    (Rn, sRn) = flat(p.Σs_new === nothing ? emissions(model).Q : p.Σs_new)
    T = length(m); mu = Vector{Float64}(undef, T); v = Vector{Float64}(undef, T)
    GC.@preserve mm yy Rn begin
        check(h, ccall((:tgp_posterior_marginals, LIB), Cint,
                       (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       h, mm.desc, yy, Rn, sRn, mu, v, C_NULL))
    end
    [Gaussian(mu[t], v[t]) for t in 1:T]
end

# marginals(model) — src/models/lgssm.jl:99-101 (data-free)
function marginals(m::B200LGSSM)
    h = handle(m.device); mm = Marshalled(m.model); T = length(m)
    mu = Vector{Float64}(undef, T); v = Vector{Float64}(undef, T)
    GC.@preserve mm begin
        check(h, ccall((:tgp_marginals, LIB), Cint, (Ptr{Cvoid}, Ref{Desc}, Ptr{Float64}, Ptr{Float64}), h, mm.desc, mu, v))
    end
    [Gaussian(mu[t], v[t]) for t in 1:T]
end

end # module
