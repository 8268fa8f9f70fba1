# NonuniformFFTsB200.jl — thin Julia host shim over libnufft_b200.so (C ABI: include/nufft_b200.h).
#
# This is the binding a NonuniformFFTs.jl maintainer would add so that `backend = B200Backend()` routes
# PlanNUFFT / set_points! / exec_type1! / exec_type2! to the B200-native library through `ccall`
# (no CUDA.jl / KernelAbstractions dispatch).  It is a literal 1:1 over the C header.  There is no Julia
# in the build image, so this file is NOT executed by the test-suite; the same ABI is exercised from
# Python ctypes (nonuniformffts.jl_b200/_lib.py), struct layout included (tests/test_abi_symbols.py).
#
# Device memory: the shim takes raw device pointers (`CuPtr`-like `Ptr{Cvoid}`); with CUDA.jl loaded one
# passes `pointer(::CuArray)`; without it, `B200Buffer` below wraps cudaMalloc/cudaMemcpy via libcudart.
module NonuniformFFTsB200

export B200Plan, set_points!, exec_type1!, exec_type2!, HalfSupport,
       KaiserBesselKernel, BackwardsKaiserBesselKernel, GaussianKernel, BSplineKernel, ESKernel

const libnufft = get(ENV, "NUFFT_B200_LIB", joinpath(@__DIR__, "..", "libnufft_b200.so"))

# ---- error codes -> the reference's exception types (src/plan.jl:545-556, src/NonuniformFFTs.jl:92-114) ----
const NUFFT_SUCCESS = Cint(0)
function check(rc::Cint)
    rc == NUFFT_SUCCESS && return nothing
    msg = unsafe_string(ccall((:nufft_last_error, libnufft), Cstring, ()))
    rc == -1 && throw(ArgumentError(msg))          # NUFFT_ERR_ARG
    rc == -2 && throw(DimensionMismatch(msg))      # NUFFT_ERR_DIM
    rc == -3 && throw(ArgumentError(msg))          # NUFFT_ERR_UNSUPPORTED
    error("libnufft_b200 [$rc]: $msg")
end

struct HalfSupport{M} end
HalfSupport(M::Integer) = HalfSupport{Int(M)}()

abstract type AbstractKernel end
struct KaiserBesselKernel <: AbstractKernel; β::Float64; end
struct BackwardsKaiserBesselKernel <: AbstractKernel; β::Float64; end
struct GaussianKernel <: AbstractKernel; ℓ::Float64; end
struct BSplineKernel <: AbstractKernel end
struct ESKernel <: AbstractKernel; β::Float64; end     # not in the reference: exp(β (sqrt(1 - y^2) - 1)), FastApproximation only
KaiserBesselKernel() = KaiserBesselKernel(NaN)
BackwardsKaiserBesselKernel() = BackwardsKaiserBesselKernel(NaN)
GaussianKernel() = GaussianKernel(NaN)
ESKernel() = ESKernel(NaN)
kernel_id(::KaiserBesselKernel) = Cint(0); kernel_id(::BackwardsKaiserBesselKernel) = Cint(1)
kernel_id(::GaussianKernel) = Cint(2);     kernel_id(::BSplineKernel) = Cint(3); kernel_id(::ESKernel) = Cint(4)
kernel_param(k::Union{KaiserBesselKernel, BackwardsKaiserBesselKernel, ESKernel}) = k.β
kernel_param(k::GaussianKernel) = k.ℓ
kernel_param(::BSplineKernel) = NaN

# mirror of `nufft_opts` (field order and types exactly as in include/nufft_b200.h)
Base.@kwdef mutable struct NufftOpts
    struct_size::UInt32 = 0
    dim::Int32 = 1
    n_modes::NTuple{3, Int64} = (1, 1, 1)
    is_complex::Int32 = 1
    dtype::Int32 = 1
    half_support::Int32 = 4
    sigma::Float64 = 2.0
    kernel::Int32 = 0
    kernel_param::Float64 = NaN
    eval_mode::Int32 = 1
    ntransforms::Int32 = 1
    fftshift::Int32 = 0
    sort_points::Int32 = 0
    gpu_method::Int32 = 0
    block_dims::NTuple{3, Int64} = (0, 0, 0)
    point_convention::Int32 = 0
    device::Int32 = -1
    stream::Ptr{Cvoid} = C_NULL
    record_timings::Int32 = 0
    spread_chunk::Int32 = 0
end

mutable struct NufftCallbacks
    struct_size::UInt32
    nu_weights::Ptr{Cvoid}
    u_factor_sep::Ptr{Ptr{Cvoid}}
    u_factor_dense::Ptr{Cvoid}
    nvrtc_src::Cstring          # optional CUDA C++ source of general callbacks (compiled with NVRTC once per plan)
    user_data::Ptr{Cvoid}       # device pointer handed to them
end

"""
    B200Plan(Z, dims; m = HalfSupport(4), σ = 2, kernel = KaiserBesselKernel(), ntransforms = 1,
             fftshift = false, kernel_evalmode = :direct, gpu_method = :auto, synchronise = false, stream = C_NULL)

Same keyword arguments as `PlanNUFFT` (src/plan.jl:467-599).
"""
mutable struct B200Plan{Z <: Number, N}
    handle::Ptr{Cvoid}
    dims::NTuple{N, Int}          # size(plan)
    ntransforms::Int
    np::Int
    points                          # keeps the point arrays alive, like points_ref[] (src/set_points.jl:45)
end

function B200Plan(::Type{Z}, dims::NTuple{N, Integer}; m = HalfSupport(4), σ::Real = 2, kernel::AbstractKernel = KaiserBesselKernel(),
        ntransforms::Integer = 1, fftshift::Bool = false, kernel_evalmode::Symbol = :direct, gpu_method::Symbol = :auto,
        block_size = nothing, stream::Ptr{Cvoid} = C_NULL, device::Integer = -1, timer::Bool = false,
        point_convention::Integer = 0) where {Z <: Number, N}
    M = m isa HalfSupport ? typeof(m).parameters[1] : Int(m)
    T = real(Z)
    o = NufftOpts()
    ccall((:nufft_opts_default, libnufft), Cint, (Ref{NufftOpts},), o) |> check
    o.dim = N
    o.n_modes = ntuple(d -> d <= N ? Int64(dims[d]) : Int64(1), 3)
    o.is_complex = Z <: Complex ? 1 : 0
    o.dtype = T === Float64 ? 1 : 0
    o.half_support = M
    o.sigma = σ
    o.kernel = kernel_id(kernel)
    o.kernel_param = kernel_param(kernel)
    o.eval_mode = kernel_evalmode === :direct ? 1 : 0
    o.ntransforms = ntransforms
    o.fftshift = fftshift
    o.gpu_method = gpu_method === :global_memory ? 1 : gpu_method === :shared_memory ? 2 :
                   gpu_method === :auto ? 0 : throw(ArgumentError("expected gpu_method ∈ (:global_memory, :shared_memory)"))
    block_size isa Tuple && (o.block_dims = ntuple(d -> d <= N ? Int64(block_size[d]) : Int64(0), 3))
    o.point_convention = point_convention
    o.device = device
    o.stream = stream
    o.record_timings = timer
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:nufft_plan_create, libnufft), Cint, (Ref{Ptr{Cvoid}}, Ref{NufftOpts}), h, o) |> check
    sz = zeros(Int64, 3); os = zeros(Int64, 3); nt = Ref{Int32}(0)
    ccall((:nufft_plan_shape, libnufft), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Int32}), h[], sz, os, nt) |> check
    p = B200Plan{Z, N}(h[], ntuple(d -> Int(sz[d]), N), Int(nt[]), -1, nothing)
    finalizer(p -> ccall((:nufft_plan_destroy, libnufft), Cint, (Ptr{Cvoid},), p.handle), p)
    p
end
B200Plan(::Type{Z}, n::Integer; kws...) where {Z} = B200Plan(Z, (n,); kws...)

Base.size(p::B200Plan) = p.dims
Base.ndims(::B200Plan{Z, N}) where {Z, N} = N
Base.eltype(::B200Plan{Z}) where {Z} = complex(Z)
ntransforms(p::B200Plan) = p.ntransforms

devptr(a) = Ptr{Cvoid}(UInt(pointer(a)))     # works for CuArray (CuPtr) and for B200Buffer

"""    set_points!(p, (xs, ys, zs))   — src/set_points.jl:33-52.  Arrays must be device arrays of real(Z)."""
function set_points!(p::B200Plan{Z, N}, xp::NTuple{N, Any}) where {Z, N}
    T = real(Z)
    all(x -> eltype(x) === T, xp) || throw(ArgumentError("input points must have the same accuracy as the created plan"))
    np = length(xp[1])
    all(x -> length(x) == np, xp) || throw(DimensionMismatch("input points must have the same length along all dimensions"))
    ptrs = Ptr{Cvoid}[devptr(x) for x in xp]
    ccall((:nufft_set_points, libnufft), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), p.handle, np, ptrs) |> check
    p.points = xp
    p.np = np
    p
end
set_points!(p::B200Plan{Z, 1}, xp::AbstractVector{<:Real}) where {Z} = set_points!(p, (xp,))

"""    set_points!(p, xp::AbstractMatrix)   — src/set_points.jl:76-88 (size(xp) == (N, Np)); also serves a
`Vector{SVector{N}}` through `reinterpret(reshape, T, xp)` (src/set_points.jl:62-74).  The reference copies the
matrix into N vectors; here the array-of-points layout is read in place by the binning kernel."""
function set_points!(p::B200Plan{Z, N}, xp::AbstractMatrix{<:Real}) where {Z, N}
    size(xp, 1) == N || throw(DimensionMismatch(lazy"expected input matrix to have dimensions ($N, Np)"))
    eltype(xp) === real(Z) || throw(ArgumentError("input points must have the same accuracy as the created plan"))
    ccall((:nufft_set_points_matrix, libnufft), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}), p.handle, size(xp, 2), devptr(xp)) |> check
    p.points = xp
    p.np = size(xp, 2)        # exec_type1! / exec_type2! validate the data vectors against it
    p
end

# Returns (Ref{NufftCallbacks} | C_NULL, roots): `roots` holds every Julia object the struct points into (the source string
# behind the Cstring, the device arrays) and must be GC.@preserve'd together with the Ref for the duration of the ccall.
function _callbacks(cb)
    cb === nothing && return (C_NULL, ())
    # cb = (nonuniform = weights::DeviceVector | nothing, uniform = factor::DeviceArray | nothing,
    #       source = CUDA C++ source of general callbacks | nothing, user_data = DeviceArray | nothing)
    src = get(cb, :source, nothing)
    src = src === nothing ? nothing : String(src)
    usr = get(cb, :user_data, nothing)
    r = Ref(NufftCallbacks(UInt32(sizeof(NufftCallbacks)),
        cb.nonuniform === nothing ? C_NULL : devptr(cb.nonuniform), C_NULL,
        cb.uniform === nothing ? C_NULL : devptr(cb.uniform),
        src === nothing ? Cstring(C_NULL) : Base.unsafe_convert(Cstring, src),
        usr === nothing ? C_NULL : devptr(usr)))
    (r, (src, usr, cb.nonuniform, cb.uniform))
end

"""    exec_type1!(ûs, p, vp; callbacks)   — src/NonuniformFFTs.jl:148-195"""
function exec_type1!(us::NTuple{C, Any}, p::B200Plan{Z}, vp::NTuple{C, Any}; callbacks = nothing) where {Z, C}
    C == p.ntransforms || throw(DimensionMismatch("wrong amount of arrays (expected a tuple of $(p.ntransforms) arrays)"))
    all(u -> eltype(u) === complex(Z), us) || throw(ArgumentError("uniform data must have the same accuracy as the created plan"))
    all(u -> size(u) == size(p), us) || throw(DimensionMismatch("wrong dimensions of array (expected dimensions $(size(p)))"))
    all(v -> length(v) == p.np, vp) || throw(DimensionMismatch("wrong length of data vector (it should match the number of points $(p.np))"))
    up = Ptr{Cvoid}[devptr(u) for u in us]; vpp = Ptr{Cvoid}[devptr(v) for v in vp]
    cb, roots = _callbacks(callbacks)
    GC.@preserve cb roots us vp ccall((:nufft_exec_type1, libnufft), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                          p.handle, up, vpp, cb === C_NULL ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, cb)) |> check
    us
end
exec_type1!(us, p::B200Plan, vp; kws...) = (exec_type1!((us,), p, (vp,); kws...); us)

"""    exec_type2!(vp, p, ûs; callbacks)   — src/NonuniformFFTs.jl:237-291"""
function exec_type2!(vp::NTuple{C, Any}, p::B200Plan{Z}, us::NTuple{C, Any}; callbacks = nothing) where {Z, C}
    C == p.ntransforms || throw(DimensionMismatch("wrong amount of data vectors (expected a tuple of $(p.ntransforms) vectors)"))
    all(u -> eltype(u) === complex(Z), us) || throw(ArgumentError("uniform data must have the same accuracy as the created plan"))
    all(u -> size(u) == size(p), us) || throw(DimensionMismatch("wrong dimensions of array (expected dimensions $(size(p)))"))
    all(v -> length(v) == p.np, vp) || throw(DimensionMismatch("wrong length of data vector (it should match the number of points $(p.np))"))
    up = Ptr{Cvoid}[devptr(u) for u in us]; vpp = Ptr{Cvoid}[devptr(v) for v in vp]
    cb, roots = _callbacks(callbacks)
    GC.@preserve cb roots us vp ccall((:nufft_exec_type2, libnufft), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}),
                          p.handle, vpp, up, cb === C_NULL ? C_NULL : Base.unsafe_convert(Ptr{Cvoid}, cb)) |> check
    vp
end
exec_type2!(vp, p::B200Plan, us; kws...) = (exec_type2!((vp,), p, (us,); kws...); vp)

# ---- minimal device buffer for hosts without CUDA.jl (cudaMalloc / cudaMemcpy through libcudart) ----
const libcudart = get(ENV, "NUFFT_B200_CUDART", "libcudart.so")
mutable struct B200Buffer{T, N} <: AbstractArray{T, N}
    ptr::Ptr{Cvoid}
    dims::NTuple{N, Int}
end
function B200Buffer{T}(dims::Integer...) where {T}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:cudaMalloc, libcudart), Cint, (Ref{Ptr{Cvoid}}, Csize_t), r, prod(dims) * sizeof(T)) == 0 || error("cudaMalloc failed")
    b = B200Buffer{T, length(dims)}(r[], map(Int, dims))
    finalizer(b -> ccall((:cudaFree, libcudart), Cint, (Ptr{Cvoid},), b.ptr), b)
end
Base.size(b::B200Buffer) = b.dims
Base.pointer(b::B200Buffer) = b.ptr
upload!(b::B200Buffer{T}, a::Array{T}) where {T} =
    (ccall((:cudaMemcpy, libcudart), Cint, (Ptr{Cvoid}, Ptr{T}, Csize_t, Cint), b.ptr, a, sizeof(a), 1) == 0 || error("cudaMemcpy H2D failed"); b)
download!(a::Array{T}, b::B200Buffer{T}) where {T} =
    (ccall((:cudaMemcpy, libcudart), Cint, (Ptr{T}, Ptr{Cvoid}, Csize_t, Cint), a, b.ptr, sizeof(a), 2) == 0 || error("cudaMemcpy D2H failed"); a)

end # module
