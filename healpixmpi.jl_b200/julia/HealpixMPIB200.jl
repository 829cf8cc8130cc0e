# HealpixMPIB200.jl — the Julia side of the drop-in: HealpixMPI.jl keeps its API (DMap, DAlm,
# MPI.Scatter!/Gather!, alm2map!, adjoint_alm2map!, strategy.jl's RR ownership) and `ccall`s into
# libhpsht_b200.so instead of Ducc0.jl.  NOT EXECUTED in the build image (no Julia there); it is the
# binding a maintainer adds next to src/sht.jl.  Every ccall mirrors a prototype of
# include/hpsht_b200.h; the Python/ctypes binding (healpixmpi.jl_b200/_lib.py) exercises the very
# same symbols and is what the test-suite runs.
#
# Two integration levels:
#   (A) one-line substitution inside src/sht.jl: replace `Ducc0.Sht.alm2leg!` etc. by
#       `HealpixMPIB200.Sht.alm2leg!` (same argument order as src/sht.jl:85,93,169,178), keep
#       communicate_alm2map!/communicate_map2alm! and MPI.Alltoallv! as they are.
#   (B) fused: `alm2map!(d_alm, d_map)` / `adjoint_alm2map!(d_map, d_alm)` (2-argument methods,
#       src/sht.jl:96-100,181-185) call one plan that owns the leg buffers and exchanges over NCCL.
module HealpixMPIB200

using MPI
import Healpix

const lib = get(ENV, "HPSHT_B200_LIB", "libhpsht_b200.so")

function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:hpsht_last_error, lib), Cstring, ()))
    # HPSHT_ERR_ARG (1) and HPSHT_ERR_EMPTY_RANK (5) are the reference's DomainErrors
    (status == 1 || status == 5) ? throw(DomainError(status, msg)) : error("hpsht error $status: $msg")
end

# ------------------------------------------------------------------------------------------------
# (A) stage level: same signatures as Ducc0.Sht.{alm2leg!,leg2alm!,leg2map!,map2leg!}
# ------------------------------------------------------------------------------------------------
module Sht
import ..lib, ..check

# alm::(nalm, ncomp) ComplexF64, leg::(nm, nring, ncomp) ComplexF64; mstart is Julia's 1-based
# info.mstart (src/alm.jl:141-151): the C side is 0-based, so subtract 1 here like Ducc0.jl does.
function alm2leg!(alm::StridedMatrix{ComplexF64}, leg::StridedArray{ComplexF64,3}, spin::Integer,
                  lmax::Integer, mval::Vector{Csize_t}, mstart::Vector{Cptrdiff_t}, lstride::Integer,
                  theta::Vector{Float64}, nthreads::Integer)
    m64 = Int64.(mval); s64 = Int64.(mstart) .- 1
    check(ccall((:hpsht_alm2leg, lib), Cint,
                (Ptr{Float64}, Int64, Ptr{Float64}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64,
                 Ptr{Float64}, Cint),
                alm, stride(alm, 2), leg, spin, lmax, length(m64), m64, s64, lstride, length(theta), theta,
                nthreads))
    leg
end

function leg2alm!(leg::StridedArray{ComplexF64,3}, alm::StridedMatrix{ComplexF64}, spin::Integer,
                  lmax::Integer, mval::Vector{Csize_t}, mstart::Vector{Cptrdiff_t}, lstride::Integer,
                  theta::Vector{Float64}, nthreads::Integer)
    m64 = Int64.(mval); s64 = Int64.(mstart) .- 1
    check(ccall((:hpsht_leg2alm, lib), Cint,
                (Ptr{Float64}, Ptr{Float64}, Int64, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64,
                 Ptr{Float64}, Cint),
                leg, alm, stride(alm, 2), spin, lmax, length(m64), m64, s64, lstride, length(theta), theta,
                nthreads))
    alm
end

# leg::(nm, nring_loc, ncomp), map::(npix_loc, ncomp); ringstart is Julia's 1-based info.rstart
function leg2map!(leg::StridedArray{ComplexF64,3}, map::StridedMatrix{Float64}, nphi::Vector{Csize_t},
                  phi0::Vector{Float64}, ringstart::Vector{Csize_t}, pixstride::Integer, nthreads::Integer)
    n64 = Int64.(nphi); r64 = Int64.(ringstart) .- 1
    check(ccall((:hpsht_leg2map, lib), Cint,
                (Ptr{Float64}, Ptr{Float64}, Int64, Cint, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64},
                 Int64, Cint),
                leg, map, stride(map, 2), size(map, 2), size(leg, 1), length(n64), n64, phi0, r64, pixstride,
                nthreads))
    map
end

function map2leg!(map::StridedMatrix{Float64}, leg::StridedArray{ComplexF64,3}, nphi::Vector{Csize_t},
                  phi0::Vector{Float64}, ringstart::Vector{Csize_t}, pixstride::Integer, nthreads::Integer)
    n64 = Int64.(nphi); r64 = Int64.(ringstart) .- 1
    check(ccall((:hpsht_map2leg, lib), Cint,
                (Ptr{Float64}, Int64, Ptr{Float64}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64},
                 Int64, Cint),
                map, stride(map, 2), leg, size(map, 2), size(leg, 1), length(n64), n64, phi0, r64, pixstride,
                nthreads))
    leg
end
end # module Sht

# ------------------------------------------------------------------------------------------------
# (B) fused plan
# ------------------------------------------------------------------------------------------------
mutable struct Plan
    handle::Ptr{Cvoid}
    function Plan(nside::Integer, lmax::Integer, comm::MPI.Comm; max_ncomp::Integer = 2, device::Integer = -1,
                  thetatot = nothing, phi0 = nothing)
        rank, nranks = MPI.Comm_rank(comm), MPI.Comm_size(comm)
        id = zeros(UInt8, 128)
        if nranks > 1
            rank == 0 && check(ccall((:hpsht_nccl_unique_id, lib), Cint, (Ptr{UInt8},), id))
            MPI.Bcast!(id, 0, comm)               # ships the NCCL unique id (HPSHT_NCCL_ID_BYTES)
        end
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:hpsht_plan_create, lib), Cint,
                    (Ref{Ptr{Cvoid}}, Int64, Int64, Cint, Cint, Ptr{UInt8}, Cint, Cint),
                    h, nside, lmax, rank, nranks, nranks > 1 ? pointer(id) : C_NULL, device, max_ncomp))
        p = new(h[])
        if thetatot !== nothing || phi0 !== nothing   # bit-identical geometry to d_map.info
            check(ccall((:hpsht_plan_set_geometry, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                        p.handle, thetatot === nothing ? C_NULL : thetatot, phi0 === nothing ? C_NULL : phi0))
        end
        finalizer(x -> ccall((:hpsht_plan_destroy, lib), Cint, (Ptr{Cvoid},), x.handle), p)
    end
end

# The geometry a plan was built with is part of the cache key (a DMap with another thetatot / phi0 must not
# reuse a stale plan).
const PLANS = Dict{Tuple{Int,Int,MPI.Comm,UInt64},Plan}()
function plan_for(nside, lmax, comm; thetatot = nothing, phi0 = nothing, kw...)
    key = (Int(nside), Int(lmax), comm, hash((thetatot, phi0)))
    get!(() -> Plan(nside, lmax, comm; thetatot = thetatot, phi0 = phi0, kw...), PLANS, key)
end

# The raw pointers handed to the library must describe exactly the shard the plan derives from
# (nside, lmax, rank, nranks): compare sizes and descriptor arrays, throw DomainError otherwise
# (src/sht.jl:76-77 style) instead of letting the device read or write out of bounds.
function check_shards(p::Plan, d_alm, d_map)
    s = [Ref{Int64}(0) for i in 1:5]
    check(ccall((:hpsht_plan_sizes, lib), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
                p.handle, s[1], s[2], s[3], s[4], s[5]))
    nm, nalm, nr, npix = s[1][], s[2][], s[3][], s[4][]
    (d_alm.info.mmax == d_alm.info.lmax) || throw(DomainError(d_alm.info.mmax, "mmax must equal lmax"))
    (size(d_alm.alm, 1) == nalm && length(d_alm.info.mval) == nm) ||
        throw(DomainError(size(d_alm.alm, 1), "DAlm shard does not match the plan (nalm_loc = $nalm, nm_loc = $nm)"))
    (size(d_map.pixels, 1) == npix && length(d_map.info.rings) == nr) ||
        throw(DomainError(size(d_map.pixels, 1), "DMap shard does not match the plan (npix_loc = $npix, nring_loc = $nr)"))
    mval = Vector{Int64}(undef, nm); mstart = Vector{Int64}(undef, nm)
    check(ccall((:hpsht_plan_alm_info, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), p.handle, mval, mstart))
    (mval == d_alm.info.mval && mstart .+ 1 == d_alm.info.mstart) ||
        throw(DomainError(0, "DAlm.info.mval / mstart differ from the plan's round-robin shard"))
    rings = Vector{Int64}(undef, nr); rstart = Vector{Int64}(undef, nr)
    check(ccall((:hpsht_plan_map_info, lib), Cint,
                (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                p.handle, rings, rstart, C_NULL, C_NULL, C_NULL, C_NULL))
    (rings == d_map.info.rings && rstart .+ 1 == d_map.info.rstart) ||
        throw(DomainError(0, "DMap.info.rings / rstart differ from the plan's round-robin shard"))
    nothing
end

# Methods to add in src/sht.jl in place of the allocating wrappers (src/sht.jl:96-100, 181-185).
# d_alm.alm::(nalm_loc, ncomp) ComplexF64 and d_map.pixels::(npix_loc, ncomp) Float64 are borrowed
# for the duration of the call; outputs are overwritten in place.
function alm2map_fused!(d_alm, d_map; nthreads::Integer = 0)
    (size(d_alm.alm, 2) == size(d_map.pixels, 2)) || throw(DomainError(0, "Number of components must match."))
    comm = (d_alm.info.comm == d_map.info.comm) ? d_alm.info.comm : throw(DomainError(0, "Communicators must match"))
    p = plan_for(d_map.info.nside, d_alm.info.lmax, comm; thetatot = d_map.info.thetatot, phi0 = d_map.info.phi0)
    check_shards(p, d_alm, d_map)
    check(ccall((:hpsht_alm2map, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Cint),
                p.handle, d_alm.alm, d_map.pixels, size(d_alm.alm, 2), nthreads))
    d_map
end

function adjoint_alm2map_fused!(d_map, d_alm; nthreads::Integer = 0)
    (size(d_alm.alm, 2) == size(d_map.pixels, 2)) || throw(DomainError(0, "Number of components must match."))
    comm = (d_alm.info.comm == d_map.info.comm) ? d_alm.info.comm : throw(DomainError(0, "Communicators must match"))
    p = plan_for(d_map.info.nside, d_alm.info.lmax, comm; thetatot = d_map.info.thetatot, phi0 = d_map.info.phi0)
    check_shards(p, d_alm, d_map)
    check(ccall((:hpsht_adjoint_alm2map, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Cint),
                p.handle, d_map.pixels, d_alm.alm, size(d_alm.alm, 2), nthreads))
    d_alm
end

# Stream-ordered forms for device-resident shards (CUDA.jl CuArrays): `alm` / `pixels` are device pointers, the
# transform is ordered after the work queued on `stream` (a CUstream / cudaStream_t) and does not block the host.
function alm2map_async!(p::Plan, d_alm_ptr::Ptr{Float64}, d_pixels_ptr::Ptr{Float64}, ncomp::Integer, stream::Ptr{Cvoid})
    check(ccall((:hpsht_alm2map_async, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Cvoid}),
                p.handle, d_alm_ptr, d_pixels_ptr, ncomp, stream))
end
function adjoint_alm2map_async!(p::Plan, d_pixels_ptr::Ptr{Float64}, d_alm_ptr::Ptr{Float64}, ncomp::Integer, stream::Ptr{Cvoid})
    check(ccall((:hpsht_adjoint_alm2map_async, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Cvoid}),
                p.handle, d_pixels_ptr, d_alm_ptr, ncomp, stream))
end
synchronize(p::Plan) = check(ccall((:hpsht_plan_synchronize, lib), Cint, (Ptr{Cvoid},), p.handle))

# Shards on the device (SURVEY 8f N1-N3).  Methods to add next to src/alm.jl:621-660, 693-713 and
# src/cl.jl:33-38; `comp` arguments are 1-based columns of d_alm.alm as in the reference.
function dot_fused(a, b; comp1::Integer = 1, comp2::Integer = 1, nside::Integer)
    comm = (a.info.comm == b.info.comm) ? a.info.comm : throw(DomainError(0, "Communicators must match"))
    p = plan_for(nside, a.info.lmax, comm)
    res = Ref{Float64}(0.0)
    check(ccall((:hpsht_plan_dot, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ref{Float64}),
                p.handle, pointer(a.alm, (comp1 - 1) * size(a.alm, 1) + 1),
                pointer(b.alm, (comp2 - 1) * size(b.alm, 1) + 1), res))
    res[]
end

# almxfl!(alm, fl) (src/alm.jl:693-713): column c of `alm` is scaled by column c of `fl`, for the first
# min(size(alm.alm, 2), size(fl, 2)) columns only -- a vector `fl` touches column 1 alone.
function almxfl_fused!(alm, fl::AbstractVecOrMat{Float64}; nside::Integer)
    p = plan_for(nside, alm.info.lmax, alm.info.comm)
    ncol = min(size(alm.alm, 2), size(fl, 2))
    for c in 1:ncol
        flc = Vector{Float64}(fl[:, c])
        check(ccall((:hpsht_plan_almxfl, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Int64, Cint),
                    p.handle, pointer(alm.alm, (c - 1) * size(alm.alm, 1) + 1), flc, length(flc), 1))
    end
    alm
end

function alm2cl_fused(a, b; comp1::Integer = 1, comp2::Integer = 1, nside::Integer)
    p = plan_for(nside, a.info.lmax, a.info.comm)
    cl = zeros(Float64, a.info.lmax + 1)
    check(ccall((:hpsht_plan_alm2cl, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}),
                p.handle, pointer(a.alm, (comp1 - 1) * size(a.alm, 1) + 1),
                pointer(b.alm, (comp2 - 1) * size(b.alm, 1) + 1), cl))
    cl
end

# Scatter! / Gather! data movement for an already described shard (d_alm.info filled as in src/alm.jl:159-176):
# global_alm is the root's Alm.alm (m-major), ignored elsewhere.
function scatter_alm_fused!(global_alm::Union{Nothing,Vector{ComplexF64}}, d_alm; root::Integer = 0, nside::Integer)
    p = plan_for(nside, d_alm.info.lmax, d_alm.info.comm)
    g = global_alm === nothing ? Ptr{ComplexF64}(C_NULL) : pointer(global_alm)
    check(ccall((:hpsht_plan_scatter_alm, lib), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cint),
                p.handle, g, d_alm.alm, 1, root))
    d_alm
end

end # module
