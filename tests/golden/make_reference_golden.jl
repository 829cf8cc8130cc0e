# make_reference_golden.jl — golden outputs of the REAL reference (HealpixMPI.jl on Ducc0) for this repo's seeded inputs.
#
# NOT RUN in the build image (no Julia / MPI / Ducc0 there).  On a machine with Julia + MPI:
#
#   python tests/golden/make_reference_inputs.py                      # writes tests/golden/ref_inputs/
#   mpiexec -n 2 julia --project tests/golden/make_reference_golden.jl tests/golden/ref_inputs n64_l191
#   mpiexec -n 1 julia --project tests/golden/make_reference_golden.jl tests/golden/ref_inputs n1024_l3071
#
# Each run writes, next to the inputs, <tag>_refmap.f64 (alm2map! of the input alm, gathered, RING order, one
# component after the other) and <tag>_refalm.c128 (adjoint_alm2map! of the input map, gathered, Alm order).
# tests/test_reference_golden.py consumes them (python tests/golden/pack_reference_golden.py packs them into the
# committed tests/golden/reference_<tag>.npz).  The calls are the reference's own public API, driven exactly as in
# README.md:32-100 and test/test_MPI_a2m.jl / test_MPI_adj_a2m_pol.jl.
using MPI
using Healpix
using HealpixMPI
using JSON

MPI.Init()
comm = MPI.COMM_WORLD
crank = MPI.Comm_rank(comm)
csize = MPI.Comm_size(comm)
root = 0

dir, casename = ARGS[1], ARGS[2]
manifest = JSON.parsefile(joinpath(dir, "manifest.json"))

for c in manifest
    c["name"] == casename || continue
    c["ranks"] == csize || error("case $(c["name"]) is meant for $(c["ranks"]) ranks, got $csize")
    tag, nside, lmax, ncomp = c["tag"], c["nside"], c["lmax"], c["ncomp"]
    nalm, npix = c["nalm"], c["npix"]
    # ---- inputs (root only) ----
    if crank == root
        alm_in = Array{ComplexF64}(undef, nalm, ncomp)
        read!(joinpath(dir, tag * "_alm.c128"), alm_in)
        map_in = Array{Float64}(undef, npix, ncomp)
        read!(joinpath(dir, tag * "_map.f64"), map_in)
        h_alms = [Alm(lmax, lmax, alm_in[:, i]) for i in 1:ncomp]
        h_maps = [HealpixMap{Float64,RingOrder}(map_in[:, i]) for i in 1:ncomp]
    else
        h_alms = nothing
        h_maps = nothing
    end
    d_alm = DAlm{RR}(comm)
    d_map = DMap{RR}(comm)
    # ---- alm2map! ----
    if ncomp == 1
        MPI.Scatter!(crank == root ? h_alms[1] : nothing, d_alm, comm)
        MPI.Scatter!(crank == root ? HealpixMap{Float64,RingOrder}(nside) : nothing, d_map, comm)
    else
        MPI.Scatter!(h_alms, d_alm, comm)                                    # [E, B]
        pm = crank == root ? PolarizedHealpixMap{Float64,RingOrder}(nside) : nothing
        MPI.Scatter!(pm, d_map, comm)                                        # Q, U of an empty map
    end
    MPI.Barrier(comm)
    alm2map!(d_alm, d_map; nthreads = 0)
    out_maps = crank == root ? [HealpixMap{Float64,RingOrder}(nside) for i in 1:ncomp] : [nothing for i in 1:ncomp]
    for i in 1:ncomp
        MPI.Gather!(d_map, out_maps[i], i)
    end
    # ---- adjoint_alm2map! on the input maps ----
    d_map2 = DMap{RR}(comm)
    d_alm2 = DAlm{RR}(comm)
    if ncomp == 1
        MPI.Scatter!(crank == root ? h_maps[1] : nothing, d_map2, comm)
        MPI.Scatter!(crank == root ? Alm(lmax, lmax) : nothing, d_alm2, comm)
    else
        pm = crank == root ? PolarizedHealpixMap{Float64,RingOrder}(zeros(npix), map_in[:, 1], map_in[:, 2]) : nothing
        MPI.Scatter!(pm, d_map2, comm)                                       # Q, U
        MPI.Scatter!(crank == root ? [Alm(lmax, lmax) for i in 1:2] : nothing, d_alm2, comm)
    end
    MPI.Barrier(comm)
    adjoint_alm2map!(d_map2, d_alm2; nthreads = 0)
    out_alms = crank == root ? [Alm(lmax, lmax) for i in 1:ncomp] : nothing
    if ncomp == 1
        MPI.Gather!(d_alm2, crank == root ? out_alms[1] : nothing, clear = true)
    else
        MPI.Gather!(d_alm2, out_alms, clear = true)
    end
    if crank == root
        write(joinpath(dir, tag * "_refmap.f64"), hcat([m.pixels for m in out_maps]...))
        write(joinpath(dir, tag * "_refalm.c128"), hcat([a.alm for a in out_alms]...))
        println("written $(tag)_refmap.f64 / $(tag)_refalm.c128")
    end
end
MPI.Finalize()
