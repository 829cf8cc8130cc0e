# reference_cpu.jl — the reference's own CPU number for BASELINE.json's configs: HealpixMPI.jl on Ducc0, driven as in
# README.md:32-100, with timers.  NOT RUN in the build image (no Julia / MPI / Ducc0 there); run it next to
# `python bench.py` on the same host to get the real reference arm:
#
#   mpiexec -n P julia --project bench/reference_cpu.jl [nside] [lmax] [reps]     (defaults 4096 8191 3)
#
# One step = spin-0 alm2map! + spin-0 adjoint_alm2map! + spin-2 alm2map! + spin-2 adjoint_alm2map! on DAlm{RR} /
# DMap{RR} (BASELINE config 4), nthreads = Sys.CPU_THREADS ÷ P per rank.  Prints one JSON line on rank 0.
using MPI
using Random
using Healpix
using HealpixMPI

MPI.Init()
comm = MPI.COMM_WORLD
crank = MPI.Comm_rank(comm)
csize = MPI.Comm_size(comm)
root = 0

NSIDE = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 4096
lmax = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 2 * NSIDE - 1
reps = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 3
nthreads = max(1, Sys.CPU_THREADS ÷ csize)

Random.seed!(4096)
if crank == root
    h_map = HealpixMap{Float64,RingOrder}(randn(Float64, nside2npix(NSIDE)))
    h_pmap = PolarizedHealpixMap{Float64,RingOrder}(randn(Float64, nside2npix(NSIDE)), randn(Float64, nside2npix(NSIDE)),
                                                    randn(Float64, nside2npix(NSIDE)))
    h_alm = Alm(lmax, lmax, randn(ComplexF64, numberOfAlms(lmax)))
    h_palm = [Alm(lmax, lmax, randn(ComplexF64, numberOfAlms(lmax))) for i in 1:2]
else
    h_map = nothing; h_pmap = nothing; h_alm = nothing; h_palm = nothing
end
d_map = DMap{RR}(comm); d_alm = DAlm{RR}(comm)
d_pmap = DMap{RR}(comm); d_palm = DAlm{RR}(comm)
MPI.Scatter!(h_map, d_map, comm)
MPI.Scatter!(h_alm, d_alm, comm)
MPI.Scatter!(h_pmap, d_pmap, comm)        # Q, U
MPI.Scatter!(h_palm, d_palm, comm)        # E, B

function step!()
    alm2map!(d_alm, d_map; nthreads = nthreads)
    adjoint_alm2map!(d_map, d_alm; nthreads = nthreads)
    alm2map!(d_palm, d_pmap; nthreads = nthreads)
    adjoint_alm2map!(d_pmap, d_palm; nthreads = nthreads)
end

step!()                                   # warm-up (compilation, Ducc0 plans)
times = Float64[]
for i in 1:reps
    MPI.Barrier(comm)
    t0 = time_ns()
    step!()
    MPI.Barrier(comm)
    push!(times, (time_ns() - t0) / 1e6)
end
tmax = MPI.Allreduce(minimum(times), MPI.MAX, comm)
if crank == root
    println("{\"impl\": \"HealpixMPI.jl + Ducc0\", \"metric\": \"ms per step (spin-0 + spin-2 alm2map!/adjoint_alm2map!)\", " *
            "\"nside\": $NSIDE, \"lmax\": $lmax, \"ranks\": $csize, \"nthreads_per_rank\": $nthreads, " *
            "\"cpu_threads\": $(Sys.CPU_THREADS), \"value\": $tmax, \"unit\": \"ms\", \"reps\": $reps}")
end
MPI.Finalize()
