/*
 * hpsht_b200.h — C ABI of the B200-native spherical-harmonic-transform library that
 * replaces the Ducc0.jl + MPI.Alltoallv! seam of HealpixMPI.jl's hot path
 * (alm2map! / adjoint_alm2map! on DAlm{RR} / DMap{RR}; reference src/sht.jl:75-94, 165-179).
 *
 * Conventions
 *   - Every function returns int: 0 = success, non-zero = error; hpsht_last_error() gives the
 *     thread-local message.  Nothing throws across the boundary, nothing calls exit().
 *   - All index arrays at this boundary are 0-BASED (the Julia shim subtracts 1 from
 *     info.mstart / info.rstart exactly like Ducc0.jl's glue does).
 *   - Arrays use the memory order Julia hands over (column-major, first index fastest):
 *       alm  (nalm_loc, ncomp)           ->  alm[comp*alm_cstride + i]        complex128 (re,im)
 *       map  (npix_loc, ncomp)           ->  map[comp*map_cstride + p]        float64
 *       leg  alm-side (nm_loc, nring, ncomp) -> leg[(comp*nring + ir)*nm_loc + mi]   complex128
 *       leg  map-side (nm_tot, nring_loc, ncomp) -> leg[(comp*nring_loc + ir)*nm_tot + m]
 *   - Every data pointer may be a HOST pointer or a DEVICE pointer (cudaMalloc'd on the
 *     plan's device); the library detects which (cudaPointerGetAttributes) and stages
 *     host buffers through pinned memory.  Small descriptor arrays (mval, mstart, theta,
 *     nphi, phi0, ringstart) are always HOST pointers.
 *   - ncomp == 1  <=> spin 0 ; ncomp == 2 <=> spin 2 with (E,B) <-> (Q,U)   (src/sht.jl:84).
 *   - nthreads is accepted for signature parity with the reference and ignored.
 *   - Calls are synchronous (device work is complete on return) unless stated otherwise.
 *   - The product path is CUDA only: without a usable sm_100 device the compute entry
 *     points fail with HPSHT_ERR_CUDA; there is no CPU fallback.
 */
#ifndef HPSHT_B200_H
#define HPSHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPSHT_VERSION_MAJOR 0
#define HPSHT_VERSION_MINOR 1

enum {
  HPSHT_OK = 0,
  HPSHT_ERR_ARG = 1,      /* invalid argument (DomainError on the Julia side)            */
  HPSHT_ERR_CUDA = 2,     /* CUDA runtime failure, or no usable device                   */
  HPSHT_ERR_NCCL = 3,     /* NCCL failure / libnccl not loadable                         */
  HPSHT_ERR_UNSUPPORTED = 4, /* e.g. ring length not a multiple of 4, lstride != 1       */
  HPSHT_ERR_EMPTY_RANK = 5,  /* a rank would own no m or no ring (src/alm.jl:168, src/map.jl:162) */
  HPSHT_ERR_INTERNAL = 6
};

/* ------------------------------------------------------------------------------------ */
/* 0. Library info / errors                                                             */
/* ------------------------------------------------------------------------------------ */
int         hpsht_version(void);              /* major*100 + minor                                 */
const char* hpsht_last_error(void);           /* thread-local, valid until the next call on thread */
int         hpsht_device_count(int* count);   /* number of CUDA devices visible (0 on a CPU box)   */

/* ------------------------------------------------------------------------------------ */
/* 1. Round-robin bookkeeping (integer-exact restatement; host only, no CUDA needed)    */
/*    replaces src/alm.jl:102-151 and src/map.jl:99-143.                                */
/* ------------------------------------------------------------------------------------ */
/* get_nm_RR        (src/alm.jl:102-105): number of m owned by task_rank                */
int hpsht_rr_nm(int64_t mmax, int task_rank, int c_size, int64_t* nm);
/* get_mval_RR      (src/alm.jl:112-119): mval[i] = task_rank + i*c_size, i<nm          */
int hpsht_rr_mval(int64_t mmax, int task_rank, int c_size, int64_t* mval /*[nm]*/);
/* get_m_tasks_RR   (src/alm.jl:127-133): task[m] = m % c_size, m=0..mmax               */
int hpsht_rr_m_tasks(int64_t mmax, int c_size, int64_t* task /*[mmax+1]*/);
/* make_mstart_complex (src/alm.jl:141-151) but 0-BASED: alm[mstart[mi] + l] = a_{l,mval[mi]}.
   (Julia's 1-based value is this + 1.)                                                */
int hpsht_rr_mstart(int64_t lmax, int64_t stride, int64_t nm, const int64_t* mval,
                    int64_t* mstart /*[nm]*/);
/* get_nrings_RR    (src/map.jl:99-102): eq_idx = 2*nside                               */
int hpsht_rr_nrings(int64_t eq_idx, int task_rank, int c_size, int64_t* nrings);
/* get_rindexes_RR  (src/map.jl:113-121): 1-BASED global ring ids, equator outwards N,S */
int hpsht_rr_rindexes(int64_t eq_idx, int task_rank, int c_size, int64_t* rings /*[nrings]*/);
/* get_rindexes_tot_RR (src/map.jl:134-143): concatenation over ranks, length 2*eq_idx-1 */
int hpsht_rr_rindexes_tot(int64_t eq_idx, int c_size, int64_t* rings /*[2*eq_idx-1]*/);
/* all-to-all counts of communicate_alm2map! (src/sht.jl:20-27), in complex elements per
   component: send[t] = nm(rank)*nrings(t), recv[t] = nrings(rank)*nm(t)                 */
int hpsht_rr_a2a_counts(int64_t mmax, int64_t eq_idx, int task_rank, int c_size,
                        int64_t* send /*[c_size]*/, int64_t* recv /*[c_size]*/);

/* HEALPix RING-scheme geometry of one ring (1-based ring id 1..4*nside-1); what
   Healpix.getringinfo!/pix2ang provide to ScatterMap! (src/map.jl:177-183).
   first_pix is the 0-based RING index of the ring's first pixel.                       */
int hpsht_healpix_ring(int64_t nside, int64_t ring, int64_t* nphi, double* theta,
                       double* phi0, int64_t* first_pix);

/* ------------------------------------------------------------------------------------ */
/* 2. Stage-level entry points — 1:1 with the four Ducc0.Sht calls of src/sht.jl        */
/* ------------------------------------------------------------------------------------ */
/* Ducc0.Sht.alm2leg!(alm, leg, spin, lmax, mval, mstart, lstride, theta, nthreads)
   (src/sht.jl:85).  leg[(comp*nring+ir)*nm + mi] = sum_l alm-term * lambda(theta[ir]).  */
int hpsht_alm2leg(const double* alm, int64_t alm_cstride, double* leg, int spin,
                  int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                  int64_t lstride, int64_t nring, const double* theta, int nthreads);
/* Ducc0.Sht.leg2alm!(leg, alm, spin, lmax, mval, mstart, lstride, theta, nthreads)
   (src/sht.jl:178): adjoint of the above.  Entries of alm with l < spin are set to 0.   */
int hpsht_leg2alm(const double* leg, double* alm, int64_t alm_cstride, int spin,
                  int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                  int64_t lstride, int64_t nring, const double* theta, int nthreads);
/* Ducc0.Sht.leg2map!(leg, map, nphi, phi0, ringstart, pixstride, nthreads) (src/sht.jl:93)
   leg is map-side: (nm, nring, ncomp) with m = 0..nm-1 contiguous.                      */
int hpsht_leg2map(const double* leg, double* map, int64_t map_cstride, int ncomp,
                  int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                  const int64_t* ringstart, int64_t pixstride, int nthreads);
/* Ducc0.Sht.map2leg!(map, leg, nphi, phi0, ringstart, pixstride, nthreads) (src/sht.jl:169) */
int hpsht_map2leg(const double* map, int64_t map_cstride, double* leg, int ncomp,
                  int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                  const int64_t* ringstart, int64_t pixstride, int nthreads);

/* ------------------------------------------------------------------------------------ */
/* 3. Fused plan — owns leg buffers, tables, streams and the NCCL communicator.          */
/*    alm2map!(d_alm, d_map) / adjoint_alm2map!(d_map, d_alm)  (src/sht.jl:75-100,165-185)*/
/* ------------------------------------------------------------------------------------ */
typedef struct hpsht_plan hpsht_plan;

#define HPSHT_NCCL_ID_BYTES 128
/* rank 0 calls this and ships the 128 bytes to the other ranks (MPI.bcast in Julia,
   torch.distributed / file store in the Python harness).                                */
int hpsht_nccl_unique_id(unsigned char id[HPSHT_NCCL_ID_BYTES]);

/* Collective over the n_ranks processes.  The RR geometry (which m, which rings, all
   descriptor arrays of AlmInfoMPI / GeomInfoMPI) is derived inside from (nside,lmax,rank,
   n_ranks) and must equal what Scatter! builds (src/alm.jl:159-176, src/map.jl:151-197).
   device < 0: use the current CUDA device.  nccl_id may be NULL iff n_ranks == 1.
   max_ncomp (1 or 2) sizes the internal leg buffers.                                    */
int hpsht_plan_create(hpsht_plan** plan, int64_t nside, int64_t lmax, int rank, int n_ranks,
                      const unsigned char* nccl_id, int device, int max_ncomp);
/* Optional: override the internally derived ring geometry with the caller's own doubles
   (d_map.info.thetatot [nring_tot] and d_map.info.phi0 [nring_loc], src/map.jl:189-196), so the
   plan sees bit-identical inputs to what the reference hands Ducc0.  Either may be NULL.
   Must be called before the first transform.                                            */
int hpsht_plan_set_geometry(hpsht_plan* plan, const double* thetatot, const double* phi0_loc);
int hpsht_plan_destroy(hpsht_plan* plan);

/* Sizes and descriptor arrays of this rank's shard (so callers can cross-check them
   against DAlm.info / DMap.info).  Any output pointer may be NULL.                      */
int hpsht_plan_sizes(const hpsht_plan* plan, int64_t* nm_loc, int64_t* nalm_loc,
                     int64_t* nring_loc, int64_t* npix_loc, int64_t* nring_tot);
int hpsht_plan_alm_info(const hpsht_plan* plan, int64_t* mval, int64_t* mstart /*0-based*/);
int hpsht_plan_map_info(const hpsht_plan* plan, int64_t* rings /*1-based ids*/,
                        int64_t* rstart /*0-based*/, int64_t* nphi, double* theta,
                        double* phi0, double* thetatot /*[nring_tot]*/);

/* alm (nalm_loc, ncomp) complex128 -> map (npix_loc, ncomp) float64.  Host or device
   pointers (see top).  Collective: every rank of the plan calls it, and every rank passes the SAME KIND of
   pointer for `map` (all host or all device) -- the pipeline a call runs (how many NCCL exchanges it issues)
   depends on it, a mix would dead-lock.  With device pointers: the library waits for the device on entry
   (the caller's producer stream is unknown to it; its own streams are non-blocking) and the call is
   complete on return.                                                                  */
int hpsht_alm2map(hpsht_plan* plan, const double* alm, double* map, int ncomp, int nthreads);
/* map -> alm, exact transpose Y^T (no quadrature weights).  Collective.                */
int hpsht_adjoint_alm2map(hpsht_plan* plan, const double* map, double* alm, int ncomp,
                          int nthreads);

/* Stream-ordered forms for solver loops that keep DAlm / DMap resident on the device (device pointers
   only): the transform is ordered after everything queued on `cuda_stream` (a cudaStream_t; NULL = the
   legacy default stream) at the time of the call, `cuda_stream` continues after it, and the host is not
   blocked.  Collective like the blocking forms.  hpsht_plan_synchronize waits for the plan's streams;
   hpsht_plan_timings implies it.                                                         */
int hpsht_alm2map_async(hpsht_plan* plan, const double* d_alm, double* d_map, int ncomp, void* cuda_stream);
int hpsht_adjoint_alm2map_async(hpsht_plan* plan, const double* d_map, double* d_alm, int ncomp,
                                void* cuda_stream);
int hpsht_plan_synchronize(hpsht_plan* plan);
/* P > 1: run device-pointer calls ring block by ring block, the leg exchange / ring FFT of a block behind the
   Legendre kernels of the next one (default on; HPSHT_DEV_BLOCKS=0 in the environment turns it off at plan
   creation).  Off = the stages run back to back, which is what the per-stage timers want.  Must be set
   identically on every rank of the plan.                                                 */
int hpsht_plan_set_dev_blocks(hpsht_plan* plan, int on);
/* The ring blocks a plan of (nside, lmax, n_ranks) cuts (host arithmetic only, no device needed): nblocks and,
   for block b and rank t, the range [lo[b*n_ranks + t], hi[b*n_ranks + t]) of rank t's local rings.  It takes
   no rank argument: every rank of a collective call derives the same cut (and the same send / receive counts
   per block) from these numbers.  lo / hi may be NULL; capacity = entries available in each.            */
int hpsht_ring_block_ranges(int64_t nside, int64_t lmax, int n_ranks, int* nblocks, int64_t* lo, int64_t* hi,
                            int capacity);

/* ---- shards on the device: scatter / gather / algebra -------------------
   All collective over the plan's ranks; host or device pointers.  `ncomp` components are moved
   together ([ncomp][...] as in hpsht_alm2map).
   Global alm: Healpix.jl Alm order, 0-based index(l, m) = m (2 lmax + 1 - m) / 2 + l, [ncomp][nalm_tot]
   complex128, nalm_tot = (lmax+1)(lmax+2)/2.  Global map: RING order, [ncomp][12 nside^2] float64.

   scatter: replaces MPI.Scatter!(alm, d_alm) / ScatterAlm! (src/alm.jl:159-186,219-307) and
   MPI.Scatter!(map, d_map) / ScatterMap! (src/map.jl:151-212,245-330); the global array is read on
   `root` only (may be NULL elsewhere); every rank receives exactly its shard.
   gather: replaces MPI.Gather! / MPI.Allgather! (src/alm.jl:313-595, src/map.jl:337-559); root >= 0:
   the global array is written on `root` only; root = -1: on every rank (Allgather!).              */
int hpsht_plan_scatter_alm(hpsht_plan* plan, const double* global_alm, double* alm, int ncomp, int root);
int hpsht_plan_gather_alm(hpsht_plan* plan, const double* alm, double* global_alm, int ncomp, int root);
int hpsht_plan_scatter_map(hpsht_plan* plan, const double* global_map, double* map, int ncomp, int root);
int hpsht_plan_gather_map(hpsht_plan* plan, const double* map, double* global_map, int ncomp, int root);
/* dot(a, b) of one component each (nalm_loc complex128): weight 1 for m = 0 (real parts only), 2 for
   m > 0, summed over all ranks (localdot + Allreduce, src/alm.jl:621-660).                         */
int hpsht_plan_dot(hpsht_plan* plan, const double* alm1, const double* alm2, double* result);
/* almxfl! kernel: alm[l, m] *= fl[l] on `ncomp` consecutive components; fl has nfl entries and is zero-extended
   to lmax + 1.  The reference (src/alm.jl:693-713) scales column c of the DAlm by column c of `fl` for the first
   min(ncomp, ncol(fl)) columns only -- the bindings loop over those columns with ncomp = 1 and the column's pointer
   (healpixmpi.jl_b200/sht.py Plan.almxfl, julia/HealpixMPIB200.jl almxfl_fused!).                    */
int hpsht_plan_almxfl(hpsht_plan* plan, double* alm, const double* fl, int64_t nfl, int ncomp);
/* alm2cl of one component each: cl[l] = sum_m w_m Re(a conj b) / (2l + 1), w = 1 (m = 0) or 2, summed
   over all ranks; cl has lmax + 1 entries (src/cl.jl:5-38).                                         */
int hpsht_plan_alm2cl(hpsht_plan* plan, const double* alm1, const double* alm2, double* cl);

/* Per-stage device times (ms, CUDA events) of the most recent fused call on this plan:
   [0] total, [1] the Legendre recurrence kernel alone, [2] whole Legendre stage (with the O(nalm)
   prep / reduce / post kernels), [3] all-to-all (+pack/unpack), [4] ring FFT, [5] H2D, [6] D2H.
   Host-pointer calls run ring block by ring block with the copies on their own stream, so there the
   stages overlap: [0] stays the whole call, [1]/[2] span first to last Legendre block (the ring FFTs of
   the blocks in between included), [4] is the last block's FFT, [5]/[6] only the copies that were not
   hidden.  The per-stage figures `bench.py` reports come from device-pointer calls.       */
#define HPSHT_NTIMES 8
int hpsht_plan_timings(const hpsht_plan* plan, double ms[HPSHT_NTIMES]);
/* Number of kernel launches issued by the most recent fused call.                       */
int hpsht_plan_launches(const hpsht_plan* plan, int64_t* n);
/* Algorithmic Legendre work of this rank for a given spin (0 or 2): number of
   (l, m, ring-pair) recurrence steps after the m <= mlim(theta) cut.         */
int hpsht_plan_legendre_steps(const hpsht_plan* plan, int spin, double* steps_culled,
                              double* steps_dense);

/* Device-memory helpers so a host language without a CUDA binding can keep shards
   resident.                                                                 */
int hpsht_device_malloc(void** dptr, size_t bytes);
int hpsht_device_free(void* dptr);
/* page-locked host memory (fast H2D/D2H for host-pointer calls) */
int hpsht_host_malloc(void** hptr, size_t bytes);
int hpsht_host_free(void* hptr);
int hpsht_memcpy_h2d(void* dptr, const void* hptr, size_t bytes);
int hpsht_memcpy_d2h(void* hptr, const void* dptr, size_t bytes);
int hpsht_device_synchronize(void);

/* FP64 FMA micro-benchmark (no FP64 entry in MEASURED_PEAKS.json): returns sustained
   TFLOP/s of dependent-chain-free DFMA on the current device.                           */
int hpsht_measure_fp64_peak(double* tflops, double* seconds_run);

#ifdef __cplusplus
}
#endif
#endif /* HPSHT_B200_H */
