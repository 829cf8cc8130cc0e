// api.cu — the C ABI (include/hpsht_b200.h): round-robin bookkeeping, the four stage-level entry
// points that mirror the Ducc0.Sht calls of reference src/sht.jl:85,93,169,178, and the fused plan
// that owns the leg buffers, the RR geometry and the NCCL exchange replacing MPI.Alltoallv!
// (src/sht.jl:32,133).
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.hpp"
#include "legendre.cuh"
#include "nccl_dyn.hpp"
#include "ringfft.cuh"
#include "rr.hpp"

namespace hpsht {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

namespace {

void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    throw Error(HPSHT_ERR_CUDA,
                "no CUDA device available: libhpsht_b200 has no CPU fallback (the product path is "
                "sm_100a kernels only)");
  }
}

// ---- exchange pack / unpack (map-side leg <-> per-peer blocks) ----
// legF[(comp*nr_loc + j)*nm_tot + m]  <->  block of source s = m % P, laid out like the reference's receive
// buffer (src/sht.jl:35-42): boff[s] + (comp*nr_loc + j)*nm_s + i,  i = m / P (local m index on s).
struct ExchGeom {
  int64_t nm_tot, nr_loc, mmax;
  int P;
};
__device__ __forceinline__ int64_t mapside_off(const ExchGeom& g, const int64_t* boff, int64_t m, int64_t j, int comp) {
  const int s = (int)(m % g.P);
  const int64_t i = m / g.P;
  const int64_t nm_s = (g.mmax + g.P - s) / g.P;
  return boff[s] + ((int64_t)comp * g.nr_loc + j) * nm_s + i;
}
// rows j0 .. j0+nj-1 of every component (the whole array: j0 = 0, nj = nr_loc)
__global__ void k_unpack(const double2* __restrict__ mapside, const int64_t* __restrict__ boff,
                         double2* __restrict__ legF, ExchGeom g, int64_t j0, int64_t nj) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.nm_tot) return;
  const int comp = (int)(blockIdx.y / nj);
  const int64_t j = j0 + blockIdx.y % nj;
  legF[(comp * g.nr_loc + j) * g.nm_tot + m] = mapside[mapside_off(g, boff, m, j, comp)];
}
__global__ void k_pack(const double2* __restrict__ legF, const int64_t* __restrict__ boff,
                       double2* __restrict__ mapside, ExchGeom g, int64_t j0, int64_t nj) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= g.nm_tot) return;
  const int comp = (int)(blockIdx.y / nj);
  const int64_t j = j0 + blockIdx.y % nj;
  mapside[mapside_off(g, boff, m, j, comp)] = legF[(comp * g.nr_loc + j) * g.nm_tot + m];
}

// ---- FP64 FMA peak probe: 8 independent chains per thread, no memory traffic ----
__global__ void k_dfma_peak(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
         a6 = seed + 6, a7 = seed + 7;
  const double x = 0.999999, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, x, y);
      a1 = fma(a1, x, y);
      a2 = fma(a2, x, y);
      a3 = fma(a3, x, y);
      a4 = fma(a4, x, y);
      a5 = fma(a5, x, y);
      a6 = fma(a6, x, y);
      a7 = fma(a7, x, y);
    }
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

struct Timer {
  cudaEvent_t ev[16];
  Timer() {
    for (auto& e : ev) cudaEventCreate(&e);
  }
  ~Timer() {
    for (auto& e : ev) cudaEventDestroy(e);
  }
};

}  // namespace

// =================================================================================================
// plan
// =================================================================================================
}  // namespace hpsht

using namespace hpsht;

// shard descriptors of one rank on the device (shards.cuh)
struct RankAlmDesc {
  std::vector<int64_t> mval, mstart0;
  int64_t nalm = 0;
  DevBuf<int64_t> d_mval, d_mstart0;
};
struct RankMapDesc {
  std::vector<int64_t> rstart0, first_pix, nphi;
  int64_t npix = 0;
  DevBuf<int64_t> d_rstart0, d_first_pix, d_nphi;
};

struct hpsht_plan {
  int64_t nside = 0, lmax = 0;
  int rank = 0, P = 1, device = 0, max_ncomp = 1;
  // RR shard description (what Scatter! builds: src/alm.jl:159-176, src/map.jl:151-197)
  std::vector<int64_t> mval, mstart, rings, rstart, nphi;
  std::vector<double> theta, phi0, thetatot;
  std::vector<int64_t> nr_of, nm_of;  // per rank
  int64_t nm_loc = 0, nalm_loc = 0, nr_loc = 0, npix_loc = 0, nr_tot = 0, nm_tot = 0;
  bool stages_ready = false;

  cudaStream_t stream = nullptr;   // Legendre stage
  cudaStream_t cstream = nullptr;  // NCCL exchange of a ring block (overlaps the Legendre work of the next block)
  cudaStream_t fstream = nullptr;  // P>1: ring FFT + pack/unpack of a block
  cudaStream_t hstream = nullptr;  // host -> device copies (and all copies of single-direction calls)
  cudaStream_t ostream = nullptr;  // device -> host copies
  LegendreStage leg;
  RingFFT fft;
  DevBuf<double> legF;      // map-side leg [ncomp][nr_loc][nm_tot] (complex)
  DevBuf<double> almside;   // P>1: per-peer blocks [comp][nr_t][nm_loc]
  DevBuf<double> mapside;   // P>1: per-peer blocks [comp][nr_loc][nm_s]
  std::vector<int64_t> aoff, boff;  // block offsets in complex elements (max_ncomp sized)
  DevBuf<int64_t> d_boff;
  DevBuf<double> stage_alm, stage_map;  // device staging for host-pointer calls
  RankAlmDesc own_alm;                  // shard ops (shards.cuh): this rank's descriptors, built lazily
  RankMapDesc own_map;
  DevBuf<double> shard_scratch;
  // Ring blocks (block set 1 of the Legendre stage): HB ranges of the global pole -> equator pair list.  Used
  // to pipeline Legendre(b) / exchange(b) / ring FFT(b) / PCIe copy(b); HB is derived from quantities that
  // are identical on every rank.
  int HB = 1;
  bool dev_blocks = true;               // P>1: device-pointer calls run block by block too (exchange overlap)
  std::vector<std::unique_ptr<RingFFT>> fftb;
  std::vector<int64_t> bj_lo, bj_hi;    // [block][rank]: local ring range of the block on that rank
  std::vector<cudaEvent_t> ev_blk, ev_x, ev_f, ev_alm;
  cudaEvent_t ev_hdone = nullptr, ev_a0 = nullptr, ev_a1 = nullptr, ev_entry = nullptr, ev_user = nullptr,
              ev_done = nullptr;
  int64_t blo(int set, int b, int t) const { return set ? bj_lo[(size_t)b * P + t] : 0; }
  int64_t bhi(int set, int b, int t) const { return set ? bj_hi[(size_t)b * P + t] : nr_of[(size_t)t]; }
  int64_t pix_at(int64_t j) const { return j < nr_loc ? rstart[(size_t)j] : npix_loc; }
  void setup_blocks();
  void exchange_block(bool synthesis, int ncomp, int set, int b, cudaStream_t st);
  ncclComm_t comm = nullptr;
  std::unique_ptr<Timer> tm;
  double ms[HPSHT_NTIMES] = {0};
  bool times_pending = false, last_synthesis = true;
  int64_t n_launch = 0;

  void build_shard();
  void ensure_stages();
  void run(bool synthesis, const double* in, double* out, int ncomp, bool async, cudaStream_t user);
  void synchronize();
  void finish_times();
};

void hpsht_plan::build_shard() {
  const int64_t eq = 2 * nside;
  HP_REQUIRE(nside >= 1 && lmax >= 0, HPSHT_ERR_ARG, "nside >= 1 and lmax >= 0 required");
  HP_REQUIRE(rank >= 0 && rank < P, HPSHT_ERR_ARG, "rank can not exceed communicator size");
  mval = rr_mval(lmax, rank, P);
  HP_REQUIRE(!mval.empty(), HPSHT_ERR_EMPTY_RANK, std::to_string(rank) + "-th task is empty.");
  mstart = rr_mstart0(lmax, 1, mval);
  rings = rr_rindexes(eq, rank, P);
  HP_REQUIRE(!rings.empty(), HPSHT_ERR_EMPTY_RANK, std::to_string(rank) + "-th MPI task has no rings.");
  nm_loc = (int64_t)mval.size();
  nalm_loc = 0;
  for (int64_t m : mval) nalm_loc += lmax + 1 - m;
  nr_loc = (int64_t)rings.size();
  nm_tot = lmax + 1;
  nr_tot = 2 * eq - 1;
  rstart.resize((size_t)nr_loc);
  nphi.resize((size_t)nr_loc);
  theta.resize((size_t)nr_loc);
  phi0.resize((size_t)nr_loc);
  npix_loc = 0;
  for (int64_t i = 0; i < nr_loc; ++i) {
    RingGeom g = healpix_ring(nside, rings[(size_t)i]);
    rstart[(size_t)i] = npix_loc;
    nphi[(size_t)i] = g.nphi;
    theta[(size_t)i] = g.theta;
    phi0[(size_t)i] = g.phi0;
    npix_loc += g.nphi;
  }
  std::vector<int64_t> tot = rr_rindexes_tot(eq, P);
  thetatot.resize(tot.size());
  for (size_t i = 0; i < tot.size(); ++i) thetatot[i] = healpix_ring(nside, tot[i]).theta;
  nr_of.resize((size_t)P);
  nm_of.resize((size_t)P);
  for (int t = 0; t < P; ++t) {
    nr_of[(size_t)t] = rr_nrings(eq, t, P);
    nm_of[(size_t)t] = rr_nm(lmax, t, P);
    HP_REQUIRE(nr_of[(size_t)t] > 0 && nm_of[(size_t)t] > 0, HPSHT_ERR_EMPTY_RANK,
               std::to_string(t) + "-th task is empty.");
  }
}

void hpsht_plan::ensure_stages() {
  if (stages_ready) return;
  // ---- leg layout seen by the Legendre kernels: ring slots in thetatot order ----
  LegLayout lay;
  lay.a0.resize((size_t)nr_tot);
  lay.nr.resize((size_t)nr_tot);
  lay.j.resize((size_t)nr_tot);
  int prio_lo = 0, prio_hi = 0;
  HP_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  if (P == 1) {
    // single rank: the Legendre stage reads/writes the map-side array directly (src/sht.jl:87-88)
    for (int64_t s = 0; s < nr_tot; ++s) {
      lay.a0[(size_t)s] = 0;
      lay.nr[(size_t)s] = nr_loc;
      lay.j[(size_t)s] = s;
    }
    legF.alloc((size_t)max_ncomp * nr_loc * nm_tot * 2);
  } else {
    // the per-destination blocks of the reference's all-to-all (src/sht.jl:20-32), all components in one
    aoff.assign((size_t)P + 1, 0);
    boff.assign((size_t)P + 1, 0);
    for (int t = 0; t < P; ++t) {
      aoff[(size_t)t + 1] = aoff[(size_t)t] + (int64_t)max_ncomp * nr_of[(size_t)t] * nm_loc;
      boff[(size_t)t + 1] = boff[(size_t)t] + (int64_t)max_ncomp * nr_loc * nm_of[(size_t)t];
    }
    int64_t s = 0;
    for (int t = 0; t < P; ++t)
      for (int64_t j = 0; j < nr_of[(size_t)t]; ++j, ++s) {
        lay.a0[(size_t)s] = aoff[(size_t)t];
        lay.nr[(size_t)s] = nr_of[(size_t)t];
        lay.j[(size_t)s] = j;
      }
    almside.alloc((size_t)aoff[(size_t)P] * 2);
    mapside.alloc((size_t)boff[(size_t)P] * 2);
    legF.alloc((size_t)max_ncomp * nr_loc * nm_tot * 2);
    d_boff.upload(boff.data(), boff.size(), stream);
    // exchange and the per-block FFT stream outrank the Legendre stream: their few CTAs must not queue
    // behind the thousands of pending CTAs of the next Legendre launch
    HP_CUDA(cudaStreamCreateWithPriority(&cstream, cudaStreamNonBlocking, prio_hi));
  }
  HP_CUDA(cudaStreamCreateWithPriority(&fstream, cudaStreamNonBlocking, prio_hi));
  HP_CUDA(cudaStreamCreateWithFlags(&hstream, cudaStreamNonBlocking));
  // One copy stream per direction.  (Several ranks of one host share its memory system -- tools/h2d_bench.py on the
  // 8 x B200 box, GB/s per rank, H2D alone vs H2D and D2H at once per direction: 2 ranks 51.7 vs 31.1, 4 ranks 35.5 vs
  // 13.8, 8 ranks 21.5 vs 7.4 -- but inside a transform the two directions rarely meet, and one shared stream made
  // the 8-rank host-pointer step slower, 161.6 -> 178.7 ms: profiles/r02_multi_gpu.md.)
  HP_CUDA(cudaStreamCreateWithFlags(&ostream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&ev_hdone, &ev_a0, &ev_a1, &ev_entry, &ev_user, &ev_done})
    HP_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  leg.setup(lmax, mval.data(), mstart.data(), nm_loc, nr_tot, thetatot.data(), lay, max_ncomp, stream);
  fft.setup(nr_loc, nphi.data(), phi0.data(), rstart.data(), nm_tot, stream);
  setup_blocks();
  for (auto* v : {&ev_blk, &ev_x, &ev_f}) {
    v->resize((size_t)std::max(HB, 1));
    for (auto& e : *v) HP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  HP_CUDA(cudaStreamSynchronize(stream));
  stages_ready = true;
}

// Host-pointer calls move 8 B per pixel and component over PCIe, ~0.4x the time the transform itself
// needs at nside 4096, and on P > 1 ranks every transform has the leg exchange in the middle.  With the
// ring pairs cut into HB tile-aligned blocks the copy / exchange / FFT of one block overlaps the Legendre
// work of the others (HPSHT_HOST_BLOCKS, 1 = off).  The blocks are ranges of the global pole -> equator
// pair list; with round-robin ring ownership every rank's share of a block is a contiguous range
// [blo, bhi) of its local rings.  Everything that decides the cut is the same on every rank (the ranks of a
// collective call must cut the same blocks): nside, lmax, P and the environment -- no rank-local quantity.
namespace {
// up to 2 ranks: blocks that grow from both ends (the copies are 0.3 - 0.6 of the compute, only the end blocks show);
// from 4 ranks on the ranks share the host's memory system, a block's copy takes about as long as its compute, and
// equal blocks pipeline best (profiles/r02_multi_gpu.md)
bool equal_ring_blocks(int P) { return P >= 4; }
int wanted_ring_blocks(int64_t nside, int P) {
  const char* eb = getenv("HPSHT_HOST_BLOCKS");
  int want = eb ? std::max(1, atoi(eb)) : (P <= 2 ? 9 : 8);
  if (!eb && 12 * nside * nside / P < (1 << 22)) want = 1;  // small maps: the copies are microseconds
  return want;
}
// [block][rank] local ring ranges of the cut `bp` (first pair of each block + end); false: exotic ring order
bool ring_block_ranges(const std::vector<PairHost>& pairs, const std::vector<int>& bp, const std::vector<int64_t>& nr_of,
                       std::vector<int64_t>& bj_lo, std::vector<int64_t>& bj_hi) {
  const int P = (int)nr_of.size(), HB = (int)bp.size() - 1;
  std::vector<int64_t> roff((size_t)P + 1, 0);  // first slot of every rank in thetatot order
  for (int t = 0; t < P; ++t) roff[(size_t)t + 1] = roff[(size_t)t] + nr_of[(size_t)t];
  bj_lo.assign((size_t)HB * P, 0);
  bj_hi.assign((size_t)HB * P, 0);
  for (int b = 0; b < HB; ++b) {
    std::vector<int64_t> lo((size_t)P, INT64_MAX), hi((size_t)P, -1), cnt((size_t)P, 0);
    for (int p = bp[(size_t)b]; p < bp[(size_t)b + 1]; ++p)
      for (int64_t slot : {pairs[(size_t)p].slotN, pairs[(size_t)p].slotS}) {
        if (slot < 0) continue;
        const int t = (int)(std::upper_bound(roff.begin(), roff.end(), slot) - roff.begin()) - 1;
        const int64_t j = slot - roff[(size_t)t];
        lo[(size_t)t] = std::min(lo[(size_t)t], j);
        hi[(size_t)t] = std::max(hi[(size_t)t], j + 1);
        ++cnt[(size_t)t];
      }
    for (int t = 0; t < P; ++t) {
      const int64_t prev = b == 0 ? nr_of[(size_t)t] : bj_lo[(size_t)(b - 1) * P + t];
      if (cnt[(size_t)t] == 0) lo[(size_t)t] = hi[(size_t)t] = prev;  // no ring of this rank in the block
      // blocks run pole -> equator = towards lower local ring indices, without gaps
      if (cnt[(size_t)t] != hi[(size_t)t] - lo[(size_t)t] || hi[(size_t)t] != prev) return false;
      bj_lo[(size_t)b * P + t] = lo[(size_t)t];
      bj_hi[(size_t)b * P + t] = hi[(size_t)t];
    }
  }
  for (int t = 0; t < P; ++t)
    if (bj_lo[(size_t)(HB - 1) * P + t] != 0) return false;
  return true;
}
}  // namespace

void hpsht_plan::setup_blocks() {
  HB = 1;
  const char* ed = getenv("HPSHT_DEV_BLOCKS");
  dev_blocks = ed ? atoi(ed) != 0 : true;
  std::vector<double> w((size_t)nr_tot);
  {
    std::vector<int64_t> tot = rr_rindexes_tot(2 * nside, P);
    for (size_t i = 0; i < tot.size(); ++i) w[i] = (double)healpix_ring(nside, tot[i]).nphi;
  }
  HB = leg.setup_blocks(wanted_ring_blocks(nside, P), w.data(), equal_ring_blocks(P));
  if (HB <= 1) return;
  std::vector<int> bp;
  // the pair list the Legendre stage cut is the one build_pairs gives for thetatot
  const std::vector<PairHost> pairs = build_pairs(lmax, nr_tot, thetatot.data());
  bp = cut_ring_blocks(pairs, wanted_ring_blocks(nside, P), w.data(), equal_ring_blocks(P));
  if (!ring_block_ranges(pairs, bp, nr_of, bj_lo, bj_hi)) {  // exotic ring order: no blocking
    HB = leg.setup_blocks(1, w.data());
    return;
  }
  fftb.clear();
  for (int b = 0; b < HB; ++b) {
    const int64_t lo = blo(1, b, rank), hi = bhi(1, b, rank);
    fftb.emplace_back(new RingFFT());
    if (hi > lo)
      fftb.back()->setup(hi - lo, nphi.data() + lo, phi0.data() + lo, rstart.data() + lo, nm_tot, stream, lo, nr_loc);
  }
}

// m-sharded <-> ring-sharded transpose of the leg array (communicate_alm2map! / communicate_map2alm!,
// src/sht.jl:5-44,106-135) for the rings of one block: per peer and component one contiguous run of rows
// on both sides, all of them in one grouped NCCL send/recv exchange.
void hpsht_plan::exchange_block(bool synthesis, int ncomp, int set, int b, cudaStream_t st) {
  NcclApi& N = NcclApi::get();
  double* A = almside.p;
  double* B = mapside.p;
  const int64_t mlo = blo(set, b, rank), mhi = bhi(set, b, rank);
  HP_NCCL(N.GroupStart());
  ncclResult_t err = ncclSuccess;
  for (int t = 0; t < P && err == ncclSuccess; ++t) {
    if (t == rank) continue;
    for (int c = 0; c < ncomp && err == ncclSuccess; ++c) {
      // alm-side: my m's x t's rings of the block ; map-side: t's m's x my rings of the block
      const size_t na = (size_t)(bhi(set, b, t) - blo(set, b, t)) * nm_loc * 2;
      const size_t nb = (size_t)(mhi - mlo) * nm_of[(size_t)t] * 2;
      double* pa = A + ((size_t)aoff[(size_t)t] + ((size_t)c * nr_of[(size_t)t] + blo(set, b, t)) * nm_loc) * 2;
      double* pb = B + ((size_t)boff[(size_t)t] + ((size_t)c * nr_loc + mlo) * nm_of[(size_t)t]) * 2;
      if (synthesis) {
        if (na) err = N.Send(pa, na, ncclDouble, t, comm, st);
        if (nb && err == ncclSuccess) err = N.Recv(pb, nb, ncclDouble, t, comm, st);
      } else {
        if (nb) err = N.Send(pb, nb, ncclDouble, t, comm, st);
        if (na && err == ncclSuccess) err = N.Recv(pa, na, ncclDouble, t, comm, st);
      }
    }
  }
  const ncclResult_t end = N.GroupEnd();  // always close the group, also on the error path
  HP_NCCL(err);
  HP_NCCL(end);
  if (mhi > mlo)
    for (int c = 0; c < ncomp; ++c) {  // own block never leaves the device
      const size_t n = (size_t)(mhi - mlo) * nm_loc * 2;
      double* pa = A + ((size_t)aoff[(size_t)rank] + ((size_t)c * nr_loc + mlo) * nm_loc) * 2;
      double* pb = B + ((size_t)boff[(size_t)rank] + ((size_t)c * nr_loc + mlo) * nm_loc) * 2;
      HP_CUDA(cudaMemcpyAsync(synthesis ? pb : pa, synthesis ? pa : pb, n * sizeof(double),
                              cudaMemcpyDeviceToDevice, st));
    }
}

// One transform.  The work is cut into the ring blocks of block set `set` (set 0: one block = everything):
//   synthesis:  Legendre(b) -> [P>1: exchange(b) -> unpack(b)] -> ring FFT(b) -> [host map: D2H copy(b)]
//   adjoint:    [host map: H2D copy(b)] -> ring FFT(b) -> [P>1: pack(b) -> exchange(b)] -> Legendre(b)
// each stage on its own stream, so block b's exchange / FFT / copy run behind the Legendre kernels of the
// next block.  All ranks of a P > 1 plan take the same path whatever kind of pointer they pass.
void hpsht_plan::run(bool synthesis, const double* in, double* out, int ncomp, bool async, cudaStream_t user) {
  HP_REQUIRE(ncomp == 1 || ncomp == 2, HPSHT_ERR_ARG, "Number of components must be 1 (spin 0) or 2 (spin 2).");
  HP_REQUIRE(ncomp <= max_ncomp, HPSHT_ERR_ARG, "plan was created with max_ncomp < ncomp");
  HP_REQUIRE(in && out, HPSHT_ERR_ARG, "null data pointer");
  HP_CUDA(cudaSetDevice(device));
  ensure_stages();
  if (times_pending && !async) finish_times();  // (a stream-ordered call never blocks: it just supersedes them)
  leg.reset_launches();
  fft.reset_launches();
  int64_t extra = 0;
  const size_t alm_d = (size_t)nalm_loc * ncomp * 2, map_d = (size_t)npix_loc * ncomp;
  const bool in_dev = is_device_pointer(in), out_dev = is_device_pointer(out);
  if (!tm) tm.reset(new Timer());
  cudaEvent_t* e = tm->ev;
  if (async) {
    // stream-ordered form: the work is ordered after what the caller has queued on `user` so far, and `user`
    // continues after it; nothing blocks the host
    HP_REQUIRE(in_dev && out_dev, HPSHT_ERR_ARG, "the stream-ordered calls take device pointers only");
    HP_CUDA(cudaEventRecord(ev_user, user));
    HP_CUDA(cudaStreamWaitEvent(stream, ev_user, 0));
  } else if (in_dev || out_dev) {
    // Device pointers: whoever produced `in` (or last touched `out`) did so on a stream this library does
    // not know -- e.g. the caller's framework filling the buffer a moment ago.  Wait for the device once on
    // entry; on return the call is complete.  (hpsht_*_async takes the caller's stream instead.)
    HP_CUDA(cudaDeviceSynchronize());
  }
  HP_CUDA(cudaEventRecord(e[0], stream));
  HP_CUDA(cudaEventRecord(ev_entry, stream));
  for (cudaStream_t s : {cstream, fstream, hstream, ostream != hstream ? ostream : (cudaStream_t) nullptr})
    if (s) HP_CUDA(cudaStreamWaitEvent(s, ev_entry, 0));

  const int spin = ncomp == 1 ? 0 : 2;
  const bool host_map = synthesis ? !out_dev : !in_dev;  // the map side travels over PCIe
  // P > 1, device pointers: block by block in the synthesis (the exchange of a block hides behind the next block's
  // Legendre kernels: -3 ... -9 % per transform on 2 and 8 GPUs), in one piece in the adjoint (there the per-block
  // launches cost more than the hidden exchange saves: +0 ... +5 %; profiles/r02_multi_gpu.md).  Host maps always go
  // block by block.  The choice depends on nothing rank-local besides the pointer kind, which all ranks must share.
  const int set = (HB > 1 && (host_map || (P > 1 && dev_blocks && synthesis))) ? 1 : 0;
  const int NB = leg.nblocks(set);
  auto fftof = [&](int b) -> RingFFT& { return set ? *fftb[(size_t)b] : fft; };
  auto plo = [&](int b) { return pix_at(blo(set, b, rank)); };
  auto phi = [&](int b) { return pix_at(bhi(set, b, rank)); };
  // Block b (pole -> equator) only touches the local m's [0, mie[b]) (m <= mlim of its rings): a host-side alm is
  // streamed in / out in the matching m-ranges [mie[b-1], mie[b]), so that only the polar block's share of the alm
  // copy (synthesis: before the first block; adjoint: after the last one) is not hidden behind Legendre work.
  std::vector<int> mie((size_t)NB, (int)nm_loc);
  for (int b = 0, hi = 0; b + 1 < NB; ++b) mie[(size_t)b] = hi = std::max(hi, leg.block_mi_end(set, b, spin));
  const bool split_alm = NB > 1 && mie[0] < (int)nm_loc && !getenv("HPSHT_NO_ALM_STREAM");
  auto a_at = [&](int mi) { return mi < (int)nm_loc ? (size_t)(mstart[(size_t)mi] + mval[(size_t)mi]) : (size_t)nalm_loc; };
  const bool stream_alm_in = synthesis && !in_dev && split_alm;
  const bool stream_alm_out = !synthesis && !out_dev && split_alm;
  if ((stream_alm_in || stream_alm_out) && (int)ev_alm.size() < NB) {
    const size_t have = ev_alm.size();
    ev_alm.resize((size_t)NB);
    for (size_t i = have; i < ev_alm.size(); ++i) HP_CUDA(cudaEventCreateWithFlags(&ev_alm[i], cudaEventDisableTiming));
  }

  const double* d_in = in;
  double* d_out = out;
  if (!in_dev) {
    DevBuf<double>& sb = synthesis ? stage_alm : stage_map;
    sb.ensure(synthesis ? alm_d : map_d);
    d_in = sb.p;
  }
  if (!out_dev) {
    DevBuf<double>& sb = synthesis ? stage_map : stage_alm;
    sb.ensure(synthesis ? map_d : alm_d);
    d_out = sb.p;
  }
  const int TB = 256;
  ExchGeom G;
  G.nm_tot = nm_tot;
  G.nr_loc = nr_loc;
  G.mmax = lmax;
  G.P = P;
  double* legA = P == 1 ? legF.p : almside.p;  // what the Legendre stage reads / writes

  if (synthesis) {
    // ---- alm in ----
    if (stream_alm_in) {
      for (int b = 0; b < NB; ++b) {
        const size_t lo = a_at(b ? mie[(size_t)b - 1] : 0), hi = a_at(mie[(size_t)b]);
        if (hi > lo)
          for (int c = 0; c < ncomp; ++c) {
            const size_t o = ((size_t)c * nalm_loc + lo) * 2;
            HP_CUDA(cudaMemcpyAsync(stage_alm.p + o, in + o, (hi - lo) * 2 * sizeof(double), cudaMemcpyHostToDevice, hstream));
          }
        HP_CUDA(cudaEventRecord(ev_alm[(size_t)b], hstream));
      }
    } else if (!in_dev) {
      HP_CUDA(cudaMemcpyAsync(stage_alm.p, in, alm_d * sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    HP_CUDA(cudaEventRecord(e[1], stream));
    // ---- alm2leg (src/sht.jl:85), block by block ----
    leg.alm2leg_zero(legA, P == 1 ? (int64_t)max_ncomp * nr_loc * nm_tot : aoff[(size_t)P], ncomp, stream);
    if (!stream_alm_in) leg.alm2leg_prep(d_in, nalm_loc, ncomp, 0, (int)nm_loc, stream);
    for (int b = 0; b < NB; ++b) {
      if (stream_alm_in) {  // the m-range this block is the first to need
        HP_CUDA(cudaStreamWaitEvent(stream, ev_alm[(size_t)b], 0));
        leg.alm2leg_prep(d_in, nalm_loc, ncomp, b ? mie[(size_t)b - 1] : 0, mie[(size_t)b], stream);
      }
      leg.alm2leg_block(set, b, legA, ncomp, stream, b == NB - 1);
      if (b == NB - 1) HP_CUDA(cudaEventRecord(e[2], stream));
      const int64_t j0 = blo(set, b, rank), nj = bhi(set, b, rank) - j0;
      if (P == 1) {
        if (b == NB - 1) HP_CUDA(cudaEventRecord(e[3], stream));
        fftof(b).leg2map(legF.p, d_out, npix_loc, ncomp, stream);  // leg2map (src/sht.jl:93)
        HP_CUDA(cudaEventRecord(ev_f[(size_t)b], stream));
        if (b == NB - 1) HP_CUDA(cudaEventRecord(e[4], stream));
      } else {
        // exchange of block b on the communication stream (src/sht.jl:90), unpack + ring FFT on the FFT stream
        HP_CUDA(cudaEventRecord(ev_blk[(size_t)b], stream));
        HP_CUDA(cudaStreamWaitEvent(cstream, ev_blk[(size_t)b], 0));
        exchange_block(true, ncomp, set, b, cstream);
        HP_CUDA(cudaEventRecord(ev_x[(size_t)b], cstream));
        HP_CUDA(cudaStreamWaitEvent(fstream, ev_x[(size_t)b], 0));
        extra += 1;
        if (nj > 0) {
          dim3 grid((unsigned)((nm_tot + TB - 1) / TB), (unsigned)(ncomp * nj));
          k_unpack<<<grid, TB, 0, fstream>>>(reinterpret_cast<const double2*>(mapside.p), d_boff.p,
                                             reinterpret_cast<double2*>(legF.p), G, j0, nj);
          HP_CUDA(cudaGetLastError());
          extra += 1;
        }
        if (b == NB - 1) HP_CUDA(cudaEventRecord(e[3], fstream));
        if (nj > 0) fftof(b).leg2map(legF.p, d_out, npix_loc, ncomp, fstream);
        HP_CUDA(cudaEventRecord(ev_f[(size_t)b], fstream));
        if (b == NB - 1) HP_CUDA(cudaEventRecord(e[4], fstream));
      }
    }
    // ---- map out ----
    if (!out_dev) {
      for (int b = 0; b < NB; ++b) {
        HP_CUDA(cudaStreamWaitEvent(ostream, ev_f[(size_t)b], 0));
        if (phi(b) > plo(b))
          for (int c = 0; c < ncomp; ++c) {
            const size_t o = (size_t)c * npix_loc + (size_t)plo(b);
            HP_CUDA(cudaMemcpyAsync(out + o, d_out + o, (size_t)(phi(b) - plo(b)) * sizeof(double),
                                    cudaMemcpyDeviceToHost, ostream));
          }
      }
      HP_CUDA(cudaEventRecord(ev_hdone, ostream));
      HP_CUDA(cudaStreamWaitEvent(stream, ev_hdone, 0));
    } else if (P > 1) {
      for (int b = 0; b < NB; ++b) HP_CUDA(cudaStreamWaitEvent(stream, ev_f[(size_t)b], 0));
    }
  } else {
    HP_CUDA(cudaEventRecord(e[1], stream));
    leg.leg2alm_begin(ncomp, stream);
    // once block b is through, the blocks still to come (b-1 .. 0) only touch the local m's below mie[b-1]: the
    // rows from there up to what has gone out already are final -- post-process them and copy them out on the
    // output stream while the next block computes
    int alm_done_from = (int)nm_loc;
    auto alm_out_after = [&](int b) {
      const int lo = b ? mie[(size_t)b - 1] : 0;
      if (lo >= alm_done_from) return;
      HP_CUDA(cudaEventRecord(ev_alm[(size_t)b], stream));
      HP_CUDA(cudaStreamWaitEvent(ostream, ev_alm[(size_t)b], 0));
      leg.leg2alm_end(d_out, nalm_loc, ncomp, ostream, lo, alm_done_from);
      for (int c = 0; c < ncomp; ++c) {
        const size_t o = ((size_t)c * nalm_loc + a_at(lo)) * 2;
        HP_CUDA(cudaMemcpyAsync(out + o, d_out + o, (a_at(alm_done_from) - a_at(lo)) * 2 * sizeof(double),
                                cudaMemcpyDeviceToHost, ostream));
      }
      alm_done_from = lo;
    };
    // the small equator-most blocks go first so that the compute starts early.  All input copies are queued before
    // anything else: with one copy stream (P > 1) an alm piece going out between two of them would otherwise hold
    // the next block's pixels back until the Legendre kernels of the current block are done.
    if (!in_dev)
      for (int b = NB - 1; b >= 0; --b) {
        if (phi(b) > plo(b))
          for (int c = 0; c < ncomp; ++c) {
            const size_t o = (size_t)c * npix_loc + (size_t)plo(b);
            HP_CUDA(cudaMemcpyAsync(stage_map.p + o, in + o, (size_t)(phi(b) - plo(b)) * sizeof(double),
                                    cudaMemcpyHostToDevice, hstream));
          }
        HP_CUDA(cudaEventRecord(ev_blk[(size_t)b], hstream));
      }
    for (int b = NB - 1; b >= 0; --b) {
      const int64_t j0 = blo(set, b, rank), nj = bhi(set, b, rank) - j0;
      if (P == 1) {
        if (!in_dev) HP_CUDA(cudaStreamWaitEvent(stream, ev_blk[(size_t)b], 0));
        fftof(b).map2leg(d_in, npix_loc, legF.p, ncomp, stream);  // map2leg (src/sht.jl:169)
        if (b == NB - 1) {
          HP_CUDA(cudaEventRecord(e[2], stream));
          HP_CUDA(cudaEventRecord(e[3], stream));
        }
      } else {
        if (!in_dev) HP_CUDA(cudaStreamWaitEvent(fstream, ev_blk[(size_t)b], 0));
        if (nj > 0) fftof(b).map2leg(d_in, npix_loc, legF.p, ncomp, fstream);
        if (b == NB - 1) HP_CUDA(cudaEventRecord(e[2], fstream));
        if (nj > 0) {
          dim3 grid((unsigned)((nm_tot + TB - 1) / TB), (unsigned)(ncomp * nj));
          k_pack<<<grid, TB, 0, fstream>>>(reinterpret_cast<const double2*>(legF.p), d_boff.p,
                                           reinterpret_cast<double2*>(mapside.p), G, j0, nj);
          HP_CUDA(cudaGetLastError());
          extra += 1;
        }
        HP_CUDA(cudaEventRecord(ev_f[(size_t)b], fstream));
        HP_CUDA(cudaStreamWaitEvent(cstream, ev_f[(size_t)b], 0));
        exchange_block(false, ncomp, set, b, cstream);  // src/sht.jl:174
        extra += 1;
        HP_CUDA(cudaEventRecord(ev_x[(size_t)b], cstream));
        if (b == NB - 1) HP_CUDA(cudaEventRecord(e[3], cstream));
        HP_CUDA(cudaStreamWaitEvent(stream, ev_x[(size_t)b], 0));
      }
      leg.leg2alm_block(set, b, legA, ncomp, stream, b == 0);  // leg2alm (src/sht.jl:178)
      if (stream_alm_out) alm_out_after(b);
    }
    if (!stream_alm_out) leg.leg2alm_end(d_out, nalm_loc, ncomp, stream, 0, (int)nm_loc);
    HP_CUDA(cudaEventRecord(e[4], stream));
    // ---- alm out ----
    if (stream_alm_out) {
      HP_CUDA(cudaEventRecord(ev_hdone, ostream));
      HP_CUDA(cudaStreamWaitEvent(stream, ev_hdone, 0));
    } else if (!out_dev) {
      HP_CUDA(cudaMemcpyAsync(out, d_out, alm_d * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
  }
  HP_CUDA(cudaEventRecord(e[5], stream));
  n_launch = leg.launches() + fft.launches() + extra;
  for (auto& f : fftb) {
    n_launch += f->launches();
    f->reset_launches();
  }
  last_synthesis = synthesis;
  times_pending = true;
  if (async) {
    HP_CUDA(cudaEventRecord(ev_done, stream));
    HP_CUDA(cudaStreamWaitEvent(user, ev_done, 0));
  } else {
    synchronize();
  }
}

void hpsht_plan::synchronize() {
  HP_CUDA(cudaSetDevice(device));
  for (cudaStream_t s : {stream, cstream, fstream, hstream, ostream != hstream ? ostream : (cudaStream_t) nullptr})
    if (s) HP_CUDA(cudaStreamSynchronize(s));
  if (times_pending) finish_times();
}

void hpsht_plan::finish_times() {
  times_pending = false;
  if (!tm) return;
  cudaEvent_t* e = tm->ev;
  HP_CUDA(cudaEventSynchronize(e[5]));
  auto el = [&](int a, int b) {
    float f = 0;
    if (cudaEventElapsedTime(&f, e[a], e[b]) != cudaSuccess) {
      cudaGetLastError();
      f = 0;
    }
    return (double)std::max(f, 0.f);
  };
  std::memset(ms, 0, sizeof(ms));
  ms[0] = el(0, 5);
  ms[5] = el(0, 1);
  ms[6] = el(4, 5);
  ms[3] = el(2, 3);
  if (last_synthesis) {
    ms[2] = el(1, 2);  // includes the O(nalm) prep kernel
    ms[4] = el(3, 4);
  } else {
    ms[4] = el(1, 2);
    ms[2] = el(3, 4);  // includes the post kernel
  }
  ms[1] = leg.main_kernel_ms();
}

// =================================================================================================
// stage-level caches (the reference calls the four Ducc0 functions with the same descriptors over
// and over from its solver loops; tables are rebuilt only when the descriptors change)
// =================================================================================================
namespace {

struct LegKey {
  int64_t lmax = -1, nm = 0, nring = 0;
  std::vector<int64_t> mval, mstart;
  std::vector<double> theta;
  bool operator==(const LegKey& o) const {
    return lmax == o.lmax && nm == o.nm && nring == o.nring && mval == o.mval && mstart == o.mstart &&
           theta == o.theta;
  }
};
struct FftKey {
  int64_t nm = 0, nring = 0;
  std::vector<int64_t> nphi, rstart;
  std::vector<double> phi0;
  bool operator==(const FftKey& o) const {
    return nm == o.nm && nring == o.nring && nphi == o.nphi && rstart == o.rstart && phi0 == o.phi0;
  }
};
struct StageCache {
  std::mutex mu;
  LegKey lkey;
  std::unique_ptr<LegendreStage> leg;
  FftKey fkey;
  std::unique_ptr<RingFFT> fft;
  DevBuf<double> d_alm, d_leg, d_map;
};
StageCache& cache() {
  static StageCache c;
  return c;
}

LegendreStage& get_leg_stage(StageCache& c, int64_t lmax, int64_t nm, const int64_t* mval,
                             const int64_t* mstart, int64_t nring, const double* theta) {
  LegKey k;
  k.lmax = lmax;
  k.nm = nm;
  k.nring = nring;
  k.mval.assign(mval, mval + nm);
  k.mstart.assign(mstart, mstart + nm);
  k.theta.assign(theta, theta + nring);
  if (!c.leg || !(c.lkey == k)) {
    c.leg.reset(new LegendreStage());
    LegLayout lay;  // reference alm-side layout (nm, nring, ncomp): leg[(comp*nring + ir)*nm + mi]
    lay.a0.assign((size_t)nring, 0);
    lay.nr.assign((size_t)nring, nring);
    lay.j.resize((size_t)nring);
    for (int64_t s = 0; s < nring; ++s) lay.j[(size_t)s] = s;
    c.leg->setup(lmax, mval, mstart, nm, nring, theta, lay, 2, 0);
    c.lkey = k;
  }
  return *c.leg;
}

RingFFT& get_fft_stage(StageCache& c, int64_t nm, int64_t nring, const int64_t* nphi,
                       const double* phi0, const int64_t* rstart) {
  FftKey k;
  k.nm = nm;
  k.nring = nring;
  k.nphi.assign(nphi, nphi + nring);
  k.rstart.assign(rstart, rstart + nring);
  k.phi0.assign(phi0, phi0 + nring);
  if (!c.fft || !(c.fkey == k)) {
    c.fft.reset(new RingFFT());
    c.fft->setup(nring, nphi, phi0, rstart, nm, 0);
    c.fkey = k;
  }
  return *c.fft;
}

int64_t alm_extent(int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart) {
  int64_t ext = 0;
  for (int64_t i = 0; i < nm; ++i) {
    HP_REQUIRE(mval[i] >= 0 && mval[i] <= lmax, HPSHT_ERR_ARG, "mval out of range");
    HP_REQUIRE(mstart[i] + mval[i] >= 0, HPSHT_ERR_ARG, "mstart points before the array");
    ext = std::max(ext, mstart[i] + lmax + 1);
  }
  return ext;
}

// stage a (possibly host) input on the device; returns the device pointer to use
const double* stage_in(const double* p, size_t n, DevBuf<double>& buf) {
  if (is_device_pointer(p)) {
    // produced on a stream this library does not know (see hpsht_plan::run): wait for the device
    HP_CUDA(cudaDeviceSynchronize());
    return p;
  }
  buf.ensure(n);
  HP_CUDA(cudaMemcpy(buf.p, p, n * sizeof(double), cudaMemcpyHostToDevice));
  return buf.p;
}
double* stage_out(double* p, size_t n, DevBuf<double>& buf) {
  if (is_device_pointer(p)) {
    HP_CUDA(cudaDeviceSynchronize());  // earlier work of the caller on this buffer must not land after ours
    return p;
  }
  buf.ensure(n);
  return buf.p;
}
void finish_out(double* user, const double* dev, size_t n) {
  HP_CUDA(cudaDeviceSynchronize());
  if (user != dev) HP_CUDA(cudaMemcpy(user, dev, n * sizeof(double), cudaMemcpyDeviceToHost));
}

}  // namespace

#include "shards.cuh"

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

int hpsht_version(void) { return HPSHT_VERSION_MAJOR * 100 + HPSHT_VERSION_MINOR; }
const char* hpsht_last_error(void) { return g_last_error.c_str(); }
int hpsht_device_count(int* count) {
  return guarded([&] {
    HP_REQUIRE(count, HPSHT_ERR_ARG, "null pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    *count = n;
  });
}

// ---- RR bookkeeping ----
#define RR_CHECK(rank, size)                                                                     \
  HP_REQUIRE((size) >= 1 && (rank) >= 0 && (rank) < (size), HPSHT_ERR_ARG,                       \
             std::to_string(rank) + " can not exceed communicator size")

int hpsht_rr_nm(int64_t mmax, int task_rank, int c_size, int64_t* nm) {
  return guarded([&] {
    RR_CHECK(task_rank, c_size);
    HP_REQUIRE(nm && mmax >= 0, HPSHT_ERR_ARG, "bad argument");
    *nm = rr_nm(mmax, task_rank, c_size);
  });
}
int hpsht_rr_mval(int64_t mmax, int task_rank, int c_size, int64_t* mval) {
  return guarded([&] {
    RR_CHECK(task_rank, c_size);
    HP_REQUIRE(mval && mmax >= 0, HPSHT_ERR_ARG, "bad argument");
    std::vector<int64_t> v = rr_mval(mmax, task_rank, c_size);
    std::copy(v.begin(), v.end(), mval);
  });
}
int hpsht_rr_m_tasks(int64_t mmax, int c_size, int64_t* task) {
  return guarded([&] {
    HP_REQUIRE(task && mmax >= 0 && c_size >= 1, HPSHT_ERR_ARG, "bad argument");
    for (int64_t m = 0; m <= mmax; ++m) task[m] = m % c_size;
  });
}
int hpsht_rr_mstart(int64_t lmax, int64_t stride, int64_t nm, const int64_t* mval, int64_t* mstart) {
  return guarded([&] {
    HP_REQUIRE(mval && mstart && nm >= 0, HPSHT_ERR_ARG, "bad argument");
    std::vector<int64_t> mv(mval, mval + nm);
    std::vector<int64_t> v = rr_mstart0(lmax, stride, mv);
    std::copy(v.begin(), v.end(), mstart);
  });
}
int hpsht_rr_nrings(int64_t eq_idx, int task_rank, int c_size, int64_t* nrings) {
  return guarded([&] {
    RR_CHECK(task_rank, c_size);
    HP_REQUIRE(nrings && eq_idx >= 1, HPSHT_ERR_ARG, "bad argument");
    *nrings = rr_nrings(eq_idx, task_rank, c_size);
  });
}
int hpsht_rr_rindexes(int64_t eq_idx, int task_rank, int c_size, int64_t* rings) {
  return guarded([&] {
    RR_CHECK(task_rank, c_size);
    HP_REQUIRE(rings && eq_idx >= 1, HPSHT_ERR_ARG, "bad argument");
    std::vector<int64_t> v = rr_rindexes(eq_idx, task_rank, c_size);
    std::copy(v.begin(), v.end(), rings);
  });
}
int hpsht_rr_rindexes_tot(int64_t eq_idx, int c_size, int64_t* rings) {
  return guarded([&] {
    HP_REQUIRE(rings && eq_idx >= 1 && c_size >= 1, HPSHT_ERR_ARG, "bad argument");
    std::vector<int64_t> v = rr_rindexes_tot(eq_idx, c_size);
    std::copy(v.begin(), v.end(), rings);
  });
}
int hpsht_rr_a2a_counts(int64_t mmax, int64_t eq_idx, int task_rank, int c_size, int64_t* send,
                        int64_t* recv) {
  return guarded([&] {
    RR_CHECK(task_rank, c_size);
    HP_REQUIRE(send && recv, HPSHT_ERR_ARG, "bad argument");
    const int64_t loc_nm = rr_nm(mmax, task_rank, c_size), loc_nr = rr_nrings(eq_idx, task_rank, c_size);
    for (int t = 0; t < c_size; ++t) {
      send[t] = loc_nm * rr_nrings(eq_idx, t, c_size);
      recv[t] = loc_nr * rr_nm(mmax, t, c_size);
    }
  });
}
int hpsht_healpix_ring(int64_t nside, int64_t ring, int64_t* nphi, double* theta, double* phi0,
                       int64_t* first_pix) {
  return guarded([&] {
    HP_REQUIRE(nside >= 1 && ring >= 1 && ring <= 4 * nside - 1, HPSHT_ERR_ARG, "ring out of range");
    RingGeom g = healpix_ring(nside, ring);
    if (nphi) *nphi = g.nphi;
    if (theta) *theta = g.theta;
    if (phi0) *phi0 = g.phi0;
    if (first_pix) *first_pix = g.first_pix;
  });
}

// ---- stage level ----
int hpsht_alm2leg(const double* alm, int64_t alm_cstride, double* leg, int spin, int64_t lmax,
                  int64_t nm, const int64_t* mval, const int64_t* mstart, int64_t lstride,
                  int64_t nring, const double* theta, int /*nthreads*/) {
  return guarded([&] {
    HP_REQUIRE(spin == 0 || spin == 2, HPSHT_ERR_UNSUPPORTED, "only spin 0 and spin 2 are supported");
    HP_REQUIRE(lstride == 1, HPSHT_ERR_UNSUPPORTED, "lstride must be 1");
    HP_REQUIRE(alm && leg && mval && mstart && theta && nm > 0 && nring > 0 && lmax >= 0,
               HPSHT_ERR_ARG, "bad argument");
    require_device();
    const int ncomp = spin == 0 ? 1 : 2;
    const int64_t ext = alm_extent(lmax, nm, mval, mstart);
    HP_REQUIRE(ncomp == 1 || alm_cstride >= ext, HPSHT_ERR_ARG, "alm_cstride too small");
    StageCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    LegendreStage& st = get_leg_stage(c, lmax, nm, mval, mstart, nring, theta);
    const size_t nalm = (size_t)(ncomp == 1 ? ext : alm_cstride + ext) * 2;
    const size_t nleg = (size_t)ncomp * nring * nm * 2;
    const double* d_alm = stage_in(alm, nalm, c.d_alm);
    double* d_leg = stage_out(leg, nleg, c.d_leg);
    st.alm2leg(d_alm, alm_cstride, d_leg, (int64_t)ncomp * nring * nm, ncomp, 0);
    finish_out(leg, d_leg, nleg);
  });
}

int hpsht_leg2alm(const double* leg, double* alm, int64_t alm_cstride, int spin, int64_t lmax,
                  int64_t nm, const int64_t* mval, const int64_t* mstart, int64_t lstride,
                  int64_t nring, const double* theta, int /*nthreads*/) {
  return guarded([&] {
    HP_REQUIRE(spin == 0 || spin == 2, HPSHT_ERR_UNSUPPORTED, "only spin 0 and spin 2 are supported");
    HP_REQUIRE(lstride == 1, HPSHT_ERR_UNSUPPORTED, "lstride must be 1");
    HP_REQUIRE(alm && leg && mval && mstart && theta && nm > 0 && nring > 0 && lmax >= 0,
               HPSHT_ERR_ARG, "bad argument");
    require_device();
    const int ncomp = spin == 0 ? 1 : 2;
    const int64_t ext = alm_extent(lmax, nm, mval, mstart);
    HP_REQUIRE(ncomp == 1 || alm_cstride >= ext, HPSHT_ERR_ARG, "alm_cstride too small");
    StageCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    LegendreStage& st = get_leg_stage(c, lmax, nm, mval, mstart, nring, theta);
    const size_t nalm = (size_t)(ncomp == 1 ? ext : alm_cstride + ext) * 2;
    const size_t nleg = (size_t)ncomp * nring * nm * 2;
    const double* d_leg = stage_in(leg, nleg, c.d_leg);
    const bool host_out = !is_device_pointer(alm);
    double* d_alm = stage_out(alm, nalm, c.d_alm);
    // entries of the caller's array that are not a_lm of an owned m must survive untouched
    if (host_out) HP_CUDA(cudaMemcpy(d_alm, alm, nalm * sizeof(double), cudaMemcpyHostToDevice));
    st.leg2alm(d_leg, d_alm, alm_cstride, ncomp, 0);
    finish_out(alm, d_alm, nalm);
  });
}

int hpsht_leg2map(const double* leg, double* map, int64_t map_cstride, int ncomp, int64_t nm,
                  int64_t nring, const int64_t* nphi, const double* phi0, const int64_t* ringstart,
                  int64_t pixstride, int /*nthreads*/) {
  return guarded([&] {
    HP_REQUIRE(pixstride == 1, HPSHT_ERR_UNSUPPORTED, "pixstride must be 1");
    HP_REQUIRE(ncomp == 1 || ncomp == 2, HPSHT_ERR_ARG, "ncomp must be 1 or 2");
    HP_REQUIRE(leg && map && nphi && phi0 && ringstart && nm > 0 && nring > 0, HPSHT_ERR_ARG,
               "bad argument");
    require_device();
    int64_t ext = 0;
    for (int64_t i = 0; i < nring; ++i) {
      HP_REQUIRE(ringstart[i] >= 0, HPSHT_ERR_ARG, "negative ringstart");
      ext = std::max(ext, ringstart[i] + nphi[i]);
    }
    HP_REQUIRE(ncomp == 1 || map_cstride >= ext, HPSHT_ERR_ARG, "map_cstride too small");
    StageCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    RingFFT& f = get_fft_stage(c, nm, nring, nphi, phi0, ringstart);
    const size_t nleg = (size_t)ncomp * nring * nm * 2;
    const size_t nmap = (size_t)(ncomp == 1 ? ext : map_cstride + ext);
    const double* d_leg = stage_in(leg, nleg, c.d_leg);
    const bool host_out = !is_device_pointer(map);
    double* d_map = stage_out(map, nmap, c.d_map);
    if (host_out) HP_CUDA(cudaMemcpy(d_map, map, nmap * sizeof(double), cudaMemcpyHostToDevice));
    f.leg2map(d_leg, d_map, map_cstride, ncomp, 0);
    finish_out(map, d_map, nmap);
  });
}

int hpsht_map2leg(const double* map, int64_t map_cstride, double* leg, int ncomp, int64_t nm,
                  int64_t nring, const int64_t* nphi, const double* phi0, const int64_t* ringstart,
                  int64_t pixstride, int /*nthreads*/) {
  return guarded([&] {
    HP_REQUIRE(pixstride == 1, HPSHT_ERR_UNSUPPORTED, "pixstride must be 1");
    HP_REQUIRE(ncomp == 1 || ncomp == 2, HPSHT_ERR_ARG, "ncomp must be 1 or 2");
    HP_REQUIRE(leg && map && nphi && phi0 && ringstart && nm > 0 && nring > 0, HPSHT_ERR_ARG,
               "bad argument");
    require_device();
    int64_t ext = 0;
    for (int64_t i = 0; i < nring; ++i) {
      HP_REQUIRE(ringstart[i] >= 0, HPSHT_ERR_ARG, "negative ringstart");
      ext = std::max(ext, ringstart[i] + nphi[i]);
    }
    HP_REQUIRE(ncomp == 1 || map_cstride >= ext, HPSHT_ERR_ARG, "map_cstride too small");
    StageCache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    RingFFT& f = get_fft_stage(c, nm, nring, nphi, phi0, ringstart);
    const size_t nleg = (size_t)ncomp * nring * nm * 2;
    const size_t nmap = (size_t)(ncomp == 1 ? ext : map_cstride + ext);
    const double* d_map = stage_in(map, nmap, c.d_map);
    double* d_leg = stage_out(leg, nleg, c.d_leg);
    f.map2leg(d_map, map_cstride, d_leg, ncomp, 0);
    finish_out(leg, d_leg, nleg);
  });
}

// ---- fused plan ----
int hpsht_nccl_unique_id(unsigned char id[HPSHT_NCCL_ID_BYTES]) {
  return guarded([&] {
    HP_REQUIRE(id, HPSHT_ERR_ARG, "null pointer");
    static_assert(sizeof(ncclUniqueId) == HPSHT_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    HP_NCCL(NcclApi::get().GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
  });
}

int hpsht_plan_create(hpsht_plan** plan, int64_t nside, int64_t lmax, int rank, int n_ranks,
                      const unsigned char* nccl_id, int device, int max_ncomp) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan pointer");
    *plan = nullptr;
    HP_REQUIRE(n_ranks >= 1, HPSHT_ERR_ARG, "n_ranks must be >= 1");
    HP_REQUIRE(max_ncomp == 1 || max_ncomp == 2, HPSHT_ERR_ARG, "max_ncomp must be 1 or 2");
    HP_REQUIRE(n_ranks == 1 || nccl_id, HPSHT_ERR_ARG, "nccl_id required when n_ranks > 1");
    std::unique_ptr<hpsht_plan> p(new hpsht_plan());
    p->nside = nside;
    p->lmax = lmax;
    p->rank = rank;
    p->P = n_ranks;
    p->max_ncomp = max_ncomp;
    p->build_shard();  // pure host work: argument errors surface before any CUDA call
    require_device();
    if (device < 0) HP_CUDA(cudaGetDevice(&device));
    p->device = device;
    HP_CUDA(cudaSetDevice(device));
    HP_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    if (n_ranks > 1) {
      ncclUniqueId u;
      std::memcpy(&u, nccl_id, sizeof(u));
      HP_NCCL(NcclApi::get().CommInitRank(&p->comm, n_ranks, u, rank));
    }
    *plan = p.release();
  });
}

int hpsht_plan_set_geometry(hpsht_plan* plan, const double* thetatot, const double* phi0_loc) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    HP_REQUIRE(!plan->stages_ready, HPSHT_ERR_ARG, "geometry must be set before the first transform");
    if (thetatot) plan->thetatot.assign(thetatot, thetatot + plan->nr_tot);
    if (phi0_loc) plan->phi0.assign(phi0_loc, phi0_loc + plan->nr_loc);
  });
}

int hpsht_plan_destroy(hpsht_plan* plan) {
  return guarded([&] {
    if (!plan) return;
    cudaSetDevice(plan->device);
    if (plan->ostream == plan->hstream) plan->ostream = nullptr;
    for (cudaStream_t st : {plan->stream, plan->cstream, plan->fstream, plan->hstream, plan->ostream})
      if (st) cudaStreamSynchronize(st);
    if (plan->comm) NcclApi::get().CommDestroy(plan->comm);
    for (auto* v : {&plan->ev_blk, &plan->ev_x, &plan->ev_f, &plan->ev_alm})
      for (auto& e : *v) cudaEventDestroy(e);
    for (cudaEvent_t e : {plan->ev_hdone, plan->ev_a0, plan->ev_a1, plan->ev_entry, plan->ev_user, plan->ev_done})
      if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {plan->cstream, plan->fstream, plan->hstream, plan->ostream, plan->stream})
      if (st) cudaStreamDestroy(st);
    delete plan;
  });
}

int hpsht_plan_sizes(const hpsht_plan* plan, int64_t* nm_loc, int64_t* nalm_loc, int64_t* nring_loc,
                     int64_t* npix_loc, int64_t* nring_tot) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    if (nm_loc) *nm_loc = plan->nm_loc;
    if (nalm_loc) *nalm_loc = plan->nalm_loc;
    if (nring_loc) *nring_loc = plan->nr_loc;
    if (npix_loc) *npix_loc = plan->npix_loc;
    if (nring_tot) *nring_tot = plan->nr_tot;
  });
}
int hpsht_plan_alm_info(const hpsht_plan* plan, int64_t* mval, int64_t* mstart) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    if (mval) std::copy(plan->mval.begin(), plan->mval.end(), mval);
    if (mstart) std::copy(plan->mstart.begin(), plan->mstart.end(), mstart);
  });
}
int hpsht_plan_map_info(const hpsht_plan* plan, int64_t* rings, int64_t* rstart, int64_t* nphi,
                        double* theta, double* phi0, double* thetatot) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    if (rings) std::copy(plan->rings.begin(), plan->rings.end(), rings);
    if (rstart) std::copy(plan->rstart.begin(), plan->rstart.end(), rstart);
    if (nphi) std::copy(plan->nphi.begin(), plan->nphi.end(), nphi);
    if (theta) std::copy(plan->theta.begin(), plan->theta.end(), theta);
    if (phi0) std::copy(plan->phi0.begin(), plan->phi0.end(), phi0);
    if (thetatot) std::copy(plan->thetatot.begin(), plan->thetatot.end(), thetatot);
  });
}

int hpsht_alm2map(hpsht_plan* plan, const double* alm, double* map, int ncomp, int /*nthreads*/) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    plan->run(true, alm, map, ncomp, false, nullptr);
  });
}
int hpsht_adjoint_alm2map(hpsht_plan* plan, const double* map, double* alm, int ncomp,
                          int /*nthreads*/) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    plan->run(false, map, alm, ncomp, false, nullptr);
  });
}
int hpsht_alm2map_async(hpsht_plan* plan, const double* d_alm, double* d_map, int ncomp, void* cuda_stream) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    plan->run(true, d_alm, d_map, ncomp, true, static_cast<cudaStream_t>(cuda_stream));
  });
}
int hpsht_adjoint_alm2map_async(hpsht_plan* plan, const double* d_map, double* d_alm, int ncomp, void* cuda_stream) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    plan->run(false, d_map, d_alm, ncomp, true, static_cast<cudaStream_t>(cuda_stream));
  });
}
int hpsht_ring_block_ranges(int64_t nside, int64_t lmax, int n_ranks, int* nblocks, int64_t* lo, int64_t* hi,
                            int capacity) {
  return guarded([&] {
    HP_REQUIRE(nblocks && nside >= 1 && lmax >= 0 && n_ranks >= 1, HPSHT_ERR_ARG, "bad argument");
    const int64_t eq = 2 * nside, nr_tot = 2 * eq - 1;
    std::vector<int64_t> tot = rr_rindexes_tot(eq, n_ranks), nr_of((size_t)n_ranks);
    std::vector<double> theta((size_t)nr_tot), w((size_t)nr_tot);
    for (size_t i = 0; i < tot.size(); ++i) {
      const RingGeom g = healpix_ring(nside, tot[i]);
      theta[i] = g.theta;
      w[i] = (double)g.nphi;
    }
    for (int t = 0; t < n_ranks; ++t) nr_of[(size_t)t] = rr_nrings(eq, t, n_ranks);
    const std::vector<PairHost> pairs = build_pairs(lmax, nr_tot, theta.data());
    std::vector<int> bp = cut_ring_blocks(pairs, wanted_ring_blocks(nside, n_ranks), w.data(), equal_ring_blocks(n_ranks));
    std::vector<int64_t> l, h;
    if (bp.size() <= 2 || !ring_block_ranges(pairs, bp, nr_of, l, h)) {
      bp = {0, (int)pairs.size()};
      l.assign((size_t)n_ranks, 0);
      h = nr_of;
    }
    *nblocks = (int)bp.size() - 1;
    HP_REQUIRE(!lo || !hi || (int64_t)l.size() <= capacity, HPSHT_ERR_ARG, "capacity too small");
    if (lo && hi) {
      std::copy(l.begin(), l.end(), lo);
      std::copy(h.begin(), h.end(), hi);
    }
  });
}
int hpsht_plan_set_dev_blocks(hpsht_plan* plan, int on) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    plan->dev_blocks = on != 0;
  });
}
int hpsht_plan_synchronize(hpsht_plan* plan) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    plan->synchronize();
  });
}

// ---- shard bookkeeping and algebra on the device (shards.cuh) ----
int hpsht_plan_scatter_alm(hpsht_plan* plan, const double* global_alm, double* alm, int ncomp, int root) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    shard_exchange(*plan, 0, 0, global_alm, alm, ncomp, root);
  });
}
int hpsht_plan_gather_alm(hpsht_plan* plan, const double* alm, double* global_alm, int ncomp, int root) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    shard_exchange(*plan, 0, 1, alm, global_alm, ncomp, root);
  });
}
int hpsht_plan_scatter_map(hpsht_plan* plan, const double* global_map, double* map, int ncomp, int root) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    shard_exchange(*plan, 1, 0, global_map, map, ncomp, root);
  });
}
int hpsht_plan_gather_map(hpsht_plan* plan, const double* map, double* global_map, int ncomp, int root) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    shard_exchange(*plan, 1, 1, map, global_map, ncomp, root);
  });
}
int hpsht_plan_dot(hpsht_plan* plan, const double* alm1, const double* alm2, double* result) {
  return guarded([&] {
    HP_REQUIRE(plan && result, HPSHT_ERR_ARG, "null pointer");
    *result = shard_dot(*plan, alm1, alm2);
  });
}
int hpsht_plan_almxfl(hpsht_plan* plan, double* alm, const double* fl, int64_t nfl, int ncomp) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    shard_almxfl(*plan, alm, fl, nfl, ncomp);
  });
}
int hpsht_plan_alm2cl(hpsht_plan* plan, const double* alm1, const double* alm2, double* cl) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    shard_alm2cl(*plan, alm1, alm2, cl);
  });
}

int hpsht_plan_timings(const hpsht_plan* plan, double ms[HPSHT_NTIMES]) {
  return guarded([&] {
    HP_REQUIRE(plan && ms, HPSHT_ERR_ARG, "null pointer");
    if (plan->times_pending) const_cast<hpsht_plan*>(plan)->synchronize();
    std::memcpy(ms, plan->ms, sizeof(plan->ms));
  });
}
int hpsht_plan_launches(const hpsht_plan* plan, int64_t* n) {
  return guarded([&] {
    HP_REQUIRE(plan && n, HPSHT_ERR_ARG, "null pointer");
    *n = plan->n_launch;
  });
}
int hpsht_plan_legendre_steps(const hpsht_plan* plan, int spin, double* steps_culled,
                              double* steps_dense) {
  return guarded([&] {
    HP_REQUIRE(plan, HPSHT_ERR_ARG, "null plan");
    // pure host arithmetic on the shard description (no device needed)
    std::vector<PairHost> pairs = build_pairs(plan->lmax, plan->nr_tot, plan->thetatot.data());
    double c = 0, d = 0;
    for (const PairHost& p : pairs)
      for (int64_t m : plan->mval) {
        const double w = (double)(plan->lmax - m + 1);
        d += w;
        if (m <= (spin ? p.mlim2 : p.mlim0)) c += w;
      }
    if (steps_culled) *steps_culled = c;
    if (steps_dense) *steps_dense = d;
  });
}

// ---- device memory helpers ----
int hpsht_device_malloc(void** dptr, size_t bytes) {
  return guarded([&] {
    HP_REQUIRE(dptr, HPSHT_ERR_ARG, "null pointer");
    require_device();
    HP_CUDA(cudaMalloc(dptr, bytes));
  });
}
int hpsht_device_free(void* dptr) {
  return guarded([&] { HP_CUDA(cudaFree(dptr)); });
}
int hpsht_host_malloc(void** hptr, size_t bytes) {
  return guarded([&] {
    HP_REQUIRE(hptr, HPSHT_ERR_ARG, "null pointer");
    require_device();
    HP_CUDA(cudaMallocHost(hptr, bytes));
  });
}
int hpsht_host_free(void* hptr) {
  return guarded([&] { HP_CUDA(cudaFreeHost(hptr)); });
}
int hpsht_memcpy_h2d(void* dptr, const void* hptr, size_t bytes) {
  return guarded([&] { HP_CUDA(cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice)); });
}
int hpsht_memcpy_d2h(void* hptr, const void* dptr, size_t bytes) {
  return guarded([&] { HP_CUDA(cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost)); });
}
int hpsht_device_synchronize(void) {
  return guarded([&] { HP_CUDA(cudaDeviceSynchronize()); });
}

int hpsht_measure_fp64_peak(double* tflops, double* seconds_run) {
  return guarded([&] {
    HP_REQUIRE(tflops, HPSHT_ERR_ARG, "null pointer");
    require_device();
    cudaDeviceProp prop;
    int dev = 0;
    HP_CUDA(cudaGetDevice(&dev));
    HP_CUDA(cudaGetDeviceProperties(&prop, dev));
    DevBuf<double> out;
    out.alloc(1);
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 20000;
    cudaEvent_t a, b;
    HP_CUDA(cudaEventCreate(&a));
    HP_CUDA(cudaEventCreate(&b));
    k_dfma_peak<<<blocks, threads>>>(out.p, 200, 0.5);  // warm-up
    double best = 0, total = 0;
    for (int rep = 0; rep < 5; ++rep) {
      HP_CUDA(cudaEventRecord(a));
      k_dfma_peak<<<blocks, threads>>>(out.p, iters, 0.5);
      HP_CUDA(cudaEventRecord(b));
      HP_CUDA(cudaEventSynchronize(b));
      float msf = 0;
      HP_CUDA(cudaEventElapsedTime(&msf, a, b));
      const double flop = 2.0 * 64.0 * (double)iters * threads * (double)blocks;
      best = std::max(best, flop / (msf * 1e-3) / 1e12);
      total += msf * 1e-3;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    if (seconds_run) *seconds_run = total;
  });
}

}  // extern "C"
