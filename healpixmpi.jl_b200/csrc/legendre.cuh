// legendre.cuh — host-visible interface of the Legendre stage (legendre.cu).
//
// Replaces Ducc0.Sht.alm2leg! / leg2alm! (reference call sites src/sht.jl:85 and src/sht.jl:178).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "common.hpp"
#include "legendre_tables.hpp"

namespace hpsht {

// Steps per TMA stage (tables are padded to this).  Must be a multiple of LEG_GROUP.
constexpr int LEG_CHUNK = 64;
// Steps between range (rescale) checks.
constexpr int LEG_GROUP = 8;

// One north/south ring pair as the kernels see it.
struct PairHost {
  double cth, sth, sh, ch;   // cos/sin(theta_N), sin/cos(theta_N / 2)
  double x2, s2, y2;         // cos^2, sin^2, 2 sin^2(theta_N / 2): each rounded once from long double
  double theta;              // colatitude of the north ring (an unpaired south ring: its mirror image)
  int64_t slotN, slotS;      // ring slots in the caller's theta list (slotS = -1: unpaired)
  int mlim0, mlim2;          // m > mlim  => culled
};

// Leg addressing.  Ring slot `slot` belongs to a block starting at a0[slot] that holds nr[slot] rings,
// the slot being ring j[slot] of that block.  Element (slot, mi, comp) lives at
//     a0 + (comp * nr + j) * nm + mi          [complex elements]
// With a0 = 0, nr = nring, j = slot this is the reference's (nm, nring, ncomp) column-major array; with one
// block per destination rank it is the reference's all-to-all send layout (src/sht.jl:21,32).
struct LegLayout {
  std::vector<int64_t> a0, nr, j;      // per ring slot
};

struct WorkItem {
  int mi;        // local m index
  int p0;        // first pair
  int np;        // number of pairs (synthesis: <= one tile; adjoint: a range walked tile by tile)
  int cls;       // accuracy class of the item (LegClass); a standard adjoint item starting in Y may run on into X
};

// launch kinds of one ring block: the two standard forms share a kernel, the difference forms have their own
enum LegKind { LK_STD = 0, LK_DP = 1, LK_DE = 2, LK_COUNT = 3 };

// Everything the Legendre kernels need for one (theta list, m list, layout).
class LegendreStage {
 public:
  LegendreStage() = default;
  ~LegendreStage();
  LegendreStage(const LegendreStage&) = delete;
  LegendreStage& operator=(const LegendreStage&) = delete;

  // theta: nring colatitudes in leg order; mval/mstart: local m's and 0-based alm offsets.
  void setup(int64_t lmax, const int64_t* mval, const int64_t* mstart, int64_t nm, int64_t nring,
             const double* theta, const LegLayout& layout, int max_ncomp, cudaStream_t stream);

  // Ring blocks: the pole -> equator pair list cut into <= `want` contiguous ranges.  For the reference's
  // round-robin ring order a block's rings are a contiguous local range on every rank, so the caller can
  // pipeline  Legendre(block) -> exchange(block) -> ring FFT(block) -> device-to-host copy(block)  and the
  // reverse for the adjoint.  Block set 0 always is the single block holding every pair; setup_blocks
  // builds set 1 and returns its number of blocks.  Call right after setup().
  int setup_blocks(int want, const double* slot_weight, bool equal = false);
  int nblocks(int set) const { return (int)blk_pair_[set].size() - 1; }
  std::vector<int64_t> block_slot_list(int set, int b) const;   // every ring slot of the block
  // number of local m's (a prefix) that block b touches at all
  int block_mi_end(int set, int b, int spin) const;

  // synthesis: alm2leg_zero (clear the leg array), alm2leg_prep over the local m's (in ranges as the alm
  // arrives), then every block once in any order
  void alm2leg_zero(double* d_leg, int64_t leg_elems, int ncomp, cudaStream_t stream);
  void alm2leg_prep(const double* d_alm, int64_t alm_cstride, int ncomp, int mi0, int mi1, cudaStream_t stream);
  void alm2leg_block(int set, int b, double* d_leg, int ncomp, cudaStream_t stream, bool last);
  // adjoint: begin, every block once (same stream), end = add the partial-sum rows and post-process the
  // local m's [mi0, mi1)
  void leg2alm_begin(int ncomp, cudaStream_t stream);
  void leg2alm_block(int set, int b, const double* d_leg, int ncomp, cudaStream_t stream, bool last);
  void leg2alm_end(double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream, int mi0 = 0, int mi1 = -1);
  // whole transforms on block set 0
  void alm2leg(const double* d_alm, int64_t alm_cstride, double* d_leg, int64_t leg_elems, int ncomp,
               cudaStream_t stream);
  void leg2alm(const double* d_leg, double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream);

  double steps_culled(int spin) const { return spin ? culled2_ : culled0_; }
  double steps_dense() const { return dense_; }
  int64_t launches() const { return launches_; }
  void reset_launches() { launches_ = 0; }
  // device time between the first and the last main (recurrence) kernel of the most recent transform; call
  // after a stream sync
  double main_kernel_ms() const;
  const ClassRanges& classes(int spin) const { return spin ? cr2_ : cr0_; }

 private:
  void ensure_spin(int spin, cudaStream_t stream);
  void fill_params(struct LegParams& P, int spin) const;
  void build_items(int spin, int set, cudaStream_t stream);
  void fork(cudaStream_t stream);
  void join(cudaStream_t stream);

  int64_t lmax_ = 0, nm_ = 0, nring_ = 0, npairs_ = 0;
  int max_ncomp_ = 1;
  std::vector<int64_t> mval_, mstart_;
  std::vector<PairHost> pairs_;
  LegLayout layout_;
  ClassRanges cr0_, cr2_;
  bool have_[2] = {false, false};
  double culled0_ = 0, culled2_ = 0, dense_ = 0;
  int64_t launches_ = 0;
  bool first_block_ = true;   // the next block launch is the first of its transform (timing event)

  // device: geometry
  DevBuf<double> d_cth_, d_sth_, d_sh_, d_ch_, d_x2_, d_s2_, d_ny2_;
  DevBuf<int> d_mlim0_, d_mlim2_;
  DevBuf<int64_t> d_a0_, d_nrr_, d_jN_, d_jS_;
  DevBuf<int64_t> d_mval_, d_mstart_;
  // device: tables, per spin index (0: spin 0, 1: spin 2)
  DevBuf<int64_t> d_off_[2], d_nstep_[2], d_coff_[2][LC_COUNT];
  DevBuf<double> d_coef_[2][LC_COUNT], d_coefa2_, d_inv_[2], d_k0_[2], d_k1_[2];   // d_inv_: invs0 / sig2
  DevBuf<double> d_stream_[2];   // prepared alm streams (4 doubles per step)
  std::vector<int64_t> nstep_h_[2], off_h_[2];
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, ev_fork_ = nullptr, ev_join_ = nullptr;
  cudaStream_t side_ = nullptr;   // the (small) difference-form launches of a block run next to the standard one
  // adjoint partial sums (4 doubles per step): rows_std_ rows of all local m's used by the standard and the
  // equatorial launches (same stream), then rows_dp_ rows of the local m's [0, mi_dp_) used by the polar launch
  DevBuf<double> d_part_;
  int rows_std_[2] = {1, 1}, rows_dp_[2] = {0, 0}, mi_dp_[2] = {0, 0};
  // work lists per spin, block set: ranges are [block][kind]
  std::vector<int> blk_pair_[2];                       // per set: first pair of each block (+ end)
  std::vector<int> rng_syn_[2][2], rng_adj_[2][2];     // [spin][set]: first item of every (block, kind) (+ end)
  DevBuf<WorkItem> d_items_syn_[2][2], d_items_adj_[2][2];
  DevBuf<int64_t> d_partoff_[2][2];
};

// Generic N/S pair detection on an arbitrary theta list (pole -> equator order).
std::vector<PairHost> build_pairs(int64_t lmax, int64_t nring, const double* theta);
// first pair of each ring block (+ the end): the pole -> equator pair list cut into <= want ranges (host only)
std::vector<int> cut_ring_blocks(const std::vector<PairHost>& pairs, int want, const double* slot_weight,
                                 bool equal = false);

}  // namespace hpsht
