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
  int64_t slotN, slotS;      // ring slots in the caller's theta list (slotS = -1: unpaired)
  int mlim0, mlim2;          // m > mlim  => culled
};

// Leg addressing.  Ring slot `slot` belongs to a block starting at a0[slot] that holds nr[slot] rings,
// the slot being ring j[slot] of that block; local m's are grouped in contiguous slabs
// (sstart[mi] = first local m index of mi's slab, nms[mi] = number of m in it).  Element
// (slot, mi, comp) lives at
//     a0 + nr * (ncs * sstart + comp * nms) + j * nms + (mi - sstart)          [complex elements]
// i.e. block layout [m-slab][component][ring][m within slab].  With one slab (sstart = 0, nms = nm) and
// a0 = 0, nr = nring, j = slot this is the reference's (nm, nring, ncomp) column-major array.
struct LegLayout {
  std::vector<int64_t> a0, nr, j;      // per ring slot
  std::vector<int64_t> sstart, nms;    // per local m (empty: single slab)
  int64_t ncs = 1;                     // components the slab blocks are sized for
};

struct WorkItem {
  int mi;        // local m index
  int p0;        // first pair of the tile
  int np;        // number of pairs in the tile (<= tile size)
  int part;      // slot in the partial-sum buffer (adjoint only)
};

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

  // alm (device, complex128 [ncomp][alm_cstride]) -> leg (device, complex128, per `layout`)
  // leg_elems: total complex elements of the leg buffer (it is zero-filled first)
  void alm2leg(const double* d_alm, int64_t alm_cstride, double* d_leg, int64_t leg_elems,
               int ncomp, cudaStream_t stream);
  // leg -> alm
  void leg2alm(const double* d_leg, double* d_alm, int64_t alm_cstride, int ncomp,
               cudaStream_t stream);
  // the same, slab by slab (lets the caller overlap the leg exchange of one m-slab with the
  // Legendre work of the next): begin, then every slab 0..nslabs()-1 in order, then (adjoint) end
  int nslabs() const { return (int)slab_first_mi_.size() - 1; }
  void alm2leg_begin(const double* d_alm, int64_t alm_cstride, double* d_leg, int64_t leg_elems,
                     int ncomp, cudaStream_t stream);
  // alm2leg_begin = alm2leg_zero (clear the leg array) + alm2leg_prep over all local m's; callers that
  // stream the alm in from the host prepare m-ranges as they arrive
  void alm2leg_zero(double* d_leg, int64_t leg_elems, int ncomp, cudaStream_t stream);
  void alm2leg_prep(const double* d_alm, int64_t alm_cstride, int ncomp, int mi0, int mi1, cudaStream_t stream);
  void alm2leg_slab(int slab, double* d_leg, int ncomp, cudaStream_t stream);
  void leg2alm_begin(int ncomp, cudaStream_t stream);
  void leg2alm_slab(int slab, const double* d_leg, int ncomp, cudaStream_t stream);
  void leg2alm_end(double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream);
  // re-records the end-of-main-kernel timing event (when slab kernels ran on other streams)
  void mark_main_end(cudaStream_t stream);

  // Ring blocks: the pole -> equator pair list cut into <= `want` contiguous ranges of about equal
  // weight (slot_weight[slot], e.g. pixels per ring).  For the reference's round-robin ring order a
  // block's rings are a contiguous local range on every rank, so the caller can pipeline  Legendre(block) -> ring FFT(block) -> device-to-host copy(block)  and
  // the reverse for the adjoint.  Call right after setup(); returns the number of blocks made.
  int setup_blocks(int want, const double* slot_weight);
  int nblocks() const { return (int)blk_pair_.size() - 1; }
  void block_slots(int b, int64_t* slot_lo, int64_t* slot_hi) const;   // hull [slot_lo, slot_hi)
  std::vector<int64_t> block_slot_list(int b) const;                   // every ring slot of the block
  // synthesis: alm2leg_begin, then every block once in any order
  void alm2leg_block(int b, double* d_leg, int ncomp, cudaStream_t stream, bool last);
  // adjoint: leg2alm_blocks_begin, every block once (same stream), leg2alm_blocks_end
  void leg2alm_blocks_begin(int ncomp, cudaStream_t stream);
  void leg2alm_block(int b, const double* d_leg, int ncomp, cudaStream_t stream, bool last);
  // reduce + post-process the local m's [mi0, mi1) (default: all)
  void leg2alm_blocks_end(double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream, int mi0 = 0,
                          int mi1 = -1);
  int block_mi_end(int b, int spin) const;

  double steps_culled(int spin) const { return spin ? culled2_ : culled0_; }
  double steps_dense() const { return dense_; }
  int64_t launches() const { return launches_; }
  void reset_launches() { launches_ = 0; }
  size_t device_bytes() const { return dev_bytes_; }
  // device time of the most recent main (recurrence) kernel alone; call after a stream sync
  double main_kernel_ms() const;

 private:
  void ensure_spin(int spin, cudaStream_t stream);
  void fill_params(struct LegParams& P, int spin, bool adjoint) const;
  std::vector<int> slab_first_mi_;
  std::vector<int> srange_s0_, srange_s2_, srange_a0_, srange_a2_;  // first item of each slab

  int64_t lmax_ = 0, nm_ = 0, nring_ = 0, npairs_ = 0;
  int max_ncomp_ = 1;
  std::vector<int64_t> mval_, mstart_;
  std::vector<PairHost> pairs_;
  LegLayout layout_;
  LegendreTables tab_;
  bool have0_ = false, have2_ = false;
  bool direct0_ = false, direct2_ = false;   // adjoint partial rows already are the per-m sums
  double culled0_ = 0, culled2_ = 0, dense_ = 0;
  int64_t launches_ = 0;
  size_t dev_bytes_ = 0;

  // device: geometry
  DevBuf<double> d_cth_, d_sth_, d_sh_, d_ch_;
  DevBuf<int> d_mlim0_, d_mlim2_;
  DevBuf<int64_t> d_a0_, d_nrr_, d_jN_, d_jS_, d_sstart_, d_nms_;
  DevBuf<int64_t> d_mval_, d_mstart_;
  // device: tables
  DevBuf<int64_t> d_off0_, d_off2_, d_nstep0_, d_nstep2_;
  DevBuf<double> d_coef0_, d_invs0_, d_k0_, d_coef2_, d_sig2_, d_kp2_, d_km2_;
  DevBuf<double> d_stream0_, d_stream2_;   // prepared alm streams (4 doubles per step)
  std::vector<int64_t> nstep0_h_, off0_h_, nstep2_h_, off2_h_;
  size_t part_steps_need_ = 0, accum_steps_need_ = 0;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  DevBuf<double> d_accum_;    // reduced adjoint sums (4 doubles per step)
  DevBuf<double> d_part_;     // adjoint per-tile partial sums
  // work lists (synthesis tiles, adjoint tiles)
  std::vector<WorkItem> items_s0_, items_s2_, items_a0_, items_a2_;
  DevBuf<WorkItem> d_items_s0_, d_items_s2_, d_items_a0_, d_items_a2_;
  std::vector<int> afirst0_, acount0_, afirst2_, acount2_;   // per mi: adjoint items
  DevBuf<int> d_afirst0_, d_acount0_, d_afirst2_, d_acount2_;
  DevBuf<int64_t> d_apartoff0_, d_apartoff2_;                // per adjoint item: offset into d_part_ (steps)
  std::vector<int64_t> apartoff0_, apartoff2_;
  // ring-block work lists
  std::vector<int> blk_pair_;                                    // first pair of each block (+ end)
  std::vector<int> brange_s0_, brange_s2_, brange_a0_, brange_a2_;
  DevBuf<WorkItem> d_bitems_s0_, d_bitems_s2_, d_bitems_a0_, d_bitems_a2_;
  DevBuf<int64_t> d_bpartoff0_, d_bpartoff2_, d_brpartoff0_, d_brpartoff2_;
  DevBuf<int> d_brfirst_, d_brcount_;
  int bgroups_ = 1;
  size_t bpart_steps_need_ = 0;
  void build_block_items(int spin, const LegendreTables& T, cudaStream_t stream);
};

// Generic N/S pair detection on an arbitrary theta list (pole -> equator order).
std::vector<PairHost> build_pairs(int64_t lmax, int64_t nring, const double* theta);

}  // namespace hpsht
