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

// Where ring slot `slot` of local-m row `mi`, component `comp` lives in the leg buffer (in complex
// elements):  base[slot] + mi*mstride[slot] + comp*cstride[slot].
struct LegLayout {
  std::vector<int64_t> base, mstride, cstride;  // per ring slot
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

  double steps_culled(int spin) const { return spin ? culled2_ : culled0_; }
  double steps_dense() const { return dense_; }
  int64_t launches() const { return launches_; }
  void reset_launches() { launches_ = 0; }
  size_t device_bytes() const { return dev_bytes_; }
  // device time of the most recent main (recurrence) kernel alone; call after a stream sync
  double main_kernel_ms() const;

 private:
  void ensure_spin(int spin, cudaStream_t stream);

  int64_t lmax_ = 0, nm_ = 0, nring_ = 0, npairs_ = 0;
  int max_ncomp_ = 1;
  std::vector<int64_t> mval_, mstart_;
  std::vector<PairHost> pairs_;
  LegLayout layout_;
  LegendreTables tab_;
  bool have0_ = false, have2_ = false;
  double culled0_ = 0, culled2_ = 0, dense_ = 0;
  int64_t launches_ = 0;
  size_t dev_bytes_ = 0;

  // device: geometry
  DevBuf<double> d_cth_, d_sth_, d_sh_, d_ch_;
  DevBuf<int> d_mlim0_, d_mlim2_;
  DevBuf<int64_t> d_offN_, d_offS_, d_mstr_, d_cstr_;
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
};

// Generic N/S pair detection on an arbitrary theta list (pole -> equator order).
std::vector<PairHost> build_pairs(int64_t lmax, int64_t nring, const double* theta);

}  // namespace hpsht
