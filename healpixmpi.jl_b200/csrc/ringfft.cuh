// ringfft.cuh — host interface of the per-ring FFT stage (ringfft.cu).
//
// Replaces Ducc0.Sht.leg2map! (reference src/sht.jl:93) and Ducc0.Sht.map2leg! (src/sht.jl:169):
// per ring, phase shift exp(+-i m phi0), alias fold of m = 0..mmax onto the ring length, and a real
// FFT of length nphi (unnormalised).  leg layout is the reference's map-side one:
// leg[(comp*nring + ir)*nm + m]; map[comp*map_cstride + ringstart[ir] + j].
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <vector>

#include "common.hpp"

namespace hpsht {

struct RingDesc {
  int64_t pixoff;   // 0-based offset of the ring's first pixel in one map component
  int64_t hoff;     // offset (complex elements) of the ring's chirp spectrum, -1 for power-of-two r
  double phi0;
  int pq;           // phi0 = pq * pi / nphi (phases from the root table), or -1: general phi0
  int ir;           // ring slot in the leg array
  int nphi;         // ring length n = 4 r
  int r;
  int M;            // Bluestein convolution length (power of two >= 2r-1), 0 for power-of-two r
  int64_t soff;     // long rings (r > 4096): offset of the ring's u|v block in the scratch, else -1
  int64_t yoff;     // long Bluestein rings: offset of the ring's y_lo block in the second scratch, else -1
};

class RingFFT {
 public:
  RingFFT() = default;
  ~RingFFT() = default;
  RingFFT(const RingFFT&) = delete;
  RingFFT& operator=(const RingFFT&) = delete;

  // slot0 / leg_nring: the nring rings are slots slot0 .. slot0+nring-1 of a leg array that holds
  // leg_nring rings per component (defaults: the whole array)
  void setup(int64_t nring, const int64_t* nphi, const double* phi0, const int64_t* ringstart,
             int64_t nm, cudaStream_t stream, int64_t slot0 = 0, int64_t leg_nring = -1);
  // device pointers
  void leg2map(const double* d_leg, double* d_map, int64_t map_cstride, int ncomp,
               cudaStream_t stream);
  void map2leg(const double* d_map, int64_t map_cstride, double* d_leg, int ncomp,
               cudaStream_t stream);

  int64_t launches() const { return launches_; }
  void reset_launches() { launches_ = 0; }
  // algorithmic HBM bytes of one component, one direction: 16*nm*nring + 8*npix
  double algorithmic_bytes() const { return 16.0 * (double)nm_ * (double)nrings_own_ + 8.0 * (double)npix_; }

 private:
  struct Class {
    int first = 0, count = 0;  // range in the sorted descriptor array
    int threads = 0;
    size_t smem = 0;
    int rmax = 0, Mmax = 0;
    bool half = false;         // power-of-two rings: one CTA per half ring
  };
  int64_t nring_ = 0, nrings_own_ = 0, nm_ = 0, npix_ = 0;
  int64_t launches_ = 0;
  void ensure_long_scratch(int ncomp);
  std::vector<RingDesc> desc_, desc_long_;
  DevBuf<RingDesc> d_desc_long_;
  DevBuf<double> d_scratch_, d_ytmp_;
  int64_t scratch_cstride_ = 0, ytmp_cstride_ = 0;
  std::vector<Class> classes_;
  DevBuf<RingDesc> d_desc_;
  DevBuf<double> d_hhat_;   // chirp spectra
};

}  // namespace hpsht
