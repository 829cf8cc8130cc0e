// ringfft.cu — batched variable-length real FFTs of the ring stage (leg2map / map2leg), sm_100a.
//
// Replaces Ducc0.Sht.leg2map! (reference src/sht.jl:93) and Ducc0.Sht.map2leg! (src/sht.jl:169).
//
// One CTA transforms one ring of one component entirely in shared memory:
//   leg2map:  load leg row (coalesced) -> phase shift exp(i m phi0) + alias fold onto the half-complex
//             spectrum X[0..n/2] -> real inverse FFT of length n = 4r -> coalesced double2 pixel stores.
//   map2leg:  the exact transpose.
// The length-n real transform is reduced to two complex transforms of length r = n/4:
//   Zc[k] = (X[k] + conj X[N-k]) + i w^k (X[k] - conj X[N-k]),  N = 2r, w = exp(2 pi i / n)
//   u[k] = Zc[k] + Zc[k+r],  v[k] = (Zc[k] - Zc[k+r]) w^{2k},   U = IDFT_r(u), V = IDFT_r(v)
//   x[4j..4j+3] = Re U[j], Im U[j], Re V[j], Im V[j]
// IDFT_r is a radix-4 in-place power-of-two FFT when r is a power of two (the whole equatorial belt)
// and a Bluestein chirp-z convolution of power-of-two length M >= 2r-1 otherwise (polar caps, where
// r runs over every integer and is frequently prime).  The forward DIF leaves its output in
// bit-reversed order, the chirp spectrum is stored in that order and the inverse DIT consumes it, so
// no reordering pass is needed inside the convolution.
// Rings are grouped by size class (threads / shared memory per CTA) at plan time.
#include "ringfft.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#ifdef HPSHT_FFT_PROFILE
// phase-level cycle accounting of k_leg2map (build with -DHPSHT_FFT_PROFILE; tools/fft_profile.py)
__device__ unsigned long long g_fft_prof[64];
#define FFT_PROF_INIT() long long prof_t = clock64(); const int prof_c = (d.M == 0 ? 0 : (d.M > 4096 ? 1 : 2)) * 16
#define FFT_PROF_INIT_C(c) long long prof_t = clock64(); const int prof_c = (c) * 16
#define FFT_PROF(i)                                                                   \
  do {                                                                                \
    __syncthreads();                                                                  \
    if (threadIdx.x == 0) {                                                           \
      const long long t2 = clock64();                                                 \
      atomicAdd(&g_fft_prof[prof_c + (i)], (unsigned long long)(t2 - prof_t));        \
      prof_t = t2;                                                                    \
    }                                                                                 \
  } while (0)
extern "C" int hpsht_debug_fft_profile(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_fft_prof, sizeof(unsigned long long) * 64);
  if (reset) {
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(g_fft_prof, z, sizeof(z));
  }
  return 0;
}
#else
#define FFT_PROF_INIT()
#define FFT_PROF_INIT_C(c)
#define FFT_PROF(i)
#endif

namespace hpsht {

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by +i (INV) or -i (forward)
template <bool INV>
__device__ __forceinline__ double2 mul_pm_i(double2 a) {
  return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// T-th roots of unity exp(-2 pi i k / T) (T a power of two), two-level shared-memory table:
// value(k) = hi[k >> 6] * lo[k & 63]; the inverse transform conjugates.
struct Tw {
  const double2* hi;
  const double2* lo;
  int T;
  template <bool INV>
  __device__ __forceinline__ double2 get(int k) const {
    const double2 a = hi[k >> 6], b = lo[k & 63];
    const double2 t = make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
    return INV ? make_double2(t.x, -t.y) : t;
  }
};

// exp(i m phi0) with the rounding error of the product m*phi0 compensated
__device__ __forceinline__ double2 phase(int m, double phi0) {
  const double dm = (double)m;
  const double p = dm * phi0;
  const double e = fma(dm, phi0, -p);
  double s, c;
  sincos(p, &s, &c);
  return make_double2(fma(-e, s, c), fma(e, c, s));
}

// n-th roots of unity through a two-level table: W(q) = exp(2 pi i q / n) = hi[q>>6] * lo[q&63]
struct Roots {
  const double2* hi;
  const double2* lo;
  __device__ __forceinline__ double2 operator()(int q) const { return cmul(hi[q >> 6], lo[q & 63]); }
};
// sign = +1: exp(+2 pi i q / n) ; sign = -1: exp(-2 pi i q / n)
__device__ __forceinline__ void build_roots(double2* hi, double2* lo, int n, double sign, int tid, int nt) {
  const int nhi = (n >> 6) + 1;
  const double inv = 1.0 / (double)n;
  for (int i = tid; i < nhi + 64; i += nt) {
    const bool is_lo = i >= nhi;
    const int q = is_lo ? (i - nhi) : (i << 6);
    double s, c;
    sincospi(2.0 * (double)q * inv, &s, &c);
    if (is_lo)
      lo[i - nhi] = make_double2(c, sign * s);
    else
      hi[i] = make_double2(c, sign * s);
  }
}

// exp(i m phi0) for all m of one ring.  HEALPix rings have phi0 = pq * pi / n with pq = 0 or 1, i.e. the
// phases are 2n-th roots of unity: omega_2n^t = W(t >> 1) * (t odd ? exp(i pi / n) : 1) from the n-th root
// table the CTA holds anyway -- ~10 flops instead of a double-precision sincos per element.  Rings whose
// phi0 is not such a multiple (pq < 0; custom geometries) take the general path.
struct Phase {
  double phi0;
  int pq;
  unsigned n2;
  double2 c1;
  Roots W;
  __device__ __forceinline__ Phase(const RingDesc& d, const Roots& w) : phi0(d.phi0), pq(d.pq), n2(2u * (unsigned)d.nphi), W(w) {
    double s, c;
    sincospi(1.0 / (double)d.nphi, &s, &c);
    c1 = make_double2(c, s);
  }
  __device__ __forceinline__ double2 operator()(int m) const {
    if (pq < 0) return phase(m, phi0);
    if (pq == 0) return make_double2(1.0, 0.0);
    // HEALPix rings have pq = 1 and, for m < 2n, no reduction at all (the 64-bit modulo costs more than the rest)
    const unsigned t = pq == 1 ? ((unsigned)m < n2 ? (unsigned)m : (unsigned)m % n2)
                               : (unsigned)(((unsigned long long)(unsigned)m * (unsigned)pq) % n2);
    const double2 w = W((int)(t >> 1));
    return (t & 1u) ? cmul(w, c1) : w;
  }
};

// ------------------------------------------------------------------------------------------------
// in-place power-of-two FFT passes over shared memory (all threads of the CTA participate)
//   DIF: natural order in, bit-reversed out.  DIT: bit-reversed in, natural out.
//   INV = false: exp(-2 pi i jk/M);  INV = true: exp(+2 pi i jk/M).  Unnormalised.
//   tw: two-level shared-memory table of the T-th roots exp(-2 pi i k / T), T >= M.
// ------------------------------------------------------------------------------------------------
// Physical position of logical element i in a swizzled work array: the low three index bits are XORed with
// bits 2..4, a permutation inside every aligned group of 8 elements.  With 16-byte elements a quarter-warp
// then hits eight distinct 4-bank groups in EVERY radix-4 pass (quarter lengths 1, 4, 16, ...), where the plain
// layout is 4-way (q = 1) and 2-way (q = 4) conflicted -- tools/proto/fft_bank_model.py reproduces the 36 %
// replay excess ncu reports for M = 8192 and shows 0 % with this map.
// SWZ = 2 (half-ring kernels of the power-of-two rings, length 2^lg >= 256) also folds the TOP three index bits in:
// the bit-reversed reads / writes that replace a reorder pass touch eight elements a multiple of 2^(lg-3) apart,
// which the plain layout and SWZ = 1 put into one 4-bank group (8-way conflict); with the top bits in the XOR they
// fall into eight distinct groups, and every pattern SWZ = 1 serves stays conflict free (the top bits are constant
// within a quarter-warp there).  ncu before: 48 % of the shared-memory wavefronts of these kernels were replays.
template <int SWZ>
__device__ __forceinline__ int swz(int i, int lg = 0) {
  return SWZ == 1 ? (i ^ ((i >> 2) & 7))
                  : (SWZ == 2 ? (i ^ (((i >> 2) ^ (i >> (lg - 3))) & 7)) : (SWZ == 3 ? (i ^ ((i >> 4) & 7)) : i));
}

// `nb` transforms of length M stored back to back (buf + b * M) advance through the passes together
template <bool INV, int SWZ = 0>
__device__ void fft_dif(double2* buf, int M, const Tw& tw, int tid, int nt, int nb = 1) {
  if (M < 2) return;
  const int lg = 31 - __clz(M);
  int L = M;
  if (lg & 1) {
    const int h = L >> 1;
    const int ts = tw.T / L;
    for (int t = tid; t < h * nb; t += nt) {
      const int i = t & (h - 1);
      double2* p = buf + (t >> (lg - 1)) * M;
      const double2 a = p[swz<SWZ>(i, lg)], b = p[swz<SWZ>(i + h, lg)];
      p[swz<SWZ>(i, lg)] = cadd(a, b);
      p[swz<SWZ>(i + h, lg)] = cmul(csub(a, b), tw.get<INV>(i * ts));
    }
    L >>= 1;
    __syncthreads();
  }
  while (L >= 4) {
    const int q = L >> 2;
    const int lq = 31 - __clz(q);
    const int ts = tw.T / L;
    for (int t = tid; t < (M >> 2) * nb; t += nt) {
      const int i = t & ((M >> 2) - 1);
      double2* p = buf + (t >> (lg - 2)) * M;
      const int j = i & (q - 1);
      const int b = ((i >> lq) * L) + j;
      const int i0 = swz<SWZ>(b, lg), i1 = swz<SWZ>(b + q, lg), i2 = swz<SWZ>(b + 2 * q, lg), i3 = swz<SWZ>(b + 3 * q, lg);
      const double2 x0 = p[i0], x1 = p[i1], x2 = p[i2], x3 = p[i3];
      const double2 w1 = tw.get<INV>(j * ts);
      const double2 w2 = make_double2(fma(w1.x, w1.x, -w1.y * w1.y), 2.0 * w1.x * w1.y);  // w1^2
      const double2 a0 = cadd(x0, x2), a2 = cmul(csub(x0, x2), w1);
      const double2 a1 = cadd(x1, x3), a3 = mul_pm_i<INV>(cmul(csub(x1, x3), w1));
      p[i0] = cadd(a0, a1);
      p[i1] = cmul(csub(a0, a1), w2);
      p[i2] = cadd(a2, a3);
      p[i3] = cmul(csub(a2, a3), w2);
    }
    L >>= 2;
    __syncthreads();
  }
}

template <bool INV, int SWZ = 0>
__device__ void fft_dit(double2* buf, int M, const Tw& tw, int tid, int nt, int nb = 1) {
  if (M < 2) return;
  const int lg = 31 - __clz(M);
  int L = 1;
  if (lg & 1) {
    for (int t = tid; t < (M >> 1) * nb; t += nt) {
      const int i = t & ((M >> 1) - 1);
      double2* p = buf + (t >> (lg - 1)) * M;
      const int ia = swz<SWZ>(2 * i, lg), ib = swz<SWZ>(2 * i + 1, lg);
      const double2 a = p[ia], b = p[ib];
      p[ia] = cadd(a, b);
      p[ib] = csub(a, b);
    }
    L = 2;
    __syncthreads();
  }
  while (L < M) {
    L <<= 2;
    const int q = L >> 2;
    const int lq = 31 - __clz(q);
    const int ts = tw.T / L;
    for (int t = tid; t < (M >> 2) * nb; t += nt) {
      const int i = t & ((M >> 2) - 1);
      double2* p = buf + (t >> (lg - 2)) * M;
      const int j = i & (q - 1);
      const int b = ((i >> lq) * L) + j;
      const int i0 = swz<SWZ>(b, lg), i1 = swz<SWZ>(b + q, lg), i2 = swz<SWZ>(b + 2 * q, lg), i3 = swz<SWZ>(b + 3 * q, lg);
      const double2 x0 = p[i0], x1 = p[i1], x2 = p[i2], x3 = p[i3];
      const double2 w1 = tw.get<INV>(j * ts);
      const double2 w2 = make_double2(fma(w1.x, w1.x, -w1.y * w1.y), 2.0 * w1.x * w1.y);  // w1^2
      const double2 t1 = cmul(x1, w2), t3 = cmul(x3, w2);
      const double2 a0 = cadd(x0, t1), a1 = csub(x0, t1);
      const double2 a2 = cadd(x2, t3), a3 = csub(x2, t3);
      const double2 s2 = cmul(a2, w1), s3 = mul_pm_i<INV>(cmul(a3, w1));
      p[i0] = cadd(a0, s2);
      p[i2] = csub(a0, s2);
      p[i1] = cadd(a1, s3);
      p[i3] = csub(a1, s3);
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------
// radix-16 passes for the Bluestein work array (M >= 256): four fused radix-2 stages per trip through shared
// memory instead of two -- a 8192-point transform is 1 radix-2 + 3 radix-16 passes instead of 1 + 6 radix-4 ones.
// Any grouping of in-place radix-2 stages gives the same (bit-reversed) DIF output order, so these transforms
// are interchangeable with fft_dif / fft_dit and the chirp spectrum table keeps its order.  Layout: swz<3>
// (i ^ ((i >> 4) & 7)): the stride-16 accesses of the smallest pass and every contiguous access are conflict free.
// ------------------------------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ void bf4_dif(double2& x0, double2& x1, double2& x2, double2& x3, double2 w1, double2 w2) {
  const double2 a0 = cadd(x0, x2), a2 = cmul(csub(x0, x2), w1);
  const double2 a1 = cadd(x1, x3), a3 = mul_pm_i<INV>(cmul(csub(x1, x3), w1));
  x0 = cadd(a0, a1);
  x1 = cmul(csub(a0, a1), w2);
  x2 = cadd(a2, a3);
  x3 = cmul(csub(a2, a3), w2);
}
template <bool INV>
__device__ __forceinline__ void bf4_dit(double2& x0, double2& x1, double2& x2, double2& x3, double2 w1, double2 w2) {
  const double2 t1 = cmul(x1, w2), t3 = cmul(x3, w2);
  const double2 a0 = cadd(x0, t1), a1 = csub(x0, t1);
  const double2 a2 = cadd(x2, t3), a3 = csub(x2, t3);
  const double2 s2 = cmul(a2, w1), s3 = mul_pm_i<INV>(cmul(a3, w1));
  x0 = cadd(a0, s2);
  x2 = csub(a0, s2);
  x1 = cadd(a1, s3);
  x3 = csub(a1, s3);
}
__device__ __forceinline__ double2 csqr(double2 w) { return make_double2(fma(w.x, w.x, -w.y * w.y), 2.0 * w.x * w.y); }
// exp(-2 pi i c / 16) (INV: conjugated), c = 1, 2, 3
template <bool INV>
__device__ __forceinline__ double2 w16(int c) {
  const double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
  const double2 v = c == 1 ? make_double2(C1, -S1) : (c == 2 ? make_double2(H, -H) : make_double2(S1, -C1));
  return INV ? make_double2(v.x, -v.y) : v;
}

template <bool INV>
__device__ void dif_pass16(double2* buf, int M, int L, const Tw& tw, int tid, int nt) {
  const int q = L >> 2, q1 = L >> 4, lq1 = 31 - __clz(q1), ts = tw.T / L;
  for (int t = tid; t < (M >> 4); t += nt) {
    const int j = t & (q1 - 1), base = ((t >> lq1) * L) + j;
    double2 e[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) e[a][c] = buf[swz<3>(base + a * q + c * q1)];
    const double2 wj = tw.get<INV>(j * ts);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double2 w1 = c == 0 ? wj : cmul(wj, w16<INV>(c));
      bf4_dif<INV>(e[0][c], e[1][c], e[2][c], e[3][c], w1, csqr(w1));
    }
    const double2 wj4 = csqr(csqr(wj)), wj8 = csqr(wj4);
#pragma unroll
    for (int a = 0; a < 4; ++a) bf4_dif<INV>(e[a][0], e[a][1], e[a][2], e[a][3], wj4, wj8);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) buf[swz<3>(base + a * q + c * q1)] = e[a][c];
  }
  __syncthreads();
}
template <bool INV>
__device__ void dit_pass16(double2* buf, int M, int L, const Tw& tw, int tid, int nt) {
  const int q = L >> 2, q1 = L >> 4, lq1 = 31 - __clz(q1), ts = tw.T / L;
  for (int t = tid; t < (M >> 4); t += nt) {
    const int j = t & (q1 - 1), base = ((t >> lq1) * L) + j;
    double2 e[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) e[a][c] = buf[swz<3>(base + a * q + c * q1)];
    const double2 wj = tw.get<INV>(j * ts);
    const double2 wj4 = csqr(csqr(wj)), wj8 = csqr(wj4);
#pragma unroll
    for (int a = 0; a < 4; ++a) bf4_dit<INV>(e[a][0], e[a][1], e[a][2], e[a][3], wj4, wj8);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double2 w1 = c == 0 ? wj : cmul(wj, w16<INV>(c));
      bf4_dit<INV>(e[0][c], e[1][c], e[2][c], e[3][c], w1, csqr(w1));
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) buf[swz<3>(base + a * q + c * q1)] = e[a][c];
  }
  __syncthreads();
}
// single radix-2 / radix-4 passes at block length L (the stages left over when log2 M is not a multiple of 4;
// they run at the LARGE block lengths, where every access is contiguous)
template <bool INV>
__device__ void dif_pass2(double2* buf, int M, int L, const Tw& tw, int tid, int nt) {
  const int h = L >> 1, lh = 31 - __clz(h), ts = tw.T / L;
  for (int t = tid; t < (M >> 1); t += nt) {
    const int i = t & (h - 1), b = ((t >> lh) * L) + i;
    const int ia = swz<3>(b), ib = swz<3>(b + h);
    const double2 a = buf[ia], c = buf[ib];
    buf[ia] = cadd(a, c);
    buf[ib] = cmul(csub(a, c), tw.get<INV>(i * ts));
  }
  __syncthreads();
}
template <bool INV>
__device__ void dit_pass2(double2* buf, int M, int L, const Tw& tw, int tid, int nt) {
  const int h = L >> 1, lh = 31 - __clz(h), ts = tw.T / L;
  for (int t = tid; t < (M >> 1); t += nt) {
    const int i = t & (h - 1), b = ((t >> lh) * L) + i;
    const int ia = swz<3>(b), ib = swz<3>(b + h);
    const double2 a = buf[ia], c = cmul(buf[ib], tw.get<INV>(i * ts));
    buf[ia] = cadd(a, c);
    buf[ib] = csub(a, c);
  }
  __syncthreads();
}
template <bool INV>
__device__ void dif_pass4(double2* buf, int M, int L, const Tw& tw, int tid, int nt) {
  const int q = L >> 2, lq = 31 - __clz(q), ts = tw.T / L;
  for (int t = tid; t < (M >> 2); t += nt) {
    const int j = t & (q - 1), b = ((t >> lq) * L) + j;
    const int i0 = swz<3>(b), i1 = swz<3>(b + q), i2 = swz<3>(b + 2 * q), i3 = swz<3>(b + 3 * q);
    double2 x0 = buf[i0], x1 = buf[i1], x2 = buf[i2], x3 = buf[i3];
    const double2 w1 = tw.get<INV>(j * ts);
    bf4_dif<INV>(x0, x1, x2, x3, w1, csqr(w1));
    buf[i0] = x0;
    buf[i1] = x1;
    buf[i2] = x2;
    buf[i3] = x3;
  }
  __syncthreads();
}
template <bool INV>
__device__ void dit_pass4(double2* buf, int M, int L, const Tw& tw, int tid, int nt) {
  const int q = L >> 2, lq = 31 - __clz(q), ts = tw.T / L;
  for (int t = tid; t < (M >> 2); t += nt) {
    const int j = t & (q - 1), b = ((t >> lq) * L) + j;
    const int i0 = swz<3>(b), i1 = swz<3>(b + q), i2 = swz<3>(b + 2 * q), i3 = swz<3>(b + 3 * q);
    double2 x0 = buf[i0], x1 = buf[i1], x2 = buf[i2], x3 = buf[i3];
    const double2 w1 = tw.get<INV>(j * ts);
    bf4_dit<INV>(x0, x1, x2, x3, w1, csqr(w1));
    buf[i0] = x0;
    buf[i1] = x1;
    buf[i2] = x2;
    buf[i3] = x3;
  }
  __syncthreads();
}
// M = 2^lg >= 16, array in the swz<3> layout.  DIF: natural in, bit-reversed out.  DIT: the reverse.
// skip_lead2: the caller has already done the leading radix-2 stage (only meaningful when log2 M is odd)
template <bool INV>
__device__ void fft_dif16(double2* buf, int M, const Tw& tw, int tid, int nt, bool skip_lead2 = false) {
  const int rem = (31 - __clz(M)) & 3;
  int L = M;
  if (rem & 1) {
    if (!skip_lead2) dif_pass2<INV>(buf, M, L, tw, tid, nt);
    L >>= 1;
  }
  if (rem & 2) {
    dif_pass4<INV>(buf, M, L, tw, tid, nt);
    L >>= 2;
  }
  for (; L >= 16; L >>= 4) dif_pass16<INV>(buf, M, L, tw, tid, nt);
}
template <bool INV>
__device__ void fft_dit16(double2* buf, int M, const Tw& tw, int tid, int nt) {
  const int rem = (31 - __clz(M)) & 3;
  const int Ltop = M >> ((rem & 1) + (rem & 2));   // block length the radix-16 passes end at
  for (int L = 16; L <= Ltop; L <<= 4) dit_pass16<INV>(buf, M, L, tw, tid, nt);
  int L = Ltop;
  if (rem & 2) {
    L <<= 2;
    dit_pass4<INV>(buf, M, L, tw, tid, nt);
  }
  if (rem & 1) {
    L <<= 1;
    dit_pass2<INV>(buf, M, L, tw, tid, nt);
  }
}

__device__ __forceinline__ int brev_idx(int i, int lg) { return lg ? (int)(__brev((unsigned)i) >> (32 - lg)) : 0; }

__device__ void bitrev_permute(double2* buf, int M, int tid, int nt) {
  if (M < 4) return;  // M = 1, 2: identity
  const int lg = 31 - __clz(M);
  for (int i = tid; i < M; i += nt) {
    const int rv = (int)(__brev((unsigned)i) >> (32 - lg));
    if (i < rv) {
      const double2 t = buf[i];
      buf[i] = buf[rv];
      buf[rv] = t;
    }
  }
  __syncthreads();
}

// U[j] = sum_k src[k] exp(+2 pi i jk / r), j,k < r.
//   power-of-two r (M == 0): in place in `A` (src must be A).
//   otherwise Bluestein: src may be A itself or another buffer; A[r..M) is clobbered.  The result goes to
//   store(j, U[j]) when a store functor is given (leg2map: straight to the pixels), else to A[0..r).
//   Caller synchronised before; with the default store the result is visible to all threads on return.
struct StoreToA {
  double2* A;
  __device__ __forceinline__ void operator()(int k, double2 v) const { A[k] = v; }
};
template <class Store>
__device__ void idft_r(double2* A, const double2* src, int r, int M, const double2* __restrict__ hhat,
                       const Roots& W, const Tw& tw, int tid, int nt, const Store& store) {
  const int two_r = 2 * r;
  // The work array is kept in a swizzled layout between the chirp multiply and the final one (swz<3> with the
  // radix-16 passes for M >= 256, swz<1> with the radix-4 passes below that).  Both boundary loops may run in place
  // (src == A, result in A[0..r)): either swizzle permutes inside aligned groups of 8 elements, which one warp
  // handles in one iteration, so "all lanes read, __syncwarp, all lanes write" is enough; the loops are uniform
  // (every thread runs every iteration).
  // chirp c[k] = exp(i pi k^2 / r) = W(2 (k^2 mod 2r))
  const bool big = M >= 256;
  // M >= 2r: the upper half of the zero-padded input is empty, so the leading radix-2 DIF stage (odd log2 M) is
  // lo[i] = a[i], hi[i] = a[i] w_M^i -- written here together with the chirp multiply instead of in a pass of its own
  const bool lead2 = big && ((31 - __clz(M)) & 1);
  const int h = M >> 1;
  for (int kb = 0; kb < (lead2 ? h : M); kb += nt) {
    const int k = kb + tid;
    double2 v = make_double2(0.0, 0.0);
    if (k < r) {
      const int q = (int)(((unsigned)k * (unsigned)k) % (unsigned)two_r);
      v = cmul(src[k], W(2 * q));
    }
    __syncwarp();
    if (lead2) {
      if (k < h) {
        A[swz<3>(k)] = v;
        A[swz<3>(k + h)] = k < r ? cmul(v, tw.get<false>(k * (tw.T / M))) : v;
      }
    } else if (k < M) {
      A[big ? swz<3>(k) : swz<1>(k)] = v;
    }
  }
  __syncthreads();
  if (big) {
    fft_dif16<false>(A, M, tw, tid, nt, lead2);
    // (multiplying the chirp spectrum in on the way into the first radix-16 DIT pass was measured slower: that pass
    // gives every thread 16 consecutive elements, i.e. global loads 256 B apart across a warp)
    for (int k = tid; k < M; k += nt) {
      const int p = swz<3>(k);
      A[p] = cmul(A[p], __ldg(&hhat[k]));
    }
    __syncthreads();
    fft_dit16<true>(A, M, tw, tid, nt);
  } else {
    fft_dif<false, 1>(A, M, tw, tid, nt);
    for (int k = tid; k < M; k += nt) {
      const int p = swz<1>(k);
      A[p] = cmul(A[p], __ldg(&hhat[k]));
    }
    __syncthreads();
    fft_dit<true, 1>(A, M, tw, tid, nt);
  }
  for (int kb = 0; kb < r; kb += nt) {
    const int k = kb + tid;
    double2 v = make_double2(0.0, 0.0);
    if (k < r) {
      const int q = (int)(((unsigned)k * (unsigned)k) % (unsigned)two_r);
      v = cmul(A[big ? swz<3>(k) : swz<1>(k)], W(2 * q));
    }
    __syncwarp();
    if (k < r) store(k, v);
  }
  __syncthreads();
}
__device__ void idft_r(double2* A, const double2* src, int r, int M, const double2* __restrict__ hhat,
                       const Roots& W, const Tw& tw, int tid, int nt) {
  if (M == 0) {  // r is a power of two
    fft_dif<true>(A, r, tw, tid, nt);
    bitrev_permute(A, r, tid, nt);
    return;
  }
  idft_r(A, src, r, M, hhat, W, tw, tid, nt, StoreToA{A});
}

// Shared-memory carve-up of one CTA.
//   A  : max(Mmax, 2 rmax + 2) complex — spectrum X[0..2r], then u|v, then the Bluestein work array
//   B  : rmax complex, only for Bluestein classes (parks v / Uhat while A is the work array)
struct SmemLayout {
  double2* A;
  double2* B;
  double2* hi;    // rmax*4/64 + 2   n-th roots (n = 4r)
  double2* lo;    // 64
  double2* hi2;   // Tmax/64 + 2     T-th roots of the power-of-two FFT (T = M or r)
  double2* lo2;   // 64
  double2* part;  // nthreads (fold scratch)
};
__host__ __device__ __forceinline__ int size_T(int rmax, int Mmax) { return Mmax > rmax ? Mmax : rmax; }
__host__ __device__ __forceinline__ int size_A(int rmax, int Mmax) {
  return Mmax > 2 * rmax + 2 ? Mmax : 2 * rmax + 2;
}
__device__ __forceinline__ SmemLayout carve(unsigned char* base, int rmax, int Mmax, int nt) {
  SmemLayout s;
  double2* p = reinterpret_cast<double2*>(base);
  s.A = p;
  p += size_A(rmax, Mmax);
  s.B = p;
  p += (Mmax > 0 ? rmax : 0);
  s.hi = p;
  p += (rmax * 4) / 64 + 2;
  s.lo = p;
  p += 64;
  s.hi2 = p;
  p += size_T(rmax, Mmax) / 64 + 2;
  s.lo2 = p;
  p += 64;
  s.part = p;
  return s;
}
size_t smem_bytes(int rmax, int Mmax, int nt) {
  return sizeof(double2) * ((size_t)size_A(rmax, Mmax) + (Mmax > 0 ? rmax : 0) + (rmax * 4) / 64 + 2 + 64 +
                           size_T(rmax, Mmax) / 64 + 2 + 64 + nt);
}

// ------------------------------------------------------------------------------------------------
// leg2map
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_leg2map(const RingDesc* __restrict__ desc, int first, const double2* __restrict__ leg,
                          double* __restrict__ map, int64_t map_cstride, int64_t nm, int64_t nring,
                          const double2* __restrict__ hhat_all,
                          int rmax, int Mmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + blockIdx.x];
  const int comp = blockIdx.y;
  const SmemLayout S = carve(smem_raw, rmax, Mmax, nt);
  const int n = d.nphi, r = d.r, N = 2 * r;
  const int mmax = (int)nm - 1;
  const double2* row = leg + ((int64_t)comp * nring + d.ir) * nm;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, d.M ? d.M : d.r};

  FFT_PROF_INIT();
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, d.M ? d.M : d.r, -1.0, tid, nt);
  __syncthreads();  // the phases below come from the root table
  const Phase PH(d, W);
  FFT_PROF(0);

  // ---- phase shift + alias fold: X[k], k = 0..N ----
  {
    const int K = N + 1;
    const int G = (K < nt) ? nt / K : 1;
    for (int idx = tid; idx < K * G; idx += nt) {
      const int k = idx % K, g = idx / K;
      double ar = 0.0, ai = 0.0;
      if (k == 0) {
        if (g == 0) ar = row[0].x;  // c_0 = 1, only Re(leg[0]) enters
        for (long long m = (long long)(g + 1) * n; m <= mmax; m += (long long)G * n) {
          const double2 t = cmul(row[m], PH((int)m));
          ar += 2.0 * t.x;
        }
      } else {
        for (long long m = k + (long long)g * n; m <= mmax; m += (long long)G * n) {
          const double2 t = cmul(row[m], PH((int)m));
          ar += t.x;
          ai += t.y;
        }
        for (long long m = (long long)(g + 1) * n - k; m <= mmax; m += (long long)G * n) {
          const double2 t = cmul(row[m], PH((int)m));
          ar += t.x;
          ai -= t.y;
        }
      }
      if (G == 1)
        S.A[k] = make_double2(ar, ai);
      else
        S.part[idx] = make_double2(ar, ai);
    }
    if (G > 1) {
      __syncthreads();
      for (int k = tid; k < K; k += nt) {
        double ar = 0.0, ai = 0.0;
        for (int g = 0; g < G; ++g) {
          ar += S.part[g * K + k].x;
          ai += S.part[g * K + k].y;
        }
        S.A[k] = make_double2(ar, ai);
      }
    }
  }
  __syncthreads();
  FFT_PROF(1);

  // ---- X -> Zc (in place, pairs k <-> N-k) ----
  for (int k = tid; k <= r; k += nt) {
    if (k == 0) {
      const double x0 = S.A[0].x, xn = S.A[N].x;
      S.A[0] = make_double2(x0 + xn, x0 - xn);
    } else if (k == r) {
      const double2 a = S.A[r];
      S.A[r] = make_double2(2.0 * a.x, -2.0 * a.y);
    } else {
      const double2 A = S.A[k], B = cconj(S.A[N - k]);
      const double2 Pp = cadd(A, B);
      const double2 D = cmul(csub(A, B), W(k));
      const double2 Q = make_double2(-D.y, D.x);  // i * w^k (A - B)
      S.A[k] = cadd(Pp, Q);
      S.A[N - k] = make_double2(Pp.x - Q.x, -(Pp.y - Q.y));  // conj(P) - conj(Q) = conj(P - Q)
    }
  }
  __syncthreads();
  // ---- Zc -> u | v (in place) ----
  for (int k = tid; k < r; k += nt) {
    const double2 a = S.A[k], b = S.A[k + r];
    S.A[k] = cadd(a, b);
    S.A[k + r] = cmul(csub(a, b), W(2 * k));
  }
  __syncthreads();
  FFT_PROF(2);

  const double2* hhat = d.hoff >= 0 ? hhat_all + d.hoff : nullptr;
  double* out = map + (int64_t)comp * map_cstride + d.pixoff;
  const bool al16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (d.M == 0) {
    // power-of-two r: both halves transformed in place (DIF: results in bit-reversed order, read back
    // through the reversed index instead of a separate permutation pass), one fully coalesced store
    fft_dif<true>(S.A, r, TW, tid, nt, 2);
    FFT_PROF(3);
    const int lgr = 31 - __clz(r);
    // ---- pixels: x[4j..4j+3] = Re U[j], Im U[j], Re V[j], Im V[j] ----
    if (al16) {
      double2* out2 = reinterpret_cast<double2*>(out);
      for (int t = tid; t < N; t += nt) out2[t] = S.A[((t & 1) ? r : 0) + brev_idx(t >> 1, lgr)];
    } else {
      for (int t = tid; t < N; t += nt) {
        const double2 v = S.A[((t & 1) ? r : 0) + brev_idx(t >> 1, lgr)];
        out[2 * t] = v.x;
        out[2 * t + 1] = v.y;
      }
    }
    FFT_PROF(4);
  } else {
    // Bluestein: A becomes the work array, v waits in B.  U and V are stored as interleaved 16-byte
    // pieces (the two stores of a 32-byte sector meet in L2).
    for (int k = tid; k < r; k += nt) S.B[k] = S.A[k + r];
    __syncthreads();
    for (int half = 0; half < 2; ++half) {
      // the final chirp multiply writes the pixels directly
      if (al16) {
        double2* out2 = reinterpret_cast<double2*>(out);
        idft_r(S.A, half == 0 ? S.A : S.B, r, d.M, hhat, W, TW, tid, nt,
               [=](int j, double2 v) { out2[2 * j + half] = v; });
      } else {
        idft_r(S.A, half == 0 ? S.A : S.B, r, d.M, hhat, W, TW, tid, nt, [=](int j, double2 v) {
          out[4 * j + 2 * half] = v.x;
          out[4 * j + 2 * half + 1] = v.y;
        });
      }
      FFT_PROF(3);
      __syncthreads();  // A is reused as the work array of the second half
      FFT_PROF(4);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// map2leg
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_map2leg(const RingDesc* __restrict__ desc, int first, const double* __restrict__ map,
                          int64_t map_cstride, double2* __restrict__ leg, int64_t nm, int64_t nring,
                          const double2* __restrict__ hhat_all,
                          int rmax, int Mmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + blockIdx.x];
  const int comp = blockIdx.y;
  const SmemLayout S = carve(smem_raw, rmax, Mmax, nt);
  const int n = d.nphi, r = d.r, N = 2 * r;
  const int mmax = (int)nm - 1;
  double2* row = leg + ((int64_t)comp * nring + d.ir) * nm;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, d.M ? d.M : d.r};

  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, d.M ? d.M : d.r, -1.0, tid, nt);

  // ---- pixels -> conj(u) | conj(v) ;  u[j] = x[4j] + i x[4j+1], v[j] = x[4j+2] + i x[4j+3] ----
  const double* in = map + (int64_t)comp * map_cstride + d.pixoff;
  const bool al16 = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  double2* vdst = d.M == 0 ? S.A + r : S.B;  // Bluestein: v parks in B while A is the work array
  // power-of-two r: stored in bit-reversed order so that the DIT passes below deliver natural order
  const int lgr = d.M == 0 ? 31 - __clz(r) : -1;
  if (al16) {
    const double2* in2 = reinterpret_cast<const double2*>(in);
    for (int t = tid; t < N; t += nt) {
      const double2 v = in2[t];
      const int j = lgr >= 0 ? brev_idx(t >> 1, lgr) : (t >> 1);
      ((t & 1) ? vdst : S.A)[j] = make_double2(v.x, -v.y);
    }
  } else {
    for (int t = tid; t < N; t += nt) {
      const int j = lgr >= 0 ? brev_idx(t >> 1, lgr) : (t >> 1);
      ((t & 1) ? vdst : S.A)[j] = make_double2(in[2 * t], -in[2 * t + 1]);
    }
  }
  __syncthreads();

  const double2* hhat = d.hoff >= 0 ? hhat_all + d.hoff : nullptr;
  if (d.M == 0) {
    fft_dit<true>(S.A, r, TW, tid, nt, 2);  // = conj(Uhat) | conj(Vhat)
  } else {
    idft_r(S.A, S.A, r, d.M, hhat, W, TW, tid, nt);  // conj(Uhat) in A[0..r)
    for (int k = tid; k < r; k += nt) {                   // swap: conj(Uhat) -> B, conj(v) -> A
      const double2 t = S.A[k];
      S.A[k] = S.B[k];
      S.B[k] = t;
    }
    __syncthreads();
    idft_r(S.A, S.A, r, d.M, hhat, W, TW, tid, nt);  // conj(Vhat) in A[0..r)
    for (int k = tid; k < r; k += nt) {                   // A[0..r) = conj(Uhat), A[r..2r) = conj(Vhat)
      const double2 vh = S.A[k];
      S.A[k] = S.B[k];
      S.A[k + r] = vh;
    }
    __syncthreads();
  }

  // ---- Z[k] = Uhat[k] + conj(w)^{2k} Vhat[k],  Z[k+r] = Uhat[k] - ...  (in place) ----
  for (int k = tid; k < r; k += nt) {
    const double2 uh = cconj(S.A[k]), vh = cconj(S.A[k + r]);
    const double2 t = cmulc(vh, W(2 * k));
    S.A[k] = cadd(uh, t);
    S.A[k + r] = csub(uh, t);
  }
  __syncthreads();
  // ---- Z -> X[0..N] (pairs k <-> N-k) ----
  for (int k = tid; k <= r; k += nt) {
    if (k == 0) {
      const double2 z = S.A[0];
      S.A[0] = make_double2(z.x + z.y, 0.0);
      S.A[N] = make_double2(z.x - z.y, 0.0);
    } else if (k == r) {
      S.A[r] = cconj(S.A[r]);
    } else {
      const double2 A = S.A[k], B = cconj(S.A[N - k]);
      const double2 Pp = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
      const double2 D = cmulc(csub(A, B), W(k));             // conj(w^k) (A - B)
      const double2 Q = make_double2(0.5 * D.y, -0.5 * D.x);  // -i/2 * D
      S.A[k] = cadd(Pp, Q);
      S.A[N - k] = make_double2(Pp.x - Q.x, -(Pp.y - Q.y));
    }
  }
  __syncthreads();
  // ---- un-alias + phase shift: leg[m] = exp(-i m phi0) Xt[m mod n] ----
  const Phase PH(d, W);
  for (int m = tid; m <= mmax; m += nt) {
    const int k = m % n;
    const double2 v = (k <= N) ? S.A[k] : cconj(S.A[n - k]);
    row[m] = cmulc(v, PH(m));
  }
}

// chirp spectrum of one r:  hhat = DIF_forward(h) / M,  h[d] = conj(c[|d|]) circularly, |d| < r
__global__ void k_build_hhat(const int* __restrict__ rs, const int* __restrict__ Ms,
                             const int64_t* __restrict__ hoffs, double2* __restrict__ hhat_all,
                             int rmax, int Mmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const SmemLayout S = carve(smem_raw, rmax, Mmax, nt);
  const int r = rs[blockIdx.x], M = Ms[blockIdx.x];
  const int n = 4 * r;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, M};
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, M, -1.0, tid, nt);
  __syncthreads();
  const double inv = 1.0 / (double)M;
  for (int k = tid; k < M; k += nt) {
    int dd = -1;
    if (k < r)
      dd = k;
    else if (k > M - r)
      dd = M - k;
    double2 v = make_double2(0.0, 0.0);
    if (dd >= 0) {
      const int q = (int)(((long long)dd * dd) % (2 * r));
      const double2 c = W(2 * q);
      v = make_double2(c.x * inv, -c.y * inv);
    }
    S.A[k] = v;
  }
  __syncthreads();
  fft_dif<false>(S.A, M, TW, tid, nt);
  double2* out = hhat_all + hoffs[blockIdx.x];
  for (int k = tid; k < M; k += nt) out[k] = S.A[k];
}

// ================================================================================================
// long rings: 4096 < r <= 8192 (nside up to 8192).  The half-complex spectrum of such a ring
// (2r+1 complex > 227 KB) does not fit one CTA's shared memory, so the transform is split into
//   leg2map:  k_long_pre  (fold + X->Zc->u|v evaluated on the fly, written to a global scratch)
//             k_long_fft  (one CTA per half ring: IDFT_r of u or v in shared memory -> pixels)
//   map2leg:  k_long_fft  (pixels -> conj -> IDFT_r -> conj -> scratch Uhat | Vhat)
//             k_long_post (Z, X, un-alias and phase shift evaluated on the fly -> leg)
// IDFT_r in shared memory: r = 8192 (the equatorial belt of nside 8192) is a plain radix-4 FFT; other
// r use Bluestein with M = 16384, done as one radix-2 decimation step around two 8192-point FFTs so
// that only 8192 complex values live in shared memory at a time:
//   forward DIF first stage on the zero-padded input a[0..r), r <= M/2:  lo[i] = a[i], hi[i] = a[i] w_M^i
//   spectrum halves are multiplied by the matching halves of the chirp spectrum, inverse 8192-point
//   FFTs give y_lo, y_hi and the wanted outputs are c[j] = y_lo[j] + conj(w_M^j) y_hi[j], j < r.
// ================================================================================================
constexpr int LONG_MH = 8192;

// X[k] of the alias-folded, phase-shifted spectrum, evaluated straight from the leg row (k in [0, n/2])
__device__ __forceinline__ double2 fold_eval(const double2* __restrict__ row, int k, int n, int mmax,
                                             const Phase& PH) {
  double ar = 0.0, ai = 0.0;
  if (k == 0) {
    ar = row[0].x;
    for (long long m = n; m <= mmax; m += n) {
      const double2 t = cmul(row[m], PH((int)m));
      ar += 2.0 * t.x;
    }
  } else {
    for (long long m = k; m <= mmax; m += n) {
      const double2 t = cmul(row[m], PH((int)m));
      ar += t.x;
      ai += t.y;
    }
    for (long long m = (long long)n - k; m <= mmax; m += n) {
      const double2 t = cmul(row[m], PH((int)m));
      ar += t.x;
      ai -= t.y;
    }
  }
  return make_double2(ar, ai);
}

// Zc[k], 0 <= k < N = 2r, from the folded spectrum
__device__ __forceinline__ double2 zc_eval(const double2* __restrict__ row, int k, int r, int n, int mmax,
                                           const Phase& PH, const Roots& W) {
  const int N = 2 * r;
  if (k == 0) {
    const double x0 = fold_eval(row, 0, n, mmax, PH).x, xn = fold_eval(row, N, n, mmax, PH).x;
    return make_double2(x0 + xn, x0 - xn);
  }
  if (k == r) {
    const double2 a = fold_eval(row, r, n, mmax, PH);
    return make_double2(2.0 * a.x, -2.0 * a.y);
  }
  const double2 A = fold_eval(row, k, n, mmax, PH), B = cconj(fold_eval(row, N - k, n, mmax, PH));
  const double2 Pp = cadd(A, B);
  const double2 D = cmul(csub(A, B), W(k));
  return make_double2(Pp.x - D.y, Pp.y + D.x);  // P + i w^k (A - B)
}

struct LongSmem {
  double2* A;    // LONG_MH
  double2* hi;   // n-th roots, n <= 32768: 512 + 2
  double2* lo;   // 64
  double2* hi2;  // LONG_MH-th roots: 128 + 2
  double2* lo2;  // 64
};
__device__ __forceinline__ LongSmem carve_long(unsigned char* base, bool with_A) {
  LongSmem s;
  double2* p = reinterpret_cast<double2*>(base);
  s.A = p;
  p += with_A ? LONG_MH : 0;
  s.hi = p;
  p += 514;
  s.lo = p;
  p += 64;
  s.hi2 = p;
  p += 130;
  s.lo2 = p;
  return s;
}
size_t long_smem_bytes(bool with_A) { return sizeof(double2) * ((with_A ? LONG_MH : 0) + 514 + 64 + 130 + 64); }

// leg2map, phase 1: u | v of every long ring into the scratch (2r complex per ring and component)
__global__ void __launch_bounds__(1024) k_long_pre(const RingDesc* __restrict__ desc, int first,
                                                   const double2* __restrict__ leg, double2* __restrict__ scratch,
                                                   int64_t scratch_cstride, int64_t nm, int64_t nring) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + blockIdx.x];
  const int comp = blockIdx.y;
  const LongSmem S = carve_long(smem_raw, false);
  const int n = d.nphi, r = d.r;
  const int mmax = (int)nm - 1;
  const double2* row = leg + ((int64_t)comp * nring + d.ir) * nm;
  const Roots W{S.hi, S.lo};
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  __syncthreads();
  const Phase PH(d, W);
  double2* uv = scratch + (int64_t)comp * scratch_cstride + d.soff;
  for (int k = tid; k < r; k += nt) {
    const double2 a = zc_eval(row, k, r, n, mmax, PH, W);
    const double2 b = zc_eval(row, k + r, r, n, mmax, PH, W);
    uv[k] = cadd(a, b);
    uv[k + r] = cmul(csub(a, b), W(2 * k));
  }
}

// one half ring: IDFT_r in shared memory.  C2R: scratch u|v -> pixels.  R2C: pixels -> scratch Uhat|Vhat.
template <bool C2R>
__global__ void __launch_bounds__(1024) k_long_fft(const RingDesc* __restrict__ desc, int first,
                                                   double2* __restrict__ scratch, int64_t scratch_cstride,
                                                   double2* __restrict__ ytmp, int64_t ytmp_cstride,
                                                   double* __restrict__ map, int64_t map_cstride,
                                                   const double2* __restrict__ hhat_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + (blockIdx.x >> 1)];
  const int half = blockIdx.x & 1;
  const int comp = blockIdx.y;
  const LongSmem S = carve_long(smem_raw, true);
  const int n = d.nphi, r = d.r;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, LONG_MH};
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, LONG_MH, -1.0, tid, nt);
  double2* uv = scratch + (int64_t)comp * scratch_cstride + d.soff + (int64_t)half * r;
  double* px = map + (int64_t)comp * map_cstride + d.pixoff;  // pixel pair (4j+2h, 4j+2h+1) = px2[2j+h]
  // HEALPix ring starts are multiples of 4 doubles from an even base; fall back to scalar access otherwise
  const bool al16 = (reinterpret_cast<uintptr_t>(px) & 15) == 0;
  auto src = [&](int k) -> double2 {
    if (C2R) return uv[k];
    if (al16) {
      const double2 v = reinterpret_cast<const double2*>(px)[2 * k + half];
      return make_double2(v.x, -v.y);
    }
    return make_double2(px[4 * k + 2 * half], -px[4 * k + 2 * half + 1]);
  };
  auto dst = [&](int j, double2 v) {
    if (C2R) {
      if (al16) {
        reinterpret_cast<double2*>(px)[2 * j + half] = v;
      } else {
        px[4 * j + 2 * half] = v.x;
        px[4 * j + 2 * half + 1] = v.y;
      }
    } else {
      uv[j] = make_double2(v.x, -v.y);  // conj(IDFT(conj x)) = forward DFT
    }
  };
  __syncthreads();
  if (d.M == 0) {  // r == LONG_MH: plain power-of-two transform
    for (int k = tid; k < r; k += nt) S.A[k] = src(k);
    __syncthreads();
    fft_dif<true>(S.A, r, TW, tid, nt);
    bitrev_permute(S.A, r, tid, nt);
    for (int j = tid; j < r; j += nt) dst(j, S.A[j]);
    return;
  }
  // Bluestein, M = 2 * LONG_MH
  const int two_r = 2 * r;
  const double2* hh = hhat_all + d.hoff;
  double2* y0 = ytmp + (int64_t)comp * ytmp_cstride + d.yoff + (int64_t)half * LONG_MH;
  const double invM = 1.0 / (double)(2 * LONG_MH);
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = tid; k < LONG_MH; k += nt) {
      double2 v = make_double2(0.0, 0.0);
      if (k < r) {
        const int q = (int)(((unsigned)k * (unsigned)k) % (unsigned)two_r);
        v = cmul(src(k), W(2 * q));
        if (pass == 1) {  // times w_M^k = exp(-2 pi i k / M)
          double s, c;
          sincospi(2.0 * (double)k * invM, &s, &c);
          v = cmul(v, make_double2(c, -s));
        }
      }
      S.A[k] = v;
    }
    __syncthreads();
    fft_dif<false>(S.A, LONG_MH, TW, tid, nt);
    for (int k = tid; k < LONG_MH; k += nt) S.A[k] = cmul(S.A[k], __ldg(&hh[pass * LONG_MH + k]));
    __syncthreads();
    fft_dit<true>(S.A, LONG_MH, TW, tid, nt);
    if (pass == 0) {
      for (int j = tid; j < r; j += nt) y0[j] = S.A[j];
      __syncthreads();
    }
  }
  for (int j = tid; j < r; j += nt) {
    double s, c;
    sincospi(2.0 * (double)j * invM, &s, &c);
    const double2 cj = cadd(y0[j], cmul(S.A[j], make_double2(c, s)));  // y_lo + conj(w_M^j) y_hi
    const int q = (int)(((unsigned)j * (unsigned)j) % (unsigned)two_r);
    dst(j, cmul(cj, W(2 * q)));
  }
}

// map2leg, final phase: leg[m] = exp(-i m phi0) Xt[m mod n] with X evaluated from Uhat | Vhat on the fly
__global__ void __launch_bounds__(1024) k_long_post(const RingDesc* __restrict__ desc, int first,
                                                    const double2* __restrict__ scratch, int64_t scratch_cstride,
                                                    double2* __restrict__ leg, int64_t nm, int64_t nring) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + blockIdx.x];
  const int comp = blockIdx.y;
  const LongSmem S = carve_long(smem_raw, false);
  const int n = d.nphi, r = d.r, N = 2 * r;
  const int mmax = (int)nm - 1;
  double2* row = leg + ((int64_t)comp * nring + d.ir) * nm;
  const Roots W{S.hi, S.lo};
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  __syncthreads();
  const Phase PH(d, W);
  const double2* uv = scratch + (int64_t)comp * scratch_cstride + d.soff;
  auto Z = [&](int k) -> double2 {  // k in [0, N]; Z[N] = Z[0]
    if (k >= N) k -= N;
    const int kk = k < r ? k : k - r;
    const double2 t = cmulc(uv[r + kk], W(2 * kk));  // conj(w)^{2kk} Vhat
    return k < r ? cadd(uv[kk], t) : csub(uv[kk], t);
  };
  for (int m = tid; m <= mmax; m += nt) {
    const int km = m % n;
    const bool up = km > N;
    const int k = up ? n - km : km;
    const double2 A = Z(k), B = cconj(Z(N - k));
    const double2 Pp = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
    const double2 D = cmulc(csub(A, B), W(k == N ? 0 : k));
    // conj(w)^N = -1: fold the sign into D for k == N
    const double sg = (k == N) ? -1.0 : 1.0;
    const double2 X = make_double2(Pp.x + sg * 0.5 * D.y, Pp.y - sg * 0.5 * D.x);  // P - i/2 conj(w^k)(A-B)
    const double2 v = up ? cconj(X) : X;
    row[m] = cmulc(v, PH(m));
  }
}

// chirp spectrum for the long Bluestein (M = 2 LONG_MH): the two halves of the first DIF stage
__global__ void __launch_bounds__(1024) k_build_hhat_long(const int* __restrict__ rs,
                                                          const int64_t* __restrict__ hoffs,
                                                          double2* __restrict__ hhat_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const LongSmem S = carve_long(smem_raw, true);
  const int r = rs[blockIdx.x];
  const int n = 4 * r;
  const int M = 2 * LONG_MH;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, LONG_MH};
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, LONG_MH, -1.0, tid, nt);
  __syncthreads();
  const double inv = 1.0 / (double)M;
  auto h = [&](int k) -> double2 {  // h[k], k in [0, M): conj(c[|d|]) / M circularly for |d| < r
    int dd = -1;
    if (k < r)
      dd = k;
    else if (k > M - r)
      dd = M - k;
    if (dd < 0) return make_double2(0.0, 0.0);
    const int q = (int)(((unsigned)dd * (unsigned)dd) % (unsigned)(2 * r));
    const double2 c = W(2 * q);
    return make_double2(c.x * inv, -c.y * inv);
  };
  double2* out = hhat_all + hoffs[blockIdx.x];
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = tid; k < LONG_MH; k += nt) {
      const double2 a = h(k), b = h(k + LONG_MH);
      if (pass == 0) {
        S.A[k] = cadd(a, b);
      } else {
        double s, c;
        sincospi(2.0 * (double)k * inv, &s, &c);
        S.A[k] = cmul(csub(a, b), make_double2(c, -s));
      }
    }
    __syncthreads();
    fft_dif<false>(S.A, LONG_MH, TW, tid, nt);
    for (int k = tid; k < LONG_MH; k += nt) out[pass * LONG_MH + k] = S.A[k];
    __syncthreads();
  }
}

// ================================================================================================
// power-of-two rings (the whole equatorial belt of a power-of-two nside): one CTA per HALF ring
//   leg2map:  u or v evaluated straight from the leg row in global memory (fold -> Zc -> u|v, no shared-memory
//             passes), one r-point transform, 16-byte pixel stores.  A half ring of r = 4096 needs 64 KB, so
//             three CTAs share an SM and the load / transform / store phases of different rings overlap
//             (with one CTA per ring and SM they ran one after the other: profiles/r02_ringfft_phases.md).
//   map2leg:  the two halves of a ring form a thread-block cluster; each transforms its half, then reads the
//             other half's spectrum through distributed shared memory and writes its half of the leg row.
// ================================================================================================
struct HalfSmem {
  double2* A;    // rmax
  double2* hi;   // n-th roots: rmax*4/64 + 2
  double2* lo;   // 64
  double2* hi2;  // r-th roots: rmax/64 + 2
  double2* lo2;  // 64
};
__device__ __forceinline__ HalfSmem carve_half(unsigned char* base, int rmax) {
  HalfSmem s;
  double2* p = reinterpret_cast<double2*>(base);
  s.A = p;
  p += rmax;
  s.hi = p;
  p += (rmax * 4) / 64 + 2;
  s.lo = p;
  p += 64;
  s.hi2 = p;
  p += rmax / 64 + 2;
  s.lo2 = p;
  return s;
}
size_t half_smem_bytes(int rmax) { return sizeof(double2) * ((size_t)rmax + (rmax * 4) / 64 + 2 + 64 + rmax / 64 + 2 + 64); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256)
    k_leg2map_half(const RingDesc* __restrict__ desc, int first, const double2* __restrict__ leg, double* __restrict__ map,
                   int64_t map_cstride, int64_t nm, int64_t nring, int rmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + (blockIdx.x >> 1)];
  const int half = blockIdx.x & 1;  // == rank in the cluster: 0 transforms u, 1 transforms v
  const int comp = blockIdx.y;
  const HalfSmem S = carve_half(smem_raw, rmax);
  const int n = d.nphi, r = d.r;
  const int mmax = (int)nm - 1;
  const double2* row = leg + ((int64_t)comp * nring + d.ir) * nm;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, r};
  FFT_PROF_INIT_C(3);
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, r, -1.0, tid, nt);
  cluster.sync();  // tables built; the partner's shared memory exists
  const Phase PH(d, W);
  const int lgr = 31 - __clz(r);
  FFT_PROF(0);
  // front end shared by the two CTAs: each evaluates Zc[k], Zc[k + r] for half of the k's and hands u[k] to the
  // u-CTA and v[k] to the v-CTA (one of the two stores goes through distributed shared memory)
  double2* Au = half == 0 ? S.A : cluster.map_shared_rank(S.A, 0);
  double2* Av = half == 1 ? S.A : cluster.map_shared_rank(S.A, 1);
  const int kh = (r + 1) / 2;
  const int k_lo = half == 0 ? 0 : kh, k_hi = half == 0 ? min(kh, r) : r;
  if (mmax < 2 * r) {
    // no aliasing (lmax < n / 2, e.g. the belt at lmax <= 2 nside - 1): X[j] = leg[j] exp(i j phi0) for j <= mmax, else 0.
    // Four independent, coalesced loads per k (ascending and descending runs of the row) -- the generic fold below
    // walks its alias loops one dependent load at a time.
    const int N = 2 * r;
    auto X = [&](int j) -> double2 {
      if (j > mmax) return make_double2(0.0, 0.0);
      const double2 v = __ldg(&row[j]);
      return j == 0 ? make_double2(v.x, 0.0) : cmul(v, PH(j));
    };
#pragma unroll 4
    for (int k = k_lo + tid; k < k_hi; k += nt) {
      double2 a, b;
      if (k == 0) {
        const double x0 = X(0).x, xn = X(N).x;
        const double2 xr = X(r);
        a = make_double2(x0 + xn, x0 - xn);
        b = make_double2(2.0 * xr.x, -2.0 * xr.y);
      } else {
        const double2 A1 = X(k), B1 = cconj(X(N - k)), A2 = X(k + r), B2 = cconj(X(r - k));
        const double2 D1 = cmul(csub(A1, B1), W(k)), D2 = cmul(csub(A2, B2), W(k + r));
        a = make_double2(A1.x + B1.x - D1.y, A1.y + B1.y + D1.x);  // P + i w^k (A - B)
        b = make_double2(A2.x + B2.x - D2.y, A2.y + B2.y + D2.x);
      }
      Au[swz<2>(k, lgr)] = cadd(a, b);
      Av[swz<2>(k, lgr)] = cmul(csub(a, b), W(2 * k));
    }
  } else {
    for (int k = k_lo + tid; k < k_hi; k += nt) {
      const double2 a = zc_eval(row, k, r, n, mmax, PH, W);
      const double2 b = zc_eval(row, k + r, r, n, mmax, PH, W);
      Au[swz<2>(k, lgr)] = cadd(a, b);
      Av[swz<2>(k, lgr)] = cmul(csub(a, b), W(2 * k));
    }
  }
  cluster.sync();
  FFT_PROF(1);
  fft_dif<true, 2>(S.A, r, TW, tid, nt);  // bit-reversed order: read back through the reversed index
  FFT_PROF(3);
  double* out = map + (int64_t)comp * map_cstride + d.pixoff;  // x[4j + 2 half], x[4j + 2 half + 1] = Re, Im
  if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    double2* out2 = reinterpret_cast<double2*>(out);
    for (int j = tid; j < r; j += nt) out2[2 * j + half] = S.A[swz<2>(brev_idx(j, lgr), lgr)];
  } else {
    for (int j = tid; j < r; j += nt) {
      const double2 v = S.A[swz<2>(brev_idx(j, lgr), lgr)];
      out[4 * j + 2 * half] = v.x;
      out[4 * j + 2 * half + 1] = v.y;
    }
  }
  FFT_PROF(4);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256)
    k_map2leg_half(const RingDesc* __restrict__ desc, int first, const double* __restrict__ map, int64_t map_cstride,
                   double2* __restrict__ leg, int64_t nm, int64_t nring, int rmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, nt = blockDim.x;
  const RingDesc d = desc[first + (blockIdx.x >> 1)];
  const int half = blockIdx.x & 1;  // == rank in the cluster
  const int comp = blockIdx.y;
  const HalfSmem S = carve_half(smem_raw, rmax);
  const int n = d.nphi, r = d.r, N = 2 * r;
  const int mmax = (int)nm - 1;
  double2* row = leg + ((int64_t)comp * nring + d.ir) * nm;
  const Roots W{S.hi, S.lo};
  const Tw TW{S.hi2, S.lo2, r};
  FFT_PROF_INIT_C(3);
  build_roots(S.hi, S.lo, n, 1.0, tid, nt);
  build_roots(S.hi2, S.lo2, r, -1.0, tid, nt);
  FFT_PROF(8);
  // pixels -> conj(u) (half 0) or conj(v) (half 1), stored bit-reversed so that the DIT delivers natural order
  const double* in = map + (int64_t)comp * map_cstride + d.pixoff;
  const int lgr = 31 - __clz(r);
  if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    const double2* in2 = reinterpret_cast<const double2*>(in);
    for (int j = tid; j < r; j += nt) {
      const double2 v = in2[2 * j + half];
      S.A[swz<2>(brev_idx(j, lgr), lgr)] = make_double2(v.x, -v.y);
    }
  } else {
    for (int j = tid; j < r; j += nt)
      S.A[swz<2>(brev_idx(j, lgr), lgr)] = make_double2(in[4 * j + 2 * half], -in[4 * j + 2 * half + 1]);
  }
  __syncthreads();
  FFT_PROF(9);
  fft_dit<true, 2>(S.A, r, TW, tid, nt);  // conj(Uhat) or conj(Vhat)
  cluster.sync();
  FFT_PROF(10);
  const double2* mine = S.A;
  const double2* other = cluster.map_shared_rank(S.A, half ^ 1);
  const double2* cu = half == 0 ? mine : other;   // conj(Uhat)
  const double2* cv = half == 0 ? other : mine;   // conj(Vhat)
  const Phase PH(d, W);
  auto Z = [&](int k) -> double2 {  // k in [0, N]; Z[N] = Z[0]
    if (k >= N) k -= N;
    const int kk = k < r ? k : k - r;
    const int ks = swz<2>(kk, lgr);
    const double2 t = cmulc(cconj(cv[ks]), W(2 * kk));  // conj(w)^{2kk} Vhat
    const double2 u = cconj(cu[ks]);
    return k < r ? cadd(u, t) : csub(u, t);
  };
  // this CTA's half of the leg row: leg[m] = exp(-i m phi0) Xt[m mod n]
  const int mh = (mmax + 2) / 2;
  const int m_lo = half == 0 ? 0 : mh, m_hi = half == 0 ? min(mh, mmax + 1) : mmax + 1;
  for (int m = m_lo + tid; m < m_hi; m += nt) {
    const int km = m % n;
    const bool up = km > N;
    const int k = up ? n - km : km;
    const double2 A = Z(k), B = cconj(Z(N - k));
    const double2 Pp = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
    const double2 D = cmulc(csub(A, B), W(k == N ? 0 : k));
    const double sg = (k == N) ? -1.0 : 1.0;  // conj(w)^N = -1
    const double2 X = make_double2(Pp.x + sg * 0.5 * D.y, Pp.y - sg * 0.5 * D.x);  // P - i/2 conj(w^k)(A-B)
    const double2 v = up ? cconj(X) : X;
    row[m] = cmulc(v, PH(m));
  }
  cluster.sync();  // the partner may still be reading this CTA's shared memory
  FFT_PROF(11);
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

// dynamic-smem opt-in is per kernel, not per RingFFT object: only ever raise it
// (the attribute is per device; one table row per device ordinal)
static void raise_smem_attr(const void* fn, size_t* cur_by_dev, size_t want) {
  int dev = 0;
  HP_CUDA(cudaGetDevice(&dev));
  size_t* cur = &cur_by_dev[dev & 63];
  if (want <= *cur) return;
  HP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
  *cur = want;
}
static size_t g_smem_leg2map[64] = {0}, g_smem_map2leg[64] = {0}, g_smem_hhat[64] = {0}, g_smem_leg2map_half[64] = {0},
              g_smem_map2leg_half[64] = {0};

void RingFFT::setup(int64_t nring, const int64_t* nphi, const double* phi0, const int64_t* ringstart,
                    int64_t nm, cudaStream_t stream, int64_t slot0, int64_t leg_nring) {
  nring_ = leg_nring > 0 ? leg_nring : nring;
  nrings_own_ = nring;
  nm_ = nm;
  npix_ = 0;
  desc_.clear();
  classes_.clear();
  HP_REQUIRE(nm >= 1 && nm < (1ll << 30), HPSHT_ERR_ARG, "invalid nm");
  std::map<int, int64_t> hoff_of_r;
  std::vector<int> rs, Ms;
  std::vector<int64_t> hoffs;
  std::vector<int> rs_long;
  std::vector<int64_t> hoffs_long;
  int64_t htotal = 0, scratch_total = 0, ytmp_total = 0;
  desc_long_.clear();
  for (int64_t i = 0; i < nring; ++i) {
    const int64_t n = nphi[i];
    HP_REQUIRE(n >= 4 && n % 4 == 0, HPSHT_ERR_UNSUPPORTED,
               "ring length must be a positive multiple of 4 (HEALPix rings always are)");
    HP_REQUIRE(n / 4 <= 8192, HPSHT_ERR_UNSUPPORTED, "ring length > 32768 (nside > 8192) is not supported");
    npix_ += n;
    RingDesc d;
    d.pixoff = ringstart[i];
    d.phi0 = phi0[i];
    {  // phi0 = pq * pi / n with an integer pq (every HEALPix ring: 0 or 1)?  The table path evaluates the
       // phases of the exact rational; the caller's double phi0 is within 2^-53 of it, so the two differ by
       // <= mmax * phi0 * 1.1e-16 rad: accepted up to 2e-15 (short polar rings keep the general path).
      const double qd = phi0[i] * (double)n / 3.14159265358979323846;
      const double qr = std::nearbyint(qd);
      const bool rational = qr >= 0.0 && qr < 2.0 * (double)n && std::fabs(qd - qr) <= 4e-15 * std::max(1.0, qr);
      d.pq = (rational && (double)(nm - 1) * std::fabs(phi0[i]) <= 16.0) ? (int)qr : -1;
      if (getenv("HPSHT_FFT_GENERAL_PHASE")) d.pq = -1;
    }
    d.ir = (int)(slot0 + i);
    d.nphi = (int)n;
    d.r = (int)(n / 4);
    const bool pow2 = (d.r & (d.r - 1)) == 0;
    d.soff = d.yoff = -1;
    if (d.r > 4096) {  // long ring: split transform through a global scratch (see k_long_*)
      d.soff = scratch_total;
      scratch_total += 2 * (int64_t)d.r;
      if (pow2) {
        d.M = 0;
        d.hoff = -1;
      } else {
        d.M = 2 * LONG_MH;
        d.yoff = ytmp_total;
        ytmp_total += 2 * (int64_t)LONG_MH;
        auto it = hoff_of_r.find(d.r);
        if (it == hoff_of_r.end()) {
          hoff_of_r[d.r] = htotal;
          rs_long.push_back(d.r);
          hoffs_long.push_back(htotal);
          d.hoff = htotal;
          htotal += d.M;
        } else {
          d.hoff = it->second;
        }
      }
      desc_long_.push_back(d);
      continue;
    }
    if (pow2) {
      d.M = 0;
      d.hoff = -1;
    } else {
      d.M = next_pow2(2 * d.r - 1);
      auto it = hoff_of_r.find(d.r);
      if (it == hoff_of_r.end()) {
        hoff_of_r[d.r] = htotal;
        rs.push_back(d.r);
        Ms.push_back(d.M);
        hoffs.push_back(htotal);
        d.hoff = htotal;
        htotal += d.M;
      } else {
        d.hoff = it->second;
      }
    }
    desc_.push_back(d);
  }
  // size classes: big CTAs for big rings; Bluestein rings need the extra work buffer
  std::stable_sort(desc_.begin(), desc_.end(), [](const RingDesc& a, const RingDesc& b) {
    const int ca = (a.M != 0) ? std::max(a.M, 2 * a.r) : 2 * a.r;
    const int cb = (b.M != 0) ? std::max(b.M, 2 * b.r) : 2 * b.r;
    if ((a.M != 0) != (b.M != 0)) return (a.M != 0) > (b.M != 0);
    return ca > cb;
  });
  // shared memory per CTA follows the largest ring of a class: keep the M = 8192 Bluestein rings (221 KB, one
  // CTA per SM) apart from the M = 4096 ones (two CTAs per SM)
  auto bucket = [](const RingDesc& d) {
    const int work = d.M ? d.M : d.r;
    const int b = work > 4096 ? 4 : (work > 2048 ? 3 : (work > 512 ? 2 : (work > 64 ? 1 : 0)));
    return (d.M ? 8 : 0) + b;
  };
  const bool use_half = !getenv("HPSHT_FFT_NO_HALF");
  for (size_t i = 0; i < desc_.size();) {
    size_t j = i;
    Class c;
    c.first = (int)i;
    const int bk = bucket(desc_[i]);
    while (j < desc_.size() && bucket(desc_[j]) == bk) {
      c.rmax = std::max(c.rmax, desc_[j].r);
      c.Mmax = std::max(c.Mmax, desc_[j].M);
      ++j;
    }
    c.count = (int)(j - i);
    const int work = std::max(c.Mmax, c.rmax);
    // the radix-16 passes hold 16 complex values per thread (k_leg2map / k_map2leg are compiled for <= 512 threads,
    // 128 registers); an 8192-point pass has 512 butterflies
    c.threads = work > 4096 ? 512 : (work > 512 ? 256 : (work > 64 ? 128 : 64));
    c.smem = smem_bytes(c.rmax, c.Mmax, c.threads);
    // power-of-two rings of some length: one CTA per half ring (k_leg2map_half / k_map2leg_half)
    c.half = use_half && c.Mmax == 0 && c.rmax >= 1024;
    if (c.half) {
      c.threads = c.rmax >= 1024 ? 256 : 128;
      c.smem = half_smem_bytes(c.rmax);
    }
    classes_.push_back(c);
    i = j;
  }
  d_desc_.upload(desc_.data(), desc_.size(), stream);
  if (!desc_long_.empty()) {
    // biggest rings first
    std::stable_sort(desc_long_.begin(), desc_long_.end(),
                     [](const RingDesc& a, const RingDesc& b) { return a.r > b.r; });
    d_desc_long_.upload(desc_long_.data(), desc_long_.size(), stream);
    scratch_cstride_ = scratch_total;
    ytmp_cstride_ = ytmp_total;
    HP_CUDA(cudaFuncSetAttribute(k_long_fft<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)long_smem_bytes(true)));
    HP_CUDA(cudaFuncSetAttribute(k_long_fft<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)long_smem_bytes(true)));
    HP_CUDA(cudaFuncSetAttribute(k_build_hhat_long, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)long_smem_bytes(true)));
  }
  size_t max_smem = 0, max_half = 0;
  for (auto& c : classes_) (c.half ? max_half : max_smem) = std::max(c.half ? max_half : max_smem, c.smem);
  HP_REQUIRE(max_smem <= 227 * 1024, HPSHT_ERR_UNSUPPORTED, "ring too long for the in-smem FFT");
  raise_smem_attr((const void*)k_leg2map, g_smem_leg2map, max_smem);
  raise_smem_attr((const void*)k_map2leg, g_smem_map2leg, max_smem);
  if (max_half) {
    raise_smem_attr((const void*)k_leg2map_half, g_smem_leg2map_half, max_half);
    raise_smem_attr((const void*)k_map2leg_half, g_smem_map2leg_half, max_half);
  }
  // chirp spectra
  d_hhat_.alloc((size_t)std::max<int64_t>(htotal, 1) * 2);
  if (!rs.empty()) {
    DevBuf<int> d_rs, d_Ms;
    DevBuf<int64_t> d_hoffs;
    d_rs.upload(rs.data(), rs.size(), stream);
    d_Ms.upload(Ms.data(), Ms.size(), stream);
    d_hoffs.upload(hoffs.data(), hoffs.size(), stream);
    int rmax = 0, Mmax = 0;
    for (size_t i = 0; i < rs.size(); ++i) {
      rmax = std::max(rmax, rs[i]);
      Mmax = std::max(Mmax, Ms[i]);
    }
    const int nt = 256;
    const size_t sm = smem_bytes(rmax, Mmax, nt);
    raise_smem_attr((const void*)k_build_hhat, g_smem_hhat, sm);
    k_build_hhat<<<(unsigned)rs.size(), nt, sm, stream>>>(
        d_rs.p, d_Ms.p, d_hoffs.p, reinterpret_cast<double2*>(d_hhat_.p),
        rmax, Mmax);
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaStreamSynchronize(stream));
  }
  if (!rs_long.empty()) {
    DevBuf<int> d_rs;
    DevBuf<int64_t> d_hoffs;
    d_rs.upload(rs_long.data(), rs_long.size(), stream);
    d_hoffs.upload(hoffs_long.data(), hoffs_long.size(), stream);
    k_build_hhat_long<<<(unsigned)rs_long.size(), 1024, long_smem_bytes(true), stream>>>(
        d_rs.p, d_hoffs.p, reinterpret_cast<double2*>(d_hhat_.p));
    HP_CUDA(cudaGetLastError());
    HP_CUDA(cudaStreamSynchronize(stream));
  }
  HP_CUDA(cudaStreamSynchronize(stream));
}

void RingFFT::ensure_long_scratch(int ncomp) {
  if (desc_long_.empty()) return;
  d_scratch_.ensure((size_t)scratch_cstride_ * 2 * ncomp);
  if (ytmp_cstride_ > 0) d_ytmp_.ensure((size_t)ytmp_cstride_ * 2 * ncomp);
}

void RingFFT::leg2map(const double* d_leg, double* d_map, int64_t map_cstride, int ncomp,
                      cudaStream_t stream) {
  for (const Class& c : classes_) {
    if (c.half) {
      k_leg2map_half<<<dim3(2u * (unsigned)c.count, (unsigned)ncomp), c.threads, c.smem, stream>>>(
          d_desc_.p, c.first, reinterpret_cast<const double2*>(d_leg), d_map, map_cstride, nm_, nring_, c.rmax);
      ++launches_;
      continue;
    }
    dim3 grid((unsigned)c.count, (unsigned)ncomp);
    k_leg2map<<<grid, c.threads, c.smem, stream>>>(
        d_desc_.p, c.first, reinterpret_cast<const double2*>(d_leg), d_map, map_cstride, nm_, nring_,
        reinterpret_cast<const double2*>(d_hhat_.p),
        c.rmax, c.Mmax);
    ++launches_;
  }
  if (!desc_long_.empty()) {
    ensure_long_scratch(ncomp);
    const unsigned nl = (unsigned)desc_long_.size();
    k_long_pre<<<dim3(nl, (unsigned)ncomp), 1024, long_smem_bytes(false), stream>>>(
        d_desc_long_.p, 0, reinterpret_cast<const double2*>(d_leg), reinterpret_cast<double2*>(d_scratch_.p),
        scratch_cstride_, nm_, nring_);
    k_long_fft<true><<<dim3(2 * nl, (unsigned)ncomp), 1024, long_smem_bytes(true), stream>>>(
        d_desc_long_.p, 0, reinterpret_cast<double2*>(d_scratch_.p), scratch_cstride_,
        reinterpret_cast<double2*>(d_ytmp_.p), ytmp_cstride_, d_map, map_cstride,
        reinterpret_cast<const double2*>(d_hhat_.p));
    launches_ += 2;
  }
  HP_CUDA(cudaGetLastError());
}

void RingFFT::map2leg(const double* d_map, int64_t map_cstride, double* d_leg, int ncomp,
                      cudaStream_t stream) {
  for (const Class& c : classes_) {
    if (c.half) {
      k_map2leg_half<<<dim3(2u * (unsigned)c.count, (unsigned)ncomp), c.threads, c.smem, stream>>>(
          d_desc_.p, c.first, d_map, map_cstride, reinterpret_cast<double2*>(d_leg), nm_, nring_, c.rmax);
      ++launches_;
      continue;
    }
    dim3 grid((unsigned)c.count, (unsigned)ncomp);
    k_map2leg<<<grid, c.threads, c.smem, stream>>>(
        d_desc_.p, c.first, d_map, map_cstride, reinterpret_cast<double2*>(d_leg), nm_, nring_,
        reinterpret_cast<const double2*>(d_hhat_.p),
        c.rmax, c.Mmax);
    ++launches_;
  }
  if (!desc_long_.empty()) {
    ensure_long_scratch(ncomp);
    const unsigned nl = (unsigned)desc_long_.size();
    k_long_fft<false><<<dim3(2 * nl, (unsigned)ncomp), 1024, long_smem_bytes(true), stream>>>(
        d_desc_long_.p, 0, reinterpret_cast<double2*>(d_scratch_.p), scratch_cstride_,
        reinterpret_cast<double2*>(d_ytmp_.p), ytmp_cstride_, const_cast<double*>(d_map), map_cstride,
        reinterpret_cast<const double2*>(d_hhat_.p));
    k_long_post<<<dim3(nl, (unsigned)ncomp), 1024, long_smem_bytes(false), stream>>>(
        d_desc_long_.p, 0, reinterpret_cast<const double2*>(d_scratch_.p), scratch_cstride_,
        reinterpret_cast<double2*>(d_leg), nm_, nring_);
    launches_ += 2;
  }
  HP_CUDA(cudaGetLastError());
}

}  // namespace hpsht
