// cpu_port.cpp — TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/sht_oracle.c header).
//
// FP64 OpenMP port of the hot path at the Ducc0 seam of reference src/sht.jl:85,93,169,178, using the
// SAME formulation as the sm_100a kernels:
//   Legendre stage (cpu_alm2leg / cpu_leg2alm): on-the-fly scaled recurrences (cos^2 two-term form for
//     spin 0, rescaled three-term form for spin +-2) in the accuracy classes of
//     csrc/legendre_tables.hpp (X / Y standard forms, DP / DE difference forms), north/south ring
//     pairing, m > mlim(theta) culling, 2^(800 k) range tracking with a warp-like block of ring pairs (4 SIMD
//     vectors for spin 0, 2 for spin 2, held in registers) advancing in lock step;
//   ring stage (cpu_leg2map / cpu_map2leg): phase shift + alias fold + a real FFT per ring through
//     one half-length complex transform (plain mixed radix, Bluestein for what is left of the length after the
//     factors 2, 3, 5); plans are built once per length and kept.
// It is (a) the "port" CPU baseline timed by bench.py and (b) a model of the kernels' arithmetic that
// can be checked against the long-double oracle without a GPU.  It is NOT ducc: expect ducc's
// hand-vectorised kernels to be faster.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "legendre_tables.hpp"

using namespace hpsht;

namespace {

typedef std::complex<double> cd;
#ifndef PORT_NV0
#define PORT_NV0 4
#endif
#ifndef PORT_NV2
#define PORT_NV2 2
#endif
constexpr int NV0 = PORT_NV0;  // SIMD vectors of ring pairs advanced together (the "warp"): spin 0
constexpr int NV2 = PORT_NV2;  //                                                          spin 2
constexpr int GROUP = 8;    // steps between range checks
constexpr int CHUNK = 32;   // table padding (multiple of GROUP)
constexpr int SC_STEP = 800;
const double SC_BIG = std::ldexp(1.0, 400);
const double SC_DOWN = std::ldexp(1.0, -800);

struct Pair {
  double cth, sth, sh, ch;  // cos/sin(theta_north), sin/cos(theta_north/2)
  double x2, s2, y2;        // cos^2, sin^2, 2 sin^2(theta/2), each rounded once from long double
  double sthl;              // sin(theta) - sth (only used by the double-double start value experiment)
  double theta;
  int64_t irN, irS;         // ring slots (irS = -1: unpaired)
  int64_t mlim0, mlim2;
};

// pair detection: theta_i + theta_j == pi within 1e-14 relative -> (north i, south j)
std::vector<Pair> make_pairs(int64_t lmax, int64_t nring, const double* theta) {
  std::vector<int64_t> idx(nring);
  for (int64_t i = 0; i < nring; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return theta[a] < theta[b]; });
  std::vector<Pair> out;
  int64_t lo = 0, hi = nring - 1;
  const double pi = 3.14159265358979323846;
  auto mk = [&](int64_t n, int64_t s) {
    Pair p;
    long double th = (long double)theta[n];
    const long double c = cosl(th), sn = sinl(th), shl = sinl(0.5L * th);
    p.cth = (double)c;
    p.sth = (double)sn;
    p.sh = (double)shl;
    p.ch = (double)cosl(0.5L * th);
    p.x2 = (double)(c * c);
    p.s2 = (double)(sn * sn);
    p.y2 = (double)(2.0L * shl * shl);
    p.sthl = (double)(sn - (long double)p.sth);
    p.theta = theta[n];
    p.irN = n;
    p.irS = s;
    p.mlim0 = mlim_rule(lmax, 0, p.sth, p.cth);
    p.mlim2 = mlim_rule(lmax, 2, p.sth, p.cth);
    return p;
  };
  while (lo <= hi) {
    int64_t a = idx[lo], b = idx[hi];
    if (lo < hi && std::fabs(theta[a] + theta[b] - pi) <= 1e-14 * pi && theta[a] < 0.5 * pi) {
      out.push_back(mk(a, b));
      ++lo;
      --hi;
    } else {
      // unpaired: take the one farther from the equator first to keep order monotone in theta
      if (std::fabs(theta[a] - 0.5 * pi) >= std::fabs(theta[b] - 0.5 * pi)) {
        out.push_back(mk(a, -1));
        ++lo;
      } else {
        out.push_back(mk(b, -1));
        --hi;
      }
    }
  }
  return out;
}

ClassRanges class_ranges(int spin, const std::vector<Pair>& pairs) {
  std::vector<double> th(pairs.size());
  for (size_t i = 0; i < pairs.size(); ++i) {
    const double t = pairs[i].theta;
    th[i] = t <= 1.5707963267948966 ? t : 3.14159265358979323846 - t;  // an unpaired south ring counts by its mirror
  }
  return make_class_ranges(spin, (int)pairs.size(), th.data());
}

// frexp for finite doubles without the library call (zero and subnormals take the library path)
inline double frexp_fast(double x, int* e) {
  uint64_t b;
  std::memcpy(&b, &x, 8);
  const int ex = (int)((b >> 52) & 0x7ff);
  if (ex == 0 || ex == 0x7ff) return std::frexp(x, e);
  *e = ex - 1022;
  b = (b & ~(0x7ffULL << 52)) | (1022ULL << 52);
  std::memcpy(&x, &b, 8);
  return x;
}

// EXPERIMENT (cpu_set_dd_start, off by default so that the port keeps modelling the kernels): the start value
// sin(theta)^m in double-double arithmetic.  Repeated squaring in double doubles the rounding error of every earlier
// squaring, i.e. the start value -- and with it the whole (m, ring) row -- is off by up to m * 1.1e-16 = 9e-13 at
// m = 8191; measured, it is what sets the worst mid-latitude rings (2.7e-13 -> 4e-14, profiles/r02_parity.md).
int g_dd_start = 0;
inline void dd_mul(double ah, double al, double bh, double bl, double& ch, double& cl) {
  const double p = ah * bh;
  const double e = std::fma(ah, bh, -p) + (ah * bl + al * bh);
  ch = p + e;
  cl = e - (ch - p);
}
inline void pow_scaled_dd(double base, double base_lo, int64_t n, double& mant, int64_t& e) {
  int be;
  double bh = frexp_fast(base, &be);
  double bl = std::ldexp(base_lo, -be);
  int64_t bexp = be;
  double mh = 1.0, ml = 0.0;
  e = 0;
  while (n > 0) {
    int t;
    if (n & 1) {
      dd_mul(mh, ml, bh, bl, mh, ml);
      e += bexp;
      const double r = frexp_fast(mh, &t);
      ml = std::ldexp(ml, -t);
      mh = r;
      e += t;
    }
    dd_mul(bh, bl, bh, bl, bh, bl);
    bexp *= 2;
    const double r = frexp_fast(bh, &t);
    bl = std::ldexp(bl, -t);
    bh = r;
    bexp += t;
    n >>= 1;
  }
  mant = mh + ml;
}
// x^n (n >= 0) as mant * 2^e with mant in [0.5,1) (or 0)
inline void pow_scaled(double base, int64_t n, double& mant, int64_t& e) {
  int be;
  double b = frexp_fast(base, &be);
  int64_t bexp = be;
  mant = 1.0;
  e = 0;
  while (n > 0) {
    if (n & 1) {
      mant *= b;
      e += bexp;
      int t;
      mant = frexp_fast(mant, &t);
      e += t;
    }
    b *= b;
    bexp *= 2;
    int t;
    b = frexp_fast(b, &t);
    bexp += t;
    n >>= 1;
  }
}

// v * 2^e -> (lam, sc) with lam = true * 2^(-800 sc), sc <= 0, |lam| < 2^400
inline void to_scaled(double v, int64_t e, double& lam, int& sc) {
  if (v == 0.0) {
    lam = 0.0;
    sc = 0;
    return;
  }
  int t;
  v = frexp_fast(v, &t);
  e += t;
  if (e >= -400) {
    sc = 0;
  } else {
    sc = -(int)((-e - 400 + SC_STEP - 1) / SC_STEP);
  }
  int64_t ee = e - (int64_t)SC_STEP * sc;
  lam = (ee < -1070) ? 0.0 : std::ldexp(v, (int)ee);
}

inline double eps_d(int64_t l, int64_t m) { return std::sqrt((double)(l * l - m * m) / (double)(4 * l * l - 1)); }

// ---------------------------------------------------------------------------------------------
// SIMD plumbing: GCC generic vectors of VL doubles (one AVX-512 register with -march=native; narrower targets get the
// same code split by the compiler).  The block kernels keep their whole state in NV such vectors held in local
// variables across the steps of a group -- as arrays on the stack every step paid a store-to-load forward on each
// accumulator chain (measured: 25 % of the FMA peak before).
// ---------------------------------------------------------------------------------------------
constexpr int VL = 8;
typedef double v8d __attribute__((vector_size(VL * sizeof(double))));
typedef long long v8i __attribute__((vector_size(VL * sizeof(long long))));
inline v8d bc(double x) { return v8d{x, x, x, x, x, x, x, x}; }
inline v8d ld8(const double* p) {
  v8d r;
  std::memcpy(&r, p, sizeof r);
  return r;
}
inline void st8(double* p, v8d v) { std::memcpy(p, &v, sizeof v); }

// out[i] = sum of the lanes of in[i], i = 0..7: three shuffle-and-add levels instead of eight horizontal sums
inline v8d reduce8(const v8d* in) {
  const v8i a1 = {0, 8, 2, 10, 4, 12, 6, 14}, b1 = {1, 9, 3, 11, 5, 13, 7, 15};
  const v8i a2 = {0, 1, 8, 9, 4, 5, 12, 13}, b2 = {2, 3, 10, 11, 6, 7, 14, 15};
  const v8i a3 = {0, 1, 2, 3, 8, 9, 10, 11}, b3 = {4, 5, 6, 7, 12, 13, 14, 15};
  v8d r[4], q[2];
  for (int k = 0; k < 4; ++k)
    r[k] = __builtin_shuffle(in[2 * k], in[2 * k + 1], a1) + __builtin_shuffle(in[2 * k], in[2 * k + 1], b1);
  for (int k = 0; k < 2; ++k)
    q[k] = __builtin_shuffle(r[2 * k], r[2 * k + 1], a2) + __builtin_shuffle(r[2 * k], r[2 * k + 1], b2);
  return __builtin_shuffle(q[0], q[1], a3) + __builtin_shuffle(q[0], q[1], b3);
}
// the sums of one group of steps: red[g][t] (t = 0..3) -> acc[4 (n0 + g) + t] += sum of lanes
inline void flush_group(const v8d* red, double* acc) {
  static_assert(GROUP == 8, "two steps (x 4 sums) per output vector");
  for (int q = 0; q < GROUP / 2; ++q) st8(acc + 8 * q, ld8(acc + 8 * q) + reduce8(red + 8 * q));
}

// one recurrence step: `u` is the value the accumulation uses; DIFF: (u, v) = (value, first difference)
template <bool DIFF>
inline void rec_step(v8d t, v8d& u, v8d& v) {
  if (DIFF) {
    v = t * u + v;
    u = u + v;
  } else {
    const v8d nx = t * u - v;
    v = u;
    u = nx;
  }
}

// does any lane that is still scaled (sc < 0) exceed the rescale threshold?
template <int WW>
inline bool needs_rescale(const double* u, const int* sc) {
  int big = 0;
#pragma omp simd reduction(| : big)
  for (int w = 0; w < WW; ++w) big |= (sc[w] < 0) & (std::fabs(u[w]) > SC_BIG);
  return big != 0;
}

// ---------------------------------------------------------------------------------------------
// spin 0, one block of <= NV*VL pairs of one class.  ADJ: leg -> acc (4 doubles / step); else stream -> leg.
// ---------------------------------------------------------------------------------------------
template <bool ADJ, bool DIFF, bool ALT, int NV>
void block_s0(const std::vector<Pair>& pairs, int64_t p0, int64_t p1, bool ysin, int64_t m, int64_t mi, int64_t nm,
              int64_t nring, int64_t npad, const double* cf, double k0, const double* stream, double* acc,
              cd* leg) {
  constexpr int WW = NV * VL;
  alignas(64) double y[WW], u[WW], v[WW], er[WW], ei[WW], orr[WW], oi[WW];
  alignas(64) double ter[WW], tei[WW], tor[WW], toi[WW];
  int sc[WW];
  bool act[WW];
  bool any_act = false;
  const int nw = (int)(p1 - p0);
  for (int w = 0; w < WW; ++w) {
    const int64_t p = std::min(p0 + w, p1 - 1);
    const Pair& pr = pairs[p];
    act[w] = w < nw && m <= pr.mlim0;
    any_act |= act[w];
    y[w] = ysin ? pr.s2 : pr.x2;
    double mant;
    int64_t e;
    if (g_dd_start) pow_scaled_dd(pr.sth, pr.sthl, m, mant, e);
    else pow_scaled(pr.sth, m, mant, e);
    to_scaled(k0 * mant, e, u[w], sc[w]);
    if (!act[w]) u[w] = 0.0, sc[w] = 0;
    v[w] = DIFF ? u[w] : 0.0;
    er[w] = ei[w] = orr[w] = oi[w] = 0.0;
    ter[w] = tei[w] = tor[w] = toi[w] = 0.0;
    if (ADJ) {
      const cd gN = act[w] ? leg[(0 * nring + pr.irN) * nm + mi] : cd(0, 0);
      const cd gS = (act[w] && pr.irS >= 0) ? leg[(0 * nring + pr.irS) * nm + mi] : cd(0, 0);
      ter[w] = gN.real() + gS.real();
      tei[w] = gN.imag() + gS.imag();
      tor[w] = pr.cth * (gN.real() - gS.real());
      toi[w] = pr.cth * (gN.imag() - gS.imag());
      if (sc[w] == 0) er[w] = ter[w], ei[w] = tei[w], orr[w] = tor[w], oi[w] = toi[w];
    }
  }
  if (!any_act) return;
  v8d Y[NV], U[NV], V[NV], ER[NV], EI[NV], OR[NV], OI[NV];
  auto load = [&]() {
    for (int k = 0; k < NV; ++k) {
      Y[k] = ld8(y + VL * k), U[k] = ld8(u + VL * k), V[k] = ld8(v + VL * k);
      ER[k] = ld8(er + VL * k), EI[k] = ld8(ei + VL * k), OR[k] = ld8(orr + VL * k), OI[k] = ld8(oi + VL * k);
    }
  };
  auto store = [&]() {
    for (int k = 0; k < NV; ++k) {
      st8(u + VL * k, U[k]), st8(v + VL * k, V[k]);
      st8(er + VL * k, ER[k]), st8(ei + VL * k, EI[k]), st8(orr + VL * k, OR[k]), st8(oi + VL * k, OI[k]);
    }
  };
  auto any_unscaled = [&]() {
    bool a = false;
    for (int w = 0; w < WW; ++w) a |= act[w] && (sc[w] == 0);
    return a;
  };
  load();
  bool any0 = any_unscaled();
  v8d red[GROUP * 4];
  for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
    if (any0) {
#pragma GCC unroll 8
      for (int g = 0; g < GROUP; ++g) {
        const int64_t n = n0 + g;
        const v8d a_ = bc(cf[2 * n]), b_ = bc(cf[2 * n + 1]);
        const double sg = (ALT && (n & 1)) ? -1.0 : 1.0;
        if (ADJ) {
          v8d s0 = bc(0.0), s1 = s0, s2 = s0, s3 = s0;
#pragma GCC unroll 4
          for (int k = 0; k < NV; ++k) {
            const v8d pv = ALT ? bc(sg) * U[k] : U[k];
            s0 = pv * ER[k] + s0, s1 = pv * EI[k] + s1, s2 = pv * OR[k] + s2, s3 = pv * OI[k] + s3;
          }
          red[4 * g + 0] = s0, red[4 * g + 1] = s1, red[4 * g + 2] = s2, red[4 * g + 3] = s3;
        } else {
          const double* s = &stream[4 * n];
          const v8d s0 = bc(sg * s[0]), s1 = bc(sg * s[1]), s2 = bc(sg * s[2]), s3 = bc(sg * s[3]);
#pragma GCC unroll 4
          for (int k = 0; k < NV; ++k) {
            ER[k] = U[k] * s0 + ER[k], EI[k] = U[k] * s1 + EI[k];
            OR[k] = U[k] * s2 + OR[k], OI[k] = U[k] * s3 + OI[k];
          }
        }
#pragma GCC unroll 4
        for (int k = 0; k < NV; ++k) rec_step<DIFF>(a_ * Y[k] + b_, U[k], V[k]);
      }
      if (ADJ) flush_group(red, acc + 4 * n0);
    } else {
#pragma GCC unroll 8
      for (int g = 0; g < GROUP; ++g) {
        const int64_t n = n0 + g;
        const v8d a_ = bc(cf[2 * n]), b_ = bc(cf[2 * n + 1]);
#pragma GCC unroll 4
        for (int k = 0; k < NV; ++k) rec_step<DIFF>(a_ * Y[k] + b_, U[k], V[k]);
      }
    }
    for (int k = 0; k < NV; ++k) st8(u + VL * k, U[k]);
    if (needs_rescale<WW>(u, sc)) {
      store();
      for (int w = 0; w < WW; ++w) {
        if (sc[w] < 0 && std::fabs(u[w]) > SC_BIG) {
          u[w] *= SC_DOWN;
          v[w] *= SC_DOWN;
          sc[w] += 1;
          if (ADJ) {
            if (sc[w] == 0) er[w] = ter[w], ei[w] = tei[w], orr[w] = tor[w], oi[w] = toi[w];
          } else {
            er[w] = ei[w] = orr[w] = oi[w] = 0.0;  // drop what was accumulated while scaled
          }
        }
      }
      load();
      any0 = any_unscaled();
    }
  }
  store();
  if (!ADJ)
    for (int w = 0; w < nw; ++w) {
      const Pair& pr = pairs[p0 + w];
      cd E(er[w], ei[w]), O(orr[w] * pr.cth, oi[w] * pr.cth);
      if (!act[w] || sc[w] < 0) E = 0, O = 0;
      leg[(0 * nring + pr.irN) * nm + mi] = E + O;
      if (pr.irS >= 0) leg[(0 * nring + pr.irS) * nm + mi] = E - O;
    }
}

// ---------------------------------------------------------------------------------------------
// spin 2, one block of <= NV*VL pairs of one class
// ---------------------------------------------------------------------------------------------
template <bool ADJ, bool DIFF, int NV>
void block_s2(const std::vector<Pair>& pairs, int64_t p0, int64_t p1, bool yform, int64_t m, int64_t mi, int64_t nm,
              int64_t nring, int64_t npad, const double* ca, const double* cf, double kp, double km,
              const double* stream, double* acc, cd* leg) {
  constexpr int WW = NV * VL;
  alignas(64) double y[WW], up[WW], vp[WW], um[WW], vm[WW];
  // synthesis: the four complex accumulators P+N, P-N, P+S, P-S; adjoint: the four complex inputs w+N, w-N, w+S,
  // w-S masked by the scale state of the lambda they multiply.  Slot order: 0 +N, 1 -N, 2 +S, 3 -S (re, im apart).
  alignas(64) double zr[4][WW], zi[4][WW];
  alignas(64) double tr[4][WW], ti[4][WW];   // adjoint: the true inputs
  int scp[WW], scm[WW];
  bool act[WW];
  bool any_act = false;
  const int nw = (int)(p1 - p0);
  const int64_t pw = m >= 2 ? m - 2 : 2 - m;
  const int q = (int)std::min<int64_t>(m, 2);
  const int64_t l0 = m > 2 ? m : 2;
  const bool odd0 = ((l0 + m) & 1) != 0;  // parity of (l+m) at the first step
  const cd I(0, 1);
  for (int w = 0; w < WW; ++w) {
    const int64_t p = std::min(p0 + w, p1 - 1);
    const Pair& pr = pairs[p];
    act[w] = w < nw && m <= pr.mlim2;
    any_act |= act[w];
    y[w] = yform ? -pr.y2 : pr.cth;
    double mant;
    int64_t e;
    if (g_dd_start) pow_scaled_dd(pr.sth, pr.sthl, pw, mant, e);
    else pow_scaled(pr.sth, pw, mant, e);
    double fs = 1.0, fc = 1.0;
    for (int k = 0; k < q; ++k) {
      fs *= pr.sh * pr.sh;
      fc *= pr.ch * pr.ch;
    }
    to_scaled(kp * mant * fs, e, up[w], scp[w]);
    to_scaled(km * mant * fc, e, um[w], scm[w]);
    if (!act[w]) up[w] = um[w] = 0.0, scp[w] = scm[w] = 0;
    vp[w] = DIFF ? up[w] : 0.0;
    vm[w] = DIFF ? um[w] : 0.0;
    for (int t = 0; t < 4; ++t) zr[t][w] = zi[t][w] = tr[t][w] = ti[t][w] = 0.0;
    if (ADJ) {
      const cd qN = act[w] ? leg[(0 * nring + pr.irN) * nm + mi] : cd(0, 0);
      const cd uN = act[w] ? leg[(1 * nring + pr.irN) * nm + mi] : cd(0, 0);
      const cd qS = (act[w] && pr.irS >= 0) ? leg[(0 * nring + pr.irS) * nm + mi] : cd(0, 0);
      const cd uS = (act[w] && pr.irS >= 0) ? leg[(1 * nring + pr.irS) * nm + mi] : cd(0, 0);
      const cd tw[4] = {qN + I * uN, qN - I * uN, qS + I * uS, qS - I * uS};
      for (int t = 0; t < 4; ++t) tr[t][w] = tw[t].real(), ti[t][w] = tw[t].imag();
      // G+ += lam+ w+N + sgn lam- w+S ; G- += lam- w-N + sgn lam+ w-S
      if (scp[w] == 0) zr[0][w] = tr[0][w], zi[0][w] = ti[0][w], zr[3][w] = tr[3][w], zi[3][w] = ti[3][w];
      if (scm[w] == 0) zr[1][w] = tr[1][w], zi[1][w] = ti[1][w], zr[2][w] = tr[2][w], zi[2][w] = ti[2][w];
    }
  }
  if (!any_act) return;
  v8d Y[NV], UP[NV], VP[NV], UM[NV], VM[NV], ZR[4][NV], ZI[4][NV];
  auto load = [&]() {
    for (int k = 0; k < NV; ++k) {
      Y[k] = ld8(y + VL * k);
      UP[k] = ld8(up + VL * k), VP[k] = ld8(vp + VL * k), UM[k] = ld8(um + VL * k), VM[k] = ld8(vm + VL * k);
      for (int t = 0; t < 4; ++t) ZR[t][k] = ld8(zr[t] + VL * k), ZI[t][k] = ld8(zi[t] + VL * k);
    }
  };
  auto store = [&]() {
    for (int k = 0; k < NV; ++k) {
      st8(up + VL * k, UP[k]), st8(vp + VL * k, VP[k]), st8(um + VL * k, UM[k]), st8(vm + VL * k, VM[k]);
      for (int t = 0; t < 4; ++t) st8(zr[t] + VL * k, ZR[t][k]), st8(zi[t] + VL * k, ZI[t][k]);
    }
  };
  auto any_unscaled = [&]() {
    bool a = false;
    for (int w = 0; w < WW; ++w) a |= act[w] && ((scp[w] == 0) | (scm[w] == 0));
    return a;
  };
  load();
  bool any0 = any_unscaled();
  v8d red[GROUP * 4];
  for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
#pragma GCC unroll 8
    for (int g = 0; g < GROUP; ++g) {
      const int64_t j = n0 + g;
      const v8d a_ = bc(ca[j]), cp = bc(cf[2 * j]), cm = bc(cf[2 * j + 1]);
      const double sgn = (odd0 ^ ((j & 1) != 0)) ? -1.0 : 1.0;
      if (any0) {
        if (ADJ) {
          v8d s0 = bc(0.0), s1 = s0, s2 = s0, s3 = s0;
          const v8d sg = bc(sgn);
#pragma GCC unroll 4
          for (int k = 0; k < NV; ++k) {
            const v8d sm_ = sg * UM[k], sp_ = sg * UP[k];
            s0 = UP[k] * ZR[0][k] + (sm_ * ZR[2][k] + s0);
            s1 = UP[k] * ZI[0][k] + (sm_ * ZI[2][k] + s1);
            s2 = UM[k] * ZR[1][k] + (sp_ * ZR[3][k] + s2);
            s3 = UM[k] * ZI[1][k] + (sp_ * ZI[3][k] + s3);
          }
          red[4 * g + 0] = s0, red[4 * g + 1] = s1, red[4 * g + 2] = s2, red[4 * g + 3] = s3;
        } else {
          const double* s = &stream[4 * j];
          const v8d s0 = bc(s[0]), s1 = bc(s[1]), s2 = bc(s[2]), s3 = bc(s[3]);
          const v8d t0 = bc(sgn * s[0]), t1 = bc(sgn * s[1]), t2 = bc(sgn * s[2]), t3 = bc(sgn * s[3]);
#pragma GCC unroll 4
          for (int k = 0; k < NV; ++k) {
            ZR[0][k] = UP[k] * s0 + ZR[0][k], ZI[0][k] = UP[k] * s1 + ZI[0][k];   // P+N
            ZR[1][k] = UM[k] * s2 + ZR[1][k], ZI[1][k] = UM[k] * s3 + ZI[1][k];   // P-N
            ZR[2][k] = UM[k] * t0 + ZR[2][k], ZI[2][k] = UM[k] * t1 + ZI[2][k];   // P+S
            ZR[3][k] = UP[k] * t2 + ZR[3][k], ZI[3][k] = UP[k] * t3 + ZI[3][k];   // P-S
          }
        }
      }
#pragma GCC unroll 4
      for (int k = 0; k < NV; ++k) {
        rec_step<DIFF>(a_ * Y[k] + cp, UP[k], VP[k]);
        rec_step<DIFF>(a_ * Y[k] + cm, UM[k], VM[k]);
      }
    }
    if (ADJ && any0) flush_group(red, acc + 4 * n0);
    for (int k = 0; k < NV; ++k) st8(up + VL * k, UP[k]), st8(um + VL * k, UM[k]);
    if (needs_rescale<WW>(up, scp) || needs_rescale<WW>(um, scm)) {
      store();
      for (int w = 0; w < WW; ++w) {
        if (scp[w] < 0 && std::fabs(up[w]) > SC_BIG) {
          up[w] *= SC_DOWN;
          vp[w] *= SC_DOWN;
          scp[w] += 1;
          if (ADJ) {
            if (scp[w] == 0) zr[0][w] = tr[0][w], zi[0][w] = ti[0][w], zr[3][w] = tr[3][w], zi[3][w] = ti[3][w];
          } else {
            zr[0][w] = zi[0][w] = zr[3][w] = zi[3][w] = 0.0;
          }
        }
        if (scm[w] < 0 && std::fabs(um[w]) > SC_BIG) {
          um[w] *= SC_DOWN;
          vm[w] *= SC_DOWN;
          scm[w] += 1;
          if (ADJ) {
            if (scm[w] == 0) zr[1][w] = tr[1][w], zi[1][w] = ti[1][w], zr[2][w] = tr[2][w], zi[2][w] = ti[2][w];
          } else {
            zr[1][w] = zi[1][w] = zr[2][w] = zi[2][w] = 0.0;
          }
        }
      }
      load();
      any0 = any_unscaled();
    }
  }
  store();
  if (!ADJ)
    for (int w = 0; w < nw; ++w) {
      const Pair& pr = pairs[p0 + w];
      cd a(zr[0][w], zi[0][w]), b(zr[1][w], zi[1][w]), c(zr[2][w], zi[2][w]), d(zr[3][w], zi[3][w]);
      if (!act[w] || scp[w] < 0) a = 0, d = 0;
      if (!act[w] || scm[w] < 0) b = 0, c = 0;
      const cd mi_(0, -1);
      leg[(0 * nring + pr.irN) * nm + mi] = a + b;
      leg[(1 * nring + pr.irN) * nm + mi] = mi_ * (a - b);
      if (pr.irS >= 0) {
        leg[(0 * nring + pr.irS) * nm + mi] = c + d;
        leg[(1 * nring + pr.irS) * nm + mi] = mi_ * (c - d);
      }
    }
}

// all classes of one m.  T holds the tables of this m alone (built by the calling thread); mi is the column of
// the leg array.
template <bool ADJ>
void run_m(int spin, const LegendreTables& T, const std::vector<Pair>& pairs, const ClassRanges& cr, int64_t mi,
           int64_t nm, int64_t nring, const double* stream, double* acc, cd* leg) {
  const int64_t m = T.mval[0];
  for (int c = 0; c < LC_COUNT; ++c) {
    if (spin != 0 && c == LC_DE) continue;
    const bool yside = (c == LC_Y || c == LC_DP);
    const int W = (spin == 0 ? NV0 : NV2) * VL;
    for (int64_t p0 = cr.lo(c); p0 < cr.hi(c); p0 += W) {
      const int64_t p1 = std::min<int64_t>(p0 + W, cr.hi(c));
      if (spin == 0) {
        const int64_t npad = round_up(T.nstep0[0], CHUNK);
        const double* cf = &T.coef0[c][(size_t)T.coff0[c][0] * 2];
        const double k0 = T.k0[0];
        if (c == LC_DE)
          block_s0<ADJ, true, true, NV0>(pairs, p0, p1, false, m, mi, nm, nring, npad, cf, k0, stream, acc, leg);
        else if (c == LC_DP)
          block_s0<ADJ, true, false, NV0>(pairs, p0, p1, true, m, mi, nm, nring, npad, cf, k0, stream, acc, leg);
        else
          block_s0<ADJ, false, false, NV0>(pairs, p0, p1, yside, m, mi, nm, nring, npad, cf, k0, stream, acc, leg);
      } else {
        const int64_t npad = round_up(T.nstep2[0], CHUNK);
        const double* ca = &T.coefa2[(size_t)T.off2[0]];
        const double* cf = &T.coef2[c][(size_t)T.coff2[c][0] * 2];
        if (c == LC_DP)
          block_s2<ADJ, true, NV2>(pairs, p0, p1, true, m, mi, nm, nring, npad, ca, cf, T.kp2[0], T.km2[0], stream,
                                   acc, leg);
        else
          block_s2<ADJ, false, NV2>(pairs, p0, p1, yside, m, mi, nm, nring, npad, ca, cf, T.kp2[0], T.km2[0],
                                    stream, acc, leg);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ring stage: plain recursive mixed-radix complex FFT (radix 4/2/3/5, generic odd radix up to 64); what is
// left of n after the small factors are split off (a ring has n = 4 r with r often prime) goes through a
// Bluestein convolution of that length only
// ---------------------------------------------------------------------------------------------
struct FftPlan {
  int64_t n = 0;
  std::vector<cd> w;          // exp(-2 pi i k / n)
  int64_t leaf = 0;           // n with the factors 2, 3, 5 removed, when that has a prime factor > 64 (else 0)
  int64_t M = 0;              // Bluestein length for the leaf (power of two >= 2 leaf - 1)
  std::vector<cd> chirp, chirpF;
  const FftPlan* sub = nullptr;  // plan of length M (owned by the plan store, shared by every leaf of that M)
};
void fft_rec(const FftPlan& P, const cd* in, int64_t istride, cd* out, int64_t n, int64_t wstride, std::vector<cd>& work);
// leaf DFT by Bluestein: out[k] = sum_j in[j istride] exp(-2 pi i jk / L)
void bluestein_leaf(const FftPlan& P, const cd* in, int64_t istride, cd* out, std::vector<cd>& work) {
  const int64_t L = P.leaf, M = P.M;
  std::vector<cd> sub_work;
  const size_t base = work.size();
  work.resize(base + (size_t)(2 * M));
  cd* a = work.data() + base;
  cd* A = a + M;
  for (int64_t k = 0; k < L; ++k) a[k] = in[k * istride] * P.chirp[(size_t)k];
  for (int64_t k = L; k < M; ++k) a[k] = cd(0, 0);
  fft_rec(*P.sub, a, 1, A, M, 1, sub_work);
  for (int64_t k = 0; k < M; ++k) a[k] = std::conj(A[k] * P.chirpF[(size_t)k]);
  fft_rec(*P.sub, a, 1, A, M, 1, sub_work);
  const double inv = 1.0 / (double)M;
  for (int64_t k = 0; k < L; ++k) out[k] = std::conj(A[k]) * inv * P.chirp[(size_t)k];
  work.resize(base);
}
void fft_rec(const FftPlan& P, const cd* in, int64_t istride, cd* out, int64_t n, int64_t wstride, std::vector<cd>& work) {
  if (n == 1) {
    out[0] = in[0];
    return;
  }
  if (P.leaf > 0 && n == P.leaf) {
    bluestein_leaf(P, in, istride, out, work);
    return;
  }
  int64_t r = 0;
  for (int64_t p : {4, 2, 3, 5}) if (n % p == 0) { r = p; break; }
  if (r == 0) for (r = 7; n % r; r += 2) {}
  const int64_t mth = n / r;
  for (int64_t k = 0; k < r; ++k) fft_rec(P, in + k * istride, istride * r, out + k * mth, mth, wstride * r, work);
  // decimation in time: X[q + mth*s] = sum_k W_n^{k q} W_r^{k s} Y_k[q]
  const cd* w = P.w.data();
  if (r == 2) {
    for (int64_t q2 = 0; q2 < mth; ++q2) {
      const cd a = out[q2], b = out[mth + q2] * w[q2 * wstride];
      out[q2] = a + b;
      out[q2 + mth] = a - b;
    }
  } else if (r == 4) {
    for (int64_t q2 = 0; q2 < mth; ++q2) {
      const cd a = out[q2], b = out[mth + q2] * w[q2 * wstride], c = out[2 * mth + q2] * w[2 * q2 * wstride],
               d = out[3 * mth + q2] * w[3 * q2 * wstride];
      const cd apc = a + c, amc = a - c, bpd = b + d, bmd = b - d;
      const cd ibmd(-bmd.imag(), bmd.real());  // i (b - d)
      out[q2] = apc + bpd;
      out[q2 + mth] = amc - ibmd;
      out[q2 + 2 * mth] = apc - bpd;
      out[q2 + 3 * mth] = amc + ibmd;
    }
  } else {
    cd t[64], o[64];
    const int64_t rs = P.n / r;
    for (int64_t q2 = 0; q2 < mth; ++q2) {
      for (int64_t k = 0; k < r; ++k) t[k] = out[k * mth + q2] * w[k * q2 * wstride];
      for (int64_t s = 0; s < r; ++s) {
        cd acc = t[0];
        for (int64_t k = 1; k < r; ++k) acc += t[k] * w[((k * s) % r) * rs];
        o[s] = acc;
      }
      for (int64_t s = 0; s < r; ++s) out[q2 + mth * s] = o[s];
    }
  }
}
// exp(-2 pi i k / n), k = 0 .. count-1
void fill_twiddles(int64_t n, int64_t count, std::vector<cd>& w);
int64_t bluestein_leaf_of(int64_t n) {  // n with 2, 3, 5 removed when that keeps a prime factor > 64, else 0
  int64_t L = n;
  for (int64_t p : {2, 3, 5})
    while (L % p == 0) L /= p;
  int64_t big = 1, t = L;
  for (int64_t p = 7; p * p <= t; p += 2)
    while (t % p == 0) big = std::max(big, p), t /= p;
  big = std::max(big, t);
  return (L > 1 && big > 64) ? L : 0;
}
int64_t bluestein_len(int64_t L) {
  int64_t M = 1;
  while (M < 2 * L - 1) M *= 2;
  return M;
}
FftPlan* make_plan(int64_t n, const FftPlan* sub) {
  FftPlan* P = new FftPlan();
  P->n = n;
  fill_twiddles(n, n, P->w);
  const int64_t L = bluestein_leaf_of(n);
  if (L) {
    P->leaf = L;
    P->M = bluestein_len(L);
    P->sub = sub;
    std::vector<cd> w2;
    fill_twiddles(2 * L, 2 * L, w2);   // exp(-i pi q / L)
    P->chirp.resize((size_t)L);
    for (int64_t k = 0; k < L; ++k) P->chirp[(size_t)k] = w2[(size_t)((k * k) % (2 * L))];
    std::vector<cd> b((size_t)P->M, cd(0, 0)), wk;
    for (int64_t k = 0; k < L; ++k) {
      b[(size_t)k] = std::conj(P->chirp[(size_t)k]);
      if (k) b[(size_t)(P->M - k)] = b[(size_t)k];
    }
    P->chirpF.resize((size_t)P->M);
    fft_rec(*P->sub, b.data(), 1, P->chirpF.data(), P->M, 1, wk);
  }
  return P;
}
// exp(-2 pi i k / n) = A[k / S] * B[k % S]: 2 sqrt(n) long-double evaluations and `count` long-double products, each
// entry still within an ulp (the product is formed in 64-bit mantissas before rounding)
void fill_twiddles(int64_t n, int64_t count, std::vector<cd>& w) {
  const long double tp = 2.0L * LT_PI / (long double)n;
  w.resize((size_t)count);
  int64_t S = 1;
  while (S * S < count) S *= 2;
  std::vector<long double> ac((size_t)((count + S - 1) / S)), as(ac.size()), bcs((size_t)S), bsn((size_t)S);
  for (size_t i = 0; i < ac.size(); ++i) ac[i] = cosl(tp * (long double)(i * S)), as[i] = sinl(tp * (long double)(i * S));
  for (int64_t j = 0; j < S; ++j) bcs[(size_t)j] = cosl(tp * (long double)j), bsn[(size_t)j] = sinl(tp * (long double)j);
  for (int64_t k = 0; k < count; ++k) {
    const size_t i = (size_t)(k / S), j = (size_t)(k % S);
    w[(size_t)k] = cd((double)(ac[i] * bcs[j] - as[i] * bsn[j]), (double)-(as[i] * bcs[j] + ac[i] * bsn[j]));
  }
}
// forward DFT (e^{-i}); inverse via conjugation by the caller
void fft_fwd(const FftPlan& P, const cd* in, cd* out, std::vector<cd>& work) {
  work.clear();
  fft_rec(P, in, 1, out, P.n, 1, work);
}

// A ring of n = 2h real points is transformed through ONE complex FFT of length h (even / odd samples packed as
// real / imaginary part) plus a split pass.
struct RingPlan {
  int64_t n = 0;
  const FftPlan* half = nullptr;   // length n / 2
  std::vector<cd> wn;              // exp(-2 pi i k / n), k = 0 .. n/2
};
// Plans live for the life of the process (a production library keeps its plans between calls too); lookups after
// prepare() are read-only.
struct PlanStore {
  std::vector<int64_t> flen, rlen;   // sorted keys
  std::vector<FftPlan*> fplan;
  std::vector<RingPlan*> rplan;
  const FftPlan* fft(int64_t n) const {
    const size_t i = (size_t)(std::lower_bound(flen.begin(), flen.end(), n) - flen.begin());
    return (i < flen.size() && flen[i] == n) ? fplan[i] : nullptr;
  }
  const RingPlan* ring(int64_t n) const {
    const size_t i = (size_t)(std::lower_bound(rlen.begin(), rlen.end(), n) - rlen.begin());
    return (i < rlen.size() && rlen[i] == n) ? rplan[i] : nullptr;
  }
  template <class T>
  static void merge(std::vector<int64_t>& keys, std::vector<T*>& vals, const std::vector<int64_t>& nk,
                    const std::vector<T*>& nv) {
    std::vector<std::pair<int64_t, T*>> all;
    for (size_t i = 0; i < keys.size(); ++i) all.emplace_back(keys[i], vals[i]);
    for (size_t i = 0; i < nk.size(); ++i) all.emplace_back(nk[i], nv[i]);
    std::sort(all.begin(), all.end());
    keys.clear(), vals.clear();
    for (auto& kv : all) keys.push_back(kv.first), vals.push_back(kv.second);
  }
  // make sure a ring plan exists for every length in nphi (all even)
  void prepare(int64_t nring, const int64_t* nphi) {
    std::vector<int64_t> want(nphi, nphi + nring);
    std::sort(want.begin(), want.end());
    want.erase(std::unique(want.begin(), want.end()), want.end());
    std::vector<int64_t> miss;
    for (int64_t n : want)
      if (!ring(n)) miss.push_back(n);
    if (miss.empty()) return;
    // the power-of-two transforms behind the Bluestein leaves first (a handful), then the half-length plans
    std::vector<int64_t> subs;
    for (int64_t n : miss) {
      const int64_t L = bluestein_leaf_of(n / 2);
      if (L && !fft(bluestein_len(L))) subs.push_back(bluestein_len(L));
    }
    std::sort(subs.begin(), subs.end());
    subs.erase(std::unique(subs.begin(), subs.end()), subs.end());
    std::vector<FftPlan*> sp(subs.size());
    for (size_t i = 0; i < subs.size(); ++i) sp[i] = make_plan(subs[i], nullptr);
    merge(flen, fplan, subs, sp);
    std::vector<int64_t> halves;
    for (int64_t n : miss)
      if (!fft(n / 2)) halves.push_back(n / 2);
    std::vector<FftPlan*> hp(halves.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < (int64_t)halves.size(); ++i) {
      const int64_t L = bluestein_leaf_of(halves[(size_t)i]);
      hp[(size_t)i] = make_plan(halves[(size_t)i], L ? fft(bluestein_len(L)) : nullptr);
    }
    merge(flen, fplan, halves, hp);
    std::vector<RingPlan*> rp(miss.size());
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)miss.size(); ++i) {
      RingPlan* R = new RingPlan();
      R->n = miss[(size_t)i];
      R->half = fft(R->n / 2);
      fill_twiddles(R->n, R->n / 2 + 1, R->wn);
      rp[(size_t)i] = R;
    }
    merge(rlen, rplan, miss, rp);
  }
};
PlanStore g_plans;

// x[0..n) real from the half-complex spectrum hc[0..n/2] (Im hc[0], Im hc[n/2] ignored):
//   x[j] = sum_{k=0}^{n-1} F[k] e^{+2 pi i jk/n},  F[k] = hc[k] (k <= n/2), F[n-k] = conj(hc[k])
// Z[k] = (F[k] + conj(F[h-k])) + i e^{+2 pi i k/n} (F[k] - conj(F[h-k])),  z = IDFT_h(Z),  x[2j] + i x[2j+1] = z[j]
void ring_c2r(const RingPlan& R, const cd* hc, double* x, std::vector<cd>& a, std::vector<cd>& b, std::vector<cd>& work) {
  const int64_t h = R.n / 2;
  a.resize((size_t)h);
  b.resize((size_t)h);
  for (int64_t k = 0; k < h; ++k) {
    const cd Fk = k == 0 ? cd(hc[0].real(), 0) : hc[k];
    const cd Fq = std::conj(k == 0 ? cd(hc[h].real(), 0) : hc[h - k]);
    const cd s = Fk + Fq, d = (Fk - Fq) * std::conj(R.wn[(size_t)k]);   // conj(wn) = e^{+2 pi i k/n}
    const cd Z = s + cd(-d.imag(), d.real());
    a[(size_t)k] = std::conj(Z);   // inverse transform through the forward one
  }
  fft_fwd(*R.half, a.data(), b.data(), work);
  for (int64_t j = 0; j < h; ++j) x[2 * j] = b[(size_t)j].real(), x[2 * j + 1] = -b[(size_t)j].imag();
}
// X[k] = sum_j x[j] e^{-2 pi i jk/n} for k = 0 .. n/2 from the real x[0..n)
void ring_r2c(const RingPlan& R, const double* x, cd* X, std::vector<cd>& a, std::vector<cd>& b, std::vector<cd>& work) {
  const int64_t h = R.n / 2;
  a.resize((size_t)h);
  b.resize((size_t)h);
  for (int64_t j = 0; j < h; ++j) a[(size_t)j] = cd(x[2 * j], x[2 * j + 1]);
  fft_fwd(*R.half, a.data(), b.data(), work);
  for (int64_t k = 0; k <= h; ++k) {
    const cd Zk = b[(size_t)(k == h ? 0 : k)], Zq = std::conj(b[(size_t)(k == 0 ? 0 : h - k)]);
    const cd E = 0.5 * (Zk + Zq), D = Zk - Zq;          // O = D / (2i)
    const cd O(0.5 * D.imag(), -0.5 * D.real());
    X[k] = E + R.wn[(size_t)k] * O;
  }
}

}  // namespace

extern "C" {

int cpu_max_threads(void) { return omp_get_max_threads(); }
void cpu_set_dd_start(int on) { g_dd_start = on; }
void cpu_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

// culled / dense step counts of the work model for a theta list and m list
void cpu_legendre_steps(int spin, int64_t lmax, int64_t nm, const int64_t* mval, int64_t nring,
                        const double* theta, double* culled, double* dense) {
  std::vector<Pair> pairs = make_pairs(lmax, nring, theta);
  double c = 0, d = 0;
  for (int64_t i = 0; i < nm; ++i) {
    int64_t m = mval[i];
    for (auto& p : pairs) {
      d += (double)(lmax - m + 1);
      if (m <= (spin ? p.mlim2 : p.mlim0)) c += (double)(lmax - m + 1);
    }
  }
  *culled = c;
  *dense = d;
}

// alm [ncomp][alm_cstride] complex128, mstart 0-based; leg [ncomp][nring][nm] complex128
void cpu_alm2leg(int spin, int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                 const double* alm_, int64_t alm_cstride, int64_t nring, const double* theta,
                 double* leg_) {
  const cd* alm = reinterpret_cast<const cd*>(alm_);
  cd* leg = reinterpret_cast<cd*>(leg_);
  const int ncomp = spin ? 2 : 1;
  std::vector<Pair> pairs = make_pairs(lmax, nring, theta);
  const ClassRanges cr = class_ranges(spin, pairs);
  std::memset(leg_, 0, sizeof(double) * 2 * (size_t)ncomp * nring * nm);
#pragma omp parallel
  {
    std::vector<double> stream;  // per-m prepared alm stream
    LegendreTables T;            // per-m coefficient tables, built by the thread that uses them
#pragma omp for schedule(dynamic, 1)
    for (int64_t mi = 0; mi < nm; ++mi) {
      const int64_t m = mval[mi];
      build_legendre_tables(lmax, &m, 1, CHUNK, spin == 0, spin != 0, T);
      if (spin == 0) {
        const int64_t nn = T.nstep0[0];
        if (nn == 0) continue;
        const double* is = &T.invs0[0];
        stream.assign((size_t)round_up(nn, CHUNK) * 4, 0.0);
        const cd* a = alm + mstart[mi];
        for (int64_t n = 0; n < nn; ++n) {  // Ae_n, Ao_n
          const int64_t l = m + 2 * n;
          const cd al = a[l];
          const cd al1 = (l + 1 <= lmax) ? a[l + 1] : cd(0, 0);
          const cd al2 = (l + 2 <= lmax) ? a[l + 2] : cd(0, 0);
          const cd A = (eps_d(l + 1, m) * al + eps_d(l + 2, m) * al2) * is[n];
          const cd B = al1 * is[n];
          stream[4 * n + 0] = A.real();
          stream[4 * n + 1] = A.imag();
          stream[4 * n + 2] = B.real();
          stream[4 * n + 3] = B.imag();
        }
      } else {
        const int64_t nl = T.nstep2[0];
        if (nl == 0) continue;
        const int64_t l0 = m > 2 ? m : 2;
        const double* sg = &T.sig2[0];
        stream.assign((size_t)round_up(nl, CHUNK) * 4, 0.0);
        const cd* ea = alm + mstart[mi];
        const cd* ba = alm + alm_cstride + mstart[mi];
        for (int64_t j = 0; j < nl; ++j) {  // alpha+ = -(E+iB) sigma/2, alpha- = -(E-iB) sigma/2
          const int64_t l = l0 + j;
          const cd e = ea[l], b = ba[l];
          const cd ib(-b.imag(), b.real());
          const cd ap = -0.5 * sg[j] * (e + ib), am = -0.5 * sg[j] * (e - ib);
          stream[4 * j + 0] = ap.real();
          stream[4 * j + 1] = ap.imag();
          stream[4 * j + 2] = am.real();
          stream[4 * j + 3] = am.imag();
        }
      }
      run_m<false>(spin, T, pairs, cr, mi, nm, nring, stream.data(), nullptr, leg);
    }
  }
}

// leg [ncomp][nring][nm] complex128 -> alm [ncomp][alm_cstride] (only l=m..lmax of owned m written;
// spin 2: l < 2 entries set to 0)
void cpu_leg2alm(int spin, int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                 const double* leg_, int64_t nring, const double* theta, double* alm_,
                 int64_t alm_cstride) {
  cd* leg = const_cast<cd*>(reinterpret_cast<const cd*>(leg_));
  cd* alm = reinterpret_cast<cd*>(alm_);
  std::vector<Pair> pairs = make_pairs(lmax, nring, theta);
  const ClassRanges cr = class_ranges(spin, pairs);
#pragma omp parallel
  {
    std::vector<double> acc;
    LegendreTables T;
#pragma omp for schedule(dynamic, 1)
    for (int64_t mi = 0; mi < nm; ++mi) {
      const int64_t m = mval[mi];
      build_legendre_tables(lmax, &m, 1, CHUNK, spin == 0, spin != 0, T);
      if (spin == 0) {
        const int64_t nn = T.nstep0[0];
        if (nn == 0) continue;
        const double* is = &T.invs0[0];
        acc.assign((size_t)round_up(nn, CHUNK) * 4, 0.0);
        run_m<true>(spin, T, pairs, cr, mi, nm, nring, nullptr, acc.data(), leg);
        // a_l (l=m+2n) = eps_{l+1} At_n/g_n + eps_l At_{n-1}/g_{n-1}; a_{l+1} = Bt_n/g_n
        cd* a = alm + mstart[mi];
        for (int64_t n = 0; n < nn; ++n) {
          const int64_t l = m + 2 * n;
          const cd At(acc[4 * n], acc[4 * n + 1]), Bt(acc[4 * n + 2], acc[4 * n + 3]);
          cd v = eps_d(l + 1, m) * is[n] * At;
          if (n > 0) v += eps_d(l, m) * is[n - 1] * cd(acc[4 * (n - 1)], acc[4 * (n - 1) + 1]);
          a[l] = v;
          if (l + 1 <= lmax) a[l + 1] = is[n] * Bt;
        }
      } else {
        cd* ea = alm + mstart[mi];
        cd* ba = alm + alm_cstride + mstart[mi];
        for (int64_t l = m; l <= lmax && l < 2; ++l) ea[l] = 0, ba[l] = 0;
        const int64_t nl = T.nstep2[0];
        if (nl == 0) continue;
        const int64_t l0 = m > 2 ? m : 2;
        const double* sg = &T.sig2[0];
        acc.assign((size_t)round_up(nl, CHUNK) * 4, 0.0);  // G+ (re,im), G- (re,im) per l
        run_m<true>(spin, T, pairs, cr, mi, nm, nring, nullptr, acc.data(), leg);
        for (int64_t j = 0; j < nl; ++j) {  // E = -sigma/2 (G+ + G-), B = i sigma/2 (G+ - G-)
          const int64_t l = l0 + j;
          const cd Gp(acc[4 * j], acc[4 * j + 1]), Gm(acc[4 * j + 2], acc[4 * j + 3]);
          ea[l] = -0.5 * sg[j] * (Gp + Gm);
          ba[l] = cd(0, 0.5 * sg[j]) * (Gp - Gm);
        }
      }
    }
  }
}

// Build (and keep) the FFT plans of these ring lengths, so that a timed region does not contain plan construction.
// Ring lengths must be even (HEALPix: n = 4 r).
int cpu_prepare_rings(int64_t nring, const int64_t* nphi) {
  for (int64_t i = 0; i < nring; ++i)
    if (nphi[i] <= 0 || (nphi[i] & 1)) return -1;
  g_plans.prepare(nring, nphi);
  return 0;
}

// leg [ncomp][nring][nm] -> map [ncomp][map_cstride]   (ringstart 0-based)
void cpu_leg2map(int ncomp, int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                 const int64_t* ringstart, const double* leg_, double* map, int64_t map_cstride) {
  const cd* leg = reinterpret_cast<const cd*>(leg_);
  if (cpu_prepare_rings(nring, nphi)) std::abort();
#pragma omp parallel
  {
    std::vector<cd> hc, a, b, work;
#pragma omp for schedule(dynamic, 4) collapse(2)
    for (int c = 0; c < ncomp; ++c)
      for (int64_t ir = 0; ir < nring; ++ir) {
        const int64_t n = nphi[ir], h = n / 2;
        const cd* row = leg + ((size_t)c * nring + ir) * nm;
        hc.assign((size_t)(h + 1), cd(0, 0));
        hc[0] = row[0].real();
        const cd step(std::cos(phi0[ir]), std::sin(phi0[ir]));
        cd cur(1, 0);
        int64_t i1 = 0;   // m mod n, kept incrementally
        for (int64_t m = 1; m < nm; ++m) {
          // e^{i m phi0}: recomputed every 64 steps, multiplied up in between
          if ((m & 63) == 1) cur = cd(std::cos(m * phi0[ir]), std::sin(m * phi0[ir]));
          else cur *= step;
          if (++i1 == n) i1 = 0;
          const cd t = row[m] * cur;
          if (i1 <= h) hc[(size_t)i1] += t;
          if (i1 == 0) hc[0] += std::conj(t);
          else if (i1 >= h) hc[(size_t)(n - i1)] += std::conj(t);
        }
        ring_c2r(*g_plans.ring(n), hc.data(), map + (size_t)c * map_cstride + ringstart[ir], a, b, work);
      }
  }
}

// map -> leg
void cpu_map2leg(int ncomp, int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                 const int64_t* ringstart, const double* map, int64_t map_cstride, double* leg_) {
  cd* leg = reinterpret_cast<cd*>(leg_);
  if (cpu_prepare_rings(nring, nphi)) std::abort();
#pragma omp parallel
  {
    std::vector<cd> X, a, b, work;
#pragma omp for schedule(dynamic, 4) collapse(2)
    for (int c = 0; c < ncomp; ++c)
      for (int64_t ir = 0; ir < nring; ++ir) {
        const int64_t n = nphi[ir], h = n / 2;
        X.resize((size_t)(h + 1));
        ring_r2c(*g_plans.ring(n), map + (size_t)c * map_cstride + ringstart[ir], X.data(), a, b, work);
        cd* row = leg + ((size_t)c * nring + ir) * nm;
        const cd step(std::cos(phi0[ir]), -std::sin(phi0[ir]));
        cd cur(1, 0);
        int64_t i1 = 0;
        for (int64_t m = 0; m < nm; ++m) {
          if ((m & 63) == 0) cur = cd(std::cos(m * phi0[ir]), -std::sin(m * phi0[ir]));
          row[m] = (i1 <= h ? X[(size_t)i1] : std::conj(X[(size_t)(n - i1)])) * cur;
          cur *= step;
          if (++i1 == n) i1 = 0;
        }
      }
  }
}

}  // extern "C"
