// cpu_port.cpp — TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/sht_oracle.c header).
//
// FP64 OpenMP port of the hot path at the Ducc0 seam of reference src/sht.jl:85,93,169,178, using the
// SAME formulation as the sm_100a kernels:
//   Legendre stage (cpu_alm2leg / cpu_leg2alm): on-the-fly scaled recurrences (cos^2 two-term form for
//     spin 0, rescaled three-term form for spin +-2) in the accuracy classes of
//     csrc/legendre_tables.hpp (X / Y standard forms, DP / DE difference forms), north/south ring
//     pairing, m > mlim(theta) culling, 2^(800 k) range tracking with a warp-like block of W ring pairs
//     advancing in lock step;
//   ring stage (cpu_leg2map / cpu_map2leg): phase shift + alias fold + a plain
//     mixed-radix / Bluestein complex FFT per ring.
// It is (a) the "port" CPU baseline timed by bench.py and (b) a model of the kernels' arithmetic that
// can be checked against the long-double oracle without a GPU.  It is NOT ducc: expect ducc's
// hand-vectorised kernels to be faster.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

#include "legendre_tables.hpp"

using namespace hpsht;

namespace {

typedef std::complex<double> cd;
constexpr int W = 16;       // ring pairs advanced together (the "warp")
constexpr int GROUP = 8;    // steps between range checks
constexpr int CHUNK = 32;   // table padding (multiple of GROUP)
constexpr int SC_STEP = 800;
const double SC_BIG = std::ldexp(1.0, 400);
const double SC_DOWN = std::ldexp(1.0, -800);

struct Pair {
  double cth, sth, sh, ch;  // cos/sin(theta_north), sin/cos(theta_north/2)
  double x2, s2, y2;        // cos^2, sin^2, 2 sin^2(theta/2), each rounded once from long double
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

// x^n (n >= 0) as mant * 2^e with mant in [0.5,1) (or 0)
inline void pow_scaled(double base, int64_t n, double& mant, int64_t& e) {
  int be;
  double b = std::frexp(base, &be);
  int64_t bexp = be;
  mant = 1.0;
  e = 0;
  while (n > 0) {
    if (n & 1) {
      mant *= b;
      e += bexp;
      int t;
      mant = std::frexp(mant, &t);
      e += t;
    }
    b *= b;
    bexp *= 2;
    int t;
    b = std::frexp(b, &t);
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
  v = std::frexp(v, &t);
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

// one recurrence step: `u` is the value the accumulation uses; DIFF: (u, v) = (value, first difference)
template <bool DIFF>
inline void rec_step(double t, double& u, double& v) {
  if (DIFF) {
    v = std::fma(t, u, v);
    u = u + v;
  } else {
    const double nx = std::fma(t, u, -v);
    v = u;
    u = nx;
  }
}

// ---------------------------------------------------------------------------------------------
// spin 0, one block of <= W pairs of one class.  ADJ: leg -> acc (4 doubles / step); else stream -> leg.
// ---------------------------------------------------------------------------------------------
template <bool ADJ, bool DIFF, bool ALT>
void block_s0(const std::vector<Pair>& pairs, int64_t p0, int64_t p1, bool ysin, int64_t m, int64_t mi, int64_t nm,
              int64_t nring, int64_t npad, const double* cf, double k0, const double* stream, double* acc,
              cd* leg) {
  double y[W], u[W], v[W], er[W], ei[W], orr[W], oi[W];
  double ter[W], tei[W], tor[W], toi[W];
  int sc[W];
  bool act[W];
  bool any_act = false;
  const int nw = (int)(p1 - p0);
  for (int w = 0; w < W; ++w) {
    const int64_t p = std::min(p0 + w, p1 - 1);
    const Pair& pr = pairs[p];
    act[w] = w < nw && m <= pr.mlim0;
    any_act |= act[w];
    y[w] = ysin ? pr.s2 : pr.x2;
    double mant;
    int64_t e;
    pow_scaled(pr.sth, m, mant, e);
    to_scaled(k0 * mant, e, u[w], sc[w]);
    if (!act[w]) u[w] = 0.0, sc[w] = 0;
    v[w] = DIFF ? u[w] : 0.0;
    er[w] = ei[w] = orr[w] = oi[w] = 0.0;
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
  for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
    bool any0 = false;
    for (int w = 0; w < W; ++w) any0 |= act[w] && (sc[w] == 0);
    for (int g = 0; g < GROUP; ++g) {
      const int64_t n = n0 + g;
      const double a_ = cf[2 * n], b_ = cf[2 * n + 1];
      const double sg = (ALT && (n & 1)) ? -1.0 : 1.0;
      if (ADJ) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        if (any0) {
#pragma omp simd reduction(+ : s0, s1, s2, s3)
          for (int w = 0; w < W; ++w) {
            const double pv = sg * u[w];
            s0 += pv * er[w];
            s1 += pv * ei[w];
            s2 += pv * orr[w];
            s3 += pv * oi[w];
          }
        }
        for (int w = 0; w < W; ++w) rec_step<DIFF>(std::fma(a_, y[w], b_), u[w], v[w]);
        double* d = &acc[4 * n];
        d[0] += s0, d[1] += s1, d[2] += s2, d[3] += s3;
      } else {
        const double* s = &stream[4 * n];
        for (int w = 0; w < W; ++w) {
          if (any0) {
            const double pv = sg * u[w];
            er[w] = std::fma(pv, s[0], er[w]);
            ei[w] = std::fma(pv, s[1], ei[w]);
            orr[w] = std::fma(pv, s[2], orr[w]);
            oi[w] = std::fma(pv, s[3], oi[w]);
          }
          rec_step<DIFF>(std::fma(a_, y[w], b_), u[w], v[w]);
        }
      }
    }
    for (int w = 0; w < W; ++w) {
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
  }
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
// spin 2, one block of <= W pairs of one class
// ---------------------------------------------------------------------------------------------
template <bool ADJ, bool DIFF>
void block_s2(const std::vector<Pair>& pairs, int64_t p0, int64_t p1, bool yform, int64_t m, int64_t mi, int64_t nm,
              int64_t nring, int64_t npad, const double* ca, const double* cf, double kp, double km,
              const double* stream, double* acc, cd* leg) {
  double y[W], up[W], vp[W], um[W], vm[W];
  cd PpN[W], PmN[W], PpS[W], PmS[W];              // synthesis accumulators
  cd twpN[W], twmN[W], twpS[W], twmS[W];          // adjoint: true inputs
  cd wpN[W], wmN[W], wpS[W], wmS[W];              // adjoint: masked by the scale state of the lam they multiply
  int scp[W], scm[W];
  bool act[W];
  bool any_act = false;
  const int nw = (int)(p1 - p0);
  const int64_t pw = m >= 2 ? m - 2 : 2 - m;
  const int q = (int)std::min<int64_t>(m, 2);
  const int64_t l0 = m > 2 ? m : 2;
  const bool odd0 = ((l0 + m) & 1) != 0;  // parity of (l+m) at the first step
  const cd I(0, 1);
  for (int w = 0; w < W; ++w) {
    const int64_t p = std::min(p0 + w, p1 - 1);
    const Pair& pr = pairs[p];
    act[w] = w < nw && m <= pr.mlim2;
    any_act |= act[w];
    y[w] = yform ? -pr.y2 : pr.cth;
    double mant;
    int64_t e;
    pow_scaled(pr.sth, pw, mant, e);
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
    PpN[w] = PmN[w] = PpS[w] = PmS[w] = 0;
    if (ADJ) {
      const cd qN = act[w] ? leg[(0 * nring + pr.irN) * nm + mi] : cd(0, 0);
      const cd uN = act[w] ? leg[(1 * nring + pr.irN) * nm + mi] : cd(0, 0);
      const cd qS = (act[w] && pr.irS >= 0) ? leg[(0 * nring + pr.irS) * nm + mi] : cd(0, 0);
      const cd uS = (act[w] && pr.irS >= 0) ? leg[(1 * nring + pr.irS) * nm + mi] : cd(0, 0);
      twpN[w] = qN + I * uN;
      twmN[w] = qN - I * uN;
      twpS[w] = qS + I * uS;
      twmS[w] = qS - I * uS;
      // G+ += lam+ w+N + sgn lam- w+S ; G- += lam- w-N + sgn lam+ w-S
      wpN[w] = scp[w] == 0 ? twpN[w] : cd(0, 0);
      wmS[w] = scp[w] == 0 ? twmS[w] : cd(0, 0);
      wmN[w] = scm[w] == 0 ? twmN[w] : cd(0, 0);
      wpS[w] = scm[w] == 0 ? twpS[w] : cd(0, 0);
    }
  }
  if (!any_act) return;
  for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
    bool any0 = false;
    for (int w = 0; w < W; ++w) any0 |= act[w] && ((scp[w] == 0) | (scm[w] == 0));
    for (int g = 0; g < GROUP; ++g) {
      const int64_t j = n0 + g;
      const double a_ = ca[j], cp = cf[2 * j], cm = cf[2 * j + 1];
      const double sgn = (odd0 ^ ((j & 1) != 0)) ? -1.0 : 1.0;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      const double* s = ADJ ? nullptr : &stream[4 * j];
      if (ADJ && any0) {
        const double* ap = reinterpret_cast<const double*>(wpN);   // (re, im) pairs
        const double* bp = reinterpret_cast<const double*>(wpS);
        const double* cp2 = reinterpret_cast<const double*>(wmN);
        const double* dp = reinterpret_cast<const double*>(wmS);
#pragma omp simd reduction(+ : s0, s1, s2, s3)
        for (int w = 0; w < W; ++w) {
          const double sm_ = sgn * um[w], sp_ = sgn * up[w];
          s0 += up[w] * ap[2 * w] + sm_ * bp[2 * w];
          s1 += up[w] * ap[2 * w + 1] + sm_ * bp[2 * w + 1];
          s2 += um[w] * cp2[2 * w] + sp_ * dp[2 * w];
          s3 += um[w] * cp2[2 * w + 1] + sp_ * dp[2 * w + 1];
        }
      }
      for (int w = 0; w < W; ++w) {
        if (any0) {
          if (ADJ) {
          } else {
            PpN[w] = cd(std::fma(up[w], s[0], PpN[w].real()), std::fma(up[w], s[1], PpN[w].imag()));
            PmN[w] = cd(std::fma(um[w], s[2], PmN[w].real()), std::fma(um[w], s[3], PmN[w].imag()));
            PpS[w] = cd(std::fma(sgn * um[w], s[0], PpS[w].real()), std::fma(sgn * um[w], s[1], PpS[w].imag()));
            PmS[w] = cd(std::fma(sgn * up[w], s[2], PmS[w].real()), std::fma(sgn * up[w], s[3], PmS[w].imag()));
          }
        }
        rec_step<DIFF>(std::fma(a_, y[w], cp), up[w], vp[w]);
        rec_step<DIFF>(std::fma(a_, y[w], cm), um[w], vm[w]);
      }
      if (ADJ) {
        double* d = &acc[4 * j];
        d[0] += s0, d[1] += s1, d[2] += s2, d[3] += s3;
      }
    }
    for (int w = 0; w < W; ++w) {
      if (scp[w] < 0 && std::fabs(up[w]) > SC_BIG) {
        up[w] *= SC_DOWN;
        vp[w] *= SC_DOWN;
        scp[w] += 1;
        if (ADJ) {
          if (scp[w] == 0) wpN[w] = twpN[w], wmS[w] = twmS[w];
        } else {
          PpN[w] = PmS[w] = 0;
        }
      }
      if (scm[w] < 0 && std::fabs(um[w]) > SC_BIG) {
        um[w] *= SC_DOWN;
        vm[w] *= SC_DOWN;
        scm[w] += 1;
        if (ADJ) {
          if (scm[w] == 0) wmN[w] = twmN[w], wpS[w] = twpS[w];
        } else {
          PmN[w] = PpS[w] = 0;
        }
      }
    }
  }
  if (!ADJ)
    for (int w = 0; w < nw; ++w) {
      const Pair& pr = pairs[p0 + w];
      cd a = PpN[w], b = PmN[w], c = PpS[w], d = PmS[w];
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

// all classes of one m
template <bool ADJ>
void run_m(int spin, const LegendreTables& T, const std::vector<Pair>& pairs, const ClassRanges& cr, int64_t mi,
           int64_t nm, int64_t nring, const double* stream, double* acc, cd* leg) {
  const int64_t m = T.mval[(size_t)mi];
  for (int c = 0; c < LC_COUNT; ++c) {
    if (spin != 0 && c == LC_DE) continue;
    const bool yside = (c == LC_Y || c == LC_DP);
    for (int64_t p0 = cr.lo(c); p0 < cr.hi(c); p0 += W) {
      const int64_t p1 = std::min<int64_t>(p0 + W, cr.hi(c));
      if (spin == 0) {
        const int64_t npad = round_up(T.nstep0[(size_t)mi], CHUNK);
        const double* cf = &T.coef0[c][(size_t)T.coff0[c][(size_t)mi] * 2];
        const double k0 = T.k0[(size_t)mi];
        if (c == LC_DE)
          block_s0<ADJ, true, true>(pairs, p0, p1, false, m, mi, nm, nring, npad, cf, k0, stream, acc, leg);
        else if (c == LC_DP)
          block_s0<ADJ, true, false>(pairs, p0, p1, true, m, mi, nm, nring, npad, cf, k0, stream, acc, leg);
        else
          block_s0<ADJ, false, false>(pairs, p0, p1, yside, m, mi, nm, nring, npad, cf, k0, stream, acc, leg);
      } else {
        const int64_t npad = round_up(T.nstep2[(size_t)mi], CHUNK);
        const double* ca = &T.coefa2[(size_t)T.off2[(size_t)mi]];
        const double* cf = &T.coef2[c][(size_t)T.coff2[c][(size_t)mi] * 2];
        if (c == LC_DP)
          block_s2<ADJ, true>(pairs, p0, p1, true, m, mi, nm, nring, npad, ca, cf, T.kp2[(size_t)mi], T.km2[(size_t)mi],
                              stream, acc, leg);
        else
          block_s2<ADJ, false>(pairs, p0, p1, yside, m, mi, nm, nring, npad, ca, cf, T.kp2[(size_t)mi],
                               T.km2[(size_t)mi], stream, acc, leg);
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
  FftPlan* sub = nullptr;     // plan of length M
  ~FftPlan() { delete sub; }
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
FftPlan* make_plan(int64_t n) {
  FftPlan* P = new FftPlan();
  P->n = n;
  const long double tp = 2.0L * LT_PI / (long double)n;
  P->w.resize((size_t)n);
  for (int64_t k = 0; k < n; ++k) P->w[(size_t)k] = cd((double)cosl(tp * k), (double)-sinl(tp * k));
  int64_t L = n;
  for (int64_t p : {2, 3, 5})
    while (L % p == 0) L /= p;
  int64_t big = 1, t = L;
  for (int64_t p = 7; p * p <= t; p += 2)
    while (t % p == 0) big = std::max(big, p), t /= p;
  big = std::max(big, t);
  if (L > 1 && big > 64) {
    P->leaf = L;
    P->M = 1;
    while (P->M < 2 * L - 1) P->M *= 2;
    P->sub = make_plan(P->M);
    P->chirp.resize((size_t)L);
    for (int64_t k = 0; k < L; ++k) {
      const int64_t kk = (k * k) % (2 * L);
      const long double a = LT_PI * (long double)kk / (long double)L;
      P->chirp[(size_t)k] = cd((double)cosl(a), (double)-sinl(a));
    }
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
// forward DFT (e^{-i}); inverse via conjugation by the caller
void fft_fwd(const FftPlan& P, const cd* in, cd* out, std::vector<cd>& work) {
  work.clear();
  fft_rec(P, in, 1, out, P.n, 1, work);
}

struct PlanCache {
  std::vector<int64_t> lens;
  std::vector<FftPlan*> plans;
  const FftPlan& get(int64_t n) const {
    const size_t i = (size_t)(std::lower_bound(lens.begin(), lens.end(), n) - lens.begin());
    return *plans[i];
  }
  ~PlanCache() {
    for (auto* p : plans) delete p;
  }
};
void build_cache(PlanCache& pc, int64_t nring, const int64_t* nphi) {
  pc.lens.assign(nphi, nphi + nring);
  std::sort(pc.lens.begin(), pc.lens.end());
  pc.lens.erase(std::unique(pc.lens.begin(), pc.lens.end()), pc.lens.end());
  pc.plans.resize(pc.lens.size());
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < (int64_t)pc.lens.size(); ++i) pc.plans[(size_t)i] = make_plan(pc.lens[(size_t)i]);
}

}  // namespace

extern "C" {

int cpu_max_threads(void) { return omp_get_max_threads(); }
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
  LegendreTables T;
  build_legendre_tables(lmax, mval, nm, CHUNK, spin == 0, spin != 0, T);
  std::memset(leg_, 0, sizeof(double) * 2 * (size_t)ncomp * nring * nm);
#pragma omp parallel
  {
    std::vector<double> stream;  // per-m prepared alm stream
#pragma omp for schedule(dynamic, 1)
    for (int64_t mi = 0; mi < nm; ++mi) {
      const int64_t m = mval[mi];
      if (spin == 0) {
        const int64_t nn = T.nstep0[mi];
        if (nn == 0) continue;
        const double* is = &T.invs0[(size_t)T.off0[mi]];
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
        const int64_t nl = T.nstep2[mi];
        if (nl == 0) continue;
        const int64_t l0 = m > 2 ? m : 2;
        const double* sg = &T.sig2[(size_t)T.off2[mi]];
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
  LegendreTables T;
  build_legendre_tables(lmax, mval, nm, CHUNK, spin == 0, spin != 0, T);
#pragma omp parallel
  {
    std::vector<double> acc;
#pragma omp for schedule(dynamic, 1)
    for (int64_t mi = 0; mi < nm; ++mi) {
      const int64_t m = mval[mi];
      if (spin == 0) {
        const int64_t nn = T.nstep0[mi];
        if (nn == 0) continue;
        const double* is = &T.invs0[(size_t)T.off0[mi]];
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
        const int64_t nl = T.nstep2[mi];
        if (nl == 0) continue;
        const int64_t l0 = m > 2 ? m : 2;
        const double* sg = &T.sig2[(size_t)T.off2[mi]];
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

// leg [ncomp][nring][nm] -> map [ncomp][map_cstride]   (ringstart 0-based)
void cpu_leg2map(int ncomp, int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                 const int64_t* ringstart, const double* leg_, double* map, int64_t map_cstride) {
  const cd* leg = reinterpret_cast<const cd*>(leg_);
  PlanCache pc;
  build_cache(pc, nring, nphi);
#pragma omp parallel
  {
    std::vector<cd> hc, full, out, work;
#pragma omp for schedule(dynamic, 4) collapse(2)
    for (int c = 0; c < ncomp; ++c)
      for (int64_t ir = 0; ir < nring; ++ir) {
        const int64_t n = nphi[ir];
        const cd* row = leg + ((size_t)c * nring + ir) * nm;
        hc.assign((size_t)(n / 2 + 1), cd(0, 0));
        hc[0] = row[0].real();
        const cd step(std::cos(phi0[ir]), std::sin(phi0[ir]));
        cd cur(1, 0);
        for (int64_t m = 1; m < nm; ++m) {
          // e^{i m phi0}: recomputed every 64 steps, multiplied up in between
          if ((m & 63) == 1) cur = cd(std::cos(m * phi0[ir]), std::sin(m * phi0[ir]));
          else cur *= step;
          const cd t = row[m] * cur;
          const int64_t i1 = m % n, i2 = (n - i1) % n;
          if (i1 <= n / 2) hc[(size_t)i1] += t;
          if (i2 <= n / 2) hc[(size_t)i2] += std::conj(t);
        }
        // ring[j] = sum_k c_k Re(hc[k] e^{2 pi i j k / n}): Hermitian-extend, conj, forward FFT, conj
        full.assign((size_t)n, cd(0, 0));
        full[0] = hc[0].real();
        for (int64_t k = 1; k <= n / 2; ++k) {
          if (2 * k == n) {
            full[(size_t)k] = hc[(size_t)k].real();
          } else {
            full[(size_t)k] = std::conj(hc[(size_t)k]);
            full[(size_t)(n - k)] = hc[(size_t)k];
          }
        }
        out.resize((size_t)n);
        fft_fwd(pc.get(n), full.data(), out.data(), work);
        double* dst = map + (size_t)c * map_cstride + ringstart[ir];
        for (int64_t j = 0; j < n; ++j) dst[j] = out[(size_t)j].real();
      }
  }
}

// map -> leg
void cpu_map2leg(int ncomp, int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                 const int64_t* ringstart, const double* map, int64_t map_cstride, double* leg_) {
  cd* leg = reinterpret_cast<cd*>(leg_);
  PlanCache pc;
  build_cache(pc, nring, nphi);
#pragma omp parallel
  {
    std::vector<cd> in, X, work;
#pragma omp for schedule(dynamic, 4) collapse(2)
    for (int c = 0; c < ncomp; ++c)
      for (int64_t ir = 0; ir < nring; ++ir) {
        const int64_t n = nphi[ir];
        const double* src = map + (size_t)c * map_cstride + ringstart[ir];
        in.resize((size_t)n);
        for (int64_t j = 0; j < n; ++j) in[(size_t)j] = src[j];
        X.resize((size_t)n);
        fft_fwd(pc.get(n), in.data(), X.data(), work);
        cd* row = leg + ((size_t)c * nring + ir) * nm;
        const cd step(std::cos(phi0[ir]), -std::sin(phi0[ir]));
        cd cur(1, 0);
        for (int64_t m = 0; m < nm; ++m) {
          if ((m & 63) == 0) cur = cd(std::cos(m * phi0[ir]), -std::sin(m * phi0[ir]));
          row[m] = X[(size_t)(m % n)] * cur;
          cur *= step;
        }
      }
  }
}

}  // extern "C"
