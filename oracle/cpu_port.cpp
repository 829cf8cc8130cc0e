// cpu_port.cpp — TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/sht_oracle.c header).
//
// FP64 OpenMP port of the Legendre stage (alm2leg / leg2alm at the Ducc0 seam of reference
// src/sht.jl:85,178) using the SAME formulation as the sm_100a kernels: on-the-fly scaled
// recurrences (cos^2 two-term form for spin 0, rescaled three-term form for spin +-2), north/south
// ring pairing, m > mlim(theta) culling, 2^(800 k) range tracking with a warp-like block of W ring
// pairs advancing in lock step.  It is (a) the "port" CPU baseline timed by bench.py and (b) a
// bit-for-bit-comparable model of the kernels' control flow that can be checked against the
// long-double oracle without a GPU.  It is NOT ducc: expect ducc's hand-vectorised kernels to be
// faster.
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

constexpr int W = 16;       // ring pairs advanced together (the "warp")
constexpr int GROUP = 8;    // steps between range checks
constexpr int CHUNK = 32;   // table padding (multiple of GROUP)
constexpr int SC_STEP = 800;
const double SC_BIG = std::ldexp(1.0, 400);
const double SC_DOWN = std::ldexp(1.0, -800);

struct Pair {
  double cth, sth, sh, ch;  // cos/sin(theta_north), sin/cos(theta_north/2)
  int64_t irN, irS;         // ring slots (irS = -1: unpaired)
  int64_t mlim0, mlim2;
};

// pair detection: theta_i + theta_j == pi within 1e-14 relative -> (north i, south j)
std::vector<Pair> make_pairs(int64_t lmax, int64_t nring, const double* theta) {
  std::vector<int64_t> idx(nring);
  for (int64_t i = 0; i < nring; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return theta[a] < theta[b]; });
  std::vector<Pair> out;
  int64_t lo = 0, hi = nring - 1;
  const double pi = 3.14159265358979323846;
  auto mk = [&](int64_t n, int64_t s) {
    Pair p;
    long double th = (long double)theta[n];
    p.cth = (double)cosl(th);
    p.sth = (double)sinl(th);
    p.sh = (double)sinl(0.5L * th);
    p.ch = (double)cosl(0.5L * th);
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

}  // namespace

extern "C" {

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
  typedef std::complex<double> cd;
  const cd* alm = reinterpret_cast<const cd*>(alm_);
  cd* leg = reinterpret_cast<cd*>(leg_);
  const int ncomp = spin ? 2 : 1;
  std::vector<Pair> pairs = make_pairs(lmax, nring, theta);
  const int64_t np = (int64_t)pairs.size();
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
        const int64_t npad = round_up(nn, CHUNK);
        const double* cf = &T.coef0[(size_t)T.off0[mi] * 2];
        const double* is = &T.invs0[(size_t)T.off0[mi]];
        // ---- prep: A_n, B_n ----
        stream.assign((size_t)npad * 4, 0.0);
        const cd* a = alm + mstart[mi];
        for (int64_t n = 0; n < nn; ++n) {
          int64_t l = m + 2 * n;
          cd al = a[l];
          cd al1 = (l + 1 <= lmax) ? a[l + 1] : cd(0, 0);
          cd al2 = (l + 2 <= lmax) ? a[l + 2] : cd(0, 0);
          double e1 = std::sqrt((double)((l + 1) * (l + 1) - m * m) / (double)(4 * (l + 1) * (l + 1) - 1));
          double e2 = std::sqrt((double)((l + 2) * (l + 2) - m * m) / (double)(4 * (l + 2) * (l + 2) - 1));
          cd A = (e1 * al + e2 * al2) * is[n];
          cd B = al1 * is[n];
          stream[4 * n + 0] = A.real();
          stream[4 * n + 1] = A.imag();
          stream[4 * n + 2] = B.real();
          stream[4 * n + 3] = B.imag();
        }
        // ---- main: blocks of W pairs ----
        for (int64_t p0 = 0; p0 < np; p0 += W) {
          double csq[W], lam1[W], lam2[W], er[W], ei[W], orr[W], oi[W];
          int sc[W];
          bool act[W];
          bool any_act = false;
          for (int w = 0; w < W; ++w) {
            int64_t p = p0 + w;
            act[w] = p < np && m <= pairs[p].mlim0;
            any_act |= act[w];
            const Pair& pr = pairs[std::min(p, np - 1)];
            csq[w] = pr.cth * pr.cth;
            double mant;
            int64_t e;
            pow_scaled(pr.sth, m, mant, e);
            to_scaled(T.k0[mi] * mant, e, lam2[w], sc[w]);
            lam1[w] = 0.0;
            er[w] = ei[w] = orr[w] = oi[w] = 0.0;
          }
          if (!any_act) continue;
          for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
            bool any0 = false;
            for (int w = 0; w < W; ++w) any0 |= (sc[w] == 0);
            if (any0) {
              for (int g = 0; g < GROUP; ++g) {
                const double a_ = cf[2 * (n0 + g)], b_ = cf[2 * (n0 + g) + 1];
                const double* s = &stream[4 * (n0 + g)];
                for (int w = 0; w < W; ++w) {
                  er[w] = std::fma(lam2[w], s[0], er[w]);
                  ei[w] = std::fma(lam2[w], s[1], ei[w]);
                  orr[w] = std::fma(lam2[w], s[2], orr[w]);
                  oi[w] = std::fma(lam2[w], s[3], oi[w]);
                  double t = std::fma(a_, csq[w], b_);
                  double nx = std::fma(t, lam2[w], lam1[w]);
                  lam1[w] = lam2[w];
                  lam2[w] = nx;
                }
              }
            } else {
              for (int g = 0; g < GROUP; ++g) {
                const double a_ = cf[2 * (n0 + g)], b_ = cf[2 * (n0 + g) + 1];
                for (int w = 0; w < W; ++w) {
                  double t = std::fma(a_, csq[w], b_);
                  double nx = std::fma(t, lam2[w], lam1[w]);
                  lam1[w] = lam2[w];
                  lam2[w] = nx;
                }
              }
            }
            for (int w = 0; w < W; ++w) {
              if (sc[w] < 0 && std::fabs(lam2[w]) > SC_BIG) {
                lam1[w] *= SC_DOWN;
                lam2[w] *= SC_DOWN;
                sc[w] += 1;
                er[w] = ei[w] = orr[w] = oi[w] = 0.0;  // drop garbage accumulated while scaled
              }
            }
          }
          for (int w = 0; w < W; ++w) {
            int64_t p = p0 + w;
            if (p >= np) break;
            const Pair& pr = pairs[p];
            cd E(er[w], ei[w]), O(orr[w] * pr.cth, oi[w] * pr.cth);
            if (!act[w] || sc[w] < 0) {
              E = 0;
              O = 0;
            }
            leg[(0 * nring + pr.irN) * nm + mi] = E + O;
            if (pr.irS >= 0) leg[(0 * nring + pr.irS) * nm + mi] = E - O;
          }
        }
      } else {
        const int64_t nl = T.nstep2[mi];
        if (nl == 0) continue;
        const int64_t l0 = m > 2 ? m : 2;
        const int64_t npad = round_up(nl, CHUNK);
        const double* cf = &T.coef2[(size_t)T.off2[mi] * 2];
        const double* sg = &T.sig2[(size_t)T.off2[mi]];
        // ---- prep: alpha+ = -(E+iB) sigma/2, alpha- = -(E-iB) sigma/2 ----
        stream.assign((size_t)npad * 4, 0.0);
        const cd* ea = alm + mstart[mi];
        const cd* ba = alm + alm_cstride + mstart[mi];
        for (int64_t j = 0; j < nl; ++j) {
          int64_t l = l0 + j;
          cd e = ea[l], b = ba[l];
          cd ib(-b.imag(), b.real());
          cd ap = -0.5 * sg[j] * (e + ib), am = -0.5 * sg[j] * (e - ib);
          stream[4 * j + 0] = ap.real();
          stream[4 * j + 1] = ap.imag();
          stream[4 * j + 2] = am.real();
          stream[4 * j + 3] = am.imag();
        }
        const int64_t pw = m >= 2 ? m - 2 : 2 - m;
        const int q = (int)std::min<int64_t>(m, 2);
        const bool odd0 = ((l0 + m) & 1) != 0;  // parity of (l+m) at the first step
        for (int64_t p0 = 0; p0 < np; p0 += W) {
          double x[W], lp1[W], lp2[W], lm1[W], lm2[W];
          // accumulators: P+N, P-N use (lam+, lam-); P+S, P-S use (lam-, lam+) with (-1)^(l+m)
          double ppn_r[W], ppn_i[W], pmn_r[W], pmn_i[W], pps_r[W], pps_i[W], pms_r[W], pms_i[W];
          int scp[W], scm[W];
          bool act[W];
          bool any_act = false;
          for (int w = 0; w < W; ++w) {
            int64_t p = p0 + w;
            act[w] = p < np && m <= pairs[p].mlim2;
            any_act |= act[w];
            const Pair& pr = pairs[std::min(p, np - 1)];
            x[w] = pr.cth;
            double mant;
            int64_t e;
            pow_scaled(pr.sth, pw, mant, e);
            double fs = 1.0, fc = 1.0;
            for (int k = 0; k < q; ++k) {
              fs *= pr.sh * pr.sh;
              fc *= pr.ch * pr.ch;
            }
            to_scaled(T.kp2[mi] * mant * fs, e, lp2[w], scp[w]);
            to_scaled(T.km2[mi] * mant * fc, e, lm2[w], scm[w]);
            lp1[w] = lm1[w] = 0.0;
            ppn_r[w] = ppn_i[w] = pmn_r[w] = pmn_i[w] = 0.0;
            pps_r[w] = pps_i[w] = pms_r[w] = pms_i[w] = 0.0;
          }
          if (!any_act) continue;
          for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
            bool any0 = false;
            for (int w = 0; w < W; ++w) any0 |= (scp[w] == 0) | (scm[w] == 0);
            for (int g = 0; g < GROUP; ++g) {
              const int64_t j = n0 + g;
              const double a_ = cf[2 * j], b_ = cf[2 * j + 1];
              const double* s = &stream[4 * j];
              const bool odd = odd0 ^ ((j & 1) != 0);
              const double sgn = odd ? -1.0 : 1.0;
              for (int w = 0; w < W; ++w) {
                if (any0) {
                  ppn_r[w] = std::fma(lp2[w], s[0], ppn_r[w]);
                  ppn_i[w] = std::fma(lp2[w], s[1], ppn_i[w]);
                  pmn_r[w] = std::fma(lm2[w], s[2], pmn_r[w]);
                  pmn_i[w] = std::fma(lm2[w], s[3], pmn_i[w]);
                  pps_r[w] = std::fma(sgn * lm2[w], s[0], pps_r[w]);
                  pps_i[w] = std::fma(sgn * lm2[w], s[1], pps_i[w]);
                  pms_r[w] = std::fma(sgn * lp2[w], s[2], pms_r[w]);
                  pms_i[w] = std::fma(sgn * lp2[w], s[3], pms_i[w]);
                }
                double tp = std::fma(x[w], a_, b_);
                double tm = std::fma(x[w], a_, -b_);
                double np_ = std::fma(tp, lp2[w], -lp1[w]);
                double nm_ = std::fma(tm, lm2[w], -lm1[w]);
                lp1[w] = lp2[w];
                lp2[w] = np_;
                lm1[w] = lm2[w];
                lm2[w] = nm_;
              }
            }
            for (int w = 0; w < W; ++w) {
              if (scp[w] < 0 && std::fabs(lp2[w]) > SC_BIG) {
                lp1[w] *= SC_DOWN;
                lp2[w] *= SC_DOWN;
                scp[w] += 1;
                ppn_r[w] = ppn_i[w] = pms_r[w] = pms_i[w] = 0.0;
              }
              if (scm[w] < 0 && std::fabs(lm2[w]) > SC_BIG) {
                lm1[w] *= SC_DOWN;
                lm2[w] *= SC_DOWN;
                scm[w] += 1;
                pmn_r[w] = pmn_i[w] = pps_r[w] = pps_i[w] = 0.0;
              }
            }
          }
          for (int w = 0; w < W; ++w) {
            int64_t p = p0 + w;
            if (p >= np) break;
            const Pair& pr = pairs[p];
            cd PpN(ppn_r[w], ppn_i[w]), PmN(pmn_r[w], pmn_i[w]);
            cd PpS(pps_r[w], pps_i[w]), PmS(pms_r[w], pms_i[w]);
            if (!act[w] || scp[w] < 0) {
              PpN = 0;
              PmS = 0;
            }
            if (!act[w] || scm[w] < 0) {
              PmN = 0;
              PpS = 0;
            }
            const cd mi_(0, -1);
            leg[(0 * nring + pr.irN) * nm + mi] = PpN + PmN;
            leg[(1 * nring + pr.irN) * nm + mi] = mi_ * (PpN - PmN);
            if (pr.irS >= 0) {
              leg[(0 * nring + pr.irS) * nm + mi] = PpS + PmS;
              leg[(1 * nring + pr.irS) * nm + mi] = mi_ * (PpS - PmS);
            }
          }
        }
      }
    }
  }
}

// leg [ncomp][nring][nm] complex128 -> alm [ncomp][alm_cstride] (only l=m..lmax of owned m written;
// spin 2: l < 2 entries set to 0)
void cpu_leg2alm(int spin, int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                 const double* leg_, int64_t nring, const double* theta, double* alm_,
                 int64_t alm_cstride) {
  typedef std::complex<double> cd;
  const cd* leg = reinterpret_cast<const cd*>(leg_);
  cd* alm = reinterpret_cast<cd*>(alm_);
  std::vector<Pair> pairs = make_pairs(lmax, nring, theta);
  const int64_t np = (int64_t)pairs.size();
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
        const int64_t npad = round_up(nn, CHUNK);
        const double* cf = &T.coef0[(size_t)T.off0[mi] * 2];
        const double* is = &T.invs0[(size_t)T.off0[mi]];
        acc.assign((size_t)npad * 4, 0.0);
        for (int64_t p0 = 0; p0 < np; p0 += W) {
          double csq[W], lam1[W], lam2[W], er[W], ei[W], orr[W], oi[W];
          double ter[W], tei[W], tor[W], toi[W];  // true inputs
          int sc[W];
          bool any_act = false;
          for (int w = 0; w < W; ++w) {
            int64_t p = p0 + w;
            bool act = p < np && m <= pairs[p].mlim0;
            any_act |= act;
            const Pair& pr = pairs[std::min(p, np - 1)];
            csq[w] = pr.cth * pr.cth;
            double mant;
            int64_t e;
            pow_scaled(pr.sth, m, mant, e);
            to_scaled(T.k0[mi] * mant, e, lam2[w], sc[w]);
            lam1[w] = 0.0;
            cd gN = act ? leg[(0 * nring + pr.irN) * nm + mi] : cd(0, 0);
            cd gS = (act && pr.irS >= 0) ? leg[(0 * nring + pr.irS) * nm + mi] : cd(0, 0);
            ter[w] = gN.real() + gS.real();
            tei[w] = gN.imag() + gS.imag();
            tor[w] = pr.cth * (gN.real() - gS.real());
            toi[w] = pr.cth * (gN.imag() - gS.imag());
            bool on = sc[w] == 0;
            er[w] = on ? ter[w] : 0.0;
            ei[w] = on ? tei[w] : 0.0;
            orr[w] = on ? tor[w] : 0.0;
            oi[w] = on ? toi[w] : 0.0;
          }
          if (!any_act) continue;
          for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
            bool any0 = false;
            for (int w = 0; w < W; ++w) any0 |= (sc[w] == 0);
            for (int g = 0; g < GROUP; ++g) {
              const double a_ = cf[2 * (n0 + g)], b_ = cf[2 * (n0 + g) + 1];
              double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
              for (int w = 0; w < W; ++w) {
                if (any0) {
                  s0 = std::fma(lam2[w], er[w], s0);
                  s1 = std::fma(lam2[w], ei[w], s1);
                  s2 = std::fma(lam2[w], orr[w], s2);
                  s3 = std::fma(lam2[w], oi[w], s3);
                }
                double t = std::fma(a_, csq[w], b_);
                double nx = std::fma(t, lam2[w], lam1[w]);
                lam1[w] = lam2[w];
                lam2[w] = nx;
              }
              double* d = &acc[4 * (n0 + g)];
              d[0] += s0;
              d[1] += s1;
              d[2] += s2;
              d[3] += s3;
            }
            for (int w = 0; w < W; ++w) {
              if (sc[w] < 0 && std::fabs(lam2[w]) > SC_BIG) {
                lam1[w] *= SC_DOWN;
                lam2[w] *= SC_DOWN;
                sc[w] += 1;
                if (sc[w] == 0) {
                  er[w] = ter[w];
                  ei[w] = tei[w];
                  orr[w] = tor[w];
                  oi[w] = toi[w];
                }
              }
            }
          }
        }
        // ---- post: a_l (l=m+2n) = eps_{l+1} At_n/s_n + eps_l At_{n-1}/s_{n-1}; a_{l+1} = Bt_n/s_n
        cd* a = alm + mstart[mi];
        for (int64_t n = 0; n < nn; ++n) {
          int64_t l = m + 2 * n;
          double e1 = std::sqrt((double)((l + 1) * (l + 1) - m * m) / (double)(4 * (l + 1) * (l + 1) - 1));
          double e0 = std::sqrt((double)(l * l - m * m) / (double)(4 * l * l - 1));
          cd At(acc[4 * n], acc[4 * n + 1]), Bt(acc[4 * n + 2], acc[4 * n + 3]);
          cd v = e1 * is[n] * At;
          if (n > 0) v += e0 * is[n - 1] * cd(acc[4 * (n - 1)], acc[4 * (n - 1) + 1]);
          a[l] = v;
          if (l + 1 <= lmax) a[l + 1] = is[n] * Bt;
        }
      } else {
        cd* ea = alm + mstart[mi];
        cd* ba = alm + alm_cstride + mstart[mi];
        for (int64_t l = m; l <= lmax && l < 2; ++l) {
          ea[l] = 0;
          ba[l] = 0;
        }
        const int64_t nl = T.nstep2[mi];
        if (nl == 0) continue;
        const int64_t l0 = m > 2 ? m : 2;
        const int64_t npad = round_up(nl, CHUNK);
        const double* cf = &T.coef2[(size_t)T.off2[mi] * 2];
        const double* sg = &T.sig2[(size_t)T.off2[mi]];
        acc.assign((size_t)npad * 4, 0.0);  // G+ (re,im), G- (re,im) per l
        const int64_t pw = m >= 2 ? m - 2 : 2 - m;
        const int q = (int)std::min<int64_t>(m, 2);
        const bool odd0 = ((l0 + m) & 1) != 0;
        for (int64_t p0 = 0; p0 < np; p0 += W) {
          double x[W], lp1[W], lp2[W], lm1[W], lm2[W];
          // inputs: w+ = Q + iU, w- = Q - iU at N and S (complex each)
          cd twpN[W], twmN[W], twpS[W], twmS[W];
          cd wpN[W], wmN[W], wpS[W], wmS[W];  // masked by the scale state of the lam they multiply
          int scp[W], scm[W];
          bool any_act = false;
          for (int w = 0; w < W; ++w) {
            int64_t p = p0 + w;
            bool act = p < np && m <= pairs[p].mlim2;
            any_act |= act;
            const Pair& pr = pairs[std::min(p, np - 1)];
            x[w] = pr.cth;
            double mant;
            int64_t e;
            pow_scaled(pr.sth, pw, mant, e);
            double fs = 1.0, fc = 1.0;
            for (int k = 0; k < q; ++k) {
              fs *= pr.sh * pr.sh;
              fc *= pr.ch * pr.ch;
            }
            to_scaled(T.kp2[mi] * mant * fs, e, lp2[w], scp[w]);
            to_scaled(T.km2[mi] * mant * fc, e, lm2[w], scm[w]);
            lp1[w] = lm1[w] = 0.0;
            cd qN = act ? leg[(0 * nring + pr.irN) * nm + mi] : cd(0, 0);
            cd uN = act ? leg[(1 * nring + pr.irN) * nm + mi] : cd(0, 0);
            cd qS = (act && pr.irS >= 0) ? leg[(0 * nring + pr.irS) * nm + mi] : cd(0, 0);
            cd uS = (act && pr.irS >= 0) ? leg[(1 * nring + pr.irS) * nm + mi] : cd(0, 0);
            const cd I(0, 1);
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
          if (!any_act) continue;
          for (int64_t n0 = 0; n0 < npad; n0 += GROUP) {
            bool any0 = false;
            for (int w = 0; w < W; ++w) any0 |= (scp[w] == 0) | (scm[w] == 0);
            for (int g = 0; g < GROUP; ++g) {
              const int64_t j = n0 + g;
              const double a_ = cf[2 * j], b_ = cf[2 * j + 1];
              const bool odd = odd0 ^ ((j & 1) != 0);
              const double sgn = odd ? -1.0 : 1.0;
              double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
              for (int w = 0; w < W; ++w) {
                if (any0) {
                  s0 = std::fma(lp2[w], wpN[w].real(), s0);
                  s1 = std::fma(lp2[w], wpN[w].imag(), s1);
                  s0 = std::fma(sgn * lm2[w], wpS[w].real(), s0);
                  s1 = std::fma(sgn * lm2[w], wpS[w].imag(), s1);
                  s2 = std::fma(lm2[w], wmN[w].real(), s2);
                  s3 = std::fma(lm2[w], wmN[w].imag(), s3);
                  s2 = std::fma(sgn * lp2[w], wmS[w].real(), s2);
                  s3 = std::fma(sgn * lp2[w], wmS[w].imag(), s3);
                }
                double tp = std::fma(x[w], a_, b_);
                double tm = std::fma(x[w], a_, -b_);
                double np_ = std::fma(tp, lp2[w], -lp1[w]);
                double nm_ = std::fma(tm, lm2[w], -lm1[w]);
                lp1[w] = lp2[w];
                lp2[w] = np_;
                lm1[w] = lm2[w];
                lm2[w] = nm_;
              }
              double* d = &acc[4 * j];
              d[0] += s0;
              d[1] += s1;
              d[2] += s2;
              d[3] += s3;
            }
            for (int w = 0; w < W; ++w) {
              if (scp[w] < 0 && std::fabs(lp2[w]) > SC_BIG) {
                lp1[w] *= SC_DOWN;
                lp2[w] *= SC_DOWN;
                scp[w] += 1;
                if (scp[w] == 0) {
                  wpN[w] = twpN[w];
                  wmS[w] = twmS[w];
                }
              }
              if (scm[w] < 0 && std::fabs(lm2[w]) > SC_BIG) {
                lm1[w] *= SC_DOWN;
                lm2[w] *= SC_DOWN;
                scm[w] += 1;
                if (scm[w] == 0) {
                  wmN[w] = twmN[w];
                  wpS[w] = twpS[w];
                }
              }
            }
          }
        }
        // ---- post: E = -sigma/2 (G+ + G-), B = i sigma/2 (G+ - G-)
        for (int64_t j = 0; j < nl; ++j) {
          int64_t l = l0 + j;
          cd Gp(acc[4 * j], acc[4 * j + 1]), Gm(acc[4 * j + 2], acc[4 * j + 3]);
          ea[l] = -0.5 * sg[j] * (Gp + Gm);
          ba[l] = cd(0, 0.5 * sg[j]) * (Gp - Gm);
        }
      }
    }
  }
}

}  // extern "C"
