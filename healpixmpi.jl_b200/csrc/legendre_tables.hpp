// legendre_tables.hpp — host-side (long double) generation of the per-m recurrence tables used by
// the sm_100a Legendre kernels (alm2leg / leg2alm; replaces the arithmetic Ducc0.Sht.alm2leg! /
// leg2alm! perform behind reference src/sht.jl:85,178).
//
// Formulation (published algorithms, re-derived here; no Plm table is ever stored):
//
//  spin 0  — two-term-in-cos^2 recurrence (Ishioka 2018).  With x = cos(theta),
//            eps_l = sqrt((l^2-m^2)/(4l^2-1)) and  x lam_l = eps_{l+1} lam_{l+1} + eps_l lam_{l-1},
//            the sequence q_n = lam_{m+2n+1}/x obeys
//              eps_{L+1} eps_{L+2} q_{n+1} = (x^2 - eps_{L+1}^2 - eps_L^2) q_n - eps_L eps_{L-1} q_{n-1},
//            L = m+2n+1.  Rescaling p_n = s_n q_n with s_{n+1} = -s_{n-1} eps_{L+1}eps_{L+2}/(eps_L eps_{L-1})
//            turns it into            p_{n+1} = (a_n x^2 + b_n) p_n + p_{n-1}        (2 FMA)
//            and the synthesis sum splits into an even and an odd part in x:
//              sum_l a_lm lam_l = sum_n p_n A_n + x sum_n p_n B_n,
//              A_n = (eps_{l+1} a_l + eps_{l+2} a_{l+2}) / s_n,  B_n = a_{l+1} / s_n,   l = m+2n.
//            North ring = even + odd, south ring = even - odd.
//
//  spin 2  — three-term recurrence of the spin-weighted functions (Goldberg convention)
//              C_{l+1} lam_{l+1} = (x + s m/(l(l+1))) lam_l - C_l lam_{l-1},
//              C_l = sqrt((l^2-m^2)(l^2-4)/(l^2(4l^2-1))),
//            rescaled with lam_l = sigma_l delta_l, sigma_{l+1} = sigma_{l-1} C_l/C_{l+1}, so that
//              delta^{+-}_{l+1} = (a_l x +- b_l) delta^{+-}_l - delta^{+-}_{l-1}          (2 FMA each)
//            a_l = sigma_l/(sigma_{l+1} C_{l+1}),  b_l = a_l * 2m/(l(l+1)).
//
// All tables are padded per m to a multiple of `chunk` steps with a = b = 0 (the recurrence then
// just swaps p_{n+1} = p_{n-1}; padded alm-stream entries are 0 so nothing is accumulated).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace hpsht {

typedef long double ldbl;
static const ldbl LT_PI = 3.14159265358979323846264338327950288419716939937510L;

struct LegendreTables {
  int64_t lmax = 0;
  int64_t chunk = 0;             // padding granularity in steps
  std::vector<int64_t> mval;     // local m values
  // ---- spin 0 ----
  std::vector<int64_t> nstep0;   // per m: number of real steps nn = (lmax-m)/2+1
  std::vector<int64_t> off0;     // per m: offset (in steps) into coef0/invs0; length nm+1
  std::vector<double> coef0;     // 2 doubles per step: a_n, b_n
  std::vector<double> invs0;     // 1/s_n per step
  std::vector<double> k0;        // per m: p_0 = k0 * sin(theta)^m
  // ---- spin 2 ----
  std::vector<int64_t> nstep2;   // per m: nl = lmax - max(m,2) + 1 (0 if negative)
  std::vector<int64_t> off2;     // per m offsets, length nm+1
  std::vector<double> coef2;     // 2 doubles per step: a_l, b_l (step l -> l+1)
  std::vector<double> sig2;      // sigma_l per step
  std::vector<double> kp2, km2;  // per m: delta^+_{l0} = kp2 * sth^pw * sin(th/2)^(2q),
                                 //        delta^-_{l0} = km2 * sth^pw * cos(th/2)^(2q)
                                 //        pw = |m-2|, q = min(m,2)
};

static inline int64_t round_up(int64_t v, int64_t g) { return (v + g - 1) / g * g; }

static inline ldbl eps2(int64_t l, int64_t m) {  // eps_l^2
  ldbl fl = (ldbl)l, fm = (ldbl)m;
  return (fl * fl - fm * fm) / (4.0L * fl * fl - 1.0L);
}
static inline ldbl c2sq(int64_t l, int64_t m) {  // C_l^2 for |s| = 2
  ldbl fl = (ldbl)l, fm = (ldbl)m;
  return (fl * fl - fm * fm) * (fl * fl - 4.0L) / (fl * fl * (4.0L * fl * fl - 1.0L));
}

// sqrt(prod_{k=1..m} (2k-1)/(2k)) ~ (pi m)^{-1/4}; safe in long double for any m of interest.
static inline ldbl sqrt_central_ratio(int64_t m) {
  ldbl p = 1.0L;
  for (int64_t k = 1; k <= m; ++k) p *= sqrtl((ldbl)(2 * k - 1) / (ldbl)(2 * k));
  return p;
}

inline void build_legendre_tables(int64_t lmax, const int64_t* mval, int64_t nm, int64_t chunk,
                                  bool want_spin0, bool want_spin2, LegendreTables& T) {
  T.lmax = lmax;
  T.chunk = chunk;
  T.mval.assign(mval, mval + nm);
  T.nstep0.assign(nm, 0);
  T.nstep2.assign(nm, 0);
  T.off0.assign(nm + 1, 0);
  T.off2.assign(nm + 1, 0);
  for (int64_t i = 0; i < nm; ++i) {
    int64_t m = mval[i];
    int64_t nn = (m <= lmax) ? (lmax - m) / 2 + 1 : 0;
    int64_t l0 = m > 2 ? m : 2;
    int64_t nl = (l0 <= lmax) ? lmax - l0 + 1 : 0;
    T.nstep0[i] = nn;
    T.nstep2[i] = nl;
    T.off0[i + 1] = T.off0[i] + (want_spin0 ? round_up(nn, chunk) : 0);
    T.off2[i + 1] = T.off2[i] + (want_spin2 ? round_up(nl, chunk) : 0);
  }
  if (want_spin0) {
    T.coef0.assign((size_t)T.off0[nm] * 2, 0.0);
    T.invs0.assign((size_t)T.off0[nm], 0.0);
    T.k0.assign(nm, 0.0);
  }
  if (want_spin2) {
    T.coef2.assign((size_t)T.off2[nm] * 2, 0.0);
    T.sig2.assign((size_t)T.off2[nm], 0.0);
    T.kp2.assign(nm, 0.0);
    T.km2.assign(nm, 0.0);
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t i = 0; i < nm; ++i) {
    const int64_t m = mval[i];
    if (want_spin0 && T.nstep0[i] > 0) {
      double* cf = &T.coef0[(size_t)T.off0[i] * 2];
      double* is = &T.invs0[(size_t)T.off0[i]];
      ldbl s = 1.0L, rho_prev = 1.0L;
      for (int64_t n = 0; n < T.nstep0[i]; ++n) {
        int64_t L = m + 2 * n + 1;
        ldbl e1 = eps2(L + 1, m), e2 = eps2(L + 2, m), e0 = eps2(L, m);
        ldbl ee = sqrtl(e1 * e2);
        ldbl rho;
        if (n == 0) {
          rho = 1.0L;
        } else {
          ldbl em1 = eps2(L - 1, m);
          rho = -ee / sqrtl(e0 * em1) / rho_prev;
        }
        ldbl a = rho / ee;
        ldbl b = -a * (e1 + e0);
        cf[2 * n] = (double)a;
        cf[2 * n + 1] = (double)b;
        is[n] = (double)(1.0L / s);
        s *= rho;
        rho_prev = rho;
      }
      // p_0 = q_0 = lam_mm / eps_{m+1} = lam_mm sqrt(2m+3)
      ldbl k = sqrtl((ldbl)(2 * m + 1) * (ldbl)(2 * m + 3) / (4.0L * LT_PI)) * sqrt_central_ratio(m);
      T.k0[i] = (double)((m & 1) ? -k : k);
    }
    if (want_spin2 && T.nstep2[i] > 0) {
      double* cf = &T.coef2[(size_t)T.off2[i] * 2];
      double* sg = &T.sig2[(size_t)T.off2[i]];
      const int64_t l0 = m > 2 ? m : 2;
      // sigma_{l0} = sigma_{l0+1} = 1 ; sigma_{l+1} = sigma_{l-1} C_l / C_{l+1}
      ldbl sig_prev = 1.0L;  // sigma_{l-1}
      ldbl sig_cur = 1.0L;   // sigma_l
      for (int64_t j = 0; j < T.nstep2[i]; ++j) {
        int64_t l = l0 + j;
        ldbl sig_next;
        if (j == 0)
          sig_next = 1.0L;
        else
          sig_next = sig_prev * sqrtl(c2sq(l, m) / c2sq(l + 1, m));
        ldbl a = sig_cur / (sig_next * sqrtl(c2sq(l + 1, m)));
        ldbl b = a * 2.0L * (ldbl)m / ((ldbl)l * (ldbl)(l + 1));
        cf[2 * j] = (double)a;
        cf[2 * j + 1] = (double)b;
        sg[j] = (double)sig_cur;
        sig_prev = sig_cur;
        sig_cur = sig_next;
      }
      if (m >= 2) {
        // (-1)^m sqrt((2m+1)/(4pi) (2m)!/((m+2)!(m-2)!)) / 2^(m-2)
        //  (2m)!/((m+2)!(m-2)!) = C(2m,m) m(m-1)/((m+1)(m+2)),  C(2m,m)/4^m = prod (2k-1)/(2k)
        ldbl r = sqrt_central_ratio(m);  // sqrt(C(2m,m)/4^m)
        ldbl k = sqrtl((ldbl)(2 * m + 1) / (4.0L * LT_PI) * (ldbl)m * (ldbl)(m - 1) /
                       ((ldbl)(m + 1) * (ldbl)(m + 2))) * r * 4.0L;  // 2^m / 2^(m-2) = 4
        if (m & 1) k = -k;
        T.kp2[i] = (double)k;
        T.km2[i] = (double)k;
      } else if (m == 1) {
        ldbl pref = sqrtl(6.0L * 1.0L * 5.0L / (4.0L * LT_PI * 24.0L));
        T.kp2[i] = (double)(-2.0L * pref);  // -pref*4*ch*sh^3 = -2 pref sth sh^2
        T.km2[i] = (double)(2.0L * pref);   //  pref*4*ch^3*sh =  2 pref sth ch^2
      } else {
        ldbl pref = sqrtl(2.0L * 2.0L * 5.0L / (4.0L * LT_PI * 24.0L));
        T.kp2[i] = (double)(1.5L * pref);  // pref*6*ch^2 sh^2 = 1.5 pref sth^2
        T.km2[i] = (double)(1.5L * pref);
      }
    }
  }
}

// The standard libsharp/ducc ring-culling rule: (m, ring) pairs with m > mlim are
// skipped (their leg is exactly 0).  Published formula; spin = 0 or 2.
static inline int64_t mlim_rule(int64_t lmax, int spin, double sth, double cth) {
  double ofs = lmax * 0.01;
  if (ofs < 100.) ofs = 100.;
  double b = -2. * spin * std::fabs(cth);
  double t1 = lmax * sth + ofs;
  double c = (double)spin * spin - t1 * t1;
  double discr = b * b - 4. * c;
  if (discr <= 0) return lmax;
  double res = (-b + std::sqrt(discr)) / 2.;
  if (res > lmax) res = (double)lmax;
  return (int64_t)(res + 0.5);
}

}  // namespace hpsht
