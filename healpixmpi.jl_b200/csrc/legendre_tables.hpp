// legendre_tables.hpp — host-side (long double) generation of the per-m recurrence tables used by
// the sm_100a Legendre kernels (alm2leg / leg2alm; replaces the arithmetic Ducc0.Sht.alm2leg! /
// leg2alm! perform behind reference src/sht.jl:85,178).
//
// Formulation (published algorithms, re-derived here; no Plm table is ever stored):
//
//  spin 0  — two-term-in-cos^2 recurrence (Ishioka 2018).  With x = cos(theta),
//            eps_l = sqrt((l^2-m^2)/(4l^2-1)),  F(l) = eps_l eps_{l+1}  and  q_n = lam_{m+2n+1}/x :
//              F(L+1) q_{n+1} = (x^2 - eps_{L+1}^2 - eps_L^2) q_n - F(L-1) q_{n-1},      L = m+2n+1.
//            Rescaling p_n = g_n q_n with g_{n+1} g_n = 4 F(L+1)·(smooth branch, see below) gives
//                         p_{n+1} = (A_n x^2 + B_n) p_n - p_{n-1}                       (2 FMA)
//            and the synthesis sum splits into an even and an odd part in x:
//              sum_l a_lm lam_l = sum_n p_n Ae_n + x sum_n p_n Ao_n,
//              Ae_n = (eps_{l+1} a_l + eps_{l+2} a_{l+2}) / g_n,  Ao_n = a_{l+1} / g_n,   l = m+2n.
//            North ring = even + odd, south ring = even - odd.
//
//  spin 2  — three-term recurrence of the spin-weighted functions (Goldberg convention)
//              C_{l+1} lam_{l+1} = (x + s m/(l(l+1))) lam_l - C_l lam_{l-1},
//              C_l = sqrt((l^2-m^2)(l^2-4)/(l^2(4l^2-1))),
//            rescaled with lam_l = sigma_l delta_l, sigma_{l+1} sigma_l C_{l+1} = 1/2 (smooth branch):
//              delta^{+-}_{l+1} = (a_l x +- b_l) delta^{+-}_l - delta^{+-}_{l-1}          (2 FMA each)
//            a_l = sigma_l/(sigma_{l+1} C_{l+1}),  b_l = a_l * 2m/(l(l+1)).
//
//  Smooth branch.  Making the coefficient of the (n-1) term exactly 1 leaves one free constant per
//  parity of n (a 2-cycle g, 1/g, g, ... of the scale factors is also a solution).  Round 1 fixed it
//  by g_0 = g_1 = 1, which leaves the 2-cycle in for m > 0.  Here the chain is run BACKWARDS from the
//  top of the l-range, started on the asymptotically smooth solution (g ~ 2 sqrt(F), sigma ~
//  (4 C_l C_{l+1})^(-1/4)), so A_n -> 4, B_n -> -2 (a_l -> 2) smoothly.  That is what makes the
//  difference forms below work.
//
//  Accuracy classes (per ring pair; tools/proto/diff_form.cpp has the numbers).  The recurrence
//  rotates a phase by 2 theta (spin 0) / theta (spin 2) per step; one rounding of the per-ring
//  quantity (x^2 or x) is a constant bias of that angle which drifts the phase linearly in n, and next
//  to the poles (and, for the two-step spin-0 recurrence, next to the equator) the characteristic
//  roots coalesce and every rounding is amplified by 1/sin(2 theta) (1/sin theta).  Therefore:
//    X   t = fma( A, x^2, B)        equator side   (spin 2: t+- = fma(a,  x, +-b))
//    Y   t = fma(-A, s^2, A + B)    pole side      (spin 2: t+- = fma(a, -y, a +- b), y = 2 sin^2(theta/2))
//        -- the per-ring quantity whose rounding matters is small on the respective side;
//    DP  difference form next to the poles:    d = fma(-A, s^2, A + B - 2);  v += d u;  u += v
//    DE  difference form next to the equator (spin 0):  w_n = (-1)^n p_n,  d = fma(-A, x^2, -(B + 2))
//        -- rounding errors of u no longer feed the growing solution; measured against __float128 the
//        recurrence error drops from 4e-10 (ring 1, nside 4096) / 5e-12 (equator) to 5e-15.
//
// All tables are padded per m to a multiple of `chunk` steps with zeros (the recurrence then just
// swaps / drifts linearly for at most chunk-1 steps; padded alm-stream entries are 0 so nothing is
// accumulated).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace hpsht {

typedef long double ldbl;
static const ldbl LT_PI = 3.14159265358979323846264338327950288419716939937510L;

enum LegClass { LC_X = 0, LC_Y = 1, LC_DP = 2, LC_DE = 3, LC_COUNT = 4 };

// Pair classes as index ranges of the pole -> equator pair list:
//   [0, n_dp) DP, [n_dp, n_y) Y, [n_y, np - n_de) X, [np - n_de, np) DE  (n_de = 0 for spin 2)
struct ClassRanges {
  int np = 0, n_dp = 0, n_y = 0, n_de = 0;
  int lo(int c) const { return c == LC_DP ? 0 : c == LC_Y ? n_dp : c == LC_X ? n_y : np - n_de; }
  int hi(int c) const { return c == LC_DP ? n_dp : c == LC_Y ? n_y : c == LC_X ? np - n_de : np; }
  int cls(int p) const { return p < n_dp ? LC_DP : p < n_y ? LC_Y : p < np - n_de ? LC_X : LC_DE; }
};
// theta_north[p]: colatitude of the north ring of pair p (ascending).  The classes are chosen by angle
// (so that a sampled ring list gets the same treatment per ring as the full list): DP below 0.1 rad, DE
// (spin 0) for |cos theta| < 0.021, Y below 45 degrees (spin 0) / 60 degrees (spin 2), where the two
// standard forms' sensitivities cross.  For long lists the boundaries are moved to multiples of the
// kernels' tile sizes (the crossovers are flat, nothing is lost).
static inline ClassRanges make_class_ranges(int spin, int np, const double* theta_north) {
  ClassRanges r;
  r.np = np;
  const double lim_y = spin == 0 ? 0.25 * (double)LT_PI : (double)LT_PI / 3.0;
  const double lim_de = std::acos(0.021);
  int ndp = 0, ny = 0, nde = 0;
  for (int p = 0; p < np; ++p) {
    if (theta_north[p] < 0.1) ++ndp;
    if (theta_north[p] < lim_y) ++ny;
    if (spin == 0 && theta_north[p] > lim_de) ++nde;
  }
  // the difference forms are accurate anywhere on their side, only slower: round their ranges up
  const int a = np >= 4096 ? 512 : (np >= 512 ? 32 : 1);
  const int ade = np >= 4096 ? 256 : (np >= 512 ? 32 : 1);
  ndp = (ndp + a - 1) / a * a;
  ny = (ny + a / 2) / a * a;
  nde = (nde + ade - 1) / ade * ade;
  r.n_de = std::min(nde, np);
  r.n_dp = std::min(ndp, np - r.n_de);
  r.n_y = std::min(std::max(ny, r.n_dp), np - r.n_de);
  return r;
}

struct LegendreTables {
  int64_t lmax = 0;
  int64_t chunk = 0;             // padding granularity in steps
  std::vector<int64_t> mval;     // local m values
  // ---- spin 0 ----
  std::vector<int64_t> nstep0;   // per m: number of real steps nn = (lmax-m)/2+1
  std::vector<int64_t> off0;     // per m: offset (in steps) into invs0 / streams; length nm+1
  std::vector<int64_t> coff0[LC_COUNT];  // per m: offset into coef0[c] (rows of m's the class never needs are empty)
  std::vector<double> coef0[LC_COUNT];   // 2 doubles per step, see the class list above
  std::vector<double> invs0;     // 1/g_n per step
  std::vector<double> k0;        // per m: p_0 = k0 * sin(theta)^m
  // ---- spin 2 ----
  std::vector<int64_t> nstep2;   // per m: nl = lmax - max(m,2) + 1 (0 if negative)
  std::vector<int64_t> off2;     // per m offsets, length nm+1
  std::vector<int64_t> coff2[LC_COUNT];
  std::vector<double> coefa2;    // a_l per step (shared by all classes; offsets off2)
  std::vector<double> coef2[LC_COUNT];   // 2 doubles per step: X {b, -b}; Y {a+b, a-b}; DP {a+b-2, a-b-2}
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

// mlimX[c]: largest m any ring pair of class c keeps (-1: class empty).  Rows of larger m are left out of
// that class's tables.  nullptr = build every class for every m.
inline void build_legendre_tables(int64_t lmax, const int64_t* mval, int64_t nm, int64_t chunk,
                                  bool want_spin0, bool want_spin2, LegendreTables& T,
                                  const int64_t* mlim0 = nullptr, const int64_t* mlim2 = nullptr) {
  T.lmax = lmax;
  T.chunk = chunk;
  T.mval.assign(mval, mval + nm);
  T.nstep0.assign(nm, 0);
  T.nstep2.assign(nm, 0);
  T.off0.assign(nm + 1, 0);
  T.off2.assign(nm + 1, 0);
  for (int c = 0; c < LC_COUNT; ++c) {
    T.coff0[c].assign(nm + 1, 0);
    T.coff2[c].assign(nm + 1, 0);
  }
  for (int64_t i = 0; i < nm; ++i) {
    int64_t m = mval[i];
    int64_t nn = (m <= lmax) ? (lmax - m) / 2 + 1 : 0;
    int64_t l0 = m > 2 ? m : 2;
    int64_t nl = (l0 <= lmax) ? lmax - l0 + 1 : 0;
    T.nstep0[i] = nn;
    T.nstep2[i] = nl;
    T.off0[i + 1] = T.off0[i] + (want_spin0 ? round_up(nn, chunk) : 0);
    T.off2[i + 1] = T.off2[i] + (want_spin2 ? round_up(nl, chunk) : 0);
    for (int c = 0; c < LC_COUNT; ++c) {
      const bool n0 = want_spin0 && (!mlim0 || m <= mlim0[c]);
      const bool n2 = want_spin2 && c != LC_DE && (!mlim2 || m <= mlim2[c]);
      T.coff0[c][i + 1] = T.coff0[c][i] + (n0 ? round_up(nn, chunk) : 0);
      T.coff2[c][i + 1] = T.coff2[c][i] + (n2 ? round_up(nl, chunk) : 0);
    }
  }
  if (want_spin0) {
    for (int c = 0; c < LC_COUNT; ++c) T.coef0[c].assign((size_t)T.coff0[c][nm] * 2, 0.0);
    T.invs0.assign((size_t)T.off0[nm], 0.0);
    T.k0.assign(nm, 0.0);
  }
  if (want_spin2) {
    T.coefa2.assign((size_t)T.off2[nm], 0.0);
    for (int c = 0; c < LC_COUNT; ++c) T.coef2[c].assign((size_t)T.coff2[c][nm] * 2, 0.0);
    T.sig2.assign((size_t)T.off2[nm], 0.0);
    T.kp2.assign(nm, 0.0);
    T.km2.assign(nm, 0.0);
  }
#pragma omp parallel
  {
    std::vector<ldbl> F, g, C, sg;
#pragma omp for schedule(dynamic, 8)
    for (int64_t i = 0; i < nm; ++i) {
      const int64_t m = mval[i];
      if (want_spin0 && T.nstep0[i] > 0) {
        const int64_t nn = T.nstep0[i];
        // Fv[k] = F(m + k) = eps_{m+k} eps_{m+k+1}, k = 0 .. 2nn+5   (F(m) = 0)
        F.resize((size_t)(2 * nn + 6));
        {
          ldbl eprev = sqrtl(eps2(m, m));
          for (int64_t k = 0; k < 2 * nn + 6; ++k) {
            const ldbl enext = sqrtl(eps2(m + k + 1, m));
            F[(size_t)k] = eprev * enext;
            eprev = enext;
          }
        }
        // g_n g_{n+1} = 4 F(L_n + 1) with L_n = m+2n+1, i.e. F index 2n+2; run backwards from the smooth top
        g.resize((size_t)(nn + 2));
        g[(size_t)nn + 1] = 2.0L * sqrtl(F[(size_t)(2 * (nn + 1) + 1)]);
        g[(size_t)nn] = 2.0L * sqrtl(F[(size_t)(2 * nn + 1)]);
        for (int64_t n = nn; n >= 1; --n)  // g_{n-1} = g_{n+1} F(L_n - 1) / F(L_n + 1)
          g[(size_t)n - 1] = g[(size_t)n + 1] * F[(size_t)(2 * n)] / F[(size_t)(2 * n + 2)];
        double* is = &T.invs0[(size_t)T.off0[i]];
        double* cf[LC_COUNT];
        for (int c = 0; c < LC_COUNT; ++c)
          cf[c] = (T.coff0[c][i + 1] > T.coff0[c][i]) ? &T.coef0[c][(size_t)T.coff0[c][i] * 2] : nullptr;
        for (int64_t n = 0; n < nn; ++n) {
          const int64_t L = m + 2 * n + 1;
          const ldbl A = (g[(size_t)n + 1] / g[(size_t)n]) / F[(size_t)(2 * n + 2)];
          const ldbl e10 = eps2(L + 1, m) + eps2(L, m);
          const ldbl B = -A * e10;
          if (cf[LC_X]) cf[LC_X][2 * n] = (double)A, cf[LC_X][2 * n + 1] = (double)B;
          if (cf[LC_Y]) cf[LC_Y][2 * n] = (double)(-A), cf[LC_Y][2 * n + 1] = (double)(A + B);
          if (cf[LC_DP]) cf[LC_DP][2 * n] = (double)(-A), cf[LC_DP][2 * n + 1] = (double)(A + B - 2.0L);
          if (cf[LC_DE]) cf[LC_DE][2 * n] = (double)(-A), cf[LC_DE][2 * n + 1] = (double)(-(B + 2.0L));
          is[n] = (double)(1.0L / g[(size_t)n]);
        }
        // p_0 = g_0 q_0 = g_0 lam_mm / eps_{m+1} = g_0 lam_mm sqrt(2m+3)
        ldbl k = sqrtl((ldbl)(2 * m + 1) * (ldbl)(2 * m + 3) / (4.0L * LT_PI)) * sqrt_central_ratio(m) * g[0];
        T.k0[i] = (double)((m & 1) ? -k : k);
      }
      if (want_spin2 && T.nstep2[i] > 0) {
        const int64_t nl = T.nstep2[i];
        const int64_t l0 = m > 2 ? m : 2;
        // C[j] = C_{l0+j}, j = 0 .. nl+2   (C[0] = 0)
        C.resize((size_t)(nl + 3));
        for (int64_t j = 0; j < nl + 3; ++j) C[(size_t)j] = sqrtl(c2sq(l0 + j, m));
        sg.resize((size_t)(nl + 2));
        auto lead = [&](int64_t j) { return 1.0L / sqrtl(2.0L * sqrtl(C[(size_t)j] * C[(size_t)j + 1])); };
        sg[(size_t)nl + 1] = lead(nl + 1);
        sg[(size_t)nl] = lead(nl);
        for (int64_t j = nl; j >= 1; --j)  // sigma_{l-1} = sigma_{l+1} C_{l+1} / C_l   (C[j] > 0 for j >= 1)
          sg[(size_t)j - 1] = sg[(size_t)j + 1] * C[(size_t)j + 1] / C[(size_t)j];
        double* ca = &T.coefa2[(size_t)T.off2[i]];
        double* s = &T.sig2[(size_t)T.off2[i]];
        double* cf[LC_COUNT];
        for (int c = 0; c < LC_COUNT; ++c)
          cf[c] = (T.coff2[c][i + 1] > T.coff2[c][i]) ? &T.coef2[c][(size_t)T.coff2[c][i] * 2] : nullptr;
        for (int64_t j = 0; j < nl; ++j) {
          const int64_t l = l0 + j;
          const ldbl a = sg[(size_t)j] / (sg[(size_t)j + 1] * C[(size_t)j + 1]);
          const ldbl b = a * 2.0L * (ldbl)m / ((ldbl)l * (ldbl)(l + 1));
          ca[j] = (double)a;
          s[j] = (double)sg[(size_t)j];
          if (cf[LC_X]) cf[LC_X][2 * j] = (double)b, cf[LC_X][2 * j + 1] = (double)(-b);
          if (cf[LC_Y]) cf[LC_Y][2 * j] = (double)(a + b), cf[LC_Y][2 * j + 1] = (double)(a - b);
          if (cf[LC_DP]) cf[LC_DP][2 * j] = (double)(a + b - 2.0L), cf[LC_DP][2 * j + 1] = (double)(a - b - 2.0L);
        }
        const ldbl s0 = sg[0];  // delta_{l0} = lam_{l0} / sigma_{l0}
        if (m >= 2) {
          // (-1)^m sqrt((2m+1)/(4pi) (2m)!/((m+2)!(m-2)!)) / 2^(m-2)
          //  (2m)!/((m+2)!(m-2)!) = C(2m,m) m(m-1)/((m+1)(m+2)),  C(2m,m)/4^m = prod (2k-1)/(2k)
          ldbl r = sqrt_central_ratio(m);  // sqrt(C(2m,m)/4^m)
          ldbl k = sqrtl((ldbl)(2 * m + 1) / (4.0L * LT_PI) * (ldbl)m * (ldbl)(m - 1) /
                         ((ldbl)(m + 1) * (ldbl)(m + 2))) * r * 4.0L / s0;  // 2^m / 2^(m-2) = 4
          if (m & 1) k = -k;
          T.kp2[i] = (double)k;
          T.km2[i] = (double)k;
        } else if (m == 1) {
          ldbl pref = sqrtl(6.0L * 1.0L * 5.0L / (4.0L * LT_PI * 24.0L)) / s0;
          T.kp2[i] = (double)(-2.0L * pref);  // -pref*4*ch*sh^3 = -2 pref sth sh^2
          T.km2[i] = (double)(2.0L * pref);   //  pref*4*ch^3*sh =  2 pref sth ch^2
        } else {
          ldbl pref = sqrtl(2.0L * 2.0L * 5.0L / (4.0L * LT_PI * 24.0L)) / s0;
          T.kp2[i] = (double)(1.5L * pref);  // pref*6*ch^2 sh^2 = 1.5 pref sth^2
          T.km2[i] = (double)(1.5L * pref);
        }
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
