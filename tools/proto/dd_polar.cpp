// Prototype for the next round: how much precision does the spin-0 recurrence need next to the poles?
// Runs the kernels' two-term-in-cos^2 recurrence  p_{n+1} = (a_n x^2 + b_n) p_n + p_{n-1}  (csrc/legendre_tables.hpp)
// for HEALPix rings in (1) plain double (what the kernels do), (2) double state with double-double
// coefficients, (3) double-double state and coefficients, (4) plain double except that x^2 = cos^2(theta) is
// carried as a double-double (one extra FMA per step), and compares each with a __float128 run of the same
// recurrence.
// Result (nside 4096, lmax 8191): the error of (1) is dominated by the ONE rounding of x^2 -- a constant
// bias of the rotation angle that drifts the phase linearly over the ~4000 steps; (4) removes it.  Build:  g++ -O2 -std=c++17 tools/proto/dd_polar.cpp -lquadmath -o /tmp/dd_polar
#include <quadmath.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef __float128 q;
struct dd {
  double hi, lo;
};
static inline dd two_sum(double a, double b) {
  double s = a + b, bb = s - a;
  return {s, (a - (s - bb)) + (b - bb)};
}
static inline dd two_prod(double a, double b) {
  double p = a * b;
  return {p, std::fma(a, b, -p)};
}
static inline dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  double e = s.lo + a.lo + b.lo;
  return two_sum(s.hi, e);
}
static inline dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  double e = p.lo + a.hi * b.lo + a.lo * b.hi;
  return two_sum(p.hi, e);
}
static inline dd to_dd(q v) {
  double h = (double)v;
  return {h, (double)(v - (q)h)};
}
static inline q eps2(int64_t l, int64_t m) { return ((q)l * l - (q)m * m) / (4 * (q)l * l - 1); }

int main() {
  const int64_t nside = 4096, lmax = 8191;
  const int rings[] = {1, 2, 4, 16, 64, 256, 1024, 4000};
  const int ms[] = {0, 1, 2, 10};
  printf("ring  m   double      dd-coef     full dd     dd x^2 only   (max_n |p_n - p_n^quad| / max_n |p_n^quad|)\n");
  for (int ring : rings) {
    const q t = (q)ring * ring / (3 * (q)nside * nside);
    const q cth = 1 - t, x2q = cth * cth;
    for (int m : ms) {
      const int64_t nn = (lmax - m) / 2 + 1;
      std::vector<q> a(nn), b(nn);
      q rho_prev = 1;
      for (int64_t n = 0; n < nn; ++n) {  // same construction as build_legendre_tables, in quad
        const int64_t L = m + 2 * n + 1;
        q e1 = eps2(L + 1, m), e2 = eps2(L + 2, m), e0 = eps2(L, m);
        q ee = sqrtq(e1 * e2);
        q rho = n == 0 ? (q)1 : -ee / sqrtq(e0 * eps2(L - 1, m)) / rho_prev;
        a[n] = rho / ee;
        b[n] = -a[n] * (e1 + e0);
        rho_prev = rho;
      }
      // the scale of p_0 is irrelevant for relative errors
      q pq1 = 0, pq2 = 1;
      double pd1 = 0, pd2 = 1;   // (1) plain double
      double pc1 = 0, pc2 = 1;   // (2) dd coefficients, double state
      dd pf1 = {0, 0}, pf2 = {1, 0};  // (3) full dd
      double px1 = 0, px2 = 1;   // (4) dd x^2 only
      const double x2d = (double)x2q;
      const dd x2dd = to_dd(x2q);
      double e1m = 0, e2m = 0, e3m = 0, e4m = 0;
      q pmax = 1;
      for (int64_t n = 0; n + 1 < nn; ++n) {
        q nq = (a[n] * x2q + b[n]) * pq2 + pq1;
        pq1 = pq2, pq2 = nq;
        double td = std::fma((double)a[n], x2d, (double)b[n]);
        double nd = std::fma(td, pd2, pd1);
        pd1 = pd2, pd2 = nd;
        dd tc = dd_add(dd_mul(to_dd(a[n]), x2dd), to_dd(b[n]));
        double nc = std::fma(tc.hi, pc2, std::fma(tc.lo, pc2, pc1));
        pc1 = pc2, pc2 = nc;
        dd nf = dd_add(dd_mul(tc, pf2), pf1);
        pf1 = pf2, pf2 = nf;
        double tx = std::fma((double)a[n], x2dd.lo, std::fma((double)a[n], x2dd.hi, (double)b[n]));
        double nx = std::fma(tx, px2, px1);
        px1 = px2, px2 = nx;
        q aq = fabsq(pq2);
        if (aq > pmax) pmax = aq;
        e1m = std::fmax(e1m, (double)fabsq((q)pd2 - pq2));
        e2m = std::fmax(e2m, (double)fabsq((q)pc2 - pq2));
        e3m = std::fmax(e3m, (double)fabsq(((q)pf2.hi + (q)pf2.lo) - pq2));
        e4m = std::fmax(e4m, (double)fabsq((q)px2 - pq2));
      }
      printf("%4d %3d   %.2e    %.2e    %.2e    %.2e\n", ring, m, e1m / (double)pmax, e2m / (double)pmax,
             e3m / (double)pmax, e4m / (double)pmax);
    }
  }
  return 0;
}
