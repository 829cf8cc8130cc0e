// Prototype (round 2): which FP64 formulation of the spin-0 / spin-2 recurrences is accurate next to the
// poles and next to the equator?  Compares against a __float128 run of the same recurrence at the exact
// double theta.
//   spin 0:  u_{n+1} = tau_n u_n - u_{n-1},  tau_n = A_n x^2 + B_n   (the kernels' recurrence without the sign
//            alternation; A_n = (-1)^n a_n, B_n = (-1)^n b_n of csrc/legendre_tables.hpp)
//     (1) tau = fma(A, x^2, B)                       -- round 1
//     (2) tau = fma(-A, s^2, A + B)  for theta < 45 deg (s^2 = sin^2 theta), (1) otherwise
//     (3) difference form: d = G - A y ; v += d u ; u += v   with  polar  G = A + B - 2, y = s^2
//                                                            and   equator G = -(B + 2), y = x^2, u -> (-1)^n u
//   spin 2:  D_{l+1} = (a_l x +- b_l) D_l - D_{l-1}
//     (1) t = fma(x, a, +-b)    (2) t = fma(-a, y, a +- b), y = 1 - x = 2 sin^2(theta/2)   (3) difference form
// Build:  g++ -O2 -std=c++17 tools/proto/diff_form.cpp -lquadmath -o /tmp/diff_form
#include <quadmath.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

typedef __float128 q;
static inline q eps2(int64_t l, int64_t m) { return ((q)l * l - (q)m * m) / (4 * (q)l * l - 1); }
static inline q c2sq(int64_t l, int64_t m) {
  q fl = l, fm = m;
  return (fl * fl - fm * fm) * (fl * fl - 4) / (fl * fl * (4 * fl * fl - 1));
}

int main(int argc, char** argv) {
  const int64_t nside = argc > 1 ? atoll(argv[1]) : 4096, lmax = argc > 2 ? atoll(argv[2]) : 8191;
  std::vector<int64_t> rings = {1, 2, 4, 16, 64, 256, 1024, 2048, 3000, 4000, 6000, 8000, 8100, 8176, 8188, 8190, 8191, 8192};
  const int ms[] = {0, 1, 2, 10, 100};
  printf("spin 0:  ring  m   std-x2      y-switch    diff-form\n");
  for (int64_t ring : rings) {
    if (ring > 2 * nside) continue;
    // HEALPix colatitude as the caller's double
    q z;
    if (ring < nside) z = 1 - (q)ring * ring / (3 * (q)nside * nside);
    else z = (q)(2 * nside - ring) * 2 / (3 * (q)nside);
    const double theta = (double)acosq(z);
    const q cth = cosq((q)theta), sth = sinq((q)theta);
    const double x2 = (double)(cth * cth), s2 = (double)(sth * sth);
    const bool polar_side = theta < 0.78539816339744830962;
    for (int m : ms) {
      const int64_t nn = (lmax - m) / 2 + 1;
      std::vector<q> A(nn), B(nn);
      q rho_prev = 1;
      for (int64_t n = 0; n < nn; ++n) {
        const int64_t L = m + 2 * n + 1;
        q e1 = eps2(L + 1, m), e2 = eps2(L + 2, m), e0 = eps2(L, m);
        q ee = sqrtq(e1 * e2);
        q rho = n == 0 ? (q)1 : -ee / sqrtq(e0 * eps2(L - 1, m)) / rho_prev;
        q a = rho / ee, b = -a * (e1 + e0);
        A[n] = (n & 1) ? -a : a;
        B[n] = (n & 1) ? -b : b;
        rho_prev = rho;
      }
      q uq1 = 0, uq2 = 1;
      double a1 = 0, a2 = 1;  // (1)
      double b1 = 0, b2 = 1;  // (2)
      double cu = 1, cv = 1;  // (3): u_0 = 1, v_0 = u_0 - u_{-1} = 1   (equator: w_n = (-1)^n u_n)
      double e1m = 0, e2m = 0, e3m = 0;
      q umax = 1;
      for (int64_t n = 0; n + 1 < nn; ++n) {
        q tq = A[n] * cth * cth + B[n];
        q nq = tq * uq2 - uq1;
        uq1 = uq2, uq2 = nq;
        double t1 = std::fma((double)A[n], x2, (double)B[n]);
        double n1 = std::fma(t1, a2, -a1);
        a1 = a2, a2 = n1;
        double t2 = polar_side ? std::fma(-(double)A[n], s2, (double)(A[n] + B[n])) : t1;
        double n2 = std::fma(t2, b2, -b1);
        b1 = b2, b2 = n2;
        double d = polar_side ? std::fma(-(double)A[n], s2, (double)(A[n] + B[n] - 2))
                              : std::fma(-(double)A[n], x2, (double)(-(B[n] + 2)));
        cv = std::fma(d, cu, cv);
        cu = cu + cv;
        const double u3 = polar_side ? cu : (((n + 1) & 1) ? -cu : cu);
        q aq = fabsq(uq2);
        if (aq > umax) umax = aq;
        e1m = std::fmax(e1m, (double)fabsq((q)a2 - uq2));
        e2m = std::fmax(e2m, (double)fabsq((q)b2 - uq2));
        e3m = std::fmax(e3m, (double)fabsq((q)u3 - uq2));
      }
      printf("       %5ld %3d   %.2e    %.2e    %.2e\n", (long)ring, m, e1m / (double)umax, e2m / (double)umax,
             e3m / (double)umax);
    }
  }
  printf("spin 2 (D+ recurrence):  ring  m   std-x       y-form      diff-form\n");
  for (int64_t ring : rings) {
    if (ring > 2 * nside) continue;
    q z;
    if (ring < nside) z = 1 - (q)ring * ring / (3 * (q)nside * nside);
    else z = (q)(2 * nside - ring) * 2 / (3 * (q)nside);
    const double theta = (double)acosq(z);
    const q cth = cosq((q)theta), shq = sinq((q)theta / 2);
    const double x = (double)cth, y = (double)(2 * shq * shq);
    for (int m : ms) {
      const int64_t l0 = m > 2 ? m : 2;
      const int64_t nl = lmax - l0 + 1;
      std::vector<q> a(nl), b(nl);
      q sp = 1, sc = 1;
      for (int64_t j = 0; j < nl; ++j) {
        int64_t l = l0 + j;
        q sn = j == 0 ? (q)1 : sp * sqrtq(c2sq(l, m) / c2sq(l + 1, m));
        a[j] = sc / (sn * sqrtq(c2sq(l + 1, m)));
        b[j] = a[j] * 2 * (q)m / ((q)l * (l + 1));
        sp = sc, sc = sn;
      }
      q uq1 = 0, uq2 = 1;
      double a1 = 0, a2 = 1, b1 = 0, b2 = 1, cu = 1, cv = 1;
      double e1m = 0, e2m = 0, e3m = 0;
      q umax = 1;
      for (int64_t j = 0; j + 1 < nl; ++j) {
        q nq = (a[j] * cth + b[j]) * uq2 - uq1;
        uq1 = uq2, uq2 = nq;
        double t1 = std::fma(x, (double)a[j], (double)b[j]);
        double n1 = std::fma(t1, a2, -a1);
        a1 = a2, a2 = n1;
        double t2 = std::fma(-(double)a[j], y, (double)(a[j] + b[j]));
        double n2 = std::fma(t2, b2, -b1);
        b1 = b2, b2 = n2;
        double d = std::fma(-(double)a[j], y, (double)(a[j] + b[j] - 2));
        cv = std::fma(d, cu, cv);
        cu = cu + cv;
        q aq = fabsq(uq2);
        if (aq > umax) umax = aq;
        e1m = std::fmax(e1m, (double)fabsq((q)a2 - uq2));
        e2m = std::fmax(e2m, (double)fabsq((q)b2 - uq2));
        e3m = std::fmax(e3m, (double)fabsq((q)cu - uq2));
      }
      printf("       %5ld %3d   %.2e    %.2e    %.2e\n", (long)ring, m, e1m / (double)umax, e2m / (double)umax,
             e3m / (double)umax);
    }
  }
  return 0;
}
