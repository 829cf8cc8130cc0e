// equator_state.cpp -- prototype behind profiles/r02_parity.md ("the spin-0 equator rings").
// (1) The equatorial difference form of legendre_tables.hpp on the equator ring with its state in double, with a
//     compensated first difference and in __float128 (same double coefficient table): they agree to 1e-15 over 4000
//     steps, so the 2.9e-13 of those rings is not the recurrence.
// (2) lambda_l = eps_{l+1} q_n + eps_l q_{n-1} for single l, with the pieces in double or __float128: each term is n
//     times lambda_l, and the error is set by eps_l / the products / the table entry 1/g_n, not by the state.
//   g++ -O2 -std=gnu++17 -fopenmp -ffp-contract=off -Ihealpixmpi.jl_b200/csrc tools/proto/equator_state.cpp -lquadmath -o /tmp/eq && /tmp/eq 0
#include <cstdio>
#include <cmath>
#include <vector>
#include <quadmath.h>
#include "legendre_tables.hpp"
using namespace hpsht;
typedef __float128 q;
int main(int argc, char** argv) {
  int64_t lmax = 8191, m = argc > 1 ? atoll(argv[1]) : 0;
  LegendreTables T;
  build_legendre_tables(lmax, &m, 1, 32, true, false, T);
  int64_t nn = T.nstep0[0];
  const double* cf = &T.coef0[LC_DE][(size_t)T.coff0[LC_DE][0] * 2];
  const double* is = &T.invs0[0];
  // exact lambda_{l,m}(pi/2) for l = m + 2n via quad recurrence (x = 0): lam_{l+2} = -(eps_{l+1}/eps_{l+2}) lam_l
  std::vector<q> lam(nn);
  {
    q p = (q)(2 * m + 1) / (4 * M_PIq);
    for (int64_t k = 1; k <= m; ++k) p *= (q)(2 * k - 1) / (q)(2 * k);
    q cur = sqrtq(p); if (m & 1) cur = -cur;
    for (int64_t n = 0; n < nn; ++n) {
      lam[n] = cur;
      int64_t l = m + 2 * n;
      q fm = m;
      auto eps = [&](int64_t L) { q fl = L; return sqrtq((fl * fl - fm * fm) / (4 * fl * fl - 1)); };
      cur = -(eps(l + 1) / eps(l + 2)) * cur;
    }
  }

  // lambda_l from the two-term form, pieces in double or quad:  lam_l = eps_{l+1} q_n + eps_l q_{n-1},
  // q_n = (-1)^n u_n / g_n
  std::vector<double> ud(nn); std::vector<q> uq(nn);
  {
    double u = T.k0[0], v = u; q U = T.k0[0], V = U;
    for (int64_t n = 0; n < nn; ++n) {
      ud[n] = u; uq[n] = U;
      double d = cf[2 * n + 1];
      v = std::fma(d, u, v); u = u + v;
      V = (q)d * U + V; U = U + V;
    }
  }
  auto epsq = [&](int64_t L) { q fl = L, fm = m; return sqrtq((fl * fl - fm * fm) / (4 * fl * fl - 1)); };
  auto epsd = [&](int64_t L) { return std::sqrt((double)(L * L - m * m) / (double)(4 * L * L - 1)); };
  for (int64_t n : {500, 1000, 2000, 3000, 3500, 4000}) {
    if (n >= nn) continue;
    int64_t l = m + 2 * n;
    q sg = (n & 1) ? -1 : 1;
    // all double (as the port)
    double t1 = epsd(l + 1) * is[n] * ((n & 1) ? -ud[n] : ud[n]);
    double t2 = epsd(l) * is[n - 1] * (((n - 1) & 1) ? -ud[n - 1] : ud[n - 1]);
    double all_double = t1 + t2;
    // quad eps, double rest
    q a1 = epsq(l + 1) * (q)is[n] * sg * (q)ud[n] + epsq(l) * (q)is[n - 1] * (-sg) * (q)ud[n - 1];
    // quad state too
    q a2 = epsq(l + 1) * (q)is[n] * sg * uq[n] + epsq(l) * (q)is[n - 1] * (-sg) * uq[n - 1];
    printf("n %ld l %ld: term/lam %.1f | all double %.2e | quad eps+arith %.2e | + quad state %.2e\n", (long)n, (long)l,
           (double)(t1 / lam[n]), (double)(((q)all_double - lam[n]) / lam[n]), (double)((a1 - lam[n]) / lam[n]),
           (double)((a2 - lam[n]) / lam[n]));
  }
}
