// equator_state.cpp -- prototype behind profiles/r02_parity.md ("the spin-0 equator rings"): the equatorial difference
// form of legendre_tables.hpp run on the equator ring with its state in double, with a compensated first difference,
// and in __float128 (same double coefficient table): they agree to 1e-15 over 4000 steps, so the 2.9e-13 of those
// rings is not the recurrence.  Also prints |v/u| and the step coefficient d.
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

  double y = 0.0;
  q u = T.k0[0], v = u;
  double ud = T.k0[0], vd = ud;
  double uc = T.k0[0], vc = uc, vcl = 0;   // compensated v
  double u3 = T.k0[0], u3p = 0;            // three-term form  u_{n+1} = (2 + d) u_n - u_{n-1}
  for (int64_t n = 0; n < nn; ++n) {
    if (n == 1 || n == 10 || n == 100 || n == 1000 || n == 2000 || n == 3000 || n == 4000)
      printf("n %ld  double-vs-quad state %.2e   compensated-v %.2e   |v/u| %.2e  d %.3e\n", (long)n, (double)(((q)ud - u) / u),
             (double)(((q)uc - u) / u), (double)fabsq(v / u), cf[2*n+1]);
    double d = std::fma(cf[2 * n], y, cf[2 * n + 1]);
    vd = std::fma(d, ud, vd); ud = ud + vd;
    { double p = d * uc, ep = std::fma(d, uc, -p), s2 = vc + p, bb = s2 - vc, es = (vc - (s2 - bb)) + (p - bb); vcl += ep + es; vc = s2; uc = uc + (vc + vcl); }
    v = (q)d * u + v; u = u + v;
  }
}
