// rr.hpp — integer-exact round-robin (RR) bookkeeping and HEALPix ring geometry (host only).
//
// Restates what the reference computes in src/alm.jl:102-151 (get_nm_RR, get_mval_RR,
// get_m_tasks_RR, make_mstart_complex) and src/map.jl:99-143 (get_nrings_RR, get_rindexes_RR,
// get_rindexes_tot_RR) plus the geometry ScatterMap! reads from Healpix.jl (src/map.jl:177-183).
// Written from the formulas, not translated line by line.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace hpsht {

// number of m owned by `rank` out of `size` for m = 0..mmax               (src/alm.jl:102-105)
static inline int64_t rr_nm(int64_t mmax, int64_t rank, int64_t size) {
  return (mmax + size - rank) / size;
}
// m values of `rank`: rank, rank+size, ...                                 (src/alm.jl:112-119)
static inline std::vector<int64_t> rr_mval(int64_t mmax, int64_t rank, int64_t size) {
  std::vector<int64_t> v((size_t)rr_nm(mmax, rank, size));
  for (size_t i = 0; i < v.size(); ++i) v[i] = rank + (int64_t)i * size;
  return v;
}
// 0-based mstart: a_{l,m} of the mi-th owned m sits at mstart[mi] + l*stride.  The running
// offset advances by lmax+1-m per m, so Julia's 1-based value is this + 1.   (src/alm.jl:141-151)
static inline std::vector<int64_t> rr_mstart0(int64_t lmax, int64_t stride,
                                              const std::vector<int64_t>& mval) {
  std::vector<int64_t> v(mval.size());
  int64_t filled = 0;
  for (size_t i = 0; i < mval.size(); ++i) {
    v[i] = stride * (filled - mval[i]);
    filled += lmax + 1 - mval[i];
  }
  return v;
}
// rings owned by `rank`; eq = 2*nside is the equator ring id; the equator belongs to rank 0,
// every other rank (and rank 0 beyond the equator) owns N/S pairs at distance rank + k*size.
static inline int64_t rr_nrings(int64_t eq, int64_t rank, int64_t size) {  // src/map.jl:99-102
  return (eq + size - 1 - rank) / size * 2 - (rank == 0 ? 1 : 0);
}
// 1-based global ring ids in local order: equator outwards, north before south (src/map.jl:113-121)
static inline std::vector<int64_t> rr_rindexes(int64_t eq, int64_t rank, int64_t size) {
  std::vector<int64_t> v;
  v.reserve((size_t)rr_nrings(eq, rank, size));
  for (int64_t d = rank; d < eq; d += size) {
    if (d == 0) {
      v.push_back(eq);
    } else {
      v.push_back(eq - d);
      v.push_back(eq + d);
    }
  }
  return v;
}
static inline std::vector<int64_t> rr_rindexes_tot(int64_t eq, int64_t size) {  // src/map.jl:134-143
  std::vector<int64_t> v;
  v.reserve((size_t)(2 * eq - 1));
  for (int64_t r = 0; r < size; ++r) {
    std::vector<int64_t> a = rr_rindexes(eq, r, size);
    v.insert(v.end(), a.begin(), a.end());
  }
  return v;
}

// HEALPix RING geometry of 1-based ring `ring` (1..4nside-1): standard formulas (Gorski et al.
// 2005).  theta/phi0 are evaluated in long double and rounded once.
struct RingGeom {
  int64_t nphi;
  int64_t first_pix;  // 0-based RING index of the first pixel
  double theta;
  double phi0;
};
static inline RingGeom healpix_ring(int64_t nside, int64_t ring) {
  const long double PI = 3.14159265358979323846264338327950288419716939937510L;
  const int64_t nr = 4 * nside - 1;
  const int64_t npix = 12 * nside * nside;
  const bool south = ring > 2 * nside;
  const int64_t n = south ? nr + 1 - ring : ring;
  RingGeom g;
  long double th;
  if (n < nside) {
    long double t = (long double)n * (long double)n / (3.0L * (long double)nside * (long double)nside);
    th = atan2l(sqrtl(t * (2.0L - t)), 1.0L - t);
    g.nphi = 4 * n;
    g.first_pix = 2 * n * (n - 1);
    g.phi0 = (double)(PI / (long double)(4 * n));
  } else {
    th = acosl((long double)(2 * nside - n) * 2.0L / (3.0L * (long double)nside));
    g.nphi = 4 * nside;
    g.first_pix = 2 * nside * (nside - 1) + (n - nside) * 4 * nside;
    g.phi0 = ((n - nside) % 2 == 0) ? (double)(PI / (long double)(4 * nside)) : 0.0;
  }
  if (south) {
    th = PI - th;
    g.first_pix = npix - g.first_pix - g.nphi;
  }
  g.theta = (double)th;
  return g;
}

}  // namespace hpsht
