// shards.cuh — device-side shard bookkeeping and algebra on a plan's shards.
// Included by api.cu after the definition of hpsht_plan.
//
//   N1  scatter: global Alm / RING map on `root`  ->  every rank's DAlm / DMap shard
//       (ScatterAlm! src/alm.jl:159-186, ScatterMap! src/map.jl:151-212, MPI.Scatter! src/alm.jl:219-307,
//        src/map.jl:245-330).  The reference broadcasts the whole object to every rank and lets each rank
//       slice it; here the root cuts every rank's shard on the device and ships exactly that shard.
//   N2  gather / allgather: shards -> global object (src/alm.jl:313-595, src/map.jl:337-559).  The
//       reference runs maxnm (maxnr) rounds of tiny collectives; here one message per rank.
//   N3  algebra: dot with the (1, 2) weights of localdot (src/alm.jl:621-660), almxfl! (src/alm.jl:693-713),
//       alm2cl (src/cl.jl:5-38).
//
// Layouts: global alm = Healpix.jl Alm order, index(l, m) = m (2 lmax + 1 - m) / 2 + l (mmax = lmax),
// [ncomp][nalm_tot] complex128; global map = RING order [ncomp][12 nside^2] float64; shards as in hpsht_alm2map.
#pragma once

namespace hpsht {
namespace shards {

// ---- alm: one CTA row per local m of the target rank; copy (dir = 0: global -> shard, 1: shard -> global)
__global__ void k_alm_shard_copy(double2* __restrict__ shard, double2* __restrict__ glob,
                                 const int64_t* __restrict__ mval, const int64_t* __restrict__ mstart0, int64_t lmax,
                                 int64_t shard_cstride, int64_t glob_cstride, int ncomp, int dir) {
  const int mi = blockIdx.y;
  const int64_t m = mval[mi];
  const int64_t l = m + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const int64_t gi = m * (2 * lmax + 1 - m) / 2 + l;
  const int64_t si = mstart0[mi] + l;
  for (int c = 0; c < ncomp; ++c) {
    if (dir == 0)
      shard[c * shard_cstride + si] = glob[c * glob_cstride + gi];
    else
      glob[c * glob_cstride + gi] = shard[c * shard_cstride + si];
  }
}

// ---- map: one CTA column per local ring of the target rank
__global__ void k_map_shard_copy(double* __restrict__ shard, double* __restrict__ glob,
                                 const int64_t* __restrict__ rstart0, const int64_t* __restrict__ first_pix,
                                 const int64_t* __restrict__ nphi, int64_t shard_cstride, int64_t glob_cstride,
                                 int ncomp, int dir) {
  const int ir = blockIdx.x;
  const int64_t n = nphi[ir], s0 = rstart0[ir], g0 = first_pix[ir];
  for (int c = 0; c < ncomp; ++c)
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      if (dir == 0)
        shard[c * shard_cstride + s0 + j] = glob[c * glob_cstride + g0 + j];
      else
        glob[c * glob_cstride + g0 + j] = shard[c * shard_cstride + s0 + j];
    }
}

// ---- dot: per local m one partial (fixed summation order => deterministic), weights 1 (m = 0, real parts
//      only) and 2 (m > 0)
__global__ void k_dot_rows(const double2* __restrict__ a, const double2* __restrict__ b,
                           const int64_t* __restrict__ mval, const int64_t* __restrict__ mstart0, int64_t lmax,
                           double* __restrict__ rowsum) {
  __shared__ double red[256];
  const int mi = blockIdx.x;
  const int64_t m = mval[mi];
  const int64_t base = mstart0[mi];
  double s = 0.0;
  for (int64_t l = m + threadIdx.x; l <= lmax; l += blockDim.x) {
    const double2 x = a[base + l], y = b[base + l];
    s += (m == 0) ? x.x * y.x : 2.0 * (x.x * y.x + x.y * y.y);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) rowsum[mi] = red[0];
}
__global__ void k_sum_rows(const double* __restrict__ rowsum, int64_t n, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += rowsum[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// ---- almxfl: alm[l, m] *= fl[l] (fl shorter than lmax + 1 is zero-extended by the caller)
__global__ void k_almxfl(double2* __restrict__ alm, const double* __restrict__ fl, const int64_t* __restrict__ mval,
                         const int64_t* __restrict__ mstart0, int64_t lmax) {
  const int mi = blockIdx.y;
  const int64_t m = mval[mi];
  const int64_t l = m + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  double2 v = alm[mstart0[mi] + l];
  const double f = fl[l];
  v.x *= f;
  v.y *= f;
  alm[mstart0[mi] + l] = v;
}

// ---- alm2cl (local part): cl[l] = sum over local m <= l of w_m Re(a conj(b)) / (2l + 1).  Two stages so
//      that the sum has a fixed order: partial sums over chunks of CL_CHUNK consecutive local m's, then
//      the chunks in ascending order.
constexpr int CL_CHUNK = 64;
__global__ void k_alm2cl_part(const double2* __restrict__ a, const double2* __restrict__ b,
                              const int64_t* __restrict__ mval, const int64_t* __restrict__ mstart0, int64_t nm,
                              int64_t lmax, double* __restrict__ part) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const int64_t mi0 = (int64_t)blockIdx.y * CL_CHUNK, mi1 = min(nm, mi0 + CL_CHUNK);
  double s = 0.0;
  for (int64_t mi = mi0; mi < mi1; ++mi) {
    const int64_t m = mval[mi];
    if (m > l) break;
    const double2 x = a[mstart0[mi] + l], y = b[mstart0[mi] + l];
    const double w = m > 0 ? 2.0 : 1.0;
    s += w * (x.x * y.x + x.y * y.y) / (double)(2 * l + 1);
  }
  part[(int64_t)blockIdx.y * (lmax + 1) + l] = s;
}
__global__ void k_alm2cl_sum(const double* __restrict__ part, int64_t nchunk, int64_t lmax, double* __restrict__ cl) {
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  double s = 0.0;
  for (int64_t c = 0; c < nchunk; ++c) s += part[c * (lmax + 1) + l];
  cl[l] = s;
}

}  // namespace shards
}  // namespace hpsht

namespace {

using hpsht::DevBuf;
using namespace hpsht::shards;

void make_alm_desc(const hpsht_plan& pl, int t, RankAlmDesc& d, cudaStream_t st) {
  d.mval = hpsht::rr_mval(pl.lmax, t, pl.P);
  d.mstart0 = hpsht::rr_mstart0(pl.lmax, 1, d.mval);
  d.nalm = 0;
  for (int64_t m : d.mval) d.nalm += pl.lmax + 1 - m;
  d.d_mval.upload(d.mval.data(), d.mval.size(), st);
  d.d_mstart0.upload(d.mstart0.data(), d.mstart0.size(), st);
}
void make_map_desc(const hpsht_plan& pl, int t, RankMapDesc& d, cudaStream_t st) {
  const std::vector<int64_t> rings = hpsht::rr_rindexes(2 * pl.nside, t, pl.P);
  d.rstart0.clear();
  d.first_pix.clear();
  d.nphi.clear();
  d.npix = 0;
  for (int64_t r : rings) {
    const hpsht::RingGeom g = hpsht::healpix_ring(pl.nside, r);
    d.rstart0.push_back(d.npix);
    d.first_pix.push_back(g.first_pix);
    d.nphi.push_back(g.nphi);
    d.npix += g.nphi;
  }
  d.d_rstart0.upload(d.rstart0.data(), d.rstart0.size(), st);
  d.d_first_pix.upload(d.first_pix.data(), d.first_pix.size(), st);
  d.d_nphi.upload(d.nphi.data(), d.nphi.size(), st);
}

const RankAlmDesc& own_alm_desc(hpsht_plan& pl) {
  if (pl.own_alm.mval.empty()) {
    make_alm_desc(pl, pl.rank, pl.own_alm, pl.stream);
    HP_CUDA(cudaStreamSynchronize(pl.stream));
  }
  return pl.own_alm;
}
const RankMapDesc& own_map_desc(hpsht_plan& pl) {
  if (pl.own_map.nphi.empty()) {
    make_map_desc(pl, pl.rank, pl.own_map, pl.stream);
    HP_CUDA(cudaStreamSynchronize(pl.stream));
  }
  return pl.own_map;
}

void launch_alm_copy(const hpsht_plan& pl, const RankAlmDesc& d, double* shard, double* glob, int ncomp, int dir,
                     cudaStream_t st) {
  const int TB = 256;
  const int64_t nalm_tot = (pl.lmax + 1) * (pl.lmax + 2) / 2;
  dim3 grid((unsigned)((pl.lmax + 1 + TB - 1) / TB), (unsigned)d.mval.size());
  k_alm_shard_copy<<<grid, TB, 0, st>>>(reinterpret_cast<double2*>(shard), reinterpret_cast<double2*>(glob),
                                        d.d_mval.p, d.d_mstart0.p, pl.lmax, d.nalm, nalm_tot, ncomp, dir);
  HP_CUDA(cudaGetLastError());
}
void launch_map_copy(const hpsht_plan& pl, const RankMapDesc& d, double* shard, double* glob, int ncomp, int dir,
                     cudaStream_t st) {
  const int64_t npix = 12 * pl.nside * pl.nside;
  k_map_shard_copy<<<(unsigned)d.nphi.size(), 256, 0, st>>>(shard, glob, d.d_rstart0.p, d.d_first_pix.p, d.d_nphi.p,
                                                            d.npix, npix, ncomp, dir);
  HP_CUDA(cudaGetLastError());
}

// which: 0 = alm (complex), 1 = map.  dir: 0 = scatter from root, 1 = gather to root (root < 0: to all)
void shard_exchange(hpsht_plan& pl, int which, int dir, const double* in, double* out, int ncomp, int root) {
  HP_REQUIRE(ncomp >= 1 && ncomp <= 3, HPSHT_ERR_ARG, "ncomp must be 1..3");
  HP_REQUIRE(root < pl.P && (root >= 0 || dir == 1), HPSHT_ERR_ARG, "invalid root");
  HP_CUDA(cudaSetDevice(pl.device));
  cudaStream_t st = pl.stream;
  const int64_t gcount = which == 0 ? (pl.lmax + 1) * (pl.lmax + 2) / 2 * 2 : 12 * pl.nside * pl.nside;  // doubles/comp
  const int64_t scount = which == 0 ? pl.nalm_loc * 2 : pl.npix_loc;
  const bool have_glob = dir == 0 ? pl.rank == root : (root < 0 || pl.rank == root);
  const double* shard_in = dir == 1 ? in : nullptr;
  double* shard_out = dir == 0 ? out : nullptr;
  const double* glob_in = dir == 0 ? in : nullptr;
  double* glob_out = dir == 1 ? out : nullptr;
  HP_REQUIRE(dir == 0 ? shard_out != nullptr : shard_in != nullptr, HPSHT_ERR_ARG, "null shard pointer");
  HP_REQUIRE(!have_glob || (dir == 0 ? glob_in != nullptr : glob_out != nullptr), HPSHT_ERR_ARG,
             "the global array can not be null on a task that owns it");
  DevBuf<double> b_shard, b_glob, b_tmp;
  // device views
  double* d_shard;
  if (dir == 1) {
    d_shard = const_cast<double*>(stage_in(shard_in, (size_t)scount * ncomp, b_shard));
  } else {
    d_shard = stage_out(shard_out, (size_t)scount * ncomp, b_shard);
  }
  double* d_glob = nullptr;
  if (have_glob) {
    if (dir == 0) {
      d_glob = const_cast<double*>(stage_in(glob_in, (size_t)gcount * ncomp, b_glob));
    } else {
      d_glob = stage_out(glob_out, (size_t)gcount * ncomp, b_glob);
      // positions no shard covers do not exist (the shards partition the object), nothing to clear
    }
  }
  hpsht::NcclApi* N = pl.P > 1 ? &hpsht::NcclApi::get() : nullptr;
  RankAlmDesc ad_peer;
  RankMapDesc md_peer;
  const RankAlmDesc* ad = nullptr;
  const RankMapDesc* md = nullptr;
  auto desc = [&](int t) -> int64_t {  // descriptor of rank t (own: cached); returns its shard doubles per comp
    if (which == 0) {
      if (t == pl.rank) {
        ad = &own_alm_desc(pl);
      } else {
        make_alm_desc(pl, t, ad_peer, st);
        ad = &ad_peer;
      }
      return ad->nalm * 2;
    }
    if (t == pl.rank) {
      md = &own_map_desc(pl);
    } else {
      make_map_desc(pl, t, md_peer, st);
      md = &md_peer;
    }
    return md->npix;
  };
  auto copy = [&](double* shard, double* glob, int d) {
    if (which == 0)
      launch_alm_copy(pl, *ad, shard, glob, ncomp, d, st);
    else
      launch_map_copy(pl, *md, shard, glob, ncomp, d, st);
  };
  if (dir == 0) {
    // ---- scatter ----
    if (pl.rank == root) {
      for (int t = 0; t < pl.P; ++t) {
        const int64_t n = desc(t);
        if (t == pl.rank) {
          copy(d_shard, d_glob, 0);
        } else {
          b_tmp.ensure((size_t)n * ncomp);
          copy(b_tmp.p, d_glob, 0);
          HP_NCCL(N->Send(b_tmp.p, (size_t)n * ncomp, ncclDouble, t, pl.comm, st));
        }
        HP_CUDA(cudaStreamSynchronize(st));  // the descriptor and b_tmp are reused by the next rank
      }
    } else {
      HP_NCCL(N->Recv(d_shard, (size_t)scount * ncomp, ncclDouble, root, pl.comm, st));
    }
    HP_CUDA(cudaStreamSynchronize(st));
    if (shard_out != d_shard)
      HP_CUDA(cudaMemcpy(shard_out, d_shard, (size_t)scount * ncomp * sizeof(double), cudaMemcpyDeviceToHost));
  } else if (root >= 0) {
    // ---- gather to root ----
    if (pl.rank == root) {
      for (int t = 0; t < pl.P; ++t) {
        const int64_t n = desc(t);
        if (t == pl.rank) {
          copy(d_shard, d_glob, 1);
        } else {
          b_tmp.ensure((size_t)n * ncomp);
          HP_NCCL(N->Recv(b_tmp.p, (size_t)n * ncomp, ncclDouble, t, pl.comm, st));
          copy(b_tmp.p, d_glob, 1);
        }
        HP_CUDA(cudaStreamSynchronize(st));
      }
      if (glob_out != d_glob)
        HP_CUDA(cudaMemcpy(glob_out, d_glob, (size_t)gcount * ncomp * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
      HP_NCCL(N->Send(d_shard, (size_t)scount * ncomp, ncclDouble, root, pl.comm, st));
      HP_CUDA(cudaStreamSynchronize(st));
    }
  } else {
    // ---- allgather: every rank broadcasts its shard once ----
    for (int t = 0; t < pl.P; ++t) {
      const int64_t n = desc(t);
      double* src = d_shard;
      if (pl.P > 1) {
        b_tmp.ensure((size_t)n * ncomp);
        HP_NCCL(N->Broadcast(t == pl.rank ? d_shard : b_tmp.p, b_tmp.p, (size_t)n * ncomp, ncclDouble, t, pl.comm, st));
        src = b_tmp.p;
      }
      copy(src, d_glob, 1);
      HP_CUDA(cudaStreamSynchronize(st));
    }
    if (glob_out != d_glob)
      HP_CUDA(cudaMemcpy(glob_out, d_glob, (size_t)gcount * ncomp * sizeof(double), cudaMemcpyDeviceToHost));
  }
}

// dot(a, b) over one component each, allreduced (src/alm.jl:655-660)
double shard_dot(hpsht_plan& pl, const double* a, const double* b) {
  HP_REQUIRE(a && b, HPSHT_ERR_ARG, "null pointer");
  HP_CUDA(cudaSetDevice(pl.device));
  cudaStream_t st = pl.stream;
  DevBuf<double> ba, bb;
  const size_t n = (size_t)pl.nalm_loc * 2;
  const double* da = stage_in(a, n, ba);
  const double* db = a == b ? da : stage_in(b, n, bb);
  const RankAlmDesc& ad = own_alm_desc(pl);
  pl.shard_scratch.ensure((size_t)pl.nm_loc + 1);
  double* rows = pl.shard_scratch.p;
  double* res = rows + pl.nm_loc;
  k_dot_rows<<<(unsigned)pl.nm_loc, 256, 0, st>>>(reinterpret_cast<const double2*>(da),
                                                   reinterpret_cast<const double2*>(db), ad.d_mval.p,
                                                   ad.d_mstart0.p, pl.lmax, rows);
  k_sum_rows<<<1, 256, 0, st>>>(rows, pl.nm_loc, res);
  HP_CUDA(cudaGetLastError());
  if (pl.P > 1) HP_NCCL(hpsht::NcclApi::get().AllReduce(res, res, 1, ncclDouble, ncclSum, pl.comm, st));
  double h = 0.0;
  HP_CUDA(cudaMemcpyAsync(&h, res, sizeof(double), cudaMemcpyDeviceToHost, st));
  HP_CUDA(cudaStreamSynchronize(st));
  return h;
}

void shard_almxfl(hpsht_plan& pl, double* alm, const double* fl, int64_t nfl, int ncomp) {
  HP_REQUIRE(alm && fl && nfl >= 0 && ncomp >= 1, HPSHT_ERR_ARG, "invalid argument");
  HP_CUDA(cudaSetDevice(pl.device));
  cudaStream_t st = pl.stream;
  DevBuf<double> ba, bf;
  const size_t n = (size_t)pl.nalm_loc * 2 * ncomp;
  double* da = const_cast<double*>(stage_in(alm, n, ba));
  // zero-extended copy of fl (src/alm.jl:699-701)
  std::vector<double> f((size_t)pl.lmax + 1, 0.0);
  if (is_device_pointer(fl)) {
    HP_CUDA(cudaMemcpy(f.data(), fl, (size_t)std::min<int64_t>(nfl, pl.lmax + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    std::copy(fl, fl + std::min<int64_t>(nfl, pl.lmax + 1), f.begin());
  }
  bf.upload(f.data(), f.size(), st);
  const RankAlmDesc& ad = own_alm_desc(pl);
  const int TB = 256;
  dim3 grid((unsigned)((pl.lmax + 1 + TB - 1) / TB), (unsigned)pl.nm_loc);
  for (int c = 0; c < ncomp; ++c)
    k_almxfl<<<grid, TB, 0, st>>>(reinterpret_cast<double2*>(da) + (size_t)c * pl.nalm_loc, bf.p, ad.d_mval.p,
                                  ad.d_mstart0.p, pl.lmax);
  HP_CUDA(cudaGetLastError());
  HP_CUDA(cudaStreamSynchronize(st));
  if (da != alm) HP_CUDA(cudaMemcpy(alm, da, n * sizeof(double), cudaMemcpyDeviceToHost));
}

void shard_alm2cl(hpsht_plan& pl, const double* a, const double* b, double* cl) {
  HP_REQUIRE(a && b && cl, HPSHT_ERR_ARG, "null pointer");
  HP_CUDA(cudaSetDevice(pl.device));
  cudaStream_t st = pl.stream;
  DevBuf<double> ba, bb, bc;
  const size_t n = (size_t)pl.nalm_loc * 2;
  const double* da = stage_in(a, n, ba);
  const double* db = a == b ? da : stage_in(b, n, bb);
  double* dc = stage_out(cl, (size_t)pl.lmax + 1, bc);
  const RankAlmDesc& ad = own_alm_desc(pl);
  const int TB = 128;
  const int64_t nchunk = (pl.nm_loc + CL_CHUNK - 1) / CL_CHUNK;
  pl.shard_scratch.ensure((size_t)nchunk * (pl.lmax + 1));
  dim3 grid((unsigned)((pl.lmax + 1 + TB - 1) / TB), (unsigned)nchunk);
  k_alm2cl_part<<<grid, TB, 0, st>>>(reinterpret_cast<const double2*>(da), reinterpret_cast<const double2*>(db),
                                     ad.d_mval.p, ad.d_mstart0.p, pl.nm_loc, pl.lmax, pl.shard_scratch.p);
  k_alm2cl_sum<<<grid.x, TB, 0, st>>>(pl.shard_scratch.p, nchunk, pl.lmax, dc);
  HP_CUDA(cudaGetLastError());
  if (pl.P > 1)
    HP_NCCL(hpsht::NcclApi::get().AllReduce(dc, dc, (size_t)pl.lmax + 1, ncclDouble, ncclSum, pl.comm, st));
  HP_CUDA(cudaStreamSynchronize(st));
  if (dc != cl) HP_CUDA(cudaMemcpy(cl, dc, ((size_t)pl.lmax + 1) * sizeof(double), cudaMemcpyDeviceToHost));
}

}  // namespace
