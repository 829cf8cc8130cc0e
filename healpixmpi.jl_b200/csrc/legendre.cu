// legendre.cu — sm_100a FP64 kernels of the Legendre stage: alm2leg / leg2alm, spin 0 and spin 2.
//
// Replaces Ducc0.Sht.alm2leg! (reference src/sht.jl:85) and Ducc0.Sht.leg2alm! (src/sht.jl:178).
//
// Kernel design (B200-first):
//   * one CTA = one work item = (one local m) x (one tile of N/S ring pairs, sorted pole->equator);
//   * every consumer thread owns R ring pairs and evaluates the scaled recurrence on the fly in
//     registers (no Plm table); north/south symmetry gives both rings of a pair from one recurrence;
//   * the per-m coefficient stream (recurrence coefficients + prepared alm) is the only data the
//     inner loop reads; thread 0 of the CTA moves it global->shared with 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx) through a 4-stage full/empty mbarrier ring, and the
//     consumers read it as warp-broadcast LDS.128;
//   * dynamic range: value = lam * 2^(800*sc), sc <= 0 tracked per ring; a warp whose lanes are all
//     still below range runs a 2-FMA "ramp" step instead of the full 6-FMA (spin 0) / 12-FMA
//     (spin 2) step; range checks every 8 steps;
//   * adjoint: per-thread accumulation over its R rings, transposed-butterfly warp-shuffle reduction
//     over the 32 lanes every 4 steps, cross-warp reduction through shared memory once per 64-step
//     chunk, deterministic per-tile partial sums in global memory (no atomics).
#include "legendre.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace hpsht {

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int NSTAGE = 4;
constexpr double SC_DOWN = 0x1p-800;  // exact power of two: rescaling never rounds
constexpr int SC_STEP = 800;

struct __align__(16) Coef {
  double a, b;
};
struct __align__(16) Strm {
  double x, y, z, w;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread is parked by the hardware until the phase completes or
// the hint (ns) expires, instead of re-issuing the probe every few cycles
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// consumer side: the data is normally there already (the producer runs NSTAGE-1 chunks ahead)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity, 1000u)) {
  }
}
// producer side: a stage is released once per chunk (microseconds); poll politely so that the spinning
// warp does not take issue slots from the FP64 warps (ncu: 23 % of all issued instructions were the
// SYNCS/BRA/YIELD of these loops before)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity, 4000u)) __nanosleep(200);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// x = m * 2^e with |m| in [0.5, 1): frexp for NORMAL non-zero doubles, done on the integer pipe so
// that the start-value computation does not compete with the recurrence for the FP64 pipe.
__device__ __forceinline__ void split_normal(double x, double& m, int& e) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  e = ((hi >> 20) & 0x7ff) - 1022;
  m = __hiloint2double((hi & 0x800fffff) | 0x3fe00000, lo);
}
// 2^k for -1022 <= k <= 1023
__device__ __forceinline__ double pow2i(int k) { return __hiloint2double((k + 1023) << 20, 0); }

// base^n (n >= 0) as mant * 2^e, mant in [0.5, 1); base must be 0 or a normal double
__device__ __forceinline__ void pow_scaled(double base, int n, double& mant, int& e) {
  mant = 1.0;
  e = 0;
  if (n == 0) return;
  if (base == 0.0) {
    mant = 0.0;
    return;
  }
  int be;
  double b;
  split_normal(base, b, be);
  while (n > 0) {
    if (n & 1) {
      int t;
      split_normal(mant * b, mant, t);
      e += be + t;
    }
    int t;
    split_normal(b * b, b, t);
    be = 2 * be + t;
    n >>= 1;
  }
}
// v * 2^e -> lam * 2^(800 sc), sc <= 0, |lam| in [2^-400, 2^400) when sc < 0
__device__ __forceinline__ void to_scaled(double v, int e, double& lam, int& sc) {
  if (v == 0.0 || fabs(v) < 1e-290) {  // zero (or hopelessly small: treated as zero)
    lam = 0.0;
    sc = 0;
    return;
  }
  int t;
  split_normal(v, v, t);
  e += t;
  sc = (e >= -400) ? 0 : -((-e - 400 + SC_STEP - 1) / SC_STEP);
  const int ee = e - SC_STEP * sc;  // in (-400, 400] when sc < 0; may be very negative when sc == 0
  lam = (ee < -1000) ? 0.0 : v * pow2i(ee);
}
__device__ __forceinline__ bool is_big(double v) {
  // |v| > 2^400  <=>  biased exponent field > 1023+400 (ties at exactly 2^400 are irrelevant)
  return ((__double2hiint(v) >> 20) & 0x7ff) > 1023 + 400;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// kernel parameter block
// ------------------------------------------------------------------------------------------------
struct LegParams {
  const WorkItem* items;
  // per local m
  const int64_t* mval;
  const int64_t* nstep;   // real steps
  const int64_t* off;     // table offset in steps (padded to LEG_CHUNK)
  const double* coef;     // 2 doubles / step
  const double* k0;       // spin0: p_0 constant ; spin2: kp2
  const double* k1;       // spin2: km2
  // per pair
  const double* cth;
  const double* sth;
  const double* sh;
  const double* ch;
  const int* mlim;
  // leg addressing: element (pair p, local m index mi, component c) of ring "N" lives at
  //   a0[p] + nrr[p] * (ncs * sstart[mi] + c * nms[mi]) + jN[p] * nms[mi] + (mi - sstart[mi])
  // i.e. per destination block [m-slab][component][ring][m within slab]; a single slab with
  // sstart = 0, nms = nm is the reference's (nm, nring, ncomp) array.
  const int64_t* a0;
  const int64_t* nrr;
  const int64_t* jN;
  const int64_t* jS;      // -1: unpaired
  const int64_t* sstart;  // per local m
  const int64_t* nms;     // per local m
  int64_t ncs;            // components the slab blocks are sized for
  // data
  const double* stream;   // 4 doubles / step (synthesis input)
  double* leg;            // complex128 (synthesis out / adjoint in)
  double* part;           // adjoint partial sums, 4 doubles / step
  const int64_t* partoff; // per item: offset in steps into part
  int accumulate;         // adjoint: 1 = the partial sums of this item already hold earlier ring blocks
};

__device__ __forceinline__ int64_t leg_rowoff(const LegParams& P, int p, int mi) {
  const int64_t ss = P.sstart[mi];
  return P.nrr[p] * (P.ncs * ss) + ((int64_t)mi - ss);
}
__device__ __forceinline__ int64_t leg_offN(const LegParams& P, int p, int mi) {
  return P.a0[p] + P.jN[p] * P.nms[mi];
}
__device__ __forceinline__ int64_t leg_offS(const LegParams& P, int p, int mi) {
  const int64_t j = P.jS[p];
  return j < 0 ? -1 : P.a0[p] + j * P.nms[mi];
}
__device__ __forceinline__ int64_t leg_cstr(const LegParams& P, int p, int mi) { return P.nrr[p] * P.nms[mi]; }

// ------------------------------------------------------------------------------------------------
// producer side of the coefficient pipeline.  One thread issues the 1-D TMA bulk copies of chunk
// `cn` (counted over the whole CTA lifetime; the per-m stream is replayed once per tile in the
// adjoint) into stage cn % NSTAGE once every consumer warp released that stage.
//   PROD == 1: a dedicated producer warp runs `pump` to completion (blocking on `empty`).
//   PROD == 2: lane 0 of the last consumer warp calls `pump` at every chunk start and group
//              boundary; it only blocks for the chunk it is about to consume itself and otherwise
//              probes `empty` without waiting, so the consumer warps stay decoupled.
// ------------------------------------------------------------------------------------------------
template <bool WITH_STREAM>
struct Pump {
  int nissued, total, nchunks;
  const double* gcoef;
  const double* gstrm;
  Coef (*s_coef)[LEG_CHUNK];
  Strm (*s_strm)[LEG_CHUNK];
  uint64_t* full;
  uint64_t* empty;

  __device__ __forceinline__ void issue(int cn) {
    const int st = cn % NSTAGE;
    const int c = cn % nchunks;
    const uint32_t bytes = LEG_CHUNK * sizeof(Coef) + (WITH_STREAM ? LEG_CHUNK * sizeof(Strm) : 0);
    mbar_arrive_expect_tx(&full[st], bytes);
    tma_load_1d(&s_coef[st][0], gcoef + (size_t)c * LEG_CHUNK * 2, LEG_CHUNK * sizeof(Coef), &full[st]);
    if (WITH_STREAM)
      tma_load_1d(&s_strm[st][0], gstrm + (size_t)c * LEG_CHUNK * 4, LEG_CHUNK * sizeof(Strm), &full[st]);
  }
  // issue everything that can go out now, at most NSTAGE-1 chunks beyond `cg`; wait only for `cg`
  __device__ __forceinline__ void run(int cg) {
    while (nissued < total && nissued <= cg + NSTAGE - 1) {
      if (nissued >= NSTAGE) {
        const int st = nissued % NSTAGE;
        const uint32_t par = ((nissued / NSTAGE) - 1) & 1;
        if (nissued <= cg)
          mbar_wait(&empty[st], par);
        else if (!mbar_test(&empty[st], par))
          break;
      }
      issue(nissued);
      ++nissued;
    }
  }
  // dedicated producer: everything, blocking
  __device__ __forceinline__ void run_all() {
    for (; nissued < total; ++nissued) {
      if (nissued >= NSTAGE) mbar_wait_relaxed(&empty[nissued % NSTAGE], ((nissued / NSTAGE) - 1) & 1);
      issue(nissued);
    }
  }
};

// ================================================================================================
// spin 0 synthesis:  leg_N/S = sum_n p_n A_n  +-  x sum_n p_n B_n
// ================================================================================================
template <int R, int NW, int PROD>
__global__ void __launch_bounds__((NW + 1) * 32) k_alm2leg_s0(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ Strm s_strm[NSTAGE][LEG_CHUNK];
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  __shared__ __align__(8) uint64_t s_empty[NSTAGE];

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nn = (int)P.nstep[mi];
  const int nsteps = (nn + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const double* gcoef = P.coef + (size_t)P.off[mi] * 2;
  const double* gstrm = P.stream + (size_t)P.off[mi] * 4;
  Pump<true> pump{0, nchunks, nchunks, gcoef, gstrm, s_coef, s_strm, s_full, s_empty};
  if (warp >= NW) {
    if (PROD == 1 && warp == NW && lane == 0) pump.run_all();
    return;
  }
  const bool issuer = (PROD == 2) && (threadIdx.x == (NW - 1) * 32);
  if (issuer) pump.run(0);

  // ---------------- consumer warps ----------------
  double csq[R], lam1[R], lam2[R];
  double er[R], ei[R], orr[R], oi[R];
  int sc[R];
  bool act[R];
  bool any_act = false;
  const double k0 = P.k0[mi];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pl = warp * 32 * R + r * 32 + lane;
    const bool valid = pl < it.np;
    const int p = it.p0 + (valid ? pl : 0);
    act[r] = valid && (m <= P.mlim[p]);
    any_act |= act[r];
    const double c = P.cth[p];
    csq[r] = c * c;
    double mant;
    int e;
    pow_scaled(P.sth[p], m, mant, e);
    to_scaled(k0 * mant, e, lam2[r], sc[r]);
    if (!act[r]) {  // culled lanes carry zeros so nothing can overflow
      lam2[r] = 0.0;
      sc[r] = 0;
    }
    lam1[r] = 0.0;
    er[r] = ei[r] = orr[r] = oi[r] = 0.0;
  }
  const bool warp_act = __any_sync(0xffffffffu, any_act);
  // warp-uniform range flags, re-voted only when a scale changes (not once per group)
  bool w_any0 = false, w_neg = false;
  auto revote = [&]() {
    bool mine0 = false, mineneg = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mine0 |= act[r] && (sc[r] == 0);
      mineneg |= (sc[r] < 0);
    }
    w_any0 = __any_sync(0xffffffffu, mine0);
    w_neg = __any_sync(0xffffffffu, mineneg);
  };
  revote();

  for (int c = 0; c < nchunks; ++c) {
    const int st = c % NSTAGE;
    if (PROD == 2) {
      if (issuer) pump.run(c);
      __syncwarp();
    }
    mbar_wait(&s_full[st], (c / NSTAGE) & 1);
    if (warp_act) {
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      for (int g = 0; g < ng; ++g) {
        const Coef* cf = &s_coef[st][g * LEG_GROUP];
        if (w_any0) {
          const Strm* sm = &s_strm[st][g * LEG_GROUP];
#pragma unroll
          for (int s = 0; s < LEG_GROUP; ++s) {
            const Coef ab = cf[s];
            const Strm al = sm[s];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              er[r] = fma(lam2[r], al.x, er[r]);
              ei[r] = fma(lam2[r], al.y, ei[r]);
              orr[r] = fma(lam2[r], al.z, orr[r]);
              oi[r] = fma(lam2[r], al.w, oi[r]);
              const double t = fma(ab.a, csq[r], ab.b);
              const double nx = fma(t, lam2[r], lam1[r]);
              lam1[r] = lam2[r];
              lam2[r] = nx;
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < LEG_GROUP; ++s) {
            const Coef ab = cf[s];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const double t = fma(ab.a, csq[r], ab.b);
              const double nx = fma(t, lam2[r], lam1[r]);
              lam1[r] = lam2[r];
              lam2[r] = nx;
            }
          }
        }
        if (w_neg) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (sc[r] < 0 && is_big(lam2[r])) {
              lam1[r] *= SC_DOWN;
              lam2[r] *= SC_DOWN;
              sc[r] += 1;
              er[r] = ei[r] = orr[r] = oi[r] = 0.0;  // drop what was accumulated while scaled
            }
          }
          revote();  // scales changed: refresh the warp-uniform flags
        }
        if (PROD == 2 && issuer) pump.run(c);  // opportunistic refill, never blocks here
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[st]);
  }

  // ---------------- epilogue ----------------
  double2* leg = reinterpret_cast<double2*>(P.leg);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pl = warp * 32 * R + r * 32 + lane;
    if (pl >= it.np) continue;
    const int p = it.p0 + pl;
    const bool on = act[r] && sc[r] == 0;
    const double c = P.cth[p];
    const double Er = on ? er[r] : 0.0, Ei = on ? ei[r] : 0.0;
    const double Or = on ? orr[r] * c : 0.0, Oi = on ? oi[r] * c : 0.0;
    const int64_t rowoff = leg_rowoff(P, p, mi);
    leg[leg_offN(P, p, mi) + rowoff] = make_double2(Er + Or, Ei + Oi);
    const int64_t oS = leg_offS(P, p, mi);
    if (oS >= 0) leg[oS + rowoff] = make_double2(Er - Or, Ei - Oi);
  }
}

// ================================================================================================
// spin 2 synthesis.  Stream per l: alpha+ = -(E+iB) sigma_l/2, alpha- = -(E-iB) sigma_l/2.
//   P+N = sum alpha+ d+ ; P-N = sum alpha- d- ; P+S = sum (-1)^(l+m) alpha+ d- ; P-S = sum (-1)^(l+m) alpha- d+
//   Q = P+ + P- ; U = -i (P+ - P-)
// ================================================================================================
template <int R, int NW, int PROD>
__global__ void __launch_bounds__((NW + 1) * 32) k_alm2leg_s2(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ Strm s_strm[NSTAGE][LEG_CHUNK];
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  __shared__ __align__(8) uint64_t s_empty[NSTAGE];

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nl = (int)P.nstep[mi];
  const int nsteps = (nl + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const double* gcoef = P.coef + (size_t)P.off[mi] * 2;
  const double* gstrm = P.stream + (size_t)P.off[mi] * 4;
  Pump<true> pump{0, nchunks, nchunks, gcoef, gstrm, s_coef, s_strm, s_full, s_empty};
  if (warp >= NW) {
    if (PROD == 1 && warp == NW && lane == 0) pump.run_all();
    return;
  }
  const bool issuer = (PROD == 2) && (threadIdx.x == (NW - 1) * 32);
  if (issuer) pump.run(0);

  double x[R], lp1[R], lp2[R], lm1[R], lm2[R];
  double pnr[R], pni[R], mnr[R], mni[R];  // P+N, P-N
  double psr[R], psi[R], msr[R], msi[R];  // P+S, P-S
  int scp[R], scm[R];
  bool act[R];
  bool any_act = false;
  const double kp = P.k0[mi], km = P.k1[mi];
  const int pw = m >= 2 ? m - 2 : 2 - m;
  const int q = m < 2 ? m : 2;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pl = warp * 32 * R + r * 32 + lane;
    const bool valid = pl < it.np;
    const int p = it.p0 + (valid ? pl : 0);
    act[r] = valid && (m <= P.mlim[p]);
    any_act |= act[r];
    x[r] = P.cth[p];
    double mant;
    int e;
    pow_scaled(P.sth[p], pw, mant, e);
    double fs = 1.0, fc = 1.0;
    const double s2 = P.sh[p] * P.sh[p], c2 = P.ch[p] * P.ch[p];
    for (int k = 0; k < q; ++k) {
      fs *= s2;
      fc *= c2;
    }
    to_scaled(kp * mant * fs, e, lp2[r], scp[r]);
    to_scaled(km * mant * fc, e, lm2[r], scm[r]);
    if (!act[r]) {
      lp2[r] = lm2[r] = 0.0;
      scp[r] = scm[r] = 0;
    }
    lp1[r] = lm1[r] = 0.0;
    pnr[r] = pni[r] = mnr[r] = mni[r] = 0.0;
    psr[r] = psi[r] = msr[r] = msi[r] = 0.0;
  }
  const bool warp_act = __any_sync(0xffffffffu, any_act);
  // warp-uniform range flags, re-voted only when a scale changes (not once per group)
  bool w_any0 = false, w_neg = false;
  auto revote = [&]() {
    bool mine0 = false, mineneg = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mine0 |= act[r] && ((scp[r] == 0) || (scm[r] == 0));
      mineneg |= (scp[r] < 0) || (scm[r] < 0);
    }
    w_any0 = __any_sync(0xffffffffu, mine0);
    w_neg = __any_sync(0xffffffffu, mineneg);
  };
  revote();

  for (int c = 0; c < nchunks; ++c) {
    const int st = c % NSTAGE;
    if (PROD == 2) {
      if (issuer) pump.run(c);
      __syncwarp();
    }
    mbar_wait(&s_full[st], (c / NSTAGE) & 1);
    if (warp_act) {
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      for (int g = 0; g < ng; ++g) {
        const Coef* cf = &s_coef[st][g * LEG_GROUP];
        if (w_any0) {
          const Strm* sm = &s_strm[st][g * LEG_GROUP];
#pragma unroll
          for (int s = 0; s < LEG_GROUP; ++s) {
            const Coef ab = cf[s];
            const Strm al = sm[s];  // x,y = alpha+ ; z,w = alpha-
#pragma unroll
            for (int r = 0; r < R; ++r) {
              pnr[r] = fma(lp2[r], al.x, pnr[r]);
              pni[r] = fma(lp2[r], al.y, pni[r]);
              mnr[r] = fma(lm2[r], al.z, mnr[r]);
              mni[r] = fma(lm2[r], al.w, mni[r]);
              if ((s & 1) == 0) {  // steps within a group start at an even offset
                psr[r] = fma(lm2[r], al.x, psr[r]);
                psi[r] = fma(lm2[r], al.y, psi[r]);
                msr[r] = fma(lp2[r], al.z, msr[r]);
                msi[r] = fma(lp2[r], al.w, msi[r]);
              } else {
                psr[r] = fma(-lm2[r], al.x, psr[r]);
                psi[r] = fma(-lm2[r], al.y, psi[r]);
                msr[r] = fma(-lp2[r], al.z, msr[r]);
                msi[r] = fma(-lp2[r], al.w, msi[r]);
              }
              const double tp = fma(x[r], ab.a, ab.b);
              const double tm = fma(x[r], ab.a, -ab.b);
              const double np_ = fma(tp, lp2[r], -lp1[r]);
              const double nm_ = fma(tm, lm2[r], -lm1[r]);
              lp1[r] = lp2[r];
              lp2[r] = np_;
              lm1[r] = lm2[r];
              lm2[r] = nm_;
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < LEG_GROUP; ++s) {
            const Coef ab = cf[s];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const double tp = fma(x[r], ab.a, ab.b);
              const double tm = fma(x[r], ab.a, -ab.b);
              const double np_ = fma(tp, lp2[r], -lp1[r]);
              const double nm_ = fma(tm, lm2[r], -lm1[r]);
              lp1[r] = lp2[r];
              lp2[r] = np_;
              lm1[r] = lm2[r];
              lm2[r] = nm_;
            }
          }
        }
        if (w_neg) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (scp[r] < 0 && is_big(lp2[r])) {
              lp1[r] *= SC_DOWN;
              lp2[r] *= SC_DOWN;
              scp[r] += 1;
              pnr[r] = pni[r] = msr[r] = msi[r] = 0.0;
            }
            if (scm[r] < 0 && is_big(lm2[r])) {
              lm1[r] *= SC_DOWN;
              lm2[r] *= SC_DOWN;
              scm[r] += 1;
              mnr[r] = mni[r] = psr[r] = psi[r] = 0.0;
            }
          }
          revote();  // scales changed: refresh the warp-uniform flags
        }
        if (PROD == 2 && issuer) pump.run(c);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[st]);
  }

  double2* leg = reinterpret_cast<double2*>(P.leg);
  const int l0 = m > 2 ? m : 2;
  const double sgnS = ((l0 + m) & 1) ? -1.0 : 1.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pl = warp * 32 * R + r * 32 + lane;
    if (pl >= it.np) continue;
    const int p = it.p0 + pl;
    const bool onp = act[r] && scp[r] == 0, onm = act[r] && scm[r] == 0;
    const double PNr = onp ? pnr[r] : 0.0, PNi = onp ? pni[r] : 0.0;
    const double MSr = onp ? msr[r] * sgnS : 0.0, MSi = onp ? msi[r] * sgnS : 0.0;
    const double MNr = onm ? mnr[r] : 0.0, MNi = onm ? mni[r] : 0.0;
    const double PSr = onm ? psr[r] * sgnS : 0.0, PSi = onm ? psi[r] * sgnS : 0.0;
    const int64_t rowoff = leg_rowoff(P, p, mi);
    const int64_t cs = leg_cstr(P, p, mi);
    const int64_t oN = leg_offN(P, p, mi) + rowoff;
    // Q = P+ + P- ; U = -i (P+ - P-) = ( Im(P+ - P-), -Re(P+ - P-) )
    leg[oN] = make_double2(PNr + MNr, PNi + MNi);
    leg[oN + cs] = make_double2(PNi - MNi, -(PNr - MNr));
    const int64_t oS0 = leg_offS(P, p, mi);
    if (oS0 >= 0) {
      const int64_t oS = oS0 + rowoff;
      leg[oS] = make_double2(PSr + MSr, PSi + MSi);
      leg[oS + cs] = make_double2(PSi - MSi, -(PSr - MSr));
    }
  }
}

// ================================================================================================
// adjoint kernels
// ================================================================================================
// transposed butterfly: on entry every lane holds 16 partial values v[0..15]; on exit lane j holds in
// v[0] the sum over all 32 lanes of value (j & 15).
__device__ __forceinline__ void warp_reduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const double send = up ? v[i] : v[i + s];
      const double keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// cross-warp reduction of one chunk and store of the per-tile partial sums
template <int NW>
__device__ __forceinline__ void flush_chunk(const double (*s_red)[LEG_CHUNK * 4], double* gpart,
                                            int tid_consumer, bool first) {
  for (int i = tid_consumer; i < LEG_CHUNK * 4; i += NW * 32) {
    double s = first ? 0.0 : gpart[i];  // same thread owns gpart[i] for the whole item
#pragma unroll
    for (int w = 0; w < NW; ++w) s += s_red[w][i];
    gpart[i] = s;
  }
}

// spin 0 adjoint: At_n = sum_pairs p_n (legN + legS), Bt_n = sum_pairs p_n x (legN - legS)
// One work item = one m and a contiguous RANGE of pairs; the CTA walks the range tile by tile and
// accumulates its own partial sums in global memory (read-modify-write by the same thread, L2
// resident), so the number of partial buffers per m is the number of items, not of tiles.
template <int R, int NW, int PROD>
__global__ void __launch_bounds__((NW + (PROD == 1 ? 1 : 0)) * 32, 2) k_leg2alm_s0(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ double s_red[2][NW][LEG_CHUNK * 4];
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  __shared__ __align__(8) uint64_t s_empty[NSTAGE];
  constexpr int TILE = R * NW * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nn = (int)P.nstep[mi];
  const int nsteps = (nn + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int ntile = (it.np + TILE - 1) / TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const double* gcoef = P.coef + (size_t)P.off[mi] * 2;
  const int total_chunks = nchunks * ntile;
  Pump<false> pump{0, total_chunks, nchunks, gcoef, nullptr, s_coef, nullptr, s_full, s_empty};
  if (warp >= NW) {
    if (PROD == 1 && warp == NW && lane == 0) pump.run_all();
    return;
  }
  const bool issuer = (PROD == 2) && (threadIdx.x == (NW - 1) * 32);
  if (issuer) pump.run(0);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double k0 = P.k0[mi];
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  const int tidc = threadIdx.x;  // consumer threads are 0 .. NW*32-1
  int cg = 0;                    // global chunk counter (pipeline position)

  for (int tile = 0; tile < ntile; ++tile) {
    const int tp0 = it.p0 + tile * TILE;
    const int tnp = min(TILE, it.np - tile * TILE);
    double csq[R], lam1[R], lam2[R];
    double er[R], ei[R], orr[R], oi[R];
    int sc[R];
    bool act[R];
    bool any_act = false;

    auto load_in = [&](int r, int p) {
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const double2 gN = leg[leg_offN(P, p, mi) + rowoff];
      const int64_t oS = leg_offS(P, p, mi);
      const double2 gS = oS >= 0 ? leg[oS + rowoff] : make_double2(0.0, 0.0);
      const double cc = P.cth[p];
      er[r] = gN.x + gS.x;
      ei[r] = gN.y + gS.y;
      orr[r] = cc * (gN.x - gS.x);
      oi[r] = cc * (gN.y - gS.y);
    };

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pl = warp * 32 * R + r * 32 + lane;
      const bool valid = pl < tnp;
      const int p = tp0 + (valid ? pl : 0);
      act[r] = valid && (m <= P.mlim[p]);
      any_act |= act[r];
      const double c = P.cth[p];
      csq[r] = c * c;
      double mant;
      int e;
      pow_scaled(P.sth[p], m, mant, e);
      to_scaled(k0 * mant, e, lam2[r], sc[r]);
      if (!act[r]) {
        lam2[r] = 0.0;
        sc[r] = 0;
      }
      lam1[r] = 0.0;
      er[r] = ei[r] = orr[r] = oi[r] = 0.0;
      if (act[r] && sc[r] == 0) load_in(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);

    for (int c = 0; c < nchunks; ++c, ++cg) {
      const int st = cg % NSTAGE;
      double* red = &s_red[cg & 1][warp][0];
      if (PROD == 2) {
        if (issuer) pump.run(cg);
        __syncwarp();
      }
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      for (int g = 0; g < LEG_CHUNK / LEG_GROUP; ++g) {
        bool did = false;
        if (warp_act && g < ng) {
          bool mine0 = false, mineneg = false;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            mine0 |= act[r] && (sc[r] == 0);
            mineneg |= (sc[r] < 0);
          }
          const bool any0 = __any_sync(0xffffffffu, mine0);
          const Coef* cf = &s_coef[st][g * LEG_GROUP];
          if (any0) {
            did = true;
#pragma unroll
            for (int h = 0; h < LEG_GROUP / 4; ++h) {
              double acc[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const Coef ab = cf[h * 4 + s];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  acc[4 * s + 0] = fma(lam2[r], er[r], acc[4 * s + 0]);
                  acc[4 * s + 1] = fma(lam2[r], ei[r], acc[4 * s + 1]);
                  acc[4 * s + 2] = fma(lam2[r], orr[r], acc[4 * s + 2]);
                  acc[4 * s + 3] = fma(lam2[r], oi[r], acc[4 * s + 3]);
                  const double t = fma(ab.a, csq[r], ab.b);
                  const double nx = fma(t, lam2[r], lam1[r]);
                  lam1[r] = lam2[r];
                  lam2[r] = nx;
                }
              }
              warp_reduce16(acc, lane);
              if (lane < 16) red[(g * LEG_GROUP + h * 4) * 4 + lane] = acc[0];
            }
          } else {
#pragma unroll
            for (int s = 0; s < LEG_GROUP; ++s) {
              const Coef ab = cf[s];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const double t = fma(ab.a, csq[r], ab.b);
                const double nx = fma(t, lam2[r], lam1[r]);
                lam1[r] = lam2[r];
                lam2[r] = nx;
              }
            }
          }
          if (__any_sync(0xffffffffu, mineneg)) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (sc[r] < 0 && is_big(lam2[r])) {
                lam1[r] *= SC_DOWN;
                lam2[r] *= SC_DOWN;
                sc[r] += 1;
                // ring enters range: fetch its leg values now (masked to 0 until then)
                if (sc[r] == 0) load_in(r, tp0 + warp * 32 * R + r * 32 + lane);
              }
            }
          }
        }
        if (PROD == 2 && issuer) pump.run(cg);
        if (!did) red[g * LEG_GROUP * 4 + lane] = 0.0;  // 8 steps x 4 values = 32 doubles
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      named_bar_sync(1, NW * 32);
      flush_chunk<NW>(s_red[cg & 1], gpart + (size_t)c * LEG_CHUNK * 4, tidc, tile == 0);
    }
  }
}

// spin 2 adjoint:  G+_l = sum_pairs [ d+ w+N + (-1)^(l+m) d- w+S ],  G-_l = sum_pairs [ d- w-N + (-1)^(l+m) d+ w-S ]
//   w+- = legQ +- i legU
template <int R, int NW, int PROD>
__global__ void __launch_bounds__((NW + (PROD == 1 ? 1 : 0)) * 32, 2) k_leg2alm_s2(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ double s_red[2][NW][LEG_CHUNK * 4];
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  __shared__ __align__(8) uint64_t s_empty[NSTAGE];
  constexpr int TILE = R * NW * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nl = (int)P.nstep[mi];
  const int nsteps = (nl + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int ntile = (it.np + TILE - 1) / TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const double* gcoef = P.coef + (size_t)P.off[mi] * 2;
  const int total_chunks = nchunks * ntile;
  Pump<false> pump{0, total_chunks, nchunks, gcoef, nullptr, s_coef, nullptr, s_full, s_empty};
  if (warp >= NW) {
    if (PROD == 1 && warp == NW && lane == 0) pump.run_all();
    return;
  }
  const bool issuer = (PROD == 2) && (threadIdx.x == (NW - 1) * 32);
  if (issuer) pump.run(0);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double kp = P.k0[mi], km = P.k1[mi];
  const int pw = m >= 2 ? m - 2 : 2 - m;
  const int q = m < 2 ? m : 2;
  const int l0 = m > 2 ? m : 2;
  const double sgnS = ((l0 + m) & 1) ? -1.0 : 1.0;
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  const int tidc = threadIdx.x;
  int cg = 0;

  for (int tile = 0; tile < ntile; ++tile) {
    const int tp0 = it.p0 + tile * TILE;
    const int tnp = min(TILE, it.np - tile * TILE);
    double x[R], lp1[R], lp2[R], lm1[R], lm2[R];
    // masked inputs: wpN (x d+), wmS (x d+ x sign) live with scp ; wmN (x d-), wpS (x d- x sign) with scm
    double wpNr[R], wpNi[R], wmSr[R], wmSi[R], wmNr[R], wmNi[R], wpSr[R], wpSi[R];
    int scp[R], scm[R];
    bool act[R];
    bool any_act = false;

    auto load_p = [&](int r, int p) {  // inputs multiplied by d+
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const int64_t cs = leg_cstr(P, p, mi);
      const double2 qN = leg[leg_offN(P, p, mi) + rowoff], uN = leg[leg_offN(P, p, mi) + rowoff + cs];
      // w+N = Q + iU = (qr - ui, qi + ur)
      wpNr[r] = qN.x - uN.y;
      wpNi[r] = qN.y + uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS + rowoff], uS = leg[oS + rowoff + cs];
        // w-S = Q - iU = (qr + ui, qi - ur)
        wmSr[r] = sgnS * (qS.x + uS.y);
        wmSi[r] = sgnS * (qS.y - uS.x);
      } else {
        wmSr[r] = wmSi[r] = 0.0;
      }
    };
    auto load_m = [&](int r, int p) {  // inputs multiplied by d-
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const int64_t cs = leg_cstr(P, p, mi);
      const double2 qN = leg[leg_offN(P, p, mi) + rowoff], uN = leg[leg_offN(P, p, mi) + rowoff + cs];
      wmNr[r] = qN.x + uN.y;
      wmNi[r] = qN.y - uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS + rowoff], uS = leg[oS + rowoff + cs];
        wpSr[r] = sgnS * (qS.x - uS.y);
        wpSi[r] = sgnS * (qS.y + uS.x);
      } else {
        wpSr[r] = wpSi[r] = 0.0;
      }
    };

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pl = warp * 32 * R + r * 32 + lane;
      const bool valid = pl < tnp;
      const int p = tp0 + (valid ? pl : 0);
      act[r] = valid && (m <= P.mlim[p]);
      any_act |= act[r];
      x[r] = P.cth[p];
      double mant;
      int e;
      pow_scaled(P.sth[p], pw, mant, e);
      double fs = 1.0, fc = 1.0;
      const double s2 = P.sh[p] * P.sh[p], c2 = P.ch[p] * P.ch[p];
      for (int k = 0; k < q; ++k) {
        fs *= s2;
        fc *= c2;
      }
      to_scaled(kp * mant * fs, e, lp2[r], scp[r]);
      to_scaled(km * mant * fc, e, lm2[r], scm[r]);
      if (!act[r]) {
        lp2[r] = lm2[r] = 0.0;
        scp[r] = scm[r] = 0;
      }
      lp1[r] = lm1[r] = 0.0;
      wpNr[r] = wpNi[r] = wmSr[r] = wmSi[r] = 0.0;
      wmNr[r] = wmNi[r] = wpSr[r] = wpSi[r] = 0.0;
      if (act[r] && scp[r] == 0) load_p(r, p);
      if (act[r] && scm[r] == 0) load_m(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);

    for (int c = 0; c < nchunks; ++c, ++cg) {
      const int st = cg % NSTAGE;
      double* red = &s_red[cg & 1][warp][0];
      if (PROD == 2) {
        if (issuer) pump.run(cg);
        __syncwarp();
      }
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      for (int g = 0; g < LEG_CHUNK / LEG_GROUP; ++g) {
        bool did = false;
        if (warp_act && g < ng) {
          bool mine0 = false, mineneg = false;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            mine0 |= act[r] && ((scp[r] == 0) || (scm[r] == 0));
            mineneg |= (scp[r] < 0) || (scm[r] < 0);
          }
          const bool any0 = __any_sync(0xffffffffu, mine0);
          const Coef* cf = &s_coef[st][g * LEG_GROUP];
          if (any0) {
            did = true;
#pragma unroll
            for (int h = 0; h < LEG_GROUP / 4; ++h) {
              double acc[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const Coef ab = cf[h * 4 + s];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  // G+ : d+ w+N  +- d- w+S ; G- : d- w-N +- d+ w-S   (sign alternates with the step)
                  acc[4 * s + 0] = fma(lp2[r], wpNr[r], acc[4 * s + 0]);
                  acc[4 * s + 1] = fma(lp2[r], wpNi[r], acc[4 * s + 1]);
                  acc[4 * s + 2] = fma(lm2[r], wmNr[r], acc[4 * s + 2]);
                  acc[4 * s + 3] = fma(lm2[r], wmNi[r], acc[4 * s + 3]);
                  if ((s & 1) == 0) {
                    acc[4 * s + 0] = fma(lm2[r], wpSr[r], acc[4 * s + 0]);
                    acc[4 * s + 1] = fma(lm2[r], wpSi[r], acc[4 * s + 1]);
                    acc[4 * s + 2] = fma(lp2[r], wmSr[r], acc[4 * s + 2]);
                    acc[4 * s + 3] = fma(lp2[r], wmSi[r], acc[4 * s + 3]);
                  } else {
                    acc[4 * s + 0] = fma(-lm2[r], wpSr[r], acc[4 * s + 0]);
                    acc[4 * s + 1] = fma(-lm2[r], wpSi[r], acc[4 * s + 1]);
                    acc[4 * s + 2] = fma(-lp2[r], wmSr[r], acc[4 * s + 2]);
                    acc[4 * s + 3] = fma(-lp2[r], wmSi[r], acc[4 * s + 3]);
                  }
                  const double tp = fma(x[r], ab.a, ab.b);
                  const double tm = fma(x[r], ab.a, -ab.b);
                  const double np_ = fma(tp, lp2[r], -lp1[r]);
                  const double nm_ = fma(tm, lm2[r], -lm1[r]);
                  lp1[r] = lp2[r];
                  lp2[r] = np_;
                  lm1[r] = lm2[r];
                  lm2[r] = nm_;
                }
              }
              warp_reduce16(acc, lane);
              if (lane < 16) red[(g * LEG_GROUP + h * 4) * 4 + lane] = acc[0];
            }
          } else {
#pragma unroll
            for (int s = 0; s < LEG_GROUP; ++s) {
              const Coef ab = cf[s];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const double tp = fma(x[r], ab.a, ab.b);
                const double tm = fma(x[r], ab.a, -ab.b);
                const double np_ = fma(tp, lp2[r], -lp1[r]);
                const double nm_ = fma(tm, lm2[r], -lm1[r]);
                lp1[r] = lp2[r];
                lp2[r] = np_;
                lm1[r] = lm2[r];
                lm2[r] = nm_;
              }
            }
          }
          if (__any_sync(0xffffffffu, mineneg)) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const int p = tp0 + warp * 32 * R + r * 32 + lane;
              if (scp[r] < 0 && is_big(lp2[r])) {
                lp1[r] *= SC_DOWN;
                lp2[r] *= SC_DOWN;
                scp[r] += 1;
                if (scp[r] == 0) load_p(r, p);
              }
              if (scm[r] < 0 && is_big(lm2[r])) {
                lm1[r] *= SC_DOWN;
                lm2[r] *= SC_DOWN;
                scm[r] += 1;
                if (scm[r] == 0) load_m(r, p);
              }
            }
          }
        }
        if (PROD == 2 && issuer) pump.run(cg);
        if (!did) red[g * LEG_GROUP * 4 + lane] = 0.0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      named_bar_sync(1, NW * 32);
      flush_chunk<NW>(s_red[cg & 1], gpart + (size_t)c * LEG_CHUNK * 4, tidc, tile == 0);
    }
  }
}

// ================================================================================================
// adjoint kernels, single-warp CTAs ("_w" variants)
//
// One CTA = one warp = one work item (one m, a contiguous range of ring pairs walked tile by tile,
// tile = 32 R pairs).  No block-level barrier exists: a warp that is still ramping up, or that has
// finished, never holds other warps' registers hostage, and the scheduler backfills with the next
// item.  The 32-lane reduction of the 16 partial sums of every 4-step group goes through a padded
// shared-memory transpose (16 STS.64 + 16 LDS.64 + 16 DADD + 1 shuffle) instead of a select-heavy
// shuffle butterfly.  Reduced values are added into the item's partial-sum array in global memory
// once per 64-step chunk (read-modify-write by the same lane; prefetched at chunk start).
// The coefficient stream is staged by TMA bulk copies issued by lane 0; with a single consumer warp
// a stage is free as soon as the warp has passed it, so only the `full` mbarriers are needed.
// ================================================================================================
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// sum over the 32 lanes of v[k], k < 16, through shared memory; every lane returns the total of value
// (lane & 15)
__device__ __forceinline__ double warp_reduce16_smem(const double (&v)[16], double (*tr)[33], int lane) {
  __syncwarp();  // readers of the previous round are done
#pragma unroll
  for (int k = 0; k < 16; ++k) tr[k][lane] = v[k];
  __syncwarp();
  const int j = lane & 15, h = (lane >> 4) * 16;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int t = 0; t < 16; t += 4) {
    s0 += tr[j][h + t];
    s1 += tr[j][h + t + 1];
    s2 += tr[j][h + t + 2];
    s3 += tr[j][h + t + 3];
  }
  double s = (s0 + s1) + (s2 + s3);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  return s;
}

struct WarpPipe {
  Coef (*s_coef)[LEG_CHUNK];
  uint64_t* full;
  const double* gcoef;
  int nchunks, total;
  __device__ __forceinline__ void issue(int cn) const {
    const int st = cn % NSTAGE;
    mbar_arrive_expect_tx(&full[st], LEG_CHUNK * sizeof(Coef));
    tma_load_1d(&s_coef[st][0], gcoef + (size_t)(cn % nchunks) * LEG_CHUNK * 2, LEG_CHUNK * sizeof(Coef),
                &full[st]);
  }
};

template <int R>
__global__ void __launch_bounds__(32, (R <= 4) ? 16 : (R <= 6 ? 12 : 1)) k_leg2alm_s0_w(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ double s_tr[16][33];
  __shared__ double s_res[LEG_CHUNK * 4];  // reduced sums of the current chunk
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  constexpr int TILE = R * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nn = (int)P.nstep[mi];
  const int nsteps = (nn + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int ntile = (it.np + TILE - 1) / TILE;
  const int lane = threadIdx.x;

  if (lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const WarpPipe pipe{s_coef, s_full, P.coef + (size_t)P.off[mi] * 2, nchunks, nchunks * ntile};
  if (lane == 0)
    for (int cn = 0; cn < NSTAGE - 1 && cn < pipe.total; ++cn) pipe.issue(cn);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double k0 = P.k0[mi];
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  int cg = 0;

  for (int tile = 0; tile < ntile; ++tile) {
    const int tp0 = it.p0 + tile * TILE;
    const int tnp = min(TILE, it.np - tile * TILE);
    double csq[R], lam1[R], lam2[R];
    double er[R], ei[R], orr[R], oi[R];
    int sc[R];
    bool act[R];
    bool any_act = false;

    auto load_in = [&](int r, int p) {
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const double2 gN = leg[leg_offN(P, p, mi) + rowoff];
      const int64_t oS = leg_offS(P, p, mi);
      const double2 gS = oS >= 0 ? leg[oS + rowoff] : make_double2(0.0, 0.0);
      const double cc = P.cth[p];
      er[r] = gN.x + gS.x;
      ei[r] = gN.y + gS.y;
      orr[r] = cc * (gN.x - gS.x);
      oi[r] = cc * (gN.y - gS.y);
    };

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pl = r * 32 + lane;
      const bool valid = pl < tnp;
      const int p = tp0 + (valid ? pl : 0);
      act[r] = valid && (m <= P.mlim[p]);
      any_act |= act[r];
      const double c = P.cth[p];
      csq[r] = c * c;
      double mant;
      int e;
      pow_scaled(P.sth[p], m, mant, e);
      to_scaled(k0 * mant, e, lam2[r], sc[r]);
      if (!act[r]) {
        lam2[r] = 0.0;
        sc[r] = 0;
      }
      lam1[r] = 0.0;
      er[r] = ei[r] = orr[r] = oi[r] = 0.0;
      if (act[r] && sc[r] == 0) load_in(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);
    // warp-uniform range flags, re-voted only when a scale changes (not once per group)
    bool w_any0 = false, w_neg = false;
    auto revote = [&]() {
      bool mine0 = false, mineneg = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        mine0 |= act[r] && (sc[r] == 0);
        mineneg |= (sc[r] < 0);
      }
      w_any0 = __any_sync(0xffffffffu, mine0);
      w_neg = __any_sync(0xffffffffu, mineneg);
    };
    revote();

    for (int c = 0; c < nchunks; ++c, ++cg) {
      const int st = cg % NSTAGE;
      if (lane == 0 && cg + NSTAGE - 1 < pipe.total) {
        fence_proxy_async();  // the stage being refilled was read (generic proxy) during chunk cg-1
        pipe.issue(cg + NSTAGE - 1);
      }
      // partial sums of the previous tiles for this chunk (element k*32 + lane), fetched early
      double* gp = gpart + (size_t)c * LEG_CHUNK * 4 + lane;
      double old[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) old[k] = (tile > 0 || P.accumulate) ? gp[k * 32] : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s_res[k * 32 + lane] = 0.0;  // groups that only ramp contribute 0
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      if (warp_act) {
#pragma unroll 1
        for (int g = 0; g < ng; ++g) {
          {
            const Coef* cf = &s_coef[st][g * LEG_GROUP];
            if (w_any0) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                double acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                  const Coef ab = cf[h * 4 + s];
#pragma unroll
                  for (int r = 0; r < R; ++r) {
                    acc[4 * s + 0] = fma(lam2[r], er[r], acc[4 * s + 0]);
                    acc[4 * s + 1] = fma(lam2[r], ei[r], acc[4 * s + 1]);
                    acc[4 * s + 2] = fma(lam2[r], orr[r], acc[4 * s + 2]);
                    acc[4 * s + 3] = fma(lam2[r], oi[r], acc[4 * s + 3]);
                    const double t = fma(ab.a, csq[r], ab.b);
                    const double nx = fma(t, lam2[r], lam1[r]);
                    lam1[r] = lam2[r];
                    lam2[r] = nx;
                  }
                }
                const double tot = warp_reduce16_smem(acc, s_tr, lane);
                // 4-step group q = 2g + h: lower half-warp keeps even q, upper half odd q
                if (lane < 16) s_res[(g * 2 + h) * 16 + lane] = tot;
              }
            } else {
#pragma unroll
              for (int s = 0; s < LEG_GROUP; ++s) {
                const Coef ab = cf[s];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const double t = fma(ab.a, csq[r], ab.b);
                  const double nx = fma(t, lam2[r], lam1[r]);
                  lam1[r] = lam2[r];
                  lam2[r] = nx;
                }
              }
            }
            if (w_neg) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                if (sc[r] < 0 && is_big(lam2[r])) {
                  lam1[r] *= SC_DOWN;
                  lam2[r] *= SC_DOWN;
                  sc[r] += 1;
                  if (sc[r] == 0) load_in(r, tp0 + r * 32 + lane);
                }
              }
              revote();  // scales changed: refresh the warp-uniform flags
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k * 32] = old[k] + s_res[k * 32 + lane];
      __syncwarp();  // s_res is zeroed again at the top of the next chunk
    }
  }
}

template <int R>
__global__ void __launch_bounds__(32, (R <= 2) ? 16 : (R <= 3 ? 12 : (R <= 4 ? 10 : 1))) k_leg2alm_s2_w(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ double s_tr[16][33];
  __shared__ double s_res[LEG_CHUNK * 4];  // reduced sums of the current chunk
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  constexpr int TILE = R * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nl = (int)P.nstep[mi];
  const int nsteps = (nl + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int ntile = (it.np + TILE - 1) / TILE;
  const int lane = threadIdx.x;

  if (lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const WarpPipe pipe{s_coef, s_full, P.coef + (size_t)P.off[mi] * 2, nchunks, nchunks * ntile};
  if (lane == 0)
    for (int cn = 0; cn < NSTAGE - 1 && cn < pipe.total; ++cn) pipe.issue(cn);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double kp = P.k0[mi], km = P.k1[mi];
  const int pw = m >= 2 ? m - 2 : 2 - m;
  const int q = m < 2 ? m : 2;
  const int l0 = m > 2 ? m : 2;
  const double sgnS = ((l0 + m) & 1) ? -1.0 : 1.0;
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  int cg = 0;

  for (int tile = 0; tile < ntile; ++tile) {
    const int tp0 = it.p0 + tile * TILE;
    const int tnp = min(TILE, it.np - tile * TILE);
    double x[R], lp1[R], lp2[R], lm1[R], lm2[R];
    double wpNr[R], wpNi[R], wmSr[R], wmSi[R], wmNr[R], wmNi[R], wpSr[R], wpSi[R];
    int scp[R], scm[R];
    bool act[R];
    bool any_act = false;

    auto load_p = [&](int r, int p) {  // inputs multiplied by d+
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const int64_t cs = leg_cstr(P, p, mi);
      const double2 qN = leg[leg_offN(P, p, mi) + rowoff], uN = leg[leg_offN(P, p, mi) + rowoff + cs];
      wpNr[r] = qN.x - uN.y;
      wpNi[r] = qN.y + uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS + rowoff], uS = leg[oS + rowoff + cs];
        wmSr[r] = sgnS * (qS.x + uS.y);
        wmSi[r] = sgnS * (qS.y - uS.x);
      } else {
        wmSr[r] = wmSi[r] = 0.0;
      }
    };
    auto load_m = [&](int r, int p) {  // inputs multiplied by d-
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const int64_t cs = leg_cstr(P, p, mi);
      const double2 qN = leg[leg_offN(P, p, mi) + rowoff], uN = leg[leg_offN(P, p, mi) + rowoff + cs];
      wmNr[r] = qN.x + uN.y;
      wmNi[r] = qN.y - uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS + rowoff], uS = leg[oS + rowoff + cs];
        wpSr[r] = sgnS * (qS.x - uS.y);
        wpSi[r] = sgnS * (qS.y + uS.x);
      } else {
        wpSr[r] = wpSi[r] = 0.0;
      }
    };

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pl = r * 32 + lane;
      const bool valid = pl < tnp;
      const int p = tp0 + (valid ? pl : 0);
      act[r] = valid && (m <= P.mlim[p]);
      any_act |= act[r];
      x[r] = P.cth[p];
      double mant;
      int e;
      pow_scaled(P.sth[p], pw, mant, e);
      double fs = 1.0, fc = 1.0;
      const double s2 = P.sh[p] * P.sh[p], c2 = P.ch[p] * P.ch[p];
      for (int k = 0; k < q; ++k) {
        fs *= s2;
        fc *= c2;
      }
      to_scaled(kp * mant * fs, e, lp2[r], scp[r]);
      to_scaled(km * mant * fc, e, lm2[r], scm[r]);
      if (!act[r]) {
        lp2[r] = lm2[r] = 0.0;
        scp[r] = scm[r] = 0;
      }
      lp1[r] = lm1[r] = 0.0;
      wpNr[r] = wpNi[r] = wmSr[r] = wmSi[r] = 0.0;
      wmNr[r] = wmNi[r] = wpSr[r] = wpSi[r] = 0.0;
      if (act[r] && scp[r] == 0) load_p(r, p);
      if (act[r] && scm[r] == 0) load_m(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);
    // warp-uniform range flags, re-voted only when a scale changes (not once per group)
    bool w_any0 = false, w_neg = false;
    auto revote = [&]() {
      bool mine0 = false, mineneg = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        mine0 |= act[r] && ((scp[r] == 0) || (scm[r] == 0));
        mineneg |= (scp[r] < 0) || (scm[r] < 0);
      }
      w_any0 = __any_sync(0xffffffffu, mine0);
      w_neg = __any_sync(0xffffffffu, mineneg);
    };
    revote();

    for (int c = 0; c < nchunks; ++c, ++cg) {
      const int st = cg % NSTAGE;
      if (lane == 0 && cg + NSTAGE - 1 < pipe.total) {
        fence_proxy_async();
        pipe.issue(cg + NSTAGE - 1);
      }
      double* gp = gpart + (size_t)c * LEG_CHUNK * 4 + lane;
      double old[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) old[k] = (tile > 0 || P.accumulate) ? gp[k * 32] : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s_res[k * 32 + lane] = 0.0;  // groups that only ramp contribute 0
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      if (warp_act) {
#pragma unroll 1
        for (int g = 0; g < ng; ++g) {
          {
            const Coef* cf = &s_coef[st][g * LEG_GROUP];
            if (w_any0) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                double acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                  const Coef ab = cf[h * 4 + s];
#pragma unroll
                  for (int r = 0; r < R; ++r) {
                    acc[4 * s + 0] = fma(lp2[r], wpNr[r], acc[4 * s + 0]);
                    acc[4 * s + 1] = fma(lp2[r], wpNi[r], acc[4 * s + 1]);
                    acc[4 * s + 2] = fma(lm2[r], wmNr[r], acc[4 * s + 2]);
                    acc[4 * s + 3] = fma(lm2[r], wmNi[r], acc[4 * s + 3]);
                    if ((s & 1) == 0) {
                      acc[4 * s + 0] = fma(lm2[r], wpSr[r], acc[4 * s + 0]);
                      acc[4 * s + 1] = fma(lm2[r], wpSi[r], acc[4 * s + 1]);
                      acc[4 * s + 2] = fma(lp2[r], wmSr[r], acc[4 * s + 2]);
                      acc[4 * s + 3] = fma(lp2[r], wmSi[r], acc[4 * s + 3]);
                    } else {
                      acc[4 * s + 0] = fma(-lm2[r], wpSr[r], acc[4 * s + 0]);
                      acc[4 * s + 1] = fma(-lm2[r], wpSi[r], acc[4 * s + 1]);
                      acc[4 * s + 2] = fma(-lp2[r], wmSr[r], acc[4 * s + 2]);
                      acc[4 * s + 3] = fma(-lp2[r], wmSi[r], acc[4 * s + 3]);
                    }
                    const double tp = fma(x[r], ab.a, ab.b);
                    const double tm = fma(x[r], ab.a, -ab.b);
                    const double np_ = fma(tp, lp2[r], -lp1[r]);
                    const double nm_ = fma(tm, lm2[r], -lm1[r]);
                    lp1[r] = lp2[r];
                    lp2[r] = np_;
                    lm1[r] = lm2[r];
                    lm2[r] = nm_;
                  }
                }
                const double tot = warp_reduce16_smem(acc, s_tr, lane);
                if (lane < 16) s_res[(g * 2 + h) * 16 + lane] = tot;
              }
            } else {
#pragma unroll
              for (int s = 0; s < LEG_GROUP; ++s) {
                const Coef ab = cf[s];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const double tp = fma(x[r], ab.a, ab.b);
                  const double tm = fma(x[r], ab.a, -ab.b);
                  const double np_ = fma(tp, lp2[r], -lp1[r]);
                  const double nm_ = fma(tm, lm2[r], -lm1[r]);
                  lp1[r] = lp2[r];
                  lp2[r] = np_;
                  lm1[r] = lm2[r];
                  lm2[r] = nm_;
                }
              }
            }
            if (w_neg) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const int p = tp0 + r * 32 + lane;
                if (scp[r] < 0 && is_big(lp2[r])) {
                  lp1[r] *= SC_DOWN;
                  lp2[r] *= SC_DOWN;
                  scp[r] += 1;
                  if (scp[r] == 0) load_p(r, p);
                }
                if (scm[r] < 0 && is_big(lm2[r])) {
                  lm1[r] *= SC_DOWN;
                  lm2[r] *= SC_DOWN;
                  scm[r] += 1;
                  if (scm[r] == 0) load_m(r, p);
                }
              }
              revote();  // scales changed: refresh the warp-uniform flags
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k * 32] = old[k] + s_res[k * 32 + lane];
      __syncwarp();  // s_res is zeroed again at the top of the next chunk
    }
  }
}

// Multi-warp variant of the single-warp adjoint: the NW warps of a CTA walk NW adjacent tiles of the same m in
// lockstep (one CTA barrier per 64-step chunk), share one coefficient stream and add their per-chunk sums in
// shared memory before the global partial-sum update -- NW times less partial-sum and coefficient traffic.
template <int R, int NW>
__global__ void __launch_bounds__(NW * 32, 12 / NW) k_leg2alm_s2_mw(LegParams P) {
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ double s_tr_all[NW][16][33];
  __shared__ double s_res_all[2][NW][LEG_CHUNK * 4];  // per warp: reduced sums of the current chunk (double-buffered)
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  constexpr int TILE = R * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nl = (int)P.nstep[mi];
  const int nsteps = (nl + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int ntile = (it.np + TILE - 1) / TILE;
  const int nround = (ntile + NW - 1) / NW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  double (*s_tr)[33] = s_tr_all[warp];

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const WarpPipe pipe{s_coef, s_full, P.coef + (size_t)P.off[mi] * 2, nchunks, nchunks * nround};
  if (tid == 0)
    for (int cn = 0; cn < NSTAGE && cn < pipe.total; ++cn) pipe.issue(cn);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double kp = P.k0[mi], km = P.k1[mi];
  const int pw = m >= 2 ? m - 2 : 2 - m;
  const int q = m < 2 ? m : 2;
  const int l0 = m > 2 ? m : 2;
  const double sgnS = ((l0 + m) & 1) ? -1.0 : 1.0;
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  int cg = 0;

  int buf = 0;
  for (int round = 0; round < nround; ++round) {
    const int tile = round * NW + warp;
    const int tp0 = it.p0 + min(tile, ntile - 1) * TILE;
    const int tnp = tile < ntile ? min(TILE, it.np - tile * TILE) : 0;  // 0: this warp idles through the round
    double x[R], lp1[R], lp2[R], lm1[R], lm2[R];
    double wpNr[R], wpNi[R], wmSr[R], wmSi[R], wmNr[R], wmNi[R], wpSr[R], wpSi[R];
    int scp[R], scm[R];
    bool act[R];
    bool any_act = false;

    auto load_p = [&](int r, int p) {  // inputs multiplied by d+
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const int64_t cs = leg_cstr(P, p, mi);
      const double2 qN = leg[leg_offN(P, p, mi) + rowoff], uN = leg[leg_offN(P, p, mi) + rowoff + cs];
      wpNr[r] = qN.x - uN.y;
      wpNi[r] = qN.y + uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS + rowoff], uS = leg[oS + rowoff + cs];
        wmSr[r] = sgnS * (qS.x + uS.y);
        wmSi[r] = sgnS * (qS.y - uS.x);
      } else {
        wmSr[r] = wmSi[r] = 0.0;
      }
    };
    auto load_m = [&](int r, int p) {  // inputs multiplied by d-
      const int64_t rowoff = leg_rowoff(P, p, mi);
      const int64_t cs = leg_cstr(P, p, mi);
      const double2 qN = leg[leg_offN(P, p, mi) + rowoff], uN = leg[leg_offN(P, p, mi) + rowoff + cs];
      wmNr[r] = qN.x + uN.y;
      wmNi[r] = qN.y - uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS + rowoff], uS = leg[oS + rowoff + cs];
        wpSr[r] = sgnS * (qS.x - uS.y);
        wpSi[r] = sgnS * (qS.y + uS.x);
      } else {
        wpSr[r] = wpSi[r] = 0.0;
      }
    };

#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int pl = r * 32 + lane;
      const bool valid = pl < tnp;
      const int p = tp0 + (valid ? pl : 0);
      act[r] = valid && (m <= P.mlim[p]);
      any_act |= act[r];
      x[r] = P.cth[p];
      double mant;
      int e;
      pow_scaled(P.sth[p], pw, mant, e);
      double fs = 1.0, fc = 1.0;
      const double s2 = P.sh[p] * P.sh[p], c2 = P.ch[p] * P.ch[p];
      for (int k = 0; k < q; ++k) {
        fs *= s2;
        fc *= c2;
      }
      to_scaled(kp * mant * fs, e, lp2[r], scp[r]);
      to_scaled(km * mant * fc, e, lm2[r], scm[r]);
      if (!act[r]) {
        lp2[r] = lm2[r] = 0.0;
        scp[r] = scm[r] = 0;
      }
      lp1[r] = lm1[r] = 0.0;
      wpNr[r] = wpNi[r] = wmSr[r] = wmSi[r] = 0.0;
      wmNr[r] = wmNi[r] = wpSr[r] = wpSi[r] = 0.0;
      if (act[r] && scp[r] == 0) load_p(r, p);
      if (act[r] && scm[r] == 0) load_m(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);
    // warp-uniform range flags, re-voted only when a scale changes (not once per group)
    bool w_any0 = false, w_neg = false;
    auto revote = [&]() {
      bool mine0 = false, mineneg = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        mine0 |= act[r] && ((scp[r] == 0) || (scm[r] == 0));
        mineneg |= (scp[r] < 0) || (scm[r] < 0);
      }
      w_any0 = __any_sync(0xffffffffu, mine0);
      w_neg = __any_sync(0xffffffffu, mineneg);
    };
    revote();

    for (int c = 0; c < nchunks; ++c, ++cg) {
      const int st = cg % NSTAGE;
      double* s_res = s_res_all[buf][warp];
      // the CTA's NW*32 threads share the 256 doubles of the chunk's partial sums: NE each
      constexpr int NE = LEG_CHUNK * 4 / (NW * 32);
      double* gp = gpart + (size_t)c * LEG_CHUNK * 4 + tid;
      double old[NE];
#pragma unroll
      for (int k = 0; k < NE; ++k) old[k] = (round > 0 || P.accumulate) ? gp[k * NW * 32] : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s_res[k * 32 + lane] = 0.0;  // groups that only ramp contribute 0
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      if (warp_act) {
#pragma unroll 1
        for (int g = 0; g < ng; ++g) {
          {
            const Coef* cf = &s_coef[st][g * LEG_GROUP];
            if (w_any0) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                double acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                  const Coef ab = cf[h * 4 + s];
#pragma unroll
                  for (int r = 0; r < R; ++r) {
                    acc[4 * s + 0] = fma(lp2[r], wpNr[r], acc[4 * s + 0]);
                    acc[4 * s + 1] = fma(lp2[r], wpNi[r], acc[4 * s + 1]);
                    acc[4 * s + 2] = fma(lm2[r], wmNr[r], acc[4 * s + 2]);
                    acc[4 * s + 3] = fma(lm2[r], wmNi[r], acc[4 * s + 3]);
                    if ((s & 1) == 0) {
                      acc[4 * s + 0] = fma(lm2[r], wpSr[r], acc[4 * s + 0]);
                      acc[4 * s + 1] = fma(lm2[r], wpSi[r], acc[4 * s + 1]);
                      acc[4 * s + 2] = fma(lp2[r], wmSr[r], acc[4 * s + 2]);
                      acc[4 * s + 3] = fma(lp2[r], wmSi[r], acc[4 * s + 3]);
                    } else {
                      acc[4 * s + 0] = fma(-lm2[r], wpSr[r], acc[4 * s + 0]);
                      acc[4 * s + 1] = fma(-lm2[r], wpSi[r], acc[4 * s + 1]);
                      acc[4 * s + 2] = fma(-lp2[r], wmSr[r], acc[4 * s + 2]);
                      acc[4 * s + 3] = fma(-lp2[r], wmSi[r], acc[4 * s + 3]);
                    }
                    const double tp = fma(x[r], ab.a, ab.b);
                    const double tm = fma(x[r], ab.a, -ab.b);
                    const double np_ = fma(tp, lp2[r], -lp1[r]);
                    const double nm_ = fma(tm, lm2[r], -lm1[r]);
                    lp1[r] = lp2[r];
                    lp2[r] = np_;
                    lm1[r] = lm2[r];
                    lm2[r] = nm_;
                  }
                }
                const double tot = warp_reduce16_smem(acc, s_tr, lane);
                if (lane < 16) s_res[(g * 2 + h) * 16 + lane] = tot;
              }
            } else {
#pragma unroll
              for (int s = 0; s < LEG_GROUP; ++s) {
                const Coef ab = cf[s];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const double tp = fma(x[r], ab.a, ab.b);
                  const double tm = fma(x[r], ab.a, -ab.b);
                  const double np_ = fma(tp, lp2[r], -lp1[r]);
                  const double nm_ = fma(tm, lm2[r], -lm1[r]);
                  lp1[r] = lp2[r];
                  lp2[r] = np_;
                  lm1[r] = lm2[r];
                  lm2[r] = nm_;
                }
              }
            }
            if (w_neg) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const int p = tp0 + r * 32 + lane;
                if (scp[r] < 0 && is_big(lp2[r])) {
                  lp1[r] *= SC_DOWN;
                  lp2[r] *= SC_DOWN;
                  scp[r] += 1;
                  if (scp[r] == 0) load_p(r, p);
                }
                if (scm[r] < 0 && is_big(lm2[r])) {
                  lm1[r] *= SC_DOWN;
                  lm2[r] *= SC_DOWN;
                  scm[r] += 1;
                  if (scm[r] == 0) load_m(r, p);
                }
              }
              revote();  // scales changed: refresh the warp-uniform flags
            }
          }
        }
      }
      __syncthreads();  // every warp's sums of this chunk are in shared memory; the coefficient stage is free
      if (tid == 0 && cg + NSTAGE < pipe.total) {
        fence_proxy_async();  // the stage was read through the generic proxy
        pipe.issue(cg + NSTAGE);
      }
#pragma unroll
      for (int k = 0; k < NE; ++k) {
        double v = old[k];
#pragma unroll
        for (int w = 0; w < NW; ++w) v += s_res_all[buf][w][k * NW * 32 + tid];
        gp[k * NW * 32] = v;
      }
      buf ^= 1;  // the other buffer was last read before the barrier above
    }
  }
}

// ================================================================================================
// prep / post kernels (HBM-bound, O(nalm))
// ================================================================================================
// spin 0 prep: stream[off+n] = { (eps_{l+1} a_l + eps_{l+2} a_{l+2}) / s_n , a_{l+1} / s_n }, l = m+2n
__device__ __forceinline__ double eps_lm(int64_t l, int64_t m) {
  return sqrt((double)(l * l - m * m) / (double)(4 * l * l - 1));
}

__global__ void k_prep_s0(const double2* __restrict__ alm, const int64_t* __restrict__ mval,
                          const int64_t* __restrict__ mstart, const int64_t* __restrict__ nstep,
                          const int64_t* __restrict__ off, const double* __restrict__ invs,
                          int64_t lmax, Strm* __restrict__ stream, int mi0) {
  const int mi = mi0 + blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nstep[mi]) return;
  const int64_t m = mval[mi];
  const int64_t l = m + 2 * n;
  const double2* a = alm + mstart[mi];
  const double2 al = a[l];
  const double2 al1 = (l + 1 <= lmax) ? a[l + 1] : make_double2(0.0, 0.0);
  const double2 al2 = (l + 2 <= lmax) ? a[l + 2] : make_double2(0.0, 0.0);
  const double e1 = eps_lm(l + 1, m), e2 = eps_lm(l + 2, m);
  const double is = invs[off[mi] + n];
  Strm o;
  o.x = fma(e1, al.x, e2 * al2.x) * is;
  o.y = fma(e1, al.y, e2 * al2.y) * is;
  o.z = al1.x * is;
  o.w = al1.y * is;
  stream[off[mi] + n] = o;
}

// spin 2 prep: alpha+ = -(E + iB) sigma/2 ; alpha- = -(E - iB) sigma/2
__global__ void k_prep_s2(const double2* __restrict__ almE, const double2* __restrict__ almB,
                          const int64_t* __restrict__ mval, const int64_t* __restrict__ mstart,
                          const int64_t* __restrict__ nstep, const int64_t* __restrict__ off,
                          const double* __restrict__ sig, Strm* __restrict__ stream, int mi0) {
  const int mi = mi0 + blockIdx.y;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nstep[mi]) return;
  const int64_t m = mval[mi];
  const int64_t l = (m > 2 ? m : 2) + j;
  const double2 e = almE[mstart[mi] + l];
  const double2 b = almB[mstart[mi] + l];
  const double h = -0.5 * sig[off[mi] + j];
  Strm o;
  // E + iB = (er - bi, ei + br) ; E - iB = (er + bi, ei - br)
  o.x = h * (e.x - b.y);
  o.y = h * (e.y + b.x);
  o.z = h * (e.x + b.y);
  o.w = h * (e.y - b.x);
  stream[off[mi] + j] = o;
}

// adjoint: sum the per-tile partials of one m (deterministic order)
__global__ void k_reduce_part(const Strm* __restrict__ part, const int64_t* __restrict__ partoff,
                              const int* __restrict__ afirst, const int* __restrict__ acount,
                              const int64_t* __restrict__ nstep, const int64_t* __restrict__ off,
                              Strm* __restrict__ accum, int mi0) {
  const int mi = mi0 + blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nstep[mi]) return;
  Strm s = {0.0, 0.0, 0.0, 0.0};
  const int f = afirst[mi], cnt = acount[mi];
  for (int i = 0; i < cnt; ++i) {
    const Strm v = part[partoff[f + i] + n];
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
  }
  accum[off[mi] + n] = s;
}

// spin 0 post: a_l = eps_{l+1} At_n/s_n + eps_l At_{n-1}/s_{n-1} (l = m+2n) ; a_{l+1} = Bt_n/s_n
__global__ void k_post_s0(const Strm* __restrict__ accum, const int64_t* __restrict__ mval,
                          const int64_t* __restrict__ mstart, const int64_t* __restrict__ nstep,
                          const int64_t* __restrict__ off, const double* __restrict__ invs,
                          int64_t lmax, double2* __restrict__ alm, int mi0) {
  const int mi = mi0 + blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nstep[mi]) return;
  const int64_t m = mval[mi];
  const int64_t l = m + 2 * n;
  const Strm cur = accum[off[mi] + n];
  const double is = invs[off[mi] + n];
  const double e1 = eps_lm(l + 1, m);
  double vr = e1 * is * cur.x, vi = e1 * is * cur.y;
  if (n > 0) {
    const Strm prv = accum[off[mi] + n - 1];
    const double isp = invs[off[mi] + n - 1];
    const double e0 = eps_lm(l, m);
    vr = fma(e0 * isp, prv.x, vr);
    vi = fma(e0 * isp, prv.y, vi);
  }
  double2* a = alm + mstart[mi];
  a[l] = make_double2(vr, vi);
  if (l + 1 <= lmax) a[l + 1] = make_double2(is * cur.z, is * cur.w);
}

// spin 2 post: E = -sigma/2 (G+ + G-) ; B = i sigma/2 (G+ - G-) ; l < 2 entries zeroed
__global__ void k_post_s2(const Strm* __restrict__ accum, const int64_t* __restrict__ mval,
                          const int64_t* __restrict__ mstart, const int64_t* __restrict__ nstep,
                          const int64_t* __restrict__ off, const double* __restrict__ sig,
                          int64_t lmax, double2* __restrict__ almE, double2* __restrict__ almB, int mi0) {
  const int mi = mi0 + blockIdx.y;
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = mval[mi];
  if (j == 0 && blockIdx.x == 0) {
    for (int64_t l = m; l < 2 && l <= lmax; ++l) {
      almE[mstart[mi] + l] = make_double2(0.0, 0.0);
      almB[mstart[mi] + l] = make_double2(0.0, 0.0);
    }
  }
  if (j >= nstep[mi]) return;
  const int64_t l = (m > 2 ? m : 2) + j;
  const Strm g = accum[off[mi] + j];
  const double h = 0.5 * sig[off[mi] + j];
  almE[mstart[mi] + l] = make_double2(-h * (g.x + g.z), -h * (g.y + g.w));
  // i * (dr, di) = (-di, dr)
  almB[mstart[mi] + l] = make_double2(-h * (g.y - g.w), h * (g.x - g.z));
}

// ================================================================================================
// host side
// ================================================================================================
std::vector<PairHost> build_pairs(int64_t lmax, int64_t nring, const double* theta) {
  std::vector<int64_t> idx((size_t)nring);
  for (int64_t i = 0; i < nring; ++i) idx[(size_t)i] = i;
  std::stable_sort(idx.begin(), idx.end(),
                   [&](int64_t a, int64_t b) { return theta[a] < theta[b]; });
  const double pi = 3.14159265358979323846;
  auto mk = [&](int64_t n, int64_t s) {
    PairHost p;
    const long double th = (long double)theta[n];
    p.cth = (double)cosl(th);
    p.sth = (double)sinl(th);
    p.sh = (double)sinl(0.5L * th);
    p.ch = (double)cosl(0.5L * th);
    p.slotN = n;
    p.slotS = s;
    p.mlim0 = (int)mlim_rule(lmax, 0, p.sth, p.cth);
    p.mlim2 = (int)mlim_rule(lmax, 2, p.sth, p.cth);
    return p;
  };
  std::vector<PairHost> out;
  out.reserve((size_t)nring / 2 + 1);
  int64_t lo = 0, hi = nring - 1;
  while (lo <= hi) {
    const int64_t a = idx[(size_t)lo], b = idx[(size_t)hi];
    if (lo < hi && std::fabs(theta[a] + theta[b] - pi) <= 1e-14 * pi && theta[a] < 0.5 * pi) {
      out.push_back(mk(a, b));
      ++lo;
      --hi;
    } else if (std::fabs(theta[a] - 0.5 * pi) >= std::fabs(theta[b] - 0.5 * pi)) {
      out.push_back(mk(a, -1));
      ++lo;
    } else {
      out.push_back(mk(b, -1));
      --hi;
    }
  }
  return out;
}

LegendreStage::~LegendreStage() {
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
}

double LegendreStage::main_kernel_ms() const {
  float f = 0.f;
  if (ev0_ && ev1_ && cudaEventElapsedTime(&f, ev0_, ev1_) != cudaSuccess) {
    cudaGetLastError();
    f = 0.f;
  }
  return (double)f;
}

namespace {
// kernel configurations
// synthesis kernel variants (R pairs per thread, NW consumer warps); selectable at run time for tuning
struct SynCfg { int R, NW; };
// tile shapes instantiated for the multi-warp kernels: R ring pairs per thread x NW consumer warps
// (round-1 tuning, tools/tune.py: larger R, fewer warps and single-warp synthesis CTAs were all slower)
static const SynCfg kSynCfgs[] = {{2, 4}, {4, 4}, {2, 8}, {3, 4}};
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static int cfg_clip(int i) { return i < 0 ? 0 : (i > 3 ? 3 : i); }
static int s0_cfg() { return cfg_clip(env_int("HPSHT_S0_CFG", 1)); }  // default <4,4>: 512 pairs / tile
static int s2_cfg() { return cfg_clip(env_int("HPSHT_S2_CFG", 0)); }  // default <2,4>: 256 pairs / tile
static int env_syn_prod() { return env_int("HPSHT_SYN_PROD", 1); }  // 1 dedicated producer warp, 2 polling
static int env_adj_prod() { return env_int("HPSHT_ADJ_PROD", 2); }
#define HP_LAUNCH_P(KERNEL, PROD, cfgidx, nitems, stream, P)                                          \
  do {                                                                                                 \
    const int xw = (PROD == 1) ? 32 : 0;                                                               \
    switch (cfgidx) {                                                                                  \
      case 0: KERNEL<2, 4, PROD><<<(unsigned)(nitems), 4 * 32 + xw, 0, stream>>>(P); break;            \
      case 1: KERNEL<4, 4, PROD><<<(unsigned)(nitems), 4 * 32 + xw, 0, stream>>>(P); break;            \
      case 2: KERNEL<2, 8, PROD><<<(unsigned)(nitems), 8 * 32 + xw, 0, stream>>>(P); break;            \
      default: KERNEL<3, 4, PROD><<<(unsigned)(nitems), 4 * 32 + xw, 0, stream>>>(P); break;           \
    }                                                                                                  \
  } while (0)
#define HP_LAUNCH(KERNEL, prod, cfgidx, nitems, stream, P)                                            \
  do {                                                                                                 \
    if ((prod) == 1) HP_LAUNCH_P(KERNEL, 1, cfgidx, nitems, stream, P);                                \
    else HP_LAUNCH_P(KERNEL, 2, cfgidx, nitems, stream, P);                                            \
  } while (0)
static int a0_cfg() { return cfg_clip(env_int("HPSHT_A0_CFG", 1)); }  // block-mode adjoint <4,4>
static int a2_cfg() { return cfg_clip(env_int("HPSHT_A2_CFG", 0)); }  // block-mode adjoint <2,4>
// adjoint kernel family: 1 = single-warp CTAs (k_leg2alm_*_w), 0 = 4-warp CTAs with a chunk barrier
static int adj_warp_mode() { return env_int("HPSHT_ADJ_WARP", 1); }
static int a0w_R() { return env_int("HPSHT_A0W_R", 8); }
static int a2w_R() { return env_int("HPSHT_A2W_R", 4); }
#define HP_LAUNCH_W(KERNEL, Rval, nitems, stream, P)                                                  \
  do {                                                                                                 \
    switch (Rval) {                                                                                    \
      case 2: KERNEL<2><<<(unsigned)(nitems), 32, 0, stream>>>(P); break;                              \
      case 3: KERNEL<3><<<(unsigned)(nitems), 32, 0, stream>>>(P); break;                              \
      case 4: KERNEL<4><<<(unsigned)(nitems), 32, 0, stream>>>(P); break;                              \
      case 6: KERNEL<6><<<(unsigned)(nitems), 32, 0, stream>>>(P); break;                              \
      default: KERNEL<8><<<(unsigned)(nitems), 32, 0, stream>>>(P); break;                             \
    }                                                                                                  \
  } while (0)
// warps per CTA of the spin-2 adjoint: 1 = k_leg2alm_s2_w (default, fastest), 2 / 4 = k_leg2alm_s2_mw<4, NW>:
// NW times less partial-sum DRAM traffic for 3 % (NW = 2) / 11 % (NW = 4) more time (nside 2048, round 1)
static int adj_nw() { return env_int("HPSHT_ADJ_NW", 1); }
#define HP_LAUNCH_MW(KERNEL, NWval, nitems, stream, P)                                                 \
  do {                                                                                                 \
    if ((NWval) == 2) KERNEL<4, 2><<<(unsigned)(nitems), 64, 0, stream>>>(P);                          \
    else KERNEL<4, 4><<<(unsigned)(nitems), 128, 0, stream>>>(P);                                      \
  } while (0)
static int adj_tile(int spin) {
  if (adj_warp_mode()) return 32 * (spin == 0 ? a0w_R() : a2w_R());
  const SynCfg c = kSynCfgs[spin == 0 ? a0_cfg() : a2_cfg()];
  return c.R * c.NW * 32;
}

// Work list: for every local m (ascending == descending cost; CTAs are handed out in blockIdx
// order) the active tiles, i.e. those holding at least one pair with mlim >= m.  Pairs are sorted
// pole -> equator and mlim grows towards the equator, so the active pairs of an m form a suffix.
//   groups == 0 : one item per active tile (synthesis)
//   groups  > 0 : the active tile range of each m is split into <= groups contiguous multi-tile
//                 items (adjoint; each item owns one partial-sum buffer)
void make_items(const std::vector<PairHost>& pairs, const std::vector<int64_t>& mval,
                const std::vector<int64_t>& nstep, bool spin2, int tile, int groups,
                std::vector<WorkItem>& items, std::vector<int>* first, std::vector<int>* count) {
  const int np = (int)pairs.size();
  const int ntile = (np + tile - 1) / tile;
  std::vector<int> tmax((size_t)ntile, -1);
  for (int t = 0; t < ntile; ++t)
    for (int p = t * tile; p < std::min(np, (t + 1) * tile); ++p)
      tmax[(size_t)t] = std::max(tmax[(size_t)t], spin2 ? pairs[(size_t)p].mlim2 : pairs[(size_t)p].mlim0);
  items.clear();
  if (first) first->assign(mval.size(), 0);
  if (count) count->assign(mval.size(), 0);
  for (size_t mi = 0; mi < mval.size(); ++mi) {
    if (first) (*first)[mi] = (int)items.size();
    if (nstep[mi] <= 0) continue;
    std::vector<int> active;
    for (int t = ntile - 1; t >= 0; --t)
      if (tmax[(size_t)t] >= mval[mi]) active.push_back(t);
    if (groups == 0) {
      for (int t : active) {
        WorkItem w;
        w.mi = (int)mi;
        w.p0 = t * tile;
        w.np = std::min(np, (t + 1) * tile) - t * tile;
        w.part = 0;
        items.push_back(w);
      }
    } else if (!active.empty()) {
      // active tiles are not necessarily contiguous for exotic theta lists; cover [tmin, tmax]
      const int tlo = *std::min_element(active.begin(), active.end());
      const int thi = *std::max_element(active.begin(), active.end());
      const int nt = thi - tlo + 1;
      const int G = std::min(groups, nt);
      for (int g = 0; g < G; ++g) {
        const int ta = tlo + (int)((int64_t)nt * g / G), tb = tlo + (int)((int64_t)nt * (g + 1) / G);
        WorkItem w;
        w.mi = (int)mi;
        w.p0 = ta * tile;
        w.np = std::min(np, tb * tile) - ta * tile;
        w.part = 0;
        items.push_back(w);
      }
    }
    if (count) (*count)[mi] = (int)items.size() - (*first)[mi];
  }
}
}  // namespace

void LegendreStage::setup(int64_t lmax, const int64_t* mval, const int64_t* mstart, int64_t nm,
                          int64_t nring, const double* theta, const LegLayout& layout, int max_ncomp,
                          cudaStream_t stream) {
  lmax_ = lmax;
  nm_ = nm;
  nring_ = nring;
  max_ncomp_ = max_ncomp;
  mval_.assign(mval, mval + nm);
  mstart_.assign(mstart, mstart + nm);
  layout_ = layout;
  HP_REQUIRE((int64_t)layout.a0.size() == nring && (int64_t)layout.nr.size() == nring &&
                 (int64_t)layout.j.size() == nring,
             HPSHT_ERR_INTERNAL, "layout size mismatch");
  // m-slabs: default = one slab holding every local m (the reference's (nm, nring, ncomp) array)
  if (layout_.sstart.empty()) {
    layout_.sstart.assign((size_t)nm, 0);
    layout_.nms.assign((size_t)nm, nm);
    layout_.ncs = max_ncomp;
  }
  HP_REQUIRE((int64_t)layout_.sstart.size() == nm && (int64_t)layout_.nms.size() == nm, HPSHT_ERR_INTERNAL,
             "slab table size mismatch");
  slab_first_mi_.clear();
  for (int64_t i = 0; i < nm; ++i)
    if (i == 0 || layout_.sstart[(size_t)i] != layout_.sstart[(size_t)i - 1]) slab_first_mi_.push_back((int)i);
  slab_first_mi_.push_back((int)nm);
  pairs_ = build_pairs(lmax, nring, theta);
  npairs_ = (int64_t)pairs_.size();
  HP_REQUIRE(npairs_ < (1ll << 30) && nm < (1ll << 30), HPSHT_ERR_UNSUPPORTED, "problem too large");

  const size_t np = (size_t)npairs_;
  std::vector<double> cth(np), sth(np), sh(np), ch(np);
  std::vector<int> ml0(np), ml2(np);
  std::vector<int64_t> a0(np), nrr(np), jN(np), jS(np);
  culled0_ = culled2_ = dense_ = 0;
  for (size_t p = 0; p < np; ++p) {
    const PairHost& q = pairs_[p];
    cth[p] = q.cth;
    sth[p] = q.sth;
    sh[p] = q.sh;
    ch[p] = q.ch;
    ml0[p] = q.mlim0;
    ml2[p] = q.mlim2;
    a0[p] = layout.a0[(size_t)q.slotN];
    nrr[p] = layout.nr[(size_t)q.slotN];
    jN[p] = layout.j[(size_t)q.slotN];
    if (q.slotS >= 0) {
      HP_REQUIRE(layout.a0[(size_t)q.slotS] == a0[p] && layout.nr[(size_t)q.slotS] == nrr[p],
                 HPSHT_ERR_INTERNAL, "N/S rings of a pair must live in the same block");
      jS[p] = layout.j[(size_t)q.slotS];
    } else {
      jS[p] = -1;
    }
    for (int64_t i = 0; i < nm; ++i) {
      const double w = (double)(lmax - mval[i] + 1);
      if (w <= 0) continue;
      dense_ += w;
      if (mval[i] <= q.mlim0) culled0_ += w;
      if (mval[i] <= q.mlim2) culled2_ += w;
    }
  }
  d_cth_.upload(cth.data(), np, stream);
  d_sth_.upload(sth.data(), np, stream);
  d_sh_.upload(sh.data(), np, stream);
  d_ch_.upload(ch.data(), np, stream);
  d_mlim0_.upload(ml0.data(), np, stream);
  d_mlim2_.upload(ml2.data(), np, stream);
  d_a0_.upload(a0.data(), np, stream);
  d_nrr_.upload(nrr.data(), np, stream);
  d_jN_.upload(jN.data(), np, stream);
  d_jS_.upload(jS.data(), np, stream);
  d_sstart_.upload(layout_.sstart.data(), (size_t)nm, stream);
  d_nms_.upload(layout_.nms.data(), (size_t)nm, stream);
  d_mval_.upload(mval_.data(), mval_.size(), stream);
  d_mstart_.upload(mstart_.data(), mstart_.size(), stream);
  HP_CUDA(cudaStreamSynchronize(stream));
  if (!ev0_) HP_CUDA(cudaEventCreate(&ev0_));
  if (!ev1_) HP_CUDA(cudaEventCreate(&ev1_));
  have0_ = have2_ = false;
}

static inline int64_t max_of(const std::vector<int64_t>& v) {
  int64_t m = 0;
  for (int64_t x : v) m = std::max(m, x);
  return m;
}

// ---- ring blocks -----------------------------------------------------------------------------
static int64_t gcd_i(int64_t a, int64_t b) { return b == 0 ? a : gcd_i(b, a % b); }

int LegendreStage::setup_blocks(int want, const double* slot_weight) {
  blk_pair_.clear();
  const int np = (int)npairs_;
  if (want < 1 || !adj_warp_mode()) want = 1;
  blk_pair_.push_back(0);
  // Preferred cut: multiples of the least common multiple of the four kernels' tile sizes (512 pairs
  // with the default configurations), so that no block ends in a partly filled tile.  The
  // equator-most block -- whose copy cannot hide behind compute -- is a single unit (6 % of the pixels
  // at nside 4096); the other want-1 blocks share the rest (nside 4096, want 6: 5 x 1536 + 512 pairs).
  int64_t align = 32;
  for (int t : {kSynCfgs[s0_cfg()].R * kSynCfgs[s0_cfg()].NW * 32, kSynCfgs[s2_cfg()].R * kSynCfgs[s2_cfg()].NW * 32,
                adj_tile(0), adj_tile(2)})
    align = align / gcd_i(align, t) * t;
  if (want > 1 && np >= 2 * align) {
    // equator-most block: one unit; the rest: want-1 blocks of equal pair count (whole units)
    const int64_t units = (np + align - 1) / align;           // last unit may be partial
    const int64_t nb = std::min<int64_t>(want - 1, units - 1);
    for (int64_t b = 1; b <= nb; ++b) {
      const int64_t p = ((units - 1) * b / nb) * align;
      if (p > blk_pair_.back() && p < np) blk_pair_.push_back((int)p);
    }
  } else if (want > 1) {
    // small problems (tests): equal-weight blocks on warp-aligned boundaries
    std::vector<double> cum((size_t)np + 1, 0.0);
    for (int p = 0; p < np; ++p) {
      double w = slot_weight[pairs_[(size_t)p].slotN];
      if (pairs_[(size_t)p].slotS >= 0) w += slot_weight[pairs_[(size_t)p].slotS];
      cum[(size_t)p + 1] = cum[(size_t)p] + w;
    }
    for (int b = 1; b < want; ++b) {
      const double target = cum[(size_t)np] * b / want;
      int p = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
      p = (p + 31) / 32 * 32;
      if (p > blk_pair_.back() && p < np) blk_pair_.push_back(p);
    }
  }
  blk_pair_.push_back(np);
  have0_ = have2_ = false;
  return nblocks();
}

std::vector<int64_t> LegendreStage::block_slot_list(int b) const {
  std::vector<int64_t> v;
  for (int p = blk_pair_[(size_t)b]; p < blk_pair_[(size_t)b + 1]; ++p) {
    v.push_back(pairs_[(size_t)p].slotN);
    if (pairs_[(size_t)p].slotS >= 0) v.push_back(pairs_[(size_t)p].slotS);
  }
  return v;
}

void LegendreStage::block_slots(int b, int64_t* slot_lo, int64_t* slot_hi) const {
  int64_t lo = INT64_MAX, hi = -1;
  for (int p = blk_pair_[(size_t)b]; p < blk_pair_[(size_t)b + 1]; ++p) {
    const PairHost& q = pairs_[(size_t)p];
    lo = std::min(lo, q.slotN);
    hi = std::max(hi, q.slotN);
    if (q.slotS >= 0) {
      lo = std::min(lo, q.slotS);
      hi = std::max(hi, q.slotS);
    }
  }
  *slot_lo = lo;
  *slot_hi = hi + 1;
}

// per block: synthesis = one item per active tile; adjoint = one item per m covering the block's
// active tiles, all blocks of an m accumulating into the same partial-sum row (offset off[mi])
void LegendreStage::build_block_items(int spin, const LegendreTables& T, cudaStream_t stream) {
  const bool s2 = spin != 0;
  const std::vector<int64_t>& nstep = s2 ? T.nstep2 : T.nstep0;
  const std::vector<int64_t>& off = s2 ? T.off2 : T.off0;
  const int stile = kSynCfgs[s2 ? s2_cfg() : s0_cfg()].R * kSynCfgs[s2 ? s2_cfg() : s0_cfg()].NW * 32;
  const int atile = adj_tile(spin);
  const int64_t total_steps = off[(size_t)nm_];
  // enough adjoint items per block launch to fill the GPU even when a rank owns few m's
  bgroups_ = (int)std::max<int64_t>(1, (4096 + nm_ - 1) / std::max<int64_t>(nm_, 1));
  std::vector<WorkItem> si, ai;
  std::vector<int64_t> apo;
  std::vector<int> rs, ra;
  for (int b = 0; b < nblocks(); ++b) {
    rs.push_back((int)si.size());
    ra.push_back((int)ai.size());
    const int pa = blk_pair_[(size_t)b], pb = blk_pair_[(size_t)b + 1];
    auto tile_max = [&](int tile) {
      const int nt = (pb - pa + tile - 1) / tile;
      std::vector<int> tm((size_t)nt, -1);
      for (int t = 0; t < nt; ++t)
        for (int p = pa + t * tile; p < std::min(pb, pa + (t + 1) * tile); ++p)
          tm[(size_t)t] = std::max(tm[(size_t)t], s2 ? pairs_[(size_t)p].mlim2 : pairs_[(size_t)p].mlim0);
      return tm;
    };
    const std::vector<int> tms = tile_max(stile), tma = tile_max(atile);
    for (size_t mi = 0; mi < mval_.size(); ++mi) {
      if (nstep[mi] <= 0) continue;
      for (int t = (int)tms.size() - 1; t >= 0; --t) {
        if (tms[(size_t)t] < mval_[mi]) continue;
        WorkItem w;
        w.mi = (int)mi;
        w.p0 = pa + t * stile;
        w.np = std::min(pb, pa + (t + 1) * stile) - w.p0;
        w.part = 0;
        si.push_back(w);
      }
      int tlo = -1, thi = -1;
      for (int t = 0; t < (int)tma.size(); ++t)
        if (tma[(size_t)t] >= mval_[mi]) {
          if (tlo < 0) tlo = t;
          thi = t;
        }
      if (tlo >= 0) {
        // <= G items per (m, block); item g of every block accumulates into row g of this m
        const int nt = thi - tlo + 1;
        const int G = std::min(bgroups_, nt);
        for (int g = 0; g < G; ++g) {
          const int ta = tlo + (int)((int64_t)nt * g / G), tb = tlo + (int)((int64_t)nt * (g + 1) / G);
          WorkItem w;
          w.mi = (int)mi;
          w.p0 = pa + ta * atile;
          w.np = std::min(pb, pa + tb * atile) - w.p0;
          w.part = 0;
          ai.push_back(w);
          apo.push_back((int64_t)g * total_steps + off[mi]);
        }
      }
    }
  }
  rs.push_back((int)si.size());
  ra.push_back((int)ai.size());
  (s2 ? d_bitems_s2_ : d_bitems_s0_).upload(si.data(), si.size(), stream);
  (s2 ? d_bitems_a2_ : d_bitems_a0_).upload(ai.data(), ai.size(), stream);
  (s2 ? d_bpartoff2_ : d_bpartoff0_).upload(apo.data(), apo.size(), stream);
  (s2 ? brange_s2_ : brange_s0_) = rs;
  (s2 ? brange_a2_ : brange_a0_) = ra;
  if (bgroups_ > 1) {  // tables for k_reduce_part over the G rows of every m
    std::vector<int64_t> rpo((size_t)nm_ * bgroups_);
    std::vector<int> rf((size_t)nm_), rc((size_t)nm_, bgroups_);
    for (int64_t mi = 0; mi < nm_; ++mi) {
      rf[(size_t)mi] = (int)(mi * bgroups_);
      for (int g = 0; g < bgroups_; ++g) rpo[(size_t)(mi * bgroups_ + g)] = (int64_t)g * total_steps + off[(size_t)mi];
    }
    (s2 ? d_brpartoff2_ : d_brpartoff0_).upload(rpo.data(), rpo.size(), stream);
    d_brfirst_.upload(rf.data(), rf.size(), stream);
    d_brcount_.upload(rc.data(), rc.size(), stream);
  }
  bpart_steps_need_ = std::max(bpart_steps_need_, (size_t)(total_steps * bgroups_));
  HP_CUDA(cudaStreamSynchronize(stream));
}

void LegendreStage::alm2leg_block(int b, double* d_leg, int ncomp, cudaStream_t stream, bool last) {
  const int spin = ncomp == 1 ? 0 : 2;
  const std::vector<int>& br = spin == 0 ? brange_s0_ : brange_s2_;
  const int first = br[(size_t)b], count = br[(size_t)b + 1] - first;
  if (count > 0) {
    LegParams P;
    fill_params(P, spin, false);
    P.items = (spin == 0 ? d_bitems_s0_.p : d_bitems_s2_.p) + first;
    P.leg = d_leg;
    if (spin == 0)
      HP_LAUNCH(k_alm2leg_s0, env_syn_prod(), s0_cfg(), count, stream, P);
    else
      HP_LAUNCH(k_alm2leg_s2, env_syn_prod(), s2_cfg(), count, stream, P);
    ++launches_;
    HP_CUDA(cudaGetLastError());
  }
  if (last) cudaEventRecord(ev1_, stream);
}

void LegendreStage::leg2alm_blocks_begin(int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  ensure_spin(spin, stream);
  HP_REQUIRE(adj_warp_mode(), HPSHT_ERR_UNSUPPORTED, "ring blocks need the single-warp adjoint kernels");
  d_part_.ensure(std::max(part_steps_need_, bpart_steps_need_) * 4);
  d_accum_.ensure(accum_steps_need_ * 4);
  const int64_t total = (spin == 0 ? off0_h_ : off2_h_)[(size_t)nm_] * bgroups_;
  HP_CUDA(cudaMemsetAsync(d_part_.p, 0, (size_t)total * 4 * sizeof(double), stream));
  cudaEventRecord(ev0_, stream);
}

void LegendreStage::leg2alm_block(int b, const double* d_leg, int ncomp, cudaStream_t stream, bool last) {
  const int spin = ncomp == 1 ? 0 : 2;
  const std::vector<int>& br = spin == 0 ? brange_a0_ : brange_a2_;
  const int first = br[(size_t)b], count = br[(size_t)b + 1] - first;
  if (count > 0) {
    LegParams P;
    fill_params(P, spin, true);
    P.items = (spin == 0 ? d_bitems_a0_.p : d_bitems_a2_.p) + first;
    P.partoff = (spin == 0 ? d_bpartoff0_.p : d_bpartoff2_.p) + first;
    P.leg = const_cast<double*>(d_leg);
    P.accumulate = 1;
    if (spin == 0)
      HP_LAUNCH_W(k_leg2alm_s0_w, a0w_R(), count, stream, P);
    else
      {
        if (adj_nw() > 1 && a2w_R() == 4)
          HP_LAUNCH_MW(k_leg2alm_s2_mw, adj_nw(), count, stream, P);
        else
          HP_LAUNCH_W(k_leg2alm_s2_w, a2w_R(), count, stream, P);
      }
    ++launches_;
    HP_CUDA(cudaGetLastError());
  }
  if (last) cudaEventRecord(ev1_, stream);
}

// the per-m rows of d_part_ are laid out like d_accum_ (offset off[mi]): post-process them directly
void LegendreStage::leg2alm_blocks_end(double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream,
                                       int mi0, int mi1) {
  const int spin = ncomp == 1 ? 0 : 2;
  if (mi1 < 0) mi1 = (int)nm_;
  if (mi1 <= mi0) return;
  const int TB = 256;
  const double* sums = d_part_.p;
  const int64_t mx = std::max<int64_t>(max_of(spin == 0 ? nstep0_h_ : nstep2_h_), 1);  // spin 2 also zeroes l < 2
  dim3 grid((unsigned)((mx + TB - 1) / TB), (unsigned)(mi1 - mi0));
  if (bgroups_ > 1) {  // add the G rows of every m (fixed order) first
    k_reduce_part<<<grid, TB, 0, stream>>>(reinterpret_cast<const Strm*>(d_part_.p),
                                           spin == 0 ? d_brpartoff0_.p : d_brpartoff2_.p, d_brfirst_.p,
                                           d_brcount_.p, spin == 0 ? d_nstep0_.p : d_nstep2_.p,
                                           spin == 0 ? d_off0_.p : d_off2_.p, reinterpret_cast<Strm*>(d_accum_.p),
                                           mi0);
    ++launches_;
    sums = d_accum_.p;
  }
  if (spin == 0)
    k_post_s0<<<grid, TB, 0, stream>>>(reinterpret_cast<const Strm*>(sums), d_mval_.p, d_mstart_.p, d_nstep0_.p,
                                       d_off0_.p, d_invs0_.p, lmax_, reinterpret_cast<double2*>(d_alm), mi0);
  else
    k_post_s2<<<grid, TB, 0, stream>>>(reinterpret_cast<const Strm*>(sums), d_mval_.p, d_mstart_.p, d_nstep2_.p,
                                       d_off2_.p, d_sig2_.p, lmax_, reinterpret_cast<double2*>(d_alm),
                                       reinterpret_cast<double2*>(d_alm) + alm_cstride, mi0);
  ++launches_;
  HP_CUDA(cudaGetLastError());
}

// number of local m's (a prefix, m ascending) that ring block b touches at all: the others are culled
// there (m > mlim of every pair of the block)
int LegendreStage::block_mi_end(int b, int spin) const {
  int ml = -1;
  for (int p = blk_pair_[(size_t)b]; p < blk_pair_[(size_t)b + 1]; ++p)
    ml = std::max(ml, spin == 0 ? pairs_[(size_t)p].mlim0 : pairs_[(size_t)p].mlim2);
  int n = 0;
  while (n < (int)nm_ && mval_[(size_t)n] <= ml) ++n;
  return n;
}

// first item of every m-slab (items are ordered by local m index)
static std::vector<int> slab_item_ranges(const std::vector<WorkItem>& items, const std::vector<int>& slab_first_mi) {
  std::vector<int> r(slab_first_mi.size(), (int)items.size());
  size_t it = 0;
  for (size_t s = 0; s + 1 < slab_first_mi.size(); ++s) {
    while (it < items.size() && items[it].mi < slab_first_mi[s]) ++it;
    r[s] = (int)it;
  }
  r.back() = (int)items.size();
  return r;
}

void LegendreStage::ensure_spin(int spin, cudaStream_t stream) {
  if ((spin == 0 && have0_) || (spin != 0 && have2_)) return;
  LegendreTables T;
  build_legendre_tables(lmax_, mval_.data(), nm_, LEG_CHUNK, spin == 0, spin != 0, T);
  size_t part_steps = 0;
  // enough adjoint items to fill the GPU several times over even when a rank owns few m's
  const int64_t want_items = adj_warp_mode() ? 8192 : 4096;
  const int adj_groups = (int)std::max<int64_t>(1, (want_items + nm_ - 1) / std::max<int64_t>(nm_, 1));
  if (spin == 0) {
    d_nstep0_.upload(T.nstep0.data(), T.nstep0.size(), stream);
    d_off0_.upload(T.off0.data(), T.off0.size(), stream);
    d_coef0_.upload(T.coef0.data(), T.coef0.size(), stream);
    d_invs0_.upload(T.invs0.data(), T.invs0.size(), stream);
    d_k0_.upload(T.k0.data(), T.k0.size(), stream);
    d_stream0_.alloc((size_t)T.off0[(size_t)nm_] * 4);
    d_stream0_.zero(stream);
    make_items(pairs_, mval_, T.nstep0, false, kSynCfgs[s0_cfg()].R * kSynCfgs[s0_cfg()].NW * 32, 0, items_s0_, nullptr, nullptr);
    make_items(pairs_, mval_, T.nstep0, false, adj_tile(0), adj_groups, items_a0_, &afirst0_, &acount0_);
    apartoff0_.resize(items_a0_.size());
    for (size_t i = 0; i < items_a0_.size(); ++i) {
      apartoff0_[i] = (int64_t)part_steps;
      part_steps += (size_t)round_up(T.nstep0[(size_t)items_a0_[i].mi], LEG_CHUNK);
    }
    d_items_s0_.upload(items_s0_.data(), items_s0_.size(), stream);
    d_items_a0_.upload(items_a0_.data(), items_a0_.size(), stream);
    d_afirst0_.upload(afirst0_.data(), afirst0_.size(), stream);
    d_acount0_.upload(acount0_.data(), acount0_.size(), stream);
    d_apartoff0_.upload(apartoff0_.data(), apartoff0_.size(), stream);
    nstep0_h_ = T.nstep0;
    off0_h_ = T.off0;
    srange_s0_ = slab_item_ranges(items_s0_, slab_first_mi_);
    srange_a0_ = slab_item_ranges(items_a0_, slab_first_mi_);
    direct0_ = true;  // one item per m whose partial row sits at off[mi]: the post kernel reads it in place
    for (int64_t i = 0; i < nm_; ++i)
      direct0_ = direct0_ && acount0_[(size_t)i] == 1 && apartoff0_[(size_t)afirst0_[(size_t)i]] == T.off0[(size_t)i];
    have0_ = true;
  } else {
    d_nstep2_.upload(T.nstep2.data(), T.nstep2.size(), stream);
    d_off2_.upload(T.off2.data(), T.off2.size(), stream);
    d_coef2_.upload(T.coef2.data(), T.coef2.size(), stream);
    d_sig2_.upload(T.sig2.data(), T.sig2.size(), stream);
    d_kp2_.upload(T.kp2.data(), T.kp2.size(), stream);
    d_km2_.upload(T.km2.data(), T.km2.size(), stream);
    d_stream2_.alloc((size_t)T.off2[(size_t)nm_] * 4);
    d_stream2_.zero(stream);
    make_items(pairs_, mval_, T.nstep2, true, kSynCfgs[s2_cfg()].R * kSynCfgs[s2_cfg()].NW * 32, 0, items_s2_, nullptr, nullptr);
    make_items(pairs_, mval_, T.nstep2, true, adj_tile(2), adj_groups, items_a2_, &afirst2_, &acount2_);
    apartoff2_.resize(items_a2_.size());
    for (size_t i = 0; i < items_a2_.size(); ++i) {
      apartoff2_[i] = (int64_t)part_steps;
      part_steps += (size_t)round_up(T.nstep2[(size_t)items_a2_[i].mi], LEG_CHUNK);
    }
    d_items_s2_.upload(items_s2_.data(), items_s2_.size(), stream);
    d_items_a2_.upload(items_a2_.data(), items_a2_.size(), stream);
    d_afirst2_.upload(afirst2_.data(), afirst2_.size(), stream);
    d_acount2_.upload(acount2_.data(), acount2_.size(), stream);
    d_apartoff2_.upload(apartoff2_.data(), apartoff2_.size(), stream);
    nstep2_h_ = T.nstep2;
    off2_h_ = T.off2;
    srange_s2_ = slab_item_ranges(items_s2_, slab_first_mi_);
    srange_a2_ = slab_item_ranges(items_a2_, slab_first_mi_);
    direct2_ = true;
    for (int64_t i = 0; i < nm_; ++i)
      direct2_ = direct2_ && acount2_[(size_t)i] == 1 && apartoff2_[(size_t)afirst2_[(size_t)i]] == T.off2[(size_t)i];
    have2_ = true;
  }
  if (nblocks() > 1) build_block_items(spin, T, stream);
  part_steps_need_ = std::max(part_steps_need_, part_steps);
  accum_steps_need_ = std::max(accum_steps_need_,
                               (size_t)(spin == 0 ? T.off0[(size_t)nm_] : T.off2[(size_t)nm_]));
  HP_CUDA(cudaStreamSynchronize(stream));  // host vectors of T die here
}

void LegendreStage::fill_params(LegParams& P, int spin, bool adjoint) const {
  std::memset(&P, 0, sizeof(P));
  P.mval = d_mval_.p;
  P.cth = d_cth_.p;
  P.sth = d_sth_.p;
  P.sh = d_sh_.p;
  P.ch = d_ch_.p;
  P.a0 = d_a0_.p;
  P.nrr = d_nrr_.p;
  P.jN = d_jN_.p;
  P.jS = d_jS_.p;
  P.sstart = d_sstart_.p;
  P.nms = d_nms_.p;
  P.ncs = layout_.ncs;
  if (spin == 0) {
    P.items = adjoint ? d_items_a0_.p : d_items_s0_.p;
    P.nstep = d_nstep0_.p;
    P.off = d_off0_.p;
    P.coef = d_coef0_.p;
    P.k0 = d_k0_.p;
    P.mlim = d_mlim0_.p;
    P.stream = d_stream0_.p;
    P.partoff = d_apartoff0_.p;
  } else {
    P.items = adjoint ? d_items_a2_.p : d_items_s2_.p;
    P.nstep = d_nstep2_.p;
    P.off = d_off2_.p;
    P.coef = d_coef2_.p;
    P.k0 = d_kp2_.p;
    P.k1 = d_km2_.p;
    P.mlim = d_mlim2_.p;
    P.stream = d_stream2_.p;
    P.partoff = d_apartoff2_.p;
  }
  P.part = d_part_.p;
}

// ---- synthesis ------------------------------------------------------------------------------
void LegendreStage::alm2leg_zero(double* d_leg, int64_t leg_elems, int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  ensure_spin(spin, stream);
  // (m, ring) combinations beyond mlim are never visited by a CTA: their leg is exactly 0
  HP_CUDA(cudaMemsetAsync(d_leg, 0, (size_t)leg_elems * 2 * sizeof(double), stream));
  cudaEventRecord(ev0_, stream);
}

// alm -> prepared per-step stream for the local m's [mi0, mi1)
void LegendreStage::alm2leg_prep(const double* d_alm, int64_t alm_cstride, int ncomp, int mi0, int mi1,
                                 cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  if (mi1 <= mi0) return;
  const int TB = 256;
  const int64_t mx = max_of(spin == 0 ? nstep0_h_ : nstep2_h_);
  if (mx <= 0) return;
  dim3 grid((unsigned)((mx + TB - 1) / TB), (unsigned)(mi1 - mi0));
  if (spin == 0)
    k_prep_s0<<<grid, TB, 0, stream>>>(reinterpret_cast<const double2*>(d_alm), d_mval_.p, d_mstart_.p,
                                       d_nstep0_.p, d_off0_.p, d_invs0_.p, lmax_,
                                       reinterpret_cast<Strm*>(d_stream0_.p), mi0);
  else
    k_prep_s2<<<grid, TB, 0, stream>>>(reinterpret_cast<const double2*>(d_alm),
                                       reinterpret_cast<const double2*>(d_alm) + alm_cstride, d_mval_.p,
                                       d_mstart_.p, d_nstep2_.p, d_off2_.p, d_sig2_.p,
                                       reinterpret_cast<Strm*>(d_stream2_.p), mi0);
  ++launches_;
  HP_CUDA(cudaGetLastError());
}

void LegendreStage::alm2leg_begin(const double* d_alm, int64_t alm_cstride, double* d_leg,
                                  int64_t leg_elems, int ncomp, cudaStream_t stream) {
  alm2leg_zero(d_leg, leg_elems, ncomp, stream);
  alm2leg_prep(d_alm, alm_cstride, ncomp, 0, (int)nm_, stream);
  cudaEventRecord(ev0_, stream);
}

void LegendreStage::alm2leg_slab(int slab, double* d_leg, int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  const std::vector<int>& sr = spin == 0 ? srange_s0_ : srange_s2_;
  const int first = sr[(size_t)slab], count = sr[(size_t)slab + 1] - first;
  if (count > 0) {
    LegParams P;
    fill_params(P, spin, false);
    P.items += first;
    P.leg = d_leg;
    if (spin == 0)
      HP_LAUNCH(k_alm2leg_s0, env_syn_prod(), s0_cfg(), count, stream, P);
    else
      HP_LAUNCH(k_alm2leg_s2, env_syn_prod(), s2_cfg(), count, stream, P);
    ++launches_;
    HP_CUDA(cudaGetLastError());
  }
  if (slab == nslabs() - 1) cudaEventRecord(ev1_, stream);
}

void LegendreStage::alm2leg(const double* d_alm, int64_t alm_cstride, double* d_leg,
                            int64_t leg_elems, int ncomp, cudaStream_t stream) {
  alm2leg_begin(d_alm, alm_cstride, d_leg, leg_elems, ncomp, stream);
  for (int s = 0; s < nslabs(); ++s) alm2leg_slab(s, d_leg, ncomp, stream);
}

// ---- adjoint --------------------------------------------------------------------------------
void LegendreStage::leg2alm_begin(int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  ensure_spin(spin, stream);
  d_part_.ensure(part_steps_need_ * 4);
  d_accum_.ensure(accum_steps_need_ * 4);
  cudaEventRecord(ev0_, stream);
}

void LegendreStage::leg2alm_slab(int slab, const double* d_leg, int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  const std::vector<int>& sr = spin == 0 ? srange_a0_ : srange_a2_;
  const int first = sr[(size_t)slab], count = sr[(size_t)slab + 1] - first;
  if (count > 0) {
    LegParams P;
    fill_params(P, spin, true);
    P.items += first;
    P.partoff += first;
    P.leg = const_cast<double*>(d_leg);
    if (spin == 0) {
      if (adj_warp_mode())
        HP_LAUNCH_W(k_leg2alm_s0_w, a0w_R(), count, stream, P);
      else
        HP_LAUNCH(k_leg2alm_s0, env_adj_prod(), a0_cfg(), count, stream, P);
    } else {
      if (adj_warp_mode())
        {
        if (adj_nw() > 1 && a2w_R() == 4)
          HP_LAUNCH_MW(k_leg2alm_s2_mw, adj_nw(), count, stream, P);
        else
          HP_LAUNCH_W(k_leg2alm_s2_w, a2w_R(), count, stream, P);
      }
      else
        HP_LAUNCH(k_leg2alm_s2, env_adj_prod(), a2_cfg(), count, stream, P);
    }
    ++launches_;
    HP_CUDA(cudaGetLastError());
  }
  if (slab == nslabs() - 1) cudaEventRecord(ev1_, stream);
}

void LegendreStage::mark_main_end(cudaStream_t stream) { cudaEventRecord(ev1_, stream); }

void LegendreStage::leg2alm_end(double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  const int TB = 256;
  const int64_t mx = std::max<int64_t>(max_of(spin == 0 ? nstep0_h_ : nstep2_h_), 1);  // spin 2 also zeroes l < 2
  dim3 grid((unsigned)((mx + TB - 1) / TB), (unsigned)nm_);
  const double* sums = d_part_.p;
  if (!(spin == 0 ? direct0_ : direct2_)) {  // several items per m: add their partial rows in fixed order
    k_reduce_part<<<grid, TB, 0, stream>>>(reinterpret_cast<const Strm*>(d_part_.p),
                                           spin == 0 ? d_apartoff0_.p : d_apartoff2_.p,
                                           spin == 0 ? d_afirst0_.p : d_afirst2_.p,
                                           spin == 0 ? d_acount0_.p : d_acount2_.p,
                                           spin == 0 ? d_nstep0_.p : d_nstep2_.p, spin == 0 ? d_off0_.p : d_off2_.p,
                                           reinterpret_cast<Strm*>(d_accum_.p), 0);
    ++launches_;
    sums = d_accum_.p;
  }
  if (spin == 0)
    k_post_s0<<<grid, TB, 0, stream>>>(reinterpret_cast<const Strm*>(sums), d_mval_.p, d_mstart_.p, d_nstep0_.p,
                                       d_off0_.p, d_invs0_.p, lmax_, reinterpret_cast<double2*>(d_alm), 0);
  else
    k_post_s2<<<grid, TB, 0, stream>>>(reinterpret_cast<const Strm*>(sums), d_mval_.p, d_mstart_.p, d_nstep2_.p,
                                       d_off2_.p, d_sig2_.p, lmax_, reinterpret_cast<double2*>(d_alm),
                                       reinterpret_cast<double2*>(d_alm) + alm_cstride, 0);
  ++launches_;
  HP_CUDA(cudaGetLastError());
}

void LegendreStage::leg2alm(const double* d_leg, double* d_alm, int64_t alm_cstride, int ncomp,
                            cudaStream_t stream) {
  leg2alm_begin(ncomp, stream);
  for (int s = 0; s < nslabs(); ++s) leg2alm_slab(s, d_leg, ncomp, stream);
  leg2alm_end(d_alm, alm_cstride, ncomp, stream);
}

}  // namespace hpsht
