// legendre.cu — sm_100a FP64 kernels of the Legendre stage: alm2leg / leg2alm, spin 0 and spin 2.
//
// Replaces Ducc0.Sht.alm2leg! (reference src/sht.jl:85) and Ducc0.Sht.leg2alm! (src/sht.jl:178).
//
// Kernel design (B200-first):
//   * a work item = (one local m) x (a range of N/S ring pairs, sorted pole->equator, all of one accuracy class
//     of legendre_tables.hpp); every thread owns R ring pairs and evaluates the scaled recurrence on the fly
//     in registers (no Plm table); north/south symmetry gives both rings of a pair from one recurrence;
//   * one recurrence step template (rec_step) serves every kernel: the standard forms X / Y differ only in the
//     coefficient table and the per-ring quantity they are fed, the difference forms DP / DE (poles / spin-0
//     equator) carry (value, first difference) instead of (value, previous value);
//   * the per-m coefficient stream (recurrence coefficients + prepared alm) is the only data the inner loop
//     reads; one thread moves it global->shared with 1-D TMA bulk copies (cp.async.bulk ...
//     mbarrier::complete_tx) through a 4-stage mbarrier ring, the consumers read it as warp-broadcast LDS.128;
//   * dynamic range: value = lam * 2^(800*sc), sc <= 0 tracked per ring; a warp whose lanes are all still
//     below range runs the recurrence alone; range checks every 8 steps;
//   * synthesis: one CTA = one tile of 32 R NW pairs, NW consumer warps + one TMA producer warp;
//   * adjoint: one CTA = one warp walking its pair range tile by tile; the 32-lane reduction of the 16 partial
//     sums of every 4 steps goes through a padded shared-memory transpose; the reduced sums of a 64-step chunk
//     are added into the item's partial-sum row in global memory (deterministic, no atomics); the post
//     kernel adds the rows of an m in fixed order and turns them into a_lm.
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
// warp does not take issue slots from the FP64 warps
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity, 4000u)) __nanosleep(200);
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
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  const int64_t* nstep;            // real steps
  const int64_t* off;              // offset in steps (padded to LEG_CHUNK) into stream / coefa / sums
  const int64_t* coff[LC_COUNT];   // per class: offset in steps into coef[class]
  const double* coef[LC_COUNT];    // per class: 2 doubles / step
  const double* coefa;             // spin 2: a_l per step
  const double* k0;                // spin0: p_0 constant ; spin2: kp2
  const double* k1;                // spin2: km2
  // per pair
  const double* yq[LC_COUNT];      // per class: the per-ring quantity the step coefficient is evaluated at
  const double* cth;
  const double* sth;
  const double* sh;
  const double* ch;
  const int* mlim;
  // leg addressing: element (pair p, local m index mi, component c) of ring "N" lives at
  //   a0[p] + (c * nrr[p] + jN[p]) * nm + mi
  const int64_t* a0;
  const int64_t* nrr;
  const int64_t* jN;
  const int64_t* jS;               // -1: unpaired
  int64_t nm;
  int n_y;                         // first pair of class X: a standard adjoint item that starts in Y switches there
  // data
  const double* stream;            // 4 doubles / step (synthesis input)
  double* leg;                     // complex128 (synthesis out / adjoint in)
  double* part;                    // adjoint partial sums, 4 doubles / step
  const int64_t* partoff;          // per item: offset in steps into part
};

namespace {

template <class T>
__device__ __forceinline__ T by_class(T const (&a)[LC_COUNT], int c) {
  return c == LC_X ? a[LC_X] : (c == LC_Y ? a[LC_Y] : (c == LC_DP ? a[LC_DP] : a[LC_DE]));
}
__device__ __forceinline__ int64_t leg_offN(const LegParams& P, int p, int mi) {
  return P.a0[p] + P.jN[p] * P.nm + mi;
}
__device__ __forceinline__ int64_t leg_offS(const LegParams& P, int p, int mi) {
  const int64_t j = P.jS[p];
  return j < 0 ? -1 : P.a0[p] + j * P.nm + mi;
}
__device__ __forceinline__ int64_t leg_cstr(const LegParams& P, int p) { return P.nrr[p] * P.nm; }

// The one recurrence step.  Standard form: (u, v) = (value, previous value), u' = t u - v.
// Difference form: (u, v) = (value, first difference) and t is the step coefficient minus 2:
// v' = v + t u, u' = u + v'.
template <bool DIFF>
__device__ __forceinline__ void rec_step(double t, double& u, double& v) {
  if (DIFF) {
    v = fma(t, u, v);
    u = u + v;
  } else {
    const double nx = fma(t, u, -v);
    v = u;
    u = nx;
  }
}

// start values: spin 0  p_0 = k0 sin^m(theta);  spin 2  delta+- at l0 = max(m, 2)
__device__ __forceinline__ void start_s0(const LegParams& P, int p, int m, double k0, bool act, double& u, int& sc) {
  double mant;
  int e;
  pow_scaled(P.sth[p], m, mant, e);
  to_scaled(k0 * mant, e, u, sc);
  if (!act) {  // culled lanes carry zeros so nothing can overflow
    u = 0.0;
    sc = 0;
  }
}
__device__ __forceinline__ void start_s2(const LegParams& P, int p, int m, double kp, double km, bool act,
                                         double& up, int& scp, double& um, int& scm) {
  const int pw = m >= 2 ? m - 2 : 2 - m;
  const int q = m < 2 ? m : 2;
  double mant;
  int e;
  pow_scaled(P.sth[p], pw, mant, e);
  double fs = 1.0, fc = 1.0;
  const double s2 = P.sh[p] * P.sh[p], c2 = P.ch[p] * P.ch[p];
  for (int k = 0; k < q; ++k) {
    fs *= s2;
    fc *= c2;
  }
  to_scaled(kp * mant * fs, e, up, scp);
  to_scaled(km * mant * fc, e, um, scm);
  if (!act) {
    up = um = 0.0;
    scp = scm = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// producer side of the synthesis coefficient pipeline: one thread of a dedicated warp issues the 1-D
// TMA bulk copies of chunk cn into stage cn % NSTAGE once every consumer warp released that stage.
// ------------------------------------------------------------------------------------------------
template <int NSRC>
struct Pump {
  const char* g[NSRC];   // global source of chunk 0
  char* s[NSRC];         // shared destination of stage 0
  uint32_t bytes[NSRC];  // per chunk (= stage stride)
  uint64_t* full;
  uint64_t* empty;
  __device__ __forceinline__ void run_all(int nchunks) {
    uint32_t tot = 0;
#pragma unroll
    for (int i = 0; i < NSRC; ++i) tot += bytes[i];
    for (int cn = 0; cn < nchunks; ++cn) {
      const int st = cn % NSTAGE;
      if (cn >= NSTAGE) mbar_wait_relaxed(&empty[st], ((cn / NSTAGE) - 1) & 1);
      mbar_arrive_expect_tx(&full[st], tot);
#pragma unroll
      for (int i = 0; i < NSRC; ++i)
        tma_load_1d(s[i] + (size_t)st * bytes[i], g[i] + (size_t)cn * bytes[i], bytes[i], &full[st]);
    }
  }
};

}  // namespace

// ================================================================================================
// spin 0 synthesis:  leg_N/S = sum_n p_n Ae_n  +-  x sum_n p_n Ao_n
// MODE 0: standard forms (class X or Y by item), 1: polar difference form, 2: equatorial difference form
// (which runs w_n = (-1)^n p_n, hence the alternating sign in the accumulation)
// ================================================================================================
template <int R, int NW, int MODE>
__global__ void __launch_bounds__((NW + 1) * 32) k_alm2leg_s0(LegParams P) {
  constexpr bool DIFF = MODE != 0, ALT = MODE == 2;
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

  if (warp >= NW) {
    if (lane == 0) {
      Pump<2> pump;
      pump.g[0] = reinterpret_cast<const char*>(by_class(P.coef, it.cls) + (size_t)by_class(P.coff, it.cls)[mi] * 2);
      pump.g[1] = reinterpret_cast<const char*>(P.stream + (size_t)P.off[mi] * 4);
      pump.s[0] = reinterpret_cast<char*>(&s_coef[0][0]);
      pump.s[1] = reinterpret_cast<char*>(&s_strm[0][0]);
      pump.bytes[0] = LEG_CHUNK * sizeof(Coef);
      pump.bytes[1] = LEG_CHUNK * sizeof(Strm);
      pump.full = s_full;
      pump.empty = s_empty;
      pump.run_all(nchunks);
    }
    return;
  }

  // ---------------- consumer warps ----------------
  double y[R], u[R], v[R];
  double er[R], ei[R], orr[R], oi[R];
  int sc[R];
  bool act[R];
  bool any_act = false;
  const double k0 = P.k0[mi];
  const double* yarr = by_class(P.yq, it.cls);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pl = warp * 32 * R + r * 32 + lane;
    const bool valid = pl < it.np;
    const int p = it.p0 + (valid ? pl : 0);
    act[r] = valid && (m <= P.mlim[p]);
    any_act |= act[r];
    y[r] = yarr[p];
    start_s0(P, p, m, k0, act[r], u[r], sc[r]);
    v[r] = DIFF ? u[r] : 0.0;
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
              const double pv = (ALT && (s & 1)) ? -u[r] : u[r];  // groups start at even n
              er[r] = fma(pv, al.x, er[r]);
              ei[r] = fma(pv, al.y, ei[r]);
              orr[r] = fma(pv, al.z, orr[r]);
              oi[r] = fma(pv, al.w, oi[r]);
              rec_step<DIFF>(fma(ab.a, y[r], ab.b), u[r], v[r]);
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < LEG_GROUP; ++s) {
            const Coef ab = cf[s];
#pragma unroll
            for (int r = 0; r < R; ++r) rec_step<DIFF>(fma(ab.a, y[r], ab.b), u[r], v[r]);
          }
        }
        if (w_neg) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (sc[r] < 0 && is_big(u[r])) {
              u[r] *= SC_DOWN;
              v[r] *= SC_DOWN;
              sc[r] += 1;
              er[r] = ei[r] = orr[r] = oi[r] = 0.0;  // drop what was accumulated while scaled
            }
          }
          revote();  // scales changed: refresh the warp-uniform flags
        }
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
    leg[leg_offN(P, p, mi)] = make_double2(Er + Or, Ei + Oi);
    const int64_t oS = leg_offS(P, p, mi);
    if (oS >= 0) leg[oS] = make_double2(Er - Or, Ei - Oi);
  }
}

// ================================================================================================
// spin 2 synthesis.  Stream per l: alpha+ = -(E+iB) sigma_l/2, alpha- = -(E-iB) sigma_l/2.
//   P+N = sum alpha+ d+ ; P-N = sum alpha- d- ; P+S = sum (-1)^(l+m) alpha+ d- ; P-S = sum (-1)^(l+m) alpha- d+
//   Q = P+ + P- ; U = -i (P+ - P-)
// MODE 0: standard forms (X: t = a x +- b;  Y: t = -a y + (a +- b)), 1: polar difference form
// ================================================================================================
template <int R, int NW, int MODE>
__global__ void __launch_bounds__((NW + 1) * 32) k_alm2leg_s2(LegParams P) {
  constexpr bool DIFF = MODE != 0;
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ Strm s_strm[NSTAGE][LEG_CHUNK];
  __shared__ __align__(16) double s_ca[NSTAGE][LEG_CHUNK];
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

  if (warp >= NW) {
    if (lane == 0) {
      Pump<3> pump;
      pump.g[0] = reinterpret_cast<const char*>(by_class(P.coef, it.cls) + (size_t)by_class(P.coff, it.cls)[mi] * 2);
      pump.g[1] = reinterpret_cast<const char*>(P.stream + (size_t)P.off[mi] * 4);
      pump.g[2] = reinterpret_cast<const char*>(P.coefa + (size_t)P.off[mi]);
      pump.s[0] = reinterpret_cast<char*>(&s_coef[0][0]);
      pump.s[1] = reinterpret_cast<char*>(&s_strm[0][0]);
      pump.s[2] = reinterpret_cast<char*>(&s_ca[0][0]);
      pump.bytes[0] = LEG_CHUNK * sizeof(Coef);
      pump.bytes[1] = LEG_CHUNK * sizeof(Strm);
      pump.bytes[2] = LEG_CHUNK * sizeof(double);
      pump.full = s_full;
      pump.empty = s_empty;
      pump.run_all(nchunks);
    }
    return;
  }

  double y[R], up[R], vp[R], um[R], vm[R];
  double pnr[R], pni[R], mnr[R], mni[R];  // P+N, P-N
  double psr[R], psi[R], msr[R], msi[R];  // P+S, P-S
  int scp[R], scm[R];
  bool act[R];
  bool any_act = false;
  const double kp = P.k0[mi], km = P.k1[mi];
  const double* yarr = by_class(P.yq, it.cls);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pl = warp * 32 * R + r * 32 + lane;
    const bool valid = pl < it.np;
    const int p = it.p0 + (valid ? pl : 0);
    act[r] = valid && (m <= P.mlim[p]);
    any_act |= act[r];
    y[r] = yarr[p];
    start_s2(P, p, m, kp, km, act[r], up[r], scp[r], um[r], scm[r]);
    vp[r] = DIFF ? up[r] : 0.0;
    vm[r] = DIFF ? um[r] : 0.0;
    pnr[r] = pni[r] = mnr[r] = mni[r] = 0.0;
    psr[r] = psi[r] = msr[r] = msi[r] = 0.0;
  }
  const bool warp_act = __any_sync(0xffffffffu, any_act);
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
    mbar_wait(&s_full[st], (c / NSTAGE) & 1);
    if (warp_act) {
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      for (int g = 0; g < ng; ++g) {
        const Coef* cf = &s_coef[st][g * LEG_GROUP];
        const double2* ca = reinterpret_cast<const double2*>(&s_ca[st][g * LEG_GROUP]);
        if (w_any0) {
          const Strm* sm = &s_strm[st][g * LEG_GROUP];
#pragma unroll
          for (int s2 = 0; s2 < LEG_GROUP; s2 += 2) {
            const double2 a2 = ca[s2 >> 1];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int s = s2 + h;
              const double a = h ? a2.y : a2.x;
              const Coef cc = cf[s];  // (c+, c-)
              const Strm al = sm[s];  // x,y = alpha+ ; z,w = alpha-
#pragma unroll
              for (int r = 0; r < R; ++r) {
                pnr[r] = fma(up[r], al.x, pnr[r]);
                pni[r] = fma(up[r], al.y, pni[r]);
                mnr[r] = fma(um[r], al.z, mnr[r]);
                mni[r] = fma(um[r], al.w, mni[r]);
                const double sm_ = h ? -um[r] : um[r];  // steps within a group start at an even offset
                const double sp_ = h ? -up[r] : up[r];
                psr[r] = fma(sm_, al.x, psr[r]);
                psi[r] = fma(sm_, al.y, psi[r]);
                msr[r] = fma(sp_, al.z, msr[r]);
                msi[r] = fma(sp_, al.w, msi[r]);
                rec_step<DIFF>(fma(a, y[r], cc.a), up[r], vp[r]);
                rec_step<DIFF>(fma(a, y[r], cc.b), um[r], vm[r]);
              }
            }
          }
        } else {
#pragma unroll
          for (int s2 = 0; s2 < LEG_GROUP; s2 += 2) {
            const double2 a2 = ca[s2 >> 1];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double a = h ? a2.y : a2.x;
              const Coef cc = cf[s2 + h];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                rec_step<DIFF>(fma(a, y[r], cc.a), up[r], vp[r]);
                rec_step<DIFF>(fma(a, y[r], cc.b), um[r], vm[r]);
              }
            }
          }
        }
        if (w_neg) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (scp[r] < 0 && is_big(up[r])) {
              up[r] *= SC_DOWN;
              vp[r] *= SC_DOWN;
              scp[r] += 1;
              pnr[r] = pni[r] = msr[r] = msi[r] = 0.0;
            }
            if (scm[r] < 0 && is_big(um[r])) {
              um[r] *= SC_DOWN;
              vm[r] *= SC_DOWN;
              scm[r] += 1;
              mnr[r] = mni[r] = psr[r] = psi[r] = 0.0;
            }
          }
          revote();
        }
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
    const int64_t cs = leg_cstr(P, p);
    const int64_t oN = leg_offN(P, p, mi);
    // Q = P+ + P- ; U = -i (P+ - P-) = ( Im(P+ - P-), -Re(P+ - P-) )
    leg[oN] = make_double2(PNr + MNr, PNi + MNi);
    leg[oN + cs] = make_double2(PNi - MNi, -(PNr - MNr));
    const int64_t oS = leg_offS(P, p, mi);
    if (oS >= 0) {
      leg[oS] = make_double2(PSr + MSr, PSi + MSi);
      leg[oS + cs] = make_double2(PSi - MSi, -(PSr - MSr));
    }
  }
}

// ================================================================================================
// adjoint kernels: one CTA = one warp = one work item (one m, a contiguous range of ring pairs walked
// tile by tile, tile = 32 R pairs).  No block-level barrier exists: a warp that is still ramping up, or
// that has finished, never holds other warps' registers hostage, and the scheduler backfills with the
// next item.  The coefficient stream is staged by TMA bulk copies issued by lane 0; with a single
// consumer warp a stage is free as soon as the warp has passed it, so only the `full` mbarriers are
// needed.  A standard (MODE 0) item that starts in class Y runs on into class X at pair P.n_y: tiles
// never straddle the boundary, the coefficient table and the per-ring quantity are switched there.
// ================================================================================================
namespace {

// The 32-lane reduction of 16 per-lane partial sums goes through a padded shared-memory transpose, split in
// two so that the second half overlaps the next group's FMAs: tr_store parks the 16 values, tr_reduce (one
// group later, other buffer in between) lets every lane add up one value's 16-lane half and combines the halves
// with one shuffle -- every lane returns the total of value (lane & 15).
__device__ __forceinline__ void tr_store(const double (&v)[16], double (*tr)[33], int lane) {
#pragma unroll
  for (int k = 0; k < 16; ++k) tr[k][lane] = v[k];
  __syncwarp();
}
// quarter t4 (0..3) of a lane's 16 entries; called once per recurrence step so that the loads and adds spread
// over the step's FMAs
struct TrSum {
  double s0, s1, s2, s3;
};
__device__ __forceinline__ void tr_quarter(TrSum& a, const double (*tr)[33], int lane, int t4) {
  const int j = lane & 15, h = (lane >> 4) * 16 + 4 * t4;
  a.s0 += tr[j][h];
  a.s1 += tr[j][h + 1];
  a.s2 += tr[j][h + 2];
  a.s3 += tr[j][h + 3];
}
__device__ __forceinline__ double tr_finish(const TrSum& a) {
  double s = (a.s0 + a.s1) + (a.s2 + a.s3);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  return s;
}
__device__ __forceinline__ double tr_reduce(const double (*tr)[33], int lane) {
  TrSum a = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int t = 0; t < 4; ++t) tr_quarter(a, tr, lane, t);
  return tr_finish(a);
}
constexpr int RES_DUMMY = LEG_CHUNK / 4;  // s_res slot (of 16 doubles) that takes the reduction of "nothing pending"

// the tiles of an adjoint item: segment A = [p0, split) in the item's class, segment B = [split, pend) in X
struct TileWalk {
  int p0, split, pend, ntA, ntile;
};
template <int TILE>
__device__ __forceinline__ TileWalk tile_walk(const WorkItem& it, int n_y, bool can_switch) {
  TileWalk w;
  w.p0 = it.p0;
  w.pend = it.p0 + it.np;
  w.split = (can_switch && it.cls == LC_Y) ? min(w.pend, max(w.p0, n_y)) : w.pend;
  w.ntA = (w.split - w.p0 + TILE - 1) / TILE;
  w.ntile = w.ntA + (w.pend - w.split + TILE - 1) / TILE;
  return w;
}

}  // namespace

// spin 0 adjoint: At_n = sum_pairs p_n (legN + legS), Bt_n = sum_pairs p_n x (legN - legS)
template <int R, int MODE>
__global__ void __launch_bounds__(32, (R <= 4) ? 16 : 1) k_leg2alm_s0(LegParams P) {
  constexpr bool DIFF = MODE != 0, ALT = MODE == 2;
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ double s_tr[2][16][33];
  __shared__ double s_res[LEG_CHUNK * 4 + 16];  // reduced sums of the current chunk (+ the dummy slot)
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  constexpr int TILE = R * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nn = (int)P.nstep[mi];
  const int nsteps = (nn + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int lane = threadIdx.x;
  const TileWalk tw = tile_walk<TILE>(it, P.n_y, MODE == 0);

  if (lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const double* gA = by_class(P.coef, it.cls) + (size_t)by_class(P.coff, it.cls)[mi] * 2;
  const double* gB = P.coef[LC_X] + (size_t)P.coff[LC_X][mi] * 2;
  const int total = nchunks * tw.ntile;
  auto issue = [&](int cn) {
    const int st = cn % NSTAGE;
    const int t = cn / nchunks, c = cn - t * nchunks;
    mbar_arrive_expect_tx(&s_full[st], LEG_CHUNK * sizeof(Coef));
    tma_load_1d(&s_coef[st][0], (t < tw.ntA ? gA : gB) + (size_t)c * LEG_CHUNK * 2, LEG_CHUNK * sizeof(Coef),
                &s_full[st]);
  };
  if (lane == 0)
    for (int cn = 0; cn < NSTAGE - 1 && cn < total; ++cn) issue(cn);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double k0 = P.k0[mi];
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  int cg = 0;
  int tb = 0;  // s_tr buffer the next group parks its sums in

  for (int tile = 0; tile < tw.ntile; ++tile) {
    const bool inA = tile < tw.ntA;
    const int tp0 = inA ? tw.p0 + tile * TILE : tw.split + (tile - tw.ntA) * TILE;
    const int tnp = min(TILE, (inA ? tw.split : tw.pend) - tp0);
    const double* yarr = by_class(P.yq, inA ? it.cls : (int)LC_X);
    double y[R], u[R], v[R];
    double er[R], ei[R], orr[R], oi[R];
    int sc[R];
    bool act[R];
    bool any_act = false;

    auto load_in = [&](int r, int p) {
      const double2 gN = leg[leg_offN(P, p, mi)];
      const int64_t oS = leg_offS(P, p, mi);
      const double2 gS = oS >= 0 ? leg[oS] : make_double2(0.0, 0.0);
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
      y[r] = yarr[p];
      start_s0(P, p, m, k0, act[r], u[r], sc[r]);
      v[r] = DIFF ? u[r] : 0.0;
      er[r] = ei[r] = orr[r] = oi[r] = 0.0;
      if (act[r] && sc[r] == 0) load_in(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);
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
      if (lane == 0 && cg + NSTAGE - 1 < total) {
        fence_proxy_async();  // the stage being refilled was read (generic proxy) during chunk cg-1
        issue(cg + NSTAGE - 1);
      }
      // partial sums of the earlier tiles / blocks for this chunk (element k*32 + lane), fetched early
      double* gp = gpart + (size_t)c * LEG_CHUNK * 4 + lane;
      double old[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) old[k] = gp[k * 32];
#pragma unroll
      for (int k = 0; k < 8; ++k) s_res[k * 32 + lane] = 0.0;  // groups that only ramp contribute 0
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      int pend = RES_DUMMY;  // 4-step group whose 16 partial sums are parked in s_tr[tb ^ 1]
      if (warp_act) {
#pragma unroll 1
        for (int g = 0; g < ng; ++g) {
          const Coef* cf = &s_coef[st][g * LEG_GROUP];
          if (w_any0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              // the previous 4-step group's reduction runs next to this group's FMAs, a quarter per step
              TrSum ts = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const Coef ab = cf[h * 4 + s];
                tr_quarter(ts, s_tr[tb ^ 1], lane, s);
                // the step's four sums over this thread's R rings go straight to the transpose buffer
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const double pv = (ALT && (s & 1)) ? -u[r] : u[r];
                  a0 = fma(pv, er[r], a0);
                  a1 = fma(pv, ei[r], a1);
                  a2 = fma(pv, orr[r], a2);
                  a3 = fma(pv, oi[r], a3);
                  rec_step<DIFF>(fma(ab.a, y[r], ab.b), u[r], v[r]);
                }
                s_tr[tb][4 * s + 0][lane] = a0;
                s_tr[tb][4 * s + 1][lane] = a1;
                s_tr[tb][4 * s + 2][lane] = a2;
                s_tr[tb][4 * s + 3][lane] = a3;
              }
              {
                const double tot = tr_finish(ts);
                if (lane < 16) s_res[pend * 16 + lane] = tot;
              }
              __syncwarp();
              pend = g * 2 + h;  // 4-step group q = 2g + h
              tb ^= 1;
            }
          } else {
#pragma unroll
            for (int s = 0; s < LEG_GROUP; ++s) {
              const Coef ab = cf[s];
#pragma unroll
              for (int r = 0; r < R; ++r) rec_step<DIFF>(fma(ab.a, y[r], ab.b), u[r], v[r]);
            }
          }
          if (w_neg) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (sc[r] < 0 && is_big(u[r])) {
                u[r] *= SC_DOWN;
                v[r] *= SC_DOWN;
                sc[r] += 1;
                if (sc[r] == 0) load_in(r, tp0 + r * 32 + lane);
              }
            }
            revote();
          }
        }
      }
      {  // the last group's reduction
        const double tot = tr_reduce(s_tr[tb ^ 1], lane);
        if (lane < 16) s_res[pend * 16 + lane] = tot;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k * 32] = old[k] + s_res[k * 32 + lane];
      __syncwarp();  // s_res is zeroed again at the top of the next chunk
    }
  }
}

// spin 2 adjoint:  G+ = sum_pairs d+ w+N + (-1)^(l+m) d- w+S ,  G- = sum_pairs d- w-N + (-1)^(l+m) d+ w-S
// with w+- = Q +- iU of the ring
template <int R, int MODE>
__global__ void __launch_bounds__(32, (R <= 4) ? 10 : 1) k_leg2alm_s2(LegParams P) {
  constexpr bool DIFF = MODE != 0;
  __shared__ Coef s_coef[NSTAGE][LEG_CHUNK];
  __shared__ __align__(16) double s_ca[NSTAGE][LEG_CHUNK];
  __shared__ double s_tr[16][33];
  __shared__ double s_res[LEG_CHUNK * 4];
  __shared__ __align__(8) uint64_t s_full[NSTAGE];
  constexpr int TILE = R * 32;

  const WorkItem it = P.items[blockIdx.x];
  const int mi = it.mi;
  const int m = (int)P.mval[mi];
  const int nl = (int)P.nstep[mi];
  const int nsteps = (nl + LEG_GROUP - 1) / LEG_GROUP * LEG_GROUP;
  const int nchunks = (nsteps + LEG_CHUNK - 1) / LEG_CHUNK;
  const int lane = threadIdx.x;
  const TileWalk tw = tile_walk<TILE>(it, P.n_y, MODE == 0);

  if (lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&s_full[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const double* gA = by_class(P.coef, it.cls) + (size_t)by_class(P.coff, it.cls)[mi] * 2;
  const double* gB = P.coef[LC_X] + (size_t)P.coff[LC_X][mi] * 2;
  const double* ga = P.coefa + (size_t)P.off[mi];
  const int total = nchunks * tw.ntile;
  auto issue = [&](int cn) {
    const int st = cn % NSTAGE;
    const int t = cn / nchunks, c = cn - t * nchunks;
    mbar_arrive_expect_tx(&s_full[st], LEG_CHUNK * (sizeof(Coef) + sizeof(double)));
    tma_load_1d(&s_coef[st][0], (t < tw.ntA ? gA : gB) + (size_t)c * LEG_CHUNK * 2, LEG_CHUNK * sizeof(Coef),
                &s_full[st]);
    tma_load_1d(&s_ca[st][0], ga + (size_t)c * LEG_CHUNK, LEG_CHUNK * sizeof(double), &s_full[st]);
  };
  if (lane == 0)
    for (int cn = 0; cn < NSTAGE - 1 && cn < total; ++cn) issue(cn);

  const double2* leg = reinterpret_cast<const double2*>(P.leg);
  const double kp = P.k0[mi], km = P.k1[mi];
  const int l0 = m > 2 ? m : 2;
  const double sgnS = ((l0 + m) & 1) ? -1.0 : 1.0;
  double* gpart = P.part + (size_t)P.partoff[blockIdx.x] * 4;
  int cg = 0;

  for (int tile = 0; tile < tw.ntile; ++tile) {
    const bool inA = tile < tw.ntA;
    const int tp0 = inA ? tw.p0 + tile * TILE : tw.split + (tile - tw.ntA) * TILE;
    const int tnp = min(TILE, (inA ? tw.split : tw.pend) - tp0);
    const double* yarr = by_class(P.yq, inA ? it.cls : (int)LC_X);
    double y[R], up[R], vp[R], um[R], vm[R];
    double wpNr[R], wpNi[R], wmSr[R], wmSi[R], wmNr[R], wmNi[R], wpSr[R], wpSi[R];
    int scp[R], scm[R];
    bool act[R];
    bool any_act = false;

    auto load_p = [&](int r, int p) {  // inputs multiplied by d+
      const int64_t cs = leg_cstr(P, p);
      const int64_t oN = leg_offN(P, p, mi);
      const double2 qN = leg[oN], uN = leg[oN + cs];
      wpNr[r] = qN.x - uN.y;
      wpNi[r] = qN.y + uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS], uS = leg[oS + cs];
        wmSr[r] = sgnS * (qS.x + uS.y);
        wmSi[r] = sgnS * (qS.y - uS.x);
      } else {
        wmSr[r] = wmSi[r] = 0.0;
      }
    };
    auto load_m = [&](int r, int p) {  // inputs multiplied by d-
      const int64_t cs = leg_cstr(P, p);
      const int64_t oN = leg_offN(P, p, mi);
      const double2 qN = leg[oN], uN = leg[oN + cs];
      wmNr[r] = qN.x + uN.y;
      wmNi[r] = qN.y - uN.x;
      const int64_t oS = leg_offS(P, p, mi);
      if (oS >= 0) {
        const double2 qS = leg[oS], uS = leg[oS + cs];
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
      y[r] = yarr[p];
      start_s2(P, p, m, kp, km, act[r], up[r], scp[r], um[r], scm[r]);
      vp[r] = DIFF ? up[r] : 0.0;
      vm[r] = DIFF ? um[r] : 0.0;
      wpNr[r] = wpNi[r] = wmSr[r] = wmSi[r] = 0.0;
      wmNr[r] = wmNi[r] = wpSr[r] = wpSi[r] = 0.0;
      if (act[r] && scp[r] == 0) load_p(r, p);
      if (act[r] && scm[r] == 0) load_m(r, p);
    }
    const bool warp_act = __any_sync(0xffffffffu, any_act);
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
      if (lane == 0 && cg + NSTAGE - 1 < total) {
        fence_proxy_async();
        issue(cg + NSTAGE - 1);
      }
      double* gp = gpart + (size_t)c * LEG_CHUNK * 4 + lane;
      double old[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) old[k] = gp[k * 32];
#pragma unroll
      for (int k = 0; k < 8; ++k) s_res[k * 32 + lane] = 0.0;
      mbar_wait(&s_full[st], (cg / NSTAGE) & 1);
      const int ng = min(LEG_CHUNK, nsteps - c * LEG_CHUNK) / LEG_GROUP;
      if (warp_act) {
#pragma unroll 1
        for (int g = 0; g < ng; ++g) {
          const Coef* cf = &s_coef[st][g * LEG_GROUP];
          const double2* ca = reinterpret_cast<const double2*>(&s_ca[st][g * LEG_GROUP]);
          if (w_any0) {
            // (the pipelined / per-step forms of the spin-0 kernel measured 1 - 3 % slower here: the 13 doubles of
            // state per ring leave the scheduler no room, profiles/r02_adjoint_experiments.md)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              double acc[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const Coef cc = cf[h * 4 + s];
                const double2 a2 = ca[h * 2 + (s >> 1)];
                const double a = (s & 1) ? a2.y : a2.x;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const double sm_ = (s & 1) ? -um[r] : um[r];
                  const double sp_ = (s & 1) ? -up[r] : up[r];
                  acc[4 * s + 0] = fma(up[r], wpNr[r], acc[4 * s + 0]);
                  acc[4 * s + 1] = fma(up[r], wpNi[r], acc[4 * s + 1]);
                  acc[4 * s + 2] = fma(um[r], wmNr[r], acc[4 * s + 2]);
                  acc[4 * s + 3] = fma(um[r], wmNi[r], acc[4 * s + 3]);
                  acc[4 * s + 0] = fma(sm_, wpSr[r], acc[4 * s + 0]);
                  acc[4 * s + 1] = fma(sm_, wpSi[r], acc[4 * s + 1]);
                  acc[4 * s + 2] = fma(sp_, wmSr[r], acc[4 * s + 2]);
                  acc[4 * s + 3] = fma(sp_, wmSi[r], acc[4 * s + 3]);
                  rec_step<DIFF>(fma(a, y[r], cc.a), up[r], vp[r]);
                  rec_step<DIFF>(fma(a, y[r], cc.b), um[r], vm[r]);
                }
              }
              __syncwarp();  // readers of the previous round are done
              tr_store(acc, s_tr, lane);
              const double tot = tr_reduce(s_tr, lane);
              if (lane < 16) s_res[(g * 2 + h) * 16 + lane] = tot;
            }
          } else {
#pragma unroll
            for (int s2 = 0; s2 < LEG_GROUP; s2 += 2) {
              const double2 a2 = ca[s2 >> 1];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const double a = h ? a2.y : a2.x;
                const Coef cc = cf[s2 + h];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  rec_step<DIFF>(fma(a, y[r], cc.a), up[r], vp[r]);
                  rec_step<DIFF>(fma(a, y[r], cc.b), um[r], vm[r]);
                }
              }
            }
          }
          if (w_neg) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const int p = tp0 + r * 32 + lane;
              if (scp[r] < 0 && is_big(up[r])) {
                up[r] *= SC_DOWN;
                vp[r] *= SC_DOWN;
                scp[r] += 1;
                if (scp[r] == 0) load_p(r, p);
              }
              if (scm[r] < 0 && is_big(um[r])) {
                um[r] *= SC_DOWN;
                vm[r] *= SC_DOWN;
                scm[r] += 1;
                if (scm[r] == 0) load_m(r, p);
              }
            }
            revote();
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) gp[k * 32] = old[k] + s_res[k * 32 + lane];
      __syncwarp();
    }
  }
}

// ================================================================================================
// prep / post kernels (HBM-bound, O(nalm))
// ================================================================================================
namespace {

// spin 0 prep: stream[off+n] = { (eps_{l+1} a_l + eps_{l+2} a_{l+2}) / g_n , a_{l+1} / g_n }, l = m+2n
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

// the partial-sum rows of one m, added in fixed order (deterministic, no atomics anywhere)
struct RowSet {
  const Strm* part;
  int nstd, ndp, mi_dp;
  int64_t std_stride, dp_base, dp_stride;
};
__device__ __forceinline__ Strm row_sum(const RowSet& S, int mi, int64_t idx) {
  Strm s = S.part[idx];
  for (int r = 1; r < S.nstd; ++r) {
    const Strm v = S.part[(int64_t)r * S.std_stride + idx];
    s.x += v.x;
    s.y += v.y;
    s.z += v.z;
    s.w += v.w;
  }
  if (mi < S.mi_dp)
    for (int r = 0; r < S.ndp; ++r) {
      const Strm v = S.part[S.dp_base + (int64_t)r * S.dp_stride + idx];
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
  return s;
}

// spin 0 post: a_l = eps_{l+1} At_n/g_n + eps_l At_{n-1}/g_{n-1} (l = m+2n) ; a_{l+1} = Bt_n/g_n
__global__ void k_post_s0(RowSet S, const int64_t* __restrict__ mval, const int64_t* __restrict__ mstart,
                          const int64_t* __restrict__ nstep, const int64_t* __restrict__ off,
                          const double* __restrict__ invs, int64_t lmax, double2* __restrict__ alm, int mi0) {
  const int mi = mi0 + blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nstep[mi]) return;
  const int64_t m = mval[mi];
  const int64_t l = m + 2 * n;
  const Strm cur = row_sum(S, mi, off[mi] + n);
  const double is = invs[off[mi] + n];
  const double e1 = eps_lm(l + 1, m);
  double vr = e1 * is * cur.x, vi = e1 * is * cur.y;
  if (n > 0) {
    const Strm prv = row_sum(S, mi, off[mi] + n - 1);
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
__global__ void k_post_s2(RowSet S, const int64_t* __restrict__ mval, const int64_t* __restrict__ mstart,
                          const int64_t* __restrict__ nstep, const int64_t* __restrict__ off,
                          const double* __restrict__ sig, int64_t lmax, double2* __restrict__ almE,
                          double2* __restrict__ almB, int mi0) {
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
  const Strm g = row_sum(S, mi, off[mi] + j);
  const double h = 0.5 * sig[off[mi] + j];
  almE[mstart[mi] + l] = make_double2(-h * (g.x + g.z), -h * (g.y + g.w));
  // i * (dr, di) = (-di, dr)
  almB[mstart[mi] + l] = make_double2(-h * (g.y - g.w), h * (g.x - g.z));
}

}  // namespace

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
    const long double c = cosl(th), sn = sinl(th), shl = sinl(0.5L * th);
    p.cth = (double)c;
    p.sth = (double)sn;
    p.sh = (double)shl;
    p.ch = (double)cosl(0.5L * th);
    p.x2 = (double)(c * c);
    p.s2 = (double)(sn * sn);
    p.y2 = (double)(2.0L * shl * shl);
    p.theta = theta[n] <= 0.5 * pi ? theta[n] : pi - theta[n];
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
  for (cudaEvent_t e : {ev0_, ev1_, ev_fork_, ev_join_})
    if (e) cudaEventDestroy(e);
  if (side_) cudaStreamDestroy(side_);
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
// Tile shapes (round-1 tuning on B200, profiles/r01_*): synthesis spin 0 <R = 4, 4 warps> = 512 pairs,
// spin 2 <2, 4> = 256 pairs; single-warp adjoint R = 8 (spin 0) / 4 (spin 2).  The equatorial class of
// long ring lists is 256 pairs (make_class_ranges), hence its 2-warp synthesis tile.
constexpr int SYN0_R = 4, SYN0_NW = 4, SYN0_DE_NW = 2, SYN2_R = 2, SYN2_NW = 4, ADJ0_R = 8, ADJ2_R = 4;
int syn_tile(int spin, int kind) {
  return spin == 0 ? 32 * SYN0_R * (kind == LK_DE ? SYN0_DE_NW : SYN0_NW) : 32 * SYN2_R * SYN2_NW;
}
int adj_tile(int spin) { return 32 * (spin == 0 ? ADJ0_R : ADJ2_R); }
int64_t max_of(const std::vector<int64_t>& v) {
  int64_t m = 0;
  for (int64_t x : v) m = std::max(m, x);
  return m;
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
  for (int64_t i = 1; i < nm; ++i)
    HP_REQUIRE(mval[i] > mval[i - 1], HPSHT_ERR_UNSUPPORTED, "mval must be strictly ascending");
  pairs_ = build_pairs(lmax, nring, theta);
  npairs_ = (int64_t)pairs_.size();
  HP_REQUIRE(npairs_ < (1ll << 30) && nm < (1ll << 30), HPSHT_ERR_UNSUPPORTED, "problem too large");

  const size_t np = (size_t)npairs_;
  std::vector<double> cth(np), sth(np), sh(np), ch(np), x2(np), s2(np), ny2(np), thn(np);
  std::vector<int> ml0(np), ml2(np);
  std::vector<int64_t> a0(np), nrr(np), jN(np), jS(np);
  culled0_ = culled2_ = dense_ = 0;
  for (size_t p = 0; p < np; ++p) {
    const PairHost& q = pairs_[p];
    cth[p] = q.cth;
    sth[p] = q.sth;
    sh[p] = q.sh;
    ch[p] = q.ch;
    x2[p] = q.x2;
    s2[p] = q.s2;
    ny2[p] = -q.y2;
    thn[p] = q.theta;
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
  cr0_ = make_class_ranges(0, (int)np, thn.data());
  cr2_ = make_class_ranges(2, (int)np, thn.data());
  d_cth_.upload(cth.data(), np, stream);
  d_sth_.upload(sth.data(), np, stream);
  d_sh_.upload(sh.data(), np, stream);
  d_ch_.upload(ch.data(), np, stream);
  d_x2_.upload(x2.data(), np, stream);
  d_s2_.upload(s2.data(), np, stream);
  d_ny2_.upload(ny2.data(), np, stream);
  d_mlim0_.upload(ml0.data(), np, stream);
  d_mlim2_.upload(ml2.data(), np, stream);
  d_a0_.upload(a0.data(), np, stream);
  d_nrr_.upload(nrr.data(), np, stream);
  d_jN_.upload(jN.data(), np, stream);
  d_jS_.upload(jS.data(), np, stream);
  d_mval_.upload(mval_.data(), mval_.size(), stream);
  d_mstart_.upload(mstart_.data(), mstart_.size(), stream);
  HP_CUDA(cudaStreamSynchronize(stream));
  if (!ev0_) HP_CUDA(cudaEventCreate(&ev0_));
  if (!ev1_) HP_CUDA(cudaEventCreate(&ev1_));
  if (!ev_fork_) HP_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  if (!ev_join_) HP_CUDA(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
  if (!side_) HP_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
  blk_pair_[0].assign(2, 0);
  blk_pair_[0][1] = (int)np;
  blk_pair_[1].clear();
  have_[0] = have_[1] = false;
}

// ---- ring blocks -----------------------------------------------------------------------------
std::vector<int> cut_ring_blocks(const std::vector<PairHost>& pairs, int want, const double* slot_weight, bool equal) {
  std::vector<int> bp;
  const int np = (int)pairs.size();
  if (want < 1) want = 1;
  bp.push_back(0);
  // Preferred cut: multiples of 256 pairs -- the tile of every kernel but the spin-0 synthesis (512), whose tiles
  // then end half filled at some block boundaries -- and never inside an accuracy class's tile (the class ranges
  // are multiples of 256 too).  Block sizes grow from both ends towards the middle (1, 2, 3, ... units): what a
  // pipelined call cannot hide is the polar block's share of the alm copy and the equator-most block's map copy,
  // and a block's copy must fit behind the compute of its (smaller) neighbour towards the end (nside 4096, want 9:
  // 256 512 768 1024 3072 1024 768 512 256; measured against units of 512 and 128 pairs in profiles/r02_e2e.md).
  int64_t align = 256;
  if (const char* eu = getenv("HPSHT_BLOCK_UNIT")) align = std::max(32, atoi(eu));
  if (want > 1 && np >= 2 * align) {
    const int64_t units = (np + align - 1) / align;  // last unit may be partial
    std::vector<int64_t> left, right;
    int64_t rem = units;
    if (equal) {
      // copy-bound callers (several ranks sharing one host's memory system: the copy of a block takes about as
      // long as its compute) want every block's copy to fit behind the next block's compute: equal blocks
      const int64_t nb = std::min<int64_t>(want, units);
      for (int64_t b = 1; b < nb; ++b) left.push_back(units * b / nb - units * (b - 1) / nb);
      left.push_back(units - units * (nb - 1) / nb);
      rem = 0;
    }
    for (int64_t k = 1; !equal && (int64_t)(left.size() + right.size()) + 3 <= want && rem >= 2 * k + 1; ++k) {
      left.push_back(k);
      right.push_back(k);
      rem -= 2 * k;
    }
    if (left.empty() && units >= 2) {  // want == 2: everything but the equator-most unit, then that unit
      left.push_back(units - 1);
      rem = 1;
    } else if (rem > 0) {
      left.push_back(rem);
      rem = 0;
    }
    int64_t at = 0;
    for (int64_t u : left) {
      at += u;
      if (at * align < np) bp.push_back((int)(at * align));
    }
    for (size_t i = right.size(); i-- > 0;) {
      at += right[i];
      if (at * align < np && (int)(at * align) > bp.back()) bp.push_back((int)(at * align));
    }
  } else if (want > 1) {
    // small problems (tests): equal-weight blocks on warp-aligned boundaries
    std::vector<double> cum((size_t)np + 1, 0.0);
    for (int p = 0; p < np; ++p) {
      double w = slot_weight[pairs[(size_t)p].slotN];
      if (pairs[(size_t)p].slotS >= 0) w += slot_weight[pairs[(size_t)p].slotS];
      cum[(size_t)p + 1] = cum[(size_t)p] + w;
    }
    for (int b = 1; b < want; ++b) {
      const double target = cum[(size_t)np] * b / want;
      int p = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
      p = (p + 31) / 32 * 32;
      if (p > bp.back() && p < np) bp.push_back(p);
    }
  }
  bp.push_back(np);
  return bp;
}

int LegendreStage::setup_blocks(int want, const double* slot_weight, bool equal) {
  blk_pair_[1] = cut_ring_blocks(pairs_, want, slot_weight, equal);
  have_[0] = have_[1] = false;
  return nblocks(1);
}

std::vector<int64_t> LegendreStage::block_slot_list(int set, int b) const {
  std::vector<int64_t> v;
  const std::vector<int>& bp = blk_pair_[set];
  for (int p = bp[(size_t)b]; p < bp[(size_t)b + 1]; ++p) {
    v.push_back(pairs_[(size_t)p].slotN);
    if (pairs_[(size_t)p].slotS >= 0) v.push_back(pairs_[(size_t)p].slotS);
  }
  return v;
}

// number of local m's (a prefix, m ascending) that ring block b touches at all: the others are culled
// there (m > mlim of every pair of the block)
int LegendreStage::block_mi_end(int set, int b, int spin) const {
  const std::vector<int>& bp = blk_pair_[set];
  int ml = -1;
  for (int p = bp[(size_t)b]; p < bp[(size_t)b + 1]; ++p)
    ml = std::max(ml, spin == 0 ? pairs_[(size_t)p].mlim0 : pairs_[(size_t)p].mlim2);
  int n = 0;
  while (n < (int)nm_ && mval_[(size_t)n] <= ml) ++n;
  return n;
}

// Work lists of one (spin, block set): for every block and launch kind
//   synthesis: one item per (local m ascending == descending cost, tile of one class holding an active pair);
//   adjoint:   per local m the active pair range of the kind inside the block, cut into <= rows contiguous
//              multi-tile items, item g accumulating into partial-sum row g of that m.
// Active = mlim >= m.  Pairs are sorted pole -> equator; exotic theta lists whose mlim is not monotone are
// covered from the first to the last active pair.
void LegendreStage::build_items(int spin, int set, cudaStream_t stream) {
  const int si = spin ? 1 : 0;
  const ClassRanges& cr = spin ? cr2_ : cr0_;
  const std::vector<int>& bp = blk_pair_[set];
  const int nb = (int)bp.size() - 1;
  const std::vector<int64_t>& nstep = nstep_h_[si];
  const std::vector<int64_t>& off = off_h_[si];
  const int64_t total_steps = off[(size_t)nm_];
  const int64_t dp_steps = off[(size_t)mi_dp_[si]];
  auto mlim_of = [&](int p) { return spin ? pairs_[(size_t)p].mlim2 : pairs_[(size_t)p].mlim0; };
  std::vector<WorkItem> syn, adj;
  std::vector<int64_t> apo;
  std::vector<int>& rs = rng_syn_[si][set];
  std::vector<int>& ra = rng_adj_[si][set];
  rs.clear();
  ra.clear();
  const int atile = adj_tile(spin);
  for (int b = 0; b < nb; ++b) {
    for (int kind = 0; kind < LK_COUNT; ++kind) {
      rs.push_back((int)syn.size());
      ra.push_back((int)adj.size());
      // classes of this kind, pole -> equator
      std::vector<int> cls;
      if (kind == LK_STD) cls = {LC_Y, LC_X};
      else if (kind == LK_DP) cls = {LC_DP};
      else cls = {LC_DE};
      const int stile = syn_tile(spin, kind);
      // synthesis tiles of the block with their largest mlim
      struct Tile { int p0, np, cls, mmax; };
      std::vector<Tile> tiles;
      int ka = INT32_MAX, kb = -1;  // pair range of the kind inside the block
      for (int c : cls) {
        const int ca = std::max(cr.lo(c), bp[(size_t)b]), cb = std::min(cr.hi(c), bp[(size_t)b + 1]);
        if (cb <= ca) continue;
        ka = std::min(ka, ca);
        kb = std::max(kb, cb);
        for (int p0 = ca; p0 < cb; p0 += stile) {
          Tile t{p0, std::min(stile, cb - p0), c, -1};
          for (int p = p0; p < p0 + t.np; ++p) t.mmax = std::max(t.mmax, mlim_of(p));
          tiles.push_back(t);
        }
      }
      if (kb <= ka) continue;
      // prefix / suffix maxima of mlim over [ka, kb): first / last active pair of an m by bisection
      const int len = kb - ka;
      std::vector<int> pmax((size_t)len), smax((size_t)len);
      for (int i = 0; i < len; ++i) pmax[(size_t)i] = std::max(i ? pmax[(size_t)i - 1] : -1, mlim_of(ka + i));
      for (int i = len - 1; i >= 0; --i) smax[(size_t)i] = std::max(i + 1 < len ? smax[(size_t)i + 1] : -1, mlim_of(ka + i));
      const int rows = kind == LK_DP ? rows_dp_[si] : rows_std_[si];
      for (size_t mi = 0; mi < mval_.size(); ++mi) {
        if (nstep[mi] <= 0) continue;
        const int m = (int)mval_[mi];
        for (int t = (int)tiles.size() - 1; t >= 0; --t)
          if (tiles[(size_t)t].mmax >= m) syn.push_back(WorkItem{(int)mi, tiles[(size_t)t].p0, tiles[(size_t)t].np, tiles[(size_t)t].cls});
        if (pmax[(size_t)len - 1] < m) continue;
        const int first = (int)(std::lower_bound(pmax.begin(), pmax.end(), m) - pmax.begin());
        // smax is non-increasing: last index with smax >= m
        int lo = 0, hi = len - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) / 2;
          if (smax[(size_t)mid] >= m) lo = mid; else hi = mid - 1;
        }
        const int plo = ka + first / atile * atile, phi = ka + lo + 1;
        HP_REQUIRE(kind != LK_DP || (int)mi < mi_dp_[si], HPSHT_ERR_INTERNAL, "polar row table too small");
        const int nt = (phi - plo + atile - 1) / atile;
        const int G = std::max(1, std::min(rows, nt));
        for (int g = 0; g < G; ++g) {
          const int ta = (int)((int64_t)nt * g / G), tb = (int)((int64_t)nt * (g + 1) / G);
          WorkItem w;
          w.mi = (int)mi;
          w.p0 = plo + ta * atile;
          w.np = std::min(phi, plo + tb * atile) - w.p0;
          w.cls = cr.cls(w.p0);
          adj.push_back(w);
          apo.push_back(kind == LK_DP ? (int64_t)rows_std_[si] * total_steps + (int64_t)g * dp_steps + off[mi]
                                      : (int64_t)g * total_steps + off[mi]);
        }
      }
    }
  }
  rs.push_back((int)syn.size());
  ra.push_back((int)adj.size());
  d_items_syn_[si][set].upload(syn.data(), syn.size(), stream);
  d_items_adj_[si][set].upload(adj.data(), adj.size(), stream);
  d_partoff_[si][set].upload(apo.data(), apo.size(), stream);
  HP_CUDA(cudaStreamSynchronize(stream));  // host vectors die here
}

void LegendreStage::ensure_spin(int spin, cudaStream_t stream) {
  const int si = spin ? 1 : 0;
  if (have_[si]) return;
  const ClassRanges& cr = spin ? cr2_ : cr0_;
  // largest m any pair of a class keeps: the class's coefficient rows of larger m are never read
  int64_t mlimc[LC_COUNT];
  for (int c = 0; c < LC_COUNT; ++c) {
    mlimc[c] = -1;
    for (int p = cr.lo(c); p < cr.hi(c); ++p)
      mlimc[c] = std::max<int64_t>(mlimc[c], spin ? pairs_[(size_t)p].mlim2 : pairs_[(size_t)p].mlim0);
  }
  LegendreTables T;
  build_legendre_tables(lmax_, mval_.data(), nm_, LEG_CHUNK, spin == 0, spin != 0, T, spin == 0 ? mlimc : nullptr,
                        spin != 0 ? mlimc : nullptr);
  const std::vector<int64_t>& nstep = spin ? T.nstep2 : T.nstep0;
  const std::vector<int64_t>& off = spin ? T.off2 : T.off0;
  d_nstep_[si].upload(nstep.data(), nstep.size(), stream);
  d_off_[si].upload(off.data(), off.size(), stream);
  for (int c = 0; c < LC_COUNT; ++c) {
    const std::vector<int64_t>& co = spin ? T.coff2[c] : T.coff0[c];
    const std::vector<double>& cf = spin ? T.coef2[c] : T.coef0[c];
    d_coff_[si][c].upload(co.data(), co.size(), stream);
    d_coef_[si][c].upload(cf.data(), cf.size(), stream);
  }
  if (spin) {
    d_coefa2_.upload(T.coefa2.data(), T.coefa2.size(), stream);
    d_inv_[si].upload(T.sig2.data(), T.sig2.size(), stream);
    d_k0_[si].upload(T.kp2.data(), T.kp2.size(), stream);
    d_k1_[si].upload(T.km2.data(), T.km2.size(), stream);
  } else {
    d_inv_[si].upload(T.invs0.data(), T.invs0.size(), stream);
    d_k0_[si].upload(T.k0.data(), T.k0.size(), stream);
  }
  d_stream_[si].alloc((size_t)off[(size_t)nm_] * 4);
  d_stream_[si].zero(stream);
  nstep_h_[si] = nstep;
  off_h_[si] = off;
  // partial-sum rows: enough adjoint items to fill the GPU several times over even when a rank owns few m's
  rows_std_[si] = (int)std::min<int64_t>(64, std::max<int64_t>(1, (8192 + nm_ - 1) / std::max<int64_t>(nm_, 1)));
  mi_dp_[si] = 0;
  while (mi_dp_[si] < (int)nm_ && mval_[(size_t)mi_dp_[si]] <= mlimc[LC_DP]) ++mi_dp_[si];
  rows_dp_[si] = mi_dp_[si] == 0 ? 0 : (int)std::min<int64_t>(16, std::max<int64_t>(1, (2048 + mi_dp_[si] - 1) / mi_dp_[si]));
  build_items(spin, 0, stream);
  if (!blk_pair_[1].empty()) build_items(spin, 1, stream);
  HP_CUDA(cudaStreamSynchronize(stream));  // host vectors of T die here
  have_[si] = true;
}

void LegendreStage::fill_params(LegParams& P, int spin) const {
  const int si = spin ? 1 : 0;
  std::memset(&P, 0, sizeof(P));
  P.mval = d_mval_.p;
  P.nstep = d_nstep_[si].p;
  P.off = d_off_[si].p;
  for (int c = 0; c < LC_COUNT; ++c) {
    P.coff[c] = d_coff_[si][c].p;
    P.coef[c] = d_coef_[si][c].p;
  }
  P.coefa = d_coefa2_.p;
  P.k0 = d_k0_[si].p;
  P.k1 = d_k1_[si].p;
  if (spin == 0) {
    P.yq[LC_X] = d_x2_.p;
    P.yq[LC_Y] = d_s2_.p;
    P.yq[LC_DP] = d_s2_.p;
    P.yq[LC_DE] = d_x2_.p;
  } else {
    P.yq[LC_X] = d_cth_.p;
    P.yq[LC_Y] = d_ny2_.p;
    P.yq[LC_DP] = d_ny2_.p;
    P.yq[LC_DE] = d_cth_.p;
  }
  P.cth = d_cth_.p;
  P.sth = d_sth_.p;
  P.sh = d_sh_.p;
  P.ch = d_ch_.p;
  P.mlim = spin ? d_mlim2_.p : d_mlim0_.p;
  P.a0 = d_a0_.p;
  P.nrr = d_nrr_.p;
  P.jN = d_jN_.p;
  P.jS = d_jS_.p;
  P.nm = nm_;
  P.n_y = (spin ? cr2_ : cr0_).n_y;
  P.stream = d_stream_[si].p;
  P.part = d_part_.p;
}

// the small difference-form launches of a block go to a side stream next to the standard launch
void LegendreStage::fork(cudaStream_t stream) {
  HP_CUDA(cudaEventRecord(ev_fork_, stream));
  HP_CUDA(cudaStreamWaitEvent(side_, ev_fork_, 0));
}
void LegendreStage::join(cudaStream_t stream) {
  HP_CUDA(cudaEventRecord(ev_join_, side_));
  HP_CUDA(cudaStreamWaitEvent(stream, ev_join_, 0));
}

// ---- synthesis ------------------------------------------------------------------------------
void LegendreStage::alm2leg_zero(double* d_leg, int64_t leg_elems, int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  ensure_spin(spin, stream);
  // (m, ring) combinations beyond mlim are never visited by a CTA: their leg is exactly 0
  HP_CUDA(cudaMemsetAsync(d_leg, 0, (size_t)leg_elems * 2 * sizeof(double), stream));
  first_block_ = true;
}

// alm -> prepared per-step stream for the local m's [mi0, mi1)
void LegendreStage::alm2leg_prep(const double* d_alm, int64_t alm_cstride, int ncomp, int mi0, int mi1,
                                 cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  const int si = spin ? 1 : 0;
  if (mi1 <= mi0) return;
  const int TB = 256;
  const int64_t mx = max_of(nstep_h_[si]);
  if (mx <= 0) return;
  dim3 grid((unsigned)((mx + TB - 1) / TB), (unsigned)(mi1 - mi0));
  if (spin == 0)
    k_prep_s0<<<grid, TB, 0, stream>>>(reinterpret_cast<const double2*>(d_alm), d_mval_.p, d_mstart_.p,
                                       d_nstep_[0].p, d_off_[0].p, d_inv_[0].p, lmax_,
                                       reinterpret_cast<Strm*>(d_stream_[0].p), mi0);
  else
    k_prep_s2<<<grid, TB, 0, stream>>>(reinterpret_cast<const double2*>(d_alm),
                                       reinterpret_cast<const double2*>(d_alm) + alm_cstride, d_mval_.p,
                                       d_mstart_.p, d_nstep_[1].p, d_off_[1].p, d_inv_[1].p,
                                       reinterpret_cast<Strm*>(d_stream_[1].p), mi0);
  ++launches_;
  HP_CUDA(cudaGetLastError());
}

void LegendreStage::alm2leg_block(int set, int b, double* d_leg, int ncomp, cudaStream_t stream, bool last) {
  const int spin = ncomp == 1 ? 0 : 2;
  const int si = spin ? 1 : 0;
  const std::vector<int>& rg = rng_syn_[si][set];
  LegParams P;
  fill_params(P, spin);
  P.leg = d_leg;
  auto count = [&](int kind) { return rg[(size_t)(b * LK_COUNT + kind) + 1] - rg[(size_t)(b * LK_COUNT + kind)]; };
  const bool side = count(LK_STD) > 0 && (count(LK_DP) > 0 || count(LK_DE) > 0);
  if (first_block_) cudaEventRecord(ev0_, stream);
  first_block_ = false;
  if (side) fork(stream);
  for (int kind : {LK_DP, LK_DE, LK_STD}) {
    const int n = count(kind);
    if (n <= 0) continue;
    cudaStream_t st = (side && kind != LK_STD) ? side_ : stream;
    P.items = d_items_syn_[si][set].p + rg[(size_t)(b * LK_COUNT + kind)];
    if (spin == 0) {
      if (kind == LK_STD) k_alm2leg_s0<SYN0_R, SYN0_NW, 0><<<(unsigned)n, (SYN0_NW + 1) * 32, 0, st>>>(P);
      else if (kind == LK_DP) k_alm2leg_s0<SYN0_R, SYN0_NW, 1><<<(unsigned)n, (SYN0_NW + 1) * 32, 0, st>>>(P);
      else k_alm2leg_s0<SYN0_R, SYN0_DE_NW, 2><<<(unsigned)n, (SYN0_DE_NW + 1) * 32, 0, st>>>(P);
    } else {
      if (kind == LK_STD) k_alm2leg_s2<SYN2_R, SYN2_NW, 0><<<(unsigned)n, (SYN2_NW + 1) * 32, 0, st>>>(P);
      else k_alm2leg_s2<SYN2_R, SYN2_NW, 1><<<(unsigned)n, (SYN2_NW + 1) * 32, 0, st>>>(P);
    }
    ++launches_;
    HP_CUDA(cudaGetLastError());
  }
  if (side) join(stream);
  if (last) cudaEventRecord(ev1_, stream);
}

void LegendreStage::alm2leg(const double* d_alm, int64_t alm_cstride, double* d_leg, int64_t leg_elems,
                            int ncomp, cudaStream_t stream) {
  alm2leg_zero(d_leg, leg_elems, ncomp, stream);
  alm2leg_prep(d_alm, alm_cstride, ncomp, 0, (int)nm_, stream);
  alm2leg_block(0, 0, d_leg, ncomp, stream, true);
}

// ---- adjoint --------------------------------------------------------------------------------
void LegendreStage::leg2alm_begin(int ncomp, cudaStream_t stream) {
  const int spin = ncomp == 1 ? 0 : 2;
  const int si = spin ? 1 : 0;
  ensure_spin(spin, stream);
  const int64_t total = off_h_[si][(size_t)nm_], dp = off_h_[si][(size_t)mi_dp_[si]];
  const size_t need = ((size_t)rows_std_[si] * total + (size_t)rows_dp_[si] * dp) * 4;
  d_part_.ensure(need);
  // every item adds to its row: start from zero
  if (need) HP_CUDA(cudaMemsetAsync(d_part_.p, 0, need * sizeof(double), stream));
  first_block_ = true;
}

void LegendreStage::leg2alm_block(int set, int b, const double* d_leg, int ncomp, cudaStream_t stream, bool last) {
  const int spin = ncomp == 1 ? 0 : 2;
  const int si = spin ? 1 : 0;
  const std::vector<int>& rg = rng_adj_[si][set];
  LegParams P;
  fill_params(P, spin);
  P.leg = const_cast<double*>(d_leg);
  auto count = [&](int kind) { return rg[(size_t)(b * LK_COUNT + kind) + 1] - rg[(size_t)(b * LK_COUNT + kind)]; };
  // the polar launch has its own partial-sum rows and runs next to the other two, which share rows and
  // therefore follow each other on the caller's stream
  const bool side = count(LK_DP) > 0 && (count(LK_STD) > 0 || count(LK_DE) > 0);
  if (first_block_) cudaEventRecord(ev0_, stream);
  first_block_ = false;
  if (side) fork(stream);
  for (int kind : {LK_DP, LK_DE, LK_STD}) {
    const int n = count(kind);
    if (n <= 0) continue;
    cudaStream_t st = (side && kind == LK_DP) ? side_ : stream;
    const int first = rg[(size_t)(b * LK_COUNT + kind)];
    P.items = d_items_adj_[si][set].p + first;
    P.partoff = d_partoff_[si][set].p + first;
    if (spin == 0) {
      if (kind == LK_STD) k_leg2alm_s0<ADJ0_R, 0><<<(unsigned)n, 32, 0, st>>>(P);
      else if (kind == LK_DP) k_leg2alm_s0<ADJ0_R, 1><<<(unsigned)n, 32, 0, st>>>(P);
      else k_leg2alm_s0<ADJ0_R, 2><<<(unsigned)n, 32, 0, st>>>(P);
    } else {
      if (kind == LK_STD) k_leg2alm_s2<ADJ2_R, 0><<<(unsigned)n, 32, 0, st>>>(P);
      else k_leg2alm_s2<ADJ2_R, 1><<<(unsigned)n, 32, 0, st>>>(P);
    }
    ++launches_;
    HP_CUDA(cudaGetLastError());
  }
  if (side) join(stream);
  if (last) cudaEventRecord(ev1_, stream);
}

void LegendreStage::leg2alm_end(double* d_alm, int64_t alm_cstride, int ncomp, cudaStream_t stream, int mi0,
                                int mi1) {
  const int spin = ncomp == 1 ? 0 : 2;
  const int si = spin ? 1 : 0;
  if (mi1 < 0) mi1 = (int)nm_;
  if (mi1 <= mi0) return;
  const int TB = 256;
  const int64_t mx = std::max<int64_t>(max_of(nstep_h_[si]), 1);  // spin 2 also zeroes l < 2
  dim3 grid((unsigned)((mx + TB - 1) / TB), (unsigned)(mi1 - mi0));
  RowSet S;
  S.part = reinterpret_cast<const Strm*>(d_part_.p);
  S.nstd = rows_std_[si];
  S.ndp = rows_dp_[si];
  S.mi_dp = mi_dp_[si];
  S.std_stride = off_h_[si][(size_t)nm_];
  S.dp_base = (int64_t)rows_std_[si] * S.std_stride;
  S.dp_stride = off_h_[si][(size_t)mi_dp_[si]];
  if (spin == 0)
    k_post_s0<<<grid, TB, 0, stream>>>(S, d_mval_.p, d_mstart_.p, d_nstep_[0].p, d_off_[0].p, d_inv_[0].p, lmax_,
                                       reinterpret_cast<double2*>(d_alm), mi0);
  else
    k_post_s2<<<grid, TB, 0, stream>>>(S, d_mval_.p, d_mstart_.p, d_nstep_[1].p, d_off_[1].p, d_inv_[1].p, lmax_,
                                       reinterpret_cast<double2*>(d_alm),
                                       reinterpret_cast<double2*>(d_alm) + alm_cstride, mi0);
  ++launches_;
  HP_CUDA(cudaGetLastError());
}

void LegendreStage::leg2alm(const double* d_leg, double* d_alm, int64_t alm_cstride, int ncomp,
                            cudaStream_t stream) {
  leg2alm_begin(ncomp, stream);
  leg2alm_block(0, 0, d_leg, ncomp, stream, true);
  leg2alm_end(d_alm, alm_cstride, ncomp, stream);
}

}  // namespace hpsht
