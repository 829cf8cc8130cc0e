/*
 * sht_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into, imported by or called from the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may use anything under oracle/).
 *
 * Extended-precision (x87 80-bit long double; __float128 for the brute-force part and the polar rings) CPU
 * restatement of the arithmetic that HealpixMPI.jl's hot path delegates to Ducc0:
 *     alm2leg! / leg2map! / map2leg! / leg2alm!        (reference src/sht.jl:85,93,169,178)
 * in the data layouts the reference passes across that seam (src/sht.jl:67-69,97-98,182-183).
 *
 * PARITY STATUS: "parity unpinned" at the 1e-12 level.  The arithmetic lives in the
 * un-vendored third-party dependency Ducc0.jl (Project.toml:11, compat "0.1.0" at
 * Project.toml:17, binary ducc0_jll version unpinned because the reference ships no
 * Manifest.toml); neither Julia, MPI nor ducc0 exist in the build image and the reference's
 * tests hold no golden vectors (they compare against serial Healpix.jl on unseeded random
 * input at rtol 1.5e-8: test/test_MPI_a2m.jl:38, test_MPI_adj_a2m.jl:40,
 * test_MPI_a2m_pol.jl:47-48, test_MPI_adj_a2m_pol.jl:52-53).  What pins this oracle instead:
 *   (1) the published definitions it restates (orthonormal associated Legendre functions with
 *       Condon-Shortley phase; Goldberg spin-weighted harmonics; HEALPix E/B -> Q/U convention
 *       2a = -(E+iB), -2a = -(E-iB));
 *   (2) a brute-force __float128 double sum over explicit sY_lm (orc_brute_synth) that shares
 *       no recurrence with the staged path;
 *   (3) analytic known answers (a00 = sqrt(4pi) -> map == 1; single-alm maps; E22 closed form);
 *   (4) third-party implementations: scipy.special.sph_harm_y (double), mpmath.spherharm at 40 digits (spin 0,
 *       agreement 5e-18), sympy's exact Wigner small-d for the spin-weighted functions and their sign convention
 *       (agreement 2e-17) -- tests/test_oracle.py;
 *   (4b) its own accuracy at l = 8191, measured against the same recurrences in __float128 (orc_lambda_quad):
 *       <= 2e-14 everywhere, ~1e-18 on the rings next to the poles, which run in __float128 (see use_quad);
 *   (5) the adjointness identity <f, Y a> = <Y^T f, a>_w with the localdot weighting
 *       (src/alm.jl:634-644).
 * Algorithms here are deliberately DIFFERENT from the product's (plain three-term recurrences
 * in l, no north/south pairing, direct O(n*m) ring sums instead of FFTs) so that agreement is
 * evidence, not tautology.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp ... -lquadmath -lm).
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long double ld;
typedef __float128 q128;

#define ORC_PI 3.14159265358979323846264338327950288419716939937510L

/* cos(theta) as an unevaluated sum hi + lo of two long doubles (lo from a __float128 evaluation).
 * Why: the recurrences multiply the running value by x = cos(theta), and d(lambda)/dx ~ l / sin(theta), so the
 * 2.7e-20 rounding of a plain long-double cosine is amplified to 1e-12 relative at l = 8191 on the innermost HEALPix
 * rings (sin(theta) ~ 2e-4) -- the very accuracy this file is used to certify.  tests/test_oracle.py measured it
 * against mpmath (6.7e-17 at l = 48, theta = 0.02) before this was added. */
static void cos_hi_lo(ld th, ld* hi, ld* lo) {
  __float128 c = cosq((__float128)th);
  *hi = (ld)c;
  *lo = (ld)(c - (__float128)*hi);
}

/* Near the poles even the compensated long-double recurrence is not good enough: rounding errors injected at every
 * step grow like l^2 for l*theta < 1 (the second solution of the recurrence is singular at the pole), measured
 * 5.6e-13 relative at l = 8191 on the first ring of nside 4096 against the __float128 recurrence, 1.4e-14 on the third
 * ring, < 1e-15 from lmax*sin(theta) = 80 on.  Rings with lmax*sin(theta) < 32 (about 20 rings per pole at the
 * headline size; they carry few m) therefore run the recurrence itself in __float128. */
static int use_quad(ld th, int64_t lmax) { return fabsl(sinl(th)) * (ld)lmax < 64.0L; }

/* ------------------------------------------------------------------------------------------
 * HEALPix RING-scheme geometry of ring `ring` (1-based, 1..4*nside-1).  Restates what
 * ScatterMap! obtains from Healpix.getringinfo!/pix2ang (src/map.jl:177-183); formulas are the
 * standard HEALPix ones (Gorski et al. 2005) as summarised in SURVEY.md 8c.
 * ---------------------------------------------------------------------------------------- */
void orc_healpix_ring(int64_t nside, int64_t ring, int64_t* nphi, double* theta, double* phi0,
                      int64_t* first_pix) {
  int64_t nr = 4 * nside - 1;
  int64_t npix = 12 * nside * nside;
  int south = ring > 2 * nside;
  int64_t n = south ? (nr + 1 - ring) : ring; /* northern mirror ring index 1..2nside */
  ld th;
  int64_t np, fp;
  double p0;
  if (n < nside) { /* polar cap */
    ld t = (ld)n * (ld)n / (3.0L * (ld)nside * (ld)nside);
    ld cth = 1.0L - t;
    ld sth = sqrtl(t * (2.0L - t));
    th = atan2l(sth, cth);
    np = 4 * n;
    fp = 2 * n * (n - 1);
    p0 = (double)(ORC_PI / (ld)(4 * n));
  } else { /* equatorial belt */
    ld cth = (ld)(2 * nside - n) * 2.0L / (3.0L * (ld)nside);
    th = acosl(cth);
    np = 4 * nside;
    fp = 2 * nside * (nside - 1) + (n - nside) * 4 * nside;
    p0 = ((n - nside) % 2 == 0) ? (double)(ORC_PI / (ld)(4 * nside)) : 0.0;
  }
  if (south) {
    th = ORC_PI - th;
    fp = npix - fp - np;
  }
  *nphi = np;
  *theta = (double)th;
  *phi0 = p0;
  *first_pix = fp;
}

/* ------------------------------------------------------------------------------------------
 * Extended-exponent helper: value = v * 2^e, v a long double, renormalised on demand.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  ld v;
  int64_t e;
} xld;

static xld x_norm(xld a) {
  if (a.v == 0.0L) {
    a.e = 0;
    return a;
  }
  int ex;
  a.v = frexpl(a.v, &ex);
  a.e += ex;
  return a;
}
static xld x_mul(xld a, xld b) {
  xld r = {a.v * b.v, a.e + b.e};
  return x_norm(r);
}
static xld x_from(ld v) {
  xld r = {v, 0};
  return x_norm(r);
}
static xld x_powi(ld base, int64_t n) { /* base^n, n >= 0, binary powering */
  xld r = x_from(1.0L), b = x_from(base);
  while (n > 0) {
    if (n & 1) r = x_mul(r, b);
    b = x_mul(b, b);
    n >>= 1;
  }
  return r;
}

/* Range management for the upward recurrences: true value = v * 2^(SCALE_STEP*sc). */
#define SCALE_STEP 8192
static const ld BIG = 0x1p+4096L;
static const ld RESCALE_DOWN = 0x1p-8192L;

/* spin 0: out[l] = lambda_lm(theta) for l = m..lmax (orthonormal, Condon-Shortley), 0 for l<m.
 * Recurrence: x*lam_l = eps_{l+1} lam_{l+1} + eps_l lam_{l-1}, eps_l = sqrt((l^2-m^2)/(4l^2-1)).
 * Values whose magnitude is below 2^-4096 are returned as 0. */
static void lam_row_spin0(int64_t lmax, int64_t m, ld th, ld* out) {
  ld x, xlo, sth = sinl(th);
  cos_hi_lo(th, &x, &xlo);
  for (int64_t l = 0; l < m && l <= lmax; ++l) out[l] = 0.0L;
  if (m > lmax) return;
  /* lam_mm = (-1)^m sqrt((2m+1)/(4pi)) * sqrt(prod_{k=1..m} (2k-1)/(2k)) * sin^m */
  xld c = x_from(sqrtl((ld)(2 * m + 1) / (4.0L * ORC_PI)));
  {
    xld p = x_from(1.0L);
    for (int64_t k = 1; k <= m; ++k) p = x_mul(p, x_from((ld)(2 * k - 1) / (ld)(2 * k)));
    /* sqrt of the product: handle odd exponent */
    if (p.e & 1) {
      p.v *= 2.0L;
      p.e -= 1;
    }
    xld s = {sqrtl(p.v), p.e / 2};
    c = x_mul(c, x_norm(s));
  }
  if (m & 1) c.v = -c.v;
  xld st = x_mul(c, x_powi(sth, m));
  if (st.v == 0.0L) {
    for (int64_t l = m; l <= lmax; ++l) out[l] = 0.0L;
    return;
  }
  /* choose sc <= 0 with st.e - SCALE_STEP*sc in (-SCALE_STEP/2, SCALE_STEP/2] */
  int64_t sc = 0;
  if (st.e < -SCALE_STEP / 2) sc = -((-st.e + SCALE_STEP / 2 - 1) / SCALE_STEP);
  ld cur = ldexpl(st.v, (int)(st.e - SCALE_STEP * sc));
  ld prev = 0.0L;
  ld m2 = (ld)m * (ld)m;
  if (use_quad(th, lmax)) { /* same recurrence, same range management, 113-bit arithmetic */
    q128 qx = cosq((q128)th), qc = cur, qp = 0, qm2 = (q128)m * m;
    for (int64_t l = m; l <= lmax; ++l) {
      out[l] = (sc == 0) ? (ld)qc : 0.0L;
      q128 fl = l, fl1 = l + 1;
      q128 el = (l == m) ? 0 : sqrtq((fl * fl - qm2) / (4 * fl * fl - 1));
      q128 el1 = sqrtq((fl1 * fl1 - qm2) / (4 * fl1 * fl1 - 1));
      q128 nx = (qx * qc - el * qp) / el1;
      qp = qc;
      qc = nx;
      if (fabsq(qc) > (q128)BIG && sc < 0) {
        qc *= (q128)RESCALE_DOWN;
        qp *= (q128)RESCALE_DOWN;
        sc += 1;
      }
    }
    return;
  }
  for (int64_t l = m; l <= lmax; ++l) {
    out[l] = (sc == 0) ? cur : 0.0L;
    ld l1 = (ld)(l + 1);
    ld eps_l = (l == m) ? 0.0L : sqrtl(((ld)l * (ld)l - m2) / (4.0L * (ld)l * (ld)l - 1.0L));
    ld eps_l1 = sqrtl((l1 * l1 - m2) / (4.0L * l1 * l1 - 1.0L));
    ld nxt = ((x * cur - eps_l * prev) + xlo * cur) / eps_l1;
    prev = cur;
    cur = nxt;
    if (fabsl(cur) > BIG && sc < 0) {
      cur *= RESCALE_DOWN;
      prev *= RESCALE_DOWN;
      sc += 1;
    }
  }
}

/* spin +-2: out[l] = {s}lambda_lm(theta), Goldberg convention:
 *   sY_lm = (-1)^m sqrt((l+m)!(l-m)!(2l+1)/(4pi (l+s)!(l-s)!)) sin^{2l}(th/2)
 *           * sum_r C(l-s,r) C(l+s,r+s-m) (-1)^{l-r-s} e^{i m phi} cot^{2r+s-m}(th/2)
 * Three-term recurrence (verified numerically against that sum):
 *   C_{l+1} lam_{l+1} = (x + s m/(l(l+1))) lam_l - C_l lam_{l-1},
 *   C_l = sqrt((l^2-m^2)(l^2-s^2)/(l^2 (4l^2-1))).
 * Start at l0 = max(m,2) with the single-term closed forms of the sum. Requires m >= 0. */
static void lam_row_spin2(int s, int64_t lmax, int64_t m, ld th, ld* out) {
  ld x, xlo;
  cos_hi_lo(th, &x, &xlo);
  ld sh = sinl(0.5L * th), ch = cosl(0.5L * th);
  int64_t l0 = m > 2 ? m : 2;
  for (int64_t l = 0; l < l0 && l <= lmax; ++l) out[l] = 0.0L;
  if (l0 > lmax) return;
  xld st;
  if (m >= 2) {
    /* (-1)^m sqrt((2m+1)/(4pi) (2m)!/((m+s)!(m-s)!)) sin^{m+s}(th/2) cos^{m-s}(th/2) */
    /* (2m)!/((m+2)!(m-2)!) = C(2m,m) * m(m-1)/((m+1)(m+2)) ; C(2m,m) = prod_{k=1..m} (m+k)/k */
    xld p = x_from(1.0L);
    for (int64_t k = 1; k <= m; ++k) p = x_mul(p, x_from((ld)(m + k) / (ld)k));
    p = x_mul(p, x_from((ld)m * (ld)(m - 1) / ((ld)(m + 1) * (ld)(m + 2))));
    p = x_mul(p, x_from((ld)(2 * m + 1) / (4.0L * ORC_PI)));
    if (p.e & 1) {
      p.v *= 2.0L;
      p.e -= 1;
    }
    xld c = {sqrtl(p.v), p.e / 2};
    c = x_norm(c);
    if (m & 1) c.v = -c.v;
    st = x_mul(c, x_mul(x_powi(sh, m + s), x_powi(ch, m - s)));
  } else {
    /* l0 = 2 > m: sY_2m from the Goldberg sum (single surviving term):
       s=+2: (-1)^m pref C(4,2-m) cos^{2-m}(th/2) sin^{2+m}(th/2)
       s=-2:        pref C(4,2+m) cos^{2+m}(th/2) sin^{2-m}(th/2)
       pref = sqrt((2+m)!(2-m)! 5/(4pi 4!)) */
    ld fact[5] = {1.0L, 1.0L, 2.0L, 6.0L, 24.0L};
    ld binom4[5] = {1.0L, 4.0L, 6.0L, 4.0L, 1.0L};
    ld pref = sqrtl(fact[2 + m] * fact[2 - m] * 5.0L / (4.0L * ORC_PI * 24.0L));
    ld v;
    if (s > 0) {
      v = pref * binom4[2 - m] * powl(ch, (ld)(2 - m)) * powl(sh, (ld)(2 + m));
      if (m & 1) v = -v;
    } else {
      v = pref * binom4[2 + m] * powl(ch, (ld)(2 + m)) * powl(sh, (ld)(2 - m));
    }
    st = x_from(v);
  }
  if (st.v == 0.0L) {
    for (int64_t l = l0; l <= lmax; ++l) out[l] = 0.0L;
    return;
  }
  int64_t sc = 0;
  if (st.e < -SCALE_STEP / 2) sc = -((-st.e + SCALE_STEP / 2 - 1) / SCALE_STEP);
  ld cur = ldexpl(st.v, (int)(st.e - SCALE_STEP * sc));
  ld prev = 0.0L;
  ld m2 = (ld)m * (ld)m, s2 = 4.0L;
  if (use_quad(th, lmax)) {
    q128 qx = cosq((q128)th), qc = cur, qp = 0, qm2 = (q128)m * m, qs2 = 4;
    for (int64_t l = l0; l <= lmax; ++l) {
      out[l] = (sc == 0) ? (ld)qc : 0.0L;
      q128 fl = l, fl1 = l + 1;
      q128 Cl = (l == l0) ? 0 : sqrtq((fl * fl - qm2) * (fl * fl - qs2) / (fl * fl * (4 * fl * fl - 1)));
      q128 Cl1 = sqrtq((fl1 * fl1 - qm2) * (fl1 * fl1 - qs2) / (fl1 * fl1 * (4 * fl1 * fl1 - 1)));
      q128 nx = ((qx + (q128)s * (q128)m / (fl * fl1)) * qc - Cl * qp) / Cl1;
      qp = qc;
      qc = nx;
      if (fabsq(qc) > (q128)BIG && sc < 0) {
        qc *= (q128)RESCALE_DOWN;
        qp *= (q128)RESCALE_DOWN;
        sc += 1;
      }
    }
    return;
  }
  for (int64_t l = l0; l <= lmax; ++l) {
    out[l] = (sc == 0) ? cur : 0.0L;
    ld fl = (ld)l, fl1 = (ld)(l + 1);
    ld Cl = (l == l0) ? 0.0L
                      : sqrtl((fl * fl - m2) * (fl * fl - s2) / (fl * fl * (4.0L * fl * fl - 1.0L)));
    ld Cl1 = sqrtl((fl1 * fl1 - m2) * (fl1 * fl1 - s2) / (fl1 * fl1 * (4.0L * fl1 * fl1 - 1.0L)));
    ld nxt = (((x + (ld)s * (ld)m / (fl * fl1)) * cur - Cl * prev) + xlo * cur) / Cl1;
    prev = cur;
    cur = nxt;
    if (fabsl(cur) > BIG && sc < 0) {
      cur *= RESCALE_DOWN;
      prev *= RESCALE_DOWN;
      sc += 1;
    }
  }
}

/* public: spin in {0, +2, -2}; out has lmax+1 entries */
void orc_lambda(int s, int64_t lmax, int64_t m, double theta, ld* out) {
  if (s == 0)
    lam_row_spin0(lmax, m, (ld)theta, out);
  else
    lam_row_spin2(s, lmax, m, (ld)theta, out);
}

/* ------------------------------------------------------------------------------------------
 * Optional: the ring-culling rule of libsharp / ducc0 [UPSTREAM, published algorithm; SURVEY 8d]:
 * (m, ring) combinations with m > mlim(theta) are skipped by the reference's backend (their leg is
 * exactly 0 / they do not contribute to a_lm).  With white-noise a_lm at lmax = 8191 the skipped
 * terms are ~1e-11 of a ring's norm, i.e. larger than the 1e-12 parity target, so "the exact sum"
 * and "what the reference computes" differ by that much.  orc_set_cut(1) makes the oracle apply the
 * same rule so that the arithmetic error of an implementation can be separated from the cut.
 * ---------------------------------------------------------------------------------------- */
static int g_apply_cut = 0;
void orc_set_cut(int on) { g_apply_cut = on; }
int64_t orc_mlim(int64_t lmax, int spin, double theta) {
  double sth = (double)sinl((ld)theta), cth = (double)cosl((ld)theta);
  double ofs = lmax * 0.01;
  if (ofs < 100.) ofs = 100.;
  double b = -2. * spin * fabs(cth);
  double t1 = lmax * sth + ofs;
  double c = (double)spin * spin - t1 * t1;
  double discr = b * b - 4. * c;
  if (discr <= 0) return lmax;
  double res = (-b + sqrt(discr)) / 2.;
  if (res > lmax) res = (double)lmax;
  return (int64_t)(res + 0.5);
}

/* ------------------------------------------------------------------------------------------
 * alm2leg  (what Ducc0.Sht.alm2leg! computes at src/sht.jl:85)
 *   alm : double complex, alm[comp*alm_cstride + mstart[mi] + l]   (mstart 0-based)
 *   leg : long double complex (re,im interleaved), leg[((comp*nring + ir)*nm + mi)]
 *   spin 0: leg = sum_l a_lm lam_lm ; (Im a_l0 participates here exactly as in ducc: the
 *           real-part selection happens in leg2map)
 *   spin 2: leg_Q = -sum_l (E F+ + i B F-), leg_U = -sum_l (B F+ - i E F-),
 *           F+- = (2lam +- -2lam)/2
 * ---------------------------------------------------------------------------------------- */
void orc_alm2leg(int spin, int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                 const double* alm, int64_t alm_cstride, int64_t nring, const double* theta,
                 ld* leg) {
  int ncomp = spin == 0 ? 1 : 2;
#pragma omp parallel
  {
    ld* lp = (ld*)malloc(sizeof(ld) * (size_t)(lmax + 1));
    ld* lm = (ld*)malloc(sizeof(ld) * (size_t)(lmax + 1));
#pragma omp for collapse(2) schedule(dynamic, 4)
    for (int64_t ir = 0; ir < nring; ++ir) {
      for (int64_t mi = 0; mi < nm; ++mi) {
        int64_t m = mval[mi];
        ld th = (ld)theta[ir];
        if (g_apply_cut && m > orc_mlim(lmax, spin, theta[ir])) {
          for (int c = 0; c < (spin == 0 ? 1 : 2); ++c) {
            ld* o = leg + 2 * ((c * nring + ir) * nm + mi);
            o[0] = 0;
            o[1] = 0;
          }
          continue;
        }
        if (spin == 0) {
          lam_row_spin0(lmax, m, th, lp);
          ld sr = 0, si = 0;
          const double* a = alm + 2 * mstart[mi];
          for (int64_t l = m; l <= lmax; ++l) {
            sr += (ld)a[2 * l] * lp[l];
            si += (ld)a[2 * l + 1] * lp[l];
          }
          ld* o = leg + 2 * ((0 * nring + ir) * nm + mi);
          o[0] = sr;
          o[1] = si;
        } else {
          lam_row_spin2(+2, lmax, m, th, lp);
          lam_row_spin2(-2, lmax, m, th, lm);
          const double* e = alm + 2 * mstart[mi];
          const double* b = alm + 2 * (alm_cstride + mstart[mi]);
          ld qr = 0, qi = 0, ur = 0, ui = 0;
          int64_t l0 = m > 2 ? m : 2;
          for (int64_t l = l0; l <= lmax; ++l) {
            ld fp = 0.5L * (lp[l] + lm[l]), fm = 0.5L * (lp[l] - lm[l]);
            ld er = e[2 * l], ei = e[2 * l + 1], br = b[2 * l], bi = b[2 * l + 1];
            /* Q = -(E F+ + i B F-) ; i*B = (-bi, br) */
            qr -= er * fp - bi * fm;
            qi -= ei * fp + br * fm;
            /* U = -(B F+ - i E F-) ; i*E = (-ei, er) */
            ur -= br * fp + ei * fm;
            ui -= bi * fp - er * fm;
          }
          ld* oq = leg + 2 * ((0 * nring + ir) * nm + mi);
          ld* ou = leg + 2 * ((1 * nring + ir) * nm + mi);
          oq[0] = qr;
          oq[1] = qi;
          ou[0] = ur;
          ou[1] = ui;
        }
      }
    }
    free(lp);
    free(lm);
  }
  (void)ncomp;
}

/* ------------------------------------------------------------------------------------------
 * leg2alm  (Ducc0.Sht.leg2alm! at src/sht.jl:178): exact transpose.
 *   spin 0: a_lm = sum_ir leg[mi,ir] lam_lm(theta_ir)
 *   spin 2: E = -sum_ir (F+ legQ + i F- legU), B = -sum_ir (F+ legU - i F- legQ)
 *   alm_out: long double complex (re,im), same indexing as alm; entries l<spin zeroed; only the
 *   entries l = m..lmax of each owned m are written.
 * ---------------------------------------------------------------------------------------- */
void orc_leg2alm(int spin, int64_t lmax, int64_t nm, const int64_t* mval, const int64_t* mstart,
                 const ld* leg, int64_t nring, const double* theta, ld* alm_out,
                 int64_t alm_cstride) {
#pragma omp parallel
  {
    ld* lp = (ld*)malloc(sizeof(ld) * (size_t)(lmax + 1));
    ld* lm = (ld*)malloc(sizeof(ld) * (size_t)(lmax + 1));
#pragma omp for schedule(dynamic, 1)
    for (int64_t mi = 0; mi < nm; ++mi) {
      int64_t m = mval[mi];
      ld* e = alm_out + 2 * mstart[mi];
      ld* b = alm_out + 2 * (alm_cstride + mstart[mi]);
      for (int64_t l = m; l <= lmax; ++l) {
        e[2 * l] = e[2 * l + 1] = 0;
        if (spin) b[2 * l] = b[2 * l + 1] = 0;
      }
      for (int64_t ir = 0; ir < nring; ++ir) {
        ld th = (ld)theta[ir];
        if (g_apply_cut && m > orc_mlim(lmax, spin, theta[ir])) continue;
        if (spin == 0) {
          lam_row_spin0(lmax, m, th, lp);
          const ld* g = leg + 2 * ((0 * nring + ir) * nm + mi);
          for (int64_t l = m; l <= lmax; ++l) {
            e[2 * l] += g[0] * lp[l];
            e[2 * l + 1] += g[1] * lp[l];
          }
        } else {
          lam_row_spin2(+2, lmax, m, th, lp);
          lam_row_spin2(-2, lmax, m, th, lm);
          const ld* q = leg + 2 * ((0 * nring + ir) * nm + mi);
          const ld* u = leg + 2 * ((1 * nring + ir) * nm + mi);
          int64_t l0 = m > 2 ? m : 2;
          for (int64_t l = l0; l <= lmax; ++l) {
            ld fp = 0.5L * (lp[l] + lm[l]), fm = 0.5L * (lp[l] - lm[l]);
            /* E -= F+ Q + i F- U ; i*U = (-ui, ur) */
            e[2 * l] -= fp * q[0] - fm * u[1];
            e[2 * l + 1] -= fp * q[1] + fm * u[0];
            /* B -= F+ U - i F- Q ; i*Q = (-qi, qr) */
            b[2 * l] -= fp * u[0] + fm * q[1];
            b[2 * l + 1] -= fp * u[1] - fm * q[0];
          }
        }
      }
    }
    free(lp);
    free(lm);
  }
}

/* ------------------------------------------------------------------------------------------
 * leg2map (Ducc0.Sht.leg2map! at src/sht.jl:93), direct summation:
 *   map[comp*map_cstride + ringstart[ir] + j] =
 *        Re sum_{m=0}^{nm-1} c_m leg[m,ir] exp(i m (phi0[ir] + 2 pi j / nphi[ir])), c_0=1, c_m=2
 *   (for m = 0 only Re(leg) contributes, by construction of the real part)
 *   leg: long double complex [(comp*nring + ir)*nm + m]; map out: long double.
 *   ring_sel (may be NULL): if given, only rings with ring_sel[ir] != 0 are computed.
 * ---------------------------------------------------------------------------------------- */
void orc_leg2map(int ncomp, int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                 const int64_t* ringstart, const ld* leg, ld* map, int64_t map_cstride,
                 const unsigned char* ring_sel) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t ir = 0; ir < nring; ++ir) {
    if (ring_sel && !ring_sel[ir]) continue;
    int64_t n = nphi[ir];
    ld* wr = (ld*)malloc(sizeof(ld) * (size_t)n);
    ld* wi = (ld*)malloc(sizeof(ld) * (size_t)n);
    for (int64_t q = 0; q < n; ++q) {
      ld a = 2.0L * ORC_PI * (ld)q / (ld)n;
      wr[q] = cosl(a);
      wi[q] = sinl(a);
    }
    for (int c = 0; c < ncomp; ++c) {
      const ld* g = leg + 2 * ((c * nring + ir) * nm);
      ld* out = map + c * map_cstride + ringstart[ir];
      for (int64_t j = 0; j < n; ++j) out[j] = 0;
      for (int64_t m = 0; m < nm; ++m) {
        ld ang = (ld)m * (ld)phi0[ir];
        ld pr = cosl(ang), pi_ = sinl(ang);
        ld tr = g[2 * m] * pr - g[2 * m + 1] * pi_;
        ld ti = g[2 * m] * pi_ + g[2 * m + 1] * pr;
        ld cm = (m == 0) ? 1.0L : 2.0L;
        tr *= cm;
        ti *= cm;
        int64_t step = m % n, idx = 0;
        for (int64_t j = 0; j < n; ++j) {
          out[j] += tr * wr[idx] - ti * wi[idx];
          idx += step;
          if (idx >= n) idx -= n;
        }
      }
    }
    free(wr);
    free(wi);
  }
}

/* map2leg (Ducc0.Sht.map2leg! at src/sht.jl:169): exact transpose of leg2map
 * without the c_m:  leg[m,ir] = sum_j map[j] exp(-i m (phi0 + 2 pi j/n)),  m = 0..nm-1.
 *   map in: double; leg out: long double complex. */
void orc_map2leg(int ncomp, int64_t nm, int64_t nring, const int64_t* nphi, const double* phi0,
                 const int64_t* ringstart, const double* map, int64_t map_cstride, ld* leg,
                 const unsigned char* ring_sel) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t ir = 0; ir < nring; ++ir) {
    if (ring_sel && !ring_sel[ir]) continue;
    int64_t n = nphi[ir];
    ld* wr = (ld*)malloc(sizeof(ld) * (size_t)n);
    ld* wi = (ld*)malloc(sizeof(ld) * (size_t)n);
    for (int64_t q = 0; q < n; ++q) {
      ld a = 2.0L * ORC_PI * (ld)q / (ld)n;
      wr[q] = cosl(a);
      wi[q] = sinl(a);
    }
    for (int c = 0; c < ncomp; ++c) {
      ld* g = leg + 2 * ((c * nring + ir) * nm);
      const double* in = map + c * map_cstride + ringstart[ir];
      for (int64_t m = 0; m < nm; ++m) {
        int64_t step = m % n, idx = 0;
        ld sr = 0, si = 0;
        for (int64_t j = 0; j < n; ++j) {
          sr += (ld)in[j] * wr[idx];
          si -= (ld)in[j] * wi[idx];
          idx += step;
          if (idx >= n) idx -= n;
        }
        ld ang = (ld)m * (ld)phi0[ir];
        ld pr = cosl(ang), pi_ = -sinl(ang);
        g[2 * m] = sr * pr - si * pi_;
        g[2 * m + 1] = sr * pi_ + si * pr;
      }
    }
    free(wr);
    free(wi);
  }
}

/* ------------------------------------------------------------------------------------------
 * Brute force in __float128: explicit Goldberg sY_lm, full sum over -l<=m<=l.
 *   alm: double complex in Healpix.jl Alm order (m-major: index m(2lmax+1-m)/2 + l), m>=0,
 *        a_{l,-m} = (-1)^m conj(a_lm); for spin 2, alm[0]=E, alm[1]=B (each nalm complex).
 *   spin 0: out[0][p] = Re sum a_lm Y_lm      (Im(a_l0) ignored: only Re of m=0 terms)
 *   spin 2: P = sum 2a_lm 2Y_lm, M = sum -2a_lm -2Y_lm, 2a=-(E+iB), -2a=-(E-iB),
 *           Q = Re (P+M)/2, U = Re (P-M)/(2i)
 * ---------------------------------------------------------------------------------------- */
static q128 q_fact(int n) {
  q128 r = 1;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}
static q128 q_binom(int n, int k) {
  if (k < 0 || k > n) return 0;
  return q_fact(n) / (q_fact(k) * q_fact(n - k));
}
static q128 q_powi(q128 b, int n) {
  q128 r = 1;
  for (int i = 0; i < n; ++i) r *= b;
  return r;
}
/* sY_lm(theta, 0), any integer m in [-l, l], l >= |s| */
static q128 q_sylm(int s, int l, int m, q128 th) {
  if (l < abs(s) || abs(m) > l) return 0;
  q128 ch = cosq(th / 2), sh = sinq(th / 2);
  q128 pref = sqrtq(q_fact(l + m) * q_fact(l - m) * (2 * l + 1) /
                    (4 * M_PIq * q_fact(l + s) * q_fact(l - s)));
  if (m & 1) pref = -pref;
  q128 tot = 0;
  for (int r = 0; r <= l - s; ++r) {
    int k = r + s - m;
    if (k < 0 || k > l + s) continue;
    int pc = 2 * r + s - m; /* power of cos(th/2) */
    int ps = 2 * l - pc;    /* power of sin(th/2) */
    if (pc < 0 || ps < 0) continue;
    q128 t = q_binom(l - s, r) * q_binom(l + s, k) * q_powi(ch, pc) * q_powi(sh, ps);
    if ((l - r - s) & 1) t = -t;
    tot += t;
  }
  return pref * tot;
}

void orc_brute_synth(int spin, int lmax, const double* alm, int64_t npts, const double* theta,
                     const double* phi, double* out) {
  int64_t nalm = (int64_t)(lmax + 1) * (lmax + 2) / 2;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < npts; ++p) {
    q128 th = theta[p], ph = phi[p];
    if (spin == 0) {
      q128 acc = 0;
      for (int m = 0; m <= lmax; ++m) {
        q128 cr = cosq(m * ph), ci = sinq(m * ph);
        for (int l = m; l <= lmax; ++l) {
          int64_t idx = (int64_t)m * (2 * lmax + 1 - m) / 2 + l;
          q128 y = q_sylm(0, l, m, th);
          q128 ar = alm[2 * idx], ai = alm[2 * idx + 1];
          if (m == 0)
            acc += ar * y;
          else
            acc += 2 * (ar * cr - ai * ci) * y;
        }
      }
      out[p] = (double)acc;
    } else {
      const double* E = alm;
      const double* B = alm + 2 * nalm;
      q128 Pr = 0, Pi = 0, Mr = 0, Mi = 0;
      for (int l = 2; l <= lmax; ++l) {
        for (int m = -l; m <= l; ++m) {
          int am = abs(m);
          int64_t idx = (int64_t)am * (2 * lmax + 1 - am) / 2 + l;
          q128 er = E[2 * idx], ei = E[2 * idx + 1], br = B[2 * idx], bi = B[2 * idx + 1];
          if (m == 0) { /* real fields: a_l0 real */
            ei = 0;
            bi = 0;
          }
          if (m < 0) { /* a_{l,-m} = (-1)^m conj(a_lm) */
            ei = -ei;
            bi = -bi;
            if (am & 1) {
              er = -er;
              ei = -ei;
              br = -br;
              bi = -bi;
            }
          }
          /* 2a = -(E + iB) = (-(er - bi), -(ei + br)) ; -2a = -(E - iB) = (-(er + bi), -(ei - br)) */
          q128 a2r = -(er - bi), a2i = -(ei + br);
          q128 am2r = -(er + bi), am2i = -(ei - br);
          q128 cr = cosq(m * ph), ci = sinq(m * ph);
          q128 yp = q_sylm(+2, l, m, th), ym = q_sylm(-2, l, m, th);
          Pr += (a2r * cr - a2i * ci) * yp;
          Pi += (a2r * ci + a2i * cr) * yp;
          Mr += (am2r * cr - am2i * ci) * ym;
          Mi += (am2r * ci + am2i * cr) * ym;
        }
      }
      /* Q = (P+M)/2 ; U = (P-M)/(2i) = -i (P-M)/2 -> Re U = Im(P-M)/2 */
      out[p] = (double)((Pr + Mr) / 2);
      out[npts + p] = (double)((Pi - Mi) / 2);
    }
  }
}

/* expose the brute-force harmonic itself for KATs (value at phi=0) */
double orc_brute_sylm(int s, int l, int m, double theta) { return (double)q_sylm(s, l, m, theta); }

/* ------------------------------------------------------------------------------------------
 * The same two recurrences in __float128 (113-bit mantissa), for checking the long-double rows above at large l where
 * no closed form is affordable (tests/test_oracle.py).  Plain quad arithmetic, no extended exponent: returns -1 when
 * the start value would leave the quad range (large m far from the equator), 0 otherwise.  out as double pairs
 * (hi, lo) so that the caller gets ~32 digits through a double interface.
 * ---------------------------------------------------------------------------------------- */
int orc_lambda_quad(int s, int64_t lmax, int64_t m, double theta, double* out_hi, double* out_lo) {
  q128 th = theta, x = cosq(th);
  int64_t l0 = s == 0 ? m : (m > 2 ? m : 2);
  for (int64_t l = 0; l <= lmax; ++l) out_hi[l] = out_lo[l] = 0.0;
  if (l0 > lmax) return 0;
  if (s != 0 && l0 > 800) return -1; /* the factorials of the explicit sum leave the quad range */
  q128 cur;
  if (s == 0) {
    q128 p = (q128)(2 * m + 1) / (4 * M_PIq);
    for (int64_t k = 1; k <= m; ++k) p *= (q128)(2 * k - 1) / (q128)(2 * k);
    cur = sqrtq(p) * powq(sinq(th), (q128)m);
    if (m & 1) cur = -cur;
  } else {
    cur = q_sylm(s, (int)l0, (int)m, th); /* explicit Goldberg sum at the first l */
  }
  if (!(fabsq(cur) > 0x1p-16000Q)) return cur == 0 && (theta == 0.0) ? 0 : -1;
  q128 prev = 0, m2 = (q128)m * m, s2 = (q128)s * s;
  for (int64_t l = l0; l <= lmax; ++l) {
    out_hi[l] = (double)cur;
    out_lo[l] = (double)(cur - (q128)out_hi[l]);
    q128 fl = l, fl1 = l + 1, nxt;
    if (s == 0) {
      q128 el = (l == l0) ? 0 : sqrtq((fl * fl - m2) / (4 * fl * fl - 1));
      q128 el1 = sqrtq((fl1 * fl1 - m2) / (4 * fl1 * fl1 - 1));
      nxt = (x * cur - el * prev) / el1;
    } else {
      q128 Cl = (l == l0) ? 0 : sqrtq((fl * fl - m2) * (fl * fl - s2) / (fl * fl * (4 * fl * fl - 1)));
      q128 Cl1 = sqrtq((fl1 * fl1 - m2) * (fl1 * fl1 - s2) / (fl1 * fl1 * (4 * fl1 * fl1 - 1)));
      nxt = ((x + (q128)s * (q128)m / (fl * fl1)) * cur - Cl * prev) / Cl1;
    }
    prev = cur;
    cur = nxt;
  }
  return 0;
}
