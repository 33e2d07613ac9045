// legendre_common.cuh -- device helpers shared by the Legendre analysis and synthesis kernels:
// scaled three-term recursions for lambda_lm (spin 0) and the spin-2 Wigner-d functions.
#pragma once
#include "hcu_common.cuh"

namespace {

constexpr int SCALE_STEP = 400;
constexpr int SCALE_HALF = 200;
#define TWO_P200 1.6069380442589903e60
#define TWO_M400 3.8725919148493183e-121

struct LamState {
  double prev, cur;
  int e;
};

__device__ __forceinline__ void pow_scaled(double x, int n, double *mant, int *ex) {
  double r = 1.0, b;
  int re = 0, be, t;
  b = frexp(x, &be);
  while (n > 0) {
    if (n & 1) {
      r *= b;
      re += be;
      r = frexp(r, &t);
      re += t;
    }
    b *= b;
    be *= 2;
    b = frexp(b, &t);
    be += t;
    n >>= 1;
  }
  r = frexp(r, &t);
  *mant = r;
  *ex = re + t;
}

__device__ __forceinline__ void set_scaled(LamState &s, double mant, int k) {
  s.prev = 0.0;
  if (mant == 0.0) {
    s.cur = 0.0;
    s.e = 0;
    return;
  }
  int e = 0;
  if (k < -SCALE_HALF) {
    int q = (-(k + SCALE_HALF) + SCALE_STEP - 1) / SCALE_STEP;
    e = -q * SCALE_STEP;
  }
  s.cur = ldexp(mant, k - e);
  s.e = e;
}

// one step of the scaled recursion q_{l+1} = ax q_l - q_{l-1} with extended-exponent rescaling
__device__ __forceinline__ void lam_advance(LamState &s, double ax) {
  double nw = fma(ax, s.cur, -s.prev);
  s.prev = s.cur;
  s.cur = nw;
  if (s.e < 0 && fabs(nw) >= TWO_P200) {
    s.cur *= TWO_M400;
    s.prev *= TWO_M400;
    s.e += SCALE_STEP;
  }
}

// starting values; cmtab[2m] = c_m, cmtab[2m+1] = c_m sqrt(m(m-1)/((m+1)(m+2)))
template <int SPIN>
__device__ __forceinline__ void lam_start(int m, const double *cmtab, double sth,
                                          double ch, double sh, LamState &sp,
                                          LamState &sm) {
  const double sign = (m & 1) ? -1.0 : 1.0;
  if (SPIN == 0) {
    double mant;
    int k, t;
    pow_scaled(sth, m, &mant, &k);
    double v = frexp(mant * cmtab[2 * m], &t);
    set_scaled(sp, sign * v, k + t);
    return;
  }
  if (m >= 2) {
    double pc, ps;
    int kc, ks, t;
    const double f = cmtab[2 * m + 1];
    // spin +2: cos^(m-2) sin^(m+2)
    pow_scaled(ch, m - 2, &pc, &kc);
    pow_scaled(sh, m + 2, &ps, &ks);
    double v = frexp(pc * ps * f, &t);
    set_scaled(sp, sign * v, kc + ks + t + m);
    // spin -2: cos^(m+2) sin^(m-2)
    pow_scaled(ch, m + 2, &pc, &kc);
    pow_scaled(sh, m - 2, &ps, &ks);
    v = frexp(pc * ps * f, &t);
    set_scaled(sm, sign * v, kc + ks + t + m);
  } else {
    const double n2 = 0.63078313050504001;  // sqrt(5/(4 pi))
    const double fac = (m == 0) ? 2.4494897427831781 : 2.0;
    double vp = sign * n2 * fac, vm = n2 * fac;
    for (int i = 0; i < 2 - m; ++i) { vp *= ch; vm *= sh; }
    for (int i = 0; i < 2 + m; ++i) { vp *= sh; vm *= ch; }
    sp.prev = 0.0; sp.cur = vp; sp.e = 0;
    sm.prev = 0.0; sm.cur = vm; sm.e = 0;
  }
}

__device__ __forceinline__ double4 ldg_d4(const double4 *p) {
  const double2 lo = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 hi = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ i64 alm_index(int lmax, int l, int m) {
  return (i64)m * (2 * lmax + 1 - m) / 2 + l;
}

// conservative estimate of the largest m that contributes at colatitude theta
__device__ __forceinline__ bool ring_is_dead(int lmax, int m, int spin, double cth, double sth) {
  double ofs = fmax(300.0, 0.03 * lmax);
  double b = -2.0 * spin * fabs(cth);
  double t1 = lmax * sth + ofs;
  double c = (double)spin * spin - t1 * t1;
  double disc = b * b - 4.0 * c;
  double res = (disc <= 0) ? lmax : (-b + sqrt(disc)) * 0.5;
  return (double)m > res;
}

#define HCU_MAX_BLOCKS 16
struct LegArgs {
  int lmax, nm, ncomp;      // ncomp: components present in `phase` rows (<= capacity of the template)
  const int *mlist;         // nullptr: m = index
  // ring pairs come in up to HCU_MAX_BLOCKS consecutive blocks [blk_rp[b], blk_rp[b+1]); phase is the
  // concatenation over the blocks of [(mi * nrp_b + rp - blk_rp[b]) * ncomp + c] * 4 (one block: the
  // plain [nm][nrp][ncomp][4] array).  grp_start[b] = first 256-ring-pair group of block b.
  const double *phase;
  int nblk;
  int grp_start[HCU_MAX_BLOCKS + 1];
  i64 blk_rp[HCU_MAX_BLOCKS + 1];
  const double *cth, *sth, *ch, *sh;  // indexed by global ring pair
  const double *coef;       // recursion coefficients, see hcu_build_coef
  const double *scale;      // s_l of the scaled recursion, lambda_l = s_l q_l
  const double *cmtab;
  const double *fl;         // nullptr or [lmax+1]
  hcu_ptrs alm;             // one complex128 row per component
  double *phase_out;        // synthesis output, same layout as `phase` with (reN, imN, reS, imS)
  double *work;             // [2] counters
};

}  // namespace
