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

// ---- helpers shared by k_legendre.cu and k_legendre2.cu ----
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Lam tile of one (j, parity): [ring 0..31][8 l], 16-byte units XOR-swizzled by the ring
__device__ __forceinline__ int swzf(int ring) { return (((ring >> 1) & 1) << 1) | ((ring >> 2) & 1); }
__device__ __forceinline__ int lam_off(int ring, int idx) {  // idx = l index within the parity, 0..7
  return ring * 8 + 2 * ((idx >> 1) ^ swzf(ring)) + (idx & 1);
}

__device__ __forceinline__ double mask_d(double v, unsigned long long msk) {
  return __longlong_as_double((long long)((unsigned long long)__double_as_longlong(v) & msk));
}
__device__ __forceinline__ double neg_d(double v) {  // sign flip on the integer pipe
  return __longlong_as_double(__double_as_longlong(v) ^ (long long)0x8000000000000000ull);
}

// prefetch loads: `asm volatile` pins them where they are written, a full sub-chunk (or chunk)
// of tensor work ahead of their first use, instead of letting ptxas sink them next to it
__device__ __forceinline__ double ldg_pin(const double *p) {
  double r;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg_pin2(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ void coef_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// split-phase CTA barrier (arrive now, wait a chunk later) for the deferred alm flush
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, int parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

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
  // optional start-state table (hcu_get_start): where every recursion chain first becomes representable and its
  // (prev, cur) there, indexed [(m * st_nrp + ring pair) * NJ + j]; nullptr: every chain starts at l0 (lam_start)
  const int *st_sub;
  const double2 *st_state;
  int st_nrp;
  hcu_ptrs alm;             // one complex128 row per component
  double *phase_out;        // synthesis output, same layout as `phase` with (reN, imN, reS, imS)
  // multi-GPU: block b of the synthesis output is written to blk_out[b] (the peer that owns those ring pairs) instead of
  // phase_out + its offset in the concatenation
  int use_blk_out;
  double *blk_out[HCU_MAX_BLOCKS];
  double *work;             // [2] counters
};

}  // namespace
