// k_legendre.cu -- FP64 Legendre stage of the spherical-harmonic analysis
// (and synthesis) on HEALPix ring pairs, spin 0 and spin 2, batched over maps.
//
// Replaces the libsharp/ducc Legendre loops behind hp.map2alm / hp.alm2map
// (heracles/healpy.py:183-189).  FP64 FMA bound.
//
// Analysis kernel design (one CTA = one m and one group of 256 ring pairs,
// 8 warps, one warp = 32 ring pairs, one lane = one ring pair):
//   phase A  every lane advances its own lambda_lm(theta) three-term recursion
//            over a chunk of LC consecutive l (scaled arithmetic while the
//            value is below 2^-200) and writes the values into a per-warp
//            shared-memory tile  Lam[parity][ring][l].
//   phase B  the same warp re-reads that tile as the A operand of a small
//            register-blocked FP64 "GEMM"  out[l][col] += Lam[l][ring] * F[ring][col]
//            where F (the ring Fourier coefficients of all maps of the batch
//            for this m, north+south and north-south combinations) was staged
//            in shared memory once per CTA.  Lane tile = 8 l x 5 columns.
//   flush    the 8 warps' partial tiles are summed through shared memory and
//            added to alm with one RED.ADD.F64 per output (x fl[l] fused).
// The recursion cost (about 4 flops per (l, ring)) is shared by all maps of the
// batch; the accumulate cost is 4 flops per (l, ring, map) for spin 0 and
// 16 per spin-2 field.
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

__device__ __forceinline__ void lam_advance(LamState &s, double ax, double g) {
  double nw = fma(ax, s.cur, -(g * s.prev));
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

struct LegArgs {
  int lmax, nm, ncomp;      // ncomp: components present in `phase` rows (<= capacity of the template)
  const int *mlist;         // nullptr: m = index
  const double *phase;      // [(mi * nrp_local + rpl) * ncomp + c] * 4
  i64 nrp_local, rp_lo;
  const double *cth, *sth, *ch, *sh;  // indexed by global ring pair
  const double *coef;       // recursion coefficients, see hcu_build_coef
  const double *cmtab;
  const double *fl;         // nullptr or [lmax+1]
  double *alm;              // complex rows, stride alm_stride (complex elements)
  i64 alm_stride;
  double *work;             // [2] counters
};

template <int SPIN, int CH>
struct Cfg {
  static constexpr int NJ = SPIN == 0 ? 1 : 2;
  static constexpr int LP = SPIN == 0 ? 32 : 16;  // l per parity per chunk
  static constexpr int LC = 2 * LP;
  static constexpr int LL = LP / 4;               // l per lane
  static constexpr int NB = SPIN == 0 ? 2 * CH : 2 * CH;  // components per batch (maps, or Q/U rows)
  static constexpr int C = 4 * CH;                // output columns (per parity)
  static constexpr int ROW = LP + 2;
  static constexpr int LAM_P = 32 * ROW + 8;
  static constexpr int LAM_J = 2 * LAM_P;
  static constexpr int LAM_WARP = NJ * LAM_J;
  static constexpr int F_ROW = 4 * NB;            // doubles per ring: NB x (re+, im+, re-, im-)
  static constexpr int F_WARP = 32 * F_ROW;
  static constexpr int WARP_SMEM = LAM_WARP + F_WARP;  // doubles
  static constexpr int NOUT = 2 * LP * C;
  static constexpr size_t SMEM_BYTES = (size_t)8 * WARP_SMEM * 8 + 64;
};

template <int SPIN, int CH>
__global__ void __launch_bounds__(256, 1) legendre_analysis_kernel(LegArgs a) {
  using K = Cfg<SPIN, CH>;
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *lam_w = smem_d + warp * K::WARP_SMEM;
  double *f_w = lam_w + K::LAM_WARP;
  int *flags = reinterpret_cast<int *>(smem_d + 8 * K::WARP_SMEM);

  const int ngroups = (int)((a.nrp_local + 255) / 256);
  const int g = blockIdx.x % ngroups;
  const int mi = blockIdx.x / ngroups;
  const int m = a.mlist ? a.mlist[mi] : mi;
  const int lmax = a.lmax;
  const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
  if (l0 > lmax) return;
  const int pb = (l0 + m) & 1;

  const i64 rpl = (i64)g * 256 + warp * 32 + lane;
  const bool valid = rpl < a.nrp_local;
  double x = 0, sth = 1, chh = 1, shh = 1;
  if (valid) {
    const i64 rp = a.rp_lo + rpl;
    x = a.cth[rp];
    sth = a.sth[rp];
    chh = a.ch[rp];
    shh = a.sh[rp];
  }
  const bool alive = valid && !ring_is_dead(lmax, m, SPIN, x, sth);
  const bool warp_alive = __any_sync(0xffffffffu, alive);
  // block-uniform early exit when no ring of this CTA can contribute
  if (__syncthreads_or(alive ? 1 : 0) == 0) return;

  // ---- stage F (ring Fourier coefficients of this m) into shared memory ----
  {
    const double *src = a.phase + ((i64)mi * a.nrp_local + (i64)g * 256 + warp * 32) * a.ncomp * 4;
    const int rows = (int)min((i64)32, a.nrp_local - ((i64)g * 256 + warp * 32));
    const int w = a.ncomp * 4;
    for (int idx = lane; idx < 32 * K::F_ROW; idx += 32) {
      int r = idx / K::F_ROW, cidx = idx - r * K::F_ROW;
      double v = 0.0;
      if (r < rows && cidx < w) v = src[(i64)r * w + cidx];
      f_w[idx] = v;
    }
  }

  LamState sp, sm;
  sp.prev = sp.cur = 0; sp.e = 0;
  sm.prev = sm.cur = 0; sm.e = 0;
  if (alive) lam_start<SPIN>(m, a.cmtab, sth, chh, shh, sp, sm);
  __syncwarp();

  // lane roles for phase B
  const int pB = lane >> 4, gB = (lane >> 2) & 3, hB = lane & 3;
  // column offsets into an F row
  int foff[CH], foff2[CH];
  double sgnP = 1.0, sgnM = 1.0;
  if (SPIN == 0) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      int col = hB * CH + i;
      foff[i] = (col >> 1) * 4 + 2 * pB + (col & 1);
      foff2[i] = 0;
    }
  } else {
    // per field 8 doubles: Q(re+, im+, re-, im-), U(re+, im+, re-, im-)
    // h = 0: E_re = -F+ Q^s_re + F- U^-s_im     h = 1: E_im = -F+ Q^s_im - F- U^-s_re
    // h = 2: B_re = -F+ U^s_re - F- Q^-s_im     h = 3: B_im = -F+ U^s_im + F- Q^-s_re
    // parity 0: s = + (Q^s = Q+, U^-s = U-); parity 1: s = -
    const int oP[4][2] = {{0, 2}, {1, 3}, {4, 6}, {5, 7}};
    const int oM[4][2] = {{7, 5}, {6, 4}, {3, 1}, {2, 0}};
    const double sM[4] = {1.0, -1.0, -1.0, 1.0};
    sgnP = -1.0;
    sgnM = sM[hB];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      foff[i] = i * 8 + oP[hB][pB];
      foff2[i] = i * 8 + oM[hB][pB];
    }
  }

  const int nchunk = (lmax - l0 + K::LC) / K::LC;
  const i64 cbase = alm_index(lmax, 0, m);  // coefficient index of (l, m) is cbase + l
  double n_rec = 0, n_acc = 0;

  for (int chk = 0; chk < nchunk; ++chk) {
    const int lstart = l0 + chk * K::LC;
    bool live = false;
    if (warp_alive) {
      // ------------------------- phase A ---------------------------------
#pragma unroll 1
      for (int s = 0; s < K::LC; s += 4) {
        double v[4], v2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int l = lstart + s + u;
          const bool inr = l <= lmax;
          if (SPIN == 0) {
            v[u] = (inr && sp.e == 0) ? sp.cur : 0.0;
            if (l < lmax) {
              const double2 cf = __ldg(reinterpret_cast<const double2 *>(a.coef) + cbase + l);
              lam_advance(sp, cf.x * x, cf.y);
            }
          } else {
            double lp = (inr && sp.e == 0) ? sp.cur : 0.0;
            double lm = (inr && sm.e == 0) ? sm.cur : 0.0;
            v[u] = 0.5 * (lp + lm);
            v2[u] = 0.5 * (lp - lm);
            if (l < lmax) {
              const double4 cf = ldg_d4(reinterpret_cast<const double4 *>(a.coef) + cbase + l);
              lam_advance(sp, fma(cf.x, x, cf.y), cf.z);
              lam_advance(sm, fma(cf.x, x, -cf.y), cf.z);
            }
          }
        }
        // steps s, s+2 have parity pb; s+1, s+3 parity 1-pb; index within parity = s/2 (+1)
        double *dst0 = lam_w + pb * K::LAM_P + lane * K::ROW + (s >> 1);
        double *dst1 = lam_w + (1 - pb) * K::LAM_P + lane * K::ROW + (s >> 1);
        *reinterpret_cast<double2 *>(dst0) = make_double2(v[0], v[2]);
        *reinterpret_cast<double2 *>(dst1) = make_double2(v[1], v[3]);
        if (SPIN != 0) {
          *reinterpret_cast<double2 *>(dst0 + K::LAM_J) = make_double2(v2[0], v2[2]);
          *reinterpret_cast<double2 *>(dst1 + K::LAM_J) = make_double2(v2[1], v2[3]);
        }
      }
      bool lane_live = alive && (sp.e == 0 || (SPIN != 0 && sm.e == 0));
      live = __any_sync(0xffffffffu, lane_live);
      n_rec += 1;
    }
    __syncwarp();

    double acc[K::LL][CH], acc2[SPIN == 0 ? 1 : K::LL][SPIN == 0 ? 1 : CH];
    if (live) {
      // ------------------------- phase B ---------------------------------
#pragma unroll
      for (int i = 0; i < K::LL; ++i)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          acc[i][j] = 0.0;
          if (SPIN != 0) acc2[i][j] = 0.0;
        }
      const double *lamp = lam_w + pB * K::LAM_P + 2 * gB;
#pragma unroll 2
      for (int k = 0; k < 32; ++k) {
        double la[K::LL], lb[SPIN == 0 ? 1 : K::LL];
#pragma unroll
        for (int i = 0; i < K::LL / 2; ++i) {
          double2 t = *reinterpret_cast<const double2 *>(lamp + k * K::ROW + 8 * i);
          la[2 * i] = t.x;
          la[2 * i + 1] = t.y;
          if (SPIN != 0) {
            double2 t2 = *reinterpret_cast<const double2 *>(lamp + K::LAM_J + k * K::ROW + 8 * i);
            lb[2 * i] = t2.x;
            lb[2 * i + 1] = t2.y;
          }
        }
        const double *fr = f_w + k * K::F_ROW;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const double fv = fr[foff[j]];
#pragma unroll
          for (int i = 0; i < K::LL; ++i) acc[i][j] = fma(la[i], fv, acc[i][j]);
          if (SPIN != 0) {
            const double fv2 = fr[foff2[j]];
#pragma unroll
            for (int i = 0; i < K::LL; ++i) acc2[i][j] = fma(lb[i], fv2, acc2[i][j]);
          }
        }
      }
      n_acc += 1;
    }
    __syncwarp();
    if (live) {
      // write this warp's partial tile over its (now consumed) lambda tile:
      // out_w[(p * LP + lidx) * C + col]
#pragma unroll
      for (int i = 0; i < K::LL; ++i) {
        const int lidx = 8 * (i >> 1) + 2 * gB + (i & 1);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          double val = (SPIN == 0) ? acc[i][j] : (sgnP * acc[i][j] + sgnM * acc2[i][j]);
          lam_w[(pB * K::LP + lidx) * K::C + hB * CH + j] = val;
        }
      }
    }
    if (lane == 0) flags[warp] = live ? 1 : 0;
    __syncthreads();
    // ------------------------- flush -------------------------------------
    {
      int fl_any = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) fl_any |= flags[w];
      if (fl_any) {
        for (int o = threadIdx.x; o < K::NOUT; o += 256) {
          // o = col * LC + s  (consecutive threads -> consecutive l)
          const int col = o / K::LC, s = o - col * K::LC;
          const int l = lstart + s;
          const int p = (s + pb) & 1, lidx = s >> 1;
          int row, ri;
          bool colvalid;
          if (SPIN == 0) {
            row = col >> 1;
            ri = col & 1;
            colvalid = row < a.ncomp;
          } else {
            const int h = col / CH, f = col - h * CH;
            row = 2 * f + (h >> 1);
            ri = h & 1;
            colvalid = (2 * f) < a.ncomp;
          }
          if (l <= lmax && colvalid) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
              if (flags[w]) sum += smem_d[w * K::WARP_SMEM + (p * K::LP + lidx) * K::C + col];
            if (a.fl) sum *= a.fl[l];
            atomicAdd(a.alm + 2 * ((i64)row * a.alm_stride + cbase + l) + ri, sum);
          }
        }
      }
    }
    __syncthreads();
  }
  if (lane == 0 && a.work && (n_rec > 0)) {
    atomicAdd(a.work, n_rec * 32.0 * K::LC);
    atomicAdd(a.work + 1, n_acc * 32.0 * K::LC);
  }
}

// ---------------------------------------------------------------------------
// synthesis: one lane = one ring pair, loops over l, alm broadcast from smem.
// phase out: [(m * nrp + rp) * ncomp + c] * 4 = (reN, imN, reS, imS)
// ---------------------------------------------------------------------------
template <int SPIN, int NB>
__global__ void __launch_bounds__(128) legendre_synthesis_kernel(LegArgs a, double *phase_out) {
  // smem: alm chunk [LCH][NB][2]
  constexpr int LCH = 64;
  __shared__ double s_alm[LCH * NB * 2];
  const int ngroups = (int)((a.nrp_local + 127) / 128);
  const int g = blockIdx.x % ngroups;
  const int m = blockIdx.x / ngroups;
  const int lmax = a.lmax;
  const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
  const i64 rp = (i64)g * 128 + threadIdx.x;
  const bool valid = rp < a.nrp_local;
  double x = 0, sth = 1, chh = 1, shh = 1;
  if (valid) {
    x = a.cth[rp];
    sth = a.sth[rp];
    chh = a.ch[rp];
    shh = a.sh[rp];
  }
  // north and south accumulators per component (complex)
  double aN[NB][2], aS[NB][2];
#pragma unroll
  for (int c = 0; c < NB; ++c) aN[c][0] = aN[c][1] = aS[c][0] = aS[c][1] = 0.0;
  if (l0 <= lmax) {
    const bool alive = valid && !ring_is_dead(lmax, m, SPIN, x, sth);
    LamState sp, sm;
    sp.prev = sp.cur = 0; sp.e = 0;
    sm.prev = sm.cur = 0; sm.e = 0;
    if (alive) lam_start<SPIN>(m, a.cmtab, sth, chh, shh, sp, sm);
    const i64 cbase = alm_index(lmax, 0, m);
    for (int lc = l0; lc <= lmax; lc += LCH) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < LCH * NB * 2; idx += 128) {
        int li = idx / (NB * 2), r = idx - li * NB * 2;
        int c = r >> 1, ri = r & 1;
        int l = lc + li;
        double v = 0.0;
        if (l <= lmax && c < a.ncomp)
          v = a.alm[2 * ((i64)c * a.alm_stride + cbase + l) + ri];
        s_alm[idx] = v;
      }
      __syncthreads();
      if (!alive) continue;
      const int lend = min(LCH, lmax - lc + 1);
      for (int li = 0; li < lend; ++li) {
        const int l = lc + li;
        const double sg = ((l + m) & 1) ? -1.0 : 1.0;
        const double *al = s_alm + li * NB * 2;
        if (SPIN == 0) {
          if (sp.e == 0) {
            const double lam = sp.cur, lams = sg * lam;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
              aN[c][0] = fma(lam, al[2 * c], aN[c][0]);
              aN[c][1] = fma(lam, al[2 * c + 1], aN[c][1]);
              aS[c][0] = fma(lams, al[2 * c], aS[c][0]);
              aS[c][1] = fma(lams, al[2 * c + 1], aS[c][1]);
            }
          }
          if (l < lmax) {
            const double2 cf = __ldg(reinterpret_cast<const double2 *>(a.coef) + cbase + l);
            lam_advance(sp, cf.x * x, cf.y);
          }
        } else {
          const double lp = (sp.e == 0) ? sp.cur : 0.0;
          const double lm = (sm.e == 0) ? sm.cur : 0.0;
          if (sp.e == 0 || sm.e == 0) {
            // accumulate P = sum 2a lam+, M = sum -2a lam-  (north); south swaps lam+-
#pragma unroll
            for (int c = 0; c < NB; c += 2) {
              const double Er = al[2 * c], Ei = al[2 * c + 1];
              const double Br = al[2 * c + 2], Bi = al[2 * c + 3];
              const double a2r = -(Er - Bi), a2i = -(Ei + Br);
              const double m2r = -(Er + Bi), m2i = -(Ei - Br);
              aN[c][0] = fma(lp, a2r, aN[c][0]);
              aN[c][1] = fma(lp, a2i, aN[c][1]);
              aN[c + 1][0] = fma(lm, m2r, aN[c + 1][0]);
              aN[c + 1][1] = fma(lm, m2i, aN[c + 1][1]);
              aS[c][0] = fma(sg * lm, a2r, aS[c][0]);
              aS[c][1] = fma(sg * lm, a2i, aS[c][1]);
              aS[c + 1][0] = fma(sg * lp, m2r, aS[c + 1][0]);
              aS[c + 1][1] = fma(sg * lp, m2i, aS[c + 1][1]);
            }
          }
          if (l < lmax) {
            const double4 cf = ldg_d4(reinterpret_cast<const double4 *>(a.coef) + cbase + l);
            lam_advance(sp, fma(cf.x, x, cf.y), cf.z);
            lam_advance(sm, fma(cf.x, x, -cf.y), cf.z);
          }
        }
      }
    }
  }
  if (!valid) return;
  for (int c = 0; c < NB && c < a.ncomp; ++c) {
    double4 o;
    if (SPIN == 0) {
      o = make_double4(aN[c][0], aN[c][1], aS[c][0], aS[c][1]);
    } else {
      // c even: Q = (P + M)/2 ; c odd: U = (P - M)/(2i)
      const int cq = c & ~1;
      const double PrN = aN[cq][0], PiN = aN[cq][1], MrN = aN[cq + 1][0], MiN = aN[cq + 1][1];
      const double PrS = aS[cq][0], PiS = aS[cq][1], MrS = aS[cq + 1][0], MiS = aS[cq + 1][1];
      if ((c & 1) == 0)
        o = make_double4(0.5 * (PrN + MrN), 0.5 * (PiN + MiN), 0.5 * (PrS + MrS), 0.5 * (PiS + MiS));
      else
        o = make_double4(0.5 * (PiN - MiN), -0.5 * (PrN - MrN), 0.5 * (PiS - MiS), -0.5 * (PrS - MrS));
    }
    *reinterpret_cast<double4 *>(phase_out + (((i64)m * a.nrp_local + rp) * a.ncomp + c) * 4) = o;
  }
}

// recursion coefficient tables: step l -> l+1 for l >= l0
//   spin 0: (alpha_l, gamma_l),  L_{l+1} = alpha x L_l - gamma L_{l-1}
//   spin 2: (alpha_l, alpha_l beta_l, gamma_l, 0),  L^{+-}_{l+1} = (alpha x +- alpha beta) L_l - gamma L_{l-1}
__global__ void coef_kernel(int lmax, int spin, double *tab) {
  const int m = blockIdx.x;
  const int s = spin;
  const int l0 = m > s ? m : s;
  const i64 base = (i64)m * (2 * lmax + 1 - m) / 2;
  for (int l = m + threadIdx.x; l <= lmax; l += blockDim.x) {
    double al = 0, ab = 0, ga = 0;
    if (l >= l0 && l < lmax) {
      const double dl = l, l1 = dl + 1.0, dm = m, ds = s;
      const double den = sqrt((l1 * l1 - dm * dm) * (l1 * l1 - ds * ds));
      al = sqrt((2 * dl + 3) / (2 * dl + 1)) * l1 * (2 * dl + 1) / den;
      const double be = (l > 0) ? (ds * dm) / (dl * l1) : 0.0;
      ab = al * be;
      ga = (l > l0) ? sqrt((2 * dl + 3) / (2 * dl - 1)) * l1 / dl *
                          sqrt((dl * dl - dm * dm) * (dl * dl - ds * ds)) / den
                    : 0.0;
    }
    if (s == 0) {
      reinterpret_cast<double2 *>(tab)[base + l] = make_double2(al, ga);
    } else {
      reinterpret_cast<double4 *>(tab)[base + l] = make_double4(al, ab, ga, 0.0);
    }
  }
}

template <int SPIN, int CH>
int launch_analysis(hcu_ctx *ctx, const LegArgs &a) {
  using K = Cfg<SPIN, CH>;
  const int ngroups = (int)((a.nrp_local + 255) / 256);
  const i64 nblocks = (i64)ngroups * a.nm;
  if (nblocks <= 0) return HCU_OK;
  HCU_CUDA(cudaFuncSetAttribute(legendre_analysis_kernel<SPIN, CH>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)K::SMEM_BYTES));
  legendre_analysis_kernel<SPIN, CH><<<(unsigned)nblocks, 256, K::SMEM_BYTES, ctx->stream>>>(a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

template <int SPIN, int NB>
int launch_synthesis(hcu_ctx *ctx, const LegArgs &a, double *phase_out) {
  const int ngroups = (int)((a.nrp_local + 127) / 128);
  const i64 nblocks = (i64)ngroups * (a.lmax + 1);
  legendre_synthesis_kernel<SPIN, NB><<<(unsigned)nblocks, 128, 0, ctx->stream>>>(a, phase_out);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

}  // namespace

int hcu_build_coef(hcu_ctx *ctx, hcu_coef *c) {
  const i64 nalm = (i64)(c->lmax + 1) * (c->lmax + 2) / 2;
  const size_t per = (c->spin == 0) ? 2 : 4;
  HCU_CUDA(cudaMalloc(&c->tab, sizeof(double) * per * nalm));
  coef_kernel<<<c->lmax + 1, 128, 0, ctx->stream>>>(c->lmax, c->spin, c->tab);
  HCU_LAUNCH_CHECK(ctx);
  // start-value normalisation in long double on the host
  std::vector<double> cm(2 * (size_t)(c->lmax + 1));
  long double v = sqrtl(1.0L / (4.0L * 3.141592653589793238462643383279502884L));
  for (int m = 0; m <= c->lmax; ++m) {
    if (m > 0) v *= sqrtl((2.0L * m + 1.0L) / (2.0L * m));
    cm[2 * m] = (double)v;
    cm[2 * m + 1] = (m >= 2) ? (double)(v * sqrtl((long double)m * (m - 1) /
                                                  ((long double)(m + 1) * (m + 2))))
                             : 0.0;
  }
  HCU_CUDA(cudaMalloc(&c->cm, sizeof(double) * cm.size()));
  HCU_CUDA(cudaMemcpyAsync(c->cm, cm.data(), sizeof(double) * cm.size(),
                           cudaMemcpyHostToDevice, ctx->stream));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));  // cm is a stack-lifetime host buffer
  return HCU_OK;
}

// maximum components per analysis launch
static int batch_capacity(int spin) { return spin == 0 ? 10 : 10; }

int hcu_legendre_analysis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                          int spin, int ncomp, const double *phase,
                          const int32_t *mlist_dev, int nm, i64 rp_lo, i64 rp_hi,
                          const double *fl_dev, double *alm, i64 alm_stride) {
  HCU_ARG(ncomp >= 1 && ncomp <= batch_capacity(spin), "legendre batch size");
  LegArgs a;
  a.lmax = lmax;
  a.nm = nm;
  a.ncomp = ncomp;
  a.mlist = mlist_dev;
  a.phase = phase;
  a.nrp_local = rp_hi - rp_lo;
  a.rp_lo = rp_lo;
  a.cth = g->cth;
  a.sth = g->sth;
  a.ch = g->ch;
  a.sh = g->sh;
  a.coef = c->tab;
  a.cmtab = c->cm;
  a.fl = fl_dev;
  a.alm = alm;
  a.alm_stride = alm_stride;
  a.work = ctx->work_counters;
  if (spin == 0) {
    const int chn = (ncomp + 1) / 2;  // 2 maps per column-group unit
    switch (chn) {
      case 1: return launch_analysis<0, 1>(ctx, a);
      case 2: return launch_analysis<0, 2>(ctx, a);
      case 3: return launch_analysis<0, 3>(ctx, a);
      case 4: return launch_analysis<0, 4>(ctx, a);
      default: return launch_analysis<0, 5>(ctx, a);
    }
  } else {
    const int nf = ncomp / 2;
    switch (nf) {
      case 1: return launch_analysis<2, 1>(ctx, a);
      case 2: return launch_analysis<2, 2>(ctx, a);
      case 3: return launch_analysis<2, 3>(ctx, a);
      case 4: return launch_analysis<2, 4>(ctx, a);
      default: return launch_analysis<2, 5>(ctx, a);
    }
  }
}

int hcu_legendre_synthesis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                           int spin, int ncomp, const double *alm,
                           i64 alm_stride, double *phase) {
  LegArgs a;
  a.lmax = lmax;
  a.nm = lmax + 1;
  a.ncomp = ncomp;
  a.mlist = nullptr;
  a.phase = nullptr;
  a.nrp_local = g->nrp;
  a.rp_lo = 0;
  a.cth = g->cth;
  a.sth = g->sth;
  a.ch = g->ch;
  a.sh = g->sh;
  a.coef = c->tab;
  a.cmtab = c->cm;
  a.fl = nullptr;
  a.alm = const_cast<double *>(alm);
  a.alm_stride = alm_stride;
  a.work = nullptr;
  if (spin == 0) {
    if (ncomp <= 1) return launch_synthesis<0, 1>(ctx, a, phase);
    if (ncomp <= 2) return launch_synthesis<0, 2>(ctx, a, phase);
    if (ncomp <= 4) return launch_synthesis<0, 4>(ctx, a, phase);
    if (ncomp <= 6) return launch_synthesis<0, 6>(ctx, a, phase);
    if (ncomp <= 8) return launch_synthesis<0, 8>(ctx, a, phase);
    HCU_ARG(ncomp <= 10, "synthesis batch size");
    return launch_synthesis<0, 10>(ctx, a, phase);
  } else {
    if (ncomp <= 2) return launch_synthesis<2, 2>(ctx, a, phase);
    if (ncomp <= 4) return launch_synthesis<2, 4>(ctx, a, phase);
    if (ncomp <= 6) return launch_synthesis<2, 6>(ctx, a, phase);
    if (ncomp <= 8) return launch_synthesis<2, 8>(ctx, a, phase);
    HCU_ARG(ncomp <= 10, "synthesis batch size");
    return launch_synthesis<2, 10>(ctx, a, phase);
  }
}
