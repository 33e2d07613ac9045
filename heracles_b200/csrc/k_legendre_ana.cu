// k_legendre_ana.cu -- FP64 Legendre stage of the spherical-harmonic ANALYSIS on
// HEALPix ring pairs, spin 0 and spin 2, batched over up to 12 components.
//
// Replaces the libsharp/ducc Legendre loop behind hp.map2alm
// (heracles/healpy.py:183-189).  Bound: FP64 pipe (DFMA and DMMA share it on
// B200: 36-37 TFLOP/s measured for either, tools/dmma_peak.cu).
//
// One CTA = one m and one group of 256 ring pairs; 8 warps; one warp = 32 ring
// pairs; in phase A one lane = one ring pair.
//   phase A  each lane advances its three-term recursion in l over a chunk of
//            LC consecutive l (scaled arithmetic while the value is below
//            2^-200; recursion coefficients of the chunk are staged once per
//            CTA in shared memory) and stores lambda into the warp's
//            shared-memory tile Lam[parity][ring k][l], XOR-swizzled so that
//            both the 128-bit stores and the fragment loads are conflict free.
//   phase B  out[l][col] += sum_k Lam[l][k] F[k][col] on the FP64 tensor path:
//            mma.sync.m8n8k4.f64 with A = Lam^T (8 l x 4 rings), B = F (4 rings
//            x 8 columns), accumulators in registers.  F holds the ring Fourier
//            coefficients of all maps of the batch for this m (north+south for
//            even l+m, north-south for odd), staged in shared memory once per
//            CTA.  One DMMA = 256 FMA for two 8-byte shared loads per lane,
//            which is what keeps the shared-memory pipe off the critical path
//            (the DFMA formulation needed ~1 wavefront per FMA instruction).
//   flush    the 8 warps' partial tiles are summed through shared memory and
//            added to alm with one RED.ADD.F64 per output (x fl[l] fused).
// Work per (l, m, ring pair): recursion ~4 flop (spin 0) / ~12 (spin 2), shared
// by the batch; accumulate 4 flop per spin-0 map, 16 per spin-2 field.
#include "legendre_common.cuh"

#ifndef HCU_ANA_WARPS
#define HCU_ANA_WARPS 8
#endif

namespace {

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// XOR swizzle of the column index inside a row of Lam / F, see header comment
__device__ __forceinline__ int swz(int k) { return ((k & 3) << 2) | (((k >> 2) & 1) << 1); }

template <int SPIN, int NBLK>
struct Cfg {
  static constexpr int NJ = SPIN == 0 ? 1 : 2;    // lambda matrices (spin 2: F+ and F-)
  static constexpr int LP = SPIN == 0 ? 32 : 16;  // l per parity per chunk
  static constexpr int LC = 2 * LP;               // l per chunk
  static constexpr int MB = LP / 8;               // 8-row m-blocks per parity
  static constexpr int C = 8 * NBLK;              // output columns per parity
  static constexpr int NCOMP = C / 2;             // components per batch (maps, or Q/U rows)
  static constexpr int TILE = 32 * LP;            // doubles of one Lam[parity] tile
  static constexpr int LAM_WARP = NJ * 2 * TILE;  // 2048 doubles for both spins
  static constexpr int F_ROW = 2 * C;             // spin 0: [parity][C]; spin 2: 8 per field
  static constexpr int F_WARP = 32 * F_ROW;
  static constexpr int WARP_SMEM = LAM_WARP + F_WARP;
  static constexpr int NOUT = 2 * LP * C;
  static constexpr int OS = LC + 2;               // column stride of the staged out tile
  static constexpr int COEF_W = SPIN == 0 ? 2 : 4;  // doubles per coefficient entry
  static constexpr int NW = HCU_ANA_WARPS;          // warps per CTA, 32 ring pairs each
  static constexpr int NT = 32 * NW;
  static constexpr int COEF_OFF = NW * WARP_SMEM;
  static constexpr int FLAG_OFF = COEF_OFF + LC * COEF_W;
  static constexpr size_t SMEM_BYTES = (size_t)(FLAG_OFF + 8) * 8;
  static_assert(C * OS <= LAM_WARP, "out tile must fit in the lambda tile");
};

template <int SPIN, int NBLK>
__global__ void __launch_bounds__(32 * HCU_ANA_WARPS, HCU_ANA_WARPS == 8 ? 1 : 2) legendre_analysis_kernel(LegArgs a) {
  using K = Cfg<SPIN, NBLK>;
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *lam_w = smem_d + warp * K::WARP_SMEM;
  double *f_w = lam_w + K::LAM_WARP;
  double *coef_s = smem_d + K::COEF_OFF;
  int *flags = reinterpret_cast<int *>(smem_d + K::FLAG_OFF);

  const int ngroups = (int)((a.nrp_local + K::NT - 1) / K::NT);
  const int g = blockIdx.x % ngroups;
  const int mi = blockIdx.x / ngroups;
  const int m = a.mlist ? a.mlist[mi] : mi;
  const int lmax = a.lmax;
  const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
  if (l0 > lmax) return;
  const int pb = (l0 + m) & 1;

  const i64 rpl = (i64)g * K::NT + warp * 32 + lane;
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
  if (__syncthreads_or(alive ? 1 : 0) == 0) return;  // no ring of this CTA contributes

  const i64 cbase = alm_index(lmax, 0, m);  // coefficient / alm index of (l, m) is cbase + l

  // coefficients of chunk `chk` -> shared (entry i belongs to l = lstart + i; zero past lmax)
  auto stage_coef = [&](int chk) {
    const int lstart = l0 + chk * K::LC;
    for (int i = threadIdx.x; i < K::LC; i += K::NT) {
      const int l = lstart + i;
      if (SPIN == 0) {
        double2 cf = make_double2(0., 0.);
        if (l < lmax) cf = __ldg(reinterpret_cast<const double2 *>(a.coef) + cbase + l);
        reinterpret_cast<double2 *>(coef_s)[i] = cf;
      } else {
        double4 cf = make_double4(0., 0., 0., 0.);
        if (l < lmax) cf = ldg_d4(reinterpret_cast<const double4 *>(a.coef) + cbase + l);
        reinterpret_cast<double4 *>(coef_s)[i] = cf;
      }
    }
  };
  stage_coef(0);

  // ---- stage F (ring Fourier coefficients of this m) into shared memory ----
  {
    const i64 row0 = (i64)g * K::NT + warp * 32;
    const double *src = a.phase + ((i64)mi * a.nrp_local + row0) * a.ncomp * 4;
    const int rows = (int)max((i64)0, min((i64)32, a.nrp_local - row0));
    const int w = a.ncomp * 4;
    for (int idx = lane; idx < 32 * K::F_ROW; idx += 32) f_w[idx] = 0.0;
    __syncwarp();
    for (int idx = lane; idx < rows * w; idx += 32) {
      const int r = idx / w, cidx = idx - r * w;
      const double v = src[idx];
      int col;
      if (SPIN == 0) {
        // (re+, im+, re-, im-) of map c  ->  [parity][2c + ri]
        const int c = cidx >> 2, q = cidx & 3;
        col = ((q >> 1) * K::C + 2 * c + (q & 1)) ^ ((r & 3) << 2);
      } else {
        col = cidx ^ ((r & 3) << 2);  // raw: 8 doubles per field (Q: re+ im+ re- im-, U: ...)
      }
      f_w[r * K::F_ROW + col] = v;
    }
  }

  LamState sp, sm;
  sp.prev = sp.cur = 0; sp.e = 0;
  sm.prev = sm.cur = 0; sm.e = 0;
  if (alive) lam_start<SPIN>(m, a.cmtab, sth, chh, shh, sp, sm);
  __syncthreads();  // coefficients of chunk 0 and the F tiles are in place

  // fragment coordinates of this lane
  const int fa = lane & 3;   // k inside a k4 block (A column / B row)
  const int fb = lane >> 2;  // A row (l) / B column
  // spin 2: signed source offsets of the B fragments, see the E/B formulas below
  //   E_re = -F+ Q^s_re + F- U^-s_im     E_im = -F+ Q^s_im - F- U^-s_re
  //   B_re = -F+ U^s_re - F- Q^-s_im     B_im = -F+ U^s_im + F- Q^-s_re
  // parity 0: s = + (Q^s = Q+, U^-s = U-); parity 1: s = -.  Column = 4 field + h.
  int b_off[2][2];     // [j][parity] offset inside the field's 8 doubles
  double b_sgn[2];     // [j]
  if (SPIN != 0) {
    const int h = fb & 3;
    const int oP[4][2] = {{0, 2}, {1, 3}, {4, 6}, {5, 7}};
    const int oM[4][2] = {{7, 5}, {6, 4}, {3, 1}, {2, 0}};
    const double sM[4] = {1.0, -1.0, -1.0, 1.0};
    b_off[0][0] = oP[h][0]; b_off[0][1] = oP[h][1];
    b_off[1][0] = oM[h][0]; b_off[1][1] = oM[h][1];
    b_sgn[0] = -1.0;
    b_sgn[1] = sM[h];
  }

  const int nchunk = (lmax - l0 + K::LC) / K::LC;
  double n_rec = 0, n_acc = 0;

  for (int chk = 0; chk < nchunk; ++chk) {
    const int lstart = l0 + chk * K::LC;
    bool live = false;
    if (warp_alive) {
      // ------------------------- phase A ---------------------------------
      const bool scaled = __any_sync(0xffffffffu, sp.e < 0 || (SPIN != 0 && sm.e < 0));
      const int ksw = swz(lane);
      double *row0 = lam_w + lane * K::LP;
#pragma unroll 2
      for (int s = 0; s < K::LC; s += 4) {
        double v[4], v2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (SPIN == 0) {
            const double2 cf = reinterpret_cast<const double2 *>(coef_s)[s + u];
            v[u] = (!scaled || sp.e == 0) ? sp.cur : 0.0;
            if (scaled) {
              lam_advance(sp, cf.x * x, cf.y);
            } else {
              const double nw = fma(cf.x * x, sp.cur, -(cf.y * sp.prev));
              sp.prev = sp.cur;
              sp.cur = nw;
            }
          } else {
            const double4 cf = reinterpret_cast<const double4 *>(coef_s)[s + u];
            const double lp = (!scaled || sp.e == 0) ? sp.cur : 0.0;
            const double lm = (!scaled || sm.e == 0) ? sm.cur : 0.0;
            v[u] = 0.5 * (lp + lm);
            v2[u] = 0.5 * (lp - lm);
            if (scaled) {
              lam_advance(sp, fma(cf.x, x, cf.y), cf.z);
              lam_advance(sm, fma(cf.x, x, -cf.y), cf.z);
            } else {
              const double np = fma(fma(cf.x, x, cf.y), sp.cur, -(cf.z * sp.prev));
              const double nm = fma(fma(cf.x, x, -cf.y), sm.cur, -(cf.z * sm.prev));
              sp.prev = sp.cur; sp.cur = np;
              sm.prev = sm.cur; sm.cur = nm;
            }
          }
        }
        // steps s, s+2 have parity pb; s+1, s+3 parity 1-pb; index within parity = s/2 (+1)
        const int col = (s >> 1) ^ ksw;
        *reinterpret_cast<double2 *>(row0 + pb * K::TILE + col) = make_double2(v[0], v[2]);
        *reinterpret_cast<double2 *>(row0 + (1 - pb) * K::TILE + col) = make_double2(v[1], v[3]);
        if (SPIN != 0) {
          *reinterpret_cast<double2 *>(row0 + (2 + pb) * K::TILE + col) = make_double2(v2[0], v2[2]);
          *reinterpret_cast<double2 *>(row0 + (3 - pb) * K::TILE + col) = make_double2(v2[1], v2[3]);
        }
      }
      const bool lane_live = alive && (sp.e == 0 || (SPIN != 0 && sm.e == 0));
      live = __any_sync(0xffffffffu, lane_live);
      n_rec += 1;
    }
    __syncwarp();

    if (live) {
      // ------------------------- phase B (DMMA) ---------------------------
      double acc[2][K::MB][NBLK][2];
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int mb = 0; mb < K::MB; ++mb)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) acc[p][mb][nb][0] = acc[p][mb][nb][1] = 0.0;
#pragma unroll 2
      for (int k4 = 0; k4 < 8; ++k4) {
        const int krow = 4 * k4 + fa;
        const int ksw = swz(krow);
        const double *lrow = lam_w + krow * K::LP;
        const double *frow = f_w + krow * K::F_ROW;
        double af[K::NJ][2][K::MB], bf[K::NJ][2][NBLK];
#pragma unroll
        for (int j = 0; j < K::NJ; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int mb = 0; mb < K::MB; ++mb)
              af[j][p][mb] = lrow[(2 * j + p) * K::TILE + ((mb * 8 + fb) ^ ksw)];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) {
            if (SPIN == 0) {
              bf[0][p][nb] = frow[(p * K::C + nb * 8 + fb) ^ (fa << 2)];
            } else {
              const int fld = nb * 2 + (fb >> 2);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                bf[j][p][nb] = b_sgn[j] * frow[(fld * 8 + b_off[j][p]) ^ (fa << 2)];
            }
          }
#pragma unroll
        for (int j = 0; j < K::NJ; ++j)
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int mb = 0; mb < K::MB; ++mb)
#pragma unroll
              for (int nb = 0; nb < NBLK; ++nb)
                dmma(acc[p][mb][nb][0], acc[p][mb][nb][1], af[j][p][mb], bf[j][p][nb]);
      }
      n_acc += 1;
      __syncwarp();
      // this warp's partial tile over its (now consumed) lambda tile, column major
      // with stride LC + 2 (conflict free for these stores and for the flush loads):
      //   out_w[col * (LC + 2) + p * LP + lidx]
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int mb = 0; mb < K::MB; ++mb)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) {
            double *o = lam_w + (nb * 8 + 2 * fa) * K::OS + p * K::LP + mb * 8 + fb;
            o[0] = acc[p][mb][nb][0];
            o[K::OS] = acc[p][mb][nb][1];
          }
    }
    if (lane == 0) flags[warp] = live ? 1 : 0;
    __syncthreads();
    // ------------------------- flush -------------------------------------
    {
      int fl_any = 0;
#pragma unroll
      for (int w = 0; w < K::NW; ++w) fl_any |= flags[w];
      if (fl_any) {
        for (int o = threadIdx.x; o < K::NOUT; o += K::NT) {
          // o = col * LC + (p * LP + lidx): consecutive threads read consecutive doubles
          const int col = o / K::LC, t = o - col * K::LC;
          const int p = t / K::LP, lidx = t - p * K::LP;
          const int l = lstart + 2 * lidx + ((p - pb) & 1);
          int row, ri;
          if (SPIN == 0) {
            row = col >> 1;
            ri = col & 1;
          } else {
            const int f = col >> 2, h = col & 3;
            row = 2 * f + (h >> 1);
            ri = h & 1;
          }
          if (l <= lmax && row < a.ncomp) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < K::NW; ++w)
              if (flags[w]) sum += smem_d[w * K::WARP_SMEM + col * K::OS + t];
            if (a.fl) sum *= a.fl[l];
            atomicAdd(a.alm.p[row] + 2 * (cbase + l) + ri, sum);
          }
        }
      }
      if (chk + 1 < nchunk) stage_coef(chk + 1);
    }
    __syncthreads();
  }
  if (lane == 0 && a.work && (n_rec > 0)) {
    atomicAdd(a.work, n_rec * 32.0 * K::LC);
    atomicAdd(a.work + 1, n_acc * 32.0 * K::LC);
  }
}

template <int SPIN, int NBLK>
int launch_analysis(hcu_ctx *ctx, const LegArgs &a) {
  using K = Cfg<SPIN, NBLK>;
  const int ngroups = (int)((a.nrp_local + K::NT - 1) / K::NT);
  const i64 nblocks = (i64)ngroups * a.nm;
  if (nblocks <= 0) return HCU_OK;
  HCU_CUDA(cudaFuncSetAttribute(legendre_analysis_kernel<SPIN, NBLK>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)K::SMEM_BYTES));
  legendre_analysis_kernel<SPIN, NBLK><<<(unsigned)nblocks, K::NT, K::SMEM_BYTES, ctx->stream>>>(a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

}  // namespace

int hcu_legendre_analysis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                          int spin, int ncomp, const double *phase,
                          const int32_t *mlist_dev, int nm, i64 rp_lo, i64 rp_hi,
                          const double *fl_dev, const hcu_ptrs &alm) {
  HCU_ARG(ncomp >= 1 && ncomp <= HCU_MAX_BATCH, "legendre batch size");
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
  a.work = ctx->work_counters;
  // 8 output columns per n-block: 4 spin-0 maps, or 2 spin-2 fields (4 Q/U rows)
  const int nblk = (ncomp + 3) / 4;
  if (spin == 0) {
    switch (nblk) {
      case 1: return launch_analysis<0, 1>(ctx, a);
      case 2: return launch_analysis<0, 2>(ctx, a);
      default: return launch_analysis<0, 3>(ctx, a);
    }
  } else {
    switch (nblk) {
      case 1: return launch_analysis<2, 1>(ctx, a);
      case 2: return launch_analysis<2, 2>(ctx, a);
      default: return launch_analysis<2, 3>(ctx, a);
    }
  }
}
