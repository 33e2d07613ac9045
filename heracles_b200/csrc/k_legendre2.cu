// k_legendre2.cu -- FP64 Legendre stage, second generation: one recursion chain per lane,
// 12 or 16 warps per SM.
//
// Replaces the libsharp/ducc Legendre loops behind hp.map2alm and (inside its default iter=3
// refinement) hp.alm2map -- heracles/healpy.py:183-189.  Bound: the FP64 pipe (DMMA 37.1 TFLOP/s,
// DFMA 36.0, ONE shared pipe; tools/dmma_peak.cu).
//
// Why a second generation: the first kernels (k_legendre.cu) keep the ring Fourier coefficients of
// 32 ring pairs x 16 columns x 2 lambda matrices in registers (128 of 252) and therefore run 2 warps per
// SM sub-partition; every warp is an in-order stream, so whatever one warp cannot issue (tile hand-over,
// flush, address arithmetic, shared-memory latency) is only hidden by ONE other warp: 59 % of the
// FP64 pipe (profiles/r01_ncu_full_legendre_analysis_spin2_c4shape.txt).  Here a warp works on 32
// "virtual rings" v, each lane running ONE scaled three-term recursion:
//     spin 0   v = ring pair (32 per warp), lambda_lm
//     spin 2   v = (j, ring pair), 16 ring pairs per warp x the two Wigner functions lambda^{+2}, lambda^{-2}
// so that spin 2 is the same skinny GEMM as spin 0 with the sum over j folded into K:
//     analysis   out[l][col] += sum_v Lam[l][v] B[v][col]      (B: 32 v x 16 columns = 32 doubles per lane)
//     synthesis  G[v][col]   += sum_l Lam[v][l] (s_l a_lm)[l][col]
// The register-resident operand halves (64 registers), the kernels fit 168 (12 warps) or 128 registers
// (16 warps per SM) and the pipe sees 3-4 independent in-order streams per sub-partition.
//
// Sub-chunk = 16 l.  Tile[t][v][8]: t = parity of the step inside the sub-chunk (NOT of l + m: the
// parity pb of the first l is folded into the operand fragments once per CTA, which makes every
// shared-memory offset of the hot loop a compile-time immediate), XOR-swizzled like the first
// generation (128-bit stores by the producer lane v, 64-bit fragment loads, conflict free).
// Values that are not representable yet (extended exponent e < 0) are stored as zeros by a
// predicated second store instead of being masked with LOP3s.
//
// Analysis flush: per sub-chunk every warp parks its 16 l x 16 column partial tile in a double-buffered
// shared tile; it is reduced over the warps one sub-chunk LATER (split-phase mbarrier, nobody waits)
// and added to alm with one RED.ADD.F64 per output (x s_l x fl[l]).
#include <type_traits>

#include "legendre_common.cuh"

namespace {

// positions (k4 step 0..7 of a live sub-chunk) at which the per-sub-chunk chores are issued, see the analysis kernel.
// The reduction comes last: it waits for the slowest warp of the previous sub-chunk.
#ifndef HCU_CHORE_STAGE
#define HCU_CHORE_STAGE 1
#endif
#ifndef HCU_CHORE_REDUCE
#define HCU_CHORE_REDUCE 5
#endif
#ifndef HCU_CHORE_SCALE
#define HCU_CHORE_SCALE 6  // after the reduction: the scale registers are reused for the current sub-chunk
#endif

constexpr int SL = 16;  // l per sub-chunk
constexpr int LC = 32;  // l per chunk (a_lm staging granularity of the synthesis)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// st.shared.v2.f64 of (a, b) when ok, of zeros otherwise
__device__ __forceinline__ void sts2_pred(unsigned addr, double a, double b, int ok) {
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.s32 q, %3, 0;\n @q st.shared.v2.f64 [%0], {%1, %2};\n"
      " @!q st.shared.v2.f64 [%0], {%4, %4};\n}" ::"r"(addr),
      "d"(a), "d"(b), "r"(ok), "d"(0.0));
}
__device__ __forceinline__ double xor_hi(double v, unsigned bits) {
  return __hiloint2double(__double2hiint(v) ^ (int)bits, __double2loint(v));
}

// ------------------------------------------------------------------------------------------
// one recursion chain per lane
// ------------------------------------------------------------------------------------------
// XSIGN: how the lambda^{-2} chain (A_l x - B_l) differs from the lambda^{+2} chain (A_l x + B_l):
//   true   the sign bit of B_l is flipped in every step (one LOP3 per step)
//   false  the chain runs the lambda^{+2} code on x' = -x: q'_l = (-1)^l q_l obeys q'_{l+1} = (A_l x' + B_l) q'_l - q'_{l-1};
//          the consumer folds (-1)^l into its operand fragments (analysis: the ring Fourier coefficients)
template <int SPIN, bool XSIGN = true>
struct Chain {
  double prev, cur, x;
  int e;
  unsigned sgn;    // spin 2, XSIGN: sign bit of the B_l term (0: lambda^{+2}, 0x80000000: lambda^{-2})
  unsigned o4[4];  // byte offset of the 128-bit store of step group q inside a tile[t]

  __device__ __forceinline__ void init(int lane) {
    prev = cur = x = 0.0;
    e = 0;
    sgn = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) o4[q] = 8u * (unsigned)(lane * 8 + 2 * (q ^ swzf(lane)));
  }
  // The coefficients of a step are fetched from shared memory TWO steps before they are used (c0 / c1 roll),
  // so that the in-order warp never waits for the load; v[] collects the four values of a step group.
  using CT = typename std::conditional<SPIN == 0, double, double2>::type;
  CT c0, c1;
  double v[4];
  __device__ __forceinline__ void begin(const double *cf) {
    c0 = reinterpret_cast<const CT *>(cf)[0];
    c1 = reinterpret_cast<const CT *>(cf)[1];
  }
  // step u (0..15) of the sub-chunk being produced; tn = shared address of its tile
  __device__ __forceinline__ void step(const double *cf, unsigned tn, int u, int ok) {
    const CT c = (u & 1) ? c1 : c0;
    if (u + 2 < SL) {
      if (u & 1)
        c1 = reinterpret_cast<const CT *>(cf)[u + 2];
      else
        c0 = reinterpret_cast<const CT *>(cf)[u + 2];
    }
    v[u & 3] = cur;
    double ax;
    if constexpr (SPIN == 0) {
      ax = c * x;
    } else {
      ax = XSIGN ? fma(c.x, x, xor_hi(c.y, sgn)) : fma(c.x, x, c.y);
    }
    const double nw = fma(ax, cur, -prev);
    prev = cur;
    cur = nw;
    if ((u & 3) == 3) {
      const int q = u >> 2;
      sts2_pred(tn + o4[q], v[0], v[2], ok);          // even steps -> tile[0]
      sts2_pred(tn + 2048u + o4[q], v[1], v[3], ok);  // odd steps  -> tile[1]
    }
  }
  __device__ __forceinline__ void step4(const double *cf, unsigned tn, int q, int ok) {
#pragma unroll
    for (int u = 0; u < 4; ++u) step(cf, tn, 4 * q + u, ok);
  }
  __device__ __forceinline__ void sub16(const double *cf, unsigned tn, int ok) {
    begin(cf);
#pragma unroll
    for (int u = 0; u < SL; ++u) step(cf, tn, u, ok);
  }
  // extended-exponent bookkeeping, once per sub-chunk (values grow by far less than 2^400 in 16 steps).
  // The magnitude test reads the exponent field on the integer pipe: a DSETP would queue behind the DMMAs.
  __device__ __forceinline__ void end_sub() {
    if (__any_sync(0xffffffffu, e < 0)) {
      const int ec = __double2hiint(cur) & 0x7ff00000, ep = __double2hiint(prev) & 0x7ff00000;
      if (e < 0 && max(ec, ep) >= 0x4c700000) {  // |.| >= 2^200
        cur *= TWO_M400;
        prev *= TWO_M400;
        e += SCALE_STEP;
      }
    }
  }
};

// coefficients of the 16 steps starting at l = lsub: global -> shared without a register hop
// (cp.async, zero fill past lmax - 1).  spin 0: 16 doubles A_l; spin 2: 16 double2 (A_l, B_l).
template <int SPIN>
__device__ __forceinline__ void stage_coef2(double *cf, const LegArgs &a, i64 cbase, int lsub, int lane) {
  if (lane < SL) {
    const int l = lsub + lane;
    const int lc = min(l, max(a.lmax - 1, 0));
    if (SPIN == 0) {
      const unsigned nbytes = (l < a.lmax) ? 8u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(cf + lane)),
                   "l"(a.coef + cbase + lc), "r"(nbytes)
                   : "memory");
    } else {
      const unsigned nbytes = (l < a.lmax) ? 16u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(cf + 2 * lane)),
                   "l"(reinterpret_cast<const double2 *>(a.coef) + cbase + lc), "r"(nbytes)
                   : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

struct Setup2 {
  int mi, m, l0, pb, nrows, nsub;
  i64 row0, cbase, nrp_b, rp_base, poff;
  double *out;
};

// CTA geometry, ring constants and recursion start values.  RW = ring pairs per warp, R = per CTA.
template <int SPIN, int NW, bool XSIGN>
__device__ __forceinline__ bool setup2(const LegArgs &a, Setup2 &s, Chain<SPIN, XSIGN> &ch, bool &alive, int warp, int lane) {
  constexpr int RW = SPIN == 0 ? 32 : 16;
  constexpr int R = NW * RW;
  const int ngroups = a.grp_start[a.nblk];
  const int g = blockIdx.x % ngroups;
  s.mi = blockIdx.x / ngroups;
  int b = 0;
  while (b + 1 < a.nblk && g >= a.grp_start[b + 1]) ++b;
  s.nrp_b = a.blk_rp[b + 1] - a.blk_rp[b];
  s.rp_base = a.blk_rp[b];
  s.poff = (i64)a.nm * a.ncomp * 4 * (a.blk_rp[b] - a.blk_rp[0]);
  s.out = a.use_blk_out ? a.blk_out[b] : a.phase_out + s.poff;
  s.m = a.mlist ? a.mlist[s.mi] : s.mi;
  s.l0 = (SPIN == 0) ? s.m : (s.m > 2 ? s.m : 2);
  s.row0 = (i64)(g - a.grp_start[b]) * R;
  s.nrows = (int)min((i64)R, s.nrp_b - s.row0);
  s.pb = (s.l0 + s.m) & 1;
  s.cbase = alm_index(a.lmax, 0, s.m);
  s.nsub = (a.lmax - s.l0 + SL) / SL;
  ch.init(lane);
  alive = false;
  if (s.l0 > a.lmax) return false;
  const int r = warp * RW + (SPIN == 0 ? lane : (lane & 15));
  const int j = SPIN == 0 ? 0 : (lane >> 4);
  double sth = 1, chh = 1, shh = 1;
  if (r < s.nrows) {
    const i64 rp = s.rp_base + s.row0 + r;
    ch.x = a.cth[rp];
    sth = a.sth[rp];
    chh = a.ch[rp];
    shh = a.sh[rp];
    alive = !ring_is_dead(a.lmax, s.m, SPIN, ch.x, sth);
  }
  if (__syncthreads_or(alive ? 1 : 0) == 0) return false;  // no ring of this CTA contributes
  if (alive) {
    LamState sp, sm;
    sp.prev = sp.cur = 0; sp.e = 0;
    sm.prev = sm.cur = 0; sm.e = 0;
    lam_start<SPIN>(s.m, a.cmtab, sth, chh, shh, sp, sm);
    const LamState &pick = (j == 0) ? sp : sm;
    ch.prev = pick.prev;
    ch.cur = pick.cur;
    ch.e = pick.e;
    if (!XSIGN && j == 1) {  // q'_l = (-1)^l q_l on x' = -x
      ch.x = -ch.x;
      if (s.l0 & 1) ch.cur = -ch.cur;
    }
  } else {
    ch.x = 0.0;
  }
  ch.sgn = (j == 0) ? 0u : 0x80000000u;
  return true;
}

// ------------------------------------------------------------------------------------------
// analysis
// ------------------------------------------------------------------------------------------
template <int SPIN, int NW, int NBLK>
struct A2Cfg {
  static constexpr int RW = SPIN == 0 ? 32 : 16;
  static constexpr int R = NW * RW;
  static constexpr int NT = 32 * NW;
  static constexpr int C = 8 * NBLK;                      // output columns
  static constexpr int TILE = 2 * 256;                    // [t][v][8]
  static constexpr int WARP = 2 * TILE + 2 * SL * 2;      // two tiles + two coefficient buffers
  static constexpr int FS = SL + 2;                       // column stride of a flush tile (conflict free)
  static constexpr int FLUSH = NW * C * FS;               // one buffer: [warp][col][FS]
  static constexpr int NOUT = SL * C;                     // outputs per sub-chunk and CTA
  static constexpr int NOPT = NOUT > NT ? NOUT / NT : 1;  // outputs per thread
  static constexpr int NSPLIT = NOUT > NT ? 1 : NT / NOUT;  // thread groups sharing the source warps of an output
  static constexpr int NSRC = NW / NSPLIT;
  static constexpr int FLAG_OFF = NW * WARP + 2 * FLUSH;  // 2 x NW ints (padded to 16), then two mbarriers
  static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)(FLAG_OFF + 16 + 2);
  static_assert(NW % NSPLIT == 0, "source warps must split evenly");
  static_assert(NOUT % NT == 0 || NOUT < NT, "flush mapping");
};

template <int SPIN, int NW, int NBLK>
__global__ void __launch_bounds__(32 * NW, 1) legendre_analysis2_kernel(LegArgs a) {
  using K = A2Cfg<SPIN, NW, NBLK>;
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Setup2 st;
  Chain<SPIN, false> ch;
  bool alive;
  if (!setup2<SPIN, NW, false>(a, st, ch, alive, warp, lane)) return;
  const int lmax = a.lmax;
  const i64 cbase = st.cbase;
  double *tiles = smem_d + warp * K::WARP;
  double *coefs = tiles + 2 * K::TILE;
  double *flush = smem_d + NW * K::WARP;
  int *flags = reinterpret_cast<int *>(smem_d + K::FLAG_OFF);
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_d + K::FLAG_OFF + 16);
  const unsigned tile_s = smem_u32(tiles);
  const bool warp_alive = __any_sync(0xffffffffu, alive);
  const int fa = lane & 3;   // k inside a k4 block (A column / B row)
  const int fb = lane >> 2;  // A row (l) / B column
  if (threadIdx.x == 0) {
    mbar_init(mbar, NW);
    mbar_init(mbar + 1, NW);
  }
  __syncthreads();

  // ---- B fragments: the ring Fourier coefficients of this warp's 32 virtual rings, resident in registers ----
  // bf[kk][t][nb]: virtual ring v = 4 kk + fa, step parity t (true parity p = pb ^ t), column nb * 8 + fb
  double bf[8][2][NBLK];
  {
    const double *src = a.phase + st.poff + ((i64)st.mi * st.nrp_b + st.row0) * a.ncomp * 4;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int v = 4 * kk + fa;
      const int r = warp * K::RW + (SPIN == 0 ? v : (v & 15));
      const int jj = SPIN == 0 ? 0 : (kk >> 2);
      const double *prow = src + (i64)r * a.ncomp * 4;
      const bool rok = warp_alive && r < st.nrows;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        const int col = nb * 8 + fb;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int p = st.pb ^ t;
          if (SPIN == 0) {
            // column 2c + ri of parity p  <-  (re+, im+, re-, im-)[2p + ri] of map c
            const int c = col >> 1, ri = col & 1;
            bf[kk][t][nb] = (rok && c < a.ncomp) ? prow[c * 4 + 2 * p + ri] : 0.0;
          } else {
            // column 4f + h, h = (E_re, E_im, B_re, B_im); raw per field: Q (re+ im+ re- im-), U (...)
            //   E_re = -F+ Q^s_re + F- U^-s_im     E_im = -F+ Q^s_im - F- U^-s_re
            //   B_re = -F+ U^s_re - F- Q^-s_im     B_im = -F+ U^s_im + F- Q^-s_re
            // with F+- = (lam+ +- lam-)/2 and s = + for parity 0, - for parity 1, so the
            // operand of lam+ is (X + Y)/2 and that of lam- is (X - Y)/2.
            const int f = col >> 2, h = col & 3;
            const int oP = 4 * (h >> 1) + (h & 1) + 2 * p;
            const int oM = 4 * (1 - (h >> 1)) + (1 - (h & 1)) + 2 * (1 - p);
            const double sM = (h == 0 || h == 3) ? 1.0 : -1.0;
            double X = 0.0, Y = 0.0;
            if (rok && 2 * f < a.ncomp) {
              X = -prow[f * 8 + oP];
              Y = sM * prow[f * 8 + oM];
            }
            // the lambda^{-2} chain delivers (-1)^l lambda (see Chain): l = l0 + 16 k + 2 idx + t
            const double sg = ((st.l0 + t) & 1) ? -0.5 : 0.5;
            bf[kk][t][nb] = jj == 0 ? 0.5 * (X + Y) : sg * (X - Y);
          }
        }
      }
    }
  }

  // fragment load offsets inside a tile[t]: virtual ring 4 kk + fa, l index fb -- two variants (kk even / odd)
  const int a_off0 = lam_off(fa, fb), a_off1 = lam_off(4 + fa, fb) - 32;
  int n_rec = 0, n_acc = 0;

  // flush assignment of this thread: NOPT outputs (lo, col) of the sub-chunk, source warps [src0, src0 + NSRC)
  const bool f_active = threadIdx.x < K::NSPLIT * K::NOUT;
  const int f_lo = threadIdx.x & (SL - 1);  // l - lsub (the same for all outputs of the thread)
  const int f_src0 = (K::NOUT > K::NT) ? 0 : (threadIdx.x / K::NOUT) * K::NSRC;
  const int f_pos = (f_lo & 1) * 8 + (f_lo >> 1);  // position inside a flush column: [t][idx]
  int f_colv[K::NOPT];
  double *f_dst[K::NOPT];
  bool f_any = false;
#pragma unroll
  for (int i = 0; i < K::NOPT; ++i) {
    const int col = ((threadIdx.x + i * K::NT) % K::NOUT) / SL;
    int row, ri;
    if (SPIN == 0) {
      row = col >> 1;
      ri = col & 1;
    } else {
      row = 2 * (col >> 2) + ((col >> 1) & 1);
      ri = col & 1;
    }
    f_colv[i] = col;
    f_dst[i] = (f_active && row < a.ncomp) ? a.alm.p[row] + 2 * cbase + ri : nullptr;
    f_any = f_any || f_dst[i] != nullptr;
  }

  // the reduction of sub-chunk s over the warps is deferred to the end of sub-chunk s + 1, so that nobody
  // waits at a barrier: partial tiles and flags are double buffered, mbar[s & 1] counts the warps
  auto reduce_sub = [&](int s, int f_l, double f_sc, double f_fl) {
    mbar_wait(mbar + (s & 1), (s >> 1) & 1);
    if (!f_any || f_l > lmax) return;
    const int *fl = flags + (s & 1) * 16;
    int any = 0;
#pragma unroll
    for (int w = 0; w < K::NSRC; ++w) any |= fl[f_src0 + w];
    if (!any) return;
    const double sc = f_sc * f_fl;
#pragma unroll
    for (int i = 0; i < K::NOPT; ++i) {
      if (f_dst[i] == nullptr) continue;
      const double *fbuf = flush + (s & 1) * K::FLUSH + f_src0 * (K::C * K::FS) + f_colv[i] * K::FS + f_pos;
      // pairwise tree: the additions queue behind other warps' DMMAs, keep the dependent chain short
      double part[K::NSRC];
#pragma unroll
      for (int w = 0; w < K::NSRC; ++w) part[w] = fbuf[w * (K::C * K::FS)];
#pragma unroll
      for (int st = 1; st < K::NSRC; st *= 2)
#pragma unroll
        for (int w = 0; w + st < K::NSRC; w += 2 * st) part[w] += part[w + st];
      atomicAdd(f_dst[i] + 2 * (i64)f_l, part[0] * sc);
    }
  };

  // ---- prologue: coefficients and tile of sub-chunk 0 ----
  stage_coef2<SPIN>(coefs, a, cbase, st.l0, lane);
  coef_wait();
  __syncwarp();
  stage_coef2<SPIN>(coefs + SL * 2, a, cbase, st.l0 + SL, lane);  // coefficients of sub-chunk 1
  bool live_cur = false, live_nxt = false;
  if (warp_alive) {
    live_cur = __any_sync(0xffffffffu, alive && ch.e == 0);
    ch.sub16(coefs, tile_s, ch.e == 0);
    ch.end_sub();
    n_rec += 1;
  }
  // scale (and window) of this thread's flush outputs of the sub-chunk whose reduction is pending
  int f_l = st.l0 + f_lo;
  double f_sc = 0.0, f_fl = 1.0;

  for (int s2 = 0; s2 < st.nsub; s2 += 2) {
#pragma unroll
    for (int sb = 0; sb < 2; ++sb) {
      const int sidx = s2 + sb;
      if (sidx >= st.nsub) break;
      const int lsub = st.l0 + sidx * SL;
      const double *tcur = tiles + sb * K::TILE;
      const unsigned tn = tile_s + (unsigned)((1 - sb) * K::TILE * 8);
      const double *ccur = coefs + (1 - sb) * (SL * 2);
      coef_wait();
      __syncwarp();  // tile `sidx` and the coefficients of sub-chunk sidx + 1 are in place
      double acc[2][NBLK][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) acc[t][nb][0] = acc[t][nb][1] = 0.0;
      // The per-sub-chunk chores -- the deferred reduction of sub-chunk sidx - 1, the scale loads for sub-chunk
      // sidx and the coefficient prefetch for sub-chunk sidx + 2 -- are issued BETWEEN the DMMAs of a live
      // sub-chunk: a warp cannot issue its next DMMA for 16 clocks anyway, and these instructions need no FP64 pipe.
      auto chore = [&](int which) {
        if (which == 0) {
#ifndef HCU_EXP_NOFLUSH
          if (sidx > 0) reduce_sub(sidx - 1, f_l, f_sc, f_fl);
#endif
        } else if (which == 1) {  // must follow chore 0: the registers now belong to sub-chunk sidx
          f_l = lsub + f_lo;
          if (f_any) {
            f_sc = ldg_pin(a.scale + cbase + min(f_l, lmax));
            if (a.fl) f_fl = ldg_pin(a.fl + min(f_l, lmax));
          }
        } else {
          // the recursion always runs one sub-chunk ahead (the one past the end is harmless: its
          // coefficients are zero and nothing reads it)
          if (warp_alive) stage_coef2<SPIN>(coefs + sb * (SL * 2), a, cbase, lsub + 2 * SL, lane);
          else asm volatile("cp.async.commit_group;" ::: "memory");
        }
      };
      if (warp_alive) {
        live_nxt = __any_sync(0xffffffffu, alive && ch.e == 0);
        const int ok = ch.e == 0;
        if (sidx + 1 < st.nsub) n_rec += 1;
        if (live_cur) {
          n_acc += 1;
          // DMMAs of sub-chunk sidx interleaved with the recursion of sub-chunk sidx + 1
          ch.begin(ccur);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const double *ta = tcur + ((kk & 1) ? a_off1 : a_off0) + 32 * kk;
            const double a0 = ta[0], a1 = ta[256];
#ifndef HCU_EXP_NOREC
            ch.step(ccur, tn, 2 * kk, ok);
#endif
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) dmma(acc[0][nb][0], acc[0][nb][1], a0, bf[kk][0][nb]);
#ifndef HCU_EXP_NOREC
            ch.step(ccur, tn, 2 * kk + 1, ok);
#endif
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) dmma(acc[1][nb][0], acc[1][nb][1], a1, bf[kk][1][nb]);
            if (kk == HCU_CHORE_STAGE) chore(2);
            if (kk == HCU_CHORE_REDUCE) chore(0);
            if (kk == HCU_CHORE_SCALE) chore(1);
          }
        } else {
          chore(2);
          ch.sub16(ccur, tn, ok);
          chore(0);
          chore(1);
        }
        ch.end_sub();
      } else {
        chore(2);
        chore(0);
        chore(1);
      }
#ifdef HCU_EXP_NOFLUSH
      {  // ablation: no partial tiles, no barrier, no reduction (results are wrong on purpose)
        double ssum = f_sc * f_fl;
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) ssum += acc[t][nb][0] + acc[t][nb][1];
        f_sc += ssum;
        live_cur = warp_alive ? live_nxt : false;
        continue;
      }
#endif
      // ---- park this sub-chunk's partial tile (its reduction happens during the NEXT sub-chunk) ----
      if (lane == 0) flags[sb * 16 + warp] = live_cur ? 1 : 0;
      {  // a warp without live values parks zeros: the reducers add all source warps unconditionally
        double *o = flush + sb * K::FLUSH + warp * (K::C * K::FS);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) {
            double *d = o + (nb * 8 + 2 * fa) * K::FS + t * 8 + fb;
            d[0] = acc[t][nb][0];
            d[K::FS] = acc[t][nb][1];
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(mbar + sb);
      live_cur = warp_alive ? live_nxt : false;
    }
  }
#ifdef HCU_EXP_NOFLUSH
  if (f_sc == 1.2345e-300) atomicAdd(a.alm.p[0], f_sc);
#else
  reduce_sub(st.nsub - 1, f_l, f_sc, f_fl);
#endif
  if (lane == 0 && a.work && n_rec > 0) {
    atomicAdd(a.work, n_rec * (double)(K::RW * SL));
    atomicAdd(a.work + 1, n_acc * (double)(K::RW * SL));
  }
}

// ------------------------------------------------------------------------------------------
// synthesis
// ------------------------------------------------------------------------------------------
template <int SPIN, int NW, int NBLK>
struct S2Cfg {
  static constexpr int RW = SPIN == 0 ? 32 : 16;
  static constexpr int R = NW * RW;
  static constexpr int NT = 32 * NW;
  static constexpr int C = 8 * NBLK;
  static constexpr int NROW = 4 * NBLK;                   // alm rows (components) staged per l
  static constexpr int BSTR = 20;                         // row stride of the a_lm tile: 4 (mod 16), >= 2 NROW
  static constexpr int TILE = 2 * 256;
  static constexpr int TPAD = 2;                          // doubles between the two t-blocks: rows r and 16 + r of a tile
                                                          // are touched by neighbouring lanes (bank conflict free with it)
  static constexpr int TSTR = 16 * BSTR + TPAD;           // stride of a t-block
  static constexpr int BT = 2 * TSTR;                     // one a_lm tile: [t][16][BSTR]
  static constexpr int WARP = 2 * TILE + 2 * SL * 2 + 2 * BT;
  static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)(NW * WARP);
  static_assert(2 * NROW <= BSTR, "a_lm row does not fit");
};

template <int SPIN, int NW, int NBLK>
__global__ void __launch_bounds__(32 * NW, 1) legendre_synthesis2_kernel(LegArgs a) {
  using K = S2Cfg<SPIN, NW, NBLK>;
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Setup2 st;
  Chain<SPIN, true> ch;
  bool alive;
  const bool active = setup2<SPIN, NW, true>(a, st, ch, alive, warp, lane);
  double *dst = st.out + ((i64)st.mi * st.nrp_b + st.row0) * a.ncomp * 4;
  if (!active) {  // every output row must be defined
    for (int i = threadIdx.x; i < st.nrows * a.ncomp * 4; i += K::NT) dst[i] = 0.0;
    return;
  }
  const int lmax = a.lmax;
  const i64 cbase = st.cbase;
  double *tiles = smem_d + warp * K::WARP;
  double *coefs = tiles + 2 * K::TILE;
  double *btiles = coefs + 2 * SL * 2;  // two a_lm tiles: rows [t][16], raw a_lm on arrival, s_l a_lm after park
  const unsigned tile_s = smem_u32(tiles);
  const bool warp_alive = __any_sync(0xffffffffu, alive);
  const int fa = lane & 3;   // k inside a k4 block (A column = l / B row = l)
  const int fb = lane >> 2;  // A row (virtual ring) / B column
  const int nchunk = (st.nsub + 1) / 2;

  if (!warp_alive) {
    for (int i = lane; i < K::RW * a.ncomp * 4; i += 32) {
      const int r = warp * K::RW + i / (a.ncomp * 4);
      if (r < st.nrows) dst[(i64)warp * K::RW * a.ncomp * 4 + i] = 0.0;
    }
    return;
  }

  // spin 0: acc[t][mb][nb]; spin 2: acc[0][mb][block], block 0 = (+2a) columns, block 1 = (-2a) columns
  constexpr int NT_ACC = SPIN == 0 ? 2 : 1;
  double acc[NT_ACC][4][NBLK][2];
#pragma unroll
  for (int i = 0; i < NT_ACC; ++i)
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) acc[i][mb][nb][0] = acc[i][mb][nb][1] = 0.0;

  // a_lm of one chunk: lane = l row of the chunk; row of the tile: (lane & 1) * 16 + (lane >> 1) = [t][idx]
  const int brow = (lane & 1) * K::TSTR + (lane >> 1) * K::BSTR;
  auto fetch_alm = [&](int chk) {  // cp.async of the raw rows (zero fill past lmax), plus s_l into the last slot
    double *bt = btiles + (chk & 1) * K::BT + brow;
    const int l = st.l0 + chk * LC + lane;
    const int lc = min(l, lmax);
    const unsigned nbytes = (l <= lmax) ? 16u : 0u;
#pragma unroll
    for (int i = 0; i < K::NROW; ++i) {
      if (i < a.ncomp) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(bt + 2 * i)),
                     "l"(reinterpret_cast<const double2 *>(a.alm.p[i]) + cbase + lc), "r"(nbytes)
                     : "memory");
      }
    }
    const unsigned nb8 = (l <= lmax) ? 8u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(bt + K::BSTR - 1)),
                 "l"(a.scale + cbase + lc), "r"(nb8)
                 : "memory");
  };
  // in place: raw a_lm -> the B operand rows (x s_l; spin 2: +2a = -(E + iB), -2a = -(E - iB))
  auto park_alm = [&](int chk) {
    double *row = btiles + (chk & 1) * K::BT + brow;
    const double sc = row[K::BSTR - 1];
    if constexpr (SPIN == 0) {
#pragma unroll
      for (int i = 0; i < K::NROW; ++i) {
        double2 v = *reinterpret_cast<double2 *>(row + 2 * i);
        if (i >= a.ncomp) v = make_double2(0., 0.);
        *reinterpret_cast<double2 *>(row + 2 * i) = make_double2(v.x * sc, v.y * sc);
      }
    } else {
      // columns: [ +2a of the NF fields | sg -2a of the NF fields ], sg = (-1)^(l+m): with the sign of the southern
      // ring folded into the -2a half, ONE operand tile serves both Wigner functions (the lambda^{-2} blocks flip
      // the sign of their A fragment instead) and two fields fill exactly one n-block
      constexpr int NF = 2 * NBLK;
      const double sg = ((st.pb ^ (lane & 1)) ? -1.0 : 1.0) * sc;
      double2 E[NF], B[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        E[f] = *reinterpret_cast<double2 *>(row + 4 * f);
        B[f] = *reinterpret_cast<double2 *>(row + 4 * f + 2);
        if (2 * f >= a.ncomp) E[f] = B[f] = make_double2(0., 0.);
      }
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        // +2a = -(E + iB), -2a = -(E - iB)
        *reinterpret_cast<double2 *>(row + 2 * f) = make_double2(-(E[f].x - B[f].y) * sc, -(E[f].y + B[f].x) * sc);
        *reinterpret_cast<double2 *>(row + 2 * NF + 2 * f) = make_double2(-(E[f].x + B[f].y) * sg, -(E[f].y - B[f].x) * sg);
      }
    }
  };

  // spin 2: sign (-1)^(l+m) of the southern ring on the A fragment: true parity p = pb ^ t
  const unsigned sflip0 = (st.pb ^ 0) ? 0x80000000u : 0u, sflip1 = (st.pb ^ 1) ? 0x80000000u : 0u;

  bool live_cur = false, live_nxt = false;
  // ---- prologue: a_lm of chunk 0 (and 1), coefficients and tile of sub-chunk 0 ----
  fetch_alm(0);
  stage_coef2<SPIN>(coefs, a, cbase, st.l0, lane);  // commits the a_lm copies of chunk 0 too
  coef_wait();
  __syncwarp();
  park_alm(0);
  stage_coef2<SPIN>(coefs + SL * 2, a, cbase, st.l0 + SL, lane);  // coefficients of sub-chunk 1
  {
    live_cur = __any_sync(0xffffffffu, alive && ch.e == 0);
    ch.sub16(coefs, tile_s, ch.e == 0);
    ch.end_sub();
  }

  for (int chk = 0; chk < nchunk; ++chk) {
    const double *bt = btiles + (chk & 1) * K::BT;
#pragma unroll
    for (int sb = 0; sb < 2; ++sb) {
      const int sidx = 2 * chk + sb;
      const int lsub = st.l0 + sidx * SL;
      const double *tcur = tiles + sb * K::TILE;
      const unsigned tn = tile_s + (unsigned)((1 - sb) * K::TILE * 8);
      const double *ccur = coefs + (1 - sb) * (SL * 2);
      coef_wait();   // coefficients of sub-chunk sidx + 1 (and, at sb = 1, the a_lm of chunk chk + 1) have landed
      __syncwarp();  // tile `sidx` is complete, everybody is done with the buffers refilled below
      // chores issued between the DMMAs of a live sub-chunk (see the analysis kernel): the coefficient prefetch for
      // sub-chunk sidx + 2; at sb = 0 the a_lm fetch of chunk chk + 1 (into the tile chunk chk - 1 used), at sb = 1 its
      // in-place conversion (the data landed before the coef_wait above)
      auto chore = [&](int which) {
        if (which == 0) {
          if (sb == 0 && chk + 1 < nchunk) fetch_alm(chk + 1);
          stage_coef2<SPIN>(coefs + sb * (SL * 2), a, cbase, lsub + 2 * SL, lane);
        } else {
          if (sb == 1 && chk + 1 < nchunk) park_alm(chk + 1);
        }
      };
      live_nxt = __any_sync(0xffffffffu, alive && ch.e == 0);
      const int ok = ch.e == 0;
      if (live_cur) {
        ch.begin(ccur);
        // k4 steps: (step parity t, half h) -> rows t*16 + sb*8 + 4h + fa of the a_lm tile, l index 4h + fa of the Lam tile
#pragma unroll
        for (int th = 0; th < 4; ++th) {
          const int t = th >> 1, h = th & 1;
          const double *brw = bt + t * K::TSTR + (sb * 8 + 4 * h + fa) * K::BSTR + fb;
          double bfr[NBLK];
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) bfr[nb] = brw[nb * 8];
#pragma unroll
          for (int mb = 0; mb < 4; ++mb) {
            const double av = tcur[t * 256 + lam_off(mb * 8 + fb, 4 * h + fa)];
            ch.step(ccur, tn, 4 * th + mb, ok);
            if constexpr (SPIN == 0) {
#pragma unroll
              for (int nb = 0; nb < NBLK; ++nb) dmma(acc[t][mb][nb][0], acc[t][mb][nb][1], av, bfr[nb]);
            } else {
              // v-blocks 0, 1: lambda^{+2} of ring pairs 0..7, 8..15; 2, 3: lambda^{-2}; columns [ +2a | sg -2a ]:
              //   j = 0:  acc[mb][+] = P_N = sum lam+ (+2a)          acc[mb][-] = M_S = sum lam+ sg (-2a)
              //   j = 1:  acc[mb][+] = P_S = sum (sg lam-) (+2a)     acc[mb][-] = M_N = sum (sg lam-) sg (-2a)
              // sg = (-1)^(l+m) goes onto the A fragment of the lambda^{-2} blocks (integer pipe)
              const double af = mb < 2 ? av : xor_hi(av, t ? sflip1 : sflip0);
#pragma unroll
              for (int nb = 0; nb < NBLK; ++nb) dmma(acc[0][mb][nb][0], acc[0][mb][nb][1], af, bfr[nb]);
            }
          }
          if (th == 0) chore(0);
          if (th == 2) chore(1);
        }
      } else {
        chore(0);
        ch.sub16(ccur, tn, ok);
        chore(1);
      }
      ch.end_sub();
      live_cur = live_nxt;
    }
  }

  // ---- results straight from the accumulator fragments ----
  if constexpr (SPIN == 0) {
    // lane holds virtual ring mb*8 + fb, columns 2 fa, 2 fa + 1 of block nb = (re, im) of map nb*4 + fa
    const double ss = st.pb ? -1.0 : 1.0;  // T_(p=0) - T_(p=1) = ss (T_(t=0) - T_(t=1))
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int r = warp * K::RW + mb * 8 + fb;
      if (r >= st.nrows) continue;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        const int c = nb * 4 + fa;
        if (c < a.ncomp) {
          const double t0r = acc[0][mb][nb][0], t0i = acc[0][mb][nb][1];
          const double t1r = acc[1][mb][nb][0], t1i = acc[1][mb][nb][1];
          *reinterpret_cast<double4 *>(dst + ((i64)r * a.ncomp + c) * 4) =
              make_double4(t0r + t1r, t0i + t1i, ss * (t0r - t1r), ss * (t0i - t1i));
        }
      }
    }
  } else {
    // ring pair rb*8 + fb (rb = 0, 1): lambda^{+2} sums in v-block rb, lambda^{-2} sums in v-block 2 + rb
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      const int r = warp * K::RW + rb * 8 + fb;
      if constexpr (NBLK == 2) {
        // columns 2 fa, 2 fa + 1 of block 0 (+2a) and of block 1 (-2a): field fa
        const int f = fa;
        if (r >= st.nrows || 2 * f >= a.ncomp) continue;
        const double PrN = acc[0][rb][0][0], PiN = acc[0][rb][0][1];
        const double MrS = acc[0][rb][1][0], MiS = acc[0][rb][1][1];
        const double PrS = acc[0][2 + rb][0][0], PiS = acc[0][2 + rb][0][1];
        const double MrN = acc[0][2 + rb][1][0], MiN = acc[0][2 + rb][1][1];
        double *d = dst + ((i64)r * a.ncomp + 2 * f) * 4;
        // Q = (P + M)/2 ; U = (P - M)/(2i)
        *reinterpret_cast<double4 *>(d) =
            make_double4(0.5 * (PrN + MrN), 0.5 * (PiN + MiN), 0.5 * (PrS + MrS), 0.5 * (PiS + MiS));
        *reinterpret_cast<double4 *>(d + 4) =
            make_double4(0.5 * (PiN - MiN), -0.5 * (PrN - MrN), 0.5 * (PiS - MiS), -0.5 * (PrS - MrS));
      } else {
        // one n-block: lanes fa = 0, 1 hold the +2a columns of fields 0, 1 (P_N in v-block rb, P_S in 2 + rb), lanes
        // fa = 2, 3 the -2a columns (M_S, M_N); partners (fa ^ 2) swap, the P lane writes the Q row, the M lane the U row
        const double x0r = acc[0][rb][0][0], x0i = acc[0][rb][0][1];
        const double x1r = acc[0][2 + rb][0][0], x1i = acc[0][2 + rb][0][1];
        const double y0r = __shfl_xor_sync(0xffffffffu, x0r, 2), y0i = __shfl_xor_sync(0xffffffffu, x0i, 2);
        const double y1r = __shfl_xor_sync(0xffffffffu, x1r, 2), y1i = __shfl_xor_sync(0xffffffffu, x1i, 2);
        const int f = fa & 1;
        if (r >= st.nrows || 2 * f >= a.ncomp) continue;
        double *d = dst + ((i64)r * a.ncomp + 2 * f) * 4;
        if (fa < 2) {  // P lane: P_N = x0, P_S = x1, M_S = y0, M_N = y1
          *reinterpret_cast<double4 *>(d) = make_double4(0.5 * (x0r + y1r), 0.5 * (x0i + y1i), 0.5 * (x1r + y0r), 0.5 * (x1i + y0i));
        } else {       // M lane: M_S = x0, M_N = x1, P_N = y0, P_S = y1
          *reinterpret_cast<double4 *>(d + 4) =
              make_double4(0.5 * (y0i - x1i), -0.5 * (y0r - x1r), 0.5 * (y1i - x0i), -0.5 * (y1r - x0r));
        }
      }
    }
  }
}

template <int SPIN, int NW, int NBLK>
int launch_analysis2(hcu_ctx *ctx, LegArgs &a, const i64 *rp_bounds) {
  using K = A2Cfg<SPIN, NW, NBLK>;
  a.grp_start[0] = 0;
  for (int b = 0; b < a.nblk; ++b)
    a.grp_start[b + 1] = a.grp_start[b] + (int)((rp_bounds[b + 1] - rp_bounds[b] + K::R - 1) / K::R);
  const i64 nblocks = (i64)a.grp_start[a.nblk] * a.nm;
  if (nblocks <= 0) return HCU_OK;
  HCU_CUDA(cudaFuncSetAttribute(legendre_analysis2_kernel<SPIN, NW, NBLK>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
  legendre_analysis2_kernel<SPIN, NW, NBLK><<<(unsigned)nblocks, K::NT, K::SMEM_BYTES, ctx->stream>>>(a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

template <int SPIN, int NW, int NBLK>
int launch_synthesis2(hcu_ctx *ctx, LegArgs &a, const i64 *rp_bounds) {
  using K = S2Cfg<SPIN, NW, NBLK>;
  a.grp_start[0] = 0;
  for (int b = 0; b < a.nblk; ++b)
    a.grp_start[b + 1] = a.grp_start[b] + (int)((rp_bounds[b + 1] - rp_bounds[b] + K::R - 1) / K::R);
  const i64 nblocks = (i64)a.grp_start[a.nblk] * a.nm;
  if (nblocks <= 0) return HCU_OK;
  HCU_CUDA(cudaFuncSetAttribute(legendre_synthesis2_kernel<SPIN, NW, NBLK>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
  legendre_synthesis2_kernel<SPIN, NW, NBLK><<<(unsigned)nblocks, K::NT, K::SMEM_BYTES, ctx->stream>>>(a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

}  // namespace

// second-generation launchers; `a` comes filled from k_legendre.cu (everything but grp_start).
// nw: warps per CTA (12 or 16; 16 only where registers / shared memory allow).
int hcu_legendre2_analysis(hcu_ctx *ctx, void *args, const i64 *rp_bounds, int spin, int ncomp, int nw) {
  LegArgs &a = *reinterpret_cast<LegArgs *>(args);
  const int nblk = (ncomp + 3) / 4 <= 1 ? 1 : 2;  // 8 output columns per n-block: 4 spin-0 maps or 2 spin-2 fields
  if (ncomp > 8) return spin == 0 ? launch_analysis2<0, 8, 4>(ctx, a, rp_bounds) : launch_analysis2<2, 8, 4>(ctx, a, rp_bounds);
  if (spin == 0) {
    if (nw == 16) return nblk == 1 ? launch_analysis2<0, 16, 1>(ctx, a, rp_bounds) : launch_analysis2<0, 16, 2>(ctx, a, rp_bounds);
    return nblk == 1 ? launch_analysis2<0, 12, 1>(ctx, a, rp_bounds) : launch_analysis2<0, 12, 2>(ctx, a, rp_bounds);
  }
  if (nw == 16) return nblk == 1 ? launch_analysis2<2, 16, 1>(ctx, a, rp_bounds) : launch_analysis2<2, 16, 2>(ctx, a, rp_bounds);
  return nblk == 1 ? launch_analysis2<2, 12, 1>(ctx, a, rp_bounds) : launch_analysis2<2, 12, 2>(ctx, a, rp_bounds);
}

int hcu_legendre2_synthesis(hcu_ctx *ctx, void *args, const i64 *rp_bounds, int spin, int ncomp, int nw) {
  LegArgs &a = *reinterpret_cast<LegArgs *>(args);
  (void)nw;
  if (ncomp > 8) {
    hcu_set_error("synthesis takes at most 8 components per launch");
    return HCU_ERR_UNSUPPORTED;
  }
  if (spin == 0) {
    const int nblk = (ncomp + 3) / 4 <= 1 ? 1 : 2;
    return nblk == 1 ? launch_synthesis2<0, 12, 1>(ctx, a, rp_bounds) : launch_synthesis2<0, 12, 2>(ctx, a, rp_bounds);
  }
  return ncomp <= 4 ? launch_synthesis2<2, 12, 1>(ctx, a, rp_bounds) : launch_synthesis2<2, 12, 2>(ctx, a, rp_bounds);
}
