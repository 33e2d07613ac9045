// k_legendre.cu -- FP64 Legendre stage of the spherical-harmonic transform on HEALPix
// ring pairs: ANALYSIS (ring Fourier coefficients -> alm) and SYNTHESIS (alm -> ring
// Fourier coefficients), spin 0 and spin 2, batched over up to 12 (spin 0) / 8 (spin 2)
// components.
//
// Replaces the libsharp/ducc Legendre loops behind hp.map2alm and (inside its default
// iter=3 refinement) hp.alm2map -- heracles/healpy.py:183-189.  Bound: the FP64 pipe
// (DMMA 37.1 TFLOP/s, DFMA 36.0, ONE shared pipe; tools/dmma_peak.cu).
//
// One CTA = one m and 256 ring pairs; 8 warps; one warp = 32 ring pairs and does BOTH
// jobs for them, software pipelined in sub-chunks of 16 l:
//   recursion  one ring pair per lane: q_{l+1} = (A_l x +- B_l) q_l - q_{l-1} with
//       lambda_l = s_l q_l (two FMAs per step; the s_l are folded into the outputs resp.
//       the staged a_lm), extended-exponent rescaling while the value is below 2^-200.
//       The 16 values of sub-chunk i+1 go to the warp's second shared-memory tile
//       Lam[j][parity][ring][8] (j: spin 2 has lambda^+2 and lambda^-2), XOR-swizzled so
//       that these 128-bit stores and both kinds of fragment loads are conflict free,
//   tensor ops  while the DMMAs (mma.sync.m8n8k4.f64) of sub-chunk i read the first tile:
//         analysis   out[l][col] += sum_ring Lam[l][ring] B[ring][col]; the B fragments
//                    (ring Fourier coefficients of the warp's 32 rings, all columns, both
//                    parities) stay in REGISTERS for the whole CTA; per chunk of 32 l the
//                    eight warps' partial tiles are reduced through a double-buffered
//                    shared tile (one CTA barrier per chunk) and added to alm with one
//                    RED.ADD.F64 per output (x s_l x fl[l]).
//         synthesis  G[ring][col] += sum_l Lam[ring][l] (s_l a_lm)[l][col]; accumulators
//                    stay in registers for the whole CTA, the a_lm chunk is prefetched a
//                    chunk ahead into a warp-private shared tile; no CTA barrier at all.
//   The recursion coefficients of the next sub-chunk arrive by cp.async (global -> shared without
//   a register hop): a register-destination load that stays outstanding for a whole sub-chunk
//   ties up a scoreboard the loop's shared-memory loads then wait on (measured: 15 %).
// The recursion and the DMMAs are written into ONE instruction stream on purpose: a DFMA
// chain in a warp of its own is starved by the scheduler as soon as two other warps of the
// same SM sub-partition keep the FP64 pipe full of DMMAs (tools/mix_peak.cu: 17 000 clk per
// dependent step), while DFMAs interleaved with the DMMAs of the same warp are free.
//
// Work per (l, m, ring pair): recursion 2 (spin 0) / 4 (spin 2) FP64 ops shared by the
// batch; accumulate 4 flop per spin-0 map, 16 per spin-2 field (SURVEY 8(d) counts the
// recursion as 4 / 12 flop, which is what the reported flop numbers use).
#include <stdlib.h>

#include "legendre_common.cuh"

namespace {

constexpr int NW = 8;        // warps per CTA
constexpr int R = 32 * NW;   // ring pairs per CTA
constexpr int SL = 16;       // l per sub-chunk (8 per parity = one m8 block)
constexpr int LC = 32;       // l per chunk (flush / a_lm staging granularity)
constexpr int NT = 32 * NW;
constexpr int FLS = 34;      // column stride of a flush tile (32 l + 2: conflict free)


// EXPERIMENT (-DHCU_PINGPONG, off): ping-pong of the two warps that share an SM sub-partition (warps w and w + 4).  The
// idea: left alone, the scheduler alternates their DMMAs one by one, so both finish the tensor part of a sub-chunk at the
// same moment and walk through the hand-over (coefficient wait, votes, rescaling, flush) together while the FP64 pipe
// idles; with a token (named barriers 1..8, bar.sync on my own, bar.arrive on the partner's) one warp streams its 64
// DMMAs while the other is in its hand-over.  Measured (nside 2048, spin 2, 4 fields): analysis 149.2 -> 157.5 ms,
// synthesis 74.9 -> 75.9 ms -- SLOWER: a warp that has the pipe to itself needs ~27 clocks per DMMA of this
// instruction mix, i.e. the two warps of a sub-partition already overlap little more than their hand-overs.
#ifdef HCU_PINGPONG
__device__ __forceinline__ void pp_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pp_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
#else
__device__ __forceinline__ void pp_wait(int) {}
__device__ __forceinline__ void pp_pass(int) {}
#endif

// st.shared.v2.f64 of (a, b) when ok, of zeros otherwise
__device__ __forceinline__ void sts2_pred1(unsigned addr, double a, double b, int ok) {
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.s32 q, %3, 0;\n @q st.shared.v2.f64 [%0], {%1, %2};\n"
      " @!q st.shared.v2.f64 [%0], {%4, %4};\n}" ::"r"(addr),
      "d"(a), "d"(b), "r"(ok), "d"(0.0));
}

template <int SPIN>
struct Rec {
  static constexpr int NJ = SPIN == 0 ? 1 : 2;
  static constexpr int TILE = NJ * 2 * 256;  // doubles of one sub-chunk tile of a warp
  LamState sp, sm;
  double x;
  unsigned long long mp, mm;  // store masks: all ones when the value is representable (e == 0)
  bool alive;
  int ssp, ssm;               // sub-chunk at whose start the chain is first representable (start-state table; 0 without)
  i64 st_idx;                 // this lane's entry of the start-state table
  int o4[4];                  // swizzled store offsets of the four step groups of a sub-chunk (lane constant)
  __device__ __forceinline__ void init_offsets(int ring) {
#pragma unroll
    for (int q = 0; q < 4; ++q) o4[q] = ring * 8 + 2 * (q ^ swzf(ring));
  }

  __device__ __forceinline__ void begin_sub() {
    mp = (sp.e == 0) ? ~0ull : 0ull;
    mm = (sm.e == 0) ? ~0ull : 0ull;
  }
  // is anything of sub-chunk s, which is about to be produced, representable in this warp?
  __device__ __forceinline__ bool sub_live(int s) const {
    return __any_sync(0xffffffffu, alive && ((sp.e == 0 && ssp <= s) || (SPIN != 0 && sm.e == 0 && ssm <= s)));
  }
  // chains whose dead zone ends at sub-chunk s pick up their state from the table (hcu_start): until then they carry
  // zeros, exactly what the extended-exponent masks produced while the dead zone was still walked
  __device__ __forceinline__ void maybe_start(const LegArgs &a, int s) {
    if (a.st_state == nullptr || !alive) return;
    if (ssp == s) {
      const double2 v = a.st_state[st_idx];
      sp.prev = v.x;
      sp.cur = v.y;
      sp.e = 0;
    }
    if (SPIN != 0 && ssm == s) {
      const double2 v = a.st_state[st_idx + 1];
      sm.prev = v.x;
      sm.cur = v.y;
      sm.e = 0;
    }
  }
  // four recursion steps s..s+3 of the current sub-chunk (s multiple of 4) and their stores.
  // Values that are not representable yet (extended exponent e < 0) are stored as zeros by a predicated second
  // store (st.shared of RZ) instead of being masked: 16 LOP3 and as many register moves less per step group.
  __device__ __forceinline__ void step4(double *tile, const double *cf, int ring, int pb, int s) {
    double vp[4], vm[4];
    double2 c2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c2[u] = reinterpret_cast<const double2 *>(cf)[s + u];  // all four loads before the chain
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vp[u] = sp.cur;
      if (SPIN == 0) {
        const double nw = fma(c2[u].x * x, sp.cur, -sp.prev);
        sp.prev = sp.cur;
        sp.cur = nw;
      } else {
        vm[u] = sm.cur;
        const double ap = fma(c2[u].x, x, c2[u].y), am = fma(c2[u].x, x, -c2[u].y);
        const double np = fma(ap, sp.cur, -sp.prev);
        const double nm = fma(am, sm.cur, -sm.prev);
        sp.prev = sp.cur; sp.cur = np;
        sm.prev = sm.cur; sm.cur = nm;
      }
    }
    // steps s, s+2 have parity pb; s+1, s+3 parity 1-pb; l index within the parity s/2, s/2+1
    const unsigned o = (unsigned)__cvta_generic_to_shared(tile) + 8u * (unsigned)o4[s >> 2];
    const int okp = mp != 0ull, okm = mm != 0ull;
    sts2_pred1(o + 2048u * pb, vp[0], vp[2], okp);
    sts2_pred1(o + 2048u * (1 - pb), vp[1], vp[3], okp);
    if (SPIN != 0) {
      sts2_pred1(o + 2048u * (2 + pb), vm[0], vm[2], okm);
      sts2_pred1(o + 2048u * (3 - pb), vm[1], vm[3], okm);
    }
  }
  // extended-exponent bookkeeping, once per sub-chunk (values grow by far less than 2^400 in 16 steps);
  // the magnitude test reads the exponent fields on the integer pipe (a DSETP would queue behind the DMMAs)
  __device__ __forceinline__ void end_sub() {
    if (__any_sync(0xffffffffu, sp.e < 0 || (SPIN != 0 && sm.e < 0))) {
      if (sp.e < 0 && max(__double2hiint(sp.cur) & 0x7ff00000, __double2hiint(sp.prev) & 0x7ff00000) >= 0x4c700000) {
        sp.cur *= TWO_M400;
        sp.prev *= TWO_M400;
        sp.e += SCALE_STEP;
      }
      if (SPIN != 0 && sm.e < 0 &&
          max(__double2hiint(sm.cur) & 0x7ff00000, __double2hiint(sm.prev) & 0x7ff00000) >= 0x4c700000) {
        sm.cur *= TWO_M400;
        sm.prev *= TWO_M400;
        sm.e += SCALE_STEP;
      }
    }
  }
};


// lanes 0..15 fetch the recursion coefficients of the 16 steps starting at l = lsub
template <int SPIN>
__device__ __forceinline__ double2 fetch_coef(const LegArgs &a, i64 cbase, int lsub, int lane) {
  double2 c2 = make_double2(0., 0.);
  if (lane < SL) {
    const int l = lsub + lane;
    const int lc = min(l, max(a.lmax - 1, 0));  // clamped: the load is unconditional
    if (SPIN == 0)
      c2.x = ldg_pin(a.coef + cbase + lc);
    else
      c2 = ldg_pin2(reinterpret_cast<const double2 *>(a.coef) + cbase + lc);
  }
  return c2;
}
// asynchronous variant: global -> shared without a register hop (cp.async, zero fill past lmax);
// completion is awaited with coef_wait() a sub-chunk later
template <int SPIN>
__device__ __forceinline__ void stage_coef_async(double *cf, const LegArgs &a, i64 cbase, int lsub, int lane) {
  if (lane < SL) {
    const int l = lsub + lane;
    const int lc = min(l, max(a.lmax - 1, 0));
    const unsigned dst = (unsigned)__cvta_generic_to_shared(cf + 2 * lane);
    if (SPIN == 0) {
      const unsigned nbytes = (l < a.lmax) ? 8u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(a.coef + cbase + lc), "r"(nbytes) : "memory");
    } else {
      const unsigned nbytes = (l < a.lmax) ? 16u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst),
                   "l"(reinterpret_cast<const double2 *>(a.coef) + cbase + lc), "r"(nbytes)
                   : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// the coefficients past lmax are zeroed here, a sub-chunk after the load was issued
__device__ __forceinline__ void park_coef(double *cf, double2 c2, int lsub, int lmax, int lane) {
  if (lane < SL) {
    if (lsub + lane >= lmax) c2 = make_double2(0., 0.);
    reinterpret_cast<double2 *>(cf)[lane] = c2;
  }
}


struct Setup {
  int g, mi, m, l0, pb, nrows, nchunk, chk0;  // chk0: first chunk in which any chain of the CTA is representable
  i64 row0, cbase;
  i64 nrp_b, rp_base, poff;  // ring pairs of this CTA's block, its first ring pair, its offset in phase
  double *out;                // base of the block in the synthesis output
};

template <int SPIN>
__device__ __forceinline__ bool setup_cta(const LegArgs &a, Setup &s, Rec<SPIN> &rec, int warp, int lane) {
  const int ngroups = a.grp_start[a.nblk];
  s.g = blockIdx.x % ngroups;
  s.mi = blockIdx.x / ngroups;
  int b = 0;
  while (b + 1 < a.nblk && s.g >= a.grp_start[b + 1]) ++b;
  s.nrp_b = a.blk_rp[b + 1] - a.blk_rp[b];
  s.rp_base = a.blk_rp[b];
  s.poff = (i64)a.nm * a.ncomp * 4 * (a.blk_rp[b] - a.blk_rp[0]);
  s.out = a.use_blk_out ? a.blk_out[b] : a.phase_out + s.poff;
  s.m = a.mlist ? a.mlist[s.mi] : s.mi;
  s.l0 = (SPIN == 0) ? s.m : (s.m > 2 ? s.m : 2);
  s.row0 = (i64)(s.g - a.grp_start[b]) * R;
  s.nrows = (int)min((i64)R, s.nrp_b - s.row0);
  s.pb = (s.l0 + s.m) & 1;
  s.cbase = alm_index(a.lmax, 0, s.m);
  s.nchunk = (a.lmax - s.l0 + LC) / LC;
  rec.sp.prev = rec.sp.cur = 0; rec.sp.e = 0;
  rec.sm.prev = rec.sm.cur = 0; rec.sm.e = 0;
  rec.x = 0;
  rec.alive = false;
  rec.init_offsets(lane);
  if (s.l0 > a.lmax) return false;
  const int r = warp * 32 + lane;
  double sth = 1, chh = 1, shh = 1;
  if (r < s.nrows) {
    const i64 rp = s.rp_base + s.row0 + r;
    rec.x = a.cth[rp];
    sth = a.sth[rp];
    chh = a.ch[rp];
    shh = a.sh[rp];
    rec.alive = !ring_is_dead(a.lmax, s.m, SPIN, rec.x, sth);
  }
  rec.ssp = rec.ssm = 0;
  rec.st_idx = 0;
  s.chk0 = 0;
  if (a.st_sub != nullptr) {
    // the dead zone was walked when the table was built: every chain starts where it first is representable
    __shared__ int s_first;
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();
    const int nsub = 2 * s.nchunk;
    int first = 0x7fffffff;
    if (rec.alive) {
      constexpr int NJ = SPIN == 0 ? 1 : 2;
      rec.st_idx = ((i64)s.m * a.st_nrp + (s.rp_base + s.row0 + r)) * NJ;
      rec.ssp = a.st_sub[rec.st_idx];
      rec.ssm = SPIN == 0 ? 0x7fffffff : a.st_sub[rec.st_idx + 1];
      first = min(rec.ssp, rec.ssm);
      rec.alive = first < nsub;
    }
    first = __reduce_min_sync(0xffffffffu, first);
    if (lane == 0 && first < nsub) atomicMin(&s_first, first);
    __syncthreads();
    if (s_first >= nsub) return false;  // nothing of this CTA ever becomes representable
    s.chk0 = s_first >> 1;
    return true;
  }
  if (__syncthreads_or(rec.alive ? 1 : 0) == 0) return false;  // no ring of this CTA contributes
  if (rec.alive) lam_start<SPIN>(s.m, a.cmtab, sth, chh, shh, rec.sp, rec.sm);
  return true;
}

// start-state table (hcu_start): one thread per (m, ring pair) walks the dead zone ONCE with exactly the arithmetic of
// Rec::step4 / end_sub (and of Chain in k_legendre2.cu), so that a chain started from the table continues bit for bit
// like one that walked there itself.  sub[.] = first sub-chunk (16 l from l0) at whose START the chain's extended
// exponent is 0; state[.] = (prev, cur) there.
template <int SPIN>
__global__ void start_table_kernel(int lmax, int nrp, const double *cth, const double *sth, const double *ch, const double *sh,
                                   const double *coef, const double *cmtab, int *sub, double2 *state) {
  constexpr int NJ = SPIN == 0 ? 1 : 2;
  const int m = blockIdx.y;
  const int rp = blockIdx.x * blockDim.x + threadIdx.x;
  if (rp >= nrp) return;
  const i64 o = ((i64)m * nrp + rp) * NJ;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    sub[o + j] = 0x7fffffff;
    state[o + j] = make_double2(0., 0.);
  }
  const int l0 = (SPIN == 0) ? m : (m > 2 ? m : 2);
  if (l0 > lmax) return;
  const double x = cth[rp];
  if (ring_is_dead(lmax, m, SPIN, x, sth[rp])) return;
  LamState sp, sm;
  sp.prev = sp.cur = 0; sp.e = 0;
  sm.prev = sm.cur = 0; sm.e = 0;
  lam_start<SPIN>(m, cmtab, sth[rp], ch[rp], sh[rp], sp, sm);
  const i64 cbase = alm_index(lmax, 0, m);
  const int nsub = 2 * ((lmax - l0 + LC) / LC);
  bool donep = false, donem = (SPIN == 0);
  for (int s = 0; s < nsub; ++s) {
    if (!donep && sp.e == 0) {
      sub[o] = s;
      state[o] = make_double2(sp.prev, sp.cur);
      donep = true;
    }
    if (SPIN != 0 && !donem && sm.e == 0) {
      sub[o + 1] = s;
      state[o + 1] = make_double2(sm.prev, sm.cur);
      donem = true;
    }
    if (donep && donem) break;
    for (int u = 0; u < SL; ++u) {
      const int l = l0 + s * SL + u;
      double cx = 0.0, cy = 0.0;  // the kernels' coefficient tiles are zero filled from lmax on
      if (l < lmax) {
        if (SPIN == 0) {
          cx = coef[cbase + l];
        } else {
          const double2 c2 = reinterpret_cast<const double2 *>(coef)[cbase + l];
          cx = c2.x;
          cy = c2.y;
        }
      }
      if (SPIN == 0) {
        const double nw = fma(cx * x, sp.cur, -sp.prev);
        sp.prev = sp.cur;
        sp.cur = nw;
      } else {
        const double ap = fma(cx, x, cy), am = fma(cx, x, -cy);
        const double np = fma(ap, sp.cur, -sp.prev);
        const double nm = fma(am, sm.cur, -sm.prev);
        sp.prev = sp.cur; sp.cur = np;
        sm.prev = sm.cur; sm.cur = nm;
      }
    }
    if (sp.e < 0 && max(__double2hiint(sp.cur) & 0x7ff00000, __double2hiint(sp.prev) & 0x7ff00000) >= 0x4c700000) {
      sp.cur *= TWO_M400;
      sp.prev *= TWO_M400;
      sp.e += SCALE_STEP;
    }
    if (SPIN != 0 && sm.e < 0 &&
        max(__double2hiint(sm.cur) & 0x7ff00000, __double2hiint(sm.prev) & 0x7ff00000) >= 0x4c700000) {
      sm.cur *= TWO_M400;
      sm.prev *= TWO_M400;
      sm.e += SCALE_STEP;
    }
  }
}

// ---------------------------------------------------------------------------------
// analysis
// ---------------------------------------------------------------------------------
template <int SPIN, int NBLK>
struct ACfg {
  static constexpr int NJ = SPIN == 0 ? 1 : 2;
  static constexpr int C = 8 * NBLK;
  static constexpr int TILE = Rec<SPIN>::TILE;
  static constexpr int WARP = 2 * TILE + 2 * SL * 2;       // two tiles + two coefficient buffers
  static constexpr int FLUSH = NW * C * FLS;               // one buffer: [warp][col][FLS]
  static constexpr int FLAG_OFF = NW * WARP + 2 * FLUSH;   // 2 x NW ints, then two mbarriers
  static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)(FLAG_OFF + NW + 2);
};

template <int SPIN, int NBLK>
__global__ void __launch_bounds__(NT, 1) legendre_analysis_kernel(LegArgs a) {
  using K = ACfg<SPIN, NBLK>;
  constexpr int NJ = K::NJ;
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Setup st;
  Rec<SPIN> rec;
  if (!setup_cta<SPIN>(a, st, rec, warp, lane)) return;
  const int lmax = a.lmax, pb = st.pb;
  const i64 cbase = st.cbase;
  double *tiles = smem_d + warp * K::WARP;
  double *coefs = tiles + 2 * K::TILE;
  double *flush = smem_d + NW * K::WARP;
  int *flags = reinterpret_cast<int *>(smem_d + K::FLAG_OFF);
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_d + K::FLAG_OFF + NW);
  const bool warp_alive = __any_sync(0xffffffffu, rec.alive);
  const int fa = lane & 3;   // k inside a k4 block (A column / B row)
  const int fb = lane >> 2;  // A row (l) / B column
  if (threadIdx.x == 0) {
    mbar_init(mbar, NW);
    mbar_init(mbar + 1, NW);
  }
  __syncthreads();
  const int pp_my = 1 + 2 * (warp & 3) + (warp >> 2), pp_other = 1 + 2 * (warp & 3) + (1 - (warp >> 2));
  if (warp >= 4) pp_pass(pp_other);  // the first turn belongs to warps 0..3

  // ---- B fragments of this warp's 32 rings, both parities, resident in registers ----
  double bf[8][NJ][2][NBLK];
  {
    const double *src = a.phase + st.poff + ((i64)st.mi * st.nrp_b + st.row0) * a.ncomp * 4;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const int r = warp * 32 + 4 * k4 + fa;
      const double *prow = src + (i64)r * a.ncomp * 4;
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        const int col = nb * 8 + fb;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          if (SPIN == 0) {
            // column 2c + ri of parity p  <-  (re+, im+, re-, im-)[2p + ri] of map c
            const int c = col >> 1, ri = col & 1;
            bf[k4][0][p][nb] = (warp_alive && r < st.nrows && c < a.ncomp) ? prow[c * 4 + 2 * p + ri] : 0.0;
          } else {
            // column 4f + h, h = (E_re, E_im, B_re, B_im); raw per field: Q (re+ im+ re- im-), U (...)
            //   E_re = -F+ Q^s_re + F- U^-s_im     E_im = -F+ Q^s_im - F- U^-s_re
            //   B_re = -F+ U^s_re - F- Q^-s_im     B_im = -F+ U^s_im + F- Q^-s_re
            // with F+- = (lam+ +- lam-)/2 and s = + for parity 0, - for parity 1, so the
            // operand of lam+ is (X + Y)/2 and that of lam- is (X - Y)/2.
            const int f = col >> 2, h = col & 3;
            // offsets inside the field's 8 doubles: oP = 4 (h >> 1) + (h & 1) + 2 p ; oM mirrors it
            const int oP = 4 * (h >> 1) + (h & 1) + 2 * p;
            const int oM = 4 * (1 - (h >> 1)) + (1 - (h & 1)) + 2 * (1 - p);
            const double sM = (h == 0 || h == 3) ? 1.0 : -1.0;
            double X = 0.0, Y = 0.0;
            if (warp_alive && r < st.nrows && 2 * f < a.ncomp) {
              X = -prow[f * 8 + oP];
              Y = sM * prow[f * 8 + oM];
            }
            bf[k4][0][p][nb] = 0.5 * (X + Y);
            bf[k4][1][p][nb] = 0.5 * (X - Y);
          }
        }
      }
    }
  }

  const int ring = lane;  // ring of this lane inside the warp's tile
  // fragment load offsets inside a (j, parity) tile: ring 4 k4 + fa, l index fb -- two variants (k4 even / odd)
  const int a_off0 = lam_off(fa, fb), a_off1 = lam_off(4 + fa, fb) - 32;
  const int nsub = 2 * st.nchunk;
  int n_rec = 0, n_acc = 0;
  // ---- prologue: the first sub-chunk of chunk chk0 (chunk 0 without a start-state table) ----
  const int chk0 = st.chk0, s0 = 2 * st.chk0;
  bool live_cur = false, live_nxt = false;
  park_coef(coefs, fetch_coef<SPIN>(a, cbase, st.l0 + s0 * SL, lane), st.l0 + s0 * SL, lmax, lane);
  stage_coef_async<SPIN>(coefs + SL * 2, a, cbase, st.l0 + (s0 + 1) * SL, lane);  // coefficients of the next sub-chunk
  __syncwarp();
  if (warp_alive) {
    rec.maybe_start(a, s0);
    live_cur = rec.sub_live(s0);
    rec.begin_sub();
#pragma unroll
    for (int s = 0; s < SL; s += 4) rec.step4(tiles, coefs, ring, pb, s);
    rec.end_sub();
    n_rec += 1;
  }
  // flush assignment of this thread: l index within the chunk tid & 31, columns (tid >> 5) + 8 i
  const int f_lidx = threadIdx.x & 31;
  const int f_p = f_lidx >> 4, f_idx = f_lidx & 15;

  // the reduction of chunk c over the eight warps is deferred to the end of chunk c + 1, so that
  // nobody waits at a barrier: partial tiles and flags are double buffered, mbar[c & 1] counts the warps
  int f_l_prev = 0;
  double f_sc_prev = 0.0;
  auto reduce_chunk = [&](int c, int f_l, double f_sc) {
    mbar_wait(mbar + (c & 1), (c >> 1) & 1);
    const double *fbuf = flush + (c & 1) * K::FLUSH;
    int lv[NW], any = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      lv[w] = flags[(c & 1) * NW + w];
      any |= lv[w];
    }
    if (any && f_l <= lmax) {
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        const int col = (threadIdx.x >> 5) + 8 * i;
        int row, ri;
        if (SPIN == 0) {
          row = col >> 1;
          ri = col & 1;
        } else {
          row = 2 * (col >> 2) + ((col >> 1) & 1);
          ri = col & 1;
        }
        if (row < a.ncomp) {
          double sum = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) sum += lv[w] ? fbuf[w * (K::C * FLS) + col * FLS + f_lidx] : 0.0;
          atomicAdd(a.alm.p[row] + 2 * (cbase + f_l) + ri, sum * f_sc);
        }
      }
    }
  };

  for (int chk = chk0; chk < st.nchunk; ++chk) {
    const int rc = chk - chk0;  // flush buffers and barrier phases count from the first chunk that is processed
    const int lstart = st.l0 + chk * LC;
    // scale of this thread's flush outputs, fetched a chunk's worth of work ahead
    const int f_l = lstart + 2 * f_idx + (f_p ^ pb);
    double f_sc = ldg_pin(a.scale + cbase + min(f_l, lmax));
    const double f_fl = a.fl ? ldg_pin(a.fl + min(f_l, lmax)) : 1.0;
    double acc[2][2][NBLK][2];  // [parity][sub-chunk][n-block][2]
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int sb = 0; sb < 2; ++sb)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb) acc[p][sb][nb][0] = acc[p][sb][nb][1] = 0.0;
    bool chunk_live = false;
#pragma unroll
    for (int sb = 0; sb < 2; ++sb) {
      const int sidx = 2 * chk + sb;
      const double *tcur = tiles + (sidx & 1) * K::TILE;
      double *tnxt = tiles + ((sidx + 1) & 1) * K::TILE;
      const double *ccur = coefs + ((sidx + 1) & 1) * (SL * 2);
      // the recursion always runs one sub-chunk ahead (the one past the end is harmless: its
      // coefficients are zero and nothing reads it) so that the hot path below is branch free
      const bool prod = warp_alive;
      coef_wait();
      __syncwarp();  // coefficients of sub-chunk sidx + 1 have landed; everybody is done with those of sidx
      if (prod) stage_coef_async<SPIN>(coefs + (sidx & 1) * (SL * 2), a, cbase, st.l0 + (sidx + 2) * SL, lane);  // tile `sidx` and the coefficients of sub-chunk sidx + 1 are in place
      if (prod) {
        rec.maybe_start(a, sidx + 1);
        live_nxt = rec.sub_live(sidx + 1);
        rec.begin_sub();
        if (sidx + 1 < nsub) n_rec += 1;
      }
      pp_wait(pp_my);
      if (live_cur) {
        chunk_live = true;
        n_acc += 1;
        // DMMAs of sub-chunk sidx interleaved with the recursion of sub-chunk sidx + 1
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          // the four recursion groups go with k4 = 0, 2, 4, 6, so that the next tile is complete one k4 step
          // before the hand-over and the tail of the DFMA chain hides behind the last DMMAs
#ifndef HCU_EXP_NOREC
#ifdef HCU_REC_BURST
          // experiment: the whole recursion of the next sub-chunk as one burst ahead of the DMMAs
          if (k4 == 0) {
#pragma unroll
            for (int s = 0; s < SL; s += 4) rec.step4(tnxt, ccur, ring, pb, s);
          }
#else
          if (!(k4 & 1)) rec.step4(tnxt, ccur, ring, pb, (k4 >> 1) * 4);
#endif
#endif
          const double *ta = tcur + ((k4 & 1) ? a_off1 : a_off0) + 32 * k4;
          double af[NJ][2];
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int p = 0; p < 2; ++p) af[j][p] = ta[(2 * j + p) * 256];
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
              for (int nb = 0; nb < NBLK; ++nb)
#ifndef HCU_EXP_NODMMA
                dmma(acc[p][sb][nb][0], acc[p][sb][nb][1], af[j][p], bf[k4][j][p][nb]);
#else
                acc[p][sb][nb][0] += af[j][p] * 1e-300 + bf[k4][j][p][nb] * 1e-300;
#endif

        }
      }
      pp_pass(pp_other);
      if (!live_cur && prod) {
#pragma unroll
        for (int s = 0; s < SL; s += 4) rec.step4(tnxt, ccur, ring, pb, s);
      }
      if (prod) rec.end_sub();
      live_cur = prod ? live_nxt : false;
    }
#ifdef HCU_EXP_NOFLUSH
    {
      double ssum = f_sc * f_fl;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int sb = 0; sb < 2; ++sb)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) ssum += acc[p][sb][nb][0] + acc[p][sb][nb][1];
      f_sc_prev += ssum;
      continue;
    }
#endif
    // ---- first the deferred reduction of the previous chunk, then park this chunk's partial tile ----
    if (rc > 0) reduce_chunk(rc - 1, f_l_prev, f_sc_prev);
    f_l_prev = f_l;
    f_sc_prev = f_sc * f_fl;
    if (lane == 0) flags[(rc & 1) * NW + warp] = chunk_live ? 1 : 0;
    if (chunk_live) {
      double *o = flush + (rc & 1) * K::FLUSH + warp * (K::C * FLS);
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int sb = 0; sb < 2; ++sb)
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) {
            double *d = o + (nb * 8 + 2 * fa) * FLS + p * 16 + sb * 8 + fb;
            d[0] = acc[p][sb][nb][0];
            d[FLS] = acc[p][sb][nb][1];
          }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(mbar + (rc & 1));
  }
#ifdef HCU_EXP_NOFLUSH
  if (f_sc_prev == 1.2345e-300) atomicAdd(a.alm.p[0], f_sc_prev);
#else
  reduce_chunk(st.nchunk - 1 - chk0, f_l_prev, f_sc_prev);
#endif
  if (lane == 0 && a.work && n_rec > 0) {
    atomicAdd(a.work, n_rec * 32.0 * SL);
    atomicAdd(a.work + 1, n_acc * 32.0 * SL);
  }
}

// ---------------------------------------------------------------------------------
// synthesis
// ---------------------------------------------------------------------------------
template <int SPIN, int NBLK>
struct SCfg {
  static constexpr int NJ = SPIN == 0 ? 1 : 2;
  static constexpr int C = 8 * NBLK;
  static constexpr int NU = SPIN == 0 ? 4 * NBLK : 4;      // maps resp. fields per batch
  static constexpr int NLOAD = SPIN == 0 ? NU : 2 * NU;    // alm rows fetched per lane and chunk
  static constexpr int BSTR = C + 4;                       // 12 or 4 (mod 16): conflict free B fragments
  static constexpr int TILE = Rec<SPIN>::TILE;
  static constexpr int WARP = 2 * TILE + 2 * SL * 2 + LC * BSTR;
  static constexpr size_t SMEM_BYTES = sizeof(double) * (size_t)(NW * WARP);
  static_assert(SPIN == 0 || NBLK == 2, "spin 2 synthesis uses the (a2 | m2) two-block column layout");
};

template <int SPIN, int NBLK>
__global__ void __launch_bounds__(NT, 1) legendre_synthesis_kernel(LegArgs a) {
  using K = SCfg<SPIN, NBLK>;
  constexpr int NJ = K::NJ;
  extern __shared__ __align__(16) double smem_d[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Setup st;
  Rec<SPIN> rec;
  const bool active = setup_cta<SPIN>(a, st, rec, warp, lane);
  double *dst = st.out + ((i64)st.mi * st.nrp_b + st.row0) * a.ncomp * 4;
  if (!active) {  // every output row must be defined
    for (int i = threadIdx.x; i < st.nrows * a.ncomp * 4; i += NT) dst[i] = 0.0;
    return;
  }
  const int lmax = a.lmax, pb = st.pb;
  const i64 cbase = st.cbase;
  double *tiles = smem_d + warp * K::WARP;
  double *coefs = tiles + 2 * K::TILE;
  double *btile = coefs + 2 * SL * 2;  // [parity][16][BSTR]: s_l a_lm of the current chunk
  const bool warp_alive = __any_sync(0xffffffffu, rec.alive);
  const int fa = lane & 3;   // k inside a k4 block (A column = l / B row = l)
  const int fb = lane >> 2;  // A row (ring) / B column

  // spin 0: acc[parity][mb][nb]; spin 2: acc[j][mb][block], block 0 = (+2a) columns, block 1 = (-2a) columns
  double acc[2][4][NBLK][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) acc[i][mb][nb][0] = acc[i][mb][nb][1] = 0.0;

  // a_lm of one chunk: lane = l row.  Fetch into registers, convert and park in btile later.
  double2 av[K::NLOAD];
  double sc_next = 0.0;
  auto fetch_alm = [&](int lstart) {
    const int l = lstart + lane;
    sc_next = ldg_pin(a.scale + cbase + min(l, lmax));
#pragma unroll
    for (int i = 0; i < K::NLOAD; ++i) {
      av[i] = make_double2(0., 0.);
      if (warp_alive && l <= lmax && i < a.ncomp)
        av[i] = ldg_pin2(reinterpret_cast<const double2 *>(a.alm.p[i]) + cbase + l);
    }
  };
  auto park_alm = [&](int lstart) {
    const int l = lstart + lane;
    const double sc = (warp_alive && l <= lmax) ? sc_next : 0.0;
    // step `lane` of the chunk has parity pb ^ (lane & 1), l index lane >> 1 within the parity
    double *row = btile + ((pb ^ (lane & 1)) * 16 + (lane >> 1)) * K::BSTR;
    if (SPIN == 0) {
#pragma unroll
      for (int i = 0; i < K::NLOAD; ++i)
        *reinterpret_cast<double2 *>(row + 2 * i) = make_double2(av[i].x * sc, av[i].y * sc);
    } else {
#pragma unroll
      for (int f = 0; f < K::NU; ++f) {
        const double2 E = av[2 * f], B = av[2 * f + 1];
        // +2a = -(E + iB) -> block 0, -2a = -(E - iB) -> block 1
        *reinterpret_cast<double2 *>(row + 2 * f) = make_double2(-(E.x - B.y) * sc, -(E.y + B.x) * sc);
        *reinterpret_cast<double2 *>(row + 8 + 2 * f) = make_double2(-(E.x + B.y) * sc, -(E.y - B.x) * sc);
      }
    }
  };

  const int ring = lane;
  bool live_cur = false, live_nxt = false;
  const int pp_my = 1 + 2 * (warp & 3) + (warp >> 2), pp_other = 1 + 2 * (warp & 3) + (1 - (warp >> 2));
  if (warp >= 4) pp_pass(pp_other);  // the first turn belongs to warps 0..3
  // ---- prologue: the first sub-chunk and the a_lm of chunk chk0 (chunk 0 without a start-state table) ----
  const int chk0 = st.chk0, s0 = 2 * st.chk0;
  fetch_alm(st.l0 + chk0 * LC);
  park_coef(coefs, fetch_coef<SPIN>(a, cbase, st.l0 + s0 * SL, lane), st.l0 + s0 * SL, lmax, lane);
  stage_coef_async<SPIN>(coefs + SL * 2, a, cbase, st.l0 + (s0 + 1) * SL, lane);  // coefficients of the next sub-chunk
  __syncwarp();
  if (warp_alive) {
    rec.maybe_start(a, s0);
    live_cur = rec.sub_live(s0);
    rec.begin_sub();
#pragma unroll
    for (int s = 0; s < SL; s += 4) rec.step4(tiles, coefs, ring, pb, s);
    rec.end_sub();
  }
  park_alm(st.l0 + chk0 * LC);

  for (int chk = chk0; chk < st.nchunk; ++chk) {
    if (chk + 1 < st.nchunk) fetch_alm(st.l0 + (chk + 1) * LC);
#pragma unroll
    for (int sb = 0; sb < 2; ++sb) {
      const int sidx = 2 * chk + sb;
      const double *tcur = tiles + (sidx & 1) * K::TILE;
      double *tnxt = tiles + ((sidx + 1) & 1) * K::TILE;
      const double *ccur = coefs + ((sidx + 1) & 1) * (SL * 2);
      const bool prod = warp_alive;  // one sub-chunk ahead, also past the end (see the analysis kernel)
      coef_wait();
      __syncwarp();  // coefficients of sub-chunk sidx + 1 have landed; everybody is done with those of sidx
      if (prod) stage_coef_async<SPIN>(coefs + (sidx & 1) * (SL * 2), a, cbase, st.l0 + (sidx + 2) * SL, lane);  // tile `sidx`, btile and the coefficients of sub-chunk sidx + 1 are in place
      if (prod) {
        rec.maybe_start(a, sidx + 1);
        live_nxt = rec.sub_live(sidx + 1);
        rec.begin_sub();
      }
      pp_wait(pp_my);
      if (live_cur) {
        // k4 steps: (parity p, half h) -> rows p*16 + sb*8 + 4h + fa of btile, l index 4h + fa of the tile
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
          const int p = ph >> 1, h = ph & 1;
          const double *brow = btile + (p * 16 + sb * 8 + 4 * h + fa) * K::BSTR + fb;
          double bfr[NBLK];
#pragma unroll
          for (int nb = 0; nb < NBLK; ++nb) bfr[nb] = brow[nb * 8];
#pragma unroll
          for (int mb = 0; mb < 4; ++mb) {
            const int off = lam_off(mb * 8 + fb, 4 * h + fa);
            if (SPIN == 0) {
              const double av0 = tcur[p * 256 + off];
#pragma unroll
              for (int nb = 0; nb < NBLK; ++nb) dmma(acc[p][mb][nb][0], acc[p][mb][nb][1], av0, bfr[nb]);
            } else {
              // acc[0][.][0] = P_N = sum lam+ (+2a)        acc[0][.][1] = M_S = sum sg lam+ (-2a)
              // acc[1][.][0] = P_S = sum sg lam- (+2a)     acc[1][.][1] = M_N = sum lam- (-2a)
              // sg = (-1)^(l+m): the sign goes onto the A fragment (integer pipe)
              const double lp = tcur[p * 256 + off], lm = tcur[(2 + p) * 256 + off];
              const double lps = p ? neg_d(lp) : lp, lms = p ? neg_d(lm) : lm;
              dmma(acc[0][mb][0][0], acc[0][mb][0][1], lp, bfr[0]);
              dmma(acc[0][mb][1][0], acc[0][mb][1][1], lps, bfr[1]);
              dmma(acc[1][mb][0][0], acc[1][mb][0][1], lms, bfr[0]);
              dmma(acc[1][mb][1][0], acc[1][mb][1][1], lm, bfr[1]);
            }
          }
          rec.step4(tnxt, ccur, ring, pb, ph * 4);
        }
      }
      pp_pass(pp_other);
      if (!live_cur && prod) {
#pragma unroll
        for (int s = 0; s < SL; s += 4) rec.step4(tnxt, ccur, ring, pb, s);
      }
      if (prod) rec.end_sub();
      live_cur = prod ? live_nxt : false;
    }
    if (chk + 1 < st.nchunk) {
      __syncwarp();  // every lane is done reading btile
      park_alm(st.l0 + (chk + 1) * LC);
    }
  }

  // ---- results straight from the accumulator fragments: lane holds ring mb*8+fb, unit fa (+4 nb) ----
  if (!warp_alive) {
    for (int i = lane; i < 32 * a.ncomp * 4; i += 32) {
      const int r = warp * 32 + i / (a.ncomp * 4);
      if (r < st.nrows) dst[(i64)warp * 32 * a.ncomp * 4 + i] = 0.0;
    }
    return;
  }
#pragma unroll
  for (int mb = 0; mb < 4; ++mb) {
    const int r = warp * 32 + mb * 8 + fb;
    if (r >= st.nrows) continue;
    if (SPIN == 0) {
      // parity 0 = (l + m) even: north = T0 + T1, south = T0 - T1
#pragma unroll
      for (int nb = 0; nb < NBLK; ++nb) {
        const int c = nb * 4 + fa;
        if (c < a.ncomp) {
          const double t0r = acc[0][mb][nb][0], t0i = acc[0][mb][nb][1];
          const double t1r = acc[1][mb][nb][0], t1i = acc[1][mb][nb][1];
          *reinterpret_cast<double4 *>(dst + ((i64)r * a.ncomp + c) * 4) =
              make_double4(t0r + t1r, t0i + t1i, t0r - t1r, t0i - t1i);
        }
      }
    } else {
      const int f = fa;
      if (2 * f < a.ncomp) {
        const double PrN = acc[0][mb][0][0], PiN = acc[0][mb][0][1];
        const double MrS = acc[0][mb][1][0], MiS = acc[0][mb][1][1];
        const double PrS = acc[1][mb][0][0], PiS = acc[1][mb][0][1];
        const double MrN = acc[1][mb][1][0], MiN = acc[1][mb][1][1];
        double *d = dst + ((i64)r * a.ncomp + 2 * f) * 4;
        // Q = (P + M)/2 ; U = (P - M)/(2i)
        *reinterpret_cast<double4 *>(d) =
            make_double4(0.5 * (PrN + MrN), 0.5 * (PiN + MiN), 0.5 * (PrS + MrS), 0.5 * (PiS + MiS));
        *reinterpret_cast<double4 *>(d + 4) =
            make_double4(0.5 * (PiN - MiN), -0.5 * (PrN - MrN), 0.5 * (PiS - MiS), -0.5 * (PrS - MrS));
      }
    }
  }
}

// recursion tables, one thread per m (sequential in l because of the running scale s_l):
//   lambda_{l+1} = (alpha_l x +- alpha_l beta_l) lambda_l - gamma_l lambda_{l-1},  l >= l0 = max(m, spin)
//   lambda_l = s_l q_l with s_{l0} = s_{l0+1} = 1, s_{l+1} = gamma_l s_{l-1}, which turns it into
//   q_{l+1} = (A_l x +- B_l) q_l - q_{l-1},  A_l = alpha_l s_l / s_{l+1},  B_l = alpha_l beta_l s_l / s_{l+1}
// tab: spin 0 double A_l; spin 2 double2 (A_l, B_l); scale: s_l.  Index cbase(m) + l.
__global__ void coef_kernel(int lmax, int spin, double *tab, double *scale) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > lmax) return;
  const int s = spin;
  const int l0 = m > s ? m : s;
  const i64 base = (i64)m * (2 * lmax + 1 - m) / 2;
  double s_prev = 1.0, s_cur = 1.0;  // s_{l-1}, s_l
  for (int l = m; l <= lmax; ++l) {
    double A = 0, B = 0, sc = 0;
    if (l >= l0) {
      sc = s_cur;
      if (l < lmax) {
        const double dl = l, l1 = dl + 1.0, dm = m, ds = s;
        const double den = sqrt((l1 * l1 - dm * dm) * (l1 * l1 - ds * ds));
        const double al = sqrt((2 * dl + 3) / (2 * dl + 1)) * l1 * (2 * dl + 1) / den;
        const double be = (l > 0) ? (ds * dm) / (dl * l1) : 0.0;
        const double ga = (l > l0) ? sqrt((2 * dl + 3) / (2 * dl - 1)) * l1 / dl *
                                         sqrt((dl * dl - dm * dm) * (dl * dl - ds * ds)) / den
                                   : 0.0;
        const double s_next = (l > l0) ? ga * s_prev : 1.0;
        A = al * s_cur / s_next;
        B = al * be * s_cur / s_next;
        s_prev = s_cur;
        s_cur = s_next;
      }
    }
    if (s == 0) {
      tab[base + l] = A;
    } else {
      reinterpret_cast<double2 *>(tab)[base + l] = make_double2(A, B);
    }
    scale[base + l] = sc;
  }
}

template <int SPIN, int NBLK>
int launch_analysis(hcu_ctx *ctx, const LegArgs &a) {
  using K = ACfg<SPIN, NBLK>;
  const i64 nblocks = (i64)a.grp_start[a.nblk] * a.nm;
  if (nblocks <= 0) return HCU_OK;
  HCU_CUDA(cudaFuncSetAttribute(legendre_analysis_kernel<SPIN, NBLK>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
  legendre_analysis_kernel<SPIN, NBLK><<<(unsigned)nblocks, NT, K::SMEM_BYTES, ctx->stream>>>(a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

template <int SPIN, int NBLK>
int launch_synthesis(hcu_ctx *ctx, const LegArgs &a) {
  using K = SCfg<SPIN, NBLK>;
  const i64 nblocks = (i64)a.grp_start[a.nblk] * a.nm;
  if (nblocks <= 0) return HCU_OK;
  HCU_CUDA(cudaFuncSetAttribute(legendre_synthesis_kernel<SPIN, NBLK>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
  legendre_synthesis_kernel<SPIN, NBLK><<<(unsigned)nblocks, NT, K::SMEM_BYTES, ctx->stream>>>(a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

void fill_args(LegArgs &a, hcu_geom *g, hcu_coef *c, int lmax, int ncomp,
               const int32_t *mlist_dev, int nm, int nblk, const i64 *rp_bounds) {
  a.lmax = lmax;
  a.nm = nm;
  a.ncomp = ncomp;
  a.mlist = mlist_dev;
  a.phase = nullptr;
  a.nblk = nblk;
  a.use_blk_out = 0;
  a.phase_out = nullptr;
  a.grp_start[0] = 0;
  for (int b = 0; b < nblk; ++b) {
    a.blk_rp[b] = rp_bounds[b];
    a.grp_start[b + 1] = a.grp_start[b] + (int)((rp_bounds[b + 1] - rp_bounds[b] + R - 1) / R);
  }
  a.blk_rp[nblk] = rp_bounds[nblk];
  a.cth = g->cth;
  a.sth = g->sth;
  a.ch = g->ch;
  a.sh = g->sh;
  a.coef = c->tab;
  a.scale = c->scale;
  a.cmtab = c->cm;
  a.fl = nullptr;
  a.phase_out = nullptr;
  a.work = nullptr;
  a.st_sub = nullptr;
  a.st_state = nullptr;
  a.st_nrp = g->nrp;
}

}  // namespace

// second-generation kernels (k_legendre2.cu)
int hcu_legendre2_analysis(hcu_ctx *ctx, void *args, const i64 *rp_bounds, int spin, int ncomp, int nw);
int hcu_legendre2_synthesis(hcu_ctx *ctx, void *args, const i64 *rp_bounds, int spin, int ncomp, int nw);

// Which kernel generation runs a batch.  Measured on B200 (profiles/r02_legendre_batch_table.txt, nside 2048, two
// analysis passes + one synthesis pass): a launch costs  a + b * columns  with b at the FP64 pipe's peak for BOTH
// generations and a latency-bound fixed part a (recursion, tile hand-over) that does not depend on the columns:
//     spin 0   <= 4 maps: gen 1;   5..8 maps: gen 2 (75.8 + 38.6 ms against 93.5 + 45.1);   9..12 maps: gen 1 (3 n-blocks)
//     spin 2   gen 1 (4 fields 159 + 76.7 ms against 160.5 + 84.5); <= 2 fields: gen 2 synthesis (one n-block)
// HCU_LEGENDRE_GEN = 1 | 2 forces one generation (A/B timing, tests); HCU_LEGENDRE_NW = 12 | 16 are the warps per CTA
// of the second-generation analysis kernel (8: experimental 16-component analysis-only batches).
static int legendre_gen_env() {
  static int gen = -1;
  if (gen < 0) {
    const char *e = getenv("HCU_LEGENDRE_GEN");
    gen = (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : 0;
  }
  return gen;
}
static int legendre_nw() {
  static int nw = -1;
  if (nw < 0) {
    const char *e = getenv("HCU_LEGENDRE_NW");
    nw = (e && atoi(e) == 16) ? 16 : (e && atoi(e) == 8) ? 8 : 12;
  }
  return nw;
}
static int legendre_gen(int spin, int ncomp, bool synthesis) {
  const int forced = legendre_gen_env();
  if (forced == 1) return 1;
  if (forced == 2) return ncomp <= 8 || legendre_nw() == 8 ? 2 : 1;
  if (spin == 0) return (ncomp >= 5 && ncomp <= 8) ? 2 : 1;
  // spin 2: the second-generation synthesis has a one-n-block variant for <= 2 fields (half the DMMAs of the
  // first generation's fixed (+2a | -2a) two-block layout)
  return (synthesis && ncomp <= 4) ? 2 : 1;
}

// components one Legendre launch takes: 12 spin-0 maps or 4 spin-2 fields (Q, U rows);
// forced second generation: 8 (HCU_LEGENDRE_NW=8: experimental 16-component analysis-only batches)
int hcu_legendre_batch(int spin) {
  if (legendre_gen_env() == 2) return legendre_nw() == 8 ? 16 : 8;
  return spin == 0 ? 12 : 8;
}

int hcu_build_coef(hcu_ctx *ctx, hcu_coef *c) {
  const i64 nalm = (i64)(c->lmax + 1) * (c->lmax + 2) / 2;
  const size_t per = (c->spin == 0) ? 1 : 2;
  HCU_CUDA(cudaMalloc(&c->tab, sizeof(double) * per * nalm));
  HCU_CUDA(cudaMalloc(&c->scale, sizeof(double) * nalm));
  coef_kernel<<<(c->lmax + 64) / 64, 64, 0, ctx->stream>>>(c->lmax, c->spin, c->tab, c->scale);
  HCU_LAUNCH_CHECK(ctx);
  // start-value normalisation in long double on the host:
  //   cm[2m] = c_m with lambda_mm = (-1)^m c_m sin^m(theta);  cm[2m+1] = c_m sqrt(m(m-1)/((m+1)(m+2)))
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

int hcu_build_start(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, hcu_start *t) {
  const size_t n = (size_t)(c->lmax + 1) * g->nrp * (c->spin == 0 ? 1 : 2);
  if (cudaMalloc(&t->sub, sizeof(int) * n) != cudaSuccess) {
    t->sub = nullptr;
    return HCU_ERR_NOMEM;
  }
  if (cudaMalloc(&t->state, sizeof(double2) * n) != cudaSuccess) {
    cudaFree(t->sub);
    t->sub = nullptr;
    t->state = nullptr;
    return HCU_ERR_NOMEM;
  }
  dim3 grid((g->nrp + 127) / 128, c->lmax + 1);
  if (c->spin == 0)
    start_table_kernel<0><<<grid, 128, 0, ctx->stream>>>(c->lmax, g->nrp, g->cth, g->sth, g->ch, g->sh, c->tab, c->cm, t->sub, t->state);
  else
    start_table_kernel<2><<<grid, 128, 0, ctx->stream>>>(c->lmax, g->nrp, g->cth, g->sth, g->ch, g->sh, c->tab, c->cm, t->sub, t->state);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

// alm[c][l, m] += sum over the ring pairs of all blocks of lambda_lm(theta) x phase, for m in mlist
int hcu_legendre_analysis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                          int spin, int ncomp, const double *phase,
                          const int32_t *mlist_dev, int nm, int nblk, const i64 *rp_bounds,
                          const double *fl_dev, const hcu_ptrs &alm) {
  HCU_ARG(ncomp >= 1 && ncomp <= hcu_legendre_batch(spin), "legendre batch size");
  HCU_ARG(nblk >= 1 && nblk <= HCU_MAX_BLOCKS, "1 <= ring-pair blocks <= 16");
  LegArgs a;
  fill_args(a, g, c, lmax, ncomp, mlist_dev, nm, nblk, rp_bounds);
  a.phase = phase;
  a.fl = fl_dev;
  a.alm = alm;
  a.work = ctx->work_counters;
  const bool gen2 = legendre_gen(spin, ncomp, false) == 2;
  hcu_start *tab = nullptr;
  if (!gen2) HCU_CHECK(hcu_get_start(ctx, g, c, &tab));  // (the second-generation kernels walk the dead zone themselves)
  if (tab) {
    a.st_sub = tab->sub;
    a.st_state = tab->state;
  }
  if (gen2) return hcu_legendre2_analysis(ctx, &a, rp_bounds, spin, ncomp, legendre_nw());
  // 8 output columns per n-block: 4 spin-0 maps, or 2 spin-2 fields (4 Q/U rows)
  const int ncolblk = (ncomp + 3) / 4;
  if (spin == 0) {
    switch (ncolblk) {
      case 1: return launch_analysis<0, 1>(ctx, a);
      case 2: return launch_analysis<0, 2>(ctx, a);
      default: return launch_analysis<0, 3>(ctx, a);
    }
  }
  return ncolblk == 1 ? launch_analysis<2, 1>(ctx, a) : launch_analysis<2, 2>(ctx, a);
}

// phase (blocked layout, see LegArgs) = (reN, imN, reS, imS) of sum_l a_lm lambda_lm(theta_rp)
int hcu_legendre_synthesis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                           int spin, int ncomp, const hcu_ptrs &alm,
                           const int32_t *mlist_dev, int nm, int nblk, const i64 *rp_bounds,
                           double *phase, double *const *block_out) {
  HCU_ARG(ncomp >= 1 && ncomp <= hcu_legendre_batch(spin), "synthesis batch size");
  HCU_ARG(nblk >= 1 && nblk <= HCU_MAX_BLOCKS, "1 <= ring-pair blocks <= 16");
  LegArgs a;
  fill_args(a, g, c, lmax, ncomp, mlist_dev, nm, nblk, rp_bounds);
  a.alm = alm;
  a.phase_out = phase;
  if (block_out) {
    a.use_blk_out = 1;
    for (int b = 0; b < nblk; ++b) a.blk_out[b] = block_out[b];
  }
  if (legendre_gen(spin, ncomp, true) == 2) return hcu_legendre2_synthesis(ctx, &a, rp_bounds, spin, ncomp, legendre_nw());
  hcu_start *tab = nullptr;
  HCU_CHECK(hcu_get_start(ctx, g, c, &tab));
  if (tab) {
    a.st_sub = tab->sub;
    a.st_state = tab->state;
  }
  if (spin == 0) {
    switch ((ncomp + 3) / 4) {
      case 1: return launch_synthesis<0, 1>(ctx, a);
      case 2: return launch_synthesis<0, 2>(ctx, a);
      default: return launch_synthesis<0, 3>(ctx, a);
    }
  }
  return launch_synthesis<2, 2>(ctx, a);
}
