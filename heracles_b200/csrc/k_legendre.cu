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
#include "legendre_common.cuh"

namespace {

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
          v = a.alm.p[c][2 * (cbase + l) + ri];
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


int hcu_legendre_synthesis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                           int spin, int ncomp, const hcu_ptrs &alm, double *phase) {
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
  a.alm = alm;
  a.work = nullptr;
  if (spin == 0) {
    if (ncomp <= 1) return launch_synthesis<0, 1>(ctx, a, phase);
    if (ncomp <= 2) return launch_synthesis<0, 2>(ctx, a, phase);
    if (ncomp <= 4) return launch_synthesis<0, 4>(ctx, a, phase);
    if (ncomp <= 6) return launch_synthesis<0, 6>(ctx, a, phase);
    if (ncomp <= 8) return launch_synthesis<0, 8>(ctx, a, phase);
    if (ncomp <= 10) return launch_synthesis<0, 10>(ctx, a, phase);
    HCU_ARG(ncomp <= 12, "synthesis batch size");
    return launch_synthesis<0, 12>(ctx, a, phase);
  } else {
    if (ncomp <= 2) return launch_synthesis<2, 2>(ctx, a, phase);
    if (ncomp <= 4) return launch_synthesis<2, 4>(ctx, a, phase);
    if (ncomp <= 6) return launch_synthesis<2, 6>(ctx, a, phase);
    if (ncomp <= 8) return launch_synthesis<2, 8>(ctx, a, phase);
    if (ncomp <= 10) return launch_synthesis<2, 10>(ctx, a, phase);
    HCU_ARG(ncomp <= 12, "synthesis batch size");
    return launch_synthesis<2, 12>(ctx, a, phase);
  }
}
