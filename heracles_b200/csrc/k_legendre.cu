// k_legendre.cu -- FP64 Legendre stage of the spherical-harmonic SYNTHESIS
// (alm -> ring Fourier coefficients) on HEALPix ring pairs, spin 0 and spin 2,
// batched over up to 12 components; plus the recursion coefficient tables
// shared with the analysis kernel (k_legendre_ana.cu).
//
// Replaces the libsharp/ducc Legendre loop behind hp.alm2map, which healpy's
// map2alm runs inside its default iter=3 refinement (heracles/healpy.py:183-189).
// FP64 FMA bound.
//
// One lane = one ring pair: it advances its own lambda_lm(theta) recursion in l
// and accumulates  b_m(theta) = sum_l a_lm lambda_lm(theta)  for every component
// of the batch in registers, so there is no cross-lane reduction.  The a_lm and
// the recursion coefficients of a chunk of l are staged once per CTA in shared
// memory and read as warp-wide broadcasts.  North and south rings share the
// recursion through lambda_lm(pi - theta) = (-1)^(l+m) lambda_lm(theta): spin 0
// keeps separate even / odd (l+m) accumulators and forms N = E + O, S = E - O at
// the end, which halves the FMAs.
#include "legendre_common.cuh"

namespace {

// phase out: [(m * nrp + rp) * ncomp + c] * 4 = (reN, imN, reS, imS)
template <int SPIN, int NB>
__global__ void __launch_bounds__(128) legendre_synthesis_kernel(LegArgs a, double *phase_out) {
  constexpr int LCH = 32;                     // l per staged chunk (even)
  constexpr int CW = SPIN == 0 ? 2 : 4;       // doubles per coefficient entry
  __shared__ __align__(16) double2 s_alm[LCH * NB];
  __shared__ __align__(16) double s_coef[LCH * CW];
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
  // spin 0: acc[q][c] with q = parity slot (even/odd step inside a chunk)
  // spin 2: acc[0] = sum lam+ 2a, acc[1] = sum lam- -2a (north); acc[2], acc[3] the southern sums
  constexpr int NACC = SPIN == 0 ? 2 : 4;
  constexpr int NCOL = SPIN == 0 ? NB : (NB + 1) / 2;
  double acc[NACC][NCOL][2];
#pragma unroll
  for (int q = 0; q < NACC; ++q)
#pragma unroll
    for (int c = 0; c < NCOL; ++c) acc[q][c][0] = acc[q][c][1] = 0.0;
  const int pb = (l0 + m) & 1;

  if (l0 <= lmax) {
    const bool alive = valid && !ring_is_dead(lmax, m, SPIN, x, sth);
    LamState sp, sm;
    sp.prev = sp.cur = 0; sp.e = 0;
    sm.prev = sm.cur = 0; sm.e = 0;
    if (alive) lam_start<SPIN>(m, a.cmtab, sth, chh, shh, sp, sm);
    const i64 cbase = alm_index(lmax, 0, m);
    for (int lc = l0; lc <= lmax; lc += LCH) {
      __syncthreads();
      // ---- stage a_lm and coefficients of l = lc .. lc + LCH - 1 ----
      if (SPIN == 0) {
        for (int idx = threadIdx.x; idx < LCH * NB; idx += 128) {
          const int li = idx / NB, c = idx - li * NB;
          const int l = lc + li;
          double2 v = make_double2(0., 0.);
          if (l <= lmax && c < a.ncomp) v = reinterpret_cast<const double2 *>(a.alm.p[c])[cbase + l];
          s_alm[idx] = v;
        }
      } else {
        for (int idx = threadIdx.x; idx < LCH * (NB / 2); idx += 128) {
          const int li = idx / (NB / 2), f = idx - li * (NB / 2);
          const int l = lc + li;
          double2 E = make_double2(0., 0.), B = make_double2(0., 0.);
          if (l <= lmax && 2 * f < a.ncomp) {
            E = reinterpret_cast<const double2 *>(a.alm.p[2 * f])[cbase + l];
            B = reinterpret_cast<const double2 *>(a.alm.p[2 * f + 1])[cbase + l];
          }
          // 2a = -(E + iB), -2a = -(E - iB)
          s_alm[li * NB + 2 * f] = make_double2(-(E.x - B.y), -(E.y + B.x));
          s_alm[li * NB + 2 * f + 1] = make_double2(-(E.x + B.y), -(E.y - B.x));
        }
      }
      for (int i = threadIdx.x; i < LCH; i += 128) {
        const int l = lc + i;
        if (SPIN == 0) {
          double2 cf = make_double2(0., 0.);
          if (l < lmax) cf = __ldg(reinterpret_cast<const double2 *>(a.coef) + cbase + l);
          reinterpret_cast<double2 *>(s_coef)[i] = cf;
        } else {
          double4 cf = make_double4(0., 0., 0., 0.);
          if (l < lmax) cf = ldg_d4(reinterpret_cast<const double4 *>(a.coef) + cbase + l);
          reinterpret_cast<double4 *>(s_coef)[i] = cf;
        }
      }
      __syncthreads();
      if (!__any_sync(0xffffffffu, alive)) continue;
      const bool scaled = __any_sync(0xffffffffu, sp.e < 0 || (SPIN != 0 && sm.e < 0));
      const bool any_live = __any_sync(0xffffffffu, alive && (sp.e == 0 || (SPIN != 0 && sm.e == 0)));
      if (!any_live) {
        // nothing representable yet in this warp: advance the recursions only
        for (int li = 0; li < LCH; ++li) {
          if (SPIN == 0) {
            const double2 cf = reinterpret_cast<const double2 *>(s_coef)[li];
            lam_advance(sp, cf.x * x, cf.y);
          } else {
            const double4 cf = reinterpret_cast<const double4 *>(s_coef)[li];
            lam_advance(sp, fma(cf.x, x, cf.y), cf.z);
            lam_advance(sm, fma(cf.x, x, -cf.y), cf.z);
          }
        }
        continue;
      }
#pragma unroll 2
      for (int li = 0; li < LCH; li += 2) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double2 *al = s_alm + (li + q) * NB;
          if (SPIN == 0) {
            const double2 cf = reinterpret_cast<const double2 *>(s_coef)[li + q];
            const double lam = (!scaled || sp.e == 0) ? sp.cur : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
              const double2 v = al[c];
              acc[q][c][0] = fma(lam, v.x, acc[q][c][0]);
              acc[q][c][1] = fma(lam, v.y, acc[q][c][1]);
            }
            if (scaled) {
              lam_advance(sp, cf.x * x, cf.y);
            } else {
              const double nw = fma(cf.x * x, sp.cur, -(cf.y * sp.prev));
              sp.prev = sp.cur;
              sp.cur = nw;
            }
          } else {
            const double4 cf = reinterpret_cast<const double4 *>(s_coef)[li + q];
            const double lp = (!scaled || sp.e == 0) ? sp.cur : 0.0;
            const double lm = (!scaled || sm.e == 0) ? sm.cur : 0.0;
            // (-1)^(l+m): slot q = 0 has parity pb
            const double sg = ((q + pb) & 1) ? -1.0 : 1.0;
            const double slp = sg * lp, slm = sg * lm;
#pragma unroll
            for (int f = 0; f < NB / 2; ++f) {
              const double2 a2 = al[2 * f], m2 = al[2 * f + 1];
              acc[0][f][0] = fma(lp, a2.x, acc[0][f][0]);
              acc[0][f][1] = fma(lp, a2.y, acc[0][f][1]);
              acc[1][f][0] = fma(lm, m2.x, acc[1][f][0]);
              acc[1][f][1] = fma(lm, m2.y, acc[1][f][1]);
              // lambda^{+2}(pi - theta) = sg lambda^{-2}(theta) and vice versa
              acc[2][f][0] = fma(slm, a2.x, acc[2][f][0]);
              acc[2][f][1] = fma(slm, a2.y, acc[2][f][1]);
              acc[3][f][0] = fma(slp, m2.x, acc[3][f][0]);
              acc[3][f][1] = fma(slp, m2.y, acc[3][f][1]);
            }
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
      }
    }
  }
  if (!valid) return;
  double *dst = phase_out + ((i64)m * a.nrp_local + rp) * a.ncomp * 4;
  if (SPIN == 0) {
    // slot q = 0 holds the terms with (l+m) parity pb
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      if (c < a.ncomp) {
        const double er = pb ? acc[1][c][0] : acc[0][c][0], ei = pb ? acc[1][c][1] : acc[0][c][1];
        const double orr = pb ? acc[0][c][0] : acc[1][c][0], oi = pb ? acc[0][c][1] : acc[1][c][1];
        *reinterpret_cast<double4 *>(dst + c * 4) = make_double4(er + orr, ei + oi, er - orr, ei - oi);
      }
    }
  } else {
#pragma unroll
    for (int f = 0; f < NB / 2; ++f) {
      if (2 * f < a.ncomp) {
        const double PrN = acc[0][f][0], PiN = acc[0][f][1], MrN = acc[1][f][0], MiN = acc[1][f][1];
        const double PrS = acc[2][f][0], PiS = acc[2][f][1], MrS = acc[3][f][0], MiS = acc[3][f][1];
        // Q = (P + M)/2 ; U = (P - M)/(2i)
        *reinterpret_cast<double4 *>(dst + (2 * f) * 4) =
            make_double4(0.5 * (PrN + MrN), 0.5 * (PiN + MiN), 0.5 * (PrS + MrS), 0.5 * (PiS + MiS));
        *reinterpret_cast<double4 *>(dst + (2 * f + 1) * 4) =
            make_double4(0.5 * (PiN - MiN), -0.5 * (PrN - MrN), 0.5 * (PiS - MiS), -0.5 * (PrS - MrS));
      }
    }
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

int hcu_legendre_synthesis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                           int spin, int ncomp, const hcu_ptrs &alm, double *phase) {
  HCU_ARG(ncomp >= 1 && ncomp <= HCU_MAX_BATCH, "synthesis batch size");
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
    return launch_synthesis<0, 12>(ctx, a, phase);
  } else {
    if (ncomp <= 2) return launch_synthesis<2, 2>(ctx, a, phase);
    if (ncomp <= 4) return launch_synthesis<2, 4>(ctx, a, phase);
    if (ncomp <= 6) return launch_synthesis<2, 6>(ctx, a, phase);
    if (ncomp <= 8) return launch_synthesis<2, 8>(ctx, a, phase);
    if (ncomp <= 10) return launch_synthesis<2, 10>(ctx, a, phase);
    return launch_synthesis<2, 12>(ctx, a, phase);
  }
}
