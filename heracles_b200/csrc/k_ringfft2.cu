// k_ringfft2.cu -- second-generation ring FFT stage: ONE fused kernel per ring-pair class and direction.
//
// Replaces the FFT half of hp.map2alm / hp.alm2map (heracles/healpy.py:183-189) for every ring pair whose
// sub-transform fits one CTA (equatorial belt at nside >= 16, polar-cap rings i <= 4096); k_ringfft.cu keeps
// the first generation for the rest (nside < 16 belts, the two-half Bluestein of nside 8192) and as HCU_RINGFFT_GEN=1.
//
// A ring pair (north ring + southern mirror, n = 4 L pixels each; L = ring number in the caps, nside in the belt)
// is ONE complex sequence z = N + i S.  A decimation-in-frequency radix-4 step at load time,
//     y_s[j] = w_n^(s j) sum_t z[j + t L] (-i)^(s t),   Z[4 k + s] = DFT_L(y_s)[k],   s = 0..3,
// leaves four length-L transforms: a direct power-of-two FFT in the belt, a Bluestein convolution (power-of-two
// length M >= 2 L - 1) in the caps.  The FFTs live in shared memory and work in REGISTER radix-16 butterflies
// (four fused radix-2 stages, constant twiddles inside, one twiddle table read per butterfly), the last forward
// pass, the Bluestein filter multiply and the first inverse pass fused in registers: a 8192-point convolution is
// 7 shared-memory round trips instead of the 13 of the first generation's radix-4 passes.  The sub-spectra go
// through a per-CTA scratch that never leaves the L2 (the kernel is persistent: grid = resident CTAs), then the CTA
// untangles N / S, folds the aliases, applies exp(-i m phi0) and the quadrature weight and writes the phase rows
// of ITS (ring pair, component) -- no intermediate array in HBM, no cuFFT, no sincospi in the loops (chirp,
// radix-4 twiddle and ring-phase factors come from tables).  The inverse mirrors this (decimation in time, so that
// the pixels of a ring are written as four contiguous runs).
//
// HBM-bound in the belt (algorithmic bytes = 8 npix + 32 nrp (lmax+1) per component), FP64 / shared-memory bound
// in the caps (the Bluestein convolution is ~6x the flops of a direct transform).
#include <stdlib.h>

#include "hcu_common.cuh"

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 conj2(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }  // a * (+i)
__device__ __forceinline__ double2 expmipi(double x) {  // exp(-i pi x)
  double s, c;
  sincospi(x, &s, &c);
  return make_double2(c, -s);
}

// shared-memory index of element e: one 16-byte pad per 16 elements, so that both the unit-stride passes and the
// 16-contiguous-elements-per-thread middle pass are free of bank conflicts
__device__ __forceinline__ int pad(int e) { return e + (e >> 4); }

// d * exp(-+ 2 pi i e16 / 16); e16 is a compile-time constant after unrolling
template <bool CONJ>
__device__ __forceinline__ double2 mulw16(double2 d, const int e16) {
  const double h = 0.70710678118654752440;
  const double c1 = 0.92387953251128673848, s1 = 0.38268343236508977173;
  if (e16 == 0) return d;
  if (e16 == 4) return CONJ ? mul_pi(d) : mul_mi(d);
  if (e16 == 2) return CONJ ? make_double2((d.x - d.y) * h, (d.x + d.y) * h) : make_double2((d.x + d.y) * h, (d.y - d.x) * h);
  if (e16 == 6) return CONJ ? make_double2((-d.x - d.y) * h, (d.x - d.y) * h) : make_double2((d.y - d.x) * h, (-d.x - d.y) * h);
  const double c = (e16 == 1) ? c1 : (e16 == 3) ? s1 : (e16 == 5) ? -s1 : -c1;
  const double s = (e16 == 1 || e16 == 7) ? s1 : c1;
  return CONJ ? make_double2(d.x * c - d.y * s, d.y * c + d.x * s) : make_double2(d.x * c + d.y * s, d.y * c - d.x * s);
}

// log2(R) in-place radix-2 decimation-in-frequency stages on R registers, twiddles w_R only
template <int R>
__device__ __forceinline__ void dif_core(double2 (&v)[R]) {
#pragma unroll
  for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((r & half) == 0) {
        const int e16 = (r & (half - 1)) * 8 / half;
        const double2 u = v[r], w = v[r + half];
        v[r] = cadd(u, w);
        v[r + half] = mulw16<false>(csub(u, w), e16);
      }
    }
  }
}
// its exact inverse up to the factor R
template <int R>
__device__ __forceinline__ void dit_core(double2 (&v)[R]) {
#pragma unroll
  for (int half = 1; half < R; half <<= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((r & half) == 0) {
        const int e16 = (r & (half - 1)) * 8 / half;
        const double2 u = v[r], t = mulw16<true>(v[r + half], e16);
        v[r] = cadd(u, t);
        v[r + half] = csub(u, t);
      }
    }
  }
}

template <int R>
__device__ __forceinline__ constexpr int brev_r(int x) {
  int y = 0;
  for (int b = 1, c = R >> 1; b < R; b <<= 1, c >>= 1)
    if (x & b) y |= c;
  return y;
}

// position r of a butterfly at offset k inside blocks of size S carries w_S^(k brev(r)) after the fused stages (the
// factor a radix-2 stage gives the upper element is common to everything that is combined with it later).
// T holds log2(R) arrays of S / R entries, T[t][k] = w_S^(k 2^t).
template <int R, bool CONJ>
__device__ __forceinline__ void twiddle_mul(double2 (&v)[R], const double2 *T, int str, int k) {
  double2 pw[R];  // pw[e] = w^e; only e = 1..R-1 used
#pragma unroll
  for (int e = 1; e < R; ++e) {
    if ((e & (e - 1)) == 0) {
      int t = 0;
#pragma unroll
      for (int b = 1; b < R; b <<= 1)
        if (b < e) ++t;
      pw[e] = T[t * str + k];
    } else {
      pw[e] = cmul(pw[e & (e - 1)], pw[e & -e]);
    }
  }
#pragma unroll
  for (int r = 1; r < R; ++r) v[r] = CONJ ? cmulc(v[r], pw[brev_r<R>(r)]) : cmul(v[r], pw[brev_r<R>(r)]);
}

template <int R>
__device__ __forceinline__ void dif_pass(double2 *a, const double2 *T, int M, int S) {
  const int str = S / R, pstr = str + (str >> 4);
  for (int b = threadIdx.x; b < M / R; b += blockDim.x) {
    const int k = b & (str - 1);
    const int base = (b - k) * R + k;
    double2 *ap = a + pad(base);  // str is a multiple of 16: element r sits r (str + str / 16) slots further
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ap[r * pstr];
    dif_core<R>(v);
    twiddle_mul<R, false>(v, T, str, k);
#pragma unroll
    for (int r = 0; r < R; ++r) ap[r * pstr] = v[r];
  }
}
template <int R>
__device__ __forceinline__ void dit_pass(double2 *a, const double2 *T, int M, int S) {
  const int str = S / R, pstr = str + (str >> 4);
  for (int b = threadIdx.x; b < M / R; b += blockDim.x) {
    const int k = b & (str - 1);
    const int base = (b - k) * R + k;
    double2 *ap = a + pad(base);  // str is a multiple of 16: element r sits r (str + str / 16) slots further
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = ap[r * pstr];
    twiddle_mul<R, true>(v, T, str, k);
    dit_core<R>(v);
#pragma unroll
    for (int r = 0; r < R; ++r) ap[r * pstr] = v[r];
  }
}

// the innermost four stages on 16 contiguous elements per thread (all twiddles constant):
// FILTER: forward stages, multiply by the Bluestein filter spectrum, inverse stages -- one round trip
template <bool FILTER>
__device__ __forceinline__ void mid16(double2 *a, const double2 *B, int M) {
  for (int b = threadIdx.x; b < M / 16; b += blockDim.x) {
    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = a[17 * b + r];
    dif_core<16>(v);
    if (FILTER) {
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = cmul(v[r], __ldg(B + 16 * b + r));
      dit_core<16>(v);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) a[17 * b + r] = v[r];
  }
}

// pass plan of a 2^p-point transform, p >= 4: radix-16 passes at S = M, M/16, ... >= 256, one radix-2/4/8 pass at
// S = 32/64/128 when p mod 4 != 0, then the middle pass at S = 16.  Twiddle tables in that order.
template <bool CONVOLVE>
__device__ __forceinline__ void fft_smem(double2 *a, const double2 *T, const double2 *B, int M) {
  int S = M;
  const double2 *t = T;
  while (S >= 256) {
    dif_pass<16>(a, t, M, S);
    __syncthreads();
    t += 4 * (S >> 4);
    S >>= 4;
  }
  // S = M / 16^passes in {16, 32, 64, 128}
  if (S == 128) dif_pass<8>(a, t, M, S);
  else if (S == 64) dif_pass<4>(a, t, M, S);
  else if (S == 32) dif_pass<2>(a, t, M, S);
  if (S > 16) __syncthreads();
  mid16<CONVOLVE>(a, B, M);
  __syncthreads();
  if (!CONVOLVE) return;
  if (S == 128) dit_pass<8>(a, t, M, S);
  else if (S == 64) dit_pass<4>(a, t, M, S);
  else if (S == 32) dit_pass<2>(a, t, M, S);
  if (S > 16) __syncthreads();
  for (int s2 = S << 4; s2 <= M; s2 <<= 4) {  // the radix-16 passes back, smallest block size first
    t -= 4 * (s2 >> 4);
    dit_pass<16>(a, t, M, s2);
    __syncthreads();
  }
}

// scratch traffic stays in the L2 (.cg), the map pixels and phase rows are touched once and stream through it (.cs)
__device__ __forceinline__ double2 ldcg2(const double2 *p) { return __ldcg(p); }
__device__ __forceinline__ void stcg2(double2 *p, double2 v) { __stcg(p, v); }
__device__ __forceinline__ double4 ldcs4(const double *p) {  // one 32-byte phase row
  const double2 a = __ldcs(reinterpret_cast<const double2 *>(p)), b = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void stcs4(double *p, double4 v) {
  __stcs(reinterpret_cast<double2 *>(p), make_double2(v.x, v.y));
  __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(v.z, v.w));
}

struct R2Args {
  int nside, lmax, ncomp;
  int ihi, nrings;  // north ring numbers ihi, ihi - 1, ... (heaviest first); belt: any order
  int Mmax;         // largest transform of this launch (sizes the shared tile); the belt has M = nside for every item
  i64 npix, ncap;
  hcu_ptrs maps;
  const double2 *bfilt;
  const i64 *boff;
  const double2 *chirp, *wtab, *wbelt, *tw;
  int tw_off[14];
  double2 *scr;
  int scrL;
  int nm;
  const int32_t *mlist, *mpos;
  const double *rw;
  i64 rp_lo, nrp_local;
  double *phase;
  hcu_rowdest dest;  // forward: where the rows go (nd == 0: phase)
};

// floor(x / d) for 0 <= x < 2^23, d >= 1, through the float reciprocal (one correction step)
__device__ __forceinline__ int div_small(int x, int d, float rd, int *rem) {
  int q = __float2int_rd((float)x * rd);
  int r = x - q * d;
  if (r < 0) {
    --q;
    r += d;
  } else if (r >= d) {
    ++q;
    r -= d;
  }
  *rem = r;
  return q;
}

// exp(-i pi m / (4 L)) from the radix-4 twiddle table wt[j] = exp(-i pi j / (2 L)), j < L, and h = exp(-i pi / (4 L))
__device__ __forceinline__ double2 ring_phase(const double2 *wt, int L, float rL, int m, double2 h) {
  int r;
  const int q = div_small(m >> 1, L, rL, &r);
  const double2 t = __ldg(wt + r);
  // times (-i)^q: (x, y), (y, -x), (-x, -y), (-y, x)
  const double ax = (q & 1) ? t.y : t.x, ay = (q & 1) ? t.x : t.y;
  const double2 e = make_double2((q & 2) ? -ax : ax, (((q + 1) & 2) ? -ay : ay));
  return (m & 1) ? cmul(e, h) : e;
}

__device__ __forceinline__ int tw2_size_dev(int p) {
  int n = 0, S = 1 << p;
  while (S >= 256) {
    n += 4 * (S >> 4);
    S >>= 4;
  }
  if (S > 16) n += 16 * (S == 128 ? 3 : S == 64 ? 2 : 1);
  return n;
}

template <bool BLU, bool INV>
__global__ void __launch_bounds__(512, 1) ring2_kernel(const R2Args A) {
  extern __shared__ double2 smem2[];
  double2 *a = smem2;
  double2 *T = smem2 + (A.Mmax + (A.Mmax >> 4));
  const int tid = threadIdx.x, NT = blockDim.x;
  double2 *scr = A.scr + (i64)blockIdx.x * 4 * A.scrL;
  const int nitems = A.nrings * A.ncomp;
  int p_loaded = -1;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int c = item % A.ncomp;
    const int ir = A.ihi - item / A.ncomp;  // north ring number, 1-based
    const bool cap = ir < A.nside;
    const int L = cap ? ir : A.nside;
    const int n = 4 * L;
    int M = L;
    if (BLU) {
      M = 16;
      while (M < 2 * L - 1) M <<= 1;
    }
    const int p = 31 - __clz(M);
    if (p != p_loaded) {  // (a barrier separates this from the previous item's last pass)
      const double2 *src = A.tw + A.tw_off[p];
      const int twn = tw2_size_dev(p);
      for (int t = tid; t < twn; t += NT) T[t] = __ldg(src + t);
      p_loaded = p;
    }
    const i64 startN = cap ? 2LL * ir * (ir - 1) : A.ncap + (i64)(ir - A.nside) * n;
    const i64 startS = A.npix - startN - n;
    const bool equator = ir == 2 * A.nside;
    const bool shifted = cap || (((ir - A.nside) & 1) == 0);
    const double2 *wt = cap ? A.wtab + (i64)ir * (ir - 1) / 2 : A.wbelt;
    const double2 *ch = BLU ? A.chirp + (i64)ir * (ir - 1) / 2 : nullptr;
    const double2 *B = BLU ? A.bfilt + A.boff[ir] : nullptr;
    const i64 rp = ir - 1;
    const int sh = 32 - p;  // direct transforms leave k at the bit-reversed position
    const double2 h = expmipi(0.25 / (double)L);
    const float rL = 1.0f / (float)L, rn = 1.0f / (float)n;
    if (!INV) {
      const double *mN = A.maps.p[c] + startN;
      const double *mS = A.maps.p[c] + startS;
      // ONE pass over the ring pair: the radix-4 step makes all four sub-sequences; the first goes to the shared
      // tile, the others wait in the scratch
#pragma unroll 2
      for (int j = tid; j < L; j += NT) {
        double2 z0, z1, z2, z3;
        z0.x = __ldcs(mN + j), z1.x = __ldcs(mN + j + L), z2.x = __ldcs(mN + j + 2 * L), z3.x = __ldcs(mN + j + 3 * L);
        if (equator) {
          z0.y = z1.y = z2.y = z3.y = 0.;
        } else {
          z0.y = __ldcs(mS + j), z1.y = __ldcs(mS + j + L), z2.y = __ldcs(mS + j + 2 * L), z3.y = __ldcs(mS + j + 3 * L);
        }
        const double2 w = __ldg(wt + j);
        const double2 e02 = cadd(z0, z2), o02 = csub(z0, z2), e13 = cadd(z1, z3), o13 = csub(z1, z3);
        double2 y0 = cadd(e02, e13), y2 = csub(e02, e13);
        double2 y1 = cadd(o02, mul_mi(o13)), y3 = cadd(o02, mul_pi(o13));
        const double2 w2 = cmul(w, w);
        y1 = cmul(y1, w);
        y2 = cmul(y2, w2);
        y3 = cmul(y3, cmul(w2, w));
        if (BLU) {
          const double2 cj = __ldg(ch + j);
          y0 = cmul(y0, cj), y1 = cmul(y1, cj), y2 = cmul(y2, cj), y3 = cmul(y3, cj);
        }
        a[pad(j)] = y0;
        stcg2(scr + L + j, y1);
        stcg2(scr + 2 * L + j, y2);
        stcg2(scr + 3 * L + j, y3);
      }
      for (int s = 0; s < 4; ++s) {
        if (s) {
#pragma unroll 4
          for (int j = tid; j < L; j += NT) a[pad(j)] = ldcg2(scr + s * L + j);
        }
        if (BLU)
          for (int j = L + tid; j < M; j += NT) a[pad(j)] = make_double2(0., 0.);
        __syncthreads();
        fft_smem<BLU>(a, T, B, M);
#pragma unroll 4
        for (int k = tid; k < L; k += NT) {
          double2 v = a[pad(k)];
          if (BLU) v = cmul(v, __ldg(ch + k));
          stcg2(scr + s * L + k, v);  // direct: position k holds Z[4 brev(k) + s]
        }
        __syncthreads();
      }
      // untangle N / S, fold to the rows of this launch, ring phase and quadrature weight
      double w = 4.0 * 3.141592653589793238462643383279502884197 / (double)A.npix;
      if (A.rw) w *= A.rw[rp];
      const i64 ooff = ((rp - A.rp_lo) * A.ncomp + c) * 4;
      const i64 rstride = A.nrp_local * A.ncomp * 4;
      const int nm = A.nm;
      auto zaddr = [&](int k) {  // where Z[k] waits in the scratch
        const int q = k >> 2;
        return scr + (k & 3) * L + (BLU ? q : (int)(__brev((unsigned)q) >> sh));
      };
      auto emit = [&](int row, double2 za, double2 zb, double2 ph) {
        // N[k] = (Z[k] + conj Z[n-k]) / 2 ; S[k] = (Z[k] - conj Z[n-k]) / (2i)
        const double2 xn = make_double2(0.5 * (za.x + zb.x), 0.5 * (za.y - zb.y));
        const double2 d = make_double2(za.x - zb.x, za.y + zb.y);
        const double2 xs = make_double2(0.5 * d.y, -0.5 * d.x);
        const double2 pp = cmul(make_double2((xn.x + xs.x) * w, (xn.y + xs.y) * w), ph);
        const double2 qq = cmul(make_double2((xn.x - xs.x) * w, (xn.y - xs.y) * w), ph);
        stcs4(hcu_row_ptr(A.dest, A.phase, row, rstride) + ooff, make_double4(pp.x, pp.y, qq.x, qq.y));
      };
      // two rows in flight per thread; the second is clamped for its loads and dropped at the store
      for (int row0 = tid; row0 < nm; row0 += 2 * NT) {
        const int row1 = row0 + NT;
        const int row1c = row1 < nm ? row1 : row0;
        const int m0 = A.mlist ? __ldg(A.mlist + row0) : row0;
        const int m1 = A.mlist ? __ldg(A.mlist + row1c) : row1c;
        int k0, k1;
        div_small(m0, n, rn, &k0);
        div_small(m1, n, rn, &k1);
        const double2 za0 = ldcg2(zaddr(k0)), zb0 = ldcg2(zaddr(k0 ? n - k0 : 0));
        const double2 za1 = ldcg2(zaddr(k1)), zb1 = ldcg2(zaddr(k1 ? n - k1 : 0));
        const double2 ph0 = shifted ? ring_phase(wt, L, rL, m0, h) : make_double2(1., 0.);
        const double2 ph1 = shifted ? ring_phase(wt, L, rL, m1, h) : make_double2(1., 0.);
        emit(row0, za0, zb0, ph0);
        if (row1 < nm) emit(row1, za1, zb1, ph1);
      }
      __syncthreads();
    } else {
      // fold the rows onto the n frequencies: Z[k] = GN[k] + i GS[k], G[n - k] = conj G[k]
      const double *prow = A.phase + ((rp - A.rp_lo) * A.ncomp + c) * 4;
      const i64 rstride = A.nrp_local * A.ncomp * 4;
      // frequency k <= n / 2 takes row k and, conjugated, row n - k (plus their aliases k + a n, a n - k on the
      // small rings); its mirror n - k is the conjugate.  The first hit of both is loaded two frequencies ahead.
      for (int k0 = tid; k0 <= 2 * L; k0 += 2 * NT) {
        double4 pa[2], pb[2];
        double2 ea[2], eb[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = k0 + u * NT;
          pa[u] = pb[u] = make_double4(0., 0., 0., 0.);
          ea[u] = eb[u] = make_double2(1., 0.);
          if (k <= 2 * L) {
            const int m2 = n - k;
            if (k <= A.lmax) {
              const int row = A.mpos ? A.mpos[k] : k;
              if (row >= 0) {
                pa[u] = ldcs4(prow + (i64)row * rstride);
                if (shifted) ea[u] = ring_phase(wt, L, rL, k, h);
              }
            }
            if (m2 <= A.lmax) {
              const int row = A.mpos ? A.mpos[m2] : m2;
              if (row >= 0) {
                pb[u] = ldcs4(prow + (i64)row * rstride);
                if (shifted) eb[u] = ring_phase(wt, L, rL, m2, h);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int k = k0 + u * NT;
          if (k > 2 * L) continue;
          double2 gn = cadd(cmulc(make_double2(pa[u].x, pa[u].y), ea[u]), conj2(cmulc(make_double2(pb[u].x, pb[u].y), eb[u])));
          double2 gs = cadd(cmulc(make_double2(pa[u].z, pa[u].w), ea[u]), conj2(cmulc(make_double2(pb[u].z, pb[u].w), eb[u])));
          for (int m = k + n; m <= A.lmax; m += n) {
            const int row = A.mpos ? A.mpos[m] : m;
            if (row < 0) continue;
            const double4 pv = ldcs4(prow + (i64)row * rstride);
            const double2 e = shifted ? ring_phase(wt, L, rL, m, h) : make_double2(1., 0.);
            gn = cadd(gn, cmulc(make_double2(pv.x, pv.y), e));
            gs = cadd(gs, cmulc(make_double2(pv.z, pv.w), e));
          }
          for (int m = 2 * n - k; m <= A.lmax; m += n) {
            const int row = A.mpos ? A.mpos[m] : m;
            if (row < 0) continue;
            const double4 pv = ldcs4(prow + (i64)row * rstride);
            const double2 e = shifted ? ring_phase(wt, L, rL, m, h) : make_double2(1., 0.);
            gn = cadd(gn, conj2(cmulc(make_double2(pv.x, pv.y), e)));
            gs = cadd(gs, conj2(cmulc(make_double2(pv.z, pv.w), e)));
          }
          stcg2(scr + (k & 3) * L + (k >> 2), make_double2(gn.x - gs.y, gn.y + gs.x));
          if (k != 0 && k != 2 * L) {
            const int k2 = n - k;
            stcg2(scr + (k2 & 3) * L + (k2 >> 2), make_double2(gn.x + gs.y, gs.x - gn.y));
          }
        }
      }
      __syncthreads();
      for (int s = 0; s < 4; ++s) {
        // y_s[j] = sum_k Z[4 k + s] e^{+2 pi i j k / L} = conj(DFT_L(conj Z_s))[j]
#pragma unroll 4
        for (int k = tid; k < L; k += NT) {
          double2 v = conj2(ldcg2(scr + s * L + k));
          if (BLU) v = cmul(v, __ldg(ch + k));
          a[pad(k)] = v;
        }
        if (BLU)
          for (int j = L + tid; j < M; j += NT) a[pad(j)] = make_double2(0., 0.);
        __syncthreads();
        fft_smem<BLU>(a, T, B, M);
#pragma unroll 4
        for (int j = tid; j < L; j += NT) {
          double2 v = a[pad(j)];
          if (BLU) v = cmul(v, __ldg(ch + j));
          stcg2(scr + s * L + j, conj2(v));
        }
        __syncthreads();
      }
      // x[j + t L] = sum_s i^(s t) conj(w_n^j)^s y_s[j]
      double *mN = A.maps.p[c] + startN;
      double *mS = A.maps.p[c] + startS;
#pragma unroll 2
      for (int j = tid; j < L; j += NT) {
        const int pj = BLU ? j : (int)(__brev((unsigned)j) >> sh);
        const double2 w = __ldg(wt + j);
        const double2 r0 = ldcg2(scr + pj), r1 = ldcg2(scr + L + pj), r2 = ldcg2(scr + 2 * L + pj),
                      r3 = ldcg2(scr + 3 * L + pj);
        const double2 w2 = cmul(w, w);
        const double2 y0 = r0;
        const double2 y1 = cmulc(r1, w);
        const double2 y2 = cmulc(r2, w2);
        const double2 y3 = cmulc(r3, cmul(w2, w));
        const double2 e02 = cadd(y0, y2), o02 = csub(y0, y2), e13 = cadd(y1, y3), o13 = csub(y1, y3);
        const double2 x0 = cadd(e02, e13), x2 = csub(e02, e13);
        const double2 x1 = cadd(o02, mul_pi(o13)), x3 = cadd(o02, mul_mi(o13));
        __stcs(mN + j, x0.x), __stcs(mN + j + L, x1.x), __stcs(mN + j + 2 * L, x2.x), __stcs(mN + j + 3 * L, x3.x);
        if (!equator)
          __stcs(mS + j, x0.y), __stcs(mS + j + L, x1.y), __stcs(mS + j + 2 * L, x2.y), __stcs(mS + j + 3 * L, x3.y);
      }
      __syncthreads();
    }
  }
}

// twiddle tables of fft_smem for M = 2^p
__global__ void tw2_kernel(int p, double2 *T) {
  const int M = 1 << p;
  int off = 0, S = M;
  while (S >= 256) {
    const int nn = S >> 4;
    for (int x = threadIdx.x; x < 4 * nn; x += blockDim.x) {
      const int t = x / nn, k = x - t * nn;
      T[off + x] = expmipi(2.0 * (double)(k << t) / (double)S);
    }
    off += 4 * nn;
    S >>= 4;
  }
  if (S > 16) {
    const int nt = S == 128 ? 3 : S == 64 ? 2 : 1;
    const int nn = 16;
    for (int x = threadIdx.x; x < nt * nn; x += blockDim.x) {
      const int t = x / nn, k = x - t * nn;
      T[off + x] = expmipi(2.0 * (double)(k << t) / (double)S);
    }
  }
}

// per cap ring i (block), j < i: chirp exp(-i pi j^2 / i) and radix-4 twiddle exp(-i pi j / (2 i))
__global__ void ring_tables_kernel(int imax, double2 *chirp, double2 *wtab) {
  const int i = blockIdx.x + 1;
  if (i > imax) return;
  const i64 off = (i64)i * (i - 1) / 2;
  for (int j = threadIdx.x; j < i; j += blockDim.x) {
    const long long r = ((long long)j * j) % (2LL * i);
    chirp[off + j] = expmipi((double)r / (double)i);
    wtab[off + j] = expmipi((double)j / (2.0 * (double)i));
  }
}
__global__ void belt_table_kernel(int nside, double2 *wbelt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < nside) wbelt[j] = expmipi((double)j / (2.0 * (double)nside));
}

int tw2_size(int p) {
  int n = 0, S = 1 << p;
  while (S >= 256) {
    n += 4 * (S >> 4);
    S >>= 4;
  }
  if (S > 16) n += 16 * (S == 128 ? 3 : S == 64 ? 2 : 1);
  return n;
}

int bluestein_M2(int i) {
  int need = 2 * i - 1, M = 16;
  while (M < need) M <<= 1;
  return M;
}

int ring2_threads(int M) {
  int t = M / 16;
  return t < 128 ? 128 : t > 512 ? 512 : t;
}

// one persistent launch over north ring numbers ihi, ihi - 1, ... (nrings of them) x components
template <bool BLU, bool INV>
int launch_group(hcu_ctx *ctx, R2Args &A, int Lmax) {
  int NT = ring2_threads(A.Mmax);
  int occ_cap = 3;  // more resident CTAs only push the scratch out of the L2
  // tuning knobs: threads per CTA of the belt launch / of the cap launch below the largest Bluestein length
  if (!BLU) {
    if (A.Mmax >= 4096) NT = 512;  // measured at nside 4096: one 512-thread CTA per SM beats two of 256 (emission and loads use all threads)
    if (const char *e = getenv("HCU_R2_BELT_NT")) NT = atoi(e) >= 32 && atoi(e) <= 512 ? atoi(e) : NT;
    if (const char *e = getenv("HCU_R2_BELT_OCC")) occ_cap = atoi(e) >= 1 ? atoi(e) : occ_cap;
  } else if (A.Mmax < 8192) {
    if (const char *e = getenv("HCU_R2_REST_NT")) NT = atoi(e) >= 32 && atoi(e) <= 512 ? atoi(e) : NT;
  }
  const size_t smem = sizeof(double2) * (size_t)(A.Mmax + (A.Mmax >> 4) + tw2_size(ilog2_host(A.Mmax)));
  HCU_CUDA(cudaFuncSetAttribute(ring2_kernel<BLU, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  HCU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring2_kernel<BLU, INV>, NT, smem));
  if (occ < 1) {
    hcu_set_error("ring FFT kernel does not fit (M = %d)", A.Mmax);
    return HCU_ERR_CUDA;
  }
  if (occ > occ_cap) occ = occ_cap;
  const i64 items = (i64)A.nrings * A.ncomp;
  i64 grid = (i64)ctx->num_sms * occ;
  if (grid > items) grid = items;
  HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_scr, sizeof(double2) * 4 * (size_t)Lmax * (size_t)(ctx->num_sms * 3)));
  A.scr = (double2 *)ctx->ws_scr.ptr;
  A.scrL = Lmax;
  // keep the scratch resident in the L2: persisting access window over the part this launch uses
  static int persist = -1;
  if (persist < 0) {
    const char *e = getenv("HCU_R2_PERSIST");
    persist = e ? atoi(e) : 0;
    if (persist) {
      int maxp = 0;
      cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
      if (maxp <= 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) != cudaSuccess) persist = 0;
      cudaGetLastError();
    }
  }
  if (persist) {
    int maxw = 0;
    cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    size_t bytes = sizeof(double2) * 4 * (size_t)Lmax * (size_t)grid;
    if ((size_t)maxw < bytes) bytes = (size_t)maxw;
    cudaStreamAttrValue av;
    av.accessPolicyWindow.base_ptr = A.scr;
    av.accessPolicyWindow.num_bytes = bytes;
    av.accessPolicyWindow.hitRatio = 1.0f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
  }
  ring2_kernel<BLU, INV><<<(unsigned)grid, NT, smem, ctx->stream>>>(A);
  HCU_LAUNCH_CHECK(ctx);
  if (persist) {
    cudaStreamAttrValue av;
    av.accessPolicyWindow.base_ptr = nullptr;
    av.accessPolicyWindow.num_bytes = 0;
    av.accessPolicyWindow.hitRatio = 0.f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
  }
  return HCU_OK;
}

}  // namespace

// largest cap ring number the second generation handles (0: none)
int hcu_ring2_imax(const hcu_geom *g) { return g->r2_imax; }

int hcu_ring2_build(hcu_ctx *ctx, hcu_geom *g, int cap_max_m) {
  g->r2_imax = 0;
  g->r2_belt = false;
  const char *e = getenv("HCU_RINGFFT_GEN");
  if (e && atoi(e) == 1) return HCU_OK;
  const int nside = (int)g->nside;
  // twiddle tables for p = 4 .. 13
  int total = 0;
  for (int p = 4; p <= 13; ++p) {
    g->r2_tw_off[p] = total;
    total += tw2_size(p);
  }
  HCU_CUDA(cudaMalloc(&g->r2_tw, sizeof(double2) * total));
  for (int p = 4; p <= 13; ++p) {
    tw2_kernel<<<1, 256, 0, ctx->stream>>>(p, g->r2_tw + g->r2_tw_off[p]);
    HCU_LAUNCH_CHECK(ctx);
  }
  if (nside >= 16 && nside <= 8192) {
    HCU_CUDA(cudaMalloc(&g->r2_wbelt, sizeof(double2) * nside));
    belt_table_kernel<<<(nside + 255) / 256, 256, 0, ctx->stream>>>(nside, g->r2_wbelt);
    HCU_LAUNCH_CHECK(ctx);
    g->r2_belt = true;
  }
  int imax = 0;
  for (int i = 1; i < nside; ++i)
    if (bluestein_M2(i) <= cap_max_m && bluestein_M2(i) <= 8192) imax = i;
  if (imax >= 1) {
    const size_t n = (size_t)imax * (imax + 1) / 2;
    HCU_CUDA(cudaMalloc(&g->r2_chirp, sizeof(double2) * n));
    HCU_CUDA(cudaMalloc(&g->r2_wtab, sizeof(double2) * n));
    ring_tables_kernel<<<imax, 256, 0, ctx->stream>>>(imax, g->r2_chirp, g->r2_wtab);
    HCU_LAUNCH_CHECK(ctx);
    g->r2_imax = imax;
  }
  return HCU_OK;
}

void hcu_ring2_free(hcu_geom *g) {
  if (g->r2_tw) cudaFree(g->r2_tw);
  if (g->r2_wbelt) cudaFree(g->r2_wbelt);
  if (g->r2_chirp) cudaFree(g->r2_chirp);
  if (g->r2_wtab) cudaFree(g->r2_wtab);
  g->r2_tw = g->r2_wbelt = g->r2_chirp = g->r2_wtab = nullptr;
}

// ring pairs [rp_a, rp_b) of the caps (north ring numbers rp + 1 in 1 .. r2_imax) or of the belt;
// inverse == false: maps -> phase rows (mlist, nm), inverse == true: phase rows (mpos) -> maps
int hcu_ring2_run(hcu_ctx *ctx, hcu_geom *g, bool inverse, bool belt, int lmax, int ncomp, const hcu_ptrs &maps,
                  const double *ring_weights, i64 rp_lo, i64 nrp_local, i64 rp_a, i64 rp_b, const int32_t *mlist,
                  int nm, const int32_t *mpos, double *phase, const hcu_rowdest *dest) {
  if (rp_a >= rp_b) return HCU_OK;
  const int nside = (int)g->nside;
  R2Args A;
  A.nside = nside;
  A.lmax = lmax;
  A.ncomp = ncomp;
  A.npix = 12LL * nside * nside;
  A.ncap = 2LL * nside * (nside - 1);
  A.maps = maps;
  A.bfilt = g->bfilt;
  A.boff = g->bfilt_off;
  A.chirp = g->r2_chirp;
  A.wtab = g->r2_wtab;
  A.wbelt = g->r2_wbelt;
  A.nm = nm;
  A.mlist = mlist;
  A.mpos = mpos;
  A.rw = ring_weights;
  A.rp_lo = rp_lo;
  A.nrp_local = nrp_local;
  A.phase = phase;
  if (dest) A.dest = *dest;
  for (int p = 0; p < 14; ++p) A.tw_off[p] = g->r2_tw_off[p];
  A.tw = g->r2_tw;
  if (belt) {
    A.Mmax = nside;
    A.ihi = (int)rp_b;
    A.nrings = (int)(rp_b - rp_a);
    return inverse ? launch_group<false, true>(ctx, A, nside) : launch_group<false, false>(ctx, A, nside);
  }
  // two launches: the rings of the largest Bluestein length (one CTA per SM), then everything below it with the
  // largest rings first -- every ring costs the same 32 (lmax + 1) bytes of phase rows, however small it is
  const int itop = (int)rp_b, ifirst = (int)rp_a + 1;
  const int Mtop = bluestein_M2(itop);
  int il = itop;
  while (il - 1 >= ifirst && bluestein_M2(il - 1) == Mtop) --il;
  A.Mmax = Mtop;
  A.ihi = itop;
  A.nrings = itop - il + 1;
  HCU_CHECK((inverse ? launch_group<true, true>(ctx, A, itop) : launch_group<true, false>(ctx, A, itop)));
  if (il > ifirst) {
    const int i2 = il - 1;
    A.Mmax = bluestein_M2(i2);
    A.ihi = i2;
    A.nrings = i2 - ifirst + 1;
    HCU_CHECK((inverse ? launch_group<true, true>(ctx, A, i2) : launch_group<true, false>(ctx, A, i2)));
  }
  return HCU_OK;
}
