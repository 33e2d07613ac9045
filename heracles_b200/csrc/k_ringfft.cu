// k_ringfft.cu -- ring FFT stage of the HEALPix spherical-harmonic transform.
//
// Replaces the FFT half of hp.map2alm / hp.alm2map (heracles/healpy.py:183-189).
//   * equatorial belt (2 nside + 1 rings of 4 nside pixels): cuFFT plan-many D2Z / Z2D
//   * polar caps (rings of 4 i pixels, i = 1..nside-1): hand-written kernel.
//     A north ring and its southern mirror are packed as ONE complex sequence
//     z = N + i S of length 4 i, split into 4 decimated sub-sequences of length
//     i, each transformed by a Bluestein (chirp-z) convolution on a power-of-two
//     radix-2 FFT held in shared memory (any i, including primes).  The final
//     radix-4 recombination, the N/S untangling, the alias fold to m <= lmax,
//     the ring phase exp(-i m phi0) and the quadrature weight are fused into
//     the post kernel that writes the m-major "phase" array for the Legendre
//     stage.
//
// phase layout: phase[((m_idx * nrp_local + rp_local) * ncomp + c) * 4 + {0,1,2,3}]
//   = (re+, im+, re-, im-),  + = north + south, - = north - south.
//
// HBM-bound: algorithmic bytes = 8 npix (map read) + 32 nrp (lmax+1) (phase write) per component.
#include <stdlib.h>

#include <algorithm>

#include "hcu_common.cuh"

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) {
  return make_double2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ double2 csub(double2 a, double2 b) {
  return make_double2(a.x - b.x, a.y - b.y);
}
// exp(-i pi x)
__device__ __forceinline__ double2 expmipi(double x) {
  double s, c;
  sincospi(x, &s, &c);
  return make_double2(c, -s);
}

__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ double2 mul_pi(double2 a) { return make_double2(-a.y, a.x); }  // a * (+i)

// natural order in -> bit-reversed order out (decimation in frequency).  Two radix-2 stages are
// fused into one pass over shared memory (4 loads, 4 stores, 2 twiddles per 4 points instead of
// 8 + 8 + 4) -- the stage is shared-memory bound; an odd log2 M starts with one plain radix-2 stage.
__device__ void fft_dif(double2 *a, const double2 *tw, int M) {
  int s = M >> 1, ts = 1;
  if (__popc(M - 1) & 1) {  // odd number of stages
    for (int p = threadIdx.x; p < (M >> 1); p += blockDim.x) {
      int k = p & (s - 1);
      int j = ((p - k) << 1) + k;
      double2 u = a[j], v = a[j + s];
      a[j] = cadd(u, v);
      a[j + s] = cmul(csub(u, v), tw[k * ts]);
    }
    __syncthreads();
    s >>= 1;
    ts <<= 1;
  }
  for (; s >= 2; s >>= 2, ts <<= 2) {
    const int h = s >> 1;
    for (int p = threadIdx.x; p < (M >> 2); p += blockDim.x) {
      const int k = p & (h - 1);
      const int j = ((p - k) << 2) + k;
      const double2 a0 = a[j], a1 = a[j + h], a2 = a[j + s], a3 = a[j + s + h];
      const double2 w1 = tw[k * ts], w2 = tw[2 * k * ts];
      const double2 u0 = cadd(a0, a2), u2 = cmul(csub(a0, a2), w1);
      const double2 u1 = cadd(a1, a3), u3 = mul_mi(cmul(csub(a1, a3), w1));
      a[j] = cadd(u0, u1);
      a[j + h] = cmul(csub(u0, u1), w2);
      a[j + s] = cadd(u2, u3);
      a[j + s + h] = cmul(csub(u2, u3), w2);
    }
    __syncthreads();
  }
}
// exact stage-by-stage inverse of fft_dif WITHOUT the 1/2 per stage
// (bit-reversed in -> natural out); the 1/M is folded into the filter table
__device__ void fft_dit_inv(double2 *a, const double2 *tw, int M) {
  const bool odd = __popc(M - 1) & 1;
  const int smax = odd ? (M >> 2) : (M >> 1);  // largest span handled by the fused passes
  for (int h = 1; 2 * h <= smax; h <<= 2) {
    const int s = 2 * h, ts = M / (2 * s);
    for (int p = threadIdx.x; p < (M >> 2); p += blockDim.x) {
      const int k = p & (h - 1);
      const int j = ((p - k) << 2) + k;
      const double2 v0 = a[j], v1 = a[j + h], v2 = a[j + s], v3 = a[j + s + h];
      const double2 w1 = tw[k * ts], w2 = tw[2 * k * ts];
      const double2 t1 = cmulc(v1, w2), t3 = cmulc(v3, w2);
      const double2 u0 = cadd(v0, t1), u1 = csub(v0, t1);
      const double2 u2 = cadd(v2, t3), u3 = csub(v2, t3);
      const double2 t2 = cmulc(u2, w1), t4 = mul_pi(cmulc(u3, w1));
      a[j] = cadd(u0, t2);
      a[j + s] = csub(u0, t2);
      a[j + h] = cadd(u1, t4);
      a[j + s + h] = csub(u1, t4);
    }
    __syncthreads();
  }
  if (odd) {
    const int s = M >> 1;
    for (int p = threadIdx.x; p < (M >> 1); p += blockDim.x) {
      int k = p & (s - 1);
      int j = ((p - k) << 1) + k;
      double2 u = a[j];
      double2 t = cmulc(a[j + s], tw[k]);
      a[j] = cadd(u, t);
      a[j + s] = csub(u, t);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void make_twiddles(double2 *tw, int M) {
  for (int k = threadIdx.x; k < (M >> 1); k += blockDim.x)
    tw[k] = expmipi(2.0 * (double)k / (double)M);
}

// chirp c[j] = exp(-i pi j^2 / n)
__device__ __forceinline__ double2 chirp(int j, int n) {
  long long r = ((long long)j * j) % (2LL * n);
  return expmipi((double)r / (double)n);
}

// Bluestein filter spectrum for sub-FFT length i: FFT_M(b)/M in bit-reversed order
__global__ void bluestein_filter_kernel(int ilo, int M, double2 *bfilt, const i64 *off) {
  extern __shared__ double2 smem[];
  double2 *a = smem;
  double2 *tw = smem + M;
  const int n = ilo + blockIdx.x;
  make_twiddles(tw, M);
  for (int j = threadIdx.x; j < M; j += blockDim.x) a[j] = make_double2(0., 0.);
  __syncthreads();
  const double inv = 1.0 / (double)M;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double2 c = chirp(j, n);
    double2 b = make_double2(c.x * inv, -c.y * inv);  // conj(c)/M
    a[j] = b;
    if (j > 0) a[M - j] = b;
  }
  __syncthreads();
  fft_dif(a, tw, M);
  double2 *out = bfilt + off[n];
  for (int j = threadIdx.x; j < M; j += blockDim.x) out[j] = a[j];
}

// ---- sub-transforms whose Bluestein length M = 2 Mh does not fit shared memory (ring number i > 4096,
// i.e. nside 8192): the FIRST radix-2 stage of the length-M transform is done "out of core".  A
// decimation-in-frequency stage splits x into u[j] = x[j] + x[j + Mh] and d[j] = (x[j] - x[j + Mh]) w_M^j,
// whose length-Mh transforms are the even and the odd frequencies -- in bit-reversed order exactly the
// first and the second half of the length-M spectrum, so the two halves are transformed, multiplied by
// their half of the filter and transformed back one after the other in the SAME shared tile, and the
// last (inverse) stage x[k] = r0[k] + conj(w_M^k) r1[k] is applied while the second half is written.
// The chirp-multiplied input is non-zero only for j < i <= Mh, so u = x and d = x w_M^j.
__device__ __forceinline__ double2 w_big(int j, int Mh) { return expmipi((double)j / (double)Mh); }  // exp(-2 pi i j / (2 Mh))

__global__ void bluestein_filter_big_kernel(int ilo, int Mh, double2 *bfilt, const i64 *off) {
  extern __shared__ double2 smem[];
  double2 *a = smem;
  double2 *tw = smem + Mh;
  const int n = ilo + blockIdx.x;
  const int M = 2 * Mh;
  make_twiddles(tw, Mh);
  const double inv = 1.0 / (double)M;
  double2 *out = bfilt + off[n];
  for (int half = 0; half < 2; ++half) {
    __syncthreads();
    for (int j = threadIdx.x; j < Mh; j += blockDim.x) {
      // b[j] = conj(chirp(j)) / M for j < n; b[M - j] = b[j]: the second half holds x1[j] = b[Mh - j] for Mh - j < n
      double2 x0 = make_double2(0., 0.), x1 = make_double2(0., 0.);
      if (j < n) {
        const double2 c = chirp(j, n);
        x0 = make_double2(c.x * inv, -c.y * inv);
      }
      if (j > 0 && Mh - j < n) {
        const double2 c = chirp(Mh - j, n);
        x1 = make_double2(c.x * inv, -c.y * inv);
      }
      a[j] = half == 0 ? cadd(x0, x1) : cmul(csub(x0, x1), w_big(j, Mh));
    }
    __syncthreads();
    fft_dif(a, tw, Mh);
    for (int j = threadIdx.x; j < Mh; j += blockDim.x) out[(i64)half * Mh + j] = a[j];
  }
}

// the convolution of one sub-sequence with the chirp filter, length M = 2 Mh:
//   load(j)      chirp-multiplied input, j < i
//   stash / fetch (k, v)   park the first half's result r0[k] (k < i) in global memory (same thread both times)
//   fin(k, r)    receives r[k], k < i
template <class Load, class Stash, class Fetch, class Fin>
__device__ __forceinline__ void bluestein_big(double2 *a, const double2 *tw, const double2 *B, int Mh, int i,
                                              Load load, Stash stash, Fetch fetch, Fin fin) {
  for (int half = 0; half < 2; ++half) {
    for (int j = threadIdx.x; j < Mh; j += blockDim.x) {
      double2 v = make_double2(0., 0.);
      if (j < i) {
        v = load(j);
        if (half) v = cmul(v, w_big(j, Mh));
      }
      a[j] = v;
    }
    __syncthreads();
    fft_dif(a, tw, Mh);
    const double2 *Bh = B + (i64)half * Mh;
    for (int j = threadIdx.x; j < Mh; j += blockDim.x) a[j] = cmul(a[j], Bh[j]);
    __syncthreads();
    fft_dit_inv(a, tw, Mh);
    for (int k = threadIdx.x; k < i; k += blockDim.x) {
      if (half == 0)
        stash(k, a[k]);
      else
        fin(k, cadd(fetch(k), cmulc(a[k], w_big(k, Mh))));
    }
    __syncthreads();
  }
}

// forward: one block per (cap ring pair i, component)
__global__ void cap_fft_fwd_kernel(int ilo, int M, i64 nside, hcu_ptrs maps,
                                   const double2 *bfilt,
                                   const i64 *off, double2 *Y, i64 ncap) {
  extern __shared__ double2 smem[];
  double2 *a = smem;
  double2 *tw = smem + M;
  const int i = ilo + blockIdx.x;
  const int c = blockIdx.y;
  const i64 npix = 12 * nside * nside;
  const i64 startN = 2LL * i * (i - 1);
  const i64 startS = npix - startN - 4LL * i;
  const double *mN = maps.p[c] + startN;
  const double *mS = maps.p[c] + startS;
  const double2 *B = bfilt + off[i];
  double2 *Yc = Y + (i64)c * ncap + startN;
  make_twiddles(tw, M);
  for (int q = 0; q < 4; ++q) {
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      double2 v = make_double2(0., 0.);
      if (j < i) v = cmul(make_double2(mN[4 * j + q], mS[4 * j + q]), chirp(j, i));
      a[j] = v;
    }
    __syncthreads();
    fft_dif(a, tw, M);
    for (int j = threadIdx.x; j < M; j += blockDim.x) a[j] = cmul(a[j], B[j]);
    __syncthreads();
    fft_dit_inv(a, tw, M);
    for (int k = threadIdx.x; k < i; k += blockDim.x)
      Yc[(i64)q * i + k] = cmul(a[k], chirp(k, i));
    __syncthreads();
  }
}

// the same for rings whose Bluestein length is 2 Mh (see bluestein_big)
__global__ void cap_fft_fwd_big_kernel(int ilo, int Mh, i64 nside, hcu_ptrs maps, const double2 *bfilt,
                                       const i64 *off, double2 *Y, i64 ncap) {
  extern __shared__ double2 smem[];
  double2 *a = smem;
  double2 *tw = smem + Mh;
  const int i = ilo + blockIdx.x;
  const int c = blockIdx.y;
  const i64 npix = 12 * nside * nside;
  const i64 startN = 2LL * i * (i - 1);
  const i64 startS = npix - startN - 4LL * i;
  const double *mN = maps.p[c] + startN;
  const double *mS = maps.p[c] + startS;
  const double2 *B = bfilt + off[i];
  double2 *Yc = Y + (i64)c * ncap + startN;
  make_twiddles(tw, Mh);
  __syncthreads();
  for (int q = 0; q < 4; ++q) {
    double2 *Yq = Yc + (i64)q * i;
    bluestein_big(
        a, tw, B, Mh, i, [&](int j) { return cmul(make_double2(mN[4 * j + q], mS[4 * j + q]), chirp(j, i)); },
        [&](int k, double2 v) { Yq[k] = v; }, [&](int k) { return Yq[k]; },
        [&](int k, double2 r) { Yq[k] = cmul(r, chirp(k, i)); });
  }
}

// Z[k] of the packed length-4i sequence from the four sub-FFTs
__device__ __forceinline__ double2 cap_combine(const double2 *Yr, int i, int k) {
  int kk = k % i;
  double2 w1 = expmipi((double)k / (2.0 * (double)i));  // exp(-2 pi i k / (4 i))
  double2 w2 = cmul(w1, w1);
  double2 w3 = cmul(w2, w1);
  double2 z = Yr[kk];
  z = cadd(z, cmul(w1, Yr[(i64)i + kk]));
  z = cadd(z, cmul(w2, Yr[2LL * i + kk]));
  z = cadd(z, cmul(w3, Yr[3LL * i + kk]));
  return z;
}

__device__ __forceinline__ void store_phase(double *o4, double2 n,
                                            double2 s, double2 ph, double w) {
  double2 p = make_double2((n.x + s.x) * w, (n.y + s.y) * w);
  double2 q = make_double2((n.x - s.x) * w, (n.y - s.y) * w);
  p = cmul(p, ph);
  q = cmul(q, ph);
  double4 *o = reinterpret_cast<double4 *>(o4);
  *o = make_double4(p.x, p.y, q.x, q.y);
}

// grid: (row blocks, cap ring pairs in range, components); row -> m through mlist (nullptr: identity)
__global__ void cap_post_kernel(i64 nside, int nm, const int32_t *mlist, int ncomp,
                                const double2 *Y, i64 ncap, const double *ring_weights,
                                i64 rp_lo, i64 nrp_local, i64 rp_first,
                                double *phase, const hcu_rowdest dest) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nm) return;
  const int m = mlist ? mlist[row] : row;
  const int comp = blockIdx.z;
  const double2 *Yc = Y + (i64)comp * ncap;
  const i64 rp = rp_first + blockIdx.y;
  const int i = (int)rp + 1;
  const int n = 4 * i;
  const double2 *Yr = Yc + 2LL * i * (i - 1);
  const int k = m % n;
  const int k2 = (n - k) % n;
  double2 z1 = cap_combine(Yr, i, k);
  double2 z2 = cap_combine(Yr, i, k2);
  // N[k] = (Z[k] + conj Z[n-k])/2 ; S[k] = (Z[k] - conj Z[n-k])/(2i)
  double2 xn = make_double2(0.5 * (z1.x + z2.x), 0.5 * (z1.y - z2.y));
  double2 d = make_double2(z1.x - z2.x, z1.y + z2.y);
  double2 xs = make_double2(0.5 * d.y, -0.5 * d.x);
  double w = 4.0 * 3.141592653589793238462643383279502884197 /
             (12.0 * (double)nside * (double)nside);
  if (ring_weights) w *= ring_weights[rp];
  double2 ph = expmipi((double)m / (4.0 * (double)i));
  store_phase(hcu_row_ptr(dest, phase, row, nrp_local * ncomp * 4) + ((rp - rp_lo) * ncomp + comp) * 4, xn, xs, ph, w);
}

// belt: X[rb][k], rb = ring - nside (0..2 nside), k = 0..2 nside
__global__ void belt_post_kernel(i64 nside, int nm, const int32_t *mlist, int ncomp, int comp,
                                 const double2 *X, const double *ring_weights,
                                 i64 rp_lo, i64 nrp_local, i64 rp_first,
                                 double *phase, const hcu_rowdest dest) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nm) return;
  const int m = mlist ? mlist[row] : row;
  const i64 rp = rp_first + blockIdx.y;
  const i64 i = rp + 1;  // north ring number, nside <= i <= 2 nside
  const int n4 = (int)(4 * nside);
  const int nk = n4 / 2 + 1;
  const i64 rbn = i - nside, rbs = 3 * nside - i;
  int k = m % n4;
  bool cj = k > n4 / 2;
  if (cj) k = n4 - k;
  double2 xn = X[rbn * nk + k];
  double2 xs = make_double2(0., 0.);
  if (rbs != rbn) xs = X[rbs * nk + k];
  if (cj) {
    xn.y = -xn.y;
    xs.y = -xs.y;
  }
  double w = 4.0 * 3.141592653589793238462643383279502884197 /
             (12.0 * (double)nside * (double)nside);
  if (ring_weights) w *= ring_weights[rp];
  double2 ph = make_double2(1., 0.);
  if (((i - nside) & 1) == 0) ph = expmipi((double)m / (4.0 * (double)nside));
  store_phase(hcu_row_ptr(dest, phase, row, nrp_local * ncomp * 4) + ((rp - rp_lo) * ncomp + comp) * 4, xn, xs, ph, w);
}

// ---------------------------------------------------------------------------
// inverse direction (synthesis): phase (b_m per ring) -> ring pixels
// phase layout for synthesis: phase[((row * nrp_local + rp - rp_lo) * ncomp + c) * 4] = (reN, imN, reS, imS),
// row = mpos[m] (nullptr: row = m; negative: this m is absent)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int row_of(const int32_t *mpos, int m) { return mpos ? mpos[m] : m; }

// belt: build the half-complex spectrum of the rings rb0 .. rb0 + gridDim.y - 1, then cuFFT Z2D
__global__ void belt_pre_inv_kernel(i64 nside, int lmax, const int32_t *mpos, int ncomp, int comp,
                                    const double *phase, i64 rp_lo, i64 nrp, i64 rb0, double2 *X) {
  const int n4 = (int)(4 * nside);
  const int nk = n4 / 2 + 1;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nk) return;
  const i64 rb = rb0 + blockIdx.y;  // 0..2 nside
  const i64 ring = rb + nside;
  const bool south = ring > 2 * nside;
  const i64 rp = (south ? 4 * nside - ring : ring) - 1 - rp_lo;
  const bool shifted = (((south ? 4 * nside - ring : ring) - nside) & 1) == 0;
  // G[k] = sum over m = k (mod n4) of c_m + sum over m = -k (mod n4), m>0, of conj(c_m)
  double2 g = make_double2(0., 0.);
  for (int m = k; m <= lmax; m += n4) {
    const int row = row_of(mpos, m);
    if (row < 0) continue;
    const double *p = phase + (((i64)row * nrp + rp) * ncomp + comp) * 4 + (south ? 2 : 0);
    double2 c = make_double2(p[0], p[1]);
    if (shifted) c = cmulc(c, expmipi((double)m / (4.0 * (double)nside)));
    g = cadd(g, c);
  }
  for (int m = n4 - k; m <= lmax; m += n4) {
    if (m == 0) continue;
    const int row = row_of(mpos, m);
    if (row < 0) continue;
    const double *p = phase + (((i64)row * nrp + rp) * ncomp + comp) * 4 + (south ? 2 : 0);
    double2 c = make_double2(p[0], p[1]);
    if (shifted) c = cmulc(c, expmipi((double)m / (4.0 * (double)nside)));
    g = cadd(g, make_double2(c.x, -c.y));
  }
  // f_j = sum_{k=0}^{n-1} G[k] e^{2 pi i jk/n}; Z2D computes sum over the half
  // spectrum assuming Hermitian symmetry, which G has by construction
  X[rb * nk + k] = g;
}

// caps: inverse of the forward scheme.  Build Z[k] = N-spectrum + i S-spectrum
// for k = 0..4i-1, split into 4 decimated inverse sub-FFTs via Bluestein.
// Z is the spectrum of z_j = fN_j + i fS_j: z_j = sum_k Z[k] e^{+2 pi i jk/n}.
// With j = 4 j' + q:  z_{4j'+q} = sum_{k'=0}^{i-1} e^{2 pi i j'k'/i}
//      [ e^{2 pi i q k'/n} sum_{s=0}^{3} Z[k' + s i] e^{2 pi i q s/4} ].
// Step 1 (cap_pre_inv_kernel): fold the m <= lmax coefficients of one ring pair
// onto the 4i frequencies, Z[k] = GN[k] + i GS[k]; grid (k blocks, cap ring pairs, comps).
__global__ void cap_pre_inv_kernel(i64 nside, int lmax, const int32_t *mpos, int ncomp, const double *phase,
                                   i64 rp_lo, i64 nrp, i64 rp_first, double2 *Z, i64 ncap) {
  const int i = (int)rp_first + blockIdx.y + 1;
  const int n = 4 * i;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int comp = blockIdx.z;
  const i64 rp = i - 1 - rp_lo;
  // G[k] folded from m = k (mod n) and, conjugated, from m = -k (mod n)
  double2 gn = make_double2(0., 0.), gs = make_double2(0., 0.);
  for (int m = k; m <= lmax; m += n) {
    const int row = row_of(mpos, m);
    if (row < 0) continue;
    const double4 p = *reinterpret_cast<const double4 *>(phase + (((i64)row * nrp + rp) * ncomp + comp) * 4);
    double2 e = expmipi((double)m / (4.0 * (double)i));
    gn = cadd(gn, cmulc(make_double2(p.x, p.y), e));
    gs = cadd(gs, cmulc(make_double2(p.z, p.w), e));
  }
  for (int m = n - k; m <= lmax; m += n) {
    if (m == 0) continue;
    const int row = row_of(mpos, m);
    if (row < 0) continue;
    const double4 p = *reinterpret_cast<const double4 *>(phase + (((i64)row * nrp + rp) * ncomp + comp) * 4);
    double2 e = expmipi((double)m / (4.0 * (double)i));
    double2 cn = cmulc(make_double2(p.x, p.y), e);
    double2 cs = cmulc(make_double2(p.z, p.w), e);
    gn = cadd(gn, make_double2(cn.x, -cn.y));
    gs = cadd(gs, make_double2(cs.x, -cs.y));
  }
  Z[(i64)comp * ncap + 2LL * i * (i - 1) + k] = make_double2(gn.x - gs.y, gn.y + gs.x);
}

// Step 2: four decimated inverse sub-FFTs per ring pair; grid (rings of one size class, comps)
__global__ void cap_fft_inv_kernel(int ilo, int M, i64 nside, const double2 *Z, i64 ncap,
                                   const double2 *bfilt, const i64 *off,
                                   hcu_ptrs maps) {
  extern __shared__ double2 smem[];
  double2 *a = smem;
  double2 *tw = smem + M;
  const int i = ilo + blockIdx.x;
  const int comp = blockIdx.y;
  const i64 npix = 12 * nside * nside;
  const i64 startN = 2LL * i * (i - 1);
  const i64 startS = npix - startN - 4LL * i;
  double *mN = maps.p[comp] + startN;
  double *mS = maps.p[comp] + startS;
  const double2 *Zr = Z + (i64)comp * ncap + startN;
  const double2 *B = bfilt + off[i];
  make_twiddles(tw, M);
  __syncthreads();
  for (int q = 0; q < 4; ++q) {
    // input of the sub-transform k' -> j' (inverse DFT = conj(DFT(conj(.))))
    for (int kp = threadIdx.x; kp < M; kp += blockDim.x) {
      double2 v = make_double2(0., 0.);
      if (kp < i) {
        double2 acc = make_double2(0., 0.);
        for (int s = 0; s < 4; ++s) {
          const double2 z = Zr[kp + s * i];
          // e^{2 pi i q s / 4}
          int r = (q * s) & 3;
          double2 zr = (r == 0) ? z
                     : (r == 1) ? make_double2(-z.y, z.x)
                     : (r == 2) ? make_double2(-z.x, -z.y)
                                : make_double2(z.y, -z.x);
          acc = cadd(acc, zr);
        }
        // times e^{2 pi i q k'/n} = conj(exp(-i pi q k' / (2 i)))
        double2 e = expmipi((double)(q * kp) / (2.0 * (double)i));
        acc = cmulc(acc, e);
        // conjugate for the inverse transform, then Bluestein pre-chirp
        v = cmul(make_double2(acc.x, -acc.y), chirp(kp, i));
      }
      a[kp] = v;
    }
    __syncthreads();
    fft_dif(a, tw, M);
    for (int j = threadIdx.x; j < M; j += blockDim.x) a[j] = cmul(a[j], B[j]);
    __syncthreads();
    fft_dit_inv(a, tw, M);
    for (int jp = threadIdx.x; jp < i; jp += blockDim.x) {
      double2 r = cmul(a[jp], chirp(jp, i));
      // undo the conjugation: z = conj(r)
      mN[4 * jp + q] = r.x;
      mS[4 * jp + q] = -r.y;
    }
    __syncthreads();
  }
}

__global__ void cap_fft_inv_big_kernel(int ilo, int Mh, i64 nside, const double2 *Z, i64 ncap,
                                       const double2 *bfilt, const i64 *off, hcu_ptrs maps) {
  extern __shared__ double2 smem[];
  double2 *a = smem;
  double2 *tw = smem + Mh;
  const int i = ilo + blockIdx.x;
  const int comp = blockIdx.y;
  const i64 npix = 12 * nside * nside;
  const i64 startN = 2LL * i * (i - 1);
  const i64 startS = npix - startN - 4LL * i;
  double *mN = maps.p[comp] + startN;
  double *mS = maps.p[comp] + startS;
  const double2 *Zr = Z + (i64)comp * ncap + startN;
  const double2 *B = bfilt + off[i];
  make_twiddles(tw, Mh);
  __syncthreads();
  for (int q = 0; q < 4; ++q) {
    bluestein_big(
        a, tw, B, Mh, i,
        [&](int kp) {  // see cap_fft_inv_kernel
          double2 acc = make_double2(0., 0.);
          for (int s = 0; s < 4; ++s) {
            const double2 z = Zr[kp + s * i];
            const int r = (q * s) & 3;
            const double2 zr = (r == 0) ? z
                             : (r == 1) ? make_double2(-z.y, z.x)
                             : (r == 2) ? make_double2(-z.x, -z.y)
                                        : make_double2(z.y, -z.x);
            acc = cadd(acc, zr);
          }
          const double2 e = expmipi((double)(q * kp) / (2.0 * (double)i));
          acc = cmulc(acc, e);
          return cmul(make_double2(acc.x, -acc.y), chirp(kp, i));
        },
        [&](int k, double2 v) { mN[4 * k + q] = v.x; mS[4 * k + q] = v.y; },
        [&](int k) { return make_double2(mN[4 * k + q], mS[4 * k + q]); },
        [&](int k, double2 rr) {
          const double2 r = cmul(rr, chirp(k, i));
          mN[4 * k + q] = r.x;
          mS[4 * k + q] = -r.y;
        });
  }
}

int bluestein_M(int i) {  // at least 16: the second generation's middle pass works on 16 points
  int need = 2 * i - 1;
  int M = 16;
  while (M < need) M <<= 1;
  return M;
}

// longest transform one CTA holds in shared memory (24 bytes per point).  HCU_CAP_MAX_M lowers it so that the
// tests can drive the two-half path of bluestein_big at small nside.
int cap_max_m() {
  static int v = 0;
  if (!v) {
    const char *e = getenv("HCU_CAP_MAX_M");
    int x = e ? atoi(e) : 8192;
    v = (x >= 4 && x <= 8192 && (x & (x - 1)) == 0) ? x : 8192;
  }
  return v;
}

int cap_threads(int M) {
  int t = M / 2;
  if (t < 32) t = 32;
  if (t > 512) t = 512;
  return t;
}

}  // namespace

int hcu_build_bluestein(hcu_ctx *ctx, hcu_geom *g) {
  const int nside = (int)g->nside;
  g->bfilt_off_h.assign(nside + 1, 0);
  i64 total = 0;
  for (int i = 1; i < nside; ++i) {
    g->bfilt_off_h[i] = total;
    total += bluestein_M(i);
  }
  if (total == 0) return hcu_ring2_build(ctx, g, cap_max_m());
  HCU_CUDA(cudaMalloc(&g->bfilt, sizeof(double2) * total));
  HCU_CUDA(cudaMalloc(&g->bfilt_off, sizeof(i64) * (nside + 1)));
  HCU_CUDA(cudaMemcpyAsync(g->bfilt_off, g->bfilt_off_h.data(), sizeof(i64) * (nside + 1),
                           cudaMemcpyHostToDevice, ctx->stream));
  int i = 1;
  while (i < nside) {
    int M = bluestein_M(i);
    int ihi = i;
    while (ihi + 1 < nside && bluestein_M(ihi + 1) == M) ++ihi;
    if (M > cap_max_m()) {
      const int Mh = M / 2;
      size_t smem = (size_t)Mh * 24;
      HCU_CUDA(cudaFuncSetAttribute(bluestein_filter_big_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bluestein_filter_big_kernel<<<ihi - i + 1, cap_threads(Mh), smem, ctx->stream>>>(i, Mh, g->bfilt, g->bfilt_off);
    } else {
      size_t smem = (size_t)M * 24;
      HCU_CUDA(cudaFuncSetAttribute(bluestein_filter_kernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bluestein_filter_kernel<<<ihi - i + 1, cap_threads(M), smem, ctx->stream>>>(
          i, M, g->bfilt, g->bfilt_off);
    }
    HCU_LAUNCH_CHECK(ctx);
    i = ihi + 1;
  }
  return hcu_ring2_build(ctx, g, cap_max_m());
}

// cuFFT plan-many over `batch` consecutive belt rings (cached per nside, batch and direction)
static int get_belt_plan(hcu_ctx *ctx, i64 nside, int batch, bool inverse, cufftHandle *out) {
  auto &cache = inverse ? ctx->belt_plan_inv : ctx->belt_plan;
  const i64 key = nside * (i64)(1 << 20) + batch;
  auto it = cache.find(key);
  if (it == cache.end()) {
    cufftHandle plan;
    int n4 = (int)(4 * nside);
    HCU_CUFFT(cufftCreate(&plan));
    size_t ws = 0;
    if (!inverse)
      HCU_CUFFT(cufftMakePlanMany(plan, 1, &n4, nullptr, 1, n4, nullptr, 1, n4 / 2 + 1,
                                  CUFFT_D2Z, batch, &ws));
    else
      HCU_CUFFT(cufftMakePlanMany(plan, 1, &n4, nullptr, 1, n4 / 2 + 1, nullptr, 1, n4,
                                  CUFFT_Z2D, batch, &ws));
    cache[key] = plan;
    it = cache.find(key);
  }
  HCU_CUFFT(cufftSetStream(it->second, ctx->stream));
  *out = it->second;
  return HCU_OK;
}

// the belt rings (rb = ring - nside) that ring pairs [belt_lo, belt_hi) touch: a northern run and its
// southern mirror; they are merged into one run when they touch or overlap (blocks around the equator)
static int belt_runs(i64 nside, i64 belt_lo, i64 belt_hi, i64 run0[2], i64 cnt[2]) {
  const i64 n_lo = belt_lo + 1 - nside, n_hi = belt_hi - nside;          // north rb range, inclusive
  i64 s_lo = 3 * nside - (belt_hi - 1) - 1, s_hi = 3 * nside - belt_lo - 1;  // south rb range, inclusive
  if (s_lo <= n_hi + 1) {  // touching / overlapping at the equator (rb = nside)
    run0[0] = n_lo;
    cnt[0] = s_hi - n_lo + 1;
    return 1;
  }
  run0[0] = n_lo; cnt[0] = n_hi - n_lo + 1;
  run0[1] = s_lo; cnt[1] = s_hi - s_lo + 1;
  return 2;
}

// comp stride of the first-generation cap workspace that covers north ring numbers <= imax
static i64 cap_stride(i64 imax) { return 2 * imax * (imax + 1); }

// first-generation cap transforms of ring pairs [lo, hi): sub-FFTs into Y, then the post kernel
static int old_caps_forward(hcu_ctx *ctx, hcu_geom *g, int ncomp, const hcu_ptrs &maps, const double *ring_weights,
                            i64 rp_lo, i64 nrp_local, i64 lo, i64 hi, const int32_t *mlist, int nm, double *phase,
                            i64 ystride, const hcu_rowdest &dest) {
  if (lo >= hi) return HCU_OK;
  const i64 nside = g->nside;
  const int mthreads = 128;
  const unsigned mblocks = (unsigned)((nm + mthreads - 1) / mthreads);
  double2 *Y = (double2 *)ctx->ws_cap.ptr;
  int i = (int)lo + 1;
  const int iend = (int)hi;  // inclusive ring number
  while (i <= iend) {
    int M = bluestein_M(i);
    int ihi = i;
    while (ihi + 1 <= iend && bluestein_M(ihi + 1) == M) ++ihi;
    dim3 grid(ihi - i + 1, ncomp);
    if (M > cap_max_m()) {
      const int Mh = M / 2;
      size_t smem = (size_t)Mh * 24;
      HCU_CUDA(cudaFuncSetAttribute(cap_fft_fwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cap_fft_fwd_big_kernel<<<grid, cap_threads(Mh), smem, ctx->stream>>>(i, Mh, nside, maps, g->bfilt,
                                                                           g->bfilt_off, Y, ystride);
    } else {
      size_t smem = (size_t)M * 24;
      HCU_CUDA(cudaFuncSetAttribute(cap_fft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cap_fft_fwd_kernel<<<grid, cap_threads(M), smem, ctx->stream>>>(i, M, nside, maps, g->bfilt, g->bfilt_off, Y,
                                                                      ystride);
    }
    HCU_LAUNCH_CHECK(ctx);
    i = ihi + 1;
  }
  dim3 grid(mblocks, (unsigned)(hi - lo), (unsigned)ncomp);
  cap_post_kernel<<<grid, mthreads, 0, ctx->stream>>>(nside, nm, mlist, ncomp, Y, ystride, ring_weights, rp_lo,
                                                      nrp_local, lo, phase, dest);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

static int old_caps_inverse(hcu_ctx *ctx, hcu_geom *g, int lmax, int ncomp, const double *phase, const int32_t *mpos,
                            i64 rp_lo, i64 nrp_local, i64 lo, i64 hi, const hcu_ptrs &maps, i64 zstride) {
  if (lo >= hi) return HCU_OK;
  const i64 nside = g->nside;
  double2 *Z = (double2 *)ctx->ws_cap.ptr;
  dim3 pgrid((unsigned)((4 * hi + 127) / 128), (unsigned)(hi - lo), (unsigned)ncomp);
  cap_pre_inv_kernel<<<pgrid, 128, 0, ctx->stream>>>(nside, lmax, mpos, ncomp, phase, rp_lo, nrp_local, lo, Z,
                                                     zstride);
  HCU_LAUNCH_CHECK(ctx);
  int i = (int)lo + 1;
  const int iend = (int)hi;
  while (i <= iend) {
    int M = bluestein_M(i);
    int ihi = i;
    while (ihi + 1 <= iend && bluestein_M(ihi + 1) == M) ++ihi;
    dim3 grid(ihi - i + 1, ncomp);
    if (M > cap_max_m()) {
      const int Mh = M / 2;
      size_t smem = (size_t)Mh * 24;
      HCU_CUDA(cudaFuncSetAttribute(cap_fft_inv_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cap_fft_inv_big_kernel<<<grid, cap_threads(Mh), smem, ctx->stream>>>(i, Mh, nside, Z, zstride, g->bfilt,
                                                                           g->bfilt_off, maps);
    } else {
      size_t smem = (size_t)M * 24;
      HCU_CUDA(cudaFuncSetAttribute(cap_fft_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cap_fft_inv_kernel<<<grid, cap_threads(M), smem, ctx->stream>>>(i, M, nside, Z, zstride, g->bfilt, g->bfilt_off,
                                                                      maps);
    }
    HCU_LAUNCH_CHECK(ctx);
    i = ihi + 1;
  }
  return HCU_OK;
}

// split of the cap ring pairs [cap_lo, cap_hi) between the generations: the second generation takes north ring
// numbers 1 .. r2_imax (ring pairs [0, r2_imax)), the first generation what lies above (a_lo..a_hi stays empty)
struct cap_split {
  i64 a_lo, a_hi;  // first generation, tiny rings
  i64 n_lo, n_hi;  // second generation
  i64 b_lo, b_hi;  // first generation, rings beyond one CTA's shared memory
  i64 ystride;     // comp stride of the first generation's workspace (0: none needed)
};
static cap_split split_caps(const hcu_geom *g, i64 cap_lo, i64 cap_hi) {
  cap_split s;
  const i64 imax = g->r2_imax;
  if (imax < 1) {
    s.a_lo = cap_lo, s.a_hi = cap_hi;
    s.n_lo = s.n_hi = s.b_lo = s.b_hi = cap_hi;
  } else {
    s.a_lo = s.a_hi = cap_lo;
    s.n_lo = cap_lo, s.n_hi = std::min<i64>(cap_hi, imax);
    s.b_lo = std::max<i64>(cap_lo, imax), s.b_hi = cap_hi;
  }
  if (s.a_lo > s.a_hi) s.a_hi = s.a_lo;
  if (s.n_lo > s.n_hi) s.n_hi = s.n_lo;
  if (s.b_lo > s.b_hi) s.b_hi = s.b_lo;
  i64 top = 0;
  if (s.a_lo < s.a_hi) top = s.a_hi;
  if (s.b_lo < s.b_hi) top = s.b_hi;
  s.ystride = top ? cap_stride(top) : 0;
  return s;
}

// forward ring FFT stage for ring pairs [rp_lo, rp_hi) of ncomp maps; phase rows follow mlist (nm rows)
int hcu_ring_fft_forward(hcu_ctx *ctx, hcu_geom *g, int lmax, int ncomp,
                         const hcu_ptrs &maps, const double *ring_weights,
                         i64 rp_lo, i64 rp_hi, const int32_t *mlist, int nm, double *phase,
                         const hcu_rowdest *destp) {
  hcu_rowdest dest;
  if (destp) dest = *destp;
  const i64 nside = g->nside;
  const i64 ncap = 2 * nside * (nside - 1);
  const i64 nrp_local = rp_hi - rp_lo;
  const int n4 = (int)(4 * nside);
  const int nk = n4 / 2 + 1;
  const int mthreads = 128;
  const unsigned mblocks = (unsigned)((nm + mthreads - 1) / mthreads);
  if (nm <= 0 || nrp_local <= 0) return HCU_OK;

  // ---- polar caps: ring pairs rp < nside - 1 ---------------------------------
  i64 cap_lo = rp_lo, cap_hi = rp_hi < nside - 1 ? rp_hi : nside - 1;
  if (cap_lo < cap_hi) {
    const cap_split cs = split_caps(g, cap_lo, cap_hi);
    if (cs.ystride) HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_cap, sizeof(double2) * (size_t)cs.ystride * ncomp));
    HCU_CHECK(old_caps_forward(ctx, g, ncomp, maps, ring_weights, rp_lo, nrp_local, cs.b_lo, cs.b_hi, mlist, nm,
                               phase, cs.ystride, dest));
    HCU_CHECK(hcu_ring2_run(ctx, g, false, false, lmax, ncomp, maps, ring_weights, rp_lo, nrp_local, cs.n_lo, cs.n_hi,
                            mlist, nm, nullptr, phase, &dest));
    HCU_CHECK(old_caps_forward(ctx, g, ncomp, maps, ring_weights, rp_lo, nrp_local, cs.a_lo, cs.a_hi, mlist, nm,
                               phase, cs.ystride, dest));
  }

  // ---- equatorial belt: ring pairs rp >= nside - 1 -------------------------------
  i64 belt_lo = rp_lo > nside - 1 ? rp_lo : nside - 1, belt_hi = rp_hi;
  if (belt_lo < belt_hi && g->r2_belt) {
    HCU_CHECK(hcu_ring2_run(ctx, g, false, true, lmax, ncomp, maps, ring_weights, rp_lo, nrp_local, belt_lo, belt_hi,
                            mlist, nm, nullptr, phase, &dest));
  } else if (belt_lo < belt_hi) {
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_belt, sizeof(double2) * (size_t)nk * (2 * nside + 1)));
    double2 *X = (double2 *)ctx->ws_belt.ptr;
    i64 run0[2], cnt[2];
    const int nrun = belt_runs(nside, belt_lo, belt_hi, run0, cnt);
    for (int c = 0; c < ncomp; ++c) {
      for (int r = 0; r < nrun; ++r) {
        cufftHandle plan;
        HCU_CHECK(get_belt_plan(ctx, nside, (int)cnt[r], false, &plan));
        HCU_CUFFT(cufftExecD2Z(plan, maps.p[c] + ncap + run0[r] * n4,
                               reinterpret_cast<cufftDoubleComplex *>(X + run0[r] * nk)));
        ctx->n_cufft++;
      }
      dim3 grid(mblocks, (unsigned)(belt_hi - belt_lo));
      belt_post_kernel<<<grid, mthreads, 0, ctx->stream>>>(
          nside, nm, mlist, ncomp, c, X, ring_weights, rp_lo, nrp_local, belt_lo, phase, dest);
      HCU_LAUNCH_CHECK(ctx);
    }
  }
  return HCU_OK;
}

// inverse ring FFT stage for ring pairs [rp_lo, rp_hi): phase rows are (reN, imN, reS, imS), the row of
// m is mpos[m] (nullptr: m).  Only the pixels of those rings are written.
int hcu_ring_fft_inverse(hcu_ctx *ctx, hcu_geom *g, int lmax, int ncomp,
                         const double *phase, const int32_t *mpos, i64 rp_lo, i64 rp_hi,
                         const hcu_ptrs &maps) {
  const i64 nside = g->nside;
  const i64 ncap = 2 * nside * (nside - 1);
  const i64 nrp_local = rp_hi - rp_lo;
  const int n4 = (int)(4 * nside);
  const int nk = n4 / 2 + 1;
  if (nrp_local <= 0) return HCU_OK;
  // caps
  i64 cap_lo = rp_lo, cap_hi = rp_hi < nside - 1 ? rp_hi : nside - 1;
  if (cap_lo < cap_hi) {
    const cap_split cs = split_caps(g, cap_lo, cap_hi);
    if (cs.ystride) HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_cap, sizeof(double2) * (size_t)cs.ystride * ncomp));
    HCU_CHECK(old_caps_inverse(ctx, g, lmax, ncomp, phase, mpos, rp_lo, nrp_local, cs.b_lo, cs.b_hi, maps, cs.ystride));
    HCU_CHECK(hcu_ring2_run(ctx, g, true, false, lmax, ncomp, maps, nullptr, rp_lo, nrp_local, cs.n_lo, cs.n_hi,
                            nullptr, 0, mpos, const_cast<double *>(phase)));
    HCU_CHECK(old_caps_inverse(ctx, g, lmax, ncomp, phase, mpos, rp_lo, nrp_local, cs.a_lo, cs.a_hi, maps, cs.ystride));
  }
  // belt
  i64 belt_lo = rp_lo > nside - 1 ? rp_lo : nside - 1, belt_hi = rp_hi;
  if (belt_lo < belt_hi && g->r2_belt) {
    HCU_CHECK(hcu_ring2_run(ctx, g, true, true, lmax, ncomp, maps, nullptr, rp_lo, nrp_local, belt_lo, belt_hi, nullptr,
                            0, mpos, const_cast<double *>(phase)));
  } else if (belt_lo < belt_hi) {
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_belt, sizeof(double2) * (size_t)nk * (2 * nside + 1)));
    double2 *X = (double2 *)ctx->ws_belt.ptr;
    i64 run0[2], cnt[2];
    const int nrun = belt_runs(nside, belt_lo, belt_hi, run0, cnt);
    for (int c = 0; c < ncomp; ++c) {
      for (int r = 0; r < nrun; ++r) {
        dim3 grid((unsigned)((nk + 127) / 128), (unsigned)cnt[r]);
        belt_pre_inv_kernel<<<grid, 128, 0, ctx->stream>>>(nside, lmax, mpos, ncomp, c, phase, rp_lo,
                                                           nrp_local, run0[r], X);
        HCU_LAUNCH_CHECK(ctx);
        cufftHandle plan;
        HCU_CHECK(get_belt_plan(ctx, nside, (int)cnt[r], true, &plan));
        HCU_CUFFT(cufftExecZ2D(plan, reinterpret_cast<cufftDoubleComplex *>(X + run0[r] * nk),
                               maps.p[c] + ncap + run0[r] * n4));
        ctx->n_cufft++;
      }
    }
  }
  return HCU_OK;
}
