// k_mapvalues.cu -- fused ang2pix (RING / NEST) + weighted scatter-add.
//
// Replaces hp.ang2pix(nside, lon, lat, lonlat=True) + numba `_map`
// (heracles/healpy.py:157-160, :58-65) and FootprintFilter's ang2pix
// (heracles/catalog/filters.py:91-94).
//
// This translation unit is compiled with -fmad=false: the pixel index must be
// bit-identical to the scalar C arithmetic healpy performs (no contraction of
// a*b+c into an FMA), see the HEALPix loc2pix algorithm restated in
// oracle/healpix_oracle.c.
//
// HBM-bound integer/byte work: one row per thread, coalesced 8-byte column
// loads, RED.ADD.F64 to the map.  Algorithmic bytes per row: 16 (lon, lat)
// + 8 nv (values) + 16 nv (read-modify-write of the map cell).
#include "hcu_common.cuh"

namespace {

__device__ __forceinline__ double fmodulo(double v1, double v2) {
  if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
  double tmp = fmod(v1, v2) + v2;
  return (tmp == v2) ? 0. : tmp;
}

__device__ __forceinline__ i64 spread_bits(i64 v) {
  unsigned long long x = (unsigned long long)v & 0xffffffffull;
  x = (x | (x << 16)) & 0x0000ffff0000ffffull;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return (i64)x;
}

__device__ __forceinline__ i64 xyf2nest(i64 ix, i64 iy, int face, int order) {
  return ((i64)face << (2 * order)) + spread_bits(ix) + (spread_bits(iy) << 1);
}

template <int NEST>
__device__ __forceinline__ i64 ang2pix_lonlat(i64 nside, int order, double lon,
                                              double lat) {
  const double PI = 3.141592653589793238462643383279502884197;
  const double HALFPI = 1.570796326794896619231321691639751442099;
  const double INV_HALFPI = 0.6366197723675813430755350534900574;
  const double TWOTHIRD = 2.0 / 3.0;
  const double DEG2RAD = 3.141592653589793238462643383279502884 / 180.0;
  double theta = HALFPI - lat * DEG2RAD;
  double phi = lon * DEG2RAD;
  if (!(theta >= 0 && theta <= PI) || !(fabs(phi) <= 1.7976931348623157e308))
    return -1;
  bool have_sth = (theta < 0.01) || (theta > 3.14159 - 0.01);
  double z = cos(theta);
  double sth = have_sth ? sin(theta) : 0.0;
  double za = fabs(z);
  double tt = fmodulo(phi * INV_HALFPI, 4.0);
  const i64 npix = 12 * nside * nside;
  const i64 ncap = 2 * nside * (nside - 1);
  if (za <= TWOTHIRD) {
    double temp1 = nside * (0.5 + tt);
    double temp2 = NEST ? nside * (z * 0.75) : nside * z * 0.75;
    i64 jp = (i64)(temp1 - temp2);
    i64 jm = (i64)(temp1 + temp2);
    if (NEST) {
      i64 ifp = jp >> order, ifm = jm >> order;
      int face = (ifp == ifm) ? (int)(ifp | 4)
                              : ((ifp < ifm) ? (int)ifp : (int)(ifm + 8));
      i64 ix = jm & (nside - 1);
      i64 iy = nside - (jp & (nside - 1)) - 1;
      return xyf2nest(ix, iy, face, order);
    } else {
      i64 nl4 = 4 * nside;
      i64 ir = nside + 1 + jp - jm;
      i64 kshift = 1 - (ir & 1);
      i64 t1 = jp + jm - nside + kshift + 1 + nl4 + nl4;
      i64 ip = (t1 >> 1) & (nl4 - 1);
      return ncap + (ir - 1) * nl4 + ip;
    }
  } else {
    double tmp = ((za < 0.99) || (!have_sth)) ? nside * sqrt(3 * (1 - za))
                                              : nside * sth / sqrt((1. + za) / 3.);
    if (NEST) {
      int ntt = (int)tt;
      if (ntt > 3) ntt = 3;
      double tp = tt - ntt;
      i64 jp = (i64)(tp * tmp);
      i64 jm = (i64)((1.0 - tp) * tmp);
      if (jp > nside - 1) jp = nside - 1;
      if (jm > nside - 1) jm = nside - 1;
      return (z >= 0) ? xyf2nest(nside - jm - 1, nside - jp - 1, ntt, order)
                      : xyf2nest(jp, jm, ntt + 8, order);
    } else {
      double tp = tt - (i64)tt;
      i64 jp = (i64)(tp * tmp);
      i64 jm = (i64)((1.0 - tp) * tmp);
      i64 ir = jp + jm + 1;
      i64 ip = (i64)(tt * ir);
      if (ip >= 4 * ir) ip = 4 * ir - 1;
      return (z > 0) ? 2 * ir * (ir - 1) + ip : npix - 2 * ir * (ir + 1) + ip;
    }
  }
}

// streaming loads: catalogue columns are read exactly once
__device__ __forceinline__ double ld_stream(const double *p) {
  return __ldcs(p);
}

template <int NEST, int NV, int AGG>
__global__ void __launch_bounds__(256)
map_values_kernel(i64 nside, int order, const double *__restrict__ lon,
                  const double *__restrict__ lat,
                  const double *__restrict__ values, i64 vstride, i64 n,
                  double *__restrict__ maps, i64 mstride,
                  unsigned long long *bad_rows, i64 *__restrict__ ipix_out) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  // the trip count is made warp-uniform so that the aggregating variant can
  // use full-mask warp primitives
  const i64 nround = (n + 31) & ~(i64)31;
  for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += stride) {
    const bool inb = j < n;
    i64 pix = -1;
    double v[NV > 0 ? NV : 1];
    if (inb) {
      double lo = ld_stream(lon + j), la = ld_stream(lat + j);
      pix = ang2pix_lonlat<NEST>(nside, order, lo, la);
      if (NV > 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = ld_stream(values + k * vstride + j);
      }
      if (NV > 0 && pix < 0) ++bad;
    }
    if (NV == 0) {
      if (inb) ipix_out[j] = pix;
      continue;
    }
    if (AGG) {
      // warp aggregation: rows of one warp that hit the same pixel are summed
      // by the lowest lane of the group; one atomic per distinct pixel
      unsigned peers = __match_any_sync(0xffffffffu, pix);
      int lane = threadIdx.x & 31;
      int leader = __ffs(peers) - 1;
      if (__popc(peers) > 1) {
        // segmented reduction over the peer group (groups are tiny in practice)
        unsigned rest = peers & ~(1u << leader);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          double acc = v[k];
          unsigned r = rest;
          while (__any_sync(peers, r != 0)) {
            int src = r ? (__ffs(r) - 1) : lane;
            double o = __shfl_sync(peers, v[k], src);
            if (r) acc += o;
            r &= r - 1;
          }
          v[k] = acc;
        }
      }
      if (lane == leader && pix >= 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) atomicAdd(maps + k * mstride + pix, v[k]);
      }
    } else {
      if (pix >= 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) atomicAdd(maps + k * mstride + pix, v[k]);
      }
    }
  }
  if (bad) atomicAdd(bad_rows, bad);
}

// ---------------------------------------------------------------------------
// one catalogue page -> position map AND shear map in one pass (hcu_map_page):
// the Positions and the Shears field of a tomographic bin read the same lon / lat / weight
// columns (heracles/fields.py:262-271 and :420-433), so ang2pix runs once and five columns
// instead of seven cross PCIe.  The running sums the Field layer keeps per page (ngal, wmean,
// w2mean, var) are reduced on the device: per thread over its rows, then per warp, then one
// atomicAdd per warp and statistic.
// ---------------------------------------------------------------------------
template <int NEST>
__global__ void __launch_bounds__(256)
map_page_kernel(i64 nside, int order, const double *__restrict__ lon, const double *__restrict__ lat,
                const double *__restrict__ w, const double *__restrict__ g1, const double *__restrict__ g2,
                i64 n, double *__restrict__ pos, double *__restrict__ she, i64 she_stride,
                double *__restrict__ stats, unsigned long long *bad_rows) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  double st[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) st[k] = 0.0;
  unsigned long long bad = 0;
  for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const double lo = ld_stream(lon + j), la = ld_stream(lat + j);
    const double wj = w ? ld_stream(w + j) : 1.0;
    double a = 0.0, b = 0.0;
    if (she) {
      a = ld_stream(g1 + j);
      b = ld_stream(g2 + j);
    }
    // CatalogPage.get raises on NaN in a requested column (catalog/base.py:114-125)
    if (lo != lo || la != la || wj != wj || a != a || b != b) {
      st[7] += 1.0;
      continue;
    }
    const i64 pix = ang2pix_lonlat<NEST>(nside, order, lo, la);
    if (pix < 0) {
      ++bad;
      continue;
    }
    if (pos) {
      atomicAdd(pos + pix, wj);
      st[0] += 1.0;
      st[1] += wj;
      st[2] += wj * wj;
    }
    if (she && wj != 0.0) {  // page.delete(page[wcol] == 0), fields.py:420-421
      const double re = wj * a, im = wj * b;  // fields.py:426
      atomicAdd(she + pix, re);
      atomicAdd(she + she_stride + pix, im);
      st[3] += 1.0;
      st[4] += wj;
      st[5] += wj * wj;
      st[6] += re * re + im * im;
    }
  }
  if (stats) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double v = st[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(stats + k, v);
    }
  }
  if (bad) atomicAdd(bad_rows, bad);
}

template <int NEST, int NV>
int launch2(hcu_ctx *ctx, i64 nside, const double *lon, const double *lat,
            const double *values, i64 vstride, i64 n, double *maps, i64 mstride,
            int flags, i64 *ipix_out) {
  int order = ilog2_host(nside);
  i64 blocks = (n + 255) / 256;
  // grid = a multiple of the SM count (8 resident 256-thread CTAs per SM), grid-stride loop
  i64 maxb = (i64)ctx->num_sms * 8;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  if (flags & HCU_MAP_AGGREGATE)
    map_values_kernel<NEST, NV, 1><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
        nside, order, lon, lat, values, vstride, n, maps, mstride, ctx->bad_rows, ipix_out);
  else
    map_values_kernel<NEST, NV, 0><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
        nside, order, lon, lat, values, vstride, n, maps, mstride, ctx->bad_rows, ipix_out);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

template <int NEST>
int launch1(hcu_ctx *ctx, i64 nside, const double *lon, const double *lat,
            const double *values, i64 vstride, int nv, i64 n, double *maps,
            i64 mstride, int flags, i64 *ipix_out) {
  switch (nv) {
    case 0: return launch2<NEST, 0>(ctx, nside, lon, lat, values, vstride, n, maps, mstride, 0, ipix_out);
    case 1: return launch2<NEST, 1>(ctx, nside, lon, lat, values, vstride, n, maps, mstride, flags, ipix_out);
    case 2: return launch2<NEST, 2>(ctx, nside, lon, lat, values, vstride, n, maps, mstride, flags, ipix_out);
    case 3: return launch2<NEST, 3>(ctx, nside, lon, lat, values, vstride, n, maps, mstride, flags, ipix_out);
    case 4: return launch2<NEST, 4>(ctx, nside, lon, lat, values, vstride, n, maps, mstride, flags, ipix_out);
  }
  hcu_set_error("map_values: nv=%d not in 0..4", nv);
  return HCU_ERR_ARG;
}

}  // namespace

// all pointers are device-accessible here; nv == 0 writes pixel indices to ipix_out
int hcu_launch_map_values(hcu_ctx *ctx, i64 nside, int scheme, const double *lon,
                          const double *lat, const double *values, i64 vstride,
                          int nv, i64 n, double *maps, i64 mstride, int flags,
                          i64 *ipix_out) {
  if (n <= 0) return HCU_OK;
  if (scheme == HCU_NEST)
    return launch1<1>(ctx, nside, lon, lat, values, vstride, nv, n, maps, mstride, flags, ipix_out);
  return launch1<0>(ctx, nside, lon, lat, values, vstride, nv, n, maps, mstride, flags, ipix_out);
}

// ---------------------------------------------------------------------------
// elementwise map arithmetic (Field-layer normalisation on the device)
// ---------------------------------------------------------------------------
namespace {
__global__ void scale_kernel(double *x, i64 n, double a) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 s = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += s) x[i] *= a;
}
__global__ void divide_kernel(double *x, i64 n, double a) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 s = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += s) x[i] /= a;
}
__global__ void add_scalar_kernel(double *x, i64 n, double a) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 s = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += s) x[i] += a;
}
__global__ void axpy_kernel(double *y, const double *x, double a, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 s = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += s) y[i] += a * x[i];
}
// out = (region map == k) ? in : 0 -- the jackknife region mask of DICES (heracles/dices/jackknife.py: _get_region_maps)
__global__ void region_select_kernel(double *out, const double *in, const double *jk, double k, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 s = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += s) out[i] = (jk[i] == k) ? in[i] : 0.0;
}
__global__ void mul_kernel(double *out, const double *a, const double *b, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 s = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += s) out[i] = a[i] * b[i];
}
// ud_grade between RING maps: each output pixel averages (degrade) or copies
// (upgrade) its NEST children / parent.
__device__ __constant__ int c_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__device__ __constant__ int c_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

__device__ i64 isqrt_dev(i64 v) {
  i64 r = (i64)sqrt((double)v + 0.5);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

__device__ void ring2xyf_dev(i64 nside, i64 pix, i64 *ix, i64 *iy, int *face) {
  const i64 npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1);
  const i64 nl2 = 2 * nside;
  i64 iring, iphi, kshift, nr;
  if (pix < ncap) {
    iring = (1 + isqrt_dev(1 + 2 * pix)) >> 1;
    iphi = (pix + 1) - 2 * iring * (iring - 1);
    kshift = 0;
    nr = iring;
    *face = (int)((iphi - 1) / nr);
  } else if (pix < (npix - ncap)) {
    i64 ip = pix - ncap;
    i64 tmp = ip / (4 * nside);
    iring = tmp + nside;
    iphi = ip - tmp * 4 * nside + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    i64 ire = tmp + 1, irm = nl2 + 1 - tmp;
    i64 ifm = (iphi - ire / 2 + nside - 1) / nside;
    i64 ifp = (iphi - irm / 2 + nside - 1) / nside;
    *face = (ifp == ifm) ? (int)(ifp | 4) : ((ifp < ifm) ? (int)ifp : (int)(ifm + 8));
  } else {
    i64 ip = npix - pix;
    iring = (1 + isqrt_dev(2 * ip - 1)) >> 1;
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
    kshift = 0;
    nr = iring;
    iring = 2 * nl2 - iring;
    *face = 8 + (int)((iphi - 1) / nr);
  }
  i64 irt = iring - ((i64)c_jrll[*face] * nside) + 1;
  i64 ipt = 2 * iphi - (i64)c_jpll[*face] * nr - kshift - 1;
  if (ipt >= nl2) ipt -= 8 * nside;
  *ix = (ipt - irt) >> 1;
  *iy = (-ipt - irt) >> 1;
}

__device__ i64 xyf2ring_dev(i64 nside, i64 ix, i64 iy, int face) {
  const i64 npix = 12 * nside * nside, ncap = 2 * nside * (nside - 1);
  i64 nl4 = 4 * nside;
  i64 jr = ((i64)c_jrll[face] * nside) - ix - iy - 1;
  i64 nr, kshift, n_before;
  if (jr < nside) {
    nr = jr;
    n_before = 2 * nr * (nr - 1);
    kshift = 0;
  } else if (jr > 3 * nside) {
    nr = nl4 - jr;
    n_before = npix - 2 * (nr + 1) * nr;
    kshift = 0;
  } else {
    nr = nside;
    n_before = ncap + (jr - nside) * nl4;
    kshift = (jr - nside) & 1;
  }
  i64 jp = ((i64)c_jpll[face] * nr + ix - iy + 1 + kshift) / 2;
  if (jp > nl4)
    jp -= nl4;
  else if (jp < 1)
    jp += nl4;
  return n_before + jp - 1;
}

// RING <-> NEST reordering of a whole map (hp.reorder): the transform works on RING maps
__global__ void reorder_kernel(i64 nside, int order, const double *__restrict__ in, double *__restrict__ out, int to_nest) {
  const i64 npix = 12 * nside * nside;
  const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;  // RING index
  if (p >= npix) return;
  i64 ix, iy;
  int face;
  ring2xyf_dev(nside, p, &ix, &iy, &face);
  const i64 q = xyf2nest(ix, iy, face, order);
  if (to_nest)
    out[q] = in[p];
  else
    out[p] = in[q];
}

__global__ void ud_grade_kernel(i64 nside_in, const double *in, i64 nside_out,
                                double *out) {
  const i64 npix_out = 12 * nside_out * nside_out;
  i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix_out) return;
  i64 ix, iy;
  int face;
  ring2xyf_dev(nside_out, p, &ix, &iy, &face);
  if (nside_in >= nside_out) {
    i64 f = nside_in / nside_out;
    double s = 0;
    for (i64 dy = 0; dy < f; ++dy)
      for (i64 dx = 0; dx < f; ++dx)
        s += in[xyf2ring_dev(nside_in, ix * f + dx, iy * f + dy, face)];
    out[p] = s / (double)(f * f);
  } else {
    i64 f = nside_out / nside_in;
    out[p] = in[xyf2ring_dev(nside_in, ix / f, iy / f, face)];
  }
}
}  // namespace

static unsigned ew_blocks(hcu_ctx *ctx, i64 n) {
  i64 b = (n + 255) / 256;
  i64 maxb = (i64)ctx->num_sms * 8;
  if (b > maxb) b = maxb;
  if (b < 1) b = 1;
  return (unsigned)b;
}

extern "C" int hcu_scale(hcu_ctx *ctx, double *x, int64_t n, double a) {
  HCU_ARG(ctx && x && n >= 0, "hcu_scale");
  if (n == 0) return HCU_OK;
  scale_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(x, n, a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

extern "C" int hcu_divide(hcu_ctx *ctx, double *x, int64_t n, double a) {
  HCU_ARG(ctx && x && n >= 0, "hcu_divide");
  if (n == 0) return HCU_OK;
  divide_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(x, n, a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

extern "C" int hcu_add_scalar(hcu_ctx *ctx, double *x, int64_t n, double a) {
  HCU_ARG(ctx && x && n >= 0, "hcu_add_scalar");
  if (n == 0) return HCU_OK;
  add_scalar_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(x, n, a);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

extern "C" int hcu_axpy(hcu_ctx *ctx, double *y, const double *x, double a, int64_t n) {
  HCU_ARG(ctx && x && y && n >= 0, "hcu_axpy");
  if (n == 0) return HCU_OK;
  axpy_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(y, x, a, n);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

int hcu_mul(hcu_ctx *ctx, double *out, const double *a, const double *b, i64 n) {
  if (n == 0) return HCU_OK;
  mul_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(out, a, b, n);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

int hcu_ud_grade_dev(hcu_ctx *ctx, i64 nside_in, const double *in, i64 nside_out, double *out) {
  i64 npix = 12 * nside_out * nside_out;
  ud_grade_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(nside_in, in, nside_out, out);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

// all pointers device-accessible; see hcu_map_page
int hcu_launch_map_page(hcu_ctx *ctx, i64 nside, int scheme, const double *lon, const double *lat, const double *w,
                        const double *g1, const double *g2, i64 n, double *pos, double *she, i64 she_stride,
                        double *stats) {
  if (n <= 0) return HCU_OK;
  const int order = ilog2_host(nside);
  i64 blocks = (n + 255) / 256;
  const i64 maxb = (i64)ctx->num_sms * 8;
  if (blocks > maxb) blocks = maxb;
  if (scheme == HCU_NEST)
    map_page_kernel<1><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nside, order, lon, lat, w, g1, g2, n, pos, she,
                                                                  she_stride, stats, ctx->bad_rows);
  else
    map_page_kernel<0><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nside, order, lon, lat, w, g1, g2, n, pos, she,
                                                                  she_stride, stats, ctx->bad_rows);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

extern "C" int hcu_reorder(hcu_ctx *ctx, int64_t nside, const double *in, double *out, int to_nest) {
  HCU_ARG(ctx && in && out && in != out, "hcu_reorder: null pointer / in-place");
  HCU_ARG(nside >= 1 && (nside & (nside - 1)) == 0, "nside must be a power of two");
  const i64 npix = 12 * nside * nside;
  reorder_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(nside, ilog2_host(nside), in, out, to_nest);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

extern "C" int hcu_region_select(hcu_ctx *ctx, double *out, const double *in, const double *jk_map, double region, int64_t n) {
  HCU_ARG(ctx && out && in && jk_map && n >= 0, "hcu_region_select");
  if (n == 0) return HCU_OK;
  region_select_kernel<<<ew_blocks(ctx, n), 256, 0, ctx->stream>>>(out, in, jk_map, region, n);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}
