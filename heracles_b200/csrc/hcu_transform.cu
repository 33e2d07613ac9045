// hcu_transform.cu -- orchestration of the spherical-harmonic transform stages
// (ring FFT <-> Legendre), batching over maps, Jacobi iterations, staging of
// host-resident inputs/outputs.  See include/heracles_cuda.h.
#include <string.h>

#include <algorithm>
#include <vector>

#include "hcu_common.cuh"

int hcu_mul(hcu_ctx *ctx, double *out, const double *a, const double *b, i64 n);
bool hcu_dev_accessible(const void *p);

namespace {

__global__ void almxfl_kernel(hcu_ptrs alm, int nrows, int lmax, const double *fl) {
  const int m = blockIdx.x;
  const i64 base = (i64)m * (2 * lmax + 1 - m) / 2;
  for (int r = blockIdx.y; r < nrows; r += gridDim.y) {
    double2 *row = reinterpret_cast<double2 *>(alm.p[r]);
    for (int l = m + threadIdx.x; l <= lmax; l += blockDim.x) {
      double2 v = row[base + l];
      const double f = fl[l];
      row[base + l] = make_double2(v.x * f, v.y * f);
    }
  }
}

// residual of a Jacobi step: out = w a - b (w == nullptr: a - b)
__global__ void sub_kernel(double *out, const double *a, const double *b, const double *w, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 s = (i64)gridDim.x * blockDim.x;
  if (w)
    for (; i < n; i += s) out[i] = w[i] * a[i] - b[i];
  else
    for (; i < n; i += s) out[i] = a[i] - b[i];
}

bool valid_nside(i64 nside) {
  return nside >= 1 && nside <= (1 << 24) && (nside & (nside - 1)) == 0;
}

int check_sht_args(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps) {
  HCU_ARG(ctx, "ctx");
  HCU_ARG(valid_nside(nside) && nside <= 8192,
          "nside must be a power of two <= 8192 (polar-cap FFT: Bluestein length 2 x 8192)");
  HCU_ARG(lmax >= 0 && lmax <= 4 * nside, "0 <= lmax <= 4 nside");
  if (spin != 0 && spin != 2) {
    hcu_set_error("spin-%d maps not yet supported", spin);
    return HCU_ERR_UNSUPPORTED;
  }
  HCU_ARG(nmaps >= 1, "nmaps >= 1");
  HCU_ARG(spin == 0 || (nmaps % 2) == 0, "spin-2 input needs (Q,U) pairs");
  return HCU_OK;
}

// small host arrays (fl, ring weights) are copied into ws_state; device arrays are used in place
int upload_small(hcu_ctx *ctx, const double *src, size_t n, size_t slot_off, const double **dev) {
  if (!src) {
    *dev = nullptr;
    return HCU_OK;
  }
  if (hcu_dev_accessible(src)) {
    *dev = src;
    return HCU_OK;
  }
  double *d = (double *)ctx->ws_state.ptr + slot_off;
  HCU_CUDA(cudaMemcpyAsync(d, src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  *dev = d;
  return HCU_OK;
}

struct StageTimer {
  hcu_ctx *ctx;
  cudaEvent_t a, b;
  float *acc;
  // only when hcu_set_timing enabled it: collect() waits on the host after every batch
  StageTimer(hcu_ctx *c, int i, float *dst) : ctx(c), a(c->ev[i]), b(c->ev[i + 1]), acc(dst) {
    if (ctx->timing) cudaEventRecord(a, ctx->stream);
  }
  void stop() {
    if (ctx->timing) cudaEventRecord(b, ctx->stream);
  }
  void collect() {
    float t = 0;
    if (ctx->timing && cudaEventSynchronize(b) == cudaSuccess && cudaEventElapsedTime(&t, a, b) == cudaSuccess)
      *acc += t;
  }
};

// one analysis pass over nmaps rows: alm[c] += A(W maps[c])
int analysis_pass(hcu_ctx *ctx, hcu_geom *g, hcu_coef *cf, int lmax, int spin, int nmaps,
                  double *const *maps, const double *rw, const double *pw, const double *fl,
                  double *const *alm) {
  const i64 nside = g->nside;
  const i64 npix = 12 * nside * nside;
  const i64 nrp = g->nrp;
  const int cap = hcu_legendre_batch(spin);
  for (int c0 = 0; c0 < nmaps; c0 += cap) {
    const int nb = std::min(cap, nmaps - c0);
    hcu_ptrs src, dst;
    for (int c = 0; c < HCU_MAX_BATCH; ++c) {
      src.p[c] = c < nb ? maps[c0 + c] : nullptr;
      dst.p[c] = c < nb ? alm[c0 + c] : nullptr;
    }
    if (pw) {
      HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_misc, sizeof(double) * npix * nb));
      for (int c = 0; c < nb; ++c) {
        double *tmp = (double *)ctx->ws_misc.ptr + (i64)c * npix;
        HCU_CHECK(hcu_mul(ctx, tmp, src.p[c], pw, npix));
        src.p[c] = tmp;
      }
    }
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_phase, sizeof(double) * 4 * (size_t)(lmax + 1) * nrp * nb));
    double *phase = (double *)ctx->ws_phase.ptr;
    StageTimer t0(ctx, 0, &ctx->sht_ms[0]);
    HCU_CHECK(hcu_ring_fft_forward(ctx, g, lmax, nb, src, rw, 0, nrp, nullptr, lmax + 1, phase));
    t0.stop();
    StageTimer t1(ctx, 2, &ctx->sht_ms[1]);
    const i64 full[2] = {0, nrp};
    HCU_CHECK(hcu_legendre_analysis(ctx, g, cf, lmax, spin, nb, phase, nullptr, lmax + 1, 1, full, fl, dst));
    t1.stop();
    t0.collect();
    t1.collect();
  }
  return HCU_OK;
}

// maps[c] = S(alm[c])
int synthesis_pass(hcu_ctx *ctx, hcu_geom *g, hcu_coef *cf, int lmax, int spin, int nmaps,
                   double *const *alm, double *const *maps) {
  const i64 nrp = g->nrp;
  const int cap = hcu_legendre_batch(spin);
  for (int c0 = 0; c0 < nmaps; c0 += cap) {
    const int nb = std::min(cap, nmaps - c0);
    hcu_ptrs src, dst;
    for (int c = 0; c < HCU_MAX_BATCH; ++c) {
      src.p[c] = c < nb ? alm[c0 + c] : nullptr;
      dst.p[c] = c < nb ? maps[c0 + c] : nullptr;
    }
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_phase, sizeof(double) * 4 * (size_t)(lmax + 1) * nrp * nb));
    double *phase = (double *)ctx->ws_phase.ptr;
    StageTimer t0(ctx, 0, &ctx->sht_ms[2]);
    const i64 full[2] = {0, nrp};
    HCU_CHECK(hcu_legendre_synthesis(ctx, g, cf, lmax, spin, nb, src, nullptr, lmax + 1, 1, full, phase));
    t0.stop();
    StageTimer t1(ctx, 2, &ctx->sht_ms[3]);
    HCU_CHECK(hcu_ring_fft_inverse(ctx, g, lmax, nb, phase, nullptr, 0, nrp, dst));
    t1.stop();
    t0.collect();
    t1.collect();
  }
  return HCU_OK;
}

}  // namespace

extern "C" int hcu_map2alm_many(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps,
                                const double *const *maps, const double *ring_weights,
                                const double *pixel_weights, int niter, const double *fl,
                                void *const *alm) {
  HCU_CHECK(check_sht_args(ctx, nside, lmax, spin, nmaps));
  HCU_ARG(maps && alm && niter >= 0, "null pointer / niter");
  HCU_CUDA(cudaSetDevice(ctx->device));
  const i64 npix = 12 * nside * nside;
  const i64 nalm = (i64)(lmax + 1) * (lmax + 2) / 2;
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));

  HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_state, sizeof(double) * (size_t)(lmax + 1 + 2 * nside + 16)));
  const double *dfl, *drw;
  HCU_CHECK(upload_small(ctx, fl, lmax + 1, 0, &dfl));
  HCU_CHECK(upload_small(ctx, ring_weights, 2 * nside, lmax + 1, &drw));

  // rows that live on the host are mirrored on the device
  std::vector<double *> dmaps(nmaps), dalm(nmaps);
  int nhost_maps = 0, nhost_alm = 0;
  for (int c = 0; c < nmaps; ++c) {
    HCU_ARG(maps[c] && alm[c], "null row pointer");
    if (!hcu_dev_accessible(maps[c])) ++nhost_maps;
    if (!hcu_dev_accessible(alm[c])) ++nhost_alm;
  }
  if (nhost_maps) HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_map, sizeof(double) * npix * nhost_maps));
  if (nhost_alm) HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_alm, sizeof(double) * 2 * nalm * nhost_alm));
  for (int c = 0, im = 0, ia = 0; c < nmaps; ++c) {
    if (hcu_dev_accessible(maps[c])) {
      dmaps[c] = const_cast<double *>(maps[c]);
    } else {
      dmaps[c] = (double *)ctx->ws_map.ptr + (i64)(im++) * npix;
      HCU_CUDA(cudaMemcpyAsync(dmaps[c], maps[c], sizeof(double) * npix, cudaMemcpyDefault, ctx->stream));
    }
    dalm[c] = hcu_dev_accessible(alm[c]) ? (double *)alm[c]
                                         : (double *)ctx->ws_alm.ptr + 2 * (i64)(ia++) * nalm;
    HCU_CUDA(cudaMemsetAsync(dalm[c], 0, sizeof(double) * 2 * nalm, ctx->stream));
  }
  const double *dpw = pixel_weights;
  // residual and pixel-weight staging live in the context (cudaMalloc / cudaFree of tens of GB per call
  // cost more than a second at nside 4096); hcu_trim releases them
  hcu_buffer &pwbuf = ctx->ws_pw, &resid = ctx->ws_resid;
  if (pixel_weights && !hcu_dev_accessible(pixel_weights)) {
    HCU_CHECK(hcu_ws_reserve(ctx, &pwbuf, sizeof(double) * npix));
    HCU_CUDA(cudaMemcpyAsync(pwbuf.ptr, pixel_weights, sizeof(double) * npix, cudaMemcpyDefault, ctx->stream));
    dpw = (double *)pwbuf.ptr;
  }
  HCU_CUDA(cudaMemsetAsync(ctx->work_counters, 0, 2 * sizeof(double), ctx->stream));
  for (int i = 0; i < 4; ++i) ctx->sht_ms[i] = 0;

  int rc = analysis_pass(ctx, g, cf, lmax, spin, nmaps, dmaps.data(), drw, dpw,
                         niter == 0 ? dfl : nullptr, dalm.data());
  if (rc == HCU_OK && niter > 0) {
    // Jacobi refinement in batches so the residual buffer stays small
    const int cap = hcu_legendre_batch(spin);
    rc = hcu_ws_reserve(ctx, &resid, sizeof(double) * npix * std::min(cap, nmaps));
    for (int c0 = 0; c0 < nmaps && rc == HCU_OK; c0 += cap) {
      const int nb = std::min(cap, nmaps - c0);
      std::vector<double *> r(nb);
      for (int c = 0; c < nb; ++c) r[c] = (double *)resid.ptr + (i64)c * npix;
      for (int it = 0; it < niter && rc == HCU_OK; ++it) {
        rc = synthesis_pass(ctx, g, cf, lmax, spin, nb, dalm.data() + c0, r.data());
        if (rc != HCU_OK) break;
        for (int c = 0; c < nb; ++c) {
          i64 b = std::min<i64>((npix + 255) / 256, (i64)ctx->num_sms * 8);
          // healpy multiplies the map by the pixel weights ONCE and iterates on the weighted map (premultiply):
          // residual = W map - S(alm), analysed with unit pixel weights; per-pass mode re-applies W in every analysis
          sub_kernel<<<(unsigned)b, 256, 0, ctx->stream>>>(r[c], dmaps[c0 + c], r[c],
                                                           ctx->weights_premultiply ? dpw : nullptr, npix);
          ctx->n_launch++;
        }
        rc = analysis_pass(ctx, g, cf, lmax, spin, nb, r.data(), drw, ctx->weights_premultiply ? nullptr : dpw, nullptr,
                           dalm.data() + c0);
      }
      if (rc == HCU_OK && dfl) {
        hcu_ptrs rows;
        for (int c = 0; c < HCU_MAX_BATCH; ++c) rows.p[c] = c < nb ? dalm[c0 + c] : nullptr;
        dim3 grid(lmax + 1, nb);
        almxfl_kernel<<<grid, 128, 0, ctx->stream>>>(rows, nb, lmax, dfl);
        ctx->n_launch++;
      }
    }
  }
  if (rc == HCU_OK)
    for (int c = 0; c < nmaps; ++c)
      if (dalm[c] != (double *)alm[c] &&
          cudaMemcpyAsync(alm[c], dalm[c], sizeof(double) * 2 * nalm, cudaMemcpyDefault, ctx->stream) != cudaSuccess)
        rc = HCU_ERR_CUDA;
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc == HCU_OK && e != cudaSuccess) {
    hcu_set_error("hcu_map2alm: %s", cudaGetErrorString(e));
    rc = HCU_ERR_CUDA;
  }
  return rc;
}

extern "C" int hcu_map2alm(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps,
                           const double *maps, int64_t map_stride,
                           const double *ring_weights, const double *pixel_weights,
                           int niter, const double *fl, void *alm_v, int64_t alm_stride) {
  HCU_ARG(ctx && maps && alm_v && nmaps >= 1, "hcu_map2alm: null pointer");
  HCU_ARG(map_stride >= 12 * nside * nside, "map_stride");
  HCU_ARG(alm_stride >= (int64_t)(lmax + 1) * (lmax + 2) / 2, "alm_stride");
  std::vector<const double *> mp(nmaps);
  std::vector<void *> ap(nmaps);
  for (int c = 0; c < nmaps; ++c) {
    mp[c] = maps + (i64)c * map_stride;
    ap[c] = (double *)alm_v + 2 * (i64)c * alm_stride;
  }
  return hcu_map2alm_many(ctx, nside, lmax, spin, nmaps, mp.data(), ring_weights, pixel_weights,
                          niter, fl, ap.data());
}

extern "C" int hcu_alm2map(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps,
                           const void *alm_v, int64_t alm_stride, double *maps,
                           int64_t map_stride) {
  HCU_CHECK(check_sht_args(ctx, nside, lmax, spin, nmaps));
  HCU_ARG(maps && alm_v, "null pointer");
  HCU_CUDA(cudaSetDevice(ctx->device));
  const i64 npix = 12 * nside * nside;
  const i64 nalm = (i64)(lmax + 1) * (lmax + 2) / 2;
  HCU_ARG(map_stride >= npix && alm_stride >= nalm, "strides");
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  std::vector<double *> dalm(nmaps), dmaps(nmaps);
  const bool host_alm = !hcu_dev_accessible(alm_v), host_maps = !hcu_dev_accessible(maps);
  if (host_alm) HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_alm, sizeof(double) * 2 * nalm * nmaps));
  if (host_maps) HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_map, sizeof(double) * npix * nmaps));
  for (int c = 0; c < nmaps; ++c) {
    const double *src = (const double *)alm_v + 2 * (i64)c * alm_stride;
    if (host_alm) {
      dalm[c] = (double *)ctx->ws_alm.ptr + 2 * (i64)c * nalm;
      HCU_CUDA(cudaMemcpyAsync(dalm[c], src, sizeof(double) * 2 * nalm, cudaMemcpyDefault, ctx->stream));
    } else {
      dalm[c] = const_cast<double *>(src);
    }
    dmaps[c] = host_maps ? (double *)ctx->ws_map.ptr + (i64)c * npix : maps + (i64)c * map_stride;
  }
  ctx->sht_ms[2] = ctx->sht_ms[3] = 0;
  HCU_CHECK(synthesis_pass(ctx, g, cf, lmax, spin, nmaps, dalm.data(), dmaps.data()));
  if (host_maps)
    for (int c = 0; c < nmaps; ++c)
      HCU_CUDA(cudaMemcpyAsync(maps + (i64)c * map_stride, dmaps[c], sizeof(double) * npix,
                               cudaMemcpyDefault, ctx->stream));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  return HCU_OK;
}

// ---- staged transform for the multi-GPU path ---------------------------------------------------
namespace {
int check_stage_args(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp, int64_t rp_lo,
                     int64_t rp_hi, const int32_t *mlist, int nm) {
  HCU_CHECK(check_sht_args(ctx, nside, lmax, spin, ncomp));
  HCU_ARG(ncomp <= hcu_legendre_batch(spin), "at most 12 (spin 0) / 8 (spin 2) components per call");
  HCU_ARG(0 <= rp_lo && rp_lo <= rp_hi && rp_hi <= 2 * nside, "ring pair range");
  HCU_ARG(nm >= 0 && nm <= lmax + 1, "0 <= nm <= lmax + 1");
  HCU_ARG(!mlist || hcu_dev_accessible(mlist), "mlist must be on the device");
  return HCU_OK;
}
}  // namespace

extern "C" int hcu_legendre_batch_size(int spin) { return hcu_legendre_batch(spin); }

extern "C" int hcu_map2phase(hcu_ctx *ctx, int64_t nside, int lmax, int ncomp,
                             const double *maps, int64_t map_stride,
                             const double *ring_weights, int64_t rp_lo, int64_t rp_hi,
                             const int32_t *mlist, int nm, double *phase) {
  HCU_CHECK(check_stage_args(ctx, nside, lmax, 0, ncomp, rp_lo, rp_hi, mlist, nm));
  HCU_ARG(maps && phase, "null pointer");
  HCU_ARG(hcu_dev_accessible(maps) && hcu_dev_accessible(phase), "device pointers required");
  HCU_ARG(!ring_weights || hcu_dev_accessible(ring_weights), "ring_weights must be on the device");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  hcu_ptrs src;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    src.p[c] = c < ncomp ? const_cast<double *>(maps) + (i64)c * map_stride : nullptr;
  return hcu_ring_fft_forward(ctx, g, lmax, ncomp, src, ring_weights, rp_lo, rp_hi, mlist, nm, phase);
}

// The same ring FFTs with the rows of every destination rank written straight into that rank's buffer (peer memory)
extern "C" int hcu_map2phase_peers(hcu_ctx *ctx, int64_t nside, int lmax, int ncomp, const double *maps,
                                   int64_t map_stride, const double *ring_weights, int64_t rp_lo, int64_t rp_hi,
                                   const int32_t *mlist, int nm, int ndest, const int32_t *row_start,
                                   double *const *dest_base) {
  HCU_CHECK(check_stage_args(ctx, nside, lmax, 0, ncomp, rp_lo, rp_hi, mlist, nm));
  HCU_ARG(maps && row_start && dest_base, "null pointer");
  HCU_ARG(ndest >= 1 && ndest <= HCU_MAX_BLOCKS, "1 <= destinations <= 16");
  HCU_ARG(row_start[0] == 0 && row_start[ndest] == nm, "row_start must cover the nm rows");
  HCU_ARG(hcu_dev_accessible(maps), "device pointers required");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  hcu_ptrs src;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    src.p[c] = c < ncomp ? const_cast<double *>(maps) + (i64)c * map_stride : nullptr;
  hcu_rowdest dest;
  dest.nd = ndest;
  for (int d = 0; d < ndest; ++d) {
    HCU_ARG(row_start[d] <= row_start[d + 1], "row_start must ascend");
    HCU_ARG(dest_base[d] || row_start[d] == row_start[d + 1], "null destination");
    dest.row_start[d] = row_start[d];
    dest.base[d] = dest_base[d];
  }
  dest.row_start[ndest] = row_start[ndest];
  return hcu_ring_fft_forward(ctx, g, lmax, ncomp, src, ring_weights, rp_lo, rp_hi, mlist, nm, nullptr, &dest);
}

extern "C" int hcu_phase2alm(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                             const double *phase, const int32_t *mlist, int nm,
                             int64_t rp_lo, int64_t rp_hi, const double *fl, void *alm,
                             int64_t alm_stride) {
  HCU_CHECK(check_stage_args(ctx, nside, lmax, spin, ncomp, rp_lo, rp_hi, mlist, nm));
  HCU_ARG(phase && alm, "null pointer");
  HCU_ARG(hcu_dev_accessible(phase) && hcu_dev_accessible(alm), "device pointers required");
  HCU_ARG(!fl || hcu_dev_accessible(fl), "fl must be on the device");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  hcu_ptrs rows;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    rows.p[c] = c < ncomp ? (double *)alm + 2 * (i64)c * alm_stride : nullptr;
  const i64 one[2] = {rp_lo, rp_hi};
  return hcu_legendre_analysis(ctx, g, cf, lmax, spin, ncomp, phase, mlist, nm, 1, one, fl, rows);
}

extern "C" int hcu_phase2alm_blocks(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                                    const double *phase, const int32_t *mlist, int nm, int nblocks,
                                    const int64_t *rp_bounds, const double *fl, void *alm,
                                    int64_t alm_stride) {
  HCU_ARG(rp_bounds && nblocks >= 1 && nblocks <= 16, "1 <= nblocks <= 16");
  for (int b = 0; b < nblocks; ++b) HCU_ARG(rp_bounds[b] <= rp_bounds[b + 1], "rp_bounds must ascend");
  HCU_CHECK(check_stage_args(ctx, nside, lmax, spin, ncomp, rp_bounds[0], rp_bounds[nblocks], mlist, nm));
  HCU_ARG(phase && alm, "null pointer");
  HCU_ARG(hcu_dev_accessible(phase) && hcu_dev_accessible(alm), "device pointers required");
  HCU_ARG(!fl || hcu_dev_accessible(fl), "fl must be on the device");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  hcu_ptrs rows;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    rows.p[c] = c < ncomp ? (double *)alm + 2 * (i64)c * alm_stride : nullptr;
  return hcu_legendre_analysis(ctx, g, cf, lmax, spin, ncomp, phase, mlist, nm, nblocks,
                               (const i64 *)rp_bounds, fl, rows);
}

extern "C" int hcu_alm2phase(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                             const void *alm, int64_t alm_stride, const int32_t *mlist, int nm,
                             int64_t rp_lo, int64_t rp_hi, double *phase) {
  HCU_CHECK(check_stage_args(ctx, nside, lmax, spin, ncomp, rp_lo, rp_hi, mlist, nm));
  HCU_ARG(phase && alm, "null pointer");
  HCU_ARG(hcu_dev_accessible(phase) && hcu_dev_accessible(alm), "device pointers required");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  hcu_ptrs rows;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    rows.p[c] = c < ncomp ? (double *)alm + 2 * (i64)c * alm_stride : nullptr;
  const i64 one[2] = {rp_lo, rp_hi};
  return hcu_legendre_synthesis(ctx, g, cf, lmax, spin, ncomp, rows, mlist, nm, 1, one, phase);
}

extern "C" int hcu_alm2phase_blocks(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                                    const void *alm, int64_t alm_stride, const int32_t *mlist, int nm,
                                    int nblocks, const int64_t *rp_bounds, double *phase) {
  HCU_ARG(rp_bounds && nblocks >= 1 && nblocks <= 16, "1 <= nblocks <= 16");
  for (int b = 0; b < nblocks; ++b) HCU_ARG(rp_bounds[b] <= rp_bounds[b + 1], "rp_bounds must ascend");
  HCU_CHECK(check_stage_args(ctx, nside, lmax, spin, ncomp, rp_bounds[0], rp_bounds[nblocks], mlist, nm));
  HCU_ARG(phase && alm, "null pointer");
  HCU_ARG(hcu_dev_accessible(phase) && hcu_dev_accessible(alm), "device pointers required");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  hcu_ptrs rows;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    rows.p[c] = c < ncomp ? (double *)alm + 2 * (i64)c * alm_stride : nullptr;
  return hcu_legendre_synthesis(ctx, g, cf, lmax, spin, ncomp, rows, mlist, nm, nblocks,
                                (const i64 *)rp_bounds, phase);
}

// hcu_alm2phase_blocks with block b written to block_out[b] (the buffer of the rank that owns those ring pairs)
extern "C" int hcu_alm2phase_peers(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp, const void *alm,
                                   int64_t alm_stride, const int32_t *mlist, int nm, int nblocks,
                                   const int64_t *rp_bounds, double *const *block_out) {
  HCU_ARG(rp_bounds && block_out && nblocks >= 1 && nblocks <= 16, "1 <= nblocks <= 16");
  for (int b = 0; b < nblocks; ++b) {
    HCU_ARG(rp_bounds[b] <= rp_bounds[b + 1], "rp_bounds must ascend");
    HCU_ARG(block_out[b] || rp_bounds[b] == rp_bounds[b + 1], "null block destination");
  }
  HCU_CHECK(check_stage_args(ctx, nside, lmax, spin, ncomp, rp_bounds[0], rp_bounds[nblocks], mlist, nm));
  HCU_ARG(alm && hcu_dev_accessible(alm), "device pointers required");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  hcu_coef *cf;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  hcu_ptrs rows;
  for (int c = 0; c < HCU_MAX_BATCH; ++c)
    rows.p[c] = c < ncomp ? (double *)alm + 2 * (i64)c * alm_stride : nullptr;
  return hcu_legendre_synthesis(ctx, g, cf, lmax, spin, ncomp, rows, mlist, nm, nblocks, (const i64 *)rp_bounds,
                                nullptr, block_out);
}

// ---- peer memory: export a cudaMalloc'ed buffer to the other ranks of the node, open theirs ----------------------
extern "C" int hcu_ipc_export(hcu_ctx *ctx, const void *ptr, void *handle64) {
  HCU_ARG(ctx && ptr && handle64, "null pointer");
  HCU_CUDA(cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  HCU_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(ptr)));
  memcpy(handle64, &h, 64);
  return HCU_OK;
}
extern "C" int hcu_ipc_open(hcu_ctx *ctx, const void *handle64, void **ptr) {
  HCU_ARG(ctx && ptr && handle64, "null pointer");
  HCU_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  HCU_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HCU_OK;
}
extern "C" int hcu_ipc_close(hcu_ctx *ctx, void *ptr) {
  HCU_ARG(ctx && ptr, "null pointer");
  HCU_CUDA(cudaSetDevice(ctx->device));
  HCU_CUDA(cudaIpcCloseMemHandle(ptr));
  return HCU_OK;
}

extern "C" int hcu_phase2map(hcu_ctx *ctx, int64_t nside, int lmax, int ncomp, const double *phase,
                             const int32_t *mpos, int64_t rp_lo, int64_t rp_hi, double *maps,
                             int64_t map_stride) {
  HCU_CHECK(check_stage_args(ctx, nside, lmax, 0, ncomp, rp_lo, rp_hi, mpos, 0));
  HCU_ARG(maps && phase, "null pointer");
  HCU_ARG(hcu_dev_accessible(maps) && hcu_dev_accessible(phase), "device pointers required");
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_geom *g;
  HCU_CHECK(hcu_get_geom(ctx, nside, &g));
  hcu_ptrs dst;
  for (int c = 0; c < HCU_MAX_BATCH; ++c) dst.p[c] = c < ncomp ? maps + (i64)c * map_stride : nullptr;
  return hcu_ring_fft_inverse(ctx, g, lmax, ncomp, phase, mpos, rp_lo, rp_hi, dst);
}

// ---- N4: catalogue -> alm without pixels (heracles/ducc.py:92-133) ------------------------------------------------
// ducc0.sht.adjoint_synthesis_general(map=values, spin, lmax, loc): alm_lm = sum_i v_i conj(sY_lm(theta_i, phi_i)).
// Every point is a "ring" of its own with one sample: its ring Fourier coefficients are v exp(-i m phi), its
// colatitude is free, and the Legendre analysis kernels (which only see cos / sin tables per ring pair) sum the
// harmonics exactly -- O(points x lmax^2), no NUFFT: meant for the catalogue sizes the discrete mapper is used with
// (examples/discrete.ipynb), not for 1e9 rows.
namespace {
__global__ void point_geom_kernel(i64 n, const double *lon, const double *lat, double *cth, double *sth, double *ch,
                                  double *sh, double *phi) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double deg = 0.017453292519943295769;
  const double theta = (90.0 - lat[i]) * deg;
  double l = fmod(lon[i], 360.0);
  if (l < 0.0) l += 360.0;  // numpy's lon % 360.0
  phi[i] = l * deg;
  double s, c;
  sincos(theta, &s, &c);
  cth[i] = c;
  sth[i] = s;
  sincos(0.5 * theta, &s, &c);
  ch[i] = c;
  sh[i] = s;
}

// phase[((m * P + pt) * ncomp + c) * 4] = v_c exp(-i m phi) as (north + south, north - south) with an empty south
__global__ void point_phase_kernel(int P, int ncomp, const double *phi, const double *values, i64 vstride, double *phase) {
  const int pt = blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= P) return;
  const int m = blockIdx.y;
  double s, c;
  sincos((double)m * phi[pt], &s, &c);
  double4 *out = reinterpret_cast<double4 *>(phase + (((i64)m * P + pt) * ncomp) * 4);
  for (int k = 0; k < ncomp; ++k) {
    const double v = values[(i64)k * vstride + pt];
    out[k] = make_double4(v * c, -v * s, v * c, -v * s);
  }
}
}  // namespace

extern "C" int hcu_points2alm(hcu_ctx *ctx, int lmax, int spin, int ncomp, int64_t npts, const double *lon,
                              const double *lat, const double *values, int64_t value_stride, void *alm,
                              int64_t alm_stride) {
  HCU_ARG(ctx && lon && lat && values && alm, "null pointer");
  HCU_ARG(lmax >= 0 && lmax <= 32768, "0 <= lmax <= 32768");
  HCU_ARG(npts >= 0, "npts >= 0");
  if (spin != 0 && spin != 2) {
    hcu_set_error("spin-%d values not yet supported", spin);
    return HCU_ERR_UNSUPPORTED;
  }
  HCU_ARG(ncomp >= 1 && ncomp <= hcu_legendre_batch(spin), "at most 12 (spin 0) / 8 (spin 2) value rows per call");
  HCU_ARG(spin == 0 || (ncomp % 2) == 0, "spin-2 input needs (Q, U) pairs");
  HCU_ARG(hcu_dev_accessible(alm), "alm must be device accessible (Mapper.create())");
  if (npts == 0) return HCU_OK;
  HCU_CUDA(cudaSetDevice(ctx->device));
  hcu_coef *cf;
  HCU_CHECK(hcu_get_coef(ctx, lmax, spin, &cf));
  // device copies of the columns + the per-point geometry
  HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_map, sizeof(double) * (size_t)npts * (size_t)(7 + ncomp)));
  double *base = (double *)ctx->ws_map.ptr;
  double *dlon = base, *dlat = base + npts, *cth = base + 2 * npts, *sth = base + 3 * npts, *ch = base + 4 * npts,
         *sh = base + 5 * npts, *phi = base + 6 * npts, *dval = base + 7 * npts;
  HCU_CUDA(cudaMemcpyAsync(dlon, lon, sizeof(double) * npts, cudaMemcpyDefault, ctx->stream));
  HCU_CUDA(cudaMemcpyAsync(dlat, lat, sizeof(double) * npts, cudaMemcpyDefault, ctx->stream));
  for (int k = 0; k < ncomp; ++k)
    HCU_CUDA(cudaMemcpyAsync(dval + (i64)k * npts, values + (i64)k * value_stride, sizeof(double) * npts,
                             cudaMemcpyDefault, ctx->stream));
  point_geom_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, ctx->stream>>>(npts, dlon, dlat, cth, sth, ch, sh, phi);
  HCU_LAUNCH_CHECK(ctx);
  hcu_ptrs rows;
  for (int c = 0; c < HCU_MAX_BATCH; ++c) rows.p[c] = c < ncomp ? (double *)alm + 2 * (i64)c * alm_stride : nullptr;
  // chunks of points whose phase array stays below ~2 GB
  i64 P = (i64)(2.0e9 / (32.0 * ncomp * (lmax + 1)));
  P = std::max<i64>(256, (P / 256) * 256);
  for (i64 p0 = 0; p0 < npts; p0 += P) {
    const int n = (int)std::min<i64>(P, npts - p0);
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_phase, sizeof(double) * 4 * (size_t)(lmax + 1) * (size_t)n * ncomp));
    double *phase = (double *)ctx->ws_phase.ptr;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)(lmax + 1));
    point_phase_kernel<<<grid, 128, 0, ctx->stream>>>(n, ncomp, phi + p0, dval + p0, npts, phase);
    HCU_LAUNCH_CHECK(ctx);
    hcu_geom g;  // a geometry of its own: the points of this chunk as ring pairs without a southern ring
    g.nside = 0;
    g.nrp = n;
    g.cth = cth + p0;
    g.sth = sth + p0;
    g.ch = ch + p0;
    g.sh = sh + p0;
    const i64 one[2] = {0, n};
    HCU_CHECK(hcu_legendre_analysis(ctx, &g, cf, lmax, spin, ncomp, phase, nullptr, lmax + 1, 1, one, nullptr, rows));
  }
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  return HCU_OK;
}
