// hcu_api.cu -- C ABI entry points, context, memory, host staging and the
// orchestration of the transform stages.  See include/heracles_cuda.h.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include <stdlib.h>

#include "hcu_common.cuh"

int hcu_mul(hcu_ctx *ctx, double *out, const double *a, const double *b, i64 n);
int hcu_ud_grade_dev(hcu_ctx *ctx, i64 nside_in, const double *in, i64 nside_out, double *out);

static thread_local char g_err[1024] = "";

void hcu_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *hcu_last_error(void) { return g_err; }
extern "C" int hcu_version(void) { return HCU_VERSION; }

extern "C" int hcu_device_count(int *count) {
  HCU_ARG(count, "count");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return HCU_OK;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int hcu_create(int device, hcu_ctx **out) {
  HCU_ARG(out, "ctx out pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    hcu_set_error("no CUDA device available (%s); heracles_cuda has no CPU path",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return HCU_ERR_NODEVICE;
  }
  HCU_ARG(device >= 0 && device < n, "device index");
  HCU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  HCU_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    hcu_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device,
                  prop.major, prop.minor);
    return HCU_ERR_UNSUPPORTED;
  }
  hcu_ctx *ctx = new hcu_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (const char *e = getenv("HCU_SLOT_ROWS")) {
    const long long v = atoll(e);
    if (v >= (1 << 12) && v <= (1 << 24)) ctx->SLOT_ROWS = v;
  }
  HCU_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  HCU_CUDA(cudaMalloc(&ctx->bad_rows, sizeof(unsigned long long)));
  HCU_CUDA(cudaMemset(ctx->bad_rows, 0, sizeof(unsigned long long)));
  HCU_CUDA(cudaMalloc(&ctx->work_counters, 2 * sizeof(double)));
  HCU_CUDA(cudaMemset(ctx->work_counters, 0, 2 * sizeof(double)));
  for (int i = 0; i < 6; ++i) HCU_CUDA(cudaEventCreate(&ctx->ev[i]));
  {
    const char *e = getenv("HCU_START_TABLE");  // 0: walk the dead zone in every pass (A/B timing)
    if (e && e[0] == '0') ctx->use_start_table = false;
  }
  *out = ctx;
  return HCU_OK;
}

static void free_buffer(hcu_buffer *b) {
  if (b->ptr) cudaFree(b->ptr);
  b->ptr = nullptr;
  b->bytes = 0;
}

extern "C" int hcu_trim(hcu_ctx *ctx) {
  HCU_ARG(ctx, "ctx");
  HCU_CUDA(cudaSetDevice(ctx->device));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  free_buffer(&ctx->ws_phase);
  free_buffer(&ctx->ws_belt);
  free_buffer(&ctx->ws_cap);
  free_buffer(&ctx->ws_scr);
  free_buffer(&ctx->ws_map);
  free_buffer(&ctx->ws_alm);
  free_buffer(&ctx->ws_misc);
  free_buffer(&ctx->ws_state);
  free_buffer(&ctx->ws_resid);
  free_buffer(&ctx->ws_pw);
  return HCU_OK;
}

extern "C" int hcu_destroy(hcu_ctx *ctx) {
  if (!ctx) return HCU_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  hcu_trim(ctx);
  for (auto &kv : ctx->geom) {
    cudaFree(kv.second.cth);
    cudaFree(kv.second.sth);
    cudaFree(kv.second.ch);
    cudaFree(kv.second.sh);
    if (kv.second.bfilt) cudaFree(kv.second.bfilt);
    if (kv.second.bfilt_off) cudaFree(kv.second.bfilt_off);
    hcu_ring2_free(&kv.second);
  }
  for (auto &kv : ctx->start) {
    if (kv.second.sub) cudaFree(kv.second.sub);
    if (kv.second.state) cudaFree(kv.second.state);
  }
  for (auto &kv : ctx->coef) {
    cudaFree(kv.second.tab);
    cudaFree(kv.second.scale);
    cudaFree(kv.second.cm);
  }
  for (auto &kv : ctx->belt_plan) cufftDestroy(kv.second);
  for (auto &kv : ctx->belt_plan_inv) cufftDestroy(kv.second);
  for (int i = 0; i < hcu_ctx::NSLOT; ++i) {
    if (ctx->slot[i].host) cudaFreeHost(ctx->slot[i].host);
    if (ctx->slot[i].dev) cudaFree(ctx->slot[i].dev);
    if (ctx->slot[i].done) cudaEventDestroy(ctx->slot[i].done);
    if (ctx->slot[i].ready) cudaEventDestroy(ctx->slot[i].ready);
  }
  cudaFree(ctx->bad_rows);
  cudaFree(ctx->work_counters);
  for (int i = 0; i < 6; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return HCU_OK;
}

extern "C" int hcu_set_stream(hcu_ctx *ctx, void *s) {
  HCU_ARG(ctx, "ctx");
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return HCU_OK;
}

extern "C" int hcu_synchronize(hcu_ctx *ctx) {
  HCU_ARG(ctx, "ctx");
  HCU_CUDA(cudaSetDevice(ctx->device));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  return HCU_OK;
}

extern "C" int hcu_launch_count(hcu_ctx *ctx, int64_t *own, int64_t *cufft) {
  HCU_ARG(ctx, "ctx");
  if (own) *own = ctx->n_launch;
  if (cufft) *cufft = ctx->n_cufft;
  return HCU_OK;
}

int hcu_ws_reserve(hcu_ctx *ctx, hcu_buffer *b, size_t bytes) {
  if (b->bytes >= bytes) return HCU_OK;
  if (b->ptr) {
    HCU_CUDA(cudaStreamSynchronize(ctx->stream));
    HCU_CUDA(cudaFree(b->ptr));
    b->ptr = nullptr;
    b->bytes = 0;
  }
  cudaError_t e = cudaMalloc(&b->ptr, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hcu_set_error("workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return HCU_ERR_NOMEM;
  }
  b->bytes = bytes;
  return HCU_OK;
}

// ---------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------
enum PtrKind { PK_PAGEABLE = 0, PK_PINNED = 1, PK_DEVICE = 2, PK_MANAGED = 3 };

static PtrKind ptr_kind(const void *p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return PK_PAGEABLE;
  }
  switch (at.type) {
    case cudaMemoryTypeHost: return PK_PINNED;
    case cudaMemoryTypeDevice: return PK_DEVICE;
    case cudaMemoryTypeManaged: return PK_MANAGED;
    default: return PK_PAGEABLE;
  }
}
static bool dev_accessible(PtrKind k) { return k == PK_DEVICE || k == PK_MANAGED; }
bool hcu_dev_accessible(const void *p) { return dev_accessible(ptr_kind(p)); }

extern "C" int hcu_malloc_managed(hcu_ctx *ctx, size_t bytes, void **ptr) {
  HCU_ARG(ctx && ptr, "hcu_malloc_managed");
  HCU_CUDA(cudaSetDevice(ctx->device));
  if (bytes == 0) bytes = 8;
  cudaError_t e = cudaMallocManaged(ptr, bytes, cudaMemAttachGlobal);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hcu_set_error("cudaMallocManaged(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return HCU_ERR_NOMEM;
  }
  // keep the pages on the device; host reads migrate on demand
  cudaMemAdvise(*ptr, bytes, cudaMemAdviseSetPreferredLocation, ctx->device);
  cudaGetLastError();
  return HCU_OK;
}

extern "C" int hcu_malloc_device(hcu_ctx *ctx, size_t bytes, void **ptr) {
  HCU_ARG(ctx && ptr, "hcu_malloc_device");
  HCU_CUDA(cudaSetDevice(ctx->device));
  if (bytes == 0) bytes = 8;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hcu_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return HCU_ERR_NOMEM;
  }
  return HCU_OK;
}

extern "C" int hcu_malloc_pinned(hcu_ctx *ctx, size_t bytes, void **ptr) {
  HCU_ARG(ctx && ptr, "hcu_malloc_pinned");
  HCU_CUDA(cudaSetDevice(ctx->device));
  if (bytes == 0) bytes = 8;
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    hcu_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return HCU_ERR_NOMEM;
  }
  return HCU_OK;
}

extern "C" int hcu_free(hcu_ctx *ctx, void *ptr) {
  HCU_ARG(ctx, "ctx");
  if (!ptr) return HCU_OK;
  HCU_CUDA(cudaSetDevice(ctx->device));
  PtrKind k = ptr_kind(ptr);
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  if (k == PK_PINNED)
    HCU_CUDA(cudaFreeHost(ptr));
  else if (k == PK_DEVICE || k == PK_MANAGED)
    HCU_CUDA(cudaFree(ptr));
  else {
    hcu_set_error("hcu_free: pointer was not allocated by this library");
    return HCU_ERR_ARG;
  }
  return HCU_OK;
}

extern "C" int hcu_prefetch(hcu_ctx *ctx, const void *ptr, size_t bytes, int to_device) {
  HCU_ARG(ctx && ptr, "hcu_prefetch");
  if (ptr_kind(ptr) != PK_MANAGED || bytes == 0) return HCU_OK;
  HCU_CUDA(cudaMemPrefetchAsync(ptr, bytes, to_device ? ctx->device : cudaCpuDeviceId, ctx->stream));
  return HCU_OK;
}

extern "C" int hcu_memset_zero(hcu_ctx *ctx, void *ptr, size_t bytes) {
  HCU_ARG(ctx && ptr, "hcu_memset_zero");
  HCU_CUDA(cudaMemsetAsync(ptr, 0, bytes, ctx->stream));
  return HCU_OK;
}

extern "C" int hcu_memcpy(hcu_ctx *ctx, void *dst, const void *src, size_t bytes) {
  HCU_ARG(ctx && dst && src, "hcu_memcpy");
  HCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
  // pageable host memory on either side: the runtime already staged it, but the
  // caller may reuse the host buffer right away only after completion
  if (ptr_kind(dst) == PK_PAGEABLE || ptr_kind(src) == PK_PAGEABLE)
    HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  return HCU_OK;
}

// ---------------------------------------------------------------------------
// catalogue pages -> device: pinned, multi-slot staging
// ---------------------------------------------------------------------------
static int ensure_slots(hcu_ctx *ctx) {
  for (int i = 0; i < hcu_ctx::NSLOT; ++i) {
    hcu_stage_slot &s = ctx->slot[i];
    if (s.dev) continue;
    size_t bytes = sizeof(double) * ctx->SLOT_ROWS * hcu_ctx::SLOT_COLS;
    HCU_CUDA(cudaHostAlloc(&s.host, bytes, cudaHostAllocDefault));
    HCU_CUDA(cudaMalloc(&s.dev, bytes));
    HCU_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    HCU_CUDA(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
  }
  if (!ctx->copy_stream) HCU_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  return HCU_OK;
}

static void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  const size_t MIN_PER_THREAD = 1 << 20;
  int nt = (int)std::min<size_t>(4, bytes / MIN_PER_THREAD);
  if (nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::thread th[4];
  size_t per = (bytes / nt + 63) & ~(size_t)63;
  for (int t = 0; t < nt; ++t) {
    size_t off = (size_t)t * per;
    if (off >= bytes) {
      nt = t;
      break;
    }
    size_t len = std::min(per, bytes - off);
    th[t] = std::thread([=]() { memcpy((char *)dst + off, (const char *)src + off, len); });
  }
  for (int t = 0; t < nt; ++t) th[t].join();
}

// Bring column `src` (rows r0..r0+nr) to the device; returns the device pointer to use.
static int stage_column(hcu_ctx *ctx, hcu_stage_slot &s, int col, const double *src,
                        PtrKind kind, i64 r0, i64 nr, const double **dev) {
  if (dev_accessible(kind)) {
    *dev = src + r0;
    return HCU_OK;
  }
  double *d = s.dev + (i64)col * ctx->SLOT_ROWS;
  // the copies run on their own stream so that they overlap the scatter kernel of the previous chunk
  if (kind == PK_PINNED) {
    HCU_CUDA(cudaMemcpyAsync(d, src + r0, sizeof(double) * nr, cudaMemcpyHostToDevice, ctx->copy_stream));
  } else {
    double *h = s.host + (i64)col * ctx->SLOT_ROWS;
    parallel_memcpy(h, src + r0, sizeof(double) * nr);
    HCU_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * nr, cudaMemcpyHostToDevice, ctx->copy_stream));
  }
  *dev = d;
  return HCU_OK;
}

static int map_values_impl(hcu_ctx *ctx, i64 nside, int scheme, const double *lon,
                           const double *lat, const double *values, i64 vstride,
                           int nv, i64 n, double *maps, i64 mstride, int flags,
                           i64 *ipix_dev) {
  const PtrKind klon = ptr_kind(lon), klat = ptr_kind(lat);
  const PtrKind kval = nv > 0 ? ptr_kind(values) : PK_DEVICE;
  const bool direct = dev_accessible(klon) && dev_accessible(klat) && dev_accessible(kval);
  if (direct)
    return hcu_launch_map_values(ctx, nside, scheme, lon, lat, values, vstride, nv, n, maps,
                                 mstride, flags, ipix_dev);
  HCU_ARG(nv <= hcu_ctx::SLOT_COLS - 2, "at most 2 value rows when staging host pages");
  HCU_CHECK(ensure_slots(ctx));
  for (i64 r0 = 0; r0 < n; r0 += ctx->SLOT_ROWS) {
    const i64 nr = std::min<i64>(ctx->SLOT_ROWS, n - r0);
    hcu_stage_slot &s = ctx->slot[ctx->next_slot];
    ctx->next_slot = (ctx->next_slot + 1) % hcu_ctx::NSLOT;
    if (s.used) HCU_CUDA(cudaEventSynchronize(s.done));  // slot (pinned + device buffer) free again
    const double *dlon, *dlat, *dval = nullptr;
    HCU_CHECK(stage_column(ctx, s, 0, lon, klon, r0, nr, &dlon));
    HCU_CHECK(stage_column(ctx, s, 1, lat, klat, r0, nr, &dlat));
    i64 dvstride = vstride;
    if (nv > 0) {
      if (dev_accessible(kval)) {
        dval = values + r0;
      } else {
        for (int v = 0; v < nv; ++v) {
          const double *tmp;
          HCU_CHECK(stage_column(ctx, s, 2 + v, values + (i64)v * vstride, kval, r0, nr, &tmp));
        }
        dval = s.dev + 2 * ctx->SLOT_ROWS;
        dvstride = ctx->SLOT_ROWS;
      }
    }
    HCU_CUDA(cudaEventRecord(s.ready, ctx->copy_stream));
    HCU_CUDA(cudaStreamWaitEvent(ctx->stream, s.ready, 0));
    HCU_CHECK(hcu_launch_map_values(ctx, nside, scheme, dlon, dlat, dval, dvstride, nv, nr,
                                    maps, mstride, flags, ipix_dev ? ipix_dev + r0 : nullptr));
    HCU_CUDA(cudaEventRecord(s.done, ctx->stream));
    s.used = true;
  }
  return HCU_OK;
}

static bool valid_nside(i64 nside) {
  return nside >= 1 && nside <= (1 << 24) && (nside & (nside - 1)) == 0;
}

int hcu_launch_map_page(hcu_ctx *ctx, i64 nside, int scheme, const double *lon, const double *lat, const double *w,
                        const double *g1, const double *g2, i64 n, double *pos, double *she, i64 she_stride,
                        double *stats);

extern "C" int hcu_map_page(hcu_ctx *ctx, int64_t nside, int scheme, const double *lon, const double *lat,
                            const double *w, const double *g1, const double *g2, int64_t n, double *pos,
                            double *she, int64_t she_stride, double *stats) {
  HCU_ARG(ctx, "ctx");
  HCU_ARG(valid_nside(nside), "nside must be a power of two");
  HCU_ARG(scheme == HCU_RING || scheme == HCU_NEST, "scheme");
  HCU_ARG(n >= 0, "n >= 0");
  if (n == 0) return HCU_OK;
  HCU_ARG(lon && lat && (pos || she), "null pointer");
  HCU_ARG(!she || (g1 && g2), "shear map without shear columns");
  HCU_ARG(!pos || dev_accessible(ptr_kind(pos)), "pos must be device or managed memory");
  HCU_ARG(!she || dev_accessible(ptr_kind(she)), "she must be device or managed memory");
  HCU_ARG(!stats || dev_accessible(ptr_kind(stats)), "stats must be device or managed memory");
  HCU_CUDA(cudaSetDevice(ctx->device));
  const double *col[5] = {lon, lat, w, she ? g1 : nullptr, she ? g2 : nullptr};
  PtrKind kind[5];
  bool direct = true;
  for (int c = 0; c < 5; ++c) {
    kind[c] = col[c] ? ptr_kind(col[c]) : PK_DEVICE;
    direct = direct && dev_accessible(kind[c]);
  }
  if (direct) return hcu_launch_map_page(ctx, nside, scheme, lon, lat, w, col[3], col[4], n, pos, she, she_stride, stats);
  HCU_CHECK(ensure_slots(ctx));
  for (i64 r0 = 0; r0 < n; r0 += ctx->SLOT_ROWS) {
    const i64 nr = std::min<i64>(ctx->SLOT_ROWS, n - r0);
    hcu_stage_slot &s = ctx->slot[ctx->next_slot];
    ctx->next_slot = (ctx->next_slot + 1) % hcu_ctx::NSLOT;
    if (s.used) HCU_CUDA(cudaEventSynchronize(s.done));
    const double *d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < 5; ++c)
      if (col[c]) HCU_CHECK(stage_column(ctx, s, c, col[c], kind[c], r0, nr, &d[c]));
    HCU_CUDA(cudaEventRecord(s.ready, ctx->copy_stream));
    HCU_CUDA(cudaStreamWaitEvent(ctx->stream, s.ready, 0));
    HCU_CHECK(hcu_launch_map_page(ctx, nside, scheme, d[0], d[1], d[2], d[3], d[4], nr, pos, she, she_stride, stats));
    HCU_CUDA(cudaEventRecord(s.done, ctx->stream));
    s.used = true;
  }
  return HCU_OK;
}

extern "C" int hcu_set_weights_mode(hcu_ctx *ctx, int per_pass) {
  HCU_ARG(ctx, "ctx");
  ctx->weights_premultiply = per_pass == 0;
  return HCU_OK;
}

extern "C" int hcu_multiply(hcu_ctx *ctx, double *out, const double *a, const double *b, int64_t n) {
  HCU_ARG(ctx && out && a && b && n >= 0, "hcu_multiply");
  return hcu_mul(ctx, out, a, b, n);
}

extern "C" int hcu_set_timing(hcu_ctx *ctx, int enabled) {
  HCU_ARG(ctx, "ctx");
  ctx->timing = enabled != 0;
  return HCU_OK;
}

extern "C" int hcu_map_values(hcu_ctx *ctx, int64_t nside, int scheme, const double *lon,
                              const double *lat, const double *values,
                              int64_t value_stride, int nv, int64_t n, double *maps,
                              int64_t map_stride, int flags) {
  HCU_ARG(ctx, "ctx");
  HCU_ARG(valid_nside(nside), "nside must be a power of two");
  HCU_ARG(scheme == HCU_RING || scheme == HCU_NEST, "scheme");
  HCU_ARG(n >= 0 && nv >= 1 && nv <= 4, "n >= 0, 1 <= nv <= 4");
  if (n == 0) return HCU_OK;
  HCU_ARG(lon && lat && values && maps, "null pointer");
  HCU_ARG(dev_accessible(ptr_kind(maps)), "maps must be device or managed memory");
  HCU_CUDA(cudaSetDevice(ctx->device));
  return map_values_impl(ctx, nside, scheme, lon, lat, values, value_stride, nv, n, maps,
                         map_stride, flags, nullptr);
}

extern "C" int hcu_ang2pix(hcu_ctx *ctx, int64_t nside, int scheme, const double *lon,
                           const double *lat, int64_t n, int64_t *ipix) {
  HCU_ARG(ctx, "ctx");
  HCU_ARG(valid_nside(nside), "nside must be a power of two");
  HCU_ARG(scheme == HCU_RING || scheme == HCU_NEST, "scheme");
  HCU_ARG(n >= 0, "n");
  if (n == 0) return HCU_OK;
  HCU_ARG(lon && lat && ipix, "null pointer");
  HCU_CUDA(cudaSetDevice(ctx->device));
  i64 *out = ipix;
  const bool host_out = !dev_accessible(ptr_kind(ipix));
  if (host_out) {
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_misc, sizeof(i64) * n));
    out = (i64 *)ctx->ws_misc.ptr;
  }
  // rows rejected by ang2pix are reported through ipix = -1, not the bad-row counter
  HCU_CHECK(map_values_impl(ctx, nside, scheme, lon, lat, nullptr, 0, 0, n, nullptr, 0, 0, out));
  if (host_out)
    HCU_CUDA(cudaMemcpyAsync(ipix, out, sizeof(i64) * n, cudaMemcpyDeviceToHost, ctx->stream));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  return HCU_OK;
}

extern "C" int hcu_bad_rows(hcu_ctx *ctx, int64_t *count) {
  HCU_ARG(ctx && count, "hcu_bad_rows");
  unsigned long long v = 0;
  HCU_CUDA(cudaMemcpyAsync(&v, ctx->bad_rows, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
  HCU_CUDA(cudaMemsetAsync(ctx->bad_rows, 0, sizeof(v), ctx->stream));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  *count = (int64_t)v;
  return HCU_OK;
}

extern "C" int hcu_ud_grade(hcu_ctx *ctx, int64_t nside_in, const double *in,
                            int64_t nside_out, double *out) {
  HCU_ARG(ctx && in && out, "hcu_ud_grade");
  HCU_ARG(valid_nside(nside_in) && valid_nside(nside_out), "nside must be a power of two");
  HCU_CUDA(cudaSetDevice(ctx->device));
  const i64 npi = 12 * nside_in * nside_in, npo = 12 * nside_out * nside_out;
  const double *din = in;
  double *dout = out;
  const bool hin = !dev_accessible(ptr_kind(in)), hout = !dev_accessible(ptr_kind(out));
  if (hin || hout) {
    HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_map, sizeof(double) * (npi + npo)));
    if (hin) {
      HCU_CUDA(cudaMemcpyAsync(ctx->ws_map.ptr, in, sizeof(double) * npi, cudaMemcpyDefault, ctx->stream));
      din = (double *)ctx->ws_map.ptr;
    }
    if (hout) dout = (double *)ctx->ws_map.ptr + npi;
  }
  HCU_CHECK(hcu_ud_grade_dev(ctx, nside_in, din, nside_out, dout));
  if (hout) {
    HCU_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * npo, cudaMemcpyDefault, ctx->stream));
    HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HCU_OK;
}

// ---------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------
int hcu_get_geom(hcu_ctx *ctx, i64 nside, hcu_geom **out) {
  auto it = ctx->geom.find(nside);
  if (it != ctx->geom.end()) {
    *out = &it->second;
    return HCU_OK;
  }
  hcu_geom g;
  g.nside = nside;
  g.nrp = (int)(2 * nside);
  std::vector<double> cth(g.nrp), sth(g.nrp), ch(g.nrp), sh(g.nrp);
  for (i64 ir = 1; ir <= g.nrp; ++ir) {
    long double z, sv;
    if (ir < nside) {
      long double tmp = (long double)(ir * ir) * 4 / (long double)(12 * nside * nside);
      z = 1 - tmp;
      sv = sqrtl(tmp * (2 - tmp));
    } else {
      z = (long double)(2 * nside - ir) * 2 / (long double)(3 * nside);
      sv = sqrtl((1 + z) * (1 - z));
    }
    long double c2 = sqrtl((1 + z) / 2);
    cth[ir - 1] = (double)z;
    sth[ir - 1] = (double)sv;
    ch[ir - 1] = (double)c2;
    sh[ir - 1] = (double)(sv / (2 * c2));
  }
  size_t bytes = sizeof(double) * g.nrp;
  HCU_CUDA(cudaMalloc(&g.cth, bytes));
  HCU_CUDA(cudaMalloc(&g.sth, bytes));
  HCU_CUDA(cudaMalloc(&g.ch, bytes));
  HCU_CUDA(cudaMalloc(&g.sh, bytes));
  HCU_CUDA(cudaMemcpy(g.cth, cth.data(), bytes, cudaMemcpyHostToDevice));
  HCU_CUDA(cudaMemcpy(g.sth, sth.data(), bytes, cudaMemcpyHostToDevice));
  HCU_CUDA(cudaMemcpy(g.ch, ch.data(), bytes, cudaMemcpyHostToDevice));
  HCU_CUDA(cudaMemcpy(g.sh, sh.data(), bytes, cudaMemcpyHostToDevice));
  ctx->geom[nside] = g;
  hcu_geom *gp = &ctx->geom[nside];
  HCU_CHECK(hcu_build_bluestein(ctx, gp));
  *out = gp;
  return HCU_OK;
}

int hcu_get_coef(hcu_ctx *ctx, int lmax, int spin, hcu_coef **out) {
  auto key = std::make_pair(lmax, spin);
  auto it = ctx->coef.find(key);
  if (it != ctx->coef.end()) {
    *out = &it->second;
    return HCU_OK;
  }
  hcu_coef c;
  c.lmax = lmax;
  c.spin = spin;
  HCU_CHECK(hcu_build_coef(ctx, &c));
  ctx->coef[key] = c;
  *out = &ctx->coef[key];
  return HCU_OK;
}

int hcu_get_start(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, hcu_start **out) {
  *out = nullptr;
  if (!ctx->use_start_table || g->nside <= 0) return HCU_OK;  // (nside 0: a geometry of free points, hcu_points2alm)
  auto key = std::make_pair(g->nside, std::make_pair(c->lmax, c->spin));
  auto it = ctx->start.find(key);
  if (it == ctx->start.end()) {
    hcu_start t;
    // 20 bytes per (m, ring pair, chain): 3.8 GB for spin 0 + 2 at nside 4096 / lmax 8192.  Built only when it leaves
    // at least 3/4 of the free memory alone; otherwise (nside 8192) the kernels walk the dead zone as before
    const size_t n = (size_t)(c->lmax + 1) * g->nrp * (c->spin == 0 ? 1 : 2);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (n * 20 < free_b / 4 && hcu_build_start(ctx, g, c, &t) != HCU_OK) {
      cudaGetLastError();
      t = hcu_start();
    }
    ctx->start[key] = t;
    it = ctx->start.find(key);
  }
  if (it->second.sub) *out = &it->second;
  return HCU_OK;
}

extern "C" int hcu_set_start_table(hcu_ctx *ctx, int enabled) {
  HCU_ARG(ctx, "ctx");
  ctx->use_start_table = enabled != 0;
  return HCU_OK;
}

// ---------------------------------------------------------------------------
// transforms
// ---------------------------------------------------------------------------
namespace {
__global__ void dfma_peak_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

extern "C" int hcu_last_sht_timing(hcu_ctx *ctx, float ms[4]) {
  HCU_ARG(ctx && ms, "hcu_last_sht_timing");
  for (int i = 0; i < 4; ++i) ms[i] = ctx->sht_ms[i];
  return HCU_OK;
}

extern "C" int hcu_last_sht_work(hcu_ctx *ctx, double *rec, double *acc) {
  HCU_ARG(ctx, "ctx");
  double h[2] = {0, 0};
  HCU_CUDA(cudaMemcpyAsync(h, ctx->work_counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  HCU_CUDA(cudaStreamSynchronize(ctx->stream));
  if (rec) *rec = h[0];
  if (acc) *acc = h[1];
  return HCU_OK;
}

extern "C" int hcu_measure_fp64_peak(hcu_ctx *ctx, double *flops) {
  HCU_ARG(ctx && flops, "hcu_measure_fp64_peak");
  HCU_CUDA(cudaSetDevice(ctx->device));
  const int blocks = ctx->num_sms * 8, threads = 256, iters = 1 << 14;
  HCU_CHECK(hcu_ws_reserve(ctx, &ctx->ws_misc, sizeof(double) * blocks * threads));
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    HCU_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double *)ctx->ws_misc.ptr, iters);
    HCU_LAUNCH_CHECK(ctx);
    HCU_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    HCU_CUDA(cudaEventSynchronize(ctx->ev[1]));
    float ms = 0;
    HCU_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    double f = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3);
    if (rep > 0 && f > best) best = f;
  }
  *flops = best;
  return HCU_OK;
}
