// hcu_common.cuh -- shared declarations of the heracles_cuda library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/heracles_cuda.h"

typedef int64_t i64;

void hcu_set_error(const char *fmt, ...);

#define HCU_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) {                                                  \
      hcu_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,        \
                    cudaGetErrorString(e_));                                  \
      return HCU_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

#define HCU_CUFFT(call)                                                       \
  do {                                                                        \
    cufftResult r_ = (call);                                                  \
    if (r_ != CUFFT_SUCCESS) {                                                \
      hcu_set_error("%s:%d: %s failed: cufft error %d", __FILE__, __LINE__,   \
                    #call, (int)r_);                                          \
      return HCU_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

#define HCU_CHECK(call)                                                       \
  do {                                                                        \
    int s_ = (call);                                                          \
    if (s_ != HCU_OK) return s_;                                              \
  } while (0)

#define HCU_ARG(cond, msg)                                                    \
  do {                                                                        \
    if (!(cond)) {                                                            \
      hcu_set_error("invalid argument: %s", msg);                             \
      return HCU_ERR_ARG;                                                     \
    }                                                                         \
  } while (0)

#define HCU_LAUNCH_CHECK(ctx)                                                 \
  do {                                                                        \
    (ctx)->n_launch++;                                                        \
    HCU_CUDA(cudaGetLastError());                                             \
  } while (0)

// up to HCU_MAX_BATCH row pointers passed to kernels by value (maps or alm rows
// of one transform batch need not be contiguous in memory)
#define HCU_MAX_BATCH 16
struct hcu_ptrs {
  double *p[HCU_MAX_BATCH];
};

// Destinations of the phase rows a ring-FFT launch writes (multi-GPU: rows are ordered by the rank that owns their m,
// and every rank's rows go STRAIGHT into that rank's buffer over NVLink peer memory): rows [row_start[d],
// row_start[d+1]) live at base[d] + (row - row_start[d]) * (nrp_local * ncomp * 4).  nd == 0: one local array.
#define HCU_MAX_BLOCKS 16
struct hcu_rowdest {
  int nd = 0;
  int row_start[HCU_MAX_BLOCKS + 1];
  double *base[HCU_MAX_BLOCKS];
};
#ifdef __CUDACC__
__device__ __forceinline__ double *hcu_row_ptr(const hcu_rowdest &rd, double *phase, int row, i64 rstride) {
  if (rd.nd == 0) return phase + (i64)row * rstride;
  int d = 0;
  while (d + 1 < rd.nd && row >= rd.row_start[d + 1]) ++d;
  return rd.base[d] + (i64)(row - rd.row_start[d]) * rstride;
}
#endif

// a growable device workspace
struct hcu_buffer {
  void *ptr = nullptr;
  size_t bytes = 0;
};

// per-nside tables
struct hcu_geom {
  i64 nside = 0;
  int nrp = 0;           // ring pairs = 2 nside
  double *cth = nullptr; // [nrp] cos(theta) of the north ring
  double *sth = nullptr; // [nrp] sin(theta)
  double *ch = nullptr;  // [nrp] cos(theta/2)
  double *sh = nullptr;  // [nrp] sin(theta/2)
  // Bluestein filter spectra for the polar-cap sub-FFTs of length i = 1..nside-1
  double2 *bfilt = nullptr; // concatenated, bit-reversed FFT_M(chirp)/M
  i64 *bfilt_off = nullptr; // [nside] offsets (device)
  std::vector<i64> bfilt_off_h;
  // second-generation ring FFT (k_ringfft2.cu): pass twiddles for 2^4 .. 2^13 points, per-ring chirp and radix-4
  // twiddle tables of the cap rings 1 .. r2_imax, radix-4 twiddles of the belt
  double2 *r2_tw = nullptr, *r2_chirp = nullptr, *r2_wtab = nullptr, *r2_wbelt = nullptr;
  int r2_tw_off[14] = {0};
  int r2_imax = 0;      // largest cap ring number handled by the second generation (0: none)
  bool r2_belt = false; // belt handled by the second generation
};

// per-(lmax, spin) recursion coefficient table
struct hcu_coef {
  int lmax = 0, spin = 0;
  double *tab = nullptr;   // scaled recursion: spin 0 double A_l; spin 2 double2 (A_l, B_l); see k_legendre.cu
  double *scale = nullptr; // s_l with lambda_l = s_l q_l
  double *cm = nullptr;    // [2 (lmax+1)] start-value normalisation
};

// per-(nside, lmax, spin) start states of the Legendre recursions: the dead zone l0 <= l < l_start(m, theta), in which
// lambda_lm is below 2^-200 and contributes nothing, is walked ONCE when the table is built instead of in every pass
struct hcu_start {
  int *sub = nullptr;        // [m][ring pair][NJ]: first sub-chunk (16 l, counted from l0) that starts representable; INT_MAX: never
  double2 *state = nullptr;  // [m][ring pair][NJ]: (prev, cur) of the scaled recursion at the start of that sub-chunk
};

struct hcu_stage_slot {
  double *host = nullptr; // pinned
  double *dev = nullptr;
  cudaEvent_t done = nullptr;  // the kernel that consumed the slot has finished
  cudaEvent_t ready = nullptr; // the H2D copies into the slot have landed
  bool used = false;
};

struct hcu_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr; // H2D staging of catalogue pages, overlapped with the scatter kernel
  cudaStream_t stream = nullptr;
  i64 n_launch = 0, n_cufft = 0;
  // staging for pageable host pages
  static const int NSLOT = 3;
  i64 SLOT_ROWS = 1 << 20;  // rows per staging slot (HCU_SLOT_ROWS at hcu_create); measured 47.5 / 50.8 / 53.1 / 53.2 GB/s host to map at 2^18 .. 2^21
  static const int SLOT_COLS = 5; // lon, lat, up to 3 more columns (hcu_map_values: 2 value rows; hcu_map_page: w, g1, g2)
  hcu_stage_slot slot[NSLOT];
  int next_slot = 0;
  unsigned long long *bad_rows = nullptr; // device counter
  // workspaces
  hcu_buffer ws_phase, ws_belt, ws_cap, ws_scr, ws_map, ws_alm, ws_misc, ws_state, ws_resid, ws_pw;
  // tables
  std::map<i64, hcu_geom> geom;
  std::map<std::pair<int, int>, hcu_coef> coef;
  std::map<std::pair<i64, std::pair<int, int>>, hcu_start> start;  // key (nside, (lmax, spin))
  bool use_start_table = true;
  std::map<i64, cufftHandle> belt_plan, belt_plan_inv;
  // timing
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float sht_ms[4] = {0, 0, 0, 0};
  bool weights_premultiply = true;  // hcu_set_weights_mode: pixel weights applied once to the map (healpy) or in every analysis pass
  bool timing = false;             // hcu_set_timing: CUDA events (and a host wait per batch) around the SHT stages
  double *work_counters = nullptr; // device [2]
};

int hcu_ws_reserve(hcu_ctx *ctx, hcu_buffer *b, size_t bytes);
int hcu_get_geom(hcu_ctx *ctx, i64 nside, hcu_geom **out);
int hcu_get_coef(hcu_ctx *ctx, int lmax, int spin, hcu_coef **out);
// start-state table for (geometry, coefficients); *out = nullptr when disabled or when it does not fit in memory
int hcu_get_start(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, hcu_start **out);
int hcu_build_start(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, hcu_start *t);

// kernels' host launchers (defined in the k_*.cu files)
int hcu_launch_map_values(hcu_ctx *ctx, i64 nside, int scheme, const double *lon,
                          const double *lat, const double *values, i64 vstride,
                          int nv, i64 n, double *maps, i64 mstride, int flags,
                          i64 *ipix_out);
int hcu_build_bluestein(hcu_ctx *ctx, hcu_geom *g);
int hcu_ring_fft_forward(hcu_ctx *ctx, hcu_geom *g, int lmax, int ncomp,
                         const hcu_ptrs &maps, const double *ring_weights,
                         i64 rp_lo, i64 rp_hi, const int32_t *mlist, int nm, double *phase,
                         const hcu_rowdest *dest = nullptr);
int hcu_ring_fft_inverse(hcu_ctx *ctx, hcu_geom *g, int lmax, int ncomp,
                         const double *phase, const int32_t *mpos, i64 rp_lo, i64 rp_hi,
                         const hcu_ptrs &maps);
// second generation (k_ringfft2.cu)
int hcu_ring2_build(hcu_ctx *ctx, hcu_geom *g, int cap_max_m);
void hcu_ring2_free(hcu_geom *g);
int hcu_ring2_run(hcu_ctx *ctx, hcu_geom *g, bool inverse, bool belt, int lmax, int ncomp, const hcu_ptrs &maps,
                  const double *ring_weights, i64 rp_lo, i64 nrp_local, i64 rp_a, i64 rp_b, const int32_t *mlist,
                  int nm, const int32_t *mpos, double *phase, const hcu_rowdest *dest = nullptr);
int hcu_build_coef(hcu_ctx *ctx, hcu_coef *c);
int hcu_legendre_batch(int spin);  // components one Legendre launch can take: 12 (spin 0), 8 (spin 2)
// (both Legendre launchers look the start-state table up themselves)
int hcu_legendre_analysis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                          int spin, int ncomp, const double *phase,
                          const int32_t *mlist_dev, int nm, int nblk, const i64 *rp_bounds,
                          const double *fl_dev, const hcu_ptrs &alm);
int hcu_legendre_synthesis(hcu_ctx *ctx, hcu_geom *g, hcu_coef *c, int lmax,
                           int spin, int ncomp, const hcu_ptrs &alm,
                           const int32_t *mlist_dev, int nm, int nblk, const i64 *rp_bounds,
                           double *phase, double *const *block_out = nullptr);

static inline int ilog2_host(i64 v) {
  int r = 0;
  while (v > 1) {
    v >>= 1;
    ++r;
  }
  return r;
}
