/*
 * heracles_cuda.h -- C ABI of the B200-native catalogue -> map -> alm -> Cl backend.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy
 * types.  Each entry point names the reference call site it replaces
 * (paths relative to the heracles-ec/heracles source tree).  The Python host
 * side (heracles_b200/) binds these with ctypes; INTEGRATION.md shows the
 * stub a Heracles maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative hcu_status on failure;
 *     hcu_last_error() returns a thread-local message for the last failure.
 *     No exceptions, no aborts cross this boundary.
 *   - "any" pointers may be pageable host, pinned host, managed or device
 *     memory; the library detects which (cudaPointerGetAttributes) and stages
 *     pageable host data through pinned double buffers.
 *   - all work is queued on the context's stream (hcu_set_stream); calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - maps are HEALPix RING-ordered float64, npix = 12 nside^2; alm are
 *     complex128 in healpy's m-major layout, idx(l,m) = m(2 lmax+1-m)/2 + l,
 *     mmax = lmax.
 *   - there is NO CPU fallback: without a CUDA device hcu_create fails.
 */
#ifndef HERACLES_CUDA_H
#define HERACLES_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCU_VERSION 100

typedef struct hcu_ctx hcu_ctx;

typedef enum {
  HCU_OK = 0,
  HCU_ERR_CUDA = -1,     /* a CUDA / cuFFT call failed */
  HCU_ERR_ARG = -2,      /* invalid argument */
  HCU_ERR_NOMEM = -3,    /* allocation failed */
  HCU_ERR_UNSUPPORTED = -4,
  HCU_ERR_NODEVICE = -5  /* no CUDA device: the library has no CPU path */
} hcu_status;

enum { HCU_RING = 0, HCU_NEST = 1 };

/* hcu_map_values flags */
enum {
  HCU_MAP_DEFAULT = 0,
  HCU_MAP_AGGREGATE = 1 /* combine equal pixels inside a warp before the atomic */
};

/* ---- library / context ------------------------------------------------ */
int hcu_version(void);
const char *hcu_last_error(void);
int hcu_device_count(int *count);
int hcu_create(int device, hcu_ctx **ctx);
int hcu_destroy(hcu_ctx *ctx);
/* use a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = the context's own
 * non-blocking stream.  To run on the legacy default stream pass cudaStreamLegacy ((void *)1). */
int hcu_set_stream(hcu_ctx *ctx, void *cuda_stream);
int hcu_synchronize(hcu_ctx *ctx);
/* number of kernels of this library (and cuFFT executions) launched so far */
int hcu_launch_count(hcu_ctx *ctx, int64_t *own_kernels, int64_t *cufft_execs);
/* release cached workspaces (tables stay) */
int hcu_trim(hcu_ctx *ctx);

/* ---- memory ------------------------------------------------------------ */
/* managed (unified) memory backs the ndarray that Mapper.create() returns
 * (heracles/healpy.py:124-142): the Field layer mutates that array in place
 * with numpy (heracles/fields.py:296,304,373,446,548), so it must be host
 * addressable while the scatter kernels update it on the device. */
int hcu_malloc_managed(hcu_ctx *ctx, size_t bytes, void **ptr);
int hcu_malloc_device(hcu_ctx *ctx, size_t bytes, void **ptr);
int hcu_malloc_pinned(hcu_ctx *ctx, size_t bytes, void **ptr);
int hcu_free(hcu_ctx *ctx, void *ptr);
int hcu_prefetch(hcu_ctx *ctx, const void *ptr, size_t bytes, int to_device);
int hcu_memset_zero(hcu_ctx *ctx, void *ptr, size_t bytes);
int hcu_memcpy(hcu_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- catalogue -> map --------------------------------------------------- */
/* hp.ang2pix(nside, lon, lat, lonlat=True[, nest]) -- heracles/healpy.py:157,
 * heracles/catalog/filters.py:91-94.  lon/lat in degrees.  ipix[j] = -1 for
 * rows whose latitude is outside [-90, 90] or not finite.  Synchronous. */
int hcu_ang2pix(hcu_ctx *ctx, int64_t nside, int scheme, const double *lon,
                const double *lat, int64_t n, int64_t *ipix);

/* HealpixMapper.map_values = hp.ang2pix + numba `_map`
 * (heracles/healpy.py:144-160, :58-65): for each row j and each of the nv
 * value rows v:  maps[v*map_stride + ipix_j] += values[v*value_stride + j].
 * lon, lat, values: "any" pointers; maps: device or managed memory.
 * Rows with invalid latitude are skipped and counted (hcu_bad_rows). */
int hcu_map_values(hcu_ctx *ctx, int64_t nside, int scheme, const double *lon,
                   const double *lat, const double *values,
                   int64_t value_stride, int nv, int64_t n, double *maps,
                   int64_t map_stride, int flags);
/* One catalogue page -> the position map AND the shear map of a tomographic bin in ONE pass.  The reference's
 * Positions and Shears fields read the same lon / lat / weight columns of every page
 * (heracles/fields.py:262-271 and :420-433) and call HealpixMapper.map_values twice; here ang2pix runs once
 * and five columns instead of seven cross PCIe:
 *     ipix = ang2pix(lon, lat);  pos[ipix] += w;  she[ipix] += w g1;  she[she_stride + ipix] += w g2
 * (re, im = w*re, w*im: fields.py:426).  w == NULL: unit weights (fields.py:265,425); pos or she may be NULL.
 * Rows with w == 0 are left out of the shear map and its sums (page.delete(page[wcol] == 0), fields.py:420-421);
 * rows with a NaN in a used column are skipped and counted (CatalogPage.get raises on them, catalog/base.py:114-125).
 * stats: NULL or device / managed float64[8], accumulated (+=) -- the running sums the Field layer keeps per page:
 *   [0] rows added to pos  [1] sum w  [2] sum w^2              (Positions: ngal, wmean, w2mean, fields.py:269-271)
 *   [3] rows added to she  [4] sum w  [5] sum w^2  [6] sum w^2 (g1^2 + g2^2)   (Shears: ..., var, fields.py:430-433)
 *   [7] rows with NaN.
 * Column pointers are "any" pointers (pageable pages go through the pinned staging slots). */
int hcu_map_page(hcu_ctx *ctx, int64_t nside, int scheme, const double *lon, const double *lat,
                 const double *w, const double *g1, const double *g2, int64_t n, double *pos,
                 double *she, int64_t she_stride, double *stats);
/* rows skipped since the last call of this function (synchronises) */
int hcu_bad_rows(hcu_ctx *ctx, int64_t *count);

/* in-place map arithmetic used by the Field layer on the created map
 * (heracles/fields.py:296 `pos /= nbar`, :304 `pos -= vis`, :446 `val /= wbar`) */
int hcu_scale(hcu_ctx *ctx, double *x, int64_t n, double a);          /* x *= a        */
int hcu_divide(hcu_ctx *ctx, double *x, int64_t n, double a);         /* x /= a        */
int hcu_axpy(hcu_ctx *ctx, double *y, const double *x, double a, int64_t n); /* y += a x */
int hcu_add_scalar(hcu_ctx *ctx, double *x, int64_t n, double a);     /* x += a        */
int hcu_multiply(hcu_ctx *ctx, double *out, const double *a, const double *b, int64_t n); /* out = a b (masks) */
/* out = (jk_map == region) ? in : 0: the jackknife region maps of DICES (heracles/dices/jackknife.py,
 * _get_region_maps: deepcopy + `_map *= (jk_map == float(jk))` per region) without leaving the device */
int hcu_region_select(hcu_ctx *ctx, double *out, const double *in, const double *jk_map, double region, int64_t n);

/* hp.reorder(map, r2n / n2r): RING <-> NEST order of a whole map (out of place; device or managed memory).
 * The transform works on RING maps; a mapper configured for NEST maps reorders before hcu_map2alm. */
int hcu_reorder(hcu_ctx *ctx, int64_t nside, const double *in, double *out, int to_nest);
/* hp.ud_grade(map, nside_out) -- heracles/healpy.py:205-209 (RING in, RING out, mean preserving) */
int hcu_ud_grade(hcu_ctx *ctx, int64_t nside_in, const double *in,
                 int64_t nside_out, double *out);

/* ---- map -> alm ---------------------------------------------------------- */
/* hp.map2alm(maps, lmax, pol, iter) + hp.almxfl -- heracles/healpy.py:183-196.
 *   spin 0: nmaps independent scalar maps  -> nmaps alm rows
 *   spin 2: nmaps must be even; rows are (Q,U) pairs -> (E,B) alm rows (the
 *           reference's zero T map, healpy.py:174-178,198-199, is not computed)
 *   ring_weights: NULL or float64[2*nside] multiplying 4pi/npix per ring pair
 *   pixel_weights: NULL or float64[npix] multiplying the map before analysis
 *   niter: Jacobi refinement steps (healpy default iter=3), alm += A(map - S(alm))
 *   fl: NULL or float64[lmax+1]; alm[l,m] *= fl[l] (pixel-window deconvolution)
 * maps / alm: device or managed memory ("any" for maps; host maps are copied). */
int hcu_map2alm(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps,
                const double *maps, int64_t map_stride,
                const double *ring_weights, const double *pixel_weights,
                int niter, const double *fl, void *alm, int64_t alm_stride);

/* How pixel_weights enter an iterated transform (niter > 0):
 *   per_pass = 0 (default, healpy's use_pixel_weights=True): the map is multiplied by the weights once and the
 *                Jacobi iterations run on the weighted map with unit weights, alm += A(W map - S(alm));
 *   per_pass = 1: the weights are part of the quadrature of every analysis pass, alm += A(W (map - S(alm))). */
int hcu_set_weights_mode(hcu_ctx *ctx, int per_pass);

/* Same transform for rows that are NOT contiguous in memory: maps[c] / alm[c]
 * are per-row pointers.  This is what lets one call batch the maps of many
 * fields (heracles/mapping.py:151-171 transforms them one at a time) so that
 * the Legendre recursion is shared by up to 12 maps (spin 0) / 4 fields (spin 2). */
int hcu_map2alm_many(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps,
                     const double *const *maps, const double *ring_weights,
                     const double *pixel_weights, int niter, const double *fl,
                     void *const *alm);

/* hp.alm2map, the synthesis used inside map2alm's iterations (also exported) */
int hcu_alm2map(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int nmaps,
                const void *alm, int64_t alm_stride, double *maps,
                int64_t map_stride);

/* Staged transform for the multi-GPU path (no reference counterpart: the
 * reference is single process; these are the four halves of hp.map2alm /
 * hp.alm2map, heracles/healpy.py:183-189, that heracles_b200/dist.py strings
 * together with NCCL exchanges in between).  Ring pairs rp = 0..2 nside-1
 * (north ring rp+1 with its southern mirror; rp = 2 nside - 1 is the equator).
 * Device pointers only; asynchronous on the context's stream.
 *   "phase" arrays are float64[nm][rp_hi - rp_lo][ncomp][4]; row r belongs to
 *   m = mlist[r] (NULL: m = r).  Analysis direction: (re, im) of north+south and
 *   north-south, multiplied by the quadrature weight and exp(-i m phi0).
 *   Synthesis direction: (re, im) of the northern and of the southern ring.
 *   hcu_map2phase  ring FFTs of ring pairs [rp_lo, rp_hi) of full-size maps.
 *   hcu_phase2alm  Legendre analysis of those ring pairs for the listed m;
 *                  ACCUMULATES (+=) into alm at the global (l, m) positions.
 *   hcu_alm2phase  Legendre synthesis for the listed m on ring pairs [rp_lo, rp_hi).
 *   hcu_phase2map  inverse ring FFTs; mpos[m] = row of m in phase (NULL: m,
 *                  negative: absent); writes only the pixels of those rings.
 * ncomp <= hcu_legendre_batch_size(spin) per call (12 for spin 0, 8 for spin 2). */
int hcu_legendre_batch_size(int spin);
int hcu_map2phase(hcu_ctx *ctx, int64_t nside, int lmax, int ncomp,
                  const double *maps, int64_t map_stride,
                  const double *ring_weights, int64_t rp_lo, int64_t rp_hi,
                  const int32_t *mlist, int nm, double *phase);
int hcu_phase2alm(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                  const double *phase, const int32_t *mlist, int nm,
                  int64_t rp_lo, int64_t rp_hi, const double *fl, void *alm,
                  int64_t alm_stride);
int hcu_alm2phase(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                  const void *alm, int64_t alm_stride, const int32_t *mlist, int nm,
                  int64_t rp_lo, int64_t rp_hi, double *phase);
/* The same two Legendre stages over SEVERAL consecutive ring-pair blocks in ONE launch: block b
 * covers [rp_bounds[b], rp_bounds[b+1]) and phase is the concatenation of the per-block arrays
 * float64[nm][rp_bounds[b+1] - rp_bounds[b]][ncomp][4] -- exactly what an all-to-all of the
 * per-rank blocks delivers (heracles_b200/dist.py).  nblocks <= 16. */
int hcu_phase2alm_blocks(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                         const double *phase, const int32_t *mlist, int nm, int nblocks,
                         const int64_t *rp_bounds, const double *fl, void *alm,
                         int64_t alm_stride);
int hcu_alm2phase_blocks(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp,
                         const void *alm, int64_t alm_stride, const int32_t *mlist, int nm,
                         int nblocks, const int64_t *rp_bounds, double *phase);
int hcu_phase2map(hcu_ctx *ctx, int64_t nside, int lmax, int ncomp, const double *phase,
                  const int32_t *mpos, int64_t rp_lo, int64_t rp_hi, double *maps,
                  int64_t map_stride);
/* The exchange between the ring-distributed and the m-distributed stage WITHOUT a collective: the producing kernel
 * writes every row straight into the buffer of the rank that consumes it, over NVLink peer memory (the reference is a
 * single process, heracles/healpy.py:183-189; north_star's "all-to-all to an m-distributed Legendre stage").
 *   hcu_map2phase_peers  hcu_map2phase whose rows [row_start[d], row_start[d+1]) -- the m owned by rank d -- go to
 *                        dest_base[d] + ((row - row_start[d]) * (rp_hi - rp_lo) + rp - rp_lo) * ncomp * 4, i.e. into
 *                        this rank's block of rank d's blocked phase array (the layout hcu_phase2alm_blocks reads).
 *   hcu_alm2phase_peers  hcu_alm2phase_blocks whose block b is written to block_out[b]: this rank's rows of rank
 *                        b's phase array (the layout hcu_phase2map reads).
 *   hcu_ipc_export / hcu_ipc_open / hcu_ipc_close  share a hcu_malloc_device buffer with the other processes of the
 *                        node (cudaIpc*; handle64 is an opaque 64-byte token). */
int hcu_map2phase_peers(hcu_ctx *ctx, int64_t nside, int lmax, int ncomp, const double *maps,
                        int64_t map_stride, const double *ring_weights, int64_t rp_lo, int64_t rp_hi,
                        const int32_t *mlist, int nm, int ndest, const int32_t *row_start,
                        double *const *dest_base);
int hcu_alm2phase_peers(hcu_ctx *ctx, int64_t nside, int lmax, int spin, int ncomp, const void *alm,
                        int64_t alm_stride, const int32_t *mlist, int nm, int nblocks,
                        const int64_t *rp_bounds, double *const *block_out);
int hcu_ipc_export(hcu_ctx *ctx, const void *ptr, void *handle64);
int hcu_ipc_open(hcu_ctx *ctx, const void *handle64, void **ptr);
int hcu_ipc_close(hcu_ctx *ctx, void *ptr);

/* ---- catalogue -> alm without pixels --------------------------------------- */
/* DiscreteMapper.map_values (heracles/ducc.py:92-133): alm[c] += ducc0.sht.adjoint_synthesis_general(map=values,
 * spin, lmax, loc=(radians(90 - lat), radians(lon % 360))), i.e. alm_lm += sum_i v_i conj(sY_lm(theta_i, phi_i)),
 * summed EXACTLY by the Legendre analysis kernels with every point as a ring of its own (O(npts lmax^2); ducc uses a
 * NUFFT with epsilon 1e-12).  lon / lat in degrees, values[c * value_stride + i], host or device; alm rows
 * (complex128, healpy order) device accessible; spin 2: rows are (Q, U) pairs -> (E, B). */
int hcu_points2alm(hcu_ctx *ctx, int lmax, int spin, int ncomp, int64_t npts, const double *lon,
                   const double *lat, const double *values, int64_t value_stride, void *alm,
                   int64_t alm_stride);

/* ---- alm -> Cl ------------------------------------------------------------ */
/* alm2cl(alm, alm2, lmax=lmax) -- heracles/twopoint.py:63-101, as a block:
 *   cl[(i*nb + j)*(lout+1) + l] = sum_m (2 - delta_m0) Re(a_i,lm conj b_j,lm) / (2l+1),
 *   lout = min(lmax_out, lmax_a, lmax_b); a, b may have different lmax. */
int hcu_alm2cl(hcu_ctx *ctx, int na, const void *a, int64_t stride_a,
               int lmax_a, int nb, const void *b, int64_t stride_b, int lmax_b,
               int lmax_out, double *cl);
/* The symmetric block of ALL pairs of nrows alm rows that live in SEPARATE allocations (rows[i]: device or managed
 * complex128[nalm], all with the same lmax): cl[(i*nrows + j)*(lout+1) + l].  angular_power_spectra calls alm2cl
 * once per pair of alm arrays (heracles/twopoint.py:198-243); this computes all of them in one launch.  nrows <= 64. */
int hcu_alm2cl_rows(hcu_ctx *ctx, int nrows, const void *const *rows, int lmax, int lmax_out, double *cl);
/* the same sum restricted to m = m_offset (mod m_step): the partial spectra of one rank of the
 * multi-GPU path, whose alm are m-distributed (summing them over the ranks gives hcu_alm2cl) */
int hcu_alm2cl_mslice(hcu_ctx *ctx, int na, const void *a, int64_t stride_a,
                      int lmax_a, int nb, const void *b, int64_t stride_b, int lmax_b,
                      int lmax_out, int m_step, int m_offset, double *cl);

/* ---- introspection --------------------------------------------------------- */
/* The Legendre kernels skip the "dead zone" l0 <= l < l_start(m, ring) -- where lambda_lm is below 2^-200 and contributes
 * nothing -- by starting every recursion from a table of start states built once per (nside, lmax, spin) with the same
 * arithmetic (20 bytes per (m, ring pair, chain); not built when it would take more than a quarter of the free memory).
 * Results are bit-identical with and without it; hcu_set_start_table(ctx, 0) disables it (A/B timing, tests). */
int hcu_set_start_table(hcu_ctx *ctx, int enabled);
/* stage timing is off by default (it costs a host wait per Legendre batch); hcu_set_timing(ctx, 1) enables it */
int hcu_set_timing(hcu_ctx *ctx, int enabled);
/* device milliseconds of the stages of the last hcu_map2alm / hcu_alm2map call
 * (CUDA events on the context stream): [0] ring FFT stage, [1] Legendre stage,
 * [2] synthesis Legendre, [3] synthesis FFT.  Synchronises. */
int hcu_last_sht_timing(hcu_ctx *ctx, float ms[4]);
/* executed Legendre work of the last analysis call: number of (l, ring-pair)
 * cells advanced by the recursion and the number that were accumulated */
int hcu_last_sht_work(hcu_ctx *ctx, double *cells_recursed, double *cells_accumulated);
/* peak FP64 FMA rate of this device measured with a register-resident DFMA
 * loop (flop/s); used as the roofline denominator of the Legendre stage */
int hcu_measure_fp64_peak(hcu_ctx *ctx, double *flops);

#ifdef __cplusplus
}
#endif
#endif /* HERACLES_CUDA_H */
