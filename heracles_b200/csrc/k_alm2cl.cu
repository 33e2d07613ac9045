// k_alm2cl.cu -- angular power spectra from alm, all pairs of a block at once.
//
// Replaces alm2cl (heracles/twopoint.py:63-101):
//   C_l = [ Re(a_l0 b_l0*) + 2 sum_{m=1..l} Re(a_lm b_lm*) ] / (2l+1)
// (the reference evaluates this as a running mean over m; the two differ by
// rounding only, <= 2e-16 of sqrt(C_l^aa C_l^bb)).
//
// HBM-bound: every a_lm / b_lm is read once per tile of TA x TB spectra.
// One thread owns one l and walks m (consecutive lanes = consecutive l =
// consecutive addresses in the m-major layout); the m range is split across
// blockIdx.y and partial sums are combined with RED.ADD.F64.
#include "hcu_common.cuh"

namespace {

constexpr int TA = 4, TB = 4;

constexpr int CL_MAX_ROWS = 64;
struct ClRows {
  const double2 *p[CL_MAX_ROWS];
};

struct ClArgs {
  const double2 *a, *b;
  i64 sa, sb;  // strides in complex elements
  int na, nb, lmax_a, lmax_b, lout, msplit;
  int mstep, moff;  // only m = moff (mod mstep) contribute (m-distributed alm of the multi-GPU path)
  double *cl;
  bool same;  // a and b are the same array with the same stride: only j >= i is computed
};

// ROWS: the rows of `a` (= `b`) are separate allocations given by a pointer table (hcu_alm2cl_rows)
template <bool ROWS>
__global__ void __launch_bounds__(128) alm2cl_kernel(ClArgs p, ClRows rows) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int ia0 = (blockIdx.z / ((p.nb + TB - 1) / TB)) * TA;
  const int ib0 = (blockIdx.z % ((p.nb + TB - 1) / TB)) * TB;
  if (p.same && ib0 + TB - 1 < ia0) return;  // strictly lower tile of a symmetric block
  if (l > p.lout) return;
  // m segment of this block
  const int mseg = (p.lout + p.msplit) / p.msplit;
  const int m0 = blockIdx.y * mseg;
  int m1 = m0 + mseg - 1;
  if (m1 > l) m1 = l;
  if (m0 > l) return;
  double acc[TA][TB];
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j] = 0.0;
  const int mfirst = m0 + ((p.moff - m0) % p.mstep + p.mstep) % p.mstep;
  for (int m = mfirst; m <= m1; m += p.mstep) {
    const i64 ia = (i64)m * (2 * p.lmax_a + 1 - m) / 2 + l;
    const i64 ib = (i64)m * (2 * p.lmax_b + 1 - m) / 2 + l;
    // the reference's m = 0 term is alm.real * alm2.real (twopoint.py:88): imaginary
    // parts of a_l0 (zero for real fields) are ignored
    const double wgt = (m == 0) ? 1.0 : 2.0;
    const double wim = (m == 0) ? 0.0 : 1.0;
    double2 av[TA], bv[TB];
#pragma unroll
    for (int i = 0; i < TA; ++i)
      av[i] = (ia0 + i < p.na) ? (ROWS ? rows.p[ia0 + i][ia] : p.a[(i64)(ia0 + i) * p.sa + ia]) : make_double2(0., 0.);
#pragma unroll
    for (int j = 0; j < TB; ++j)
      bv[j] = (ib0 + j < p.nb) ? (ROWS ? rows.p[ib0 + j][ib] : p.b[(i64)(ib0 + j) * p.sb + ib]) : make_double2(0., 0.);
#pragma unroll
    for (int i = 0; i < TA; ++i)
#pragma unroll
      for (int j = 0; j < TB; ++j)
        acc[i][j] += wgt * (av[i].x * bv[j].x + wim * (av[i].y * bv[j].y));
  }
  const double norm = 1.0 / (2.0 * l + 1.0);
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const int ga = ia0 + i, gb = ib0 + j;
      if (ga < p.na && gb < p.nb) {
        if (p.same && gb < ga) continue;
        const double v = acc[i][j] * norm;
        atomicAdd(p.cl + ((i64)ga * p.nb + gb) * (p.lout + 1) + l, v);
        if (p.same && gb != ga)
          atomicAdd(p.cl + ((i64)gb * p.nb + ga) * (p.lout + 1) + l, v);
      }
    }
}

}  // namespace

static int alm2cl_impl(hcu_ctx *ctx, int na, const void *a, int64_t stride_a, int lmax_a, int nb,
                       const void *b, int64_t stride_b, int lmax_b, int lmax_out, int mstep, int moff,
                       double *cl) {
  HCU_ARG(ctx && a && b && cl, "hcu_alm2cl: null pointer");
  HCU_ARG(na >= 1 && nb >= 1 && lmax_a >= 0 && lmax_b >= 0 && lmax_out >= 0, "hcu_alm2cl: sizes");
  HCU_ARG(mstep >= 1 && moff >= 0 && moff < mstep, "hcu_alm2cl: 0 <= m_offset < m_step");
  int lout = lmax_out;
  if (lmax_a < lout) lout = lmax_a;
  if (lmax_b < lout) lout = lmax_b;
  HCU_CUDA(cudaMemsetAsync(cl, 0, sizeof(double) * (size_t)na * nb * (lout + 1), ctx->stream));
  ClArgs p;
  p.a = (const double2 *)a;
  p.b = (const double2 *)b;
  p.sa = stride_a;
  p.sb = stride_b;
  p.na = na;
  p.nb = nb;
  p.lmax_a = lmax_a;
  p.lmax_b = lmax_b;
  p.lout = lout;
  p.mstep = mstep;
  p.moff = moff;
  p.same = (a == b) && (stride_a == stride_b) && (na == nb) && (lmax_a == lmax_b);
  p.cl = cl;
  // enough m segments to fill the device a few times over
  const int lblocks = (lout + 128) / 128;
  const int tiles = ((na + TA - 1) / TA) * ((nb + TB - 1) / TB);
  int msplit = (ctx->num_sms * 8 + lblocks * tiles - 1) / (lblocks * tiles);
  if (msplit < 1) msplit = 1;
  if (msplit > lout + 1) msplit = lout + 1;
  if (msplit > 64) msplit = 64;
  p.msplit = msplit;
  dim3 grid(lblocks, msplit, tiles);
  alm2cl_kernel<false><<<grid, 128, 0, ctx->stream>>>(p, ClRows());
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

// the symmetric block of ALL pairs of `nrows` alm rows that live in separate allocations: one launch instead of one
// hcu_alm2cl per pair of arrays (heracles/twopoint.py:198-243 calls alm2cl once per pair of alm arrays)
extern "C" int hcu_alm2cl_rows(hcu_ctx *ctx, int nrows, const void *const *rows, int lmax, int lmax_out, double *cl) {
  HCU_ARG(ctx && rows && cl, "hcu_alm2cl_rows: null pointer");
  HCU_ARG(nrows >= 1 && nrows <= CL_MAX_ROWS, "hcu_alm2cl_rows: 1 <= nrows <= 64");
  HCU_ARG(lmax >= 0 && lmax_out >= 0, "hcu_alm2cl_rows: lmax");
  const int lout = lmax_out < lmax ? lmax_out : lmax;
  HCU_CUDA(cudaMemsetAsync(cl, 0, sizeof(double) * (size_t)nrows * nrows * (lout + 1), ctx->stream));
  ClRows r;
  for (int i = 0; i < CL_MAX_ROWS; ++i) r.p[i] = i < nrows ? (const double2 *)rows[i] : nullptr;
  for (int i = 0; i < nrows; ++i) HCU_ARG(rows[i], "hcu_alm2cl_rows: null row");
  ClArgs p;
  p.a = p.b = nullptr;
  p.sa = p.sb = 0;
  p.na = p.nb = nrows;
  p.lmax_a = p.lmax_b = lmax;
  p.lout = lout;
  p.mstep = 1;
  p.moff = 0;
  p.same = true;
  p.cl = cl;
  const int lblocks = (lout + 128) / 128;
  const int tiles = ((nrows + TA - 1) / TA) * ((nrows + TB - 1) / TB);
  int msplit = (ctx->num_sms * 8 + lblocks * tiles - 1) / (lblocks * tiles);
  if (msplit < 1) msplit = 1;
  if (msplit > lout + 1) msplit = lout + 1;
  if (msplit > 64) msplit = 64;
  p.msplit = msplit;
  dim3 grid(lblocks, msplit, tiles);
  alm2cl_kernel<true><<<grid, 128, 0, ctx->stream>>>(p, r);
  HCU_LAUNCH_CHECK(ctx);
  return HCU_OK;
}

extern "C" int hcu_alm2cl(hcu_ctx *ctx, int na, const void *a, int64_t stride_a,
                          int lmax_a, int nb, const void *b, int64_t stride_b,
                          int lmax_b, int lmax_out, double *cl) {
  return alm2cl_impl(ctx, na, a, stride_a, lmax_a, nb, b, stride_b, lmax_b, lmax_out, 1, 0, cl);
}

// partial spectra from the m = m_offset (mod m_step) only: what a rank of the multi-GPU path owns
extern "C" int hcu_alm2cl_mslice(hcu_ctx *ctx, int na, const void *a, int64_t stride_a,
                                 int lmax_a, int nb, const void *b, int64_t stride_b,
                                 int lmax_b, int lmax_out, int m_step, int m_offset, double *cl) {
  return alm2cl_impl(ctx, na, a, stride_a, lmax_a, nb, b, stride_b, lmax_b, lmax_out, m_step, m_offset, cl);
}
