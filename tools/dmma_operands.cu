// tools/dmma_operands.cu -- does the FP64 tensor pipe reach its peak when every DMMA has DIFFERENT A / B operands?
// (tools/dmma_peak.cu feeds all DMMAs the same two registers.)  Variants per iteration of 32 DMMAs:
//   0  same a, same b              1  same a, 16 different b registers
//   2  8 different a, 16 different b (a used for 4 DMMAs)   3  like 2 but a re-loaded from shared memory (LDS.64) each time
//   4  like 3 plus a dependent DFMA chain of 16 links and 16 independent DFMAs interleaved (the Legendre mix)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void kern(double *out, int iters) {
  __shared__ double tile[32 * 64];
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) tile[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double c[8];
  for (int i = 0; i < 8; ++i) c[i] = threadIdx.x * 1e-9 + i;
  double b[16], a[8];
  for (int i = 0; i < 16; ++i) b[i] = 0.999999 + 1e-7 * i + 1e-9 * lane;
  for (int i = 0; i < 8; ++i) a[i] = 1.0000001 + 1e-7 * i + 1e-9 * lane;
  double q0 = 1.0, q1 = 0.5, x = 0.3 + 1e-3 * lane;
  const double *tp = tile + (warp & 3) * 256 + lane;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      double a0, a1;
      if (MODE == 0) { a0 = a[0]; a1 = a[0]; }
      else if (MODE == 1) { a0 = a[0]; a1 = a[0]; }
      else if (MODE == 2) { a0 = a[kk]; a1 = a[(kk + 3) & 7]; }
      else { a0 = tp[32 * kk]; a1 = tp[32 * kk + 1024]; }
      if (MODE == 4) {
        const double ax = fma(a[kk], x, b[kk]);
        const double nw = fma(ax, q1, -q0);
        q0 = q1; q1 = nw;
      }
      const double b0 = MODE == 0 ? b[0] : b[2 * kk], b1 = MODE == 0 ? b[0] : b[2 * kk + 1];
      dmma(c[0], c[1], a0, b0);
      dmma(c[2], c[3], a0, b1);
      if (MODE == 4) {
        const double ax = fma(a[(kk + 1) & 7], x, b[kk + 8]);
        const double nw = fma(ax, q1, -q0);
        q0 = q1; q1 = nw;
      }
      dmma(c[4], c[5], a1, b0);
      dmma(c[6], c[7], a1, b1);
    }
  }
  double s = q1;
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(int warps) {
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 1 << 13;
  kern<MODE><<<148, 32 * warps>>>(out, 64);
  cudaEventRecord(e0);
  kern<MODE><<<148, 32 * warps>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 2.0 * 256 * 32 * (double)iters * 148 * warps;
  double clk = ms * 1e-3 * 1.965e9 / iters / (warps / 4.0) / 32;
  printf("mode %d warps/SM=%2d: %8.3f ms  %6.2f TF/s DMMA   %5.1f clk per DMMA per SM sub-partition\n", MODE, warps, ms, fl / ms / 1e9, clk);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) { run<0>(w); run<1>(w); run<2>(w); run<3>(w); run<4>(w); }
  return 0;
}
