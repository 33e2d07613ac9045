// tools/dmma_lat.cu -- DMMA (mma.sync.m8n8k4.f64) throughput vs number of independent accumulator
// chains per warp and warps per SM sub-partition: what dependency distance the Legendre kernels need.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void kern(double *out, int iters) {
  double c[2 * NACC];
  for (int i = 0; i < 2 * NACC; ++i) c[i] = threadIdx.x * 1e-9 + i;
  double a = 1.0000001, b = 0.999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16 / NACC; ++r)
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
  }
  double s = 0;
  for (int i = 0; i < 2 * NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(int warps) {
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 1 << 13;
  kern<NACC><<<148, 32 * warps>>>(out, 64);
  cudaEventRecord(e0);
  kern<NACC><<<148, 32 * warps>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = 16.0 * iters * warps * 148;
  double clk_per = ms * 1e-3 * 1.965e9 / (16.0 * iters) ;  // per DMMA per warp
  printf("chains=%d warps/SM=%2d: %6.2f TF/s   %.1f clk per DMMA per warp\n", NACC, warps, n * 512 / ms / 1e9, clk_per);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w); }
  return 0;
}
