// tools/switch_peak.cu -- does alternating DMMA and DFMA at fine granularity cost FP64-pipe cycles?
// Per iteration 16 DMMAs and 16 DFMAs, issued as groups of G DMMAs followed by G DFMAs (G = 1, 2, 4, 8, 16).
// All accumulators independent, so only the instruction mix pattern changes.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dfma(double &f, double a, double b) {
  asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f) : "d"(a), "d"(b));
}
template <int G>
__global__ void kern(double *out, int iters) {
  double c[32], f[16];
  for (int i = 0; i < 32; ++i) c[i] = threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-7 + i;
  double a = 1.0000001, b = 0.999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 16 / G; ++g) {
#pragma unroll
      for (int i = 0; i < G; ++i) dmma(c[2 * (g * G + i)], c[2 * (g * G + i) + 1], a, b);
#pragma unroll
      for (int i = 0; i < G; ++i) dfma(f[g * G + i], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 32; ++i) s += c[i];
  for (int i = 0; i < 16; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int G>
void run(int warps) {
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 1 << 13;
  kern<G><<<148, 32 * warps>>>(out, 64);
  cudaEventRecord(e0);
  kern<G><<<148, 32 * warps>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double clk = ms * 1e-3 * 1.965e9 / iters;  // per iteration per warp-slot
  double ideal = (16 * 16 + 16 * 2) * (warps / 4.0);
  printf("G=%2d warps/SM=%d: %7.1f clk per iteration (pipe-ideal %5.0f) -> %4.1f %% of the pipe\n", G, warps, clk, ideal, 100 * ideal / clk);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w); }
  return 0;
}
