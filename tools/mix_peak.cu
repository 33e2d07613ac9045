// tools/mix_peak.cu -- how does a dependent DFMA chain (the Legendre recursion of the producer warps)
// progress while other warps of the same SM keep the FP64 pipe full of DMMAs?
// 12 warps per CTA, 1 CTA per SM: warps 0-3 run ILP independent dependent chains of DFMAs,
// warps 4-11 issue independent DMMAs until the producers are done.  Prints the clocks per
// dependent DFMA step seen by the chain warps and the DMMA rate achieved meanwhile.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP, int NCONS>
__global__ void __launch_bounds__(384, 1) mix_kernel(double *out, long long *clk, int steps, int with_dmma) {
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) done = 0;
  __syncthreads();
  if (warp < 4) {
    double cur[ILP], prev[ILP];
    for (int i = 0; i < ILP; ++i) { cur[i] = 1.0 + threadIdx.x * 1e-9 + i; prev[i] = 0.5; }
    const double a = 1.0000001;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        double nw = fma(a, cur[i], -prev[i]);
        prev[i] = cur[i];
        cur[i] = nw;
      }
    }
    long long t1 = clock64();
    double sum = 0;
    for (int i = 0; i < ILP; ++i) sum += cur[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 16 + warp] = t1 - t0;
    __syncwarp();
    if (threadIdx.x == 0) done = 1;
  } else if (warp < 4 + NCONS) {
    double c[24];
    for (int i = 0; i < 24; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double a = 1.0000001, b = 0.999999;
    long long n = 0;
    long long t0 = clock64();
    while (with_dmma && !done) {
#pragma unroll
      for (int i = 0; i < 12; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
      n += 12;
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 24; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) { clk[blockIdx.x * 16 + warp] = t1 - t0; clk[148 * 16 + blockIdx.x * 16 + warp] = n; }
  }
}

template <int ILP, int NCONS>
void run(int with_dmma) {
  double *out; long long *clk;
  cudaMalloc(&out, sizeof(double) * 148 * 384);
  cudaMalloc(&clk, sizeof(long long) * 148 * 32);
  cudaMemset(clk, 0, sizeof(long long) * 148 * 32);
  const int steps = 20000;
  mix_kernel<ILP, NCONS><<<148, 384>>>(out, clk, steps, with_dmma);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double pc = 0, cc = 0, nd = 0;
  for (int w = 0; w < 4; ++w) pc += h[w];
  for (int w = 4; w < 4 + NCONS; ++w) { cc += h[w]; nd += h[148 * 16 + w]; }
  pc /= 4;
  printf("ILP=%d cons=%d dmma=%d: %.1f clk per chain step (%.1f clk per DFMA instr); DMMA pipe use %.1f%%\n", ILP, NCONS,
         with_dmma, pc / steps, pc / steps / ILP, NCONS ? 100.0 * nd * 16.0 / 4.0 / (cc / NCONS) : 0.0);
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<1, 8>(0); run<1, 8>(1); run<2, 8>(1); run<4, 8>(1); run<8, 8>(1);
  run<1, 4>(1); run<2, 4>(1); run<4, 4>(1); run<8, 4>(1);
  return 0;
}
