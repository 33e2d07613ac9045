// tools/red_peak.cu -- microbenchmark: RED.ADD.F64 throughput to global memory for the
// access pattern of the Legendre analysis flush (each warp adds 32 consecutive doubles;
// `dup` warps hit the same 256-byte segment back to back).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void red_kernel(double *buf, long nseg, int iters, int dup) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long group = warp / dup;
  unsigned long long s = group * 0x9E3779B97F4A7C15ull + 12345;
  for (int it = 0; it < iters; ++it) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    long seg = (long)((s >> 20) % (unsigned long long)nseg);
    atomicAdd(buf + seg * 32 + lane, 1.0);
  }
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const long nseg = 1L << 21;  // 512 MB
  double *buf;
  cudaMalloc(&buf, nseg * 32 * sizeof(double));
  cudaMemset(buf, 0, nseg * 32 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int dup : {1, 8}) {
    for (int warps : {8, 16, 32}) {
      int blocks = p.multiProcessorCount * (32 / warps) , iters = 4096;
      red_kernel<<<blocks, warps * 32>>>(buf, nseg, 16, dup);
      cudaEventRecord(e0);
      red_kernel<<<blocks, warps * 32>>>(buf, nseg, iters, dup);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double n = (double)blocks * warps * 32 * iters;
      printf("dup=%d warps/CTA=%2d blocks=%d: %8.3f ms  %8.2f G RED.F64/s  (%7.1f GB/s of operands)\n", dup, warps,
             blocks, ms, n / ms / 1e6, n * 8 / ms / 1e6);
    }
  }
  return 0;
}
