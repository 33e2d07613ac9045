// tools/dmma_peak.cu -- microbenchmark: FP64 tensor-core (mma.sync m8n8k4 f64) vs DFMA throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 dmma only, 1 dfma only, 2 both interleaved
__global__ void kern(double *out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
  double f[8];
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-7 + i;
  double a = 1.0000001, b = 0.999999;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  for (int i = 0; i < 8; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int warps, int sms) {
  int threads = 32 * warps, blocks = sms, iters = 1 << 15;
  double *out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<MODE><<<blocks, threads>>>(out, 100);
  cudaEventRecord(e0);
  kern<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double mma_fl = (MODE != 1) ? 2.0 * 256 * 8 * (double)iters * blocks * warps : 0;
  double fma_fl = (MODE != 0) ? 2.0 * 8 * 32 * (double)iters * blocks * warps : 0;
  printf("%-10s warps/SM=%2d  %8.3f ms  dmma %7.2f TF/s  dfma %7.2f TF/s\n", name, warps, ms,
         mma_fl / ms / 1e9, fma_fl / ms / 1e9);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  for (int w : {4, 8, 16, 32}) {
    run<0>("dmma", w, sms);
    run<1>("dfma", w, sms);
    run<2>("both", w, sms);
  }
  return 0;
}
