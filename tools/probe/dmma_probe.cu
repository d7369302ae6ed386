// Does B200 run FP64 tensor MMAs (mma.sync m8n8k4 f64, SASS DMMA) on hardware separate from the FP64
// vector pipe?  Three launches: DFMA only, DMMA only, half the warps each.  If the mixed launch finishes
// in about max(t_dfma, t_dmma)/... the units are separate; if it takes the sum they share the datapath.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE> __global__ void probe(double *sink, int iters) {
  const int warp = threadIdx.x >> 5;
  const bool tensor = MODE == 1 || (MODE == 2 && (warp & 1));
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i * 1e-3;
  if (tensor) {
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);
    }
  } else {
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(c[i], b, a);
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r += c[i];
  if (r == 123.456) sink[0] = r;
}

template <int MODE> float run(int iters) {
  double *sink; cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<148 * 2, 512>>>(sink, iters);
  cudaEventRecord(e0);
  probe<MODE><<<148 * 2, 512>>>(sink, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(sink);
  return ms;
}

int main() {
  const int iters = 1 << 14;
  // per thread per iteration: DFMA mode 16 FMA; DMMA mode 8 mma x 256 FMA / 32 threads = 64 FMA
  const double threads = 148.0 * 2 * 512;
  float t0 = run<0>(iters), t1 = run<1>(iters), t2 = run<2>(iters);
  double f0 = threads * iters * 16 * 2 / (t0 * 1e-3) / 1e12;
  double f1 = threads * iters * 64 * 2 / (t1 * 1e-3) / 1e12;
  double f2 = threads * iters * (8 + 32) * 2 / (t2 * 1e-3) / 1e12;   // half the warps each
  printf("{\"dfma_only_ms\": %.3f, \"dfma_tflops\": %.2f, \"dmma_only_ms\": %.3f, \"dmma_tflops\": %.2f, "
         "\"mixed_ms\": %.3f, \"mixed_tflops\": %.2f, \"mixed_if_shared_ms\": %.3f, \"mixed_if_separate_ms\": %.3f}\n",
         t0, f0, t1, f1, t2, f2, 0.5f * (t0 + t1), 0.5f * (t0 > t1 ? t0 : t1));
  return 0;
}
