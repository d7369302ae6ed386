// What does a DFMA cost on B200 as a function of where its operands come from?  Eight independent
// accumulators per thread, 4 warps per sub-partition, register-only:
//   m0  d_i = a_i * b_i + d_i      three distinct register operands per instruction
//   m1  d_i = r   * b_i + d_i      multiplier shared by consecutive instructions (operand-reuse cache)
//   m2  d_i = r   * s   + d_i      both multiplicands shared
//   m3  d_i = a_i * K   + d_i      K a compile-time constant (constant bank / immediate)
//   m4  d_i = a_i + d_i            DADD, two register operands
//   m5  d_i = a_i * 2 - d_i        DFMA with an immediate
//   m6  d_i = a_i * b_i + c_i      three distinct, destination separate (no accumulate)
// Reports SM cycles per FP64 instruction per warp-slot (2.0 = the pipe's issue rate).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operand_probe dfma_operand_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void probe(double *out, long long *cyc, int iters, double seed) {
  double a[8], b[8], c[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = 1.0 + seed * (threadIdx.x + i);
    b[i] = 1.0 - seed * (threadIdx.x + 2 * i);
    c[i] = seed * i;
    d[i] = seed * (i + 3);
  }
  const double r = a[0], s = b[0];
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[i]));
        if (MODE == 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(r), "d"(b[i]));
        if (MODE == 2) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(r), "d"(s));
        if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, 0d3FEFFFFFFFFFFFF0, %0;" : "+d"(d[i]) : "d"(a[i]));
        if (MODE == 4) asm volatile("add.rn.f64 %0, %1, %0;" : "+d"(d[i]) : "d"(a[i]));
        if (MODE == 5) asm volatile("fma.rn.f64 %0, %1, 0d4000000000000000, %0;" : "+d"(d[i]) : "d"(a[i]));
        if (MODE == 6) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d[i]) : "d"(a[i]), "d"(b[i]), "d"(c[i]));
      }
    }
  }
  long long t1 = clock64();
  double q = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) q += d[i] + a[i] + b[i] + c[i];
  if (q == 123.456) out[0] = q;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> double run(int w, int iters) {
  double *out; long long *cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
  probe<MODE><<<148, 128 * w>>>(out, cyc, iters, 1e-9);
  probe<MODE><<<148, 128 * w>>>(out, cyc, iters, 1e-9);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double t = 0; for (int i = 0; i < 148; i++) t += h[i];
  cudaFree(out); cudaFree(cyc);
  return t / 148 / iters / 32 / w;
}

int main() {
  const int it = 1 << 12;
  printf("{\"unit\": \"SM cycles per FP64 instruction per warp-slot at 1 / 2 / 4 warps per sub-partition\",\n");
#define ROW(name, MODE) printf(" \"%s\": [%.2f, %.2f, %.2f],\n", name, run<MODE>(1, it), run<MODE>(2, it), run<MODE>(4, it));
  ROW("m0_three_distinct_accumulate", 0)
  ROW("m1_shared_multiplier", 1)
  ROW("m2_both_multiplicands_shared", 2)
  ROW("m3_constant_multiplicand", 3)
  ROW("m4_dadd", 4)
  ROW("m5_dfma_immediate", 5)
  ROW("m6_three_distinct_separate_destination", 6)
  printf(" \"end\": 0}\n");
  return 0;
}
