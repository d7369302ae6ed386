// What does the FP64 pipe of B200 sustain on the blind rotation's REAL instruction streams, with no memory
// traffic at all?  Register-only loops over the kernels' own butterflies (brs_core.cuh / br_core.cuh):
//   r4      : the 24-FMA radix-4 of the 128-thread kernel (4 complex points per thread)
//   r4x3    : three independent radix-4s per iteration (what digit-interleaving exposes)
//   r8mac   : half radix-8 + two radix-4 + 8 complex MACs = one digit of one thread, constants in registers
//   dft8    : the 64-thread kernel's radix-8 (DADD/DMUL heavy) + twiddle multiplication
// Reports SM cycles per FP64 instruction at 1 / 2 / 4 warps per sub-partition (2.0 = pipe saturated).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../rs_tfhe_b200/csrc -o fp64_stream_probe fp64_stream_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "br_core.cuh"
#include "brs_core.cuh"

using br::mk;

// radix-4 with the FMAs grouped by shared multiplier (operand-reuse cache): same dataflow as brs::r4
__device__ __forceinline__ void r4g(cplx (&u)[4], cplx r1, cplx r2) {
  using brs::fma_;
  double t0 = fma_(r2.x, u[2].x, u[0].x), t1 = fma_(r2.x, u[2].y, u[0].y);
  double t2 = fma_(r2.x, u[3].x, u[1].x), t3 = fma_(r2.x, u[3].y, u[1].y);
  cplx ep, em, fp, fm;
  ep.x = fma_(-r2.y, u[2].y, t0); ep.y = fma_(r2.y, u[2].x, t1);
  fp.x = fma_(-r2.y, u[3].y, t2); fp.y = fma_(r2.y, u[3].x, t3);
  em.x = fma_(2.0, u[0].x, -ep.x); em.y = fma_(2.0, u[0].y, -ep.y);
  fm.x = fma_(2.0, u[1].x, -fp.x); fm.y = fma_(2.0, u[1].y, -fp.y);
  // (ep + r1 fp, ep - r1 fp), (em + q fm, em - q fm), q = -i r1 = (r1.y, -r1.x)
  t0 = fma_(r1.x, fp.x, ep.x); t1 = fma_(r1.x, fp.y, ep.y);
  t2 = fma_(r1.x, fm.y, em.x); t3 = fma_(-r1.x, fm.x, em.y);
  cplx y0, y1;
  y0.x = fma_(-r1.y, fp.y, t0); y0.y = fma_(r1.y, fp.x, t1);
  y1.x = fma_(r1.y, fm.x, t2); y1.y = fma_(r1.y, fm.y, t3);
  u[2] = mk(fma_(2.0, ep.x, -y0.x), fma_(2.0, ep.y, -y0.y));
  u[3] = mk(fma_(2.0, em.x, -y1.x), fma_(2.0, em.y, -y1.y));
  u[0] = y0; u[1] = y1;
}

template <int MODE> __global__ void probe(double *out, long long *cyc, int iters, double seed) {
  cplx y[3][4], c[4], acc[2][4], u8[8];
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int k = 0; k < 4; k++) y[d][k] = mk(seed + threadIdx.x * 1e-6 + k, seed * 0.5 + d);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    c[k] = mk(cos(0.1 * (k + 1) + threadIdx.x), sin(0.1 * (k + 1) + threadIdx.x));
    acc[0][k] = acc[1][k] = mk(0.0, 0.0);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) u8[k] = mk(seed + k, seed - k);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
      brs::r4<false>(y[0], c[0], c[1]);
      brs::r4<false>(y[0], c[2], c[3]);
    } else if (MODE == 1) {
#pragma unroll
      for (int d = 0; d < 3; d++) brs::r4<false>(y[d], c[0], c[1]);
#pragma unroll
      for (int d = 0; d < 3; d++) brs::r4<false>(y[d], c[2], c[3]);
    } else if (MODE == 2) {
      cplx v[4];
      brs::r8_half<false>(u8, c[0], c[1], c[2], v);
      brs::r4<false>(v, c[0], c[1]);
      brs::r4<false>(v, c[2], c[3]);
#pragma unroll
      for (int kd = 0; kd < 4; kd++) {
        br::cfma(acc[0][kd], v[kd], c[kd]);
        br::cfma(acc[1][kd], v[kd], c[3 - kd]);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) { u8[k] = v[k]; u8[4 + k] = acc[0][k]; }
    } else if (MODE == 4) {
      r4g(y[0], c[0], c[1]);
      r4g(y[0], c[2], c[3]);
    } else if (MODE == 5) {
#pragma unroll
      for (int d = 0; d < 3; d++) r4g(y[d], c[0], c[1]);
#pragma unroll
      for (int d = 0; d < 3; d++) r4g(y[d], c[2], c[3]);
    } else {
      br::dft8<false>(u8);
#pragma unroll
      for (int k = 0; k < 8; k++) u8[k] = br::cmul(u8[k], c[k & 3]);
    }
  }
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int k = 0; k < 4; k++) r += y[d][k].x + y[d][k].y;
#pragma unroll
  for (int k = 0; k < 4; k++) r += acc[0][k].x + acc[1][k].y;
#pragma unroll
  for (int k = 0; k < 8; k++) r += u8[k].x + u8[k].y;
  if (r == 123.456) out[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> double run(int w, int iters) {
  double *out; long long *cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
  probe<MODE><<<148, 128 * w>>>(out, cyc, iters, 1e-3);
  probe<MODE><<<148, 128 * w>>>(out, cyc, iters, 1e-3);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; i++) s += h[i];
  cudaFree(out); cudaFree(cyc);
  return s / 148 / iters;
}

int main(int argc, char **argv) {
  const int it = 1 << 12;
  // FP64 instructions per iteration per thread (count them in the SASS: cuobjdump -sass | grep -c)
  const int n0 = argc > 1 ? atoi(argv[1]) : 48, n1 = argc > 2 ? atoi(argv[2]) : 144, n2 = argc > 3 ? atoi(argv[3]) : 120,
            n3 = argc > 4 ? atoi(argv[4]) : 96;
  printf("{\"unit\": \"SM cycles per iteration at 1/2/4 warps per sub-partition; (cycles per FP64 instruction x warps)\",\n");
#define ROW(name, MODE, N)                                                                              \
  {                                                                                                     \
    double a = run<MODE>(1, it), b = run<MODE>(2, it), c = run<MODE>(4, it);                            \
    printf(" \"%s\": {\"fp64_per_iter\": %d, \"cycles\": [%.1f, %.1f, %.1f], \"cycles_per_fp64_per_warp\": [%.2f, %.2f, %.2f]},\n", \
           name, N, a, b, c, a / N, b / N / 2, c / N / 4);                                              \
  }
  ROW("r4_twice", 0, n0)
  ROW("r4_three_digits_twice", 1, n1)
  ROW("digit_r8half_r4_r4_mac", 2, n2)
  ROW("dft8_plus_twiddles", 3, n3)
  ROW("r4_grouped_twice", 4, n0)
  ROW("r4_grouped_three_digits_twice", 5, n1)
  printf(" \"end\": 0}\n");
  return 0;
}
