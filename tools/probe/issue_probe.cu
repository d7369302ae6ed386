// How does the FP64 pipe of B200 share the issue port?  One CTA per SM, W warps per sub-partition, each
// iteration = 16 independent DFMAs interleaved with K instructions of another kind (integer multiply-add,
// logic op, 32-bit move through the FMA pipe, shared-memory load).  Reports SM cycles per iteration:
// if the other instructions hide in the second cycle of each DFMA the time stays at 32 W; if they take
// their own issue cycle it is (32 + K) W.  Also the dependent-issue latency of DFMA (chain of 1) and what
// one warp alone sustains at ILP 2/4/8/16.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { K_NONE, K_IMAD, K_LOP, K_FMOV, K_LDS, K_LDS128, K_FFMA };

template <int KIND> __device__ __forceinline__ void other(unsigned &x, unsigned y, float &f, const char *sm, double &sink) {
  if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y));
  if (KIND == K_LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y));
  if (KIND == K_FMOV) asm volatile("fma.rn.f32 %0, %0, 1.0, 0.0;" : "+f"(f));
  if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f));
  if (KIND == K_LDS) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(sm) + (x & 0xffc))); x ^= v; }
  if (KIND == K_LDS128) {
    double a, b;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x * 16) & 0xff0)));
    sink += a + b;
  }
}

// PER = other instructions per DFMA, in eighths (0, 4 = one per two DFMAs, 8 = one each, 16 = two each)
template <int KIND, int PER8, int ILP> __global__ void probe(double *out, long long *cyc, int iters) {
  __shared__ char sm[4096];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) ((unsigned *)sm)[i] = i;
  __syncthreads();
  double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i * 1e-3;
  unsigned x[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; f[i] = 1.0f + i; }
  double sink = 0;
  unsigned y = blockIdx.x | 1;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i % ILP]) : "d"(b), "d"(a));
      if (PER8 == 4 && (i & 1)) other<KIND>(x[i & 7], y, f[i & 7], sm, sink);
      if (PER8 >= 8) other<KIND>(x[i & 7], y, f[i & 7], sm, sink);
      if (PER8 >= 16) other<KIND>(x[(i + 4) & 7], y, f[(i + 4) & 7], sm, sink);
    }
  }
  long long t1 = clock64();
  double r = sink;
#pragma unroll
  for (int i = 0; i < 16; i++) r += c[i];
#pragma unroll
  for (int i = 0; i < 8; i++) r += x[i] + f[i];
  if (r == 123.456) out[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND, int PER8, int ILP> double run(int warps_per_smsp, int iters) {
  double *out; long long *cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
  probe<KIND, PER8, ILP><<<148, 128 * warps_per_smsp>>>(out, cyc, iters);
  probe<KIND, PER8, ILP><<<148, 128 * warps_per_smsp>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; i++) s += h[i];
  cudaFree(out); cudaFree(cyc);
  return s / 148 / iters;
}

#define ROW(name, KIND, PER8)                                                                       \
  printf(" \"%s\": [%.1f, %.1f, %.1f],\n", name, run<KIND, PER8, 16>(1, it), run<KIND, PER8, 16>(2, it), \
         run<KIND, PER8, 16>(4, it));

int main() {
  const int it = 1 << 12;
  printf("{\"unit\": \"SM cycles per iteration of 16 DFMA (+ others) per warp, at 1 / 2 / 4 warps per sub-partition\",\n");
  ROW("dfma_only", K_NONE, 0)
  ROW("dfma+imad_1_per_2", K_IMAD, 4)
  ROW("dfma+imad_1_per_1", K_IMAD, 8)
  ROW("dfma+imad_2_per_1", K_IMAD, 16)
  ROW("dfma+lop_1_per_1", K_LOP, 8)
  ROW("dfma+lop_2_per_1", K_LOP, 16)
  ROW("dfma+fmul_mov_1_per_1", K_FMOV, 8)
  ROW("dfma+ffma_1_per_1", K_FFMA, 8)
  ROW("dfma+lds32_1_per_2", K_LDS, 4)
  ROW("dfma+lds32_1_per_1", K_LDS, 8)
  ROW("dfma+lds128_1_per_2", K_LDS128, 4)
  printf(" \"one_warp_ilp\": {\"1\": %.1f, \"2\": %.1f, \"4\": %.1f, \"8\": %.1f, \"16\": %.1f},\n",
         run<K_NONE, 0, 1>(1, it), run<K_NONE, 0, 2>(1, it), run<K_NONE, 0, 4>(1, it), run<K_NONE, 0, 8>(1, it),
         run<K_NONE, 0, 16>(1, it));
  printf(" \"two_warps_ilp\": {\"1\": %.1f, \"2\": %.1f, \"4\": %.1f},\n", run<K_NONE, 0, 1>(2, it), run<K_NONE, 0, 2>(2, it),
         run<K_NONE, 0, 4>(2, it));
  printf(" \"four_warps_ilp\": {\"1\": %.1f, \"2\": %.1f, \"4\": %.1f}}\n", run<K_NONE, 0, 1>(4, it), run<K_NONE, 0, 2>(4, it),
         run<K_NONE, 0, 4>(4, it));
  return 0;
}
