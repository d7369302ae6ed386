// Throughput of legacy mma.sync INT8 (m16n8k32, u8 x u8 -> s32) on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) imma(int *out, int iters, uint32_t seed) {
  int acc[16][4];
#pragma unroll
  for (int t = 0; t < 16; t++) for (int c = 0; c < 4; c++) acc[t][c] = 0;
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < 16; t++) {
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(acc[t][0]), "+r"(acc[t][1]), "+r"(acc[t][2]), "+r"(acc[t][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + t), "r"(b1));
    }
    a0 += 1;
  }
  int s = 0;
#pragma unroll
  for (int t = 0; t < 16; t++) for (int c = 0; c < 4; c++) s += acc[t][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000, blocks = 148 * 4;
  imma<<<blocks, 256>>>(out, 10, 1); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); imma<<<blocks, 256>>>(out, iters, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * 8 /*warps*/ * iters * 16 * (16.0 * 8 * 32 * 2);
  printf("status %s  %.2f ms  %.1f TOPS (int8 mma.sync m16n8k32)\n", cudaGetErrorString(cudaGetLastError()), ms, ops / ms / 1e9);
  return 0;
}
