// TMEM as per-thread private scratch: correctness + cost of tcgen05.st/ld 32x32b.x32
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(512, 1) probe(uint32_t *out, long long *cyc, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s;
  // warp w owns lanes 32*(w%4).., columns 32*(w/4)..+31  (16 warps -> 4 column blocks of 32)
  const uint32_t taddr = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 32u;
  uint32_t r[32], q[32];
#pragma unroll
  for (int i = 0; i < 32; i++) r[i] = threadIdx.x * 1000u + i;
  tmem_st32(taddr, r);
  tmem_wait_st();
  __syncthreads();
  tmem_ld32(taddr, q);
  tmem_wait_ld();
  uint32_t bad = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) bad += (q[i] != threadIdx.x * 1000u + i);
  out[blockIdx.x * blockDim.x + threadIdx.x] = bad;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[gridDim.x * blockDim.x] = base;
  __syncthreads();
  // timing: all 16 warps ld (then st) their 32 columns `iters` times
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    tmem_ld32(taddr, q);
    tmem_wait_ld();
    acc += q[it & 31];
  }
  long long t1 = clock64();
  for (int it = 0; it < iters; it++) {
    r[0] = acc + it;
    tmem_st32(taddr, r);
    tmem_wait_st();
  }
  long long t2 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = acc; }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
  uint32_t *out; long long *cyc;
  const int blocks = 148, threads = 512, iters = 2000;
  cudaMalloc(&out, (blocks * threads + 1) * 4);
  cudaMalloc(&cyc, 3 * 8);
  probe<<<blocks, threads>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[148 * 512 + 1]; long long hc[3];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  long bad = 0; for (int i = 0; i < blocks * threads; i++) bad += h[i];
  printf("mismatched words: %ld  tmem_base=0x%x\n", bad, h[blocks * threads]);
  // per SM per iteration: 16 warps x 32 lanes x 128 B = 64 KB
  printf("ld: %.1f cyc/iter (16 warps x 4 KB = 64 KB) -> %.1f B/clk/SM\n", (double)hc[0] / iters, 65536.0 * iters / hc[0]);
  printf("st: %.1f cyc/iter -> %.1f B/clk/SM\n", (double)hc[1] / iters, 65536.0 * iters / hc[1]);
  return 0;
}
