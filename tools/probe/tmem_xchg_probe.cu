// Probe: TMEM as an intra-warp exchange medium.  Store with tcgen05.st.32x32b.x32 (thread t -> lane t,
// register c -> column c), load back with tcgen05.ld.16x256b.x4 at lane bases 0 and 16.  Predicted:
//   register 4*rep + 2*h + b of the load at lane base Lb holds (lane Lb + t/4 + 8h, column 8*rep + 2*(t%4) + b).
// Then: cost of such round trips alone, of DFMA work alone, and of both interleaved (8 warps).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define R16(a) "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void ld16x256_x4(uint32_t taddr, uint32_t (&a)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : R16(a) : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(256, 1) probe(uint32_t *out, long long *cyc, double *sink, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tq = tmem_base_s + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 64u;
  uint32_t r[32], a[16], b[16];
  for (int c = 0; c < 32; c++) r[c] = (uint32_t)(lane * 64 + c);
  st32(tq, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  ld16x256_x4(tq, a);
  ld16x256_x4(tq + (16u << 16), b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; i++) { out[(warp * 32 + lane) * 32 + i] = a[i]; out[(warp * 32 + lane) * 32 + 16 + i] = b[i]; }
  __syncthreads();

  double f[8];
  for (int i = 0; i < 8; i++) f[i] = lane + i;
  const double m = 1.0000001, c0 = 1e-9;
  // phase 0: round trips only; 1: DFMA only (256 per iteration); 2: both
  for (int phase = 0; phase < 3; phase++) {
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (phase != 1) {
        st32(tq, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        ld16x256_x4(tq, a);
        ld16x256_x4(tq + (16u << 16), b);
      }
      if (phase != 0) {
#pragma unroll
        for (int u = 0; u < 32; u++)
#pragma unroll
          for (int i = 0; i < 8; i++) f[i] = fma(f[i], m, c0);
      }
      if (phase != 1) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[5]), "+r"(b[3]), "+r"(b[15]) :: "memory");
#pragma unroll
        for (int i = 0; i < 16; i++) { r[i] ^= a[i]; r[16 + i] ^= b[i]; }
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[phase] = t1 - t0;
  }
  // phase 3/4: warps 0-3 DFMA only, warps 4-7 idle (3) or continuous TMEM round trips (4)
  for (int phase = 3; phase < 5; phase++) {
    __syncthreads();
    long long t0 = clock64();
    if (warp < 4) {
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 32; u++)
#pragma unroll
          for (int i = 0; i < 8; i++) f[i] = fma(f[i], m, c0);
      }
    } else if (phase == 4) {
      for (int it = 0; it < 4 * iters; it++) {
        st32(tq, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        ld16x256_x4(tq, a);
        ld16x256_x4(tq + (16u << 16), b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a[0]), "+r"(a[5]), "+r"(b[3]), "+r"(b[15]) :: "memory");
#pragma unroll
        for (int i = 0; i < 16; i++) { r[i] ^= a[i]; r[16 + i] ^= b[i]; }
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[phase] = t1 - t0;
    if (threadIdx.x == 128) cyc[phase + 2] = t1 - t0;
  }
  double s = 0; for (int i = 0; i < 8; i++) s += f[i];
  uint32_t x = 0; for (int i = 0; i < 32; i++) x ^= r[i];
  sink[threadIdx.x] = s + (double)x;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}
int main() {
  uint32_t *d_out; long long *d_cyc, hc[8]; double *d_sink;
  static uint32_t h[8 * 32 * 32];
  cudaMalloc(&d_out, sizeof(h)); cudaMalloc(&d_cyc, 64); cudaMalloc(&d_sink, 256 * 8);
  const int iters = 4000;
  probe<<<1, 256>>>(d_out, d_cyc, d_sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, d_cyc, 64, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int w = 0; w < 8; w++) for (int t = 0; t < 32; t++) for (int half = 0; half < 2; half++) for (int i = 0; i < 16; i++) {
    int rep = i >> 2, hh = (i >> 1) & 1, bb = i & 1;
    int lanesrc = half * 16 + t / 4 + 8 * hh, col = 8 * rep + 2 * (t % 4) + bb;
    bad += (h[(w * 32 + t) * 32 + half * 16 + i] != (uint32_t)(lanesrc * 64 + col));
  }
  printf("mapping mismatches: %ld of %d\n", bad, 8 * 32 * 32);
  printf("thread 5 regs (lane,col): ");
  for (int i = 0; i < 8; i++) printf("(%u,%u) ", h[5 * 32 + i] / 64, h[5 * 32 + i] % 64);
  printf("\n");
  printf("round trips only : %.1f cyc/iter (8 warps; 32 KB st + 32 KB ld per iter -> %.0f B/clk each way)\n", (double)hc[0] / iters, 32768.0 * iters / hc[0]);
  printf("DFMA only        : %.1f cyc/iter (8 warps x 256 DFMA -> %.1f DFMA/clk/SM)\n", (double)hc[1] / iters, 8.0 * 32 * 256 * iters / hc[1]);
  printf("both interleaved : %.1f cyc/iter\n", (double)hc[2] / iters);
  printf("4 DFMA warps alone: %.1f cyc/iter ; with 4 TMEM warps streaming: %.1f cyc/iter (TMEM warps: %.1f cyc per round trip)\n",
         (double)hc[3] / iters, (double)hc[4] / iters, (double)hc[6] / (4.0 * iters));
  return 0;
}
