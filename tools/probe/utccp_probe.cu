// Probe: tcgen05.cp 64x128b.warpx2::02_13 as a shared-memory -> TMEM broadcast copy.
// Checks (1) the source layout (64 rows x 16 B, no swizzle, 8-row core matrices 128 B apart),
// (2) which TMEM lanes receive which rows (rows 0-31 -> lane quadrants 0 and 2, rows 32-63 ->
// quadrants 1 and 3), (3) the cost of copying one 16 KB key row (16 copies) and of reading it
// back with tcgen05.ld.32x32b.x16 from 8 warps.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

__global__ void __launch_bounds__(256, 1) probe(uint32_t *out /*[8 warps][32 lanes][64 cols]*/, long long *cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t *row = reinterpret_cast<uint32_t *>(smem);            // 16 KB: [16 slices][64 rows][4 words]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 4096; i += 256) row[i] = 0xA0000000u + i;  // word index as payload
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s;
  uint32_t parity = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; s++) {
      uint64_t desc = make_desc(smem_u32(row) + s * 1024, 128, 128);
      asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(base + 64 + 4 * s), "l"(desc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, parity); parity ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = base + (((uint32_t)(warp & 3) * 32u) << 16) + 64;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; i++) out[(warp * 32 + lane) * 64 + c0 + i] = r[i];
  }
  __syncthreads();

  // ---- timing 1: copy only (16 copies + commit + wait per iteration) ----
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int s = 0; s < 16; s++) {
        uint64_t desc = make_desc(smem_u32(row) + s * 1024, 128, 128);
        asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(base + 64 + 64 * (it & 3) + 4 * s), "l"(desc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, parity); parity ^= 1;
  }
  long long t1 = clock64();
  __syncthreads();
  // ---- timing 2: 8 warps each read the 64-column row (4 x ld.x16) ----
  uint32_t sink = 0;
  long long t2 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; i++) sink ^= r[i];
    }
  }
  long long t3 = clock64();
  __syncthreads();
  // ---- timing 3: copy (thread 0) concurrent with LDS.128 traffic from the other 7 warps ----
  long long t4 = clock64();
  if (warp == 0) {
    for (int it = 0; it < iters; it++) {
      if (lane == 0) {
        for (int s = 0; s < 16; s++) {
          uint64_t desc = make_desc(smem_u32(row) + s * 1024, 128, 128);
          asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(base + 64 + 64 * (it & 3) + 4 * s), "l"(desc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, parity);
      }
      parity ^= 1;
      __syncwarp();
    }
  } else {
    const uint4 *p = reinterpret_cast<const uint4 *>(smem);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {   // 16 LDS.128 per warp = 64 wavefronts; 7 warps -> 448 per iteration
        uint4 v = p[(k * 32 + lane + it) & 1023];
        sink ^= v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  }
  long long t5 = clock64();
  __syncthreads();
  // ---- timing 4: the same LDS traffic without the copy ----
  long long t6 = clock64();
  if (warp != 0) {
    const uint4 *p = reinterpret_cast<const uint4 *>(smem);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        uint4 v = p[(k * 32 + lane + it) & 1023];
        sink ^= v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  }
  long long t7 = clock64();
  if (threadIdx.x == 32) cyc[4] = t7 - t6;
  if (sink == 0x12345678u) out[0] = sink;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; }
  if (threadIdx.x == 32) cyc[3] = t5 - t4;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
  uint32_t *d_out; long long *d_cyc;
  static uint32_t h[8 * 32 * 64];
  long long hc[5];
  cudaMalloc(&d_out, sizeof(h)); cudaMalloc(&d_cyc, sizeof(hc));
  cudaMemset(d_out, 0xFF, sizeof(h));
  const int iters = 2000;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe<<<1, 256, 16384>>>(d_out, d_cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, d_cyc, sizeof(hc), cudaMemcpyDeviceToHost);
  // expectation: warp w lane l column c = word of slice s=c/4, row (w&1)*32+l, word c%4
  long bad = 0;
  for (int w = 0; w < 8; w++) for (int l = 0; l < 32; l++) for (int c = 0; c < 64; c++) {
    uint32_t exp = 0xA0000000u + (c / 4) * 256 + (((w & 1) * 32 + l) * 4) + (c % 4);
    bad += (h[(w * 32 + l) * 64 + c] != exp);
  }
  printf("layout mismatches: %ld of %d\n", bad, 8 * 32 * 64);
  for (int w = 0; w < 4; w++)
    printf("warp %d lane 0: %08x %08x %08x %08x | %08x ;  lane 5: %08x %08x\n", w, h[(w * 32) * 64], h[(w * 32) * 64 + 1],
           h[(w * 32) * 64 + 2], h[(w * 32) * 64 + 3], h[(w * 32) * 64 + 4], h[(w * 32 + 5) * 64], h[(w * 32 + 5) * 64 + 1]);
  printf("copy 16 KB row: %.1f cyc/row -> %.1f B/clk\n", (double)hc[0] / iters, 16384.0 * iters / hc[0]);
  printf("8 warps read 64 cols: %.1f cyc/iter -> %.1f B/clk/SM\n", (double)hc[1] / iters, 8 * 32 * 256.0 * iters / hc[1]);
  printf("concurrent: copy warp %.1f cyc/row ; LDS warps %.1f cyc/iter (448 wavefronts) ; LDS alone %.1f cyc/iter\n", (double)hc[2] / iters, (double)hc[3] / iters, (double)hc[4] / iters);
  return 0;
}
