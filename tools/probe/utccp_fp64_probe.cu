// Probe: how much does a stream of tcgen05.cp (shared -> TMEM) slow concurrent FP64 work?
// Warps 1..7 run independent DFMA chains; thread 0 issues copies of one shape back to back
// (a commit + wait every 16).  Reports DFMA-loop cycles with and without the copy stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
template <int SHAPE> __device__ __forceinline__ void cp(uint32_t taddr, uint64_t desc) {
  if (SHAPE == 0) asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
  if (SHAPE == 1) asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
  if (SHAPE == 2) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
  if (SHAPE == 3) asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
template <int SHAPE>
__global__ void __launch_bounds__(256, 1) probe(long long *cyc, double *sink, int iters, int with_copy) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 8192; i += 256) reinterpret_cast<uint32_t *>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s;
  if (warp == 0) {
    if (threadIdx.x == 0 && with_copy) {
      uint32_t parity = 0; long long n = 0;
      long long t0 = clock64();
      while (!stop) {
        for (int s = 0; s < 16; s++) cp<SHAPE>(base + 64 + 8 * s, make_desc(smem_u32(smem) + (s & 7) * 1024, 128, SHAPE == 2 ? 256 : 128));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, parity); parity ^= 1; n += 16;
      }
      cyc[2] = clock64() - t0; cyc[3] = n;
    }
  } else {
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
      }
    }
    long long t1 = clock64();
    sink[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 32) cyc[0] = t1 - t0;
    asm volatile("bar.sync 1, 224;");
    if (threadIdx.x == 32) stop = 1;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}
template <int SHAPE> void run(const char *name, int bytes) {
  long long *d_cyc, h[4]; double *d_sink;
  cudaMalloc(&d_cyc, 32); cudaMalloc(&d_sink, 256 * 8);
  const int iters = 20000;
  cudaFuncSetAttribute(probe<SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  double base_cyc = 0;
  for (int wc = 0; wc < 2; wc++) {
    cudaMemset(d_cyc, 0, 32);
    probe<SHAPE><<<1, 256, 32768>>>(d_cyc, d_sink, iters, wc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost);
    if (!wc) { base_cyc = (double)h[0]; printf("%-22s fp64 alone: %.0f cyc (7 warps x %d DFMA -> %.2f DFMA/clk/SM) [%s]\n", name, base_cyc, iters * 64, 7.0 * 32 * iters * 64 / h[0], cudaGetErrorString(e)); }
    else printf("%-22s fp64 with copies: %.0f cyc (x%.3f); %lld copies in %lld cyc = %.1f cyc/copy, %.1f B/clk [%s]\n", name, (double)h[0], h[0] / base_cyc, h[3], h[2], (double)h[2] / h[3], (double)bytes * h[3] / h[2], cudaGetErrorString(e));
  }
}
int main() {
  run<0>("64x128b.warpx2::02_13", 1024);
  run<1>("128x128b", 2048);
  run<2>("128x256b", 4096);
  run<3>("32x128b.warpx4", 512);
  return 0;
}
