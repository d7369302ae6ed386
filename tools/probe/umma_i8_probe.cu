// Probe: tcgen05.mma kind::i8 with A from TMEM (TS form) and B from shared memory (K-major, no
// swizzle): D[128][256] (s32, TMEM) = A[128][32] (u8) * B[256][32]^T (u8).  Checks the operand
// layouts assumed by a tcgen05 key-switch kernel against a CPU reference.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(const uint32_t *A /*[128][8] words*/, const uint8_t *Btile /*8 KB canonical*/,
                                                int *D /*[128][256]*/, int a_variant) {
  __shared__ __align__(128) uint8_t sB[8192];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 8192 / 16; i += 128)
    reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(Btile)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s;
  // A row of this thread (lane = row) -> 8 columns starting at column 256
  {
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = A[threadIdx.x * 8 + i];
    const uint32_t taddr = base + (((uint32_t)(warp & 3) * 32u) << 16) + 256u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x == 0) {
    // smem descriptor for B: K-major, SWIZZLE_NONE: core matrix 8 rows x 16 B; LBO (next 16 B of K) = 128 B,
    // SBO (next 8 rows of N) = 256 B; version 1
    uint64_t bdesc = 0;
    bdesc |= (uint64_t)((smem_u32(sB) & 0x3FFFF) >> 4);
    bdesc |= (uint64_t)(128 >> 4) << 16;
    bdesc |= (uint64_t)(256 >> 4) << 32;
    bdesc |= (uint64_t)1 << 46;
    // instruction descriptor: c=S32 (2<<4), a/b unsigned 8-bit (0), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_tmem = base, a_tmem = base + 256u;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {  // wait for the MMA
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!ok);
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t taddr = base + (((uint32_t)(warp & 3) * 32u) << 16);
  for (int c0 = 0; c0 < 256; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) D[threadIdx.x * 256 + c0 + i] = (int)r[i];
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
  static uint8_t hA[128][32], hB[256][32], hBt[8192];
  static int hD[128 * 256], ref[128][256];
  srand(1);
  for (int m = 0; m < 128; m++) for (int k = 0; k < 32; k++) hA[m][k] = rand() & 0xFF;
  for (int n = 0; n < 256; n++) for (int k = 0; k < 32; k++) hB[n][k] = rand() & 0xFF;
  for (int m = 0; m < 128; m++) for (int n = 0; n < 256; n++) { int s = 0; for (int k = 0; k < 32; k++) s += hA[m][k] * hB[n][k]; ref[m][n] = s; }
  // canonical K-major no-swizzle tile: [n1 = n/8][k1 = k/16][r0 = n%8][16 B]
  for (int n = 0; n < 256; n++) for (int k = 0; k < 32; k++)
    hBt[(n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + (k % 16)] = hB[n][k];
  uint32_t *dA; uint8_t *dB; int *dD;
  cudaMalloc(&dA, 128 * 32); cudaMalloc(&dB, 8192); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, 128 * 32, cudaMemcpyHostToDevice);   // row m: 32 bytes = 8 little-endian words
  cudaMemcpy(dB, hBt, 8192, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, sizeof(hD));
  probe<<<1, 128>>>(dA, dB, dD, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int m = 0; m < 128; m++) for (int n = 0; n < 256; n++) bad += (hD[m * 256 + n] != ref[m][n]);
  printf("mismatches: %ld of %d\n", bad, 128 * 256);
  printf("D[0][0..3] = %d %d %d %d   ref = %d %d %d %d\n", hD[0], hD[1], hD[2], hD[3], ref[0][0], ref[0][1], ref[0][2], ref[0][3]);
  printf("D[1][0..3] = %d %d %d %d   ref = %d %d %d %d\n", hD[256], hD[257], hD[258], hD[259], ref[1][0], ref[1][1], ref[1][2], ref[1][3]);
  printf("D[37][100..101] = %d %d   ref = %d %d\n", hD[37 * 256 + 100], hD[37 * 256 + 101], ref[37][100], ref[37][101]);
  return 0;
}
