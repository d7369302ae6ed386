// keyswitch_mma.cu -- K4 on the tensor pipe for the gate parameter sets (basebit = 2).
//
// identity_key_switching (reference src/trgsw.rs:332-360) is
//     out = (0,..,0,b) - sum_{i<N} sum_{j<t} KSK[i][j][digit_j(a_i + PREC_OFFSET)]
// i.e. a product of a one-hot selection matrix with the key:  S[ct][(i,j,k)] = (digit == k, k != 0),
// out[ct][x] = init - sum_q S[ct][q] * KSK[q][x]  (mod 2^32).  Splitting every KSK word into its
// four bytes makes this an exact u8 x u8 -> s32 GEMM (each accumulator sums <= N*t bytes < 2^24);
// the byte planes are recombined with wrapping shifts in the epilogue, so the result is
// bit-identical to the reference's wrapping subtractions.
//
//   M = ciphertexts (64 per CTA = 4 m16 tiles; two CTAs per SM), N = 4 byte planes x 8 words per warp,
//   K = (i-block, j, ii, k): 32 per mma.sync.m16n8k32 = 8 coefficients x 4 digit values.
// A fragments are built in registers from the digits (one 32-bit register = the one-hot over
// k = 0..3 of one (ct, i, j)); B fragments come from a key copy pre-tiled at upload into exactly
// the fragment order and are streamed global -> shared with cp.async (8 k-steps deep, private
// slots per thread, no block barrier); accumulators are s32 in registers.
#include "kernels.h"

namespace {

#ifndef KM_MT_DEF
#define KM_MT_DEF 4
#endif
#ifndef KM_CTAS_DEF
#define KM_CTAS_DEF 2
#endif
constexpr int KM_MT = KM_MT_DEF;       // m16 tiles per CTA
constexpr int KM_WARPS = 8;    // each warp owns 8 words (32 byte columns) of the output
constexpr int KM_DEPTH = 8;    // cp.async ring depth in k-steps

__device__ __forceinline__ void mma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                       uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int Nw> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(Nw) : "memory");
}
// one-hot over k = 0..3 in the four bytes of a register.  Digit 0 sets byte 0, which meets the
// k = 0 byte of the B fragment -- forced to zero at upload -- so it selects nothing (trgsw.rs:351).
// sh3 = (30 - 2j) - 3: the digit lands pre-multiplied by 8.
__device__ __forceinline__ uint32_t onehot(uint32_t abar, uint32_t sh3) {
  return 1u << ((abar >> sh3) & 0x18u);
}

__global__ void __launch_bounds__(KM_WARPS * 32, KM_CTAS_DEF) ks_mma_kernel(const KsMmaArgs a) {
  extern __shared__ __align__(16) uint8_t km_smem[];
  uint4(*ring)[2][KM_WARPS * 32] =
      reinterpret_cast<uint4(*)[2][KM_WARPS * 32]>(km_smem);  // [slot][b0|b1][thread] : 64 KB
  const uint32_t N = br::kN;
  const uint32_t ncta_n = a.nxg / KM_WARPS;
  const uint32_t mtile = blockIdx.x / ncta_n, ncta = blockIdx.x % ncta_n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g8 = lane >> 2, tig = lane & 3;
  const uint32_t t = a.iks_t;
  const uint32_t prec = 1u << (32 - (1 + 2 * t));
  const size_t ct_base = (size_t)mtile * (KM_MT * 16);
  const uint32_t xg = ncta * KM_WARPS + warp;

  int acc[KM_MT][4][4];
#pragma unroll
  for (int m = 0; m < KM_MT; m++)
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[m][p][c] = 0;

  // B stream: k-step s = iblk * t + j; this thread's two 16-byte pieces per k-step
  const size_t row4 = (size_t)a.nxg * 8;  // uint4 per (iblk, j, ii)
  const uint4 *wbase = reinterpret_cast<const uint4 *>(a.w) + (size_t)xg * 8 + g8;
  const uint32_t ksteps = (N / 8) * t;
  auto issue = [&](uint32_t s) {
    if (s < ksteps) {
      const uint4 *p0 = wbase + ((size_t)s * 8 + tig) * row4;
      cp_async16(&ring[s % KM_DEPTH][0][threadIdx.x], p0);
      cp_async16(&ring[s % KM_DEPTH][1][threadIdx.x], p0 + 4 * row4);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < KM_DEPTH - 1; s++) issue(s);

  // digits: abar[m][r*2+h] for ct = ct_base + 16m + g8 + 8r, i = 8*iblk + tig + 4h
  auto load_ab = [&](uint32_t iblk, uint32_t (&ab)[KM_MT][4]) {
#pragma unroll
    for (int m = 0; m < KM_MT; m++)
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const size_t ct = ct_base + 16 * m + g8 + 8 * r;
        const uint32_t *e = a.ext + ct * (N + 1) + 8 * iblk + tig;
        const bool ok = ct < a.count && iblk < N / 8;
        ab[m][r] = (ok ? __ldg(e) : 0u) + prec;          // a1: row g8+8 -> index 1
        ab[m][2 + r] = (ok ? __ldg(e + 4) : 0u) + prec;  // a2/a3: second half of the k-step
      }
  };
  uint32_t ab[KM_MT][4], abn[KM_MT][4];
  load_ab(0, ab);

  uint32_t s = 0;
  for (uint32_t iblk = 0; iblk < N / 8; iblk++) {
    load_ab(iblk + 1, abn);  // prefetch next block's digits (masked past the end)
    for (uint32_t j = 0; j < t; j++, s++) {
      issue(s + KM_DEPTH - 1);
      cp_async_wait<KM_DEPTH - 1>();
      const uint4 B0 = ring[s % KM_DEPTH][0][threadIdx.x];
      const uint4 B1 = ring[s % KM_DEPTH][1][threadIdx.x];
      const uint32_t sh = 27 - 2 * j;
#pragma unroll
      for (int m = 0; m < KM_MT; m++) {
        const uint32_t a0 = onehot(ab[m][0], sh), a1 = onehot(ab[m][1], sh);
        const uint32_t a2 = onehot(ab[m][2], sh), a3 = onehot(ab[m][3], sh);
        mma_u8(acc[m][0], a0, a1, a2, a3, B0.x, B1.x);
        mma_u8(acc[m][1], a0, a1, a2, a3, B0.y, B1.y);
        mma_u8(acc[m][2], a0, a1, a2, a3, B0.z, B1.z);
        mma_u8(acc[m][3], a0, a1, a2, a3, B0.w, B1.w);
      }
    }
#pragma unroll
    for (int m = 0; m < KM_MT; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) ab[m][q] = abn[m][q];
  }
  cp_async_wait<0>();

  // epilogue: recombine byte planes (wrapping), out = init - sum   (trgsw.rs:343,353-355)
#pragma unroll
  for (int m = 0; m < KM_MT; m++)
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const size_t ct = ct_base + 16 * m + g8 + 8 * r;
      if (ct >= a.count) continue;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t x = xg * 8 + tig * 2 + c;
        if (x > a.n) continue;
        const int ci = 2 * r + c;
        const uint32_t sum = (uint32_t)acc[m][0][ci] + ((uint32_t)acc[m][1][ci] << 8) +
                             ((uint32_t)acc[m][2][ci] << 16) + ((uint32_t)acc[m][3][ci] << 24);
        const uint32_t init = (x == a.n) ? a.ext[ct * (N + 1) + N] : 0u;
        a.out[ct * (a.n + 1) + x] = init - sum;
      }
    }
}

// blob KSK rows (u32[rows + 1][stride], reference row order key.rs:102-122) -> B-fragment order.
// dst index = ((((iblk*t + j)*8 + ii)*nxg + xg)*8 + g8)*4 + p ; value byte k = byte p of
// KSK[((8*iblk+ii)*t + j)*4 + k][8*xg + g8]  (k = 0 forced to 0: never selected)
__global__ void ksk_mma_relayout_kernel(const uint32_t *__restrict__ src, uint32_t stride,
                                        uint32_t *__restrict__ dst, uint32_t n, uint32_t t, uint32_t nxg,
                                        size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const uint32_t p = idx & 3, g8 = (idx >> 2) & 7;
  size_t rest = idx >> 5;
  const uint32_t xg = rest % nxg; rest /= nxg;
  const uint32_t ii = rest & 7; rest >>= 3;
  const uint32_t j = rest % t;
  const uint32_t iblk = (uint32_t)(rest / t);
  const uint32_t x = xg * 8 + g8, i = iblk * 8 + ii;
  uint32_t v = 0;
  if (x <= n) {
    const size_t row0 = ((size_t)i * t + j) * 4;
#pragma unroll
    for (uint32_t k = 1; k < 4; k++) {
      const uint32_t wv = src[(row0 + k) * stride + x];
      v |= ((wv >> (8 * p)) & 0xFFu) << (8 * k);
    }
  }
  dst[idx] = v;
}

}  // namespace

cudaError_t ks_mma_launch(const KsMmaArgs &args, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  if (args.nxg % KM_WARPS != 0) return cudaErrorInvalidValue;
  const size_t mtiles = (args.count + KM_MT * 16 - 1) / (KM_MT * 16);
  const unsigned grid = (unsigned)(mtiles * (args.nxg / KM_WARPS));
  const int smem = KM_DEPTH * 2 * KM_WARPS * 32 * 16;
  {  // per device and cheap: set on every launch (engines may live on several GPUs)
    cudaError_t e = cudaFuncSetAttribute(ks_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  ks_mma_kernel<<<grid, KM_WARPS * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t ksk_mma_relayout_launch(const uint32_t *blob_rows, uint32_t stride, uint32_t *dst, uint32_t n,
                                    uint32_t t, cudaStream_t stream) {
  const size_t total = ks_mma_words(n, t);
  ksk_mma_relayout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(blob_rows, stride, dst, n, t,
                                                                              ks_mma_nxg(n), total);
  return cudaGetLastError();
}
