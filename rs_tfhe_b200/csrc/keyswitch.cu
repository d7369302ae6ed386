// keyswitch.cu -- K4: identity key switch of extracted level-1 samples.
//
// Replaces trgsw::identity_key_switching (reference src/trgsw.rs:332-360):
//   out = (0,...,0,b) - sum_{i<N} sum_{j<t} KSK[i][j][digit_j(a_i + PREC_OFFSET)]
// Pure wrapping u32 arithmetic, so any summation order is bit-exact.
//
// A CTA owns KS_B ciphertexts and the whole output width: thread x keeps words
// 4x..4x+3 of all KS_B outputs in registers and walks (i, j) in the same order
// as every other CTA, so at any moment the whole grid touches the same few
// hundred KB of the key (L2-resident; HBM sees the key once per wave).  Rows are
// padded to a multiple of 4 words at upload so each row read is one coalesced
// 128-bit load per thread; digit 0 reads a private all-zero row instead of the
// caller's k=0 rows (which the reference never reads), keeping loads
// unconditional and batched.
#include "kernels.h"

namespace {

constexpr int KS_B = 8;

__global__ void __launch_bounds__(320) keyswitch_kernel(const KsArgs a) {
  extern __shared__ uint32_t s_src[];  // [KS_B][1024] : a_i + PREC_OFFSET
  const uint32_t N = br::kN;
  const size_t ct0 = (size_t)blockIdx.x * KS_B;
  const uint32_t prec = 1u << (32 - (1 + a.basebit * a.iks_t));
  for (int b = 0; b < KS_B; b++) {
    const size_t ct = ct0 + b;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
      s_src[b * N + i] = ct < a.count ? a.ext[ct * (N + 1) + i] + prec : 0u;
  }
  __syncthreads();

  const uint32_t stride4 = a.stride >> 2;
  const uint32_t x4 = threadIdx.x;
  if (x4 < stride4) {
    uint4 acc[KS_B];
#pragma unroll
    for (int b = 0; b < KS_B; b++) acc[b] = make_uint4(0, 0, 0, 0);
    const uint4 *ksk4 = reinterpret_cast<const uint4 *>(a.ksk) + x4;
    const uint32_t mask = (1u << a.basebit) - 1u;
    const uint32_t t = a.iks_t;
    const bool full = ct0 + KS_B <= a.count;
    for (uint32_t i = 0; i < N; i++) {
      uint32_t ab[KS_B];
#pragma unroll
      for (int b = 0; b < KS_B; b++) ab[b] = s_src[b * N + i];
      for (uint32_t j = 0; j < t; j++) {
        const uint32_t sh = 32 - (j + 1) * a.basebit;
        const uint32_t row0 = (i * t + j) << a.basebit;
        uint4 v[KS_B];
#pragma unroll
        for (int b = 0; b < KS_B; b++) {
          uint32_t k = (ab[b] >> sh) & mask;
          uint32_t row = (k != 0 && (full || ct0 + b < a.count)) ? row0 + k : a.zero_row;
          v[b] = __ldg(ksk4 + (size_t)row * stride4);
        }
#pragma unroll
        for (int b = 0; b < KS_B; b++) {
          acc[b].x += v[b].x; acc[b].y += v[b].y; acc[b].z += v[b].z; acc[b].w += v[b].w;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < KS_B; b++) {
      const size_t ct = ct0 + b;
      if (ct >= a.count) break;
      uint32_t *o = a.out + ct * (a.n + 1);
      const uint32_t vals[4] = {acc[b].x, acc[b].y, acc[b].z, acc[b].w};
#pragma unroll
      for (int c = 0; c < 4; c++) {
        uint32_t x = 4 * x4 + c;
        if (x <= a.n) {
          uint32_t init = (x == a.n) ? a.ext[ct * (N + 1) + N] : 0u;  // res.b = src.b (trgsw.rs:343)
          o[x] = init - vals[c];
        }
      }
    }
  }
}

}  // namespace

cudaError_t ks_launch(const KsArgs &args, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  const uint32_t stride4 = args.stride >> 2;
  const int threads = (int)((stride4 + 31) & ~31u);
  if (threads > 320) return cudaErrorInvalidValue;
  const int smem = KS_B * br::kN * 4;
  const unsigned grid = (unsigned)((args.count + KS_B - 1) / KS_B);
  keyswitch_kernel<<<grid, threads, smem, stream>>>(args);
  return cudaGetLastError();
}
