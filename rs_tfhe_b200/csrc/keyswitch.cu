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
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int KS_B = 8;

__global__ void __launch_bounds__(320) keyswitch_kernel(const KsArgs a) {
  extern __shared__ uint32_t s_src[];  // [KS_B][n_in] : a_i + PREC_OFFSET
  const uint32_t N = a.n_in;  // 1024 for the bootstrap key switch, n for proxy re-encryption
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

// ---- base-4 variant (gate parameter sets: basebit = 2) --------------------------------
// For each (i, j) the three candidate rows k = 1,2,3 are loaded ONCE per CTA and reused by
// all KS4_B ciphertexts of the CTA: every ciphertext adds the row its digit selects (a
// CTA-uniform branch, so no divergence).  Compared with the generic kernel this cuts L1/L2
// row traffic by KS4_B*(3/4)/3 = 4x and the per-element instruction count by ~30 %.
constexpr int KS4_B = 16;

template <int T>
__global__ void __launch_bounds__(192) keyswitch_base4_kernel(const KsArgs a) {
  extern __shared__ uint32_t s_src[];  // [KS4_B][1024] : a_i + PREC_OFFSET
  const uint32_t N = br::kN;
  const size_t ct0 = (size_t)blockIdx.x * KS4_B;
  const uint32_t prec = 1u << (32 - (1 + 2 * T));
  for (int b = 0; b < KS4_B; b++) {
    const size_t ct = ct0 + b;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
      s_src[b * N + i] = ct < a.count ? a.ext[ct * (N + 1) + i] + prec : prec;  // digits of prec are 0
  }
  __syncthreads();

  const uint32_t stride4 = a.stride >> 2;
  const uint32_t x4 = threadIdx.x;
  if (x4 >= stride4) return;
  uint4 acc[KS4_B];
#pragma unroll
  for (int b = 0; b < KS4_B; b++) acc[b] = make_uint4(0, 0, 0, 0);
  const uint4 *ksk4 = reinterpret_cast<const uint4 *>(a.ksk) + x4;
  for (uint32_t i = 0; i < N; i++) {
    uint32_t ab[KS4_B];
#pragma unroll
    for (int b = 0; b < KS4_B; b++) ab[b] = s_src[b * N + i];
    const uint4 *rowp = ksk4 + (size_t)(i * T * 4) * stride4;
#pragma unroll
    for (int j = 0; j < T; j++) {
      const uint4 r1 = __ldg(rowp + (size_t)(4 * j + 1) * stride4);
      const uint4 r2 = __ldg(rowp + (size_t)(4 * j + 2) * stride4);
      const uint4 r3 = __ldg(rowp + (size_t)(4 * j + 3) * stride4);
#pragma unroll
      for (int b = 0; b < KS4_B; b++) {
        const uint32_t k = (ab[b] >> (30 - 2 * j)) & 3u;  // CTA-uniform
        if (k == 1) { acc[b].x += r1.x; acc[b].y += r1.y; acc[b].z += r1.z; acc[b].w += r1.w; }
        else if (k == 2) { acc[b].x += r2.x; acc[b].y += r2.y; acc[b].z += r2.z; acc[b].w += r2.w; }
        else if (k == 3) { acc[b].x += r3.x; acc[b].y += r3.y; acc[b].z += r3.z; acc[b].w += r3.w; }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < KS4_B; b++) {
    const size_t ct = ct0 + b;
    if (ct >= a.count) break;
    uint32_t *o = a.out + ct * (a.n + 1);
    const uint32_t vals[4] = {acc[b].x, acc[b].y, acc[b].z, acc[b].w};
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t x = 4 * x4 + c;
      if (x <= a.n) {
        uint32_t init = (x == a.n) ? a.ext[ct * (N + 1) + N] : 0u;
        o[x] = init - vals[c];
      }
    }
  }
}

// ---- latency variant: a handful of ciphertexts (dependent PBS chains) -----------------------------------
// The tensor-core key switch streams the whole 103 MB key whatever the batch size (0.2 ms); one ciphertext
// needs 6912 of its rows (19 MB).  Here the N*t (i, j) terms of one ciphertext are split over `parts` CTAs:
// each CTA adds up the rows its terms select (128-bit loads, eight rows in flight per thread) and subtracts its
// partial sum from the output with atomics -- wrapping u32 arithmetic, so any order is bit-exact.  The output
// must be zero on entry (the launcher clears it); part 0 adds res.b = src.b (trgsw.rs:343).
__global__ void __launch_bounds__(320) keyswitch_small_kernel(const KsArgs a, uint32_t parts) {
  extern __shared__ uint32_t s_row[];                 // selected row per term of this CTA (zero_row: digit 0)
  const uint32_t N = a.n_in, t = a.iks_t;
  const size_t ct = blockIdx.y;
  const uint32_t terms = N * t;
  const uint32_t t0 = (uint32_t)((uint64_t)terms * blockIdx.x / parts), t1 = (uint32_t)((uint64_t)terms * (blockIdx.x + 1) / parts);
  const uint32_t prec = 1u << (32 - (1 + a.basebit * t));
  const uint32_t mask = (1u << a.basebit) - 1u;
  const uint32_t *src = a.ext + ct * (N + 1);
  for (uint32_t x = t0 + threadIdx.x; x < t1; x += blockDim.x) {
    const uint32_t i = x / t, j = x % t;
    const uint32_t k = ((src[i] + prec) >> (32 - (j + 1) * a.basebit)) & mask;
    s_row[x - t0] = k ? ((x << a.basebit) + k) : a.zero_row;
  }
  __syncthreads();
  const uint32_t stride4 = a.stride >> 2;
  const uint32_t x4 = threadIdx.x;
  if (x4 >= stride4) return;
  const uint4 *ksk4 = reinterpret_cast<const uint4 *>(a.ksk) + x4;
  uint4 acc = make_uint4(0, 0, 0, 0);
  const uint32_t cnt = t1 - t0;
  uint32_t x = 0;
  for (; x + 8 <= cnt; x += 8) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = __ldg(ksk4 + (size_t)s_row[x + u] * stride4);
#pragma unroll
    for (int u = 0; u < 8; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  for (; x < cnt; x++) {
    const uint4 v = __ldg(ksk4 + (size_t)s_row[x] * stride4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  uint32_t *o = a.out + ct * (a.n + 1);
  const uint32_t vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const uint32_t w = 4 * x4 + c;
    if (w <= a.n) {
      uint32_t d = 0u - vals[c];
      if (blockIdx.x == 0 && w == a.n) d += src[N];
      if (d) atomicAdd(o + w, d);
    }
  }
}

template <int T> cudaError_t launch_base4(const KsArgs &args, cudaStream_t stream) {
  const int smem = KS4_B * br::kN * 4;
  {  // per device and cheap: set on every launch (engines may live on several GPUs)
    cudaError_t e = cudaFuncSetAttribute(keyswitch_base4_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  const uint32_t stride4 = args.stride >> 2;
  const int threads = (int)((stride4 + 31) & ~31u);
  const unsigned grid = (unsigned)((args.count + KS4_B - 1) / KS4_B);
  keyswitch_base4_kernel<T><<<grid, threads, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace

cudaError_t ks_small_launch(const KsArgs &args, int num_sms, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  const uint32_t stride4 = args.stride >> 2;
  const int threads = (int)((stride4 + 31) & ~31u);       // one thread per 4 output words, as in the row-walk kernels
  if (threads > 320) return cudaErrorInvalidValue;
  uint32_t parts = (uint32_t)((size_t)num_sms * 2 / args.count);
  const uint32_t terms = args.n_in * args.iks_t;
  if (parts < 1) parts = 1;
  if (parts > terms / 8) parts = terms / 8;
  const uint32_t per = (terms + parts - 1) / parts + 1;
  cudaError_t e = cudaMemsetAsync(args.out, 0, args.count * (size_t)(args.n + 1) * 4, stream);
  if (e != cudaSuccess) return e;
  keyswitch_small_kernel<<<dim3(parts, (unsigned)args.count), threads, per * 4, stream>>>(args, parts);
  return cudaGetLastError();
}

cudaError_t ks_launch(const KsArgs &args, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  const uint32_t stride4 = args.stride >> 2;
  const int threads = (int)((stride4 + 31) & ~31u);
  if (threads > 320) return cudaErrorInvalidValue;
  static const bool generic_only = getenv("TFHE_KS_GENERIC") != nullptr;
  if (args.basebit == 2 && threads <= 192 && !generic_only && args.n_in == br::kN) {
    if (args.iks_t == 7) return launch_base4<7>(args, stream);
    if (args.iks_t == 8) return launch_base4<8>(args, stream);
    if (args.iks_t == 9) return launch_base4<9>(args, stream);
  }
  const int smem = KS_B * (int)args.n_in * 4;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((args.count + KS_B - 1) / KS_B);
  keyswitch_kernel<<<grid, threads, smem, stream>>>(args);
  return cudaGetLastError();
}
