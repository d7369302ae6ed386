// keygen.cu -- K6: cloud-key generation on the device.
//
// Replaces (reference, file:line): key::gen_bootstrapping_key (src/key.rs:128-156) =
// per level-0 key bit a TRGSW encryption (src/trgsw.rs:29-49: 2l TRLWE encryptions of zero,
// src/trlwe.rs:30-52, plus mu*Bg^-(r+1) on a[0] / b[0]) followed by the Fourier conversion
// (src/trgsw.rs:58-68, src/trlwe.rs:91-96); and key::gen_key_switching_key (src/key.rs:102-122).
// The reference draws from an unseeded thread_rng, so there is nothing to match bit for bit:
// the generator here is the ChaCha20 block function used counter-based under a 256-bit key (OS
// entropy unless the caller asks for a reproducible test key), and the tests
// check the *structure* (b - a*s1 = noise + message with the right noise level) exactly.
//
// One CTA of 64 threads produces one TRGSW row: it draws a, forms a*s1 with the same
// three-pass negacyclic transform as the blind rotation (forward(a) x spectrum(s1) -> inverse),
// adds noise and the gadget term, transforms b, and writes both spectra straight into the
// device BSK layout (value = reference value / 1024), so no re-layout pass follows.
#include "kernels.h"

using namespace br;

namespace {

// ---- ChaCha20 block function (RFC 8439) as a counter-based generator: one call = 4 x u32 --------
// The a-words of every key row are published verbatim, so the generator must be a CSPRNG: with a
// non-cryptographic generator, recovering its state from the masks would reveal every noise term and
// with it the secret keys.  key = 256 bits (OS entropy by default, engine.cu), block counter and
// nonce = the caller's (row, index, stream tag) tuple; the first four words of the block are used.
struct ChaChaKey { uint32_t k[8]; };
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define CHACHA_QR(a, b, c, d)          \
  a += b; d ^= a; d = rotl32(d, 16);   \
  c += d; b ^= c; b = rotl32(b, 12);   \
  a += b; d ^= a; d = rotl32(d, 8);    \
  c += d; b ^= c; b = rotl32(b, 7);
__device__ __forceinline__ uint4 philox(uint4 ctr, const ChaChaKey &key) {   // (name kept: call sites)
  uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
  uint32_t x4 = key.k[0], x5 = key.k[1], x6 = key.k[2], x7 = key.k[3];
  uint32_t x8 = key.k[4], x9 = key.k[5], x10 = key.k[6], x11 = key.k[7];
  uint32_t x12 = ctr.x, x13 = ctr.y, x14 = ctr.z, x15 = ctr.w;
#pragma unroll
  for (int r = 0; r < 10; r++) {   // 20 rounds = 10 double rounds
    CHACHA_QR(x0, x4, x8, x12) CHACHA_QR(x1, x5, x9, x13) CHACHA_QR(x2, x6, x10, x14) CHACHA_QR(x3, x7, x11, x15)
    CHACHA_QR(x0, x5, x10, x15) CHACHA_QR(x1, x6, x11, x12) CHACHA_QR(x2, x7, x8, x13) CHACHA_QR(x3, x4, x9, x14)
  }
  return make_uint4(x0 + 0x61707865u, x1 + 0x3320646eu, x2 + 0x79622d32u, x3 + 0x6b206574u);
}
#undef CHACHA_QR
// utils.rs:9-12
__device__ __forceinline__ uint32_t f64_to_torus_dev(double d) {
  return (uint32_t)(unsigned long long)(long long)(fmod(d, 1.0) * 4294967296.0);
}
// utils.rs:22-38: f64_to_torus(N(0, alpha)); Box-Muller on two 32-bit uniforms
__device__ __forceinline__ uint32_t gaussian_torus(uint32_t u0, uint32_t u1, double alpha) {
  const double a = ((double)u0 + 0.5) * (1.0 / 4294967296.0);
  const double b = ((double)u1 + 0.5) * (1.0 / 4294967296.0);
  const double g = sqrt(-2.0 * log(a)) * cospi(2.0 * b);
  return f64_to_torus_dev(g * alpha);
}

struct KgArgs {
  const cplx *tw_a, *tw_b;
  const uint32_t *s0, *s1;   // device copies of the secret key (0/1 words)
  cplx *s1_spec;             // [64][8] spectrum of s1 in device order, pre-scaled by 1/512
  cplx *bsk;                 // device BSK layout
  uint32_t n, l, bgbit;
  double alpha;
  ChaChaKey seed;
};

__device__ __forceinline__ void load_tw(const KgArgs &a, int tid, cplx (&ta)[8], cplx (&tb)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) { ta[k] = a.tw_a[tid * 8 + k]; tb[k] = a.tw_b[(tid & 7) * 8 + k]; }
}

// spectrum of s1 (klemsa forward transform, unscaled) / 512, device order [v][k2]
__global__ void __launch_bounds__(64) kg_s1_spectrum_kernel(const KgArgs a) {
  __shared__ __align__(16) cplx exch[kExchStride];
  const int tid = threadIdx.x;
  cplx ta[8], tb[8];
  load_tw(a, tid, ta, tb);
  uint32_t xr[8], xi[8];
#pragma unroll
  for (int m = 0; m < 8; m++) { xr[m] = a.s1[64 * m + tid]; xi[m] = a.s1[64 * m + tid + kHalf]; }
  fwd_pass_a_i32(tid, xr, xi, ta, exch);
  __syncthreads();
  fwd_pass_b<1>(tid, tb, exch);
  __syncthreads();
  cplx out[8];
  fwd_pass_c(tid, exch, out);
#pragma unroll
  for (int k2 = 0; k2 < 8; k2++)
    a.s1_spec[tid * 8 + k2] = mk(out[k2].x * (1.0 / 512.0), out[k2].y * (1.0 / 512.0));
}

// one TRGSW row (i, r) per CTA
__global__ void __launch_bounds__(64) kg_bsk_row_kernel(const KgArgs a) {
  __shared__ __align__(16) cplx exch[2 * kExchStride];
  __shared__ uint32_t prod[2 * kN];  // inv_pass_a accumulates both outputs; [0] = a*s1
  const int tid = threadIdx.x;
  const uint32_t row = blockIdx.x;             // i * 2l + r
  const uint32_t i = row / (2 * a.l), r = row % (2 * a.l);
  cplx ta[8], tb[8];
  load_tw(a, tid, ta, tb);

  // a: uniform torus polynomial (trlwe.rs:38); noise e (trlwe.rs:40-44)
  uint32_t ar[8], ai[8], er[8], ei[8];
#pragma unroll
  for (int m = 0; m < 8; m++) {
    uint4 x = philox(make_uint4(row, 64 * m + tid, 0x42534B00u, 0), a.seed);
    uint4 y = philox(make_uint4(row, 64 * m + tid, 0x42534B01u, 0), a.seed);
    ar[m] = x.x; ai[m] = x.y;
    er[m] = gaussian_torus(x.z, x.w, a.alpha);
    ei[m] = gaussian_torus(y.x, y.y, a.alpha);
  }
  for (int x = tid; x < 2 * kN; x += 64) prod[x] = 0;

  // a * s1: forward(a) . spectrum(s1)/512 -> inverse   (trlwe.rs:45, klemsa.rs:152-174)
  fwd_pass_a_i32(tid, ar, ai, ta, exch);
  __syncthreads();
  fwd_pass_b<1>(tid, tb, exch);
  __syncthreads();
  cplx fa[8];
  fwd_pass_c(tid, exch, fa);
  cplx acc[2][8];
#pragma unroll
  for (int k2 = 0; k2 < 8; k2++) {
    acc[0][k2] = cmul(fa[k2], a.s1_spec[tid * 8 + k2]);
    acc[1][k2] = mk(0.0, 0.0);
  }
  __syncthreads();
  inv_pass_c(tid, tb, acc, exch);
  __syncthreads();
  inv_pass_b(tid, exch);
  __syncthreads();
  inv_pass_a<true>(tid, ta, exch, prod);   // |a*s1| < 2^41: exact, ties impossible
  __syncthreads();

  // b = e + a*s1 (trlwe.rs:47-49); gadget term mu * Bg^-(r'+1) on a[0] (r < l) or b[0]
  // (trgsw.rs:44-47), mu = s0[i]
  uint32_t br_[8], bi_[8];
#pragma unroll
  for (int m = 0; m < 8; m++) {
    br_[m] = er[m] + prod[64 * m + tid];
    bi_[m] = ei[m] + prod[64 * m + tid + kHalf];
  }
  const uint32_t rr = r < a.l ? r : r - a.l;
  const uint32_t gad = a.s0[i] * (1u << (32 - (rr + 1) * a.bgbit));  // f64_to_torus(Bg^-(rr+1))
  if (tid == 0) {
    if (r < a.l) ar[0] += gad; else br_[0] += gad;
  }

  // Fourier conversion of both halves straight into the device BSK layout (x 1/512)
  cplx *dst = a.bsk + (size_t)row * kChunkCplx;
  fwd_pass_a_i32(tid, ar, ai, ta, exch);
  fwd_pass_a_i32(tid, br_, bi_, ta, exch + kExchStride);
  __syncthreads();
  fwd_pass_b<2>(tid, tb, exch);
  __syncthreads();
#pragma unroll
  for (int o = 0; o < 2; o++) {
    cplx f[8];
    fwd_pass_c(tid, exch + o * kExchStride, f);
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++)
      dst[(k2 * 2 + o) * 64 + tid] = mk(f[k2].x * (1.0 / 512.0), f[k2].y * (1.0 / 512.0));
  }
}

// key.rs:102-122 in the reference row layout u32[N*t*base][n+1]; one warp per row, k = 0 rows zero
__global__ void kg_ksk_kernel(const uint32_t *__restrict__ s0, const uint32_t *__restrict__ s1,
                              uint32_t *__restrict__ ksk, uint32_t n, uint32_t basebit, uint32_t t,
                              double alpha, ChaChaKey seed) {
  const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t base = 1u << basebit;
  if (row >= br::kN * t * base) return;
  const uint32_t k = row & (base - 1), j = (row >> basebit) % t, i = (row >> basebit) / t;
  uint32_t *dst = ksk + (size_t)row * (n + 1);
  if (k == 0) {
    for (uint32_t x = lane; x <= n; x += 32) dst[x] = 0;
    return;
  }
  uint32_t inner = 0;
  for (uint32_t x0 = 0; x0 < n; x0 += 128) {   // 4 words per Philox call per lane
    uint4 v = philox(make_uint4(row, x0 / 128 * 32 + lane, 0x4B534B00u, 0), seed);
    const uint32_t vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t x = x0 + 4 * lane + c;
      if (x < n) { dst[x] = vals[c]; inner += vals[c] * s0[x]; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) inner += __shfl_xor_sync(0xffffffffu, inner, o);
  if (lane == 0) {
    uint4 v = philox(make_uint4(row, 0xFFFFFFFFu, 0x4B534B01u, 0), seed);
    // mu = k * s1[i] / 2^((j+1)*basebit)  (key.rs:112-113): an exact power-of-two fraction
    const uint32_t mu = (k * s1[i]) << (32 - (j + 1) * basebit);
    dst[n] = inner + gaussian_torus(v.x, v.y, alpha) + mu;
  }
}

}  // namespace

cudaError_t keygen_launch(const cplx *tw_a, const cplx *tw_b, const uint32_t *d_s0,
                          const uint32_t *d_s1, cplx *d_s1_spec, cplx *d_bsk, uint32_t *d_ksk_ref,
                          uint32_t n, uint32_t l, uint32_t bgbit, uint32_t basebit, uint32_t t,
                          double alpha_lv0, double alpha_lv1, const uint32_t key256[8], cudaStream_t stream) {
  KgArgs a{};
  a.tw_a = tw_a; a.tw_b = tw_b; a.s0 = d_s0; a.s1 = d_s1; a.s1_spec = d_s1_spec; a.bsk = d_bsk;
  a.n = n; a.l = l; a.bgbit = bgbit; a.alpha = alpha_lv1;
  for (int i = 0; i < 8; i++) a.seed.k[i] = key256[i];
  ChaChaKey ksk_key = a.seed;
  ksk_key.k[7] ^= 0x4B534B4Bu;   // independent stream for the key-switching key
  kg_s1_spectrum_kernel<<<1, 64, 0, stream>>>(a);
  kg_bsk_row_kernel<<<n * 2 * l, 64, 0, stream>>>(a);
  const uint32_t rows = br::kN * t * (1u << basebit);
  kg_ksk_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(d_s0, d_s1, d_ksk_ref, n, basebit, t, alpha_lv0, ksk_key);
  return cudaGetLastError();
}
