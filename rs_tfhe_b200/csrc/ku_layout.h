// ku_layout.h -- operand layout of the tcgen05 key switch (keyswitch_umma.cu).
//
// __host__ __device__ index arithmetic only, shared by the relayout kernel, the GEMM kernel and
// the CPU model (emu.cpp -> tests/test_ks_umma_layout.py), so the packing can be checked against
// the oracle's key switch without a GPU.
//
// GEMM view of identity_key_switching (reference src/trgsw.rs:332-360), P = 2^basebit:
//   D[ct][col] = sum_K A[ct][K] * B[col][K],   K = P*q + k,  q = i*t + j  (coefficient i, digit j)
//   A[ct][Pq+k] = [digit_j(a_i + PREC_OFFSET) == k]            (one-hot, built on the fly)
//   B[col][Pq+k] = byte `plane` of KSK[P q + k][x], col = 4*xl + plane, x = 120*ntile + xl
//                  (k = 0 forced to 0: a zero digit subtracts nothing, trgsw.rs:351)
//   out[ct][x] = (x == n ? b : 0) - sum_plane D[ct][4 xl + plane] << (8 plane)   (wrapping)
// One CTA owns 128 ciphertexts x 480 columns (120 output words); a pipeline stage is K = 64
// (64 / P (i,j) pairs = two K=32 MMA steps; basebit 2..6), and an MMA B operand is a 240 x 32 byte
// tile in the canonical K-major / no-swizzle order [n/8][k/16][n%8][16 B] (SBO 256 B, LBO 128 B).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KU_HD __host__ __device__ __forceinline__
#else
#define KU_HD inline
#endif

namespace ku {

constexpr int kM = 128;          // ciphertexts per CTA = TMEM lanes
constexpr int kWords = 120;      // output words per CTA
constexpr int kCols = 4 * kWords;  // accumulator columns (byte planes)
constexpr int kHalf = kCols / 2;   // N of one MMA
constexpr int kStepK = 32;       // K of one MMA
constexpr int kStageK = 64;      // K of one pipeline stage (64 / 2^basebit pairs)
constexpr int kTileBytes = kHalf * kStepK;       // 7680: one B operand
constexpr int kStageBytes = 4 * kTileBytes;      // 30720: [kstep 2][half 2] tiles
constexpr int kRing = 1024;      // level-1 dimension (input coefficients)

KU_HD uint32_t n_tiles(uint32_t n) { return (n + 1 + kWords - 1) / kWords; }
KU_HD uint32_t n_stages(uint32_t t, uint32_t basebit) {
  return (((uint32_t)kRing * t) << basebit) / kStageK;
}
KU_HD size_t key_words(uint32_t n, uint32_t t, uint32_t basebit) {
  return (size_t)n_tiles(n) * n_stages(t, basebit) * (kStageBytes / 4);
}

// Destination word `idx` of the device key (its 4 bytes are k = k0..k0+3 of one (column, pair))
struct Src {
  uint32_t row0;   // source row of the first byte: P*q + k0 (rows row0..row0+3, same pair)
  uint32_t k0;     // digit value of the first byte (multiple of 4)
  uint32_t x;      // output word (column of the caller's KSK rows); may exceed n (padding)
  uint32_t plane;  // byte of the source word
};
KU_HD Src decode(size_t idx, uint32_t t, uint32_t basebit) {
  const uint32_t words_per_stage = kStageBytes / 4, words_per_tile = kTileBytes / 4;
  const uint32_t in_stage = (uint32_t)(idx % words_per_stage);
  const size_t stage_lin = idx / words_per_stage;
  const uint32_t st = (uint32_t)(stage_lin % n_stages(t, basebit));
  const uint32_t nt = (uint32_t)(stage_lin / n_stages(t, basebit));
  const uint32_t tile = in_stage / words_per_tile, w = in_stage % words_per_tile;
  const uint32_t kstep = tile >> 1, half = tile & 1;
  // canonical tile, in words: n1 * 64 + k1 * 32 + r0 * 4 + (k % 16) / 4
  const uint32_t n1 = w >> 6, k1 = (w >> 5) & 1, r0 = (w >> 2) & 7, cw = w & 3;
  const uint32_t col = half * kHalf + n1 * 8 + r0;
  Src s;
  s.row0 = st * kStageK + kstep * kStepK + k1 * 16 + cw * 4;   // K index = P*q + k = source row
  s.k0 = s.row0 & ((1u << basebit) - 1u);
  s.x = nt * kWords + (col >> 2);
  s.plane = col & 3;
  return s;
}

}  // namespace ku
