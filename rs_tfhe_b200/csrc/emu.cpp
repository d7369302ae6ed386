// emu.cpp -- CPU emulation of the blind-rotation kernel's per-thread phases.
//
// TEST INFRASTRUCTURE: compiled with g++ into libtfhe_emu.so and driven by
// tests/test_emulator.py.  It executes exactly the __host__ __device__ functions
// of br_core.cuh that blind_rotate.cu runs on the GPU, one "thread" at a time
// with the barriers replaced by loop boundaries, so that the index, twiddle and
// layout logic is checked against the oracle without a GPU.  Never used by the
// product path.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "br_core.cuh"

using namespace br;

namespace {

struct Twiddles { cplx ta[8]; cplx tb[8]; };

void make_twiddles(int tid, Twiddles &tw) {
  for (int k0 = 0; k0 < 8; k0++) {
    int idx = ((tid * (1 - 4 * k0)) % 2048 + 2048) % 2048;
    double a = M_PI * (double)idx / 1024.0;
    tw.ta[k0] = mk(std::cos(a), std::sin(a));
  }
  cplx c[8];
  for (int x = 0; x < 8; x++) {
    int idx = ((tid & 7) * x) % 64;
    double a = -2.0 * M_PI * (double)idx / 64.0;
    c[x] = mk(std::cos(a), std::sin(a));
  }
  expand_tb(c[1], c[2], c[4], tw.tb);  // the kernels rebuild tb from three base values
}

// one sub-round of the forward transform: digits [D0, D0+ND) of polynomial p
template <int L, int BGBIT, int D0, int ND>
void fwd_round(int p, const uint32_t (*t_re)[8], const uint32_t (*t_im)[8], const Twiddles *tw,
               cplx *exch, const cplx *row, cplx (*racc)[2][8]) {
  for (int t = 0; t < kGroup; t++) fwd_pass_a<BGBIT, D0, ND>(t, t_re[t], t_im[t], tw[t].ta, exch);
  for (int t = 0; t < kGroup; t++) fwd_pass_b<ND>(t, tw[t].tb, exch);
  for (int d = 0; d < ND; d++)
    for (int t = 0; t < kGroup; t++)
      fwd_pass_c_mac(t, exch + d * kExchStride, row + (p * L + D0 + d) * kChunkCplx, racc[t]);
}

template <int L, int BGBIT, bool EXACT>
void run(uint32_t n, uint32_t offset, const double *bsk_ref, const uint32_t *tv,
         const uint32_t *lwe, int steps, uint32_t *out) {
  constexpr int L2 = 2 * L;
  constexpr int NBUF = 2;  // the G=6 kernel's budget: digits go through in sub-rounds of <= 2
  std::vector<uint32_t> acc(2 * kN);
  std::vector<cplx> exch(NBUF * kExchStride);
  std::vector<cplx> row(L2 * kChunkCplx);
  static Twiddles tw[kGroup];
  static cplx racc[kGroup][2][8];
  for (int t = 0; t < kGroup; t++) make_twiddles(t, tw[t]);

  uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)lwe[n] + (1u << 20)) >> 21));
  for (int o = 0; o < 2; o++)
    for (int j = 0; j < kN; j++) acc[o * kN + j] = rot_coeff(tv + o * kN, j, b_tilda);

  uint32_t count = steps < 0 ? n : (uint32_t)steps;
  for (uint32_t i = 0; i < count; i++) {
    // upload permutation of BSK[i] (same index math as the relayout kernel)
    for (int r = 0; r < L2; r++)
      for (int k2 = 0; k2 < 8; k2++)
        for (int o = 0; o < 2; o++)
          for (int v = 0; v < 64; v++) {
            const double *src = bsk_ref + (((size_t)i * L2 + r) * 2 + o) * kN;
            int k = bin_of(v, k2);
            row[r * kChunkCplx + (k2 * 2 + o) * 64 + v] =
                mk(src[k] * (1.0 / 1024.0), src[k + kHalf] * (1.0 / 1024.0));
          }
    uint32_t abar = (uint32_t)(lwe[i] + (1u << 20)) >> 21;
    memset(racc, 0, sizeof(racc));
    static uint32_t t_re[kGroup][8], t_im[kGroup][8];
    for (int p = 0; p < 2; p++) {
      for (int t = 0; t < kGroup; t++) load_t(t, acc.data() + p * kN, abar, offset, t_re[t], t_im[t]);
      if (L == 3) {
        fwd_round<L, BGBIT, 0, 2>(p, t_re, t_im, tw, exch.data(), row.data(), racc);
        fwd_round<L, BGBIT, (L == 3 ? 2 : 0), 1>(p, t_re, t_im, tw, exch.data(), row.data(), racc);
      } else {
        fwd_round<L, BGBIT, 0, (L == 3 ? 1 : L)>(p, t_re, t_im, tw, exch.data(), row.data(), racc);
      }
    }
    for (int t = 0; t < kGroup; t++) inv_pass_c(t, tw[t].tb, racc[t], exch.data());
    for (int t = 0; t < kGroup; t++) inv_pass_b(t, exch.data());
    for (int t = 0; t < kGroup; t++) inv_pass_a<EXACT>(t, tw[t].ta, exch.data(), acc.data());
  }
  memcpy(out, acc.data(), 2 * kN * sizeof(uint32_t));
}

}  // namespace

extern "C" int emu_blind_rotate(uint32_t n, uint32_t l, uint32_t bgbit, uint32_t offset,
                                const double *bsk_ref, const uint32_t *tv, const uint32_t *lwe,
                                int steps, uint32_t *out_trlwe) {
  if (l == 3 && bgbit == 6) run<3, 6, true>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 2 && bgbit == 10) run<2, 10, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 1 && bgbit == 18) run<1, 18, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 1 && bgbit == 22) run<1, 22, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 1 && bgbit == 23) run<1, 23, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else return -1;
  return 0;
}

// ---- 128-thread kernel (blind_rotate_s.cu / brs_core.cuh) ----------------------------------------
// Tensor memory is modelled per warp as rows x value slots with the access shapes the kernel uses:
//   forward exchange  tcgen05.st.32x32b.x16 (lane l' writes slot s to row l')  then two
//                     tcgen05.ld.16x256b.x2 (lane l reads value 2H+h from row 16H + (l>>2) + 8h, slot l&3)
//   inverse exchange  the mirror image (st.16x256b.x2, ld.32x32b.x16)
// (shapes verified on hardware by tools/probe/tmem_xchg_probe.cu).
#include "brs_core.cuh"

namespace {

struct WarpTmem { cplx v[32][4]; };

void emu_xchg_fwd(cplx (*y)[4] /*[128][4]*/) {
  for (int W = 0; W < 4; W++) {
    WarpTmem tm;
    for (int l = 0; l < 32; l++)
      for (int s = 0; s < 4; s++) tm.v[l][s] = y[32 * W + l][s];
    for (int l = 0; l < 32; l++)
      for (int H = 0; H < 2; H++)
        for (int h = 0; h < 2; h++) y[32 * W + l][2 * H + h] = tm.v[16 * H + (l >> 2) + 8 * h][l & 3];
  }
}
void emu_xchg_inv(cplx (*u)[4]) {
  for (int W = 0; W < 4; W++) {
    WarpTmem tm;
    for (int l = 0; l < 32; l++)
      for (int H = 0; H < 2; H++)
        for (int h = 0; h < 2; h++) tm.v[16 * H + (l >> 2) + 8 * h][l & 3] = u[32 * W + l][2 * H + h];
    for (int l = 0; l < 32; l++)
      for (int s = 0; s < 4; s++) u[32 * W + l][s] = tm.v[l][s];
  }
}

template <int L, int BGBIT, bool EXACT>
void run_s(uint32_t n, uint32_t offset, const double *bsk_ref, const uint32_t *tv, const uint32_t *lwe,
           int steps, uint32_t *out) {
  constexpr int L2 = 2 * L;
  constexpr int kT = brs::kT;
  std::vector<uint32_t> acc(2 * kN);
  std::vector<cplx> exch(L * kHalf > 2 * 8 * brs::kInvPitch ? L * kHalf : 2 * 8 * brs::kInvPitch);
  std::vector<cplx> row(L2 * brs::kRowCplx);
  static cplx tw[kT][brs::kTwPerThread];
  static cplx racc[kT][2][4];
  static cplx y[kT][4];
  for (int t = 0; t < kT; t++) brs::make_tw(t, tw[t]);

  uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)lwe[n] + (1u << 20)) >> 21));
  for (int o = 0; o < 2; o++)
    for (int j = 0; j < kN; j++) acc[o * kN + j] = rot_coeff(tv + o * kN, j, b_tilda);

  uint32_t count = steps < 0 ? n : (uint32_t)steps;
  for (uint32_t i = 0; i < count; i++) {
    // upload permutation of BSK[i] (same index math as bsk_permute_s_kernel)
    for (int r = 0; r < L2; r++)
      for (int kd = 0; kd < 4; kd++)
        for (int o = 0; o < 2; o++)
          for (int t = 0; t < kT; t++) {
            const double *src = bsk_ref + (((size_t)i * L2 + r) * 2 + o) * kN;
            int k = brs::bin_of(t, kd);
            row[r * brs::kRowCplx + brs::row_index(kd, o, t)] =
                mk(src[k] * (1.0 / 1024.0), src[k + kHalf] * (1.0 / 1024.0));
          }
    uint32_t abar = (uint32_t)(lwe[i] + (1u << 20)) >> 21;
    memset(racc, 0, sizeof(racc));
    static uint32_t t_re[kT][4], t_im[kT][4];
    for (int p = 0; p < 2; p++) {
      for (int t = 0; t < kT; t++) brs::load_t(t, acc.data() + p * kN, abar, offset, t_re[t], t_im[t]);
      for (int d = 0; d < L; d++)
        for (int t = 0; t < kT; t++)
          brs::fwd_pass_a<BGBIT, false>(t, d, t_re[t], t_im[t], exch.data() + d * kHalf);
      for (int d = 0; d < L; d++) {
        for (int t = 0; t < kT; t++)
          brs::fwd_pass_b(t, exch.data() + d * kHalf, tw[t][brs::TW_B], tw[t][brs::TW_B + 1],
                          tw[t][brs::TW_B + 2], y[t]);
        emu_xchg_fwd(y);
        for (int t = 0; t < kT; t++) brs::r4<false>(y[t], tw[t][brs::TW_C], tw[t][brs::TW_C + 1]);
        emu_xchg_fwd(y);
        for (int t = 0; t < kT; t++) brs::r4<false>(y[t], tw[t][brs::TW_D], tw[t][brs::TW_D + 1]);
        const cplx *rw = row.data() + (p * L + d) * brs::kRowCplx;
        for (int t = 0; t < kT; t++)
          for (int kd = 0; kd < 4; kd++) {
            cfma(racc[t][0][kd], y[t][kd], rw[brs::row_index(kd, 0, t)]);
            cfma(racc[t][1][kd], y[t][kd], rw[brs::row_index(kd, 1, t)]);
          }
      }
    }
    for (int o = 0; o < 2; o++) {
      for (int t = 0; t < kT; t++) { brs::r4_plain<true>(racc[t][o]); for (int k = 0; k < 4; k++) y[t][k] = racc[t][o][k]; }
      emu_xchg_inv(y);
      for (int t = 0; t < kT; t++) brs::r4<true>(y[t], tw[t][brs::TW_CI], tw[t][brs::TW_CI + 1]);
      emu_xchg_inv(y);
      for (int t = 0; t < kT; t++) brs::r4<true>(y[t], tw[t][brs::TW_BI], tw[t][brs::TW_BI + 1]);
      for (int t = 0; t < kT; t++) brs::inv_store_b(t, y[t], exch.data() + o * (8 * brs::kInvPitch));
    }
    for (int o = 0; o < 2; o++)
      for (int t = 0; t < kT; t++) {
        cplx ut[4];
        for (int k = 0; k < 4; k++) ut[k] = tw[t][brs::TW_UT + k];
        brs::inv_pass_a<EXACT, false>(t, exch.data() + o * (8 * brs::kInvPitch), tw[t][brs::TW_AI],
                                      tw[t][brs::TW_AI + 1], tw[t][brs::TW_AI + 2], ut, acc.data() + o * kN);
      }
  }
  memcpy(out, acc.data(), 2 * kN * sizeof(uint32_t));
}

}  // namespace

extern "C" int emu_blind_rotate_s(uint32_t n, uint32_t l, uint32_t bgbit, uint32_t offset,
                                  const double *bsk_ref, const uint32_t *tv, const uint32_t *lwe,
                                  int steps, uint32_t *out_trlwe) {
  if (l == 3 && bgbit == 6) run_s<3, 6, true>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 2 && bgbit == 10) run_s<2, 10, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 1 && bgbit == 18) run_s<1, 18, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 1 && bgbit == 22) run_s<1, 22, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else if (l == 1 && bgbit == 23) run_s<1, 23, false>(n, offset, bsk_ref, tv, lwe, steps, out_trlwe);
  else return -1;
  return 0;
}
// the twiddle table the engine uploads (cplx[128][20] as doubles)
extern "C" void emu_s_twiddles(double *out) {
  for (int t = 0; t < brs::kT; t++) {
    cplx tw[brs::kTwPerThread];
    brs::make_tw(t, tw);
    for (int k = 0; k < brs::kTwPerThread; k++) { out[(t * brs::kTwPerThread + k) * 2] = tw[k].x; out[(t * brs::kTwPerThread + k) * 2 + 1] = tw[k].y; }
  }
}

// ---- model of the tcgen05 key switch (keyswitch_umma.cu) -----------------------------------
// Same index arithmetic as the device code (ku_layout.h): emu_ku_build_key is the relayout
// kernel; emu_ku_key_switch walks CTA tiles, pipeline stages, K steps and accumulator halves as the
// kernel does, builds the one-hot words with the builder warps' expression, and reads the B
// operand through the canonical K-major / no-swizzle addressing the MMA applies
// (B[row][k] at (row/8)*256 + (k/16)*128 + (row%8)*16 + k%16).  Checked against the oracle's
// identity_key_switching in tests/test_ks_umma_layout.py.
#include "ku_layout.h"

extern "C" size_t emu_ku_key_words(uint32_t n, uint32_t t, uint32_t basebit) {
  return ku::key_words(n, t, basebit);
}

extern "C" void emu_ku_build_key(const uint32_t *ksk_ref /*[1024*t*2^basebit][n+1]*/, uint32_t n,
                                 uint32_t t, uint32_t basebit, uint32_t *dst) {
  const size_t total = ku::key_words(n, t, basebit);
  for (size_t idx = 0; idx < total; idx++) {
    const ku::Src s = ku::decode(idx, t, basebit);
    uint32_t v = 0;
    if (s.x <= n)
      for (uint32_t b = 0; b < 4; b++) {
        if (s.k0 + b == 0) continue;
        const uint32_t wv = ksk_ref[((size_t)s.row0 + b) * (n + 1) + s.x];
        v |= ((wv >> (8 * s.plane)) & 0xFFu) << (8 * b);
      }
    dst[idx] = v;
  }
}

extern "C" int emu_ku_key_switch(const uint32_t *key_words, const uint32_t *ext /*[count][1025]*/,
                                 size_t count, uint32_t n, uint32_t t, uint32_t bb,
                                 uint32_t *out /*[count][n+1]*/) {
  if (bb < 2 || bb > 6) return -1;
  const uint8_t *key = reinterpret_cast<const uint8_t *>(key_words);
  const uint32_t nst = ku::n_stages(t, bb), prec = 1u << (32 - (1 + bb * t));
  const uint32_t P = 1u << bb, pps = ku::kStageK / P, spb = 16 * t / pps;
  const size_t mtiles = (count + ku::kM - 1) / ku::kM;
  std::vector<int32_t> D(ku::kCols);
  for (uint32_t nt = 0; nt < ku::n_tiles(n); nt++)
    for (size_t mt = 0; mt < mtiles; mt++)
      for (int row = 0; row < ku::kM; row++) {
        const size_t ct = mt * ku::kM + row;
        if (ct >= count) break;
        std::fill(D.begin(), D.end(), 0);
        uint32_t st = 0;
        for (uint32_t blk = 0; blk < ku::kRing / 16; blk++) {     // builder: 16 coefficients
          uint32_t ab[16];
          for (int c = 0; c < 16; c++) ab[c] = ext[ct * (ku::kRing + 1) + 16 * blk + c] + prec;
          for (uint32_t s = 0; s < spb; s++, st++) {              // one pipeline stage
            uint32_t r[16];
            if (bb == 2) {
              for (int c = 0; c < 16; c++) {
                const uint32_t qb = 16 * s + c, il = qb / t, j = qb % t;
                r[c] = 1u << ((ab[il] >> (27 - 2 * j)) & 0x18u);
              }
            } else {
              for (uint32_t pp = 0; pp < pps; pp++) {
                const uint32_t qb = pps * s + pp, il = qb / t, j = qb % t;
                const uint32_t digit = (ab[il] >> (32 - (j + 1) * bb)) & (P - 1);
                const uint32_t word = digit >> 2, bit = 1u << ((digit & 3u) * 8);
                for (uint32_t wq = 0; wq < P / 4; wq++) r[pp * (P / 4) + wq] = (word == wq) ? bit : 0u;
              }
            }
            const uint8_t *stage = key + ((size_t)nt * nst + st) * ku::kStageBytes;
            for (int ks = 0; ks < 2; ks++)
              for (int h = 0; h < 2; h++) {
                const uint8_t *tile = stage + (ks * 2 + h) * ku::kTileBytes;
                for (int k = 0; k < ku::kStepK; k++) {
                  const uint32_t a = (r[ks * 8 + k / 4] >> (8 * (k % 4))) & 0xFFu;  // A[row][k]
                  if (!a) continue;
                  for (int nr = 0; nr < ku::kHalf; nr++)
                    D[h * ku::kHalf + nr] +=
                        (int32_t)(a * tile[(nr / 8) * 256 + (k / 16) * 128 + (nr % 8) * 16 + k % 16]);
                }
              }
          }
        }
        for (int xl = 0; xl < ku::kWords; xl++) {                 // epilogue
          const uint32_t x = nt * ku::kWords + xl;
          if (x > n) break;
          const uint32_t sum = (uint32_t)D[4 * xl] + ((uint32_t)D[4 * xl + 1] << 8) +
                               ((uint32_t)D[4 * xl + 2] << 16) + ((uint32_t)D[4 * xl + 3] << 24);
          const uint32_t init = (x == n) ? ext[ct * (ku::kRing + 1) + ku::kRing] : 0u;
          out[ct * (n + 1) + x] = init - sum;
        }
      }
  return 0;
}
