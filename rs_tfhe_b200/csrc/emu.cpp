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
