// blind_rotate.cu -- K0+K3: gate pre-combination + persistent blind rotation.
//
// Replaces (reference, file:line under rs-tfhe):
//   gates.rs:366-373 (batch prep), trgsw.rs:198-274 (blind_rotate[_with_testvec]),
//   trgsw.rs:174-196 (cmux), :77-142 (external product), :144-171 (decomposition),
//   :307-330 (X^k), fft/klemsa.rs:88-150 (transforms), trlwe.rs:106-136 (extract).
//
// Shape: one persistent CTA per SM.  A CTA owns G ciphertexts at a time, 64
// threads each (br_core.cuh), and walks the n CMUX steps; the whole external
// product of a step stays on chip (accumulator + exchange buffers in shared
// memory, spectra and MAC accumulators in registers).  A dedicated producer warp
// (in its own warpgroup, so setmaxnreg can hand its registers to the consumers) streams the Fourier-domain bootstrapping key through a ring of 16 KB stages
// with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); every staged
// row is consumed by all G ciphertext groups before its slot is released, so the
// key crosses L2->SM once per CTA per step regardless of G.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "br_core.cuh"
#include "kernels.h"

using namespace br;

#ifndef BR_PRODUCER_SLEEP_NS
#define BR_PRODUCER_SLEEP_NS 256
#endif
#ifndef BR_DEFAULT_VARIANT
#define BR_DEFAULT_VARIANT 8
#endif

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// producer-side wait: back off between polls so the spin does not steal issue slots
// from the consumer warps sharing its SM sub-partition
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(BR_PRODUCER_SLEEP_NS);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void group_sync(int g) {
  asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory");
}

// gates.rs:54-150: out = ca*a + cb*b, b-word += off
__constant__ int32_t c_gate_ca[TFHE_GATE_COUNT] = {-1, 1, 1, 1, 1, -1, -1, 1, -1, 1};
__constant__ int32_t c_gate_cb[TFHE_GATE_COUNT] = {-1, 1, 1, 2, -2, -1, 1, -1, 1, -1};
__constant__ uint32_t c_gate_off[TFHE_GATE_COUNT] = {0x20000000u, 0xE0000000u, 0x20000000u,
                                                     0x40000000u, 0xC0000000u, 0xE0000000u,
                                                     0xE0000000u, 0xE0000000u, 0x20000000u,
                                                     0x20000000u};

constexpr int kStageBytes = kChunkCplx * 16;  // 16 KB: one BSK row

// ---- TMEM as per-thread private scratch (pass-A twiddles) --------------------------
// Warp w may touch TMEM lanes 32*(w%4)..+31; a 32x32b access gives every thread of the
// warp its own lane, i.e. private storage with a data path separate from shared memory.
__device__ __forceinline__ void tmem_st_ta(uint32_t taddr, const cplx (&ta)[8]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r[4 * i + 0] = (uint32_t)__double2loint(ta[i].x); r[4 * i + 1] = (uint32_t)__double2hiint(ta[i].x);
    r[4 * i + 2] = (uint32_t)__double2loint(ta[i].y); r[4 * i + 3] = (uint32_t)__double2hiint(ta[i].y);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_ta(uint32_t taddr, cplx (&ta)[8]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ta[i].x = __hiloint2double((int)r[4 * i + 1], (int)r[4 * i + 0]);
    ta[i].y = __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]);
  }
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// park / restore one output's 8 complex accumulators (32 words) at TMEM column `col`
__device__ __forceinline__ void park8(uint32_t taddr, const cplx (&a)[8]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r[4 * i + 0] = (uint32_t)__double2loint(a[i].x); r[4 * i + 1] = (uint32_t)__double2hiint(a[i].x);
    r[4 * i + 2] = (uint32_t)__double2loint(a[i].y); r[4 * i + 3] = (uint32_t)__double2hiint(a[i].y);
  }
  tmem_st32(taddr, r);
}
__device__ __forceinline__ void unpark8(uint32_t taddr, cplx (&a)[8]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i].x = __hiloint2double((int)r[4 * i + 1], (int)r[4 * i + 0]);
    a[i].y = __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]);
  }
}

template <int L, int NBUF> struct Cfg {
  static constexpr int kAccBytes = 2 * kN * 4;
  static constexpr int kExchBytes = NBUF * kExchStride * 16;
  static constexpr int kAbarBytes = 2432;  // u16[n], n <= 1216
  static constexpr int kGroupBytes = kAccBytes + kExchBytes + kAbarBytes;
};

// Kernel variants (same per-thread code, different residency):
//   V1: G=4 groups, 3 exchange buffers, twiddles in registers   (consumers 232 regs)
//   V2: G=6 groups, 2 exchange buffers (digits in sub-rounds of <=2), pass-A twiddles in
//       TMEM, pass-B twiddles rebuilt from three base values     (consumers 160 regs)
//   MAGIC: int<->double conversions of the exact regime as 2^52-biased bit patterns + one DADD
//       (FP64 pipe) instead of I2F/F2I (quarter-rate conversion pipe).
template <int L, int BGBIT, int G, int STAGES, int NBUF, bool TMEM_TW, int REGS_CONS, int REGS_PROD,
          bool PARK = false, bool MAGIC = false>
__global__ void __launch_bounds__(((2 * G + 3) / 4) * 128 + 128, 1) blind_rotate_kernel(const BrArgs args) {
  static_assert(!MAGIC || (L == 3 && BGBIT == 6), "MAGIC conversions need the exact regime");
  using C = Cfg<L, NBUF>;
  constexpr int PW = ((2 * G + 3) / 4) * 4;  // first producer warp: its own warpgroup
  constexpr int L2 = 2 * L;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  static_assert(NBUF >= 2 && (L <= NBUF || (L == 3 && NBUF == 2)), "unsupported buffer plan");
  constexpr int ND0 = L <= NBUF ? L : 2;   // digits in the first sub-round
  constexpr int ND1 = L - ND0;             // and in the second (0 or 1)
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint8_t *groups = smem + STAGES * kStageBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(groups + G * C::kGroupBytes);
  uint64_t *empty = full + STAGES;
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(empty + STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n;
  const uint32_t grid = gridDim.x;
  const uint32_t per_round = grid * G;
  const uint32_t rounds = (uint32_t)((args.count + per_round - 1) / per_round);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2 * G);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (TMEM_TW && warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (TMEM_TW) asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (TMEM_TW) asm volatile("tcgen05.fence::after_thread_sync;");

  // Register budget: the SM sub-partition hosting the producer warpgroup also hosts
  // consumer warps, so the launch is compiled at 65536/blockDim registers/thread and
  // rebalanced here (SASS: USETMAXREG).
  if (warp >= PW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PROD));
    // ===== producer: stream BSK rows (i, r) for every round =====
    if (warp == PW && lane == 0) {
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk);
      uint32_t stage = 0, parity = 0;
      const uint32_t rows = n * L2;
      for (uint32_t rd = 0; rd < rounds; rd++) {
        for (uint32_t row = 0; row < rows; row++) {
          mbar_wait_backoff(&empty[stage], parity ^ 1);
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tma_load_1d(reinterpret_cast<uint8_t *>(ring) + stage * kStageBytes,
                      src0 + (size_t)row * kStageBytes, kStageBytes, &full[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumers: group g owns one ciphertext per round =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_CONS));
  if (warp >= 2 * G) return;  // padding warps of a partially filled consumer warpgroup (odd G)
  const int g = warp >> 1;
  const int tid = threadIdx.x & 63;
  uint8_t *gbase = groups + g * C::kGroupBytes;
  uint32_t *acc = reinterpret_cast<uint32_t *>(gbase);
  cplx *exch = reinterpret_cast<cplx *>(gbase + C::kAccBytes);
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(gbase + C::kAccBytes + C::kExchBytes);

  // twiddles: V1 keeps both sets in registers; V2 parks ta in TMEM and keeps 3 tb bases
  cplx ta_reg[TMEM_TW ? 1 : 8], tb_reg[TMEM_TW ? 1 : 8];
  cplx tb1, tb2, tb4;
  uint32_t taddr = 0;
  {
    const cplx *twb = args.tw_b + (tid & 7) * 8;
    tb1 = twb[1]; tb2 = twb[2]; tb4 = twb[4];
    if constexpr (TMEM_TW) {
      cplx ta[8];
#pragma unroll
      for (int k = 0; k < 8; k++) ta[k] = args.tw_a[tid * 8 + k];
      taddr = *tmem_base_s + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 96u;
      tmem_st_ta(taddr, ta);
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) { ta_reg[k] = args.tw_a[tid * 8 + k]; tb_reg[k] = twb[k]; }
    }
  }
#define BR_GET_TA(dst)                                   \
  cplx dst[8];                                           \
  if constexpr (TMEM_TW) tmem_ld_ta(taddr, dst);         \
  else { _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) dst[k_] = ta_reg[k_]; }
#define BR_GET_TB(dst)                                   \
  cplx dst[8];                                           \
  if constexpr (TMEM_TW) expand_tb(tb1, tb2, tb4, dst);  \
  else { _Pragma("unroll") for (int k_ = 0; k_ < 8; k_++) dst[k_] = tb_reg[k_]; }
#define BR_MAC_DIGITS(ND)                                                                   \
  _Pragma("unroll") for (int d = 0; d < (ND); d++) {                                        \
    mbar_wait(&full[stage], parity);                                                        \
    fwd_pass_c_mac(tid, exch + d * kExchStride, ring + stage * kChunkCplx, racc);           \
    __syncwarp();                                                                           \
    if (lane == 0) mbar_arrive(&empty[stage]);                                              \
    if (++stage == STAGES) { stage = 0; parity ^= 1; }                                      \
  }

  const uint32_t w = n + 1;
  uint32_t stage = 0, parity = 0;

  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = ((size_t)rd * G + g) * grid + blockIdx.x;
    const bool active = ct < args.count;

    if (active) {
      // ---- K0: linear pre-combination + modulus switch (trgsw.rs:202-203, 210-211)
      uint32_t ca = 1, cb = 0, off = 0;
      const uint32_t *A, *B;
      if (args.op >= 0 || args.ops) {
        int op = args.ops ? (int)args.ops[ct] : args.op;
        ca = (uint32_t)c_gate_ca[op]; cb = (uint32_t)c_gate_cb[op]; off = c_gate_off[op];
        A = args.in + ct * 2 * w;
        B = A + w;
      } else {
        A = args.in + ct * w;
        B = A;
      }
      for (uint32_t i = tid; i < n; i += 64) {
        uint32_t v = ca * A[i] + cb * B[i];
        abar_s[i] = (uint16_t)((uint32_t)(v + (1u << 20)) >> 21);
      }
      uint32_t bw = ca * A[n] + cb * B[n] + off;
      uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)bw + (1u << 20)) >> 21));
      const int tvi = args.tv_index ? args.tv_index[ct] : args.tv_default;
      const uint32_t *tv = args.tv + (size_t)tvi * 2 * kN;
      for (int x = tid; x < 2 * kN; x += 64)
        acc[x] = rot_coeff(tv + (x & ~(kN - 1)), x & (kN - 1), b_tilda);
    }
    group_sync(g);

    for (uint32_t i = 0; i < n; i++) {
      if (active) {
        cplx racc[2][8];
        const uint32_t abar = abar_s[i];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
          for (int k = 0; k < 8; k++) racc[o][k] = mk(0.0, 0.0);
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
          uint32_t t_re[8], t_im[8];
          load_t(tid, acc + p * kN, abar, args.offset, t_re, t_im);
          {
            BR_GET_TA(ta)
            fwd_pass_a<BGBIT, 0, ND0, MAGIC>(tid, t_re, t_im, ta, exch);
          }
          group_sync(g);
          {
            BR_GET_TB(tb)
            fwd_pass_b<ND0>(tid, tb, exch);
          }
          group_sync(g);
          if (PARK && p == 1) {   // accumulators were parked in TMEM during poly b's passes A/B
            unpark8(taddr + 32, racc[0]);
            unpark8(taddr + 64, racc[1]);
          }
          BR_MAC_DIGITS(ND0)
          if (PARK && p == 0 && ND1 == 0) {
            park8(taddr + 32, racc[0]);
            park8(taddr + 64, racc[1]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          group_sync(g);
          if constexpr (ND1 > 0) {
            {
              BR_GET_TA(ta)
              fwd_pass_a<BGBIT, ND0, ND1, MAGIC>(tid, t_re, t_im, ta, exch);
            }
            group_sync(g);
            {
              BR_GET_TB(tb)
              fwd_pass_b<ND1>(tid, tb, exch);
            }
            group_sync(g);
            BR_MAC_DIGITS(ND1)
            group_sync(g);
          }
        }
        {
          BR_GET_TB(tb)
          inv_pass_c(tid, tb, racc, exch);
        }
        group_sync(g);
        inv_pass_b(tid, exch);
        group_sync(g);
        {
          BR_GET_TA(ta)
          inv_pass_a<EXACT, MAGIC>(tid, ta, exch, acc);
        }
        group_sync(g);
      } else {
        // idle group: keep the ring's phase accounting in lock step
        for (int c = 0; c < L2; c++) {
          mbar_wait(&full[stage], parity);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
      }
    }

    if (active) {
      // ---- epilogue: TRLWE, or fused sample extraction (trlwe.rs:106-136)
      if (args.out_mode == BR_OUT_TRLWE) {
        uint32_t *o = args.out + ct * 2 * kN;
        for (int x = tid; x < 2 * kN; x += 64) o[x] = acc[x];
      } else {
        const uint32_t m = args.out_mode == BR_OUT_EXTRACT ? (uint32_t)kN : n;
        uint32_t *o = args.out + ct * (m + 1);
        for (uint32_t x = tid; x <= m; x += 64) {
          uint32_t v;
          if (x == 0) v = acc[0];
          else if (x == m) v = acc[kN];
          else v = ~acc[m - x];
          o[x] = v;
        }
      }
    }
    group_sync(g);
  }
#undef BR_GET_TA
#undef BR_GET_TB
#undef BR_MAC_DIGITS
  if constexpr (TMEM_TW) {
    asm volatile("bar.sync 15, %0;" ::"n"(G * 64) : "memory");  // all consumers done with TMEM
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_base_s));
  }
}

// completion of every outstanding tcgen05.ld of this thread; the loaded registers are threaded
// through as in/out operands so no use can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// ---- intra-warp FFT exchange through tensor memory ---------------------------------------------
// The pass B -> C exchange (and C' -> B' in the inverse) swaps the 3 register-index bits with 3
// lane bits of the SAME warp.  Instead of a padded shared-memory transpose + group barrier it runs as
//   tcgen05.st.32x32b.x32  (thread = TMEM lane, register = column)
//   tcgen05.ld.16x256b.x4  (twice: lanes 0-15 / 16-31)     -> swaps lane bits (4,3) with two
//                                                              register bits, rotates lanes 2..0 up
//   one shfl.xor 16 stage on half of the values             -> swaps the third bit
// (inverse: shuffle stage, tcgen05.st.16x256b.x4 twice, tcgen05.ld.32x32b.x32).  Measured on B200
// (tools/probe/tmem_xchg_probe.cu): the pair moves 8 KB per warp in ~29 cycles of throughput and
// does not touch the shared-memory pipe, which is the kernel's tightest resource; it needs no
// barrier because both directions stay inside one warp.  Which half a lane keeps and which it trades
// depends on a lane bit; instead of selecting registers (48 selects per exchange) the choice is
// absorbed into signs: dft8s (br_core.cuh) delivers its outputs with slots s and s^4 swapped for
// sg = -1, an input swapped that way yields odd outputs negated, and those signs ride on the
// permuted key (forward) and on the inverse pass-A twiddles (inverse).  tools/model/xchg_model.py
// is the numpy model of this index/sign algebra.  Thread maps of a group (W = warp in group, l = lane):
//   pass A / A' : T = j0 + 8 j1                              (unchanged; rows of the exchange buffer
//                                                             are kS = 73 apart so pass B is bank-clean)
//   pass B / B' : W = k0[2], l = (j0[1] j0[0] j0[2] k0[1] k0[0]),  registers j1 -> slot s = k1 ^ 4 l[2]
//   pass C / C' : W = k0[2], l = (k1[2] k0[1] k0[0] k1[1] k1[0]),  registers: slot p = j0 ^ 4 l[4] -> k2
// so bin k0 + 8 k1 + 64 k2 sits in thread T = 32 W + l, register k2, times (-1)^(k2 l[4]): the key is
// permuted and signed to that order once at upload (bsk_permute_kernel, BrArgs::bsk2).
constexpr int kS = 73;                    // exchange-buffer row pitch (complex) of this kernel
constexpr int kXStride = 8 * kS;          // 584 complex per buffer

__device__ __forceinline__ void tmem_st16x256_x4(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16x256_x4(uint32_t taddr, uint32_t (&a)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
        "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ uint32_t cword(const cplx &c, int im, int hi) {
  const double d = im ? c.y : c.x;
  return (uint32_t)(hi ? __double2hiint(d) : __double2loint(d));
}
__device__ __forceinline__ cplx shfl16(const cplx &c) {
  return mk(__shfl_xor_sync(0xffffffffu, c.x, 16), __shfl_xor_sync(0xffffffffu, c.y, 16));
}
// per-thread table of 8 complex constants parked in 32 TMEM columns
__device__ __forceinline__ void tmem_park8(uint32_t taddr, const cplx (&t)[8]) { tmem_st_ta(taddr, t); }
// 32x32b column of (slot s, im, hi): 16 s[2] + 8 im + 4 s[1] + 2 s[0] + hi
// forward: slots s of pass-B threads -> slots p of pass-C threads
// (one 32-column store: four 8-column stores measured 3 % slower, tools/exp_variants.py)
__device__ __forceinline__ void xchg_fwd(uint32_t tq, cplx (&v)[8]) {
  {
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 32; c++) {
      const int sl = ((c >> 4) & 1) * 4 + ((c >> 2) & 1) * 2 + ((c >> 1) & 1);
      r[c] = cword(v[sl], (c >> 3) & 1, c & 1);
    }
    tmem_st32(tq, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t a[2][16];
  tmem_ld16x256_x4(tq, a[0]);
  tmem_ld16x256_x4(tq + (16u << 16), a[1]);
  tmem_wait_ld16(a[0]);
  tmem_wait_ld16(a[1]);
  // register 4*(2z+im) + 2h + hi of half H: z = 0 is the half this lane keeps, z = 1 the half it trades
#pragma unroll
  for (int H = 0; H < 2; H++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      v[2 * H + h] = mk(__hiloint2double((int)a[H][2 * h + 1], (int)a[H][2 * h]),
                        __hiloint2double((int)a[H][4 + 2 * h + 1], (int)a[H][4 + 2 * h]));
      v[4 + 2 * H + h] = shfl16(mk(__hiloint2double((int)a[H][8 + 2 * h + 1], (int)a[H][8 + 2 * h]),
                                   __hiloint2double((int)a[H][12 + 2 * h + 1], (int)a[H][12 + 2 * h])));
    }
}
// inverse: slots p of pass-C threads -> slots s of pass-B threads
__device__ __forceinline__ void xchg_inv(uint32_t tq, cplx (&u)[8]) {
#pragma unroll
  for (int H = 0; H < 2; H++) {
    uint32_t r[16];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const cplx z0 = u[2 * H + h], z1 = shfl16(u[4 + 2 * H + h]);
      r[2 * h] = cword(z0, 0, 0); r[2 * h + 1] = cword(z0, 0, 1);
      r[4 + 2 * h] = cword(z0, 1, 0); r[4 + 2 * h + 1] = cword(z0, 1, 1);
      r[8 + 2 * h] = cword(z1, 0, 0); r[8 + 2 * h + 1] = cword(z1, 0, 1);
      r[12 + 2 * h] = cword(z1, 1, 0); r[12 + 2 * h + 1] = cword(z1, 1, 1);
    }
    tmem_st16x256_x4(tq + ((uint32_t)(16 * H) << 16), r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t r[32];
  tmem_ld32(tq, r);
  tmem_wait_ld32(r);
#pragma unroll
  for (int sl = 0; sl < 8; sl++) {
    const int c = 16 * (sl >> 2) + 4 * ((sl >> 1) & 1) + 2 * (sl & 1);
    u[sl] = mk(__hiloint2double((int)r[c + 1], (int)r[c]), __hiloint2double((int)r[c + 9], (int)r[c + 8]));
  }
}

template <int L, int BGBIT, int STAGES, bool MAGIC>
__global__ void __launch_bounds__(384, 1) blind_rotate_kernel_x(const BrArgs args) {
  constexpr int G = 4, PW = 8, L2 = 2 * L;
  constexpr int kTmemCols = 256;   // per warp: 4 parked tables + 3 exchange blocks of 32 columns
  constexpr int kXBlk = 32;        // distance between the digits' exchange blocks
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  constexpr int kAccBytes = 2 * kN * 4, kExchBytes = 3 * kXStride * 16, kAbarBytes = 2432;
  constexpr int kGroupBytes = kAccBytes + kExchBytes + kAbarBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint8_t *groups = smem + STAGES * kStageBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(groups + G * kGroupBytes);
  uint64_t *empty = full + STAGES;
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(empty + STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n;
  const uint32_t grid = gridDim.x;
  const uint32_t rounds = (uint32_t)((args.count + grid * G - 1) / (grid * G));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2 * G); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = *tmem_base_s;
  if (warp >= PW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == PW && lane == 0) {
      // ===== producer: stream key rows (i, r), permuted layout =====
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk2);
      uint32_t stage = 0, parity = 0;
      const uint32_t rows = n * L2;
      for (uint32_t rd = 0; rd < rounds; rd++)
        for (uint32_t row = 0; row < rows; row++) {
          mbar_wait_backoff(&empty[stage], parity ^ 1);
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tma_load_1d(reinterpret_cast<uint8_t *>(ring) + stage * kStageBytes, src0 + (size_t)row * kStageBytes,
                      kStageBytes, &full[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
    }
    return;
  }
  // ===== consumers: group g owns one ciphertext per round =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g = warp >> 1;
  const int tid = threadIdx.x & 63;               // pass A / A' thread, and the key-row slot of pass C
  const int l4 = (lane >> 4) & 1, l3 = (lane >> 3) & 1, l2 = (lane >> 2) & 1;
  const int b_j0 = 4 * l2 + 2 * l4 + l3, b_k0 = 4 * (warp & 1) + (lane & 3);   // pass B / B' coordinates
  const int c_k1 = 4 * l4 + (lane & 3);                                       // pass C / C' coordinate
  const double sg_b = l2 ? -1.0 : 1.0, sg_c = l4 ? -1.0 : 1.0;
  uint8_t *gbase = groups + g * kGroupBytes;
  uint32_t *acc = reinterpret_cast<uint32_t *>(gbase);
  cplx *exch = reinterpret_cast<cplx *>(gbase + kAccBytes);
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(gbase + kAccBytes + kExchBytes);
  // TMEM columns of this warp: four parked twiddle tables of 32 columns (pass A, signed pass A',
  // pass B by slot, pass C' by slot), then three exchange blocks
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * kTmemCols;
  const uint32_t t_tai = taddr + 32, t_tbf = taddr + 64, t_tbi = taddr + 96, tq = taddr + 128;
  {
    cplx t[8];
#pragma unroll
    for (int k = 0; k < 8; k++) t[k] = args.tw_a[tid * 8 + k];
    tmem_park8(taddr, t);
    // pass B' hands pass A' its inputs negated where j1 is odd and j0[2] is set (see dft8s)
    const double sg_a = (((tid >> 3) & 1) && ((tid >> 2) & 1)) ? -1.0 : 1.0;
#pragma unroll
    for (int k = 0; k < 8; k++) t[k] = mk(t[k].x * sg_a, t[k].y * sg_a);
    tmem_park8(t_tai, t);
#pragma unroll
    for (int sl = 0; sl < 8; sl++) t[sl] = args.tw_b[b_j0 * 8 + (sl ^ (4 * l2))];   // w64^(j0 k1), k1 = s ^ 4 l2
    tmem_park8(t_tbf, t);
#pragma unroll
    for (int sl = 0; sl < 8; sl++) t[sl] = args.tw_b[c_k1 * 8 + (sl ^ (4 * l4))];   // w64^(j0 k1), j0 = p ^ 4 l4
    tmem_park8(t_tbi, t);
  }
  const uint32_t w = n + 1;
  uint32_t stage = 0, parity = 0;
  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = ((size_t)rd * G + g) * grid + blockIdx.x;
    const bool active = ct < args.count;
    if (active) {
      // ---- K0: linear pre-combination + modulus switch (trgsw.rs:202-203, 210-211)
      uint32_t ca = 1, cb = 0, off = 0;
      const uint32_t *A, *B;
      if (args.op >= 0 || args.ops) {
        int op = args.ops ? (int)args.ops[ct] : args.op;
        ca = (uint32_t)c_gate_ca[op]; cb = (uint32_t)c_gate_cb[op]; off = c_gate_off[op];
        A = args.in + ct * 2 * w;
        B = A + w;
      } else {
        A = args.in + ct * w;
        B = A;
      }
      for (uint32_t i = tid; i < n; i += 64) {
        uint32_t v = ca * A[i] + cb * B[i];
        abar_s[i] = (uint16_t)((uint32_t)(v + (1u << 20)) >> 21);
      }
      uint32_t bw = ca * A[n] + cb * B[n] + off;
      uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)bw + (1u << 20)) >> 21));
      const int tvi = args.tv_index ? args.tv_index[ct] : args.tv_default;
      const uint32_t *tv = args.tv + (size_t)tvi * 2 * kN;
      for (int x = tid; x < 2 * kN; x += 64)
        acc[x] = rot_coeff(tv + (x & ~(kN - 1)), x & (kN - 1), b_tilda);
    }
    group_sync(g);
    for (uint32_t i = 0; i < n; i++) {
      if (active) {
        cplx racc[2][8];
        const uint32_t abar = abar_s[i];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
          for (int k = 0; k < 8; k++) racc[o][k] = mk(0.0, 0.0);
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
          {
            uint32_t t_re[8], t_im[8];
            load_t(tid, acc + p * kN, abar, args.offset, t_re, t_im);
            cplx ta[8];
            tmem_ld_ta(taddr, ta);
            fwd_pass_a<BGBIT, 0, L, MAGIC, kS, kXStride>(tid, t_re, t_im, ta, exch);
          }
          group_sync(g);
          cplx tb[8];
          tmem_ld_ta(t_tbf, tb);
#pragma unroll
          for (int d = 0; d < L; d++) {
            // pass B over j1, exchange through TMEM, pass C over j0, MAC against key row (p, d)
            const cplx *e = exch + d * kXStride + b_k0 * kS + b_j0;
            cplx v[8];
#pragma unroll
            for (int j1 = 0; j1 < 8; j1++) v[j1] = e[j1 * 9];
            dft8s<false>(v, sg_b);
#pragma unroll
            for (int sl = 0; sl < 8; sl++) v[sl] = cmul(v[sl], tb[sl]);
            xchg_fwd(tq + kXBlk * d, v);
            dft8<false>(v);
            mbar_wait(&full[stage], parity);
            const cplx *bsk_row = ring + stage * kChunkCplx;
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) {
              cfma(racc[0][k2], v[k2], bsk_row[(k2 * 2 + 0) * 64 + tid]);
              cfma(racc[1][k2], v[k2], bsk_row[(k2 * 2 + 1) * 64 + tid]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
          }
          group_sync(g);   // everyone has read this polynomial's pass-A output
        }
        {
          cplx tbi[8];
          tmem_ld_ta(t_tbi, tbi);
#pragma unroll
          for (int o = 0; o < 2; o++) {
            // pass C' over k2, exchange back, pass B' over k1
            dft8s<true>(racc[o], sg_c);
#pragma unroll
            for (int sl = 0; sl < 8; sl++) racc[o][sl] = cmulc(racc[o][sl], tbi[sl]);
            xchg_inv(tq + kXBlk * o, racc[o]);
            dft8<true>(racc[o]);
            cplx *e = exch + o * kXStride + b_k0 * kS + b_j0;
#pragma unroll
            for (int j1 = 0; j1 < 8; j1++) e[j1 * 9] = racc[o][j1];
          }
        }
        group_sync(g);
        {
          cplx ta[8];
          tmem_ld_ta(t_tai, ta);
          inv_pass_a<EXACT, MAGIC, kS, kXStride>(tid, ta, exch, acc);
        }
        group_sync(g);
      } else {
        // idle group: keep the ring's phase accounting in lock step
        for (int c = 0; c < L2; c++) {
          mbar_wait(&full[stage], parity);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
      }
    }
    if (active) {
      // ---- epilogue: TRLWE, or fused sample extraction (trlwe.rs:106-136)
      if (args.out_mode == BR_OUT_TRLWE) {
        uint32_t *o = args.out + ct * 2 * kN;
        for (int x = tid; x < 2 * kN; x += 64) o[x] = acc[x];
      } else {
        const uint32_t m = args.out_mode == BR_OUT_EXTRACT ? (uint32_t)kN : n;
        uint32_t *o = args.out + ct * (m + 1);
        for (uint32_t x = tid; x <= m; x += 64) {
          uint32_t v;
          if (x == 0) v = acc[0];
          else if (x == m) v = acc[kN];
          else v = ~acc[m - x];
          o[x] = v;
        }
      }
    }
    group_sync(g);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("bar.sync 15, %0;" ::"n"(G * 64) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}
template <int L, int BGBIT, bool MAGIC_REQ>
cudaError_t launch_x(const BrArgs &args, int num_sms, cudaStream_t stream) {
  constexpr bool MAGIC = MAGIC_REQ && L == 3 && BGBIT == 6;
  constexpr int STAGES = 4, G = 4;
  if (!args.bsk2) return cudaErrorInvalidValue;   // engine did not build the permuted key
  auto kern = blind_rotate_kernel_x<L, BGBIT, STAGES, MAGIC>;
  const int smem = STAGES * kStageBytes + G * (2 * kN * 4 + 3 * kXStride * 16 + 2432) + 2 * STAGES * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = (int)(args.count < (size_t)num_sms ? args.count : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, 384, smem, stream>>>(args);
  return cudaGetLastError();
}

// ---- V4: six groups per SM ------------------------------------------------------------
// The 2l digit polynomials of a step (rows 0..2l-1 of BSK[i]) go through the two exchange
// buffers in PAIRS (a0,a1 | a2,b0 | b1,b2 at l=3), so a step has 3l+3 group barriers.  Between
// MAC phases the 2x8 complex accumulators are parked in TMEM next to the pass-A twiddles, so
// passes A and B run with ~100 live registers and the whole kernel fits 160 registers/thread
// without spills -- which is what lets 12 consumer warps (3 per sub-partition) stay resident.
template <int L, int BGBIT>
__global__ void __launch_bounds__(6 * 64 + 128, 1) blind_rotate_kernel_v4(const BrArgs args) {
  constexpr int G = 6, STAGES = 3, NBUF = 2;
  using C = Cfg<L, NBUF>;
  constexpr int L2 = 2 * L;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint8_t *groups = smem + STAGES * kStageBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(groups + G * C::kGroupBytes);
  uint64_t *empty = full + STAGES;
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(empty + STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n;
  const uint32_t grid = gridDim.x;
  const uint32_t per_round = grid * G;
  const uint32_t rounds = (uint32_t)((args.count + per_round - 1) / per_round);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2 * G);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  if (warp >= 2 * G) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == 2 * G && lane == 0) {
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk);
      uint32_t stage = 0, parity = 0;
      const uint32_t rows = n * L2;
      for (uint32_t rd = 0; rd < rounds; rd++)
        for (uint32_t row = 0; row < rows; row++) {
          mbar_wait_backoff(&empty[stage], parity ^ 1);
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tma_load_1d(reinterpret_cast<uint8_t *>(ring) + stage * kStageBytes,
                      src0 + (size_t)row * kStageBytes, kStageBytes, &full[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
    }
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
  const int g = warp >> 1;
  const int tid = threadIdx.x & 63;
  uint8_t *gbase = groups + g * C::kGroupBytes;
  uint32_t *acc = reinterpret_cast<uint32_t *>(gbase);
  cplx *exch = reinterpret_cast<cplx *>(gbase + C::kAccBytes);
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(gbase + C::kAccBytes + C::kExchBytes);

  // TMEM columns of this thread: [0,32) pass-A twiddles, [32,64) racc[0], [64,96) racc[1]
  const uint32_t taddr = *tmem_base_s + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 96u;
  cplx tb1, tb2, tb4;
  {
    const cplx *twb = args.tw_b + (tid & 7) * 8;
    tb1 = twb[1]; tb2 = twb[2]; tb4 = twb[4];
    cplx ta[8];
#pragma unroll
    for (int k = 0; k < 8; k++) ta[k] = args.tw_a[tid * 8 + k];
    tmem_st_ta(taddr, ta);
  }

  const uint32_t w = n + 1;
  uint32_t stage = 0, parity = 0;

  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = ((size_t)rd * G + g) * grid + blockIdx.x;
    const bool active = ct < args.count;

    if (active) {
      uint32_t ca = 1, cb = 0, off = 0;
      const uint32_t *A, *B;
      if (args.op >= 0 || args.ops) {
        int op = args.ops ? (int)args.ops[ct] : args.op;
        ca = (uint32_t)c_gate_ca[op]; cb = (uint32_t)c_gate_cb[op]; off = c_gate_off[op];
        A = args.in + ct * 2 * w;
        B = A + w;
      } else {
        A = args.in + ct * w;
        B = A;
      }
      for (uint32_t i = tid; i < n; i += 64) {
        uint32_t v = ca * A[i] + cb * B[i];
        abar_s[i] = (uint16_t)((uint32_t)(v + (1u << 20)) >> 21);
      }
      uint32_t bw = ca * A[n] + cb * B[n] + off;
      uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)bw + (1u << 20)) >> 21));
      const int tvi = args.tv_index ? args.tv_index[ct] : args.tv_default;
      const uint32_t *tv = args.tv + (size_t)tvi * 2 * kN;
      for (int x = tid; x < 2 * kN; x += 64)
        acc[x] = rot_coeff(tv + (x & ~(kN - 1)), x & (kN - 1), b_tilda);
    }
    group_sync(g);

    for (uint32_t i = 0; i < n; i++) {
      if (active) {
        const uint32_t abar = abar_s[i];
        // sub-round s: rows 2s, 2s+1 of BSK[i]; row r = digit (r % L) of polynomial (r / L).
        // The loop stays rolled (one copy of the passes in the instruction cache).
        cplx racc[2][8];
#pragma unroll 1
        for (int s = 0; s < L; s++) {
          {
            cplx ta[8];
            tmem_ld_ta(taddr, ta);
            uint32_t t_re[8], t_im[8];
            const int p0 = (2 * s) / L, p1 = (2 * s + 1) / L;
            load_t(tid, acc + p0 * kN, abar, args.offset, t_re, t_im);
            fwd_pass_a_rt<BGBIT>(tid, t_re, t_im, ta, exch, (2 * s) % L);
            if (p1 != p0) load_t(tid, acc + p1 * kN, abar, args.offset, t_re, t_im);
            fwd_pass_a_rt<BGBIT>(tid, t_re, t_im, ta, exch + kExchStride, (2 * s + 1) % L);
          }
          group_sync(g);
          {
            cplx tb[8];
            expand_tb(tb1, tb2, tb4, tb);
            fwd_pass_b<2>(tid, tb, exch);
          }
          group_sync(g);
          if (s == 0) {
#pragma unroll
            for (int o = 0; o < 2; o++)
#pragma unroll
              for (int k = 0; k < 8; k++) racc[o][k] = mk(0.0, 0.0);
          } else {
            unpark8(taddr + 32, racc[0]);
            unpark8(taddr + 64, racc[1]);
          }
#pragma unroll
          for (int q = 0; q < 2; q++) {
            mbar_wait(&full[stage], parity);
            fwd_pass_c_mac(tid, exch + q * kExchStride, ring + stage * kChunkCplx, racc);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
          }
          if (s < L - 1) {
            park8(taddr + 32, racc[0]);
            park8(taddr + 64, racc[1]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
          group_sync(g);  // everyone is done reading exch
        }
        {
          cplx tb[8];
          expand_tb(tb1, tb2, tb4, tb);
          inv_pass_c(tid, tb, racc, exch);
        }
        group_sync(g);
        inv_pass_b(tid, exch);
        group_sync(g);
        {
          cplx ta[8];
          tmem_ld_ta(taddr, ta);
          inv_pass_a<EXACT>(tid, ta, exch, acc);
        }
        group_sync(g);
      } else {
        for (int c = 0; c < L2; c++) {
          mbar_wait(&full[stage], parity);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
      }
    }

    if (active) {
      if (args.out_mode == BR_OUT_TRLWE) {
        uint32_t *o = args.out + ct * 2 * kN;
        for (int x = tid; x < 2 * kN; x += 64) o[x] = acc[x];
      } else {
        const uint32_t m = args.out_mode == BR_OUT_EXTRACT ? (uint32_t)kN : n;
        uint32_t *o = args.out + ct * (m + 1);
        for (uint32_t x = tid; x <= m; x += 64) {
          uint32_t v;
          if (x == 0) v = acc[0];
          else if (x == m) v = acc[kN];
          else v = ~acc[m - x];
          o[x] = v;
        }
      }
    }
    group_sync(g);
  }
  asm volatile("bar.sync 15, %0;" ::"n"(G * 64) : "memory");
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_base_s));
}

template <int L, int BGBIT>
cudaError_t launch_v4(const BrArgs &args, int num_sms, cudaStream_t stream) {
  constexpr int G = 6, STAGES = 3;
  auto kern = blind_rotate_kernel_v4<L, BGBIT>;
  const int smem = STAGES * kStageBytes + G * Cfg<L, 2>::kGroupBytes + 2 * STAGES * 8 + 16;
  {  // per device and cheap: set on every launch (engines may live on several GPUs)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int grid = (int)(args.count < (size_t)num_sms ? args.count : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, G * 64 + 128, smem, stream>>>(args);
  return cudaGetLastError();
}

// ---- latency variant: ONE ciphertext per CTA, its 2l digit transforms spread over l groups ---
// For small batches (dependent PBS chains, BASELINE config 4) the throughput kernel leaves a
// ciphertext to one 64-thread group that walks the whole chain (11 k cycles per CMUX).  Here group
// g of l takes rows 2g, 2g+1 of BSK[i] (digit pair), the groups' partial spectra are summed
// through shared memory, and groups 0 and 1 run the two inverse transforms in parallel.  The
// ring has one stage per row (stage r <-> row r), each consumed by exactly one group, so the
// next step's rows stream in while the inverse runs.  256 threads (l<=3 groups + producer warp),
// no register rebalancing needed.
template <int L, int BGBIT>
__global__ void __launch_bounds__(256, 1) blind_rotate_latency_kernel(const BrArgs args) {
  constexpr int NG = L;            // cooperating groups
  constexpr int L2 = 2 * L;
  constexpr int STAGES = L2;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  constexpr int kAccBytes = 2 * kN * 4;
  constexpr int kExchBytes = 2 * kExchStride * 16;  // two buffers per group
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint32_t *acc = reinterpret_cast<uint32_t *>(smem + STAGES * kStageBytes);
  uint8_t *exch_base = smem + STAGES * kStageBytes + kAccBytes;
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(exch_base + NG * kExchBytes);
  uint64_t *full = reinterpret_cast<uint64_t *>(exch_base + NG * kExchBytes + 2432);
  uint64_t *empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n;
  const uint32_t grid = gridDim.x;
  const uint32_t rounds = (uint32_t)((args.count + grid - 1) / grid);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 2);   // the two warps of the one group that owns this row
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= 2 * NG) {
    if (warp == 2 * NG && lane == 0) {
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk);
      uint32_t parity = 0;
      for (uint32_t rd = 0; rd < rounds; rd++)
        for (uint32_t i = 0; i < n; i++) {
          for (uint32_t r = 0; r < (uint32_t)L2; r++) {
            mbar_wait_backoff(&empty[r], parity ^ 1);
            mbar_arrive_expect_tx(&full[r], kStageBytes);
            tma_load_1d(reinterpret_cast<uint8_t *>(ring) + r * kStageBytes,
                        src0 + ((size_t)i * L2 + r) * kStageBytes, kStageBytes, &full[r]);
            // a lone ciphertext streams the key cold from HBM: pull the next step's row into L2 now
            if (i + 1 < n)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                               src0 + ((size_t)(i + 1) * L2 + r) * kStageBytes),
                           "n"(kStageBytes)
                           : "memory");
          }
          parity ^= 1;
        }
    }
    return;
  }

  const int g = warp >> 1;
  const int tid = threadIdx.x & 63;
  const int ctid = threadIdx.x;  // 0 .. 64*NG-1 among consumers
  cplx *exch = reinterpret_cast<cplx *>(exch_base + g * kExchBytes);
  cplx ta[8], tb[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { ta[k] = args.tw_a[tid * 8 + k]; tb[k] = args.tw_b[(tid & 7) * 8 + k]; }
  auto cta_sync = [&]() { asm volatile("bar.sync 8, %0;" ::"n"(64 * NG) : "memory"); };

  const uint32_t w = n + 1;
  uint32_t parity = 0;
  const int p0 = (2 * g) / L, p1 = (2 * g + 1) / L;       // polynomials of this group's two rows
  const int d0 = (2 * g) % L, d1 = (2 * g + 1) % L;       // and their digits

  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = (size_t)rd * grid + blockIdx.x;
    const bool active = ct < args.count;
    if (active) {
      uint32_t ca = 1, cb = 0, off = 0;
      const uint32_t *A, *B;
      if (args.op >= 0 || args.ops) {
        int op = args.ops ? (int)args.ops[ct] : args.op;
        ca = (uint32_t)c_gate_ca[op]; cb = (uint32_t)c_gate_cb[op]; off = c_gate_off[op];
        A = args.in + ct * 2 * w;
        B = A + w;
      } else {
        A = args.in + ct * w;
        B = A;
      }
      for (uint32_t i = ctid; i < n; i += 64 * NG) {
        uint32_t v = ca * A[i] + cb * B[i];
        abar_s[i] = (uint16_t)((uint32_t)(v + (1u << 20)) >> 21);
      }
      uint32_t bw = ca * A[n] + cb * B[n] + off;
      uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)bw + (1u << 20)) >> 21));
      const int tvi = args.tv_index ? args.tv_index[ct] : args.tv_default;
      const uint32_t *tv = args.tv + (size_t)tvi * 2 * kN;
      for (int x = ctid; x < 2 * kN; x += 64 * NG)
        acc[x] = rot_coeff(tv + (x & ~(kN - 1)), x & (kN - 1), b_tilda);
    }
    cta_sync();

    for (uint32_t i = 0; i < n; i++) {
      if (active) {
        const uint32_t abar = abar_s[i];
        {
          uint32_t t_re[8], t_im[8];
          load_t(tid, acc + p0 * kN, abar, args.offset, t_re, t_im);
          fwd_pass_a_rt<BGBIT>(tid, t_re, t_im, ta, exch, d0);
          if (p1 != p0) load_t(tid, acc + p1 * kN, abar, args.offset, t_re, t_im);
          fwd_pass_a_rt<BGBIT>(tid, t_re, t_im, ta, exch + kExchStride, d1);
        }
        group_sync(g);
        fwd_pass_b<2>(tid, tb, exch);
        group_sync(g);
        cplx racc[2][8];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
          for (int k = 0; k < 8; k++) racc[o][k] = mk(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int row = 2 * g + q;
          mbar_wait(&full[row], parity);
          fwd_pass_c_mac(tid, exch + q * kExchStride, ring + row * kChunkCplx, racc);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[row]);
        }
        if (NG > 1) {
          group_sync(g);  // the group is done reading its pass-C inputs
          // publish this group's partial spectra: [o][v*9 + k2] in its own two buffers
#pragma unroll
          for (int o = 0; o < 2; o++)
#pragma unroll
            for (int k = 0; k < 8; k++) exch[o * kExchStride + tid * 9 + k] = racc[o][k];
          cta_sync();
          if (g < 2) {
            // group o = g sums the partials of output o and runs its inverse transform
#pragma unroll
            for (int k = 0; k < 8; k++) racc[0][k] = mk(0.0, 0.0);
#pragma unroll
            for (int gg = 0; gg < NG; gg++) {
              const cplx *src = reinterpret_cast<const cplx *>(exch_base + gg * kExchBytes) +
                                g * kExchStride + tid * 9;
#pragma unroll
              for (int k = 0; k < 8; k++) racc[0][k] = cadd(racc[0][k], src[k]);
            }
          }
          cta_sync();  // all partials consumed before anyone overwrites an exchange buffer
          if (g < 2) {
            cplx *e = exch;  // this group's buffer 0
            dft8<true>(racc[0]);
            e[tid * 9] = racc[0][0];
#pragma unroll
            for (int j0 = 1; j0 < 8; j0++) e[tid * 9 + j0] = cmulc(racc[0][j0], tb[j0]);
            group_sync(g);
            {
              const int k0 = tid >> 3, j0 = tid & 7;
              cplx *eb = e + k0 * 72 + j0;
              cplx v[8];
#pragma unroll
              for (int k1 = 0; k1 < 8; k1++) v[k1] = eb[k1 * 9];
              dft8<true>(v);
#pragma unroll
              for (int j1 = 0; j1 < 8; j1++) eb[j1 * 9] = v[j1];
            }
            group_sync(g);
            {
              const cplx *ea = e + tid + (tid >> 3);
              cplx v[8];
#pragma unroll
              for (int k0 = 0; k0 < 8; k0++) v[k0] = cmulc(ea[k0 * 72], ta[k0]);
              dft8<true>(v);
              uint32_t *ap = acc + g * kN;
#define BR_STORE(M)                                                       \
  {                                                                       \
    cplx y = twist_out<M>(v[M]);                                          \
    ap[64 * M + tid] += round_torus<EXACT>(y.x);                          \
    ap[64 * M + tid + kHalf] += round_torus<EXACT>(y.y);                  \
  }
              BR_STORE(0) BR_STORE(1) BR_STORE(2) BR_STORE(3) BR_STORE(4) BR_STORE(5) BR_STORE(6) BR_STORE(7)
#undef BR_STORE
            }
          }
          cta_sync();
        } else {
          group_sync(g);
          inv_pass_c(tid, tb, racc, exch);
          group_sync(g);
          inv_pass_b(tid, exch);
          group_sync(g);
          inv_pass_a<EXACT>(tid, ta, exch, acc);
          group_sync(g);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int row = 2 * g + q;
          mbar_wait(&full[row], parity);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[row]);
        }
      }
      parity ^= 1;
    }

    if (active) {
      if (args.out_mode == BR_OUT_TRLWE) {
        uint32_t *o = args.out + ct * 2 * kN;
        for (int x = ctid; x < 2 * kN; x += 64 * NG) o[x] = acc[x];
      } else {
        const uint32_t m = args.out_mode == BR_OUT_EXTRACT ? (uint32_t)kN : n;
        uint32_t *o = args.out + ct * (m + 1);
        for (uint32_t x = ctid; x <= m; x += 64 * NG) {
          uint32_t v;
          if (x == 0) v = acc[0];
          else if (x == m) v = acc[kN];
          else v = ~acc[m - x];
          o[x] = v;
        }
      }
    }
    cta_sync();
  }
}

template <int L, int BGBIT>
cudaError_t launch_latency(const BrArgs &args, int num_sms, cudaStream_t stream) {
  auto kern = blind_rotate_latency_kernel<L, BGBIT>;
  const int smem = 2 * L * kStageBytes + 2 * kN * 4 + L * 2 * kExchStride * 16 + 2432 + 4 * L * 8 + 16;
  {  // per device and cheap: set on every launch (engines may live on several GPUs)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int grid = (int)(args.count < (size_t)num_sms ? args.count : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, 256, smem, stream>>>(args);
  return cudaGetLastError();
}

template <int L, int BGBIT, int G, int STAGES, int NBUF, bool TMEM_TW, int RC, int RP, bool PARK = false,
          bool MAGIC_REQ = false>
cudaError_t launch_v(const BrArgs &args, int num_sms, cudaStream_t stream) {
  constexpr bool MAGIC = MAGIC_REQ && L == 3 && BGBIT == 6;
  auto kern = blind_rotate_kernel<L, BGBIT, G, STAGES, NBUF, TMEM_TW, RC, RP, PARK, MAGIC>;
  const int smem = STAGES * kStageBytes + G * Cfg<L, NBUF>::kGroupBytes + 2 * STAGES * 8 + 16;
  {  // per device and cheap: set on every launch (engines may live on several GPUs)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int grid = (int)(args.count < (size_t)num_sms ? args.count : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, ((2 * G + 3) / 4) * 128 + 128, smem, stream>>>(args);
  return cudaGetLastError();
}

int br_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TFHE_BR_VARIANT");
    v = e ? atoi(e) : BR_DEFAULT_VARIANT;
    if (v < 1 || v > 9) v = BR_DEFAULT_VARIANT;
  }
  return v;
}

// batches this small go to the latency kernel (one ciphertext per SM, l groups each)
int br_latency_threshold(int num_sms) {
  static int thr = -2;
  if (thr == -2) {
    const char *e = getenv("TFHE_BR_LATENCY_MAX");
    thr = e ? atoi(e) : -1;
  }
  return thr >= 0 ? thr : num_sms;  // measured crossover: 148 gates 2.55 vs 4.04 ms, 296 gates 5.04 vs 4.41 ms
}

// ciphertexts [base, base + n) of a launch
BrArgs br_slice(const BrArgs &a, size_t base, size_t n) {
  BrArgs s = a;
  const size_t w = a.n + 1;
  const bool gate = a.op >= 0 || a.ops;
  s.in = a.in + base * (gate ? 2 * w : w);
  if (a.ops) s.ops = a.ops + base;
  if (a.tv_index) s.tv_index = a.tv_index + base;
  const size_t out_words = a.out_mode == BR_OUT_TRLWE ? 2 * (size_t)kN
                           : a.out_mode == BR_OUT_EXTRACT ? (size_t)kN + 1 : w;
  s.out = a.out + base * out_words;
  s.count = n;
  return s;
}

template <int L, int BGBIT>
cudaError_t launch_t(const BrArgs &args, int num_sms, cudaStream_t stream) {
  if (L > 1 && args.count <= (size_t)br_latency_threshold(num_sms))
    return launch_latency<L, BGBIT>(args, num_sms, stream);
  // 9: 128 threads per ciphertext, all-FMA radix 4/8/4/4 passes (blind_rotate_s.cu), every round
  if (br_variant() == 9) return br_launch_s(L, BGBIT, args, num_sms, stream);
  // 8 (default): full persistent rounds (4 ciphertexts per SM) on the 64-thread kernel -- pass B<->C
  // exchange through tensor memory + one shuffle stage, permuted key layout -- and a last round that
  // fills at most 3 of the 4 slots per SM on the 128-thread kernel, which is faster at partial residency
  // (B200, one round of 148/296/444/592 ciphertexts: 2.76/3.63/4.65/5.86 ms against 4.34/4.43/5.70/5.67 ms).
  if (br_variant() == 8) {
    const size_t round = (size_t)num_sms * 4;
    const size_t full = args.count / round * round, tail = args.count - full;
    if (tail == 0 || tail > (size_t)num_sms * 3 || !args.bsk3)
      return launch_x<L, BGBIT, true>(args, num_sms, stream);
    if (full) {
      cudaError_t e = launch_x<L, BGBIT, true>(br_slice(args, 0, full), num_sms, stream);
      if (e != cudaSuccess) return e;
    }
    return br_launch_s(L, BGBIT, br_slice(args, full, tail), num_sms, stream);
  }
  // 7: variant 3 with the 2^52-bias conversions
  if (br_variant() == 7) return launch_v<L, BGBIT, 4, 4, 3, true, 232, 40, false, true>(args, num_sms, stream);
  if (br_variant() == 4) return launch_v4<L, BGBIT>(args, num_sms, stream);
  if (br_variant() == 6)  // five groups per SM at 160 registers (3 exchange buffers, 2-stage ring)
    return launch_v<L, BGBIT, 5, 2, 3, true, 160, 24, true>(args, num_sms, stream);
  if (br_variant() == 5)  // variant 3 + MAC accumulators parked in TMEM across poly b's passes A/B
    return launch_v<L, BGBIT, 4, 4, 3, true, 232, 40, true>(args, num_sms, stream);
  if (br_variant() == 2)
    return launch_v<L, BGBIT, 6, 3, 2, true, 160, 24>(args, num_sms, stream);
  if (br_variant() == 3)  // V1 residency, but twiddles out of the register file (more ILP room)
    return launch_v<L, BGBIT, 4, 4, 3, true, 232, 40>(args, num_sms, stream);
  return launch_v<L, BGBIT, 4, 4, 3, false, 232, 40>(args, num_sms, stream);
}

}  // namespace

bool br_uses_permuted_key() { return br_variant() == 8; }
bool br_uses_s_key() { return br_variant() == 9 || br_variant() == 8; }

bool br_supported(uint32_t l, uint32_t bgbit) {
  return (l == 3 && bgbit == 6) || (l == 2 && bgbit == 10) || (l == 1 && bgbit == 18) ||
         (l == 1 && bgbit == 22) || (l == 1 && bgbit == 23);
}

cudaError_t br_launch(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms,
                      cudaStream_t stream, int *launched) {
  if (launched) {   // kernels this call puts on the stream (the default shape may split off a tail launch)
    const size_t round = (size_t)num_sms * 4, tail = args.count % round;
    const bool split = br_variant() == 8 && args.count > round && tail != 0 && tail <= (size_t)num_sms * 3 &&
                       args.bsk3 && !(l > 1 && args.count <= (size_t)br_latency_threshold(num_sms));
    *launched = args.count == 0 ? 0 : split ? 2 : 1;
  }
  if (args.count == 0) return cudaSuccess;
  if (l == 3 && bgbit == 6) return launch_t<3, 6>(args, num_sms, stream);
  if (l == 2 && bgbit == 10) return launch_t<2, 10>(args, num_sms, stream);
  if (l == 1 && bgbit == 18) return launch_t<1, 18>(args, num_sms, stream);
  if (l == 1 && bgbit == 22) return launch_t<1, 22>(args, num_sms, stream);
  if (l == 1 && bgbit == 23) return launch_t<1, 23>(args, num_sms, stream);
  return cudaErrorInvalidValue;
}
