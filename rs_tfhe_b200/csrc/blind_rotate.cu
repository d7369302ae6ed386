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

#include "br_core.cuh"
#include "kernels.h"

using namespace br;

#ifndef BR_PRODUCER_SLEEP_NS
#define BR_PRODUCER_SLEEP_NS 256
#endif
#ifndef BR_STAGGER_NS
#define BR_STAGGER_NS 0
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

template <int L> struct Cfg {
  static constexpr int NBUF = L > 2 ? L : 2;
  static constexpr int kAccBytes = 2 * kN * 4;
  static constexpr int kExchBytes = NBUF * kExchStride * 16;
  static constexpr int kAbarBytes = 2432;  // u16[n], n <= 1216
  static constexpr int kGroupBytes = kAccBytes + kExchBytes + kAbarBytes;
};

template <int L, int BGBIT, int G, int STAGES>
__global__ void __launch_bounds__(G * 64 + 128, 1) blind_rotate_kernel(const BrArgs args) {
  using C = Cfg<L>;
  constexpr int L2 = 2 * L;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint8_t *groups = smem + STAGES * kStageBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(groups + G * C::kGroupBytes);
  uint64_t *empty = full + STAGES;

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
  __syncthreads();

  // Register budget: the SM sub-partition hosting the producer warpgroup also hosts
  // consumer warps, so the launch is compiled at 168 registers/thread and rebalanced
  // here: the producer warpgroup shrinks to 40, the consumer warpgroups grow to 232.
  if (warp >= 2 * G) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // ===== producer: stream BSK rows (i, r) for every round =====
    if (warp == 2 * G && lane == 0) {
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
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g = warp >> 1;
  const int tid = threadIdx.x & 63;
  uint8_t *gbase = groups + g * C::kGroupBytes;
  uint32_t *acc = reinterpret_cast<uint32_t *>(gbase);
  cplx *exch = reinterpret_cast<cplx *>(gbase + C::kAccBytes);
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(gbase + C::kAccBytes + C::kExchBytes);

  Twiddles tw;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    tw.ta[k] = args.tw_a[tid * 8 + k];
    tw.tb[k] = args.tw_b[(tid & 7) * 8 + k];
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
    if (BR_STAGGER_NS > 0 && rd == 0) __nanosleep(g * BR_STAGGER_NS);

    for (uint32_t i = 0; i < n; i++) {
      cplx racc[2][8];
      if (active) {
        const uint32_t abar = abar_s[i];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
          for (int k = 0; k < 8; k++) racc[o][k] = mk(0.0, 0.0);
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
          fwd_pass_a<L, BGBIT>(tid, acc + p * kN, abar, args.offset, tw, exch);
          group_sync(g);
          fwd_pass_b<L>(tid, tw, exch);
          group_sync(g);
#pragma unroll
          for (int d = 0; d < L; d++) {
            mbar_wait(&full[stage], parity);
            fwd_pass_c_mac(tid, exch + d * kExchStride, ring + stage * kChunkCplx, racc);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
          }
          group_sync(g);
        }
        inv_pass_c(tid, tw, racc, exch);
        group_sync(g);
        inv_pass_b(tid, exch);
        group_sync(g);
        inv_pass_a<EXACT>(tid, tw, exch, acc);
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
}

template <int L, int BGBIT>
cudaError_t launch_t(const BrArgs &args, int num_sms, cudaStream_t stream) {
  constexpr int G = 4, STAGES = 4;
  auto kern = blind_rotate_kernel<L, BGBIT, G, STAGES>;
  const int smem = STAGES * kStageBytes + G * Cfg<L>::kGroupBytes + 2 * STAGES * 8;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  size_t groups = (args.count + G - 1) / G;
  int grid = (int)(groups < (size_t)num_sms ? groups : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, G * 64 + 128, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace

bool br_supported(uint32_t l, uint32_t bgbit) {
  return (l == 3 && bgbit == 6) || (l == 2 && bgbit == 10) || (l == 1 && bgbit == 18) ||
         (l == 1 && bgbit == 22) || (l == 1 && bgbit == 23);
}

cudaError_t br_launch(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms,
                      cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  if (l == 3 && bgbit == 6) return launch_t<3, 6>(args, num_sms, stream);
  if (l == 2 && bgbit == 10) return launch_t<2, 10>(args, num_sms, stream);
  if (l == 1 && bgbit == 18) return launch_t<1, 18>(args, num_sms, stream);
  if (l == 1 && bgbit == 22) return launch_t<1, 22>(args, num_sms, stream);
  if (l == 1 && bgbit == 23) return launch_t<1, 23>(args, num_sms, stream);
  return cudaErrorInvalidValue;
}
