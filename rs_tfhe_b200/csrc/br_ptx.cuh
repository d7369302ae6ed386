// br_ptx.cuh -- device helpers shared by the blind-rotation kernels: mbarrier / TMA bulk copy /
// named barriers, and the gate prologue (K0) and extraction epilogue every kernel shape uses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "br_core.cuh"
#include "kernels.h"

namespace brp {

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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// producer-side wait: back off between polls so the spin does not steal issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
template <int THREADS> __device__ __forceinline__ void named_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory");
}

// gates.rs:54-150: out = ca*a + cb*b, b-word += off
static __constant__ int32_t c_gate_ca[TFHE_GATE_COUNT] = {-1, 1, 1, 1, 1, -1, -1, 1, -1, 1};
static __constant__ int32_t c_gate_cb[TFHE_GATE_COUNT] = {-1, 1, 1, 2, -2, -1, 1, -1, 1, -1};
static __constant__ uint32_t c_gate_off[TFHE_GATE_COUNT] = {0x20000000u, 0xE0000000u, 0x20000000u,
                                                     0x40000000u, 0xC0000000u, 0xE0000000u,
                                                     0xE0000000u, 0xE0000000u, 0x20000000u,
                                                     0x20000000u};

// K0 for ciphertext ct by a group of THREADS threads (tid = index within the group):
// linear pre-combination (gates.rs:366-373) + modulus switch (trgsw.rs:202-203, 210-211) into
// abar_s[0..n), and acc = X^b~ * testvec (trgsw.rs:204-207).
template <int THREADS>
__device__ __forceinline__ void prologue(const BrArgs &args, size_t ct, int tid, uint16_t *abar_s,
                                         uint32_t *acc) {
  const uint32_t n = args.n, w = n + 1;
  uint32_t ca = 1, cb = 0, off = 0;
  const uint32_t *A, *B;
  if (args.op >= 0 || args.ops) {
    int op = args.ops ? (int)args.ops[ct] : args.op;
    if ((unsigned)op >= (unsigned)TFHE_GATE_COUNT) op = 0;   // validated on the host; never index out of range
    ca = (uint32_t)c_gate_ca[op]; cb = (uint32_t)c_gate_cb[op]; off = c_gate_off[op];
    A = args.in + ct * 2 * w;
    B = A + w;
  } else {
    A = args.in + ct * w;
    B = A;
  }
  for (uint32_t i = tid; i < n; i += THREADS) {
    uint32_t v = ca * A[i] + cb * B[i];
    abar_s[i] = (uint16_t)((uint32_t)(v + (1u << 20)) >> 21);
  }
  const uint32_t bw = ca * A[n] + cb * B[n] + off;
  const uint32_t b_tilda = (uint32_t)(2 * br::kN - (((uint64_t)bw + (1u << 20)) >> 21));
  const int tvi = args.tv_index ? args.tv_index[ct] : args.tv_default;
  const uint32_t *tv = args.tv + (size_t)tvi * 2 * br::kN;
  for (int x = tid; x < 2 * br::kN; x += THREADS)
    acc[x] = br::rot_coeff(tv + (x & ~(br::kN - 1)), x & (br::kN - 1), b_tilda);
}

// TRLWE out, or fused sample extraction (trlwe.rs:106-136)
template <int THREADS>
__device__ __forceinline__ void epilogue(const BrArgs &args, size_t ct, int tid, const uint32_t *acc) {
  if (args.out_mode == BR_OUT_TRLWE) {
    uint32_t *o = args.out + ct * 2 * br::kN;
    for (int x = tid; x < 2 * br::kN; x += THREADS) o[x] = acc[x];
  } else {
    const uint32_t m = args.out_mode == BR_OUT_EXTRACT ? (uint32_t)br::kN : args.n;
    uint32_t *o = args.out + ct * (m + 1);
    for (uint32_t x = tid; x <= m; x += THREADS) {
      uint32_t v;
      if (x == 0) v = acc[0];
      else if (x == m) v = acc[br::kN];
      else v = ~acc[m - x];
      o[x] = v;
    }
  }
}

}  // namespace brp
