// blind_rotate_s.cu -- K0+K3, throughput shape: persistent blind rotation with 128 threads per
// ciphertext (16 FP64 warps per SM, 4 per sub-partition).
//
// Replaces (reference, file:line under rs-tfhe):
//   gates.rs:366-373 (batch prep), trgsw.rs:198-274 (blind_rotate[_with_testvec]),
//   trgsw.rs:174-196 (cmux), :77-142 (external product), :144-171 (decomposition),
//   :307-330 (X^k), fft/klemsa.rs:88-150 (transforms), trlwe.rs:106-136 (extract).
//
// One persistent CTA per SM owns 4 ciphertexts at a time; ciphertext g is walked through the n CMUX
// steps by warps 4g..4g+3, one on each SM sub-partition, so every sub-partition always holds four
// FP64 warps in four independent phases (the 64-thread shape of round 1 held two and left the FP64
// pipe idle 37 % of the cycles).  Per thread: 4 complex points of the transform in flight and
// 2 x 4 complex MAC accumulators (brs_core.cuh has the index maps and the all-FMA butterflies).
// Data movement of one transform:
//   pass A -> B   shared memory (crosses warps); pass B is the split radix-8: both lanes of a pair
//                 read the same 8 inputs (broadcast) and each produces 4 outputs
//   pass B -> C, C -> D and back   TENSOR MEMORY, inside each warp: tcgen05.st.32x32b.x16 (thread =
//                 TMEM lane, 16 columns = 4 complex) + two tcgen05.ld.16x256b.x2 swap the two register
//                 index bits with lane bits (4,3) and rotate the other lane bits up by two; no barrier,
//                 no shared-memory traffic, no shuffles
//   pass B' -> A' shared memory (padded rows); A' is the split radix-8 of the inverse
// The Fourier-domain key streams through a ring of 16 KB stages filled by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx) issued by a producer warp; a stage is released when all 16
// consumer warps have used it, so each key row crosses L2->SM once per CTA per step.
// Per-thread transform constants (20 complex) live in tensor memory next to the exchange blocks.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "br_ptx.cuh"
#include "brs_core.cuh"
#include "brs_tmem.cuh"
#include "kernels.h"

using namespace br;
using namespace brp;
using namespace brt;

namespace {

constexpr int kStageBytes = brs::kRowCplx * 16;   // 16 KB: one key row
constexpr int kG = 4;                             // ciphertexts resident per CTA
constexpr int kConsumers = kG * brs::kT;          // 512 threads
constexpr int kThreads = kConsumers + 128;        // + a producer warpgroup (one warp per sub-partition)
// Register plan.  A sub-partition's file holds 512 registers per lane and hosts 4 consumer warps + 1 warp of
// the producer warpgroup.  setmaxnreg only moves registers inside the CTA's LAUNCH allocation
// (5 warps x 96 = 480 per lane: the largest multiple-of-8 count that fits), so consumers can grow to
// (480 - 24) / 4 = 114 -> 112.  (Asking for 120 blocks forever: measured, the kernel hangs.)
constexpr int kRegsCons = 112, kRegsProd = 24;

template <int L, int BGBIT, int STAGES, bool MAGIC>
__global__ void __launch_bounds__(kThreads, 1) blind_rotate_kernel_s(const BrArgs args) {
  constexpr int L2 = 2 * L;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  constexpr int kAccBytes = 2 * kN * 4;
  constexpr int kInvBytes = 8 * brs::kInvPitch * 16;                    // one output's inverse exchange buffer
  constexpr int kFwdBytes = kHalf * 16;                                  // one digit's forward exchange buffer
  constexpr int kExchBytes = (L * kFwdBytes > 2 * kInvBytes) ? L * kFwdBytes : 2 * kInvBytes;
  constexpr int kAbarBytes = 2432;
  constexpr int kGroupBytes = kAccBytes + kExchBytes + kAbarBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint8_t *groups = smem + STAGES * kStageBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(groups + kG * kGroupBytes);
  uint64_t *empty = full + STAGES;
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(empty + STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n;
  const uint32_t grid = gridDim.x;
  const uint32_t rounds = (uint32_t)((args.count + (size_t)grid * kG - 1) / ((size_t)grid * kG));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4 * kG); }
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
  if (warp >= 4 * kG) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProd));
    if (warp == 4 * kG && lane == 0) {
      // ===== producer: stream key rows (i, r) in the kernel's thread order =====
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk3);
      uint32_t stage = 0, parity = 0;
      const uint32_t rows = n * L2;
      for (uint32_t rd = 0; rd < rounds; rd++)
        for (uint32_t row = 0; row < rows; row++) {
          mbar_wait_backoff(&empty[stage], parity ^ 1, 128);
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tma_load_1d(reinterpret_cast<uint8_t *>(ring) + stage * kStageBytes, src0 + (size_t)row * kStageBytes,
                      kStageBytes, &full[stage]);
          if (++stage == STAGES) { stage = 0; parity ^= 1; }
        }
    }
    return;
  }
  // ===== consumers: warps 4g..4g+3 own ciphertext g of the round =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCons));
  const int g = warp >> 2;
  const int T = threadIdx.x & (brs::kT - 1);
  uint8_t *gbase = groups + g * kGroupBytes;
  uint32_t *acc = reinterpret_cast<uint32_t *>(gbase);
  cplx *exch = reinterpret_cast<cplx *>(gbase + kAccBytes);
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(gbase + kAccBytes + kExchBytes);
  // TMEM columns of this warp: 5 blocks of 16 (constants), then two exchange blocks
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)g * 128u;
  const uint32_t t_b = taddr, t_cd = taddr + 16, t_cbi = taddr + 32, t_ai = taddr + 48, t_ut = taddr + 64;
  const uint32_t tq0 = taddr + 80, tq1 = taddr + 96;
  {
    const cplx *tw = args.tw_s + (size_t)T * brs::kTwPerThread;
#pragma unroll
    for (int blk = 0; blk < 5; blk++) {
      cplx t[4];
#pragma unroll
      for (int k = 0; k < 4; k++) t[k] = tw[4 * blk + k];
      tm_store4(taddr + 16 * blk, t);
    }
    tm_wait_st();
  }
  uint32_t stage = 0, parity = 0;
  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = ((size_t)rd * kG + g) * grid + blockIdx.x;
    const bool active = ct < args.count;
    if (active) prologue<brs::kT>(args, ct, T, abar_s, acc);
    named_sync<brs::kT>(g + 1);
    for (uint32_t i = 0; i < n; i++) {
      if (active) {
        cplx racc[2][4];
        const uint32_t abar = abar_s[i];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
          for (int k = 0; k < 4; k++) racc[o][k] = mk(0.0, 0.0);
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
          {
            uint32_t t_re[4], t_im[4];
            brs::load_t(T, acc + p * kN, abar, args.offset, t_re, t_im);
#pragma unroll
            for (int d = 0; d < L; d++) brs::fwd_pass_a<BGBIT, MAGIC>(T, d, t_re, t_im, exch + d * kHalf);
          }
          named_sync<brs::kT>(g + 1);
#pragma unroll
          for (int d = 0; d < L; d++) {
            cplx y[4];
            {
              cplx tb[4];
              tm_load4(t_b, tb);
              brs::fwd_pass_b(T, exch + d * kHalf, tb[0], tb[1], tb[2], y);
            }
            xchg_fwd(tq0, y);
            cplx tcd[4];
            tm_load4(t_cd, tcd);
            brs::r4<false>(y, tcd[0], tcd[1]);
            xchg_fwd(tq1, y);
            brs::r4<false>(y, tcd[2], tcd[3]);
            mbar_wait(&full[stage], parity);
            const cplx *row = ring + stage * brs::kRowCplx + T;
#pragma unroll
            for (int kd = 0; kd < 4; kd++) {
              cfma(racc[0][kd], y[kd], row[(kd * 2 + 0) * brs::kT]);
              cfma(racc[1][kd], y[kd], row[(kd * 2 + 1) * brs::kT]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; parity ^= 1; }
          }
          named_sync<brs::kT>(g + 1);   // everyone has read this polynomial's pass-A output
        }
        {
          cplx ti[4];
          tm_load4(t_cbi, ti);
#pragma unroll
          for (int o = 0; o < 2; o++) {
            brs::r4_plain<true>(racc[o]);
            xchg_inv(tq0, racc[o]);
            brs::r4<true>(racc[o], ti[0], ti[1]);
            xchg_inv(tq1, racc[o]);
            brs::r4<true>(racc[o], ti[2], ti[3]);
            brs::inv_store_b(T, racc[o], exch + o * (8 * brs::kInvPitch));
          }
        }
        named_sync<brs::kT>(g + 1);
        {
          cplx ta[4], ut[4];
          tm_load4(t_ai, ta);
          tm_load4(t_ut, ut);
#pragma unroll
          for (int o = 0; o < 2; o++)
            brs::inv_pass_a<EXACT, MAGIC>(T, exch + o * (8 * brs::kInvPitch), ta[0], ta[1], ta[2], ut, acc + o * kN);
        }
        named_sync<brs::kT>(g + 1);
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
    if (active) epilogue<brs::kT>(args, ct, T, acc);
    named_sync<brs::kT>(g + 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("bar.sync 15, %0;" ::"n"(kConsumers) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int L, int BGBIT, int STAGES>
cudaError_t launch_s(const BrArgs &args_in, int num_sms, cudaStream_t stream) {
  constexpr bool MAGIC = (L == 3 && BGBIT == 6);
  if (!args_in.bsk3 || !args_in.tw_s) return cudaErrorInvalidValue;   // engine did not build this kernel's key order
  const BrArgs &args = args_in;
  auto kern = blind_rotate_kernel_s<L, BGBIT, STAGES, MAGIC>;
  constexpr int kInvBytes = 8 * brs::kInvPitch * 16, kFwdBytes = kHalf * 16;
  constexpr int kExchBytes = (L * kFwdBytes > 2 * kInvBytes) ? L * kFwdBytes : 2 * kInvBytes;
  const int smem = STAGES * kStageBytes + kG * (2 * kN * 4 + kExchBytes + 2432) + 2 * STAGES * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = (int)(args.count < (size_t)num_sms ? args.count : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, kThreads, smem, stream>>>(args);
  return cudaGetLastError();
}

// ---- latency shape: ONE ciphertext per CTA, spread over l groups of 128 threads ---------------------
// Dependent PBS chains (BASELINE config 4: examples/lut_add_two_numbers.rs:80-157) and batches of at
// most one ciphertext per SM are bound by the latency of a single blind rotation.  Group g of l takes
// digit g of BOTH accumulator polynomials (key rows g and l + g of BSK[i]): two forward transforms and
// two MACs instead of 2l, the groups' partial spectra meet in shared memory, and groups 0 and 1 run the
// two inverse transforms in parallel.  The ring has one stage per key row, each consumed by exactly
// one group, so the next step's rows stream in during the inverse (and the producer prefetches the
// step after into L2: a lone ciphertext streams the key cold from HBM).  l * 128 + 32 threads,
// 128 registers each without rebalancing (4 warps per sub-partition).
template <int L, int BGBIT, bool MAGIC>
__global__ void __launch_bounds__(L * brs::kT + 32, 1) blind_rotate_latency_s_kernel(const BrArgs args) {
  constexpr int NG = L, L2 = 2 * L, STAGES = L2;
  constexpr int kCons = NG * brs::kT;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  constexpr int kAccBytes = 2 * kN * 4;
  constexpr int kInvBytes = 8 * brs::kInvPitch * 16;
  constexpr int kExchBytes = 2 * kHalf * 16;                      // two forward buffers (>= one inverse buffer)
  constexpr int kPartBytes = 2 * 4 * brs::kT * 16;                // one group's partial spectra [o][kd][T]
  static_assert(kInvBytes <= kExchBytes, "inverse buffer aliases the forward buffers");
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint32_t *acc = reinterpret_cast<uint32_t *>(smem + STAGES * kStageBytes);
  uint8_t *exch_base = smem + STAGES * kStageBytes + kAccBytes;
  cplx *part = reinterpret_cast<cplx *>(exch_base + NG * kExchBytes);
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(exch_base + NG * (kExchBytes + kPartBytes));
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(abar_s) + 2432);
  uint64_t *empty = full + STAGES;
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(empty + STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n, grid = gridDim.x;
  const uint32_t rounds = (uint32_t)((args.count + grid - 1) / grid);
  if (threadIdx.x == 0) {
    for (int r = 0; r < STAGES; r++) { mbar_init(&full[r], 1); mbar_init(&empty[r], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_512(tmem_base_s);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = *tmem_base_s;
  if (warp >= 4 * NG) {
    if (lane == 0) {
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk3);
      uint32_t parity = 0;
      for (uint32_t rd = 0; rd < rounds; rd++)
        for (uint32_t i = 0; i < n; i++) {
          for (uint32_t r = 0; r < (uint32_t)L2; r++) {
            mbar_wait_backoff(&empty[r], parity ^ 1, 64);
            mbar_arrive_expect_tx(&full[r], kStageBytes);
            tma_load_1d(reinterpret_cast<uint8_t *>(ring) + r * kStageBytes,
                        src0 + ((size_t)i * L2 + r) * kStageBytes, kStageBytes, &full[r]);
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
  const int g = warp >> 2;
  const int T = threadIdx.x & (brs::kT - 1);
  const int ctid = threadIdx.x;                      // 0 .. kCons-1 among the consumers
  cplx *exch = reinterpret_cast<cplx *>(exch_base + g * kExchBytes);
  auto cta_sync = [&]() { asm volatile("bar.sync 8, %0;" ::"n"(kCons) : "memory"); };
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)g * 128u;
  const uint32_t t_b = taddr, t_cd = taddr + 16, t_cbi = taddr + 32, t_ai = taddr + 48, t_ut = taddr + 64;
  const uint32_t tq0 = taddr + 80, tq1 = taddr + 96;
  {
    const cplx *tw = args.tw_s + (size_t)T * brs::kTwPerThread;
#pragma unroll
    for (int blk = 0; blk < 5; blk++) {
      cplx t[4];
#pragma unroll
      for (int k = 0; k < 4; k++) t[k] = tw[4 * blk + k];
      tm_store4(taddr + 16 * blk, t);
    }
    tm_wait_st();
  }
  uint32_t parity = 0;
  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = (size_t)rd * grid + blockIdx.x;
    const bool active = ct < args.count;
    if (active) prologue<kCons>(args, ct, ctid, abar_s, acc);
    cta_sync();
    for (uint32_t i = 0; i < n; i++) {
      if (active) {
        const uint32_t abar = abar_s[i];
        cplx racc[2][4];
#pragma unroll
        for (int o = 0; o < 2; o++)
#pragma unroll
          for (int k = 0; k < 4; k++) racc[o][k] = mk(0.0, 0.0);
#pragma unroll
        for (int p = 0; p < 2; p++) {   // digit g of both accumulator polynomials
          uint32_t t_re[4], t_im[4];
          brs::load_t(T, acc + p * kN, abar, args.offset, t_re, t_im);
          brs::fwd_pass_a<BGBIT, MAGIC>(T, g, t_re, t_im, exch + p * kHalf);
        }
        named_sync<brs::kT>(g + 1);
        {
          cplx tb[4], tcd[4];
          tm_load4(t_b, tb);
          tm_load4(t_cd, tcd);
#pragma unroll
          for (int p = 0; p < 2; p++) {
            const int row = p * L + g;
            cplx y[4];
            brs::fwd_pass_b(T, exch + p * kHalf, tb[0], tb[1], tb[2], y);
            xchg_fwd(tq0, y);
            brs::r4<false>(y, tcd[0], tcd[1]);
            xchg_fwd(tq1, y);
            brs::r4<false>(y, tcd[2], tcd[3]);
            mbar_wait(&full[row], parity);
            const cplx *rw = ring + row * brs::kRowCplx + T;
#pragma unroll
            for (int kd = 0; kd < 4; kd++) {
              cfma(racc[0][kd], y[kd], rw[(kd * 2 + 0) * brs::kT]);
              cfma(racc[1][kd], y[kd], rw[(kd * 2 + 1) * brs::kT]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[row]);
          }
        }
        if (NG > 1) {
          // publish this group's partial spectra; group o sums output o and runs its inverse
          cplx *mine = part + (size_t)g * (2 * 4 * brs::kT);
#pragma unroll
          for (int o = 0; o < 2; o++)
#pragma unroll
            for (int kd = 0; kd < 4; kd++) mine[(o * 4 + kd) * brs::kT + T] = racc[o][kd];
          cta_sync();   // partials visible; every group is past its pass-B reads of the exchange buffers
          if (g < 2) {
            cplx s[4];
#pragma unroll
            for (int kd = 0; kd < 4; kd++) s[kd] = g ? racc[1][kd] : racc[0][kd];   // no dynamic register indexing
#pragma unroll
            for (int gg = 0; gg < NG; gg++) {
              if (gg == g) continue;
              const cplx *src = part + (size_t)gg * (2 * 4 * brs::kT) + (size_t)g * 4 * brs::kT + T;
#pragma unroll
              for (int kd = 0; kd < 4; kd++) s[kd] = cadd(s[kd], src[kd * brs::kT]);
            }
            cplx ti[4];
            tm_load4(t_cbi, ti);
            brs::r4_plain<true>(s);
            xchg_inv(tq0, s);
            brs::r4<true>(s, ti[0], ti[1]);
            xchg_inv(tq1, s);
            brs::r4<true>(s, ti[2], ti[3]);
            brs::inv_store_b(T, s, exch);
            named_sync<brs::kT>(g + 1);
            cplx ta[4], ut[4];
            tm_load4(t_ai, ta);
            tm_load4(t_ut, ut);
            brs::inv_pass_a<EXACT, MAGIC>(T, exch, ta[0], ta[1], ta[2], ut, acc + g * kN);
          }
          cta_sync();
        } else {
          named_sync<brs::kT>(g + 1);   // pass-B reads done: the exchange buffers may be rewritten
          cplx ti[4];
          tm_load4(t_cbi, ti);
#pragma unroll
          for (int o = 0; o < 2; o++) {
            brs::r4_plain<true>(racc[o]);
            xchg_inv(tq0, racc[o]);
            brs::r4<true>(racc[o], ti[0], ti[1]);
            xchg_inv(tq1, racc[o]);
            brs::r4<true>(racc[o], ti[2], ti[3]);
            brs::inv_store_b(T, racc[o], exch);
            named_sync<brs::kT>(g + 1);
            cplx ta[4], ut[4];
            tm_load4(t_ai, ta);
            tm_load4(t_ut, ut);
            brs::inv_pass_a<EXACT, MAGIC>(T, exch, ta[0], ta[1], ta[2], ut, acc + o * kN);
            named_sync<brs::kT>(g + 1);
          }
        }
      } else {
#pragma unroll
        for (int p = 0; p < 2; p++) {
          const int row = p * L + g;
          mbar_wait(&full[row], parity);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[row]);
        }
      }
      parity ^= 1;
    }
    if (active) epilogue<kCons>(args, ct, ctid, acc);
    cta_sync();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cta_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) tmem_dealloc_512(tbase);
}

// ---- cluster latency shape: ONE ciphertext per thread-block CLUSTER of two CTAs -----------------------
// For chains of dependent bootstraps a lone blind rotation is bound by what one SM can issue per step
// (measured: every phase of the one-CTA latency shape is issue / FP64-pipe bound).  Two CTAs on two SMs
// split the step: CTA c owns accumulator polynomial c -- it rotates / decomposes / transforms only that
// polynomial (digit g in group g: key row c*l + g), so the forward work per SM halves -- and output c:
// the partial spectra for the OTHER output go straight into the peer's shared memory
// (st.async, distributed shared memory, completion counted on the peer's mbarrier);
// each CTA then sums 2l partials for its own output, runs ONE inverse transform and updates its
// own polynomial, which is all it needs for the next step: no accumulator data ever crosses.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// 16 bytes into the peer's shared memory as an asynchronous distributed-shared-memory write that counts them on
// the PEER's mbarrier when it lands: no cluster-scope fence and no separate remote arrive on the sender's
// critical path (with st.shared::cluster + fence.acq_rel.cluster + a remote arrive the step took 1700 cycles more)
__device__ __forceinline__ void st_async_remote(uint32_t addr, cplx v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr),
               "d"(v.x), "d"(v.y), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int L, int BGBIT, bool MAGIC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(L * brs::kT + 32, 1)
blind_rotate_cluster_kernel(const BrArgs args) {
  constexpr int NG = L, STAGES = L;
  constexpr int kCons = NG * brs::kT;
  constexpr bool EXACT = (L == 3 && BGBIT == 6);
  constexpr int kInvBytes = 8 * brs::kInvPitch * 16;              // >= one forward buffer (8 KB)
  constexpr int kSlotCplx = 4 * brs::kT;                           // one group's partial spectrum of one output
  constexpr int kPartBytes = 2 * L * kSlotCplx * 16;               // 2l slots: l local + l from the peer
  extern __shared__ __align__(128) uint8_t smem[];
  cplx *ring = reinterpret_cast<cplx *>(smem);
  uint32_t *acc = reinterpret_cast<uint32_t *>(smem + STAGES * kStageBytes);        // OUR polynomial only
  uint8_t *exch_base = smem + STAGES * kStageBytes + kN * 4;
  cplx *part = reinterpret_cast<cplx *>(exch_base + NG * kInvBytes);               // [parity 2][2l][4][128]
  // part: [parity 2][2l local slots][4][128], then [parity 2] reduced spectra received from the peer
  uint16_t *abar_s = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(part) + 2 * kPartBytes + 2 * kSlotCplx * 16);
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(abar_s) + 2432);
  uint64_t *empty = full + STAGES;
  uint64_t *xbar = empty + STAGES;                                 // [2]: the peer's reduced spectrum of this parity landed
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(xbar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = args.n;
  const uint32_t rank = cluster_rank(), peer = rank ^ 1u;
  const uint32_t n_clusters = gridDim.x >> 1, cluster = blockIdx.x >> 1;
  const uint32_t rounds = (uint32_t)((args.count + n_clusters - 1) / n_clusters);
  if (threadIdx.x == 0) {
    for (int r = 0; r < STAGES; r++) { mbar_init(&full[r], 1); mbar_init(&empty[r], 4); }
    mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1);   // one local arrive.expect_tx per use; the peer's stores carry the bytes
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_512(tmem_base_s);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  cluster_sync_all();            // the peer's mbarriers exist before anyone arrives on them remotely
  const uint32_t tbase = *tmem_base_s;
  if (warp >= 4 * NG) {
    if (lane == 0) {
      const uint8_t *src0 = reinterpret_cast<const uint8_t *>(args.bsk3);
      uint32_t parity = 0;
      for (uint32_t rd = 0; rd < rounds; rd++)
        for (uint32_t i = 0; i < n; i++) {
          for (uint32_t r = 0; r < (uint32_t)L; r++) {
            const size_t row = (size_t)i * 2 * L + rank * L + r;
            mbar_wait_backoff(&empty[r], parity ^ 1, 64);
            mbar_arrive_expect_tx(&full[r], kStageBytes);
            tma_load_1d(reinterpret_cast<uint8_t *>(ring) + r * kStageBytes, src0 + row * kStageBytes, kStageBytes,
                        &full[r]);
            if (i + 1 < n)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src0 + (row + 2 * L) * kStageBytes),
                           "n"(kStageBytes)
                           : "memory");
          }
          parity ^= 1;
        }
    }
    return;
  }
  const int g = warp >> 2;
  const int T = threadIdx.x & (brs::kT - 1);
  const int ctid = threadIdx.x;
  cplx *exch = reinterpret_cast<cplx *>(exch_base + g * kInvBytes);
  auto cta_sync = [&]() { asm volatile("bar.sync 8, %0;" ::"n"(kCons) : "memory"); };
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)g * 128u;
  const uint32_t t_b = taddr, t_cd = taddr + 16, t_cbi = taddr + 32, t_ai = taddr + 48, t_ut = taddr + 64;
  const uint32_t tq0 = taddr + 80, tq1 = taddr + 96;
  {
    const cplx *tw = args.tw_s + (size_t)T * brs::kTwPerThread;
#pragma unroll
    for (int blk = 0; blk < 5; blk++) {
      cplx t[4];
#pragma unroll
      for (int k = 0; k < 4; k++) t[k] = tw[4 * blk + k];
      tm_store4(taddr + 16 * blk, t);
    }
    tm_wait_st();
  }
  const uint32_t part_remote = map_to_cta(smem_u32(part), peer);
  const uint32_t xbar_remote = map_to_cta(smem_u32(xbar), peer);
  uint32_t parity = 0, step = 0;     // step counts CMUX steps over all rounds: partial buffers alternate by step
  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = (size_t)rd * n_clusters + cluster;
    const bool active = ct < args.count;      // both CTAs of a cluster agree
    if (active) {
      // K0 for our polynomial: abar for every step, acc = (X^b~ * testvec)[rank]
      const uint32_t w = n + 1;
      uint32_t ca = 1, cb = 0, off = 0;
      const uint32_t *A, *B;
      if (args.op >= 0 || args.ops) {
        int op = args.ops ? (int)args.ops[ct] : args.op;
        if ((unsigned)op >= (unsigned)TFHE_GATE_COUNT) op = 0;
        ca = (uint32_t)c_gate_ca[op]; cb = (uint32_t)c_gate_cb[op]; off = c_gate_off[op];
        A = args.in + ct * 2 * w; B = A + w;
      } else {
        A = args.in + ct * w; B = A;
      }
      for (uint32_t x = ctid; x < n; x += kCons) {
        const uint32_t v = ca * A[x] + cb * B[x];
        abar_s[x] = (uint16_t)((uint32_t)(v + (1u << 20)) >> 21);
      }
      const uint32_t bw = ca * A[n] + cb * B[n] + off;
      const uint32_t b_tilda = (uint32_t)(2 * kN - (((uint64_t)bw + (1u << 20)) >> 21));
      const int tvi = args.tv_index ? args.tv_index[ct] : args.tv_default;
      const uint32_t *tv = args.tv + ((size_t)tvi * 2 + rank) * kN;
      for (int x = ctid; x < kN; x += kCons) acc[x] = rot_coeff(tv, x, b_tilda);
    }
    cta_sync();
    for (uint32_t i = 0; i < n; i++, step++) {
      const uint32_t par = step & 1u, xphase = (step >> 1) & 1u;
      if (active) {
        const uint32_t abar = abar_s[i];
        // arm this step's receive barrier: the peer's reduced spectrum (4 x 128 complex) arrives as st.async writes
        if (ctid == 0) mbar_arrive_expect_tx(&xbar[par], (uint32_t)(kSlotCplx * 16));
        cplx rown[4];
        {
          uint32_t t_re[4], t_im[4];
          brs::load_t(T, acc, abar, args.offset, t_re, t_im);
          brs::fwd_pass_a<BGBIT, MAGIC>(T, g, t_re, t_im, exch);
        }
        named_sync<brs::kT>(g + 1);
        {
          cplx tb[4], tcd[4], y[4];
          tm_load4(t_b, tb);
          tm_load4(t_cd, tcd);
          brs::fwd_pass_b(T, exch, tb[0], tb[1], tb[2], y);
          xchg_fwd(tq0, y);
          brs::r4<false>(y, tcd[0], tcd[1]);
          xchg_fwd(tq1, y);
          brs::r4<false>(y, tcd[2], tcd[3]);
          mbar_wait(&full[g], parity);
          const cplx *rw = ring + g * brs::kRowCplx + T;
          // The PEER's output first: its partial crosses distributed shared memory and is the longest leg of the
          // step, so it is MACed, reduced over the l groups and on its way before this CTA's own output is touched.
          cplx *mine = part + (size_t)par * (2 * L * kSlotCplx) + (size_t)g * kSlotCplx + T;
          {
            cplx rp[4];
#pragma unroll
            for (int kd = 0; kd < 4; kd++) {
              rp[kd] = mk(0.0, 0.0);
              cfma(rp[kd], y[kd], rw[(size_t)(kd * 2 + (int)peer) * brs::kT]);
            }
#pragma unroll
            for (int kd = 0; kd < 4; kd++) mine[(size_t)L * kSlotCplx + kd * brs::kT] = rp[kd];   // peer's: slots l..2l-1
          }
          asm volatile("bar.sync 9, %0;" ::"n"(kCons) : "memory");   // every group's partial of the peer's output is in place
          {
            // all l x 128 threads share the reduce-and-ship: bin b of the 512; the group that runs the inverse
            // afterwards (group 0) gets the fewest
            const cplx *src = part + (size_t)par * (2 * L * kSlotCplx) + (size_t)L * kSlotCplx;
            const uint32_t dst = part_remote + (uint32_t)((2 * (size_t)2 * L * kSlotCplx + (size_t)par * kSlotCplx) * 16);
            for (int bin = (ctid + kCons - brs::kT) % kCons; bin < kSlotCplx; bin += kCons) {
              cplx v = src[bin];
#pragma unroll
              for (int sl = 1; sl < L; sl++) v = cadd(v, src[(size_t)sl * kSlotCplx + bin]);
              st_async_remote(dst + (uint32_t)(bin * 16), v, xbar_remote + 8u * par);
            }
          }
#pragma unroll
          for (int kd = 0; kd < 4; kd++) {
            rown[kd] = mk(0.0, 0.0);
            cfma(rown[kd], y[kd], rw[(size_t)(kd * 2 + (int)rank) * brs::kT]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[g]);
          if (g != 0) {
#pragma unroll
            for (int kd = 0; kd < 4; kd++) mine[kd * brs::kT] = rown[kd];                           // ours: slots 1..l-1
          }
        }
        cta_sync();   // our own output's partials visible; forward exchange buffers free
        if (g == 0) {
          cplx s[4];
          const cplx *src = part + (size_t)par * (2 * L * kSlotCplx) + T;
#pragma unroll
          for (int kd = 0; kd < 4; kd++) s[kd] = rown[kd];
#pragma unroll
          for (int sl = 1; sl < L; sl++)
#pragma unroll
            for (int kd = 0; kd < 4; kd++) s[kd] = cadd(s[kd], src[(size_t)sl * kSlotCplx + kd * brs::kT]);
          mbar_wait_cluster(&xbar[par], xphase);
          const cplx *rem = part + 2 * (size_t)2 * L * kSlotCplx + (size_t)par * kSlotCplx + T;   // the peer's reduced spectrum
#pragma unroll
          for (int kd = 0; kd < 4; kd++) s[kd] = cadd(s[kd], rem[kd * brs::kT]);
          cplx ti[4];
          tm_load4(t_cbi, ti);
          brs::r4_plain<true>(s);
          xchg_inv(tq0, s);
          brs::r4<true>(s, ti[0], ti[1]);
          xchg_inv(tq1, s);
          brs::r4<true>(s, ti[2], ti[3]);
          brs::inv_store_b(T, s, exch);
          named_sync<brs::kT>(1);
          cplx ta[4], ut[4];
          tm_load4(t_ai, ta);
          tm_load4(t_ut, ut);
          brs::inv_pass_a<EXACT, MAGIC>(T, exch, ta[0], ta[1], ta[2], ut, acc);
        }
        cta_sync();
      } else {
        mbar_wait(&full[g], parity);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[g]);
      }
      parity ^= 1;
    }
    if (active) {
      // epilogue: every mode's output splits cleanly by polynomial (trlwe.rs:106-136)
      if (args.out_mode == BR_OUT_TRLWE) {
        uint32_t *o = args.out + ct * 2 * kN + rank * kN;
        for (int x = ctid; x < kN; x += kCons) o[x] = acc[x];
      } else {
        const uint32_t m = args.out_mode == BR_OUT_EXTRACT ? (uint32_t)kN : n;
        uint32_t *o = args.out + ct * (m + 1);
        if (rank == 0) {
          for (uint32_t x = ctid; x < m; x += kCons) o[x] = x == 0 ? acc[0] : ~acc[m - x];
        } else if (ctid == 0) {
          o[m] = acc[0];
        }
      }
    }
    cta_sync();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cta_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) tmem_dealloc_512(tbase);
}

template <int L, int BGBIT>
cudaError_t launch_cluster(const BrArgs &args, int num_sms, cudaStream_t stream) {
  constexpr bool MAGIC = (L == 3 && BGBIT == 6);
  if (!args.bsk3 || !args.tw_s) return cudaErrorInvalidValue;
  auto kern = blind_rotate_cluster_kernel<L, BGBIT, MAGIC>;
  const int smem = L * kStageBytes + kN * 4 + L * (8 * brs::kInvPitch * 16) + 2 * (2 * L * 4 * brs::kT * 16) +
                   2 * (4 * brs::kT * 16) + 2432 + (2 * L + 2) * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int clusters = (int)(args.count < (size_t)(num_sms / 2) ? args.count : (size_t)(num_sms / 2));
  if (clusters < 1) clusters = 1;
  kern<<<2 * clusters, L * brs::kT + 32, smem, stream>>>(args);   // __cluster_dims__(2): pairs of CTAs
  return cudaGetLastError();
}

template <int L, int BGBIT>
cudaError_t launch_latency_s(const BrArgs &args, int num_sms, cudaStream_t stream) {
  constexpr bool MAGIC = (L == 3 && BGBIT == 6);
  if (!args.bsk3 || !args.tw_s) return cudaErrorInvalidValue;
  auto kern = blind_rotate_latency_s_kernel<L, BGBIT, MAGIC>;
  const int smem = 2 * L * kStageBytes + 2 * kN * 4 + L * (2 * kHalf * 16 + 2 * 4 * brs::kT * 16) + 2432 +
                   2 * 2 * L * 8 + 16;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = (int)(args.count < (size_t)num_sms ? args.count : (size_t)num_sms);
  if (grid < 1) grid = 1;
  kern<<<grid, L * brs::kT + 32, smem, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace

cudaError_t br_launch_cluster(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  if (l == 3 && bgbit == 6) return launch_cluster<3, 6>(args, num_sms, stream);
  if (l == 2 && bgbit == 10) return launch_cluster<2, 10>(args, num_sms, stream);
  if (l == 1 && bgbit == 18) return launch_cluster<1, 18>(args, num_sms, stream);
  if (l == 1 && bgbit == 22) return launch_cluster<1, 22>(args, num_sms, stream);
  if (l == 1 && bgbit == 23) return launch_cluster<1, 23>(args, num_sms, stream);
  return cudaErrorInvalidValue;
}

cudaError_t br_launch_latency_s(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  if (l == 3 && bgbit == 6) return launch_latency_s<3, 6>(args, num_sms, stream);
  if (l == 2 && bgbit == 10) return launch_latency_s<2, 10>(args, num_sms, stream);
  if (l == 1 && bgbit == 18) return launch_latency_s<1, 18>(args, num_sms, stream);
  if (l == 1 && bgbit == 22) return launch_latency_s<1, 22>(args, num_sms, stream);
  if (l == 1 && bgbit == 23) return launch_latency_s<1, 23>(args, num_sms, stream);
  return cudaErrorInvalidValue;
}

cudaError_t br_launch_s(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  if (l == 3 && bgbit == 6) return launch_s<3, 6, 4>(args, num_sms, stream);
  if (l == 2 && bgbit == 10) return launch_s<2, 10, 4>(args, num_sms, stream);
  if (l == 1 && bgbit == 18) return launch_s<1, 18, 4>(args, num_sms, stream);
  if (l == 1 && bgbit == 22) return launch_s<1, 22, 4>(args, num_sms, stream);
  if (l == 1 && bgbit == 23) return launch_s<1, 23, 4>(args, num_sms, stream);
  return cudaErrorInvalidValue;
}
