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
#include "br_ptx.cuh"
#include "kernels.h"

using namespace br;
using namespace brp;

#ifndef BR_PRODUCER_SLEEP_NS
#define BR_PRODUCER_SLEEP_NS 256
#endif
#ifndef BR_DEFAULT_VARIANT
#define BR_DEFAULT_VARIANT 8
#endif

namespace {

__device__ __forceinline__ void group_sync(int g) { named_sync<64>(g + 1); }
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  brp::mbar_wait_backoff(bar, parity, BR_PRODUCER_SLEEP_NS);
}

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
  uint32_t stage = 0, parity = 0;
  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = ((size_t)rd * G + g) * grid + blockIdx.x;
    const bool active = ct < args.count;
    if (active) prologue<64>(args, ct, tid, abar_s, acc);   // K0 (br_ptx.cuh)
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
    if (active) epilogue<64>(args, ct, tid, acc);
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

  uint32_t parity = 0;
  const int p0 = (2 * g) / L, p1 = (2 * g + 1) / L;       // polynomials of this group's two rows
  const int d0 = (2 * g) % L, d1 = (2 * g + 1) % L;       // and their digits

  for (uint32_t rd = 0; rd < rounds; rd++) {
    const size_t ct = (size_t)rd * grid + blockIdx.x;
    const bool active = ct < args.count;
    if (active) prologue<64 * NG>(args, ct, ctid, abar_s, acc);   // K0 (br_ptx.cuh)
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

    if (active) epilogue<64 * NG>(args, ct, ctid, acc);
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

int br_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TFHE_BR_VARIANT");
    v = e ? atoi(e) : BR_DEFAULT_VARIANT;
    if (v != 8 && v != 9) v = BR_DEFAULT_VARIANT;
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
  if (args.count <= (size_t)br_latency_threshold(num_sms)) {
    // TFHE_BR_LATENCY_KERNEL=x selects the round-1 shape (l groups of 64 threads, gadgets with l > 1)
    static const bool old_shape = [] { const char *e = getenv("TFHE_BR_LATENCY_KERNEL"); return e && e[0] == 'x'; }();
    // Up to one ciphertext per SM PAIR runs on a 2-CTA cluster: CTA c owns accumulator polynomial c, the partial
    // spectra of the other output cross distributed shared memory as st.async writes that count their bytes on
    // the peer's mbarrier (no cluster fence, no remote arrive: 1.48 ms per PBS at 128 bits against 2.00 ms on one
    // SM; with a fence + remote arrive it was 2.10 -- profiles/r2_experiments.json).  TFHE_BR_CLUSTER=0 turns it off.
    static const bool use_cluster = [] { const char *e = getenv("TFHE_BR_CLUSTER"); return !(e && e[0] == '0'); }();
    if (args.bsk3 && !old_shape && use_cluster && args.count <= (size_t)(num_sms / 2))
      return br_launch_cluster(L, BGBIT, args, num_sms, stream);
    if (args.bsk3 && !old_shape) return br_launch_latency_s(L, BGBIT, args, num_sms, stream);
    if (L > 1) return launch_latency<L, BGBIT>(args, num_sms, stream);
  }
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
    // a tail of at most one ciphertext per SM goes to the latency shapes (2-CTA cluster / one CTA per ciphertext:
    // 1.5 / 2.0 ms against 2.8 ms for the throughput kernel with one slot of four filled)
    const BrArgs t = br_slice(args, full, tail);
    if (full && tail <= (size_t)br_latency_threshold(num_sms)) return launch_t<L, BGBIT>(t, num_sms, stream);
    return br_launch_s(L, BGBIT, t, num_sms, stream);
  }
  return cudaErrorInvalidValue;
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
                       args.bsk3 && !(args.count <= (size_t)br_latency_threshold(num_sms));
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
