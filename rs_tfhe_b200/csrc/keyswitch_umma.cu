// keyswitch_umma.cu -- K4 on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM
// accumulators) for the parameter sets with basebit 2..6 (gate sets t = 7 / 8 / 9, UINT1-6).
//
// Same exact integer GEMM as keyswitch_mma.cu (reference src/trgsw.rs:332-360):
//   out[ct][x] = (x == n ? b : 0) - sum_{i,j} KSK[i][j][digit_j(a_i + PREC_OFFSET)][x]
// with the key split into byte planes and the digit selection as a one-hot u8 matrix, so
// u8 x u8 -> s32 accumulation is exact (each sum < N*t*255 < 2^24) and the wrapping recombination
// of the planes gives the reference's bits.  What changes is where it runs:
//   * one CTA = 128 ciphertexts (TMEM lanes) x 480 accumulator columns (120 output words) in
//     tensor memory: two tcgen05.mma.cta_group::1.kind::i8 of M128 x N240 x K32 per K step;
//   * A (one-hot) never exists in memory: the four builder warps compute it from the digits and
//     write it straight into TMEM (tcgen05.st.32x32b.x16, thread = ciphertext row), the MMA reads
//     it from there (A-from-TMEM form), double buffered in the 32 columns the accumulators leave;
//   * B (key bytes) is pre-tiled at key load into the canonical K-major / no-swizzle operand
//     order (ku_layout.h), so a pipeline stage (K = 64: 30 KB) is ONE contiguous 1-D TMA bulk copy
//     into a 6-stage ring; tcgen05.commit releases ring slots and A buffers;
//   * epilogue: tcgen05.ld the accumulators, recombine planes, transpose through a warp-private
//     shared-memory tile so the global stores are coalesced.
// Warp roles: 0-3 build A / run the epilogue, 4 issues the MMAs (one elected lane), 5 streams B.
// CTAs of one wave share the same N tile and walk K in step, so the key tile (17.7 MB at 128-bit)
// is read from HBM once and otherwise served by L2.
// Operand layouts (instruction descriptor, shared-memory descriptor, A-in-TMEM rows) were verified
// bit for bit on B200 with tools/probe/umma_i8_probe.cu before this kernel was written.
#include "kernels.h"
#include "ku_layout.h"

namespace {

constexpr int KU_BSTAGES = 6;
constexpr int KU_THREADS = 192;
constexpr int KU_XPITCH = 33;
constexpr int KU_ACOL = ku::kCols;   // first TMEM column of the A buffers (2 x 16 columns)
constexpr int KU_SMEM = KU_BSTAGES * ku::kStageBytes + 4 * 32 * KU_XPITCH * 4 + 17 * 8 + 16;

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
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// shared-memory operand descriptor: K-major, SWIZZLE_NONE, core matrix = 8 rows x 16 B,
// LBO (next 16 B of K) = 128 B, SBO (next 8 rows of N) = 256 B, descriptor version 1
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]^T, u8 x u8 -> s32
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32w(uint32_t taddr, uint32_t (&r)[32]) {
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
}

template <int BB, int T>
__global__ void __launch_bounds__(KU_THREADS, 1) ks_umma_kernel(const KsUmmaArgs a) {
  constexpr int P = 1 << BB;                 // K bytes per (coefficient, digit) pair: one-hot over the digit value
  constexpr int PPS = ku::kStageK / P;       // pairs per pipeline stage
  constexpr int SPB = 16 * T / PPS;          // stages per block of 16 coefficients
  constexpr uint32_t NST = (ku::kRing / 16) * SPB;
  static_assert(BB >= 2 && BB <= 6 && (16 * T) % PPS == 0, "unsupported key-switch base");
  extern __shared__ __align__(128) uint8_t ku_smem[];
  uint8_t *bst = ku_smem;
  uint32_t *xp = reinterpret_cast<uint32_t *>(ku_smem + KU_BSTAGES * ku::kStageBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(xp + 4 * 32 * KU_XPITCH);
  uint64_t *b_full = bars, *b_empty = bars + KU_BSTAGES, *a_full = bars + 2 * KU_BSTAGES;
  uint64_t *a_empty = a_full + 2, *d_full = a_empty + 2;
  uint32_t *tmem_s = reinterpret_cast<uint32_t *>(d_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t mtiles = (uint32_t)((a.count + ku::kM - 1) / ku::kM);
  const uint32_t nt = blockIdx.x / mtiles, mt = blockIdx.x % mtiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < KU_BSTAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_s;

  if (warp == 5) {
    // ===== B producer: one contiguous stage per bulk copy =====
    if (lane == 0) {
      const uint8_t *src = a.key + (size_t)nt * NST * ku::kStageBytes;
      uint32_t s = 0, ph = 0;
      for (uint32_t st = 0; st < NST; st++) {
        mbar_wait(&b_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&b_full[s], ku::kStageBytes);
        tma_load_1d(bst + s * ku::kStageBytes, src + (size_t)st * ku::kStageBytes, ku::kStageBytes,
                    &b_full[s]);
        if (++s == KU_BSTAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // c = S32 (2 << 4), a/b = unsigned 8-bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (2u << 4) | ((uint32_t)(ku::kHalf >> 3) << 17) | ((uint32_t)(ku::kM >> 4) << 24);
      uint32_t bs = 0, bph = 0, as = 0, aph = 0;
      for (uint32_t st = 0; st < NST; st++) {
        mbar_wait(&a_full[as], aph);
        mbar_wait(&b_full[bs], bph);
        tc_fence_after();
        const uint32_t sb = smem_u32(bst + bs * ku::kStageBytes);
#pragma unroll
        for (uint32_t ks = 0; ks < 2; ks++)
#pragma unroll
          for (uint32_t h = 0; h < 2; h++)
            umma_i8_ts(tbase + h * ku::kHalf, tbase + KU_ACOL + as * 16 + ks * 8,
                       b_desc(sb + (ks * 2 + h) * ku::kTileBytes), idesc, (st | ks) != 0 ? 1u : 0u);
        tc_commit(&b_empty[bs]);
        tc_commit(&a_empty[as]);
        if (++bs == KU_BSTAGES) { bs = 0; bph ^= 1; }
        if (++as == 2) { as = 0; aph ^= 1; }
      }
      tc_commit(d_full);
    }
    __syncwarp();
  } else {
    // ===== A builders: thread = ciphertext row warp*32 + lane =====
    constexpr uint32_t prec = 1u << (32 - (1 + BB * T));   // trgsw.rs:338
    const size_t ct0 = (size_t)mt * ku::kM + (size_t)warp * 32;
    uint32_t *xw = xp + warp * 32 * KU_XPITCH;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t pre[32];
    auto fetch = [&](uint32_t sb) {   // coefficients 32 sb .. 32 sb + 31 of the warp's rows, coalesced
#pragma unroll
      for (int rr = 0; rr < 32; rr++) {
        const size_t c = ct0 + rr;
        pre[rr] = (c < a.count) ? __ldg(a.ext + c * (ku::kRing + 1) + 32 * sb + lane) : 0u;
      }
    };
    fetch(0);
    uint32_t as = 0, aph = 0;
    for (uint32_t sb = 0; sb < ku::kRing / 32; sb++) {
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 32; rr++) xw[rr * KU_XPITCH + lane] = pre[rr];
      __syncwarp();
      if (sb + 1 < ku::kRing / 32) fetch(sb + 1);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint32_t ab[16];
#pragma unroll
        for (int c = 0; c < 16; c++) ab[c] = xw[lane * KU_XPITCH + half * 16 + c] + prec;
        // 16 coefficients x T digits = SPB stages of PPS pairs; pair PPS s + pp = (coefficient il, digit j).
        // One-hot over the digit value k in the pair's P bytes (byte k <-> K index P pair + k); digit 0
        // meets the zeroed k = 0 key bytes.
#pragma unroll
        for (int s = 0; s < SPB; s++) {
          uint32_t r[16];
          if (BB == 2) {
#pragma unroll
            for (int c = 0; c < 16; c++) {
              const int qb = 16 * s + c, il = qb / T, j = qb % T;
              r[c] = 1u << ((ab[il] >> (27 - 2 * j)) & 0x18u);
            }
          } else {
#pragma unroll
            for (int pp = 0; pp < PPS; pp++) {
              const int qb = PPS * s + pp, il = qb / T, j = qb % T;
              const uint32_t digit = (ab[il] >> (32 - (j + 1) * BB)) & (uint32_t)(P - 1);
              const uint32_t word = digit >> 2, bit = 1u << ((digit & 3u) * 8);
#pragma unroll
              for (int wq = 0; wq < P / 4; wq++) r[pp * (P / 4) + wq] = (word == (uint32_t)wq) ? bit : 0u;
            }
          }
          mbar_wait(&a_empty[as], aph ^ 1);
          tc_fence_after();
          tmem_st16(lane_base + KU_ACOL + as * 16, r);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[as]);
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      }
    }
    // ===== epilogue: recombine byte planes, out = init - sum (trgsw.rs:343,353-355) =====
    mbar_wait(d_full, 0);
    tc_fence_after();
    const uint32_t x_base = nt * ku::kWords;
#pragma unroll 1
    for (int grp = 0; grp < 4; grp++) {
      const int nch = grp < 3 ? 4 : 3;   // 120 words = 3 x 32 + 24
      __syncwarp();
#pragma unroll 1
      for (int ch = 0; ch < nch; ch++) {
        uint32_t r[32];
        tmem_ld32w(lane_base + (uint32_t)(grp * 4 + ch) * 32, r);
#pragma unroll
        for (int wv = 0; wv < 8; wv++)
          xw[lane * KU_XPITCH + ch * 8 + wv] =
              r[4 * wv] + (r[4 * wv + 1] << 8) + (r[4 * wv + 2] << 16) + (r[4 * wv + 3] << 24);
      }
      __syncwarp();
      const uint32_t x = x_base + grp * 32 + lane;
      if (lane < nch * 8 && x <= a.n) {
        for (int rr = 0; rr < 32; rr++) {
          const size_t c = ct0 + rr;
          if (c >= a.count) break;
          const uint32_t init = (x == a.n) ? a.ext[c * (ku::kRing + 1) + ku::kRing] : 0u;
          a.out[c * (a.n + 1) + x] = init - xw[rr * KU_XPITCH + lane];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

// blob KSK rows (u32[rows + 1][stride], reference row order key.rs:102-122) -> operand tiles
__global__ void ksk_umma_relayout_kernel(const uint32_t *__restrict__ rows, uint32_t stride,
                                         uint32_t *__restrict__ dst, uint32_t n, uint32_t t,
                                         uint32_t basebit, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const ku::Src s = ku::decode(idx, t, basebit);
  uint32_t v = 0;
  if (s.x <= n) {
#pragma unroll
    for (uint32_t b = 0; b < 4; b++) {
      if (s.k0 + b == 0) continue;
      const uint32_t wv = rows[((size_t)s.row0 + b) * stride + s.x];
      v |= ((wv >> (8 * s.plane)) & 0xFFu) << (8 * b);
    }
  }
  dst[idx] = v;
}

template <int BB, int T> cudaError_t launch_t(const KsUmmaArgs &args, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(ks_umma_kernel<BB, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, KU_SMEM);
  if (e != cudaSuccess) return e;
  const size_t mtiles = (args.count + ku::kM - 1) / ku::kM;
  const unsigned grid = (unsigned)(mtiles * ku::n_tiles(args.n));
  ks_umma_kernel<BB, T><<<grid, KU_THREADS, KU_SMEM, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace

// (basebit, t) of src/params.rs:91-404: gate sets and UINT1 (2; 7/8/9), UINT2 (4; 3), UINT4 (5; 3),
// UINT3 (6; 2), UINT5/6 (6; 3).  UINT7/8 (basebit 7) stay on the row-walk kernel.
bool ks_umma_supported(uint32_t basebit, uint32_t t) {
  return (basebit == 2 && t >= 7 && t <= 9) || (basebit == 4 && t == 3) || (basebit == 5 && t == 3) ||
         (basebit == 6 && (t == 2 || t == 3));
}
size_t ks_umma_key_bytes(uint32_t n, uint32_t t, uint32_t basebit) { return ku::key_words(n, t, basebit) * 4; }

cudaError_t ks_umma_launch(const KsUmmaArgs &args, cudaStream_t stream) {
  if (args.count == 0) return cudaSuccess;
  switch (args.basebit * 16 + args.iks_t) {
    case 2 * 16 + 7: return launch_t<2, 7>(args, stream);
    case 2 * 16 + 8: return launch_t<2, 8>(args, stream);
    case 2 * 16 + 9: return launch_t<2, 9>(args, stream);
    case 4 * 16 + 3: return launch_t<4, 3>(args, stream);
    case 5 * 16 + 3: return launch_t<5, 3>(args, stream);
    case 6 * 16 + 2: return launch_t<6, 2>(args, stream);
    case 6 * 16 + 3: return launch_t<6, 3>(args, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t ksk_umma_relayout_launch(const uint32_t *blob_rows, uint32_t stride, uint8_t *dst,
                                     uint32_t n, uint32_t t, uint32_t basebit, cudaStream_t stream) {
  const size_t total = ku::key_words(n, t, basebit);
  ksk_umma_relayout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      blob_rows, stride, reinterpret_cast<uint32_t *>(dst), n, t, basebit, total);
  return cudaGetLastError();
}
