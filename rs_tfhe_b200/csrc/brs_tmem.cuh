// brs_tmem.cuh -- tensor-memory primitives of the 128-thread kernels (blind_rotate_s.cu, fft_seam.cu):
// per-thread constant blocks and the intra-warp transform exchanges (tcgen05.st/ld shape pairs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "br_core.cuh"

namespace brt {

using br::mk;

#define R16(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
               "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
#define W16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define RW16(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), \
                "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
#define R8(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
#define W8(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
#define RW8(v) "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])

__device__ __forceinline__ void tm_st32_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      R16(r)
      : "memory");
}
__device__ __forceinline__ void tm_ld32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : W16(r)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tm_st16x256_x2(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), R8(r)
               : "memory");
}
__device__ __forceinline__ void tm_ld16x256_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : W8(r)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the waits carry the loaded registers so the uses cannot be hoisted above them
__device__ __forceinline__ void tm_wait_ld16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : RW16(r) : : "memory");
}
__device__ __forceinline__ void tm_wait_ld8x2(uint32_t (&a)[8], uint32_t (&b)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : RW8(a), RW8(b) : : "memory");
}
__device__ __forceinline__ uint32_t lo32(double d) { return (uint32_t)__double2loint(d); }
__device__ __forceinline__ uint32_t hi32(double d) { return (uint32_t)__double2hiint(d); }
__device__ __forceinline__ double mkd(uint32_t hi, uint32_t lo) { return __hiloint2double((int)hi, (int)lo); }

// 4 complex constants of this thread from 16 TMEM columns
__device__ __forceinline__ void tm_load4(uint32_t taddr, cplx (&t)[4]) {
  uint32_t r[16];
  tm_ld32_x16(taddr, r);
  tm_wait_ld16(r);
#pragma unroll
  for (int i = 0; i < 4; i++) t[i] = mk(mkd(r[4 * i + 1], r[4 * i]), mkd(r[4 * i + 3], r[4 * i + 2]));
}
__device__ __forceinline__ void tm_store4(uint32_t taddr, const cplx (&t)[4]) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    r[4 * i] = lo32(t[i].x); r[4 * i + 1] = hi32(t[i].x);
    r[4 * i + 2] = lo32(t[i].y); r[4 * i + 3] = hi32(t[i].y);
  }
  tm_st32_x16(taddr, r);
}

// Exchange (forward direction): value slot s of writer lane l' -> reader lane (l'[2:0], s), register
// (l'[4], l'[3]).  Column of (slot s, im, word b) = 8 im + 2 s + b.
__device__ __forceinline__ void xchg_fwd(uint32_t tq, cplx (&v)[4]) {
  {
    uint32_t r[16];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      r[2 * s] = lo32(v[s].x); r[2 * s + 1] = hi32(v[s].x);
      r[8 + 2 * s] = lo32(v[s].y); r[8 + 2 * s + 1] = hi32(v[s].y);
    }
    tm_st32_x16(tq, r);
  }
  tm_wait_st();
  uint32_t a0[8], a1[8];
  tm_ld16x256_x2(tq, a0);
  tm_ld16x256_x2(tq + (16u << 16), a1);
  tm_wait_ld8x2(a0, a1);
  // register 4 rep + 2 h + b of half H: rep = im, value index 2 H + h
#pragma unroll
  for (int h = 0; h < 2; h++) {
    v[h] = mk(mkd(a0[2 * h + 1], a0[2 * h]), mkd(a0[4 + 2 * h + 1], a0[4 + 2 * h]));
    v[2 + h] = mk(mkd(a1[2 * h + 1], a1[2 * h]), mkd(a1[4 + 2 * h + 1], a1[4 + 2 * h]));
  }
}
// Exchange (inverse direction): value 2H + h of writer lane l -> reader lane (H, h, l[4:2]), slot l[1:0]
__device__ __forceinline__ void xchg_inv(uint32_t tq, cplx (&u)[4]) {
#pragma unroll
  for (int H = 0; H < 2; H++) {
    uint32_t r[8];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const cplx z = u[2 * H + h];
      r[2 * h] = lo32(z.x); r[2 * h + 1] = hi32(z.x);
      r[4 + 2 * h] = lo32(z.y); r[4 + 2 * h + 1] = hi32(z.y);
    }
    tm_st16x256_x2(tq + ((uint32_t)(16 * H) << 16), r);
  }
  tm_wait_st();
  uint32_t r[16];
  tm_ld32_x16(tq, r);
  tm_wait_ld16(r);
#pragma unroll
  for (int s = 0; s < 4; s++) u[s] = mk(mkd(r[2 * s + 1], r[2 * s]), mkd(r[8 + 2 * s + 1], r[8 + 2 * s]));
}


// TMEM allocation for a CTA of 128-thread groups: warp 0 allocates all 512 columns; the caller
// follows with a CTA-wide barrier + tcgen05 fences (tmem_alloc_publish).
__device__ __forceinline__ void tmem_alloc_512(uint32_t *slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
      (uint32_t)__cvta_generic_to_shared(slot_in_smem)));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t tbase) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

}  // namespace brt
