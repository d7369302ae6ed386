// fft_seam.cu -- the FFTProcessor seam (reference: trait src/fft/mod.rs:80-107, the active
// implementation src/fft/klemsa.rs:88-174) as standalone batch kernels:
//   MODE_IFFT     ifft::<1024>      torus u32[1024] (read as i32) -> f64[1024] = 2*F, re[0..512) | im[0..512)
//   MODE_FFT      fft::<1024>       f64[1024] -> torus u32[1024]   (x0.5, inverse, untwist, /512, round half away)
//   MODE_POLYMUL  poly_mul::<1024>  a (x) b mod X^1024+1            (ifft, ifft, product x0.5, fft)
// They run the SAME per-thread passes and tensor-memory exchanges as the blind rotation
// (brs_core.cuh, brs_tmem.cuh): 128 threads per polynomial, 4 polynomials per CTA.  Outside the hot
// path these exist so the transforms can be tested and timed in isolation (the reference's own FFT
// tests, src/fft/mod.rs:118-255, and micro-benches, benches/gate_benchmarks.rs:92-126).
#include <cuda_runtime.h>
#include <stdint.h>

#include "br_ptx.cuh"
#include "brs_core.cuh"
#include "brs_tmem.cuh"
#include "kernels.h"

using namespace br;
using namespace brp;
using namespace brt;

namespace {

constexpr int kG = 4;
constexpr int kSeamThreads = kG * brs::kT;

// forward transform of the polynomial x (i32 view) of this group: leaves bins brs::bin_of(T, kd) in y[kd]
__device__ __forceinline__ void seam_forward(int T, int g, const uint32_t *x, cplx *exch, uint32_t t_b,
                                             uint32_t t_cd, uint32_t tq0, uint32_t tq1, cplx (&y)[4]) {
  uint32_t x_re[4], x_im[4];
#pragma unroll
  for (int a = 0; a < 4; a++) { x_re[a] = x[128 * a + T]; x_im[a] = x[128 * a + T + kHalf]; }
  brs::fwd_pass_a_i32(T, x_re, x_im, exch);
  named_sync<brs::kT>(g + 1);
  cplx tb[4], tcd[4];
  tm_load4(t_b, tb);
  brs::fwd_pass_b(T, exch, tb[0], tb[1], tb[2], y);
  xchg_fwd(tq0, y);
  tm_load4(t_cd, tcd);
  brs::r4<false>(y, tcd[0], tcd[1]);
  xchg_fwd(tq1, y);
  brs::r4<false>(y, tcd[2], tcd[3]);
  named_sync<brs::kT>(g + 1);   // the exchange buffer may be rewritten
}

// inverse transform of the spectrum s[kd] (already scaled) into out[0..1024) (torus, rounded)
__device__ __forceinline__ void seam_inverse(int T, int g, cplx (&s)[4], cplx *exch, uint32_t *stage,
                                             uint32_t t_cbi, uint32_t t_ai, uint32_t t_ut, uint32_t tq0,
                                             uint32_t tq1) {
  for (int x = T; x < kN; x += brs::kT) stage[x] = 0u;
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
  brs::inv_pass_a<false, false>(T, exch, ta[0], ta[1], ta[2], ut, stage);
  named_sync<brs::kT>(g + 1);
}

template <int MODE>
__global__ void __launch_bounds__(kSeamThreads, 1)
fft_seam_kernel(const cplx *__restrict__ tw_s, const void *__restrict__ in_a, const uint32_t *__restrict__ in_b,
                void *__restrict__ out, size_t count) {
  constexpr int kExchBytes = 8 * brs::kInvPitch * 16;   // >= 512 complex
  extern __shared__ __align__(128) uint8_t smem[];      // kG x (exchange buffer + output staging), TMEM base slot
  uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(smem + kG * (kExchBytes + kN * 4));
  const int warp = threadIdx.x >> 5;
  // compute-sanitizer's synccheck models the tcgen05 allocation / wait instructions as barrier operations and
  // reports "Missing init" for a kernel that uses them without ever initialising an mbarrier (the blind-rotation
  // kernels have their TMA ring).  One never-used 8-byte barrier keeps the whole library synccheck-clean.
  if (threadIdx.x == 0) mbar_init(reinterpret_cast<uint64_t *>(tmem_base_s + 2), 1);
  if (warp == 0) tmem_alloc_512(tmem_base_s);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = *tmem_base_s;
  const int g = warp >> 2, T = threadIdx.x & (brs::kT - 1);
  cplx *exch = reinterpret_cast<cplx *>(smem + g * (kExchBytes + kN * 4));
  uint32_t *stage = reinterpret_cast<uint32_t *>(smem + g * (kExchBytes + kN * 4) + kExchBytes);
  const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)g * 128u;
  const uint32_t t_b = taddr, t_cd = taddr + 16, t_cbi = taddr + 32, t_ai = taddr + 48, t_ut = taddr + 64;
  const uint32_t tq0 = taddr + 80, tq1 = taddr + 96;
  {
    const cplx *tw = tw_s + (size_t)T * brs::kTwPerThread;
#pragma unroll
    for (int blk = 0; blk < 5; blk++) {
      cplx t[4];
#pragma unroll
      for (int k = 0; k < 4; k++) t[k] = tw[4 * blk + k];
      tm_store4(taddr + 16 * blk, t);
    }
    tm_wait_st();
  }
  const size_t quads = (count + kG - 1) / kG;
  for (size_t q = blockIdx.x; q < quads; q += gridDim.x) {
    const size_t p = q * kG + g;
    const bool active = p < count;      // whole groups only: barriers below are per group
    if (!active) continue;
    if (MODE == MODE_IFFT) {
      cplx y[4];
      seam_forward(T, g, static_cast<const uint32_t *>(in_a) + p * kN, exch, t_b, t_cd, tq0, tq1, y);
      double *o = static_cast<double *>(out) + p * kN;
#pragma unroll
      for (int kd = 0; kd < 4; kd++) {   // klemsa.rs:110-114: x2, re | im split
        const int k = brs::bin_of(T, kd);
        o[k] = y[kd].x * 2.0;
        o[k + kHalf] = y[kd].y * 2.0;
      }
    } else {
      cplx s[4];
      if (MODE == MODE_FFT) {
        const double *f = static_cast<const double *>(in_a) + p * kN;
#pragma unroll
        for (int kd = 0; kd < 4; kd++) {   // klemsa.rs:126 (x0.5) and :136 (1/512): one exact scaling
          const int k = brs::bin_of(T, kd);
          s[kd] = mk(f[k] * (1.0 / 1024.0), f[k + kHalf] * (1.0 / 1024.0));
        }
      } else {
        cplx ya[4], yb[4];
        seam_forward(T, g, static_cast<const uint32_t *>(in_a) + p * kN, exch, t_b, t_cd, tq0, tq1, ya);
        seam_forward(T, g, in_b + p * kN, exch, t_b, t_cd, tq0, tq1, yb);
#pragma unroll
        for (int kd = 0; kd < 4; kd++) {   // (2Fa)(2Fb) x 0.5 (klemsa.rs:169-170), then the fft scaling 1/1024
          const cplx pr = cmul(ya[kd], yb[kd]);
          s[kd] = mk(pr.x * (1.0 / 512.0), pr.y * (1.0 / 512.0));
        }
      }
      seam_inverse(T, g, s, exch, stage, t_cbi, t_ai, t_ut, tq0, tq1);
      uint32_t *o = static_cast<uint32_t *>(out) + p * kN;
      for (int x = T; x < kN; x += brs::kT) o[x] = stage[x];
      named_sync<brs::kT>(g + 1);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) tmem_dealloc_512(tbase);
}

}  // namespace

cudaError_t fft_seam_launch(int mode, const cplx *tw_s, const void *in_a, const uint32_t *in_b, void *out,
                            size_t count, int num_sms, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const size_t quads = (count + kG - 1) / kG;
  const int grid = (int)(quads < (size_t)num_sms ? quads : (size_t)num_sms);
  const int smem = kG * (8 * brs::kInvPitch * 16 + kN * 4) + 16;
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kSeamThreads, smem, stream>>>(tw_s, in_a, in_b, out, count);
    return cudaGetLastError();
  };
  if (mode == MODE_IFFT) return go(fft_seam_kernel<MODE_IFFT>);
  if (mode == MODE_FFT) return go(fft_seam_kernel<MODE_FFT>);
  if (mode == MODE_POLYMUL) return go(fft_seam_kernel<MODE_POLYMUL>);
  return cudaErrorInvalidValue;
}
