// aux.cu -- key re-layout at upload, device-side LUT generation, sample extraction.
#include "kernels.h"
#include "brs_core.cuh"

namespace {

// Reference TRGSWLv1FFT image (trgsw.rs:52-68; f64[n][2l][2][1024], re|im split,
// natural bin order, x2-scaled per klemsa.rs:110-114) -> device order
// cplx[n][2l][8 k2][2 o][64 v], scaled by 1/1024 (folds klemsa.rs:112,126,136 and
// trgsw.rs:137-140; power-of-two scaling is exact).
__global__ void bsk_relayout_kernel(const double *__restrict__ src, cplx *__restrict__ dst,
                                    size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = idx & 63;
  const int o = (idx >> 6) & 1;
  const int k2 = (idx >> 7) & 7;
  const size_t ir = idx >> 10;  // i * 2l + r
  const int k = br::bin_of(v, k2);
  const double *s = src + (ir * 2 + o) * br::kN;
  dst[idx] = br::mk(s[k] * (1.0 / 1024.0), s[k + br::kHalf] * (1.0 / 1024.0));
}

// Device-order key rows -> the thread order of the TMEM-exchange kernel (blind_rotate.cu): slot
// T = 32 W + l of a (k2, o) slice holds bin k0 + 8 k1 + 64 k2 with k0 = 4 W + l[3:2],
// k1 = 4 l[4] + l[1:0], negated where k2 is odd and l[4] is set (the kernel's pass C produces those
// bins negated); the standard order keeps that bin in slot 8 k0 + k1.
__global__ void bsk_permute_kernel(const cplx *__restrict__ src, cplx *__restrict__ dst, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int T = idx & 63;
  const int k0 = 4 * (T >> 5) + ((T >> 2) & 3), k1 = 4 * ((T >> 4) & 1) + (T & 3);
  const cplx v = src[(idx & ~(size_t)63) + 8 * k0 + k1];
  const bool neg = ((idx >> 7) & 1) && ((T >> 4) & 1);   // idx = ((row * 8 + k2) * 2 + o) * 64 + T
  dst[idx] = neg ? br::mk(-v.x, -v.y) : v;
}

// Device-order key rows -> the order of the 128-thread kernel (blind_rotate_s.cu): slot T of a
// (kd, o) slice holds bin brs::bin_of(T, kd); the standard order keeps bin k0 + 8 k1 + 64 k2 in slot
// 8 k0 + k1 of slice (k2, o).
__global__ void bsk_permute_s_kernel(const cplx *__restrict__ src, cplx *__restrict__ dst, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // ((row * 4 + kd) * 2 + o) * 128 + T
  if (idx >= total) return;
  const int T = idx & 127, o = (idx >> 7) & 1, kd = (idx >> 8) & 3;
  const size_t row = idx >> 10;
  const int k = brs::bin_of(T, kd);
  const int k0 = k & 7, k1 = (k >> 3) & 7, k2 = k >> 6;
  dst[idx] = src[((row * 8 + k2) * 2 + o) * 64 + 8 * k0 + k1];
}

// Reference KSK image (key.rs:102-122: u32[N*t*2^basebit][n+1]) -> rows padded to
// `stride` words (zero fill) plus one trailing all-zero row.
__global__ void ksk_relayout_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst,
                                    uint32_t rows, uint32_t n, uint32_t stride) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = ((size_t)rows + 1) * stride;
  if (idx >= total) return;
  uint32_t row = (uint32_t)(idx / stride), x = (uint32_t)(idx % stride);
  dst[idx] = (row < rows && x <= n) ? src[(size_t)row * (n + 1) + x] : 0u;
}

__device__ __forceinline__ uint32_t div_round(uint32_t a, uint32_t b) { return (a + b / 2) / b; }

// lut/generator.rs:89-137 (+ encoder.rs:66-73, utils.rs:9-12): one thread per coefficient.
__global__ void lut_generate_kernel(const uint32_t *__restrict__ f_table, uint32_t m, double scale,
                                    uint32_t *__restrict__ tv_slot) {
  const uint32_t N = br::kN;
  const uint32_t i = threadIdx.x;
  const uint32_t offset = div_round(N, 2 * m);
  const uint32_t srci = (i + offset) % N;
  // box x with div_round(x*N, m) <= srci < div_round((x+1)*N, m)
  uint32_t x = (uint32_t)(((uint64_t)srci * m) / N);
  while (x + 1 < m && div_round((x + 1) * N, m) <= srci) x++;
  while (x > 0 && div_round(x * N, m) > srci) x--;
  uint32_t val = 0;
  if (div_round(x * N, m) <= srci && srci < div_round((x + 1) * N, m)) {
    uint32_t msg = f_table[x] % m;
    double d = fmod((double)msg * scale, 1.0) * 4294967296.0;
    val = (uint32_t)(unsigned long long)(long long)d;
  }
  if (i >= N - offset) val = 0u - val;
  tv_slot[i] = 0u;        // poly.a
  tv_slot[N + i] = val;   // poly.b
}

// trlwe.rs:106-120 with k = 0
__global__ void extract_kernel(const uint32_t *__restrict__ trlwe, uint32_t *__restrict__ ext,
                               size_t count) {
  const uint32_t N = br::kN;
  size_t ct = blockIdx.x;
  if (ct >= count) return;
  const uint32_t *a = trlwe + ct * 2 * N;
  uint32_t *o = ext + ct * (N + 1);
  for (uint32_t x = threadIdx.x; x <= N; x += blockDim.x) {
    uint32_t v;
    if (x == 0) v = a[0];
    else if (x == N) v = a[N];
    else v = ~a[N - x];
    o[x] = v;
  }
}

// Roofline denominator: sustained DFMA rate of this GPU right now (8 independent
// chains per thread, no memory traffic).  Diagnostic only -- never on the hot path.
__global__ void fp64_probe_kernel(double *sink, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999999, c = 1e-12;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456) sink[0] = r;  // never true; keeps the chains alive
}

// The same stream with three DISTINCT register operands per DFMA (d_i = a_i * b_i + d_i): what the
// blind rotation's butterflies and MACs look like to the register file -- every lane holds its own
// twiddle and key value.  On B200 this runs at 2/3 of the rate above: the FP64 unit receives one
// 64-bit operand per lane per cycle (tools/probe/dfma_operand_probe.cu, profiles/r2_fp64_operand_probe.json).
__global__ void fp64_probe3_kernel(double *sink, int iters) {
  double a[8], b[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    b[i] = 1.0 - 1e-9 * (threadIdx.x + 2 * i);
    d[i] = 1e-9 * (i + 3);
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(a[i]), "d"(b[i]));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += d[i] + a[i] + b[i];
  if (r == 123.456) sink[0] = r;
}

}  // namespace

cudaError_t fp64_probe_launch(double *d_sink, int blocks, int iters, bool three_operands, cudaStream_t stream) {
  if (three_operands) fp64_probe3_kernel<<<blocks, 256, 0, stream>>>(d_sink, iters);
  else fp64_probe_kernel<<<blocks, 256, 0, stream>>>(d_sink, iters);
  return cudaGetLastError();
}

cudaError_t bsk_relayout_launch(const double *src_ref, cplx *dst, uint32_t n, uint32_t l2,
                                cudaStream_t stream) {
  size_t total = (size_t)n * l2 * br::kChunkCplx;
  bsk_relayout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src_ref, dst, total);
  return cudaGetLastError();
}
cudaError_t bsk_permute_launch(const cplx *src, cplx *dst, size_t rows, cudaStream_t stream) {
  const size_t total = rows * br::kChunkCplx;
  bsk_permute_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, total);
  return cudaGetLastError();
}

cudaError_t bsk_permute_s_launch(const cplx *src, cplx *dst, size_t rows, cudaStream_t stream) {
  const size_t total = rows * brs::kRowCplx;
  bsk_permute_s_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, total);
  return cudaGetLastError();
}

cudaError_t ksk_relayout_launch(const uint32_t *src_ref, uint32_t *dst, uint32_t rows, uint32_t n,
                                uint32_t stride, cudaStream_t stream) {
  size_t total = ((size_t)rows + 1) * stride;
  ksk_relayout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src_ref, dst, rows, n,
                                                                          stride);
  return cudaGetLastError();
}
cudaError_t lut_generate_launch(const uint32_t *d_f_table, uint32_t modulus, double scale,
                                uint32_t *d_tv_slot, cudaStream_t stream) {
  lut_generate_kernel<<<1, br::kN, 0, stream>>>(d_f_table, modulus, scale, d_tv_slot);
  return cudaGetLastError();
}
cudaError_t extract_launch(const uint32_t *d_trlwe, uint32_t *d_ext, size_t count,
                           cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  extract_kernel<<<(unsigned)count, 256, 0, stream>>>(d_trlwe, d_ext, count);
  return cudaGetLastError();
}
