// br_core.cuh -- per-thread building blocks of the blind-rotation kernel.
//
// Everything here is __host__ __device__: blind_rotate.cu strings the phases
// together with named barriers and a TMA ring on the GPU; emu.cpp runs the very
// same functions thread by thread on the CPU so the index/twiddle logic can be
// checked against the oracle without a GPU (tests/test_emulator.py).
//
// One ciphertext is owned by a group of 64 threads.  A negacyclic transform of
// a degree-1024 torus polynomial is a twisted 512-point complex FFT
// (reference: src/fft/klemsa.rs:88-150), done here as three radix-8 passes
// (512 = 8*8*8) with the twist folded into the pass-A twiddles:
//   forward  (decimation in frequency): A (over j2) -> B (over j1) -> C (over j0)
//   inverse  (decimation in time)     : C'(over k2) -> B'(over k1) -> A'(over k0)
// The spectrum stays in the order the passes leave it in (thread v=k0*8+k1,
// register k2 holds bin k0+8*k1+64*k2); the bootstrapping key is permuted to
// that order once at upload, so no reordering pass exists.  All power-of-two
// scale factors of the reference (x2 in klemsa.rs:112, x0.5 in trgsw.rs:137,
// x0.5 and 1/512 in klemsa.rs:126,136) are folded into the uploaded key
// (exact: scaling a double by 2^k never rounds).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BR_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
typedef double2 cplx;
#else
#include <cmath>
#define BR_HD inline
struct alignas(16) cplx { double x, y; };
#endif

namespace br {

constexpr int kN = 1024;        // ring degree
constexpr int kHalf = 512;      // complex points per transform
constexpr int kGroup = 64;      // threads per ciphertext
constexpr int kExchStride = 576; // 512 complex + one 16-byte pad every 8 (bank spread)
constexpr int kChunkCplx = 8 * 2 * 64; // one BSK row in device order: [k2][o][v]

BR_HD cplx mk(double x, double y) { cplx c; c.x = x; c.y = y; return c; }
BR_HD cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
BR_HD cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
BR_HD cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
BR_HD cplx cmulc(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
BR_HD void cfma(cplx &acc, cplx a, cplx b) {
  acc.x += a.x * b.x; acc.x -= a.y * b.y;
  acc.y += a.x * b.y; acc.y += a.y * b.x;
}

// position p (0..511) -> padded slot in an exchange buffer
BR_HD int phys(int p) { return p + (p >> 3); }

// ---- radix-8 butterfly -----------------------------------------------------
// INV=false: X[k] = sum_m x[m] e^{-2 pi i mk/8};  INV=true: conjugate kernel.
template <bool INV> BR_HD cplx mul_i(cplx a) {  // * (-i) forward, * (+i) inverse
  return INV ? mk(-a.y, a.x) : mk(a.y, -a.x);
}
// * sqrt(2) e^{-+ i pi/4} and * sqrt(2) e^{-+ 3 i pi/4}: the 1/sqrt(2) is applied later inside an FMA
template <bool INV> BR_HD cplx mul_w1u(cplx a) {
  return INV ? mk(a.x - a.y, a.x + a.y) : mk(a.x + a.y, a.y - a.x);
}
template <bool INV> BR_HD cplx mul_w3u(cplx a) {
  return INV ? mk(-a.x - a.y, a.x - a.y) : mk(a.y - a.x, -a.x - a.y);
}
// c + s*q and c - s*q as FMAs
BR_HD cplx cfma_s(cplx c, double s, cplx q) { return mk(c.x + s * q.x, c.y + s * q.y); }

template <bool INV> BR_HD void dft8(cplx (&v)[8]) {
  const double s = 0.70710678118654752440;
  cplx a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  cplx a1 = cadd(v[1], v[5]), p5 = mul_w1u<INV>(csub(v[1], v[5]));
  cplx a2 = cadd(v[2], v[6]), a6 = mul_i<INV>(csub(v[2], v[6]));
  cplx a3 = cadd(v[3], v[7]), p7 = mul_w3u<INV>(csub(v[3], v[7]));
  cplx b0 = cadd(a0, a2), b2 = csub(a0, a2);
  cplx b1 = cadd(a1, a3), b3 = mul_i<INV>(csub(a1, a3));
  v[0] = cadd(b0, b1); v[4] = csub(b0, b1);
  v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
  // odd outputs: the two 1/sqrt(2) twiddles are folded into the last butterfly stage
  cplx c0 = cadd(a4, a6), c2 = csub(a4, a6);
  cplx q1 = cadd(p5, p7), q3 = mul_i<INV>(csub(p5, p7));
  v[1] = cfma_s(c0, s, q1); v[5] = cfma_s(c0, -s, q1);
  v[3] = cfma_s(c2, s, q3); v[7] = cfma_s(c2, -s, q3);
}

// Same butterfly with a per-thread sign sg = +-1 in the last stage: sg = -1 delivers the outputs
// with slots k and k^4 swapped (equivalently: the transform of the input with its odd elements
// negated).  Same instruction count -- the four plain complex add/sub pairs become FMAs.
template <bool INV> BR_HD void dft8s(cplx (&v)[8], double sg) {
  const double s = 0.70710678118654752440 * sg;
  cplx a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  cplx a1 = cadd(v[1], v[5]), p5 = mul_w1u<INV>(csub(v[1], v[5]));
  cplx a2 = cadd(v[2], v[6]), a6 = mul_i<INV>(csub(v[2], v[6]));
  cplx a3 = cadd(v[3], v[7]), p7 = mul_w3u<INV>(csub(v[3], v[7]));
  cplx b0 = cadd(a0, a2), b2 = csub(a0, a2);
  cplx b1 = cadd(a1, a3), b3 = mul_i<INV>(csub(a1, a3));
  v[0] = cfma_s(b0, sg, b1); v[4] = cfma_s(b0, -sg, b1);
  v[2] = cfma_s(b2, sg, b3); v[6] = cfma_s(b2, -sg, b3);
  cplx c0 = cadd(a4, a6), c2 = csub(a4, a6);
  cplx q1 = cadd(p5, p7), q3 = mul_i<INV>(csub(p5, p7));
  v[1] = cfma_s(c0, s, q1); v[5] = cfma_s(c0, -s, q1);
  v[3] = cfma_s(c2, s, q3); v[7] = cfma_s(c2, -s, q3);
}

// (re + i im) * omega^(64 M) and its conjugate form, with the trivial cases spelled out
// (M = 0: nothing; M = 4: one real factor) so no multiply-by-one/zero is issued
template <int M> BR_HD cplx pre_w();
template <int M> BR_HD cplx twist_in(double re, double im) {
  if (M == 0) return mk(re, im);
  if (M == 4) { const double s = 0.70710678118654752440; return mk((re - im) * s, (re + im) * s); }
  return cmul(mk(re, im), pre_w<M>());
}
template <int M> BR_HD cplx twist_out(cplx a) {  // a * conj(omega^(64 M))
  if (M == 0) return a;
  if (M == 4) { const double s = 0.70710678118654752440; return mk((a.x + a.y) * s, (a.y - a.x) * s); }
  return cmulc(a, pre_w<M>());
}

// omega^(64 m) = e^{i pi m/16}: the compile-time part of the twist
template <int M> BR_HD cplx pre_w() {
  constexpr double c[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613,
                           0.83146961230254523708, 0.70710678118654752440,
                           0.55557023301960222474, 0.38268343236508977173,
                           0.19509032201612826785};
  constexpr double s[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173,
                           0.55557023301960222474, 0.70710678118654752440,
                           0.83146961230254523708, 0.92387953251128675613,
                           0.98078528040323044913};
  return mk(c[M], s[M]);
}

// ---- rounding: Rust `x.round() as i64 as u32` (klemsa.rs:145-146) ------------
// EXACT regime (l*2^bgbit small, e.g. l=3,Bg=64): the value is an integer plus an
// error << 0.5, so ties cannot occur and round-to-nearest-even == half-away.
template <bool EXACT> BR_HD uint32_t round_torus(double y) {
#if defined(__CUDA_ARCH__)
  long long q = __double2ll_rn(y);
  if (!EXACT) {
    double d = y - __ll2double_rn(q);  // exact for |y| < 2^52, 0 beyond
    if (d == 0.5 && y > 0.0) q += 1;
    if (d == -0.5 && y < 0.0) q -= 1;
  }
  return (uint32_t)(unsigned long long)q;
#else
  double r = std::round(y);  // half away from zero
  long long q;
  if (r >= 9223372036854775807.0) q = INT64_MAX;
  else if (r <= -9223372036854775808.0) q = INT64_MIN;
  else q = (long long)r;
  return (uint32_t)(unsigned long long)q;
#endif
}

// Exact-regime conversions without the conversion pipe: a small unsigned u becomes the double
// 2^52 + u by bit pattern, one exact DADD removes the bias; y + 1.5*2^52 leaves round-to-nearest-
// even(y) mod 2^32 in the low mantissa word (|y| < 2^51; the exact regime is bounded by 2^48.6).
BR_HD double biased_to_double(uint32_t u, double bias) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(0x43300000, (int)u) - bias;
#else
  return (4503599627370496.0 + (double)u) - bias;
#endif
}
BR_HD uint32_t round_torus_magic(double y) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2loint(y + 6755399441055744.0);
#else
  return round_torus<true>(y);
#endif
}

// ---- rotate-and-subtract (trgsw.rs:212-215 + :183-186 fused) -----------------
// (X^abar * acc - acc)[j], with the reference's Torus::MAX - x wrap (trgsw.rs:318,322)
BR_HD uint32_t rot_diff(const uint32_t *accp, int j, uint32_t abar) {
  uint32_t idx = ((uint32_t)j - abar) & 2047u;
  uint32_t r = accp[idx & 1023u];
  if (idx & 1024u) r = ~r;
  return r - accp[j];
}
// (X^k * tv)[j] for k in [0, 2N] (trgsw.rs:204-207, :307-330)
BR_HD uint32_t rot_coeff(const uint32_t *tv, int j, uint32_t k) {
  uint32_t idx = ((uint32_t)j - k) & 2047u;
  uint32_t r = tv[idx & 1023u];
  return (idx & 1024u) ? ~r : r;
}

// ---- twiddles ------------------------------------------------------------------
// ta[k0] = e^{i pi r (1-4 k0)/1024}, r = tid          (pass A post- / pass A' pre-twiddle)
// tb[x]  = e^{-2 pi i (tid&7) x/64}                    (pass B post- / pass C' post-twiddle)
// tb is a geometric sequence: it can be rebuilt from tb[1], tb[2], tb[4].
BR_HD void expand_tb(cplx c1, cplx c2, cplx c4, cplx (&tb)[8]) {
  tb[0] = mk(1.0, 0.0);
  tb[1] = c1; tb[2] = c2; tb[4] = c4;
  tb[3] = cmul(c1, c2);
  tb[5] = cmul(c1, c4);
  tb[6] = cmul(c2, c4);
  tb[7] = cmul(tb[3], c4);
}

// ---- forward passes -------------------------------------------------------------
// Rotate-subtract + decomposition offset for the 16 coefficients a thread owns
// (trgsw.rs:212-215,183-186 fused with the `+ offset` of trgsw.rs:159-160).
BR_HD void load_t(int tid, const uint32_t *accp, uint32_t abar, uint32_t offset,
                  uint32_t (&t_re)[8], uint32_t (&t_im)[8]) {
#pragma unroll
  for (int m = 0; m < 8; m++) {
    int j = 64 * m + tid;
    t_re[m] = rot_diff(accp, j, abar) + offset;
    t_im[m] = rot_diff(accp, j + kHalf, abar) + offset;
  }
}

// Pass A for digits [D0, D0+ND) of one polynomial (decomposition trgsw.rs:161-167 fused
// with the twist, klemsa.rs:96-103): thread tid owns complex points 64m+tid; digit d goes
// to exchange buffer d-D0.
template <int BGBIT, int D0, int ND, bool MAGIC = false, int S = 72, int STRIDE = kExchStride>
BR_HD void fwd_pass_a(int tid, const uint32_t (&t_re)[8], const uint32_t (&t_im)[8],
                      const cplx (&ta)[8], cplx *exch) {
  constexpr uint32_t MASK = (1u << BGBIT) - 1u;
  constexpr int HALFBG = 1 << (BGBIT - 1);
  constexpr double BIAS = 4503599627370496.0 + (double)HALFBG;
#pragma unroll
  for (int d = D0; d < D0 + ND; d++) {
    const int sh = 32 - (d + 1) * BGBIT;
    cplx v[8];
#define BR_LOAD(M)                                                              \
  if (MAGIC) {                                                                  \
    v[M] = twist_in<M>(biased_to_double((t_re[M] >> sh) & MASK, BIAS),          \
                       biased_to_double((t_im[M] >> sh) & MASK, BIAS));         \
  } else {                                                                      \
    int dre = (int)((t_re[M] >> sh) & MASK) - HALFBG;                            \
    int dim = (int)((t_im[M] >> sh) & MASK) - HALFBG;                            \
    v[M] = twist_in<M>((double)dre, (double)dim);                               \
  }
    BR_LOAD(0) BR_LOAD(1) BR_LOAD(2) BR_LOAD(3) BR_LOAD(4) BR_LOAD(5) BR_LOAD(6) BR_LOAD(7)
#undef BR_LOAD
    dft8<false>(v);
    cplx *e = exch + (d - D0) * STRIDE + tid + (tid >> 3);
#pragma unroll
    for (int k0 = 0; k0 < 8; k0++) e[k0 * S] = cmul(v[k0], ta[k0]);
  }
}

// Pass A on full 32-bit torus coefficients (key generation: trlwe.rs:91-96 / klemsa.rs:88-117
// applied to a TRLWE row instead of a digit polynomial).  x_re[m] = coefficient 64m+tid,
// x_im[m] = coefficient 64m+tid+512, read as i32 (klemsa.rs:97-98).
BR_HD void fwd_pass_a_i32(int tid, const uint32_t (&x_re)[8], const uint32_t (&x_im)[8],
                          const cplx (&ta)[8], cplx *exch_buf) {
  cplx v[8];
#define BR_LOAD(M) v[M] = twist_in<M>((double)(int32_t)x_re[M], (double)(int32_t)x_im[M]);
  BR_LOAD(0) BR_LOAD(1) BR_LOAD(2) BR_LOAD(3) BR_LOAD(4) BR_LOAD(5) BR_LOAD(6) BR_LOAD(7)
#undef BR_LOAD
  dft8<false>(v);
  cplx *e = exch_buf + tid + (tid >> 3);
#pragma unroll
  for (int k0 = 0; k0 < 8; k0++) e[k0 * 72] = cmul(v[k0], ta[k0]);
}

// Pass C alone: thread v = (k0, k1) ends with bins k0+8*k1+64*k2 in out[k2].
BR_HD void fwd_pass_c(int tid, const cplx *exch_d, cplx (&out)[8]) {
  const cplx *e = exch_d + tid * 9;
#pragma unroll
  for (int j0 = 0; j0 < 8; j0++) out[j0] = e[j0];
  dft8<false>(out);
}

// Same for ONE digit chosen at run time (keeps the V4 kernel's sub-round loop rolled).
template <int BGBIT>
BR_HD void fwd_pass_a_rt(int tid, const uint32_t (&t_re)[8], const uint32_t (&t_im)[8],
                         const cplx (&ta)[8], cplx *exch_buf, int d) {
  constexpr uint32_t MASK = (1u << BGBIT) - 1u;
  constexpr int HALFBG = 1 << (BGBIT - 1);
  const int sh = 32 - (d + 1) * BGBIT;
  cplx v[8];
#define BR_LOAD(M)                                                              \
  {                                                                             \
    int dre = (int)((t_re[M] >> sh) & MASK) - HALFBG;                            \
    int dim = (int)((t_im[M] >> sh) & MASK) - HALFBG;                            \
    v[M] = twist_in<M>((double)dre, (double)dim);                               \
  }
  BR_LOAD(0) BR_LOAD(1) BR_LOAD(2) BR_LOAD(3) BR_LOAD(4) BR_LOAD(5) BR_LOAD(6) BR_LOAD(7)
#undef BR_LOAD
  dft8<false>(v);
  cplx *e = exch_buf + tid + (tid >> 3);
#pragma unroll
  for (int k0 = 0; k0 < 8; k0++) e[k0 * 72] = cmul(v[k0], ta[k0]);
}

// Pass B: thread u = (k0, j0) transforms over j1, twiddle e^{-2 pi i j0 k1/64}.
template <int NB> BR_HD void fwd_pass_b(int tid, const cplx (&tb)[8], cplx *exch) {
  const int k0 = tid >> 3, j0 = tid & 7;
#pragma unroll
  for (int d = 0; d < NB; d++) {
    cplx *e = exch + d * kExchStride + k0 * 72 + j0;
    cplx v[8];
#pragma unroll
    for (int j1 = 0; j1 < 8; j1++) v[j1] = e[j1 * 9];
    dft8<false>(v);
    e[0] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < 8; k1++) e[k1 * 9] = cmul(v[k1], tb[k1]);
  }
}

// Pass C for one digit + pointwise MAC against one BSK row (trgsw.rs:118-142):
// thread v = (k0, k1) ends with bins k0+8*k1+64*k2 in registers and accumulates
// both output spectra.  bsk_row is one device-order row: [k2][o][v].
BR_HD void fwd_pass_c_mac(int tid, const cplx *exch_d, const cplx *bsk_row, cplx (&acc)[2][8]) {
  const cplx *e = exch_d + tid * 9;
  cplx v[8];
#pragma unroll
  for (int j0 = 0; j0 < 8; j0++) v[j0] = e[j0];
  dft8<false>(v);
#pragma unroll
  for (int k2 = 0; k2 < 8; k2++) {
    cfma(acc[0][k2], v[k2], bsk_row[(k2 * 2 + 0) * 64 + tid]);
    cfma(acc[1][k2], v[k2], bsk_row[(k2 * 2 + 1) * 64 + tid]);
  }
}

// ---- inverse passes ---------------------------------------------------------------
BR_HD void inv_pass_c(int tid, const cplx (&tb)[8], cplx (&acc)[2][8], cplx *exch) {
#pragma unroll
  for (int o = 0; o < 2; o++) {
    dft8<true>(acc[o]);
    cplx *e = exch + o * kExchStride + tid * 9;
    e[0] = acc[o][0];
#pragma unroll
    for (int j0 = 1; j0 < 8; j0++) e[j0] = cmulc(acc[o][j0], tb[j0]);
  }
}

BR_HD void inv_pass_b(int tid, cplx *exch) {
  const int k0 = tid >> 3, j0 = tid & 7;
#pragma unroll
  for (int o = 0; o < 2; o++) {
    cplx *e = exch + o * kExchStride + k0 * 72 + j0;
    cplx v[8];
#pragma unroll
    for (int k1 = 0; k1 < 8; k1++) v[k1] = e[k1 * 9];
    dft8<true>(v);
#pragma unroll
    for (int j1 = 0; j1 < 8; j1++) e[j1 * 9] = v[j1];
  }
}

// Pass A' + untwist + torus rounding (klemsa.rs:136-147) + accumulator update
// (trgsw.rs:190-193).  acc points at this ciphertext's u32[2][1024].
template <bool EXACT, bool MAGIC = false, int S = 72, int STRIDE = kExchStride>
BR_HD void inv_pass_a(int tid, const cplx (&ta)[8], const cplx *exch, uint32_t *acc) {
#pragma unroll
  for (int o = 0; o < 2; o++) {
    const cplx *e = exch + o * STRIDE + tid + (tid >> 3);
    cplx v[8];
#pragma unroll
    for (int k0 = 0; k0 < 8; k0++) v[k0] = cmulc(e[k0 * S], ta[k0]);
    dft8<true>(v);
    uint32_t *ap = acc + o * kN;
#define BR_STORE(M)                                                       \
  {                                                                       \
    cplx y = twist_out<M>(v[M]);                                          \
    if (MAGIC) {                                                          \
      ap[64 * M + tid] += round_torus_magic(y.x);                         \
      ap[64 * M + tid + kHalf] += round_torus_magic(y.y);                 \
    } else {                                                              \
      ap[64 * M + tid] += round_torus<EXACT>(y.x);                        \
      ap[64 * M + tid + kHalf] += round_torus<EXACT>(y.y);                \
    }                                                                     \
  }
    BR_STORE(0) BR_STORE(1) BR_STORE(2) BR_STORE(3) BR_STORE(4) BR_STORE(5) BR_STORE(6) BR_STORE(7)
#undef BR_STORE
  }
}

// ---- key layout -----------------------------------------------------------------
// Device BSK: cplx[n][2l][8 (k2)][2 (o: a,b)][64 (v)], value = reference value / 1024.
// Reference: f64[n][2l][2 (o)][1024], re[0..512) | im[0..512), natural bin order.
BR_HD size_t bsk_dev_index(int i, int r, int l2, int k2, int o, int v) {
  return ((((size_t)i * l2 + r) * 8 + k2) * 2 + o) * 64 + v;
}
BR_HD int bin_of(int v, int k2) { return (v >> 3) + 8 * (v & 7) + 64 * k2; }

}  // namespace br
