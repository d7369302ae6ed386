// brs_core.cuh -- per-thread building blocks of the 128-thread blind-rotation kernel
// (blind_rotate_s.cu).  Everything is __host__ __device__: the kernel strings the phases together
// with named barriers, tensor-memory exchanges and a TMA ring; emu.cpp runs the very same functions
// thread by thread on the CPU (tensor memory modelled as an array) so the index, twiddle and layout
// logic is checked against the oracle without a GPU (tests/test_emulator.py).
//
// One ciphertext is owned by 128 threads (4 warps on 4 different SM sub-partitions), 4 complex
// points per thread.  The negacyclic transform (reference: src/fft/klemsa.rs:88-150) is a twisted
// 512-point complex DFT; with w(x) = e^{i pi x/1024}:
//   forward   X_k = sum_j z_j w(j (1 - 4k)),      z_j = x_j + i x_{j+512}
//   inverse   y_j = w(-j) sum_k S_k w(4 j k),     coefficient j = round(Re y_j), j+512 = round(Im y_j)
// Both run most-significant-digit first,
//   forward  j = 128a + 16b + 4c + d  ->  k = ka + 4kb + 32kc + 128kd   radices 4, 8, 4, 4  (passes A B C D)
//   inverse  k = (ka + 4s) + 8m + 32kc + 128kd (kb = s + 2m)  ->  j = j0 + 4j1 + 16j2 + 64j3
//                                                                 radices 4, 4, 4, 8  (passes D' C' B' A')
// so every pass evaluates a short polynomial  y_kappa = sum_x u_x (rho W_R^kappa)^x  whose ratio rho
// carries the twist AND the inter-pass twiddles of the thread: there is no separate twiddle
// multiplication anywhere.  Each radix-2 butterfly (a + r b, a - r b) costs 6 FMAs (4 for the sum,
// 2 for 2a - sum).  The radix-8 passes (B after the shared-memory exchange, A' before the accumulator
// update) are split over a pair of lanes: both lanes read all 8 inputs (same addresses, a shared-memory
// broadcast) and each evaluates the half of the outputs whose lowest index bit is its sigma.
// FP64 instructions per CMUX at l = 3: 149 504 per ciphertext (the 64-thread radix-8 kernel: 159 000).
//
// All power-of-two scale factors of the reference (x2 klemsa.rs:112, x0.5 trgsw.rs:137, x0.5 and
// 1/512 klemsa.rs:126,136) are folded into the uploaded key (exact).
#pragma once
#include <stdint.h>

#include <cmath>

#include "br_core.cuh"

namespace brs {

using br::mk;
using br::kN;
using br::kHalf;

constexpr int kT = 128;              // threads per ciphertext
constexpr int kRowCplx = 4 * 2 * kT; // one key row in kernel order: [kd][o][T]
constexpr int kInvPitch = 65;        // complex per row of the inverse exchange buffer (bank spread)
constexpr int kTwPerThread = 20;     // complex constants per thread (layout below)

// ---- butterflies ---------------------------------------------------------------------------------
BR_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
// a + r*b
BR_HD cplx bfp(cplx a, cplx b, cplx r) {
  return mk(fma_(-r.y, b.y, fma_(r.x, b.x, a.x)), fma_(r.y, b.x, fma_(r.x, b.y, a.y)));
}
// (a + r*b, a - r*b)
BR_HD void bf(cplx a, cplx b, cplx r, cplx &p, cplx &m) {
  p = bfp(a, b, r);
  m = mk(fma_(2.0, a.x, -p.x), fma_(2.0, a.y, -p.y));
}
template <bool INV> BR_HD cplx rot_i(cplx r) {  // r * (-i) forward, r * (+i) inverse
  return INV ? mk(-r.y, r.x) : mk(r.y, -r.x);
}

// radix-4: y_kappa = sum_x u_x (rho W4^kappa)^x, W4 = -i (forward) / +i (inverse); r1 = rho, r2 = rho^2
template <bool INV> BR_HD void r4(cplx (&u)[4], cplx r1, cplx r2) {
  cplx ep, em, fp, fm;
  bf(u[0], u[2], r2, ep, em);
  bf(u[1], u[3], r2, fp, fm);
  bf(ep, fp, r1, u[0], u[2]);
  bf(em, fm, rot_i<INV>(r1), u[1], u[3]);
}
// plain radix-4 (rho = 1)
template <bool INV> BR_HD void r4_plain(cplx (&u)[4]) {
  cplx ep = br::cadd(u[0], u[2]), em = br::csub(u[0], u[2]);
  cplx fp = br::cadd(u[1], u[3]), fm = rot_i<INV>(br::csub(u[1], u[3]));
  u[0] = br::cadd(ep, fp); u[2] = br::csub(ep, fp);
  u[1] = br::cadd(em, fm); u[3] = br::csub(em, fm);
}
// half of a radix-8: kappa = sigma + 2 s0 + 4 s1, output slot s = s0 + 2 s1.  The lane's sigma is
// folded into its constants: c1 = (-1)^sigma rho^4, c2 = rho^2 W4^sigma, c3 = rho W8^sigma.
template <bool INV> BR_HD void r8_half(const cplx (&u)[8], cplx c1, cplx c2, cplx c3, cplx (&y)[4]) {
  cplx s0 = bfp(u[0], u[4], c1), s1 = bfp(u[1], u[5], c1);
  cplx s2 = bfp(u[2], u[6], c1), s3 = bfp(u[3], u[7], c1);
  cplx r0p, r0m, r1p, r1m;
  bf(s0, s2, c2, r0p, r0m);
  bf(s1, s3, c2, r1p, r1m);
  bf(r0p, r1p, c3, y[0], y[2]);
  bf(r0m, r1m, rot_i<INV>(c3), y[1], y[3]);
}

// ---- thread <-> index maps (T = 32 W + lane; W = warp within the ciphertext group) ----------------
// pass A   : T = j' = 16b + 4c + d                       registers a      -> ka
// pass B   : W = ka, lane = (c1 c0 d1 d0 s)              8 inputs b       -> slot sB (kb = s + 2 sB)
// pass C   : W = ka, lane = (d1 d0 s sB1 sB0)            registers c      -> kc
// pass D   : W = ka, lane = (s sB1 sB0 kc1 kc0)          registers d      -> kd      (MAC, pass D')
// pass C'  : W = ka, lane = (j01 j00 s sB1 sB0)          registers kc     -> j1
// pass B'  : W = ka, lane = (j11 j10 j01 j00 s)          registers sB (m) -> j2
// pass A'  : T = 2 jlow + s', jlow = j0 + 4 j1 + 16 j2   8 inputs ka + 4s -> slot n (j3 = s' + 2n)
BR_HD int bin_of(int T, int kd) {
  return (T >> 5) + 4 * (((T >> 4) & 1) + 2 * ((T >> 2) & 3)) + 32 * (T & 3) + 128 * kd;
}
BR_HD size_t row_index(int kd, int o, int T) { return (size_t)(kd * 2 + o) * kT + T; }

// ---- per-thread constants --------------------------------------------------------------------------
// [0..2] pass B c1,c2,c3   [3] spare   [4,5] pass C rho, rho^2   [6,7] pass D rho, rho^2
// [8,9] pass C' rho, rho^2 [10,11] pass B' rho, rho^2   [12..14] pass A' c1,c2,c3   [15] spare
// [16..19] untwist w(-(jlow + 64 (s' + 2n))), n = 0..3
enum { TW_B = 0, TW_C = 4, TW_D = 6, TW_CI = 8, TW_BI = 10, TW_AI = 12, TW_UT = 16 };

inline cplx omega(long x) {   // e^{i pi x/1024}, exact quadrant reduction
  long r = ((x % 2048) + 2048) % 2048;
  long q = r / 512, f = r % 512;           // angle = q*pi/2 + f*pi/1024
  double a = M_PI * (double)f / 1024.0;
  double c = std::cos(a), s = std::sin(a);
  if (f == 0) { c = 1.0; s = 0.0; }
  if (f == 256) { c = s = 0.70710678118654752440; }
  switch (q) {
    case 0: return mk(c, s);
    case 1: return mk(-s, c);
    case 2: return mk(-c, -s);
    default: return mk(s, -c);
  }
}
inline void make_tw(int T, cplx (&tw)[kTwPerThread]) {
  const int W = T >> 5, lane = T & 31;
  for (auto &t : tw) t = mk(0.0, 0.0);
  {  // pass B: ka = W, sigma = lane & 1
    const long ka = W, s = lane & 1;
    cplx c1 = omega(64 * (1 - 4 * ka));
    if (s) c1 = mk(-c1.x, -c1.y);
    tw[TW_B + 0] = c1;
    tw[TW_B + 1] = omega(32 * (1 - 4 * ka) - 512 * s);
    tw[TW_B + 2] = omega(16 * (1 - 4 * ka) - 256 * s);
  }
  {  // pass C: lane = (d1 d0 s sB1 sB0)
    const long ka = W, kb = ((lane >> 2) & 1) + 2 * (lane & 3);
    const long e = 4 * (1 - 4 * ka - 16 * kb);
    tw[TW_C + 0] = omega(e); tw[TW_C + 1] = omega(2 * e);
  }
  {  // pass D: lane = (s sB1 sB0 kc1 kc0)
    const long ka = W, kb = ((lane >> 4) & 1) + 2 * ((lane >> 2) & 3), kc = lane & 3;
    const long e = 1 - 4 * (ka + 4 * kb + 32 * kc);
    tw[TW_D + 0] = omega(e); tw[TW_D + 1] = omega(2 * e);
  }
  {  // pass C': j0 = lane >> 3
    const long j0 = lane >> 3;
    tw[TW_CI + 0] = omega(128 * j0); tw[TW_CI + 1] = omega(256 * j0);
  }
  {  // pass B': lane = (j11 j10 j01 j00 s)
    const long j1 = lane >> 3, j0 = (lane >> 1) & 3;
    const long e = 32 * (j0 + 4 * j1);
    tw[TW_BI + 0] = omega(e); tw[TW_BI + 1] = omega(2 * e);
  }
  {  // pass A': T = 2 jlow + s'
    const long jl = T >> 1, s = T & 1;
    cplx c1 = omega(16 * jl);
    if (s) c1 = mk(-c1.x, -c1.y);
    tw[TW_AI + 0] = c1;
    tw[TW_AI + 1] = omega(8 * jl + 512 * s);
    tw[TW_AI + 2] = omega(4 * jl + 256 * s);
    for (long n = 0; n < 4; n++) tw[TW_UT + n] = omega(-(jl + 64 * (s + 2 * n)));
  }
}

// ---- forward pass A -----------------------------------------------------------------------------------
// Rotate-subtract + decomposition offset for the 8 coefficients a thread owns
// (trgsw.rs:212-215,183-186 fused with the `+ offset` of trgsw.rs:159-160).
BR_HD void load_t(int T, const uint32_t *accp, uint32_t abar, uint32_t offset, uint32_t (&t_re)[4],
                  uint32_t (&t_im)[4]) {
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int j = 128 * a + T;
    t_re[a] = br::rot_diff(accp, j, abar) + offset;
    t_im[a] = br::rot_diff(accp, j + kHalf, abar) + offset;
  }
}

// Pass A of digit d (decomposition trgsw.rs:161-167 fused with the twist klemsa.rs:96-103): the
// twist over a is the geometric ratio e^{i pi/8}, a compile-time constant.  Output ka -> row ka.
template <int BGBIT, bool MAGIC>
BR_HD void fwd_pass_a(int T, int d, const uint32_t (&t_re)[4], const uint32_t (&t_im)[4], cplx *exch_d) {
  constexpr uint32_t MASK = (1u << BGBIT) - 1u;
  constexpr int HALFBG = 1 << (BGBIT - 1);
  constexpr double BIAS = 4503599627370496.0 + (double)HALFBG;
  const int sh = 32 - (d + 1) * BGBIT;
  cplx u[4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    if (MAGIC) {
      u[a] = mk(br::biased_to_double((t_re[a] >> sh) & MASK, BIAS),
                br::biased_to_double((t_im[a] >> sh) & MASK, BIAS));
    } else {
      u[a] = mk((double)((int)((t_re[a] >> sh) & MASK) - HALFBG),
                (double)((int)((t_im[a] >> sh) & MASK) - HALFBG));
    }
  }
  const cplx r1 = mk(0.92387953251128675613, 0.38268343236508977173);   // e^{i pi/8}
  const cplx r2 = mk(0.70710678118654752440, 0.70710678118654752440);   // e^{i pi/4}
  r4<false>(u, r1, r2);
#pragma unroll
  for (int ka = 0; ka < 4; ka++) exch_d[ka * kT + T] = u[ka];
}
// Pass A on full 32-bit torus coefficients read as i32 (key generation: trlwe.rs:91-96)
BR_HD void fwd_pass_a_i32(int T, const uint32_t (&x_re)[4], const uint32_t (&x_im)[4], cplx *exch_d) {
  cplx u[4];
#pragma unroll
  for (int a = 0; a < 4; a++) u[a] = mk((double)(int32_t)x_re[a], (double)(int32_t)x_im[a]);
  const cplx r1 = mk(0.92387953251128675613, 0.38268343236508977173);
  const cplx r2 = mk(0.70710678118654752440, 0.70710678118654752440);
  r4<false>(u, r1, r2);
#pragma unroll
  for (int ka = 0; ka < 4; ka++) exch_d[ka * kT + T] = u[ka];
}

// Pass B: reads its 8 inputs (b) from the exchange buffer; both lanes of a sigma pair read the same
// addresses.  Output slot sB <-> kb = sigma + 2 sB.
BR_HD void fwd_pass_b(int T, const cplx *exch_d, cplx c1, cplx c2, cplx c3, cplx (&y)[4]) {
  const int W = T >> 5, lane = T & 31;
  const cplx *e = exch_d + W * kT + (lane >> 1);
  cplx u[8];
#pragma unroll
  for (int b = 0; b < 8; b++) u[b] = e[16 * b];
  r8_half<false>(u, c1, c2, c3, y);
}

// ---- inverse -----------------------------------------------------------------------------------------
// Pass B' output (j2 in registers) -> inverse exchange buffer row ka + 4 s, column jlow
BR_HD void inv_store_b(int T, const cplx (&u)[4], cplx *exch_o) {
  const int W = T >> 5, lane = T & 31;
  const int j1 = lane >> 3, j0 = (lane >> 1) & 3, s = lane & 1;
  cplx *e = exch_o + (W + 4 * s) * kInvPitch + j0 + 4 * j1;
#pragma unroll
  for (int j2 = 0; j2 < 4; j2++) e[16 * j2] = u[j2];
}
// Pass A' + untwist + torus rounding (klemsa.rs:136-147) + accumulator update (trgsw.rs:190-193)
template <bool EXACT, bool MAGIC>
BR_HD void inv_pass_a(int T, const cplx *exch_o, cplx c1, cplx c2, cplx c3, const cplx (&ut)[4],
                      uint32_t *acc_o) {
  const int jl = T >> 1, s = T & 1;
  const cplx *e = exch_o + jl;
  cplx u[8], y[4];
#pragma unroll
  for (int k = 0; k < 8; k++) u[k] = e[k * kInvPitch];
  r8_half<true>(u, c1, c2, c3, y);
#pragma unroll
  for (int n = 0; n < 4; n++) {
    const cplx z = br::cmul(y[n], ut[n]);
    const int j = jl + 64 * (s + 2 * n);
    if (MAGIC) {
      acc_o[j] += br::round_torus_magic(z.x);
      acc_o[j + kHalf] += br::round_torus_magic(z.y);
    } else {
      acc_o[j] += br::round_torus<EXACT>(z.x);
      acc_o[j + kHalf] += br::round_torus<EXACT>(z.y);
    }
  }
}

}  // namespace brs
