/*
 * tfhe_oracle.c -- see tfhe_oracle.h.  TEST INFRASTRUCTURE ONLY (parity checker
 * + CPU baseline).  Plain C11 + OpenMP; no dependency on the product.
 *
 * All citations are file:line under /root/reference.
 */
#include "tfhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define N ORC_N
#define N2 (ORC_N / 2)

/* ------------------------------------------------------------------ params */

typedef struct { const char *name; orc_params p; } named_params;

/* src/params.rs:91-404 (values only; the reference executes just the 128-bit
 * set, params.rs:426-469 -- the others are "same algorithm, other constants"). */
static const named_params k_sets[] = {
    {"80", {550, 1024, 3, 6, 2, 7, 5.0e-5, 3.73e-8}},
    {"110", {630, 1024, 3, 6, 2, 8, 3.0517578125e-05, 2.9802322387695313e-8}},
    {"128", {700, 1024, 3, 6, 2, 9, 2.0e-5, 2.0e-8}},
    {"uint1", {700, 1024, 2, 10, 2, 8, 2.0e-05, 2.0e-08}},
    {"uint2", {687, 1024, 1, 18, 4, 3, 0.00002120846893069972, 0.0000000000023184122752704995}},
    {"uint3", {820, 1024, 1, 23, 6, 2, 0.0000025167616095979554, 0.0000000000000002220446049250313}},
    {"uint4", {820, 1024, 1, 22, 5, 3, 0.0000025167616095979554, 0.0000000000000002220446049250313}},
    {"uint5", {1071, 1024, 1, 22, 6, 3, 7.08822676541043e-8, 2.2204460492503131e-17}},
    {"uint6", {1071, 1024, 1, 22, 6, 3, 7.08822676541043e-8, 2.2204460492503131e-17}},
    {"uint7", {1160, 1024, 1, 22, 7, 3, 1.9662200074984027e-8, 2.2204460492503131e-17}},
    {"uint8", {1160, 1024, 1, 22, 7, 3, 1.9662200074984027e-8, 2.2204460492503131e-17}},
};

int orc_params_by_name(const char *name, orc_params *out) {
  for (size_t i = 0; i < sizeof(k_sets) / sizeof(k_sets[0]); i++) {
    if (strcmp(name, k_sets[i].name) == 0) {
      *out = k_sets[i].p;
      return 0;
    }
  }
  return -1;
}

size_t orc_ksk_words(const orc_params *p) {
  return (size_t)N * p->iks_t * ((size_t)1 << p->basebit) * (p->n + 1);
}
size_t orc_bsk_doubles(const orc_params *p) {
  return (size_t)p->n * 2 * p->l * 2 * N;
}

/* --------------------------------------------------------------------- RNG */

static uint64_t splitmix64(uint64_t *x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

void orc_rng_seed(orc_rng *r, uint64_t seed) {
  uint64_t x = seed;
  for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&x);
  r->has_spare = 0;
  r->spare = 0.0;
}
static uint64_t rng_u64(orc_rng *r) { /* xoshiro256** */
  uint64_t *s = r->s;
  uint64_t result = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t; s[3] = rotl64(s[3], 45);
  return result;
}
uint32_t orc_rng_u32(orc_rng *r) { return (uint32_t)(rng_u64(r) >> 32); }
static double rng_unit(orc_rng *r) { return ((rng_u64(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
double orc_rng_normal(orc_rng *r, double sigma) { /* Box-Muller */
  if (r->has_spare) { r->has_spare = 0; return r->spare * sigma; }
  double u = rng_unit(r), v = rng_unit(r);
  double m = sqrt(-2.0 * log(u));
  r->spare = m * sin(6.283185307179586476925286766559 * v);
  r->has_spare = 1;
  return m * cos(6.283185307179586476925286766559 * v) * sigma;
}

/* ----------------------------------------------------------------- scalars */

/* utils.rs:9-12: (d % 1.0) * 2^32, cast to i64 (truncating), wrap to u32 */
uint32_t orc_f64_to_torus(double d) {
  double t = fmod(d, 1.0) * 4294967296.0;
  return (uint32_t)(int64_t)t;
}
double orc_torus_to_f64(uint32_t t) { return (double)t / 4294967296.0; }

/* key.rs:78-89 */
uint32_t orc_decomposition_offset(const orc_params *p) {
  uint32_t offset = 0, bg = 1u << p->bgbit;
  for (uint32_t i = 0; i < p->l; i++)
    offset += (bg / 2) * (1u << (32 - (i + 1) * p->bgbit));
  return offset;
}
/* trgsw.rs:345 */
uint32_t orc_prec_offset(const orc_params *p) {
  return 1u << (32 - (1 + p->basebit * p->iks_t));
}

/* --------------------------------------------------------------------- FFT */

/* 512-point complex FFT standing in for rustfft (klemsa.rs:62-65): iterative
 * radix-2 decimation-in-time, split re/im, exact-table twiddles. */
static double g_cos[N2 / 2], g_sin[N2 / 2]; /* e^{-2 pi i k/512}, k<256 */
static double g_tw_re[N2], g_tw_im[N2];     /* klemsa.rs:49-58: e^{i pi k/N} */
static uint16_t g_rev[N2];

__attribute__((constructor)) static void orc_init_tables(void) {
  for (int k = 0; k < N2 / 2; k++) {
    double a = 2.0 * M_PI * (double)k / (double)N2;
    g_cos[k] = cos(a);
    g_sin[k] = -sin(a);
  }
  double unit = M_PI / (double)N; /* klemsa.rs:52 */
  for (int i = 0; i < N2; i++) {
    double a = (double)i * unit;
    g_tw_re[i] = cos(a);
    g_tw_im[i] = sin(a);
  }
  for (int i = 0; i < N2; i++) {
    int r = 0;
    for (int b = 0; b < 9; b++) r |= ((i >> b) & 1) << (8 - b);
    g_rev[i] = (uint16_t)r;
  }
}

/* sign=-1 forward (e^{-2 pi i jk/512}), +1 inverse; unnormalised, in place */
static void fft512(double *re, double *im, int sign) {
  for (int i = 0; i < N2; i++) {
    int r = g_rev[i];
    if (r > i) {
      double t = re[i]; re[i] = re[r]; re[r] = t;
      t = im[i]; im[i] = im[r]; im[r] = t;
    }
  }
  for (int half = 1; half < N2; half <<= 1) {
    int step = N2 / (2 * half);
    for (int base = 0; base < N2; base += 2 * half) {
      for (int k = 0; k < half; k++) {
        double wr = g_cos[k * step];
        double wi = sign < 0 ? g_sin[k * step] : -g_sin[k * step];
        int i0 = base + k, i1 = i0 + half;
        double xr = re[i1] * wr - im[i1] * wi;
        double xi = re[i1] * wi + im[i1] * wr;
        re[i1] = re[i0] - xr; im[i1] = im[i0] - xi;
        re[i0] += xr; im[i0] += xi;
      }
    }
  }
}

/* klemsa.rs:88-117: torus -> Fourier ("ifft" in the reference's naming) */
void orc_ifft(const uint32_t *in, double *out) {
  double re[N2], im[N2];
  for (int i = 0; i < N2; i++) {
    double in_re = (double)(int32_t)in[i];
    double in_im = (double)(int32_t)in[i + N2];
    re[i] = in_re * g_tw_re[i] - in_im * g_tw_im[i];
    im[i] = in_re * g_tw_im[i] + in_im * g_tw_re[i];
  }
  fft512(re, im, -1);
  for (int i = 0; i < N2; i++) {
    out[i] = re[i] * 2.0;
    out[i + N2] = im[i] * 2.0;
  }
}

/* Rust `f64::round` = half away from zero; `as i64` saturates; `as u32` wraps */
static inline uint32_t round_to_torus(double v, double *max_frac) {
  double r = round(v);
  if (max_frac) {
    double f = fabs(v - r);
    if (f > *max_frac) *max_frac = f;
  }
  int64_t q;
  if (r >= 9223372036854775807.0) q = INT64_MAX;
  else if (r <= -9223372036854775808.0) q = INT64_MIN;
  else q = (int64_t)r;
  return (uint32_t)q;
}

/* klemsa.rs:119-150: Fourier -> torus ("fft" in the reference's naming) */
static void fft_impl(const double *in, uint32_t *out, double *max_frac) {
  double re[N2], im[N2];
  for (int i = 0; i < N2; i++) {
    re[i] = in[i] * 0.5;
    im[i] = in[i + N2] * 0.5;
  }
  fft512(re, im, +1);
  const double norm = 1.0 / (double)N2;
  for (int i = 0; i < N2; i++) {
    double w_re = g_tw_re[i], w_im = g_tw_im[i];
    double tmp_re = (re[i] * w_re + im[i] * w_im) * norm;
    double tmp_im = (im[i] * w_re - re[i] * w_im) * norm;
    out[i] = round_to_torus(tmp_re, max_frac);
    out[i + N2] = round_to_torus(tmp_im, max_frac);
  }
}
void orc_fft(const double *in, uint32_t *out) { fft_impl(in, out, NULL); }

/* klemsa.rs:152-174 */
void orc_poly_mul(const uint32_t *a, const uint32_t *b, uint32_t *out) {
  double fa[N], fb[N], fr[N];
  orc_ifft(a, fa);
  orc_ifft(b, fb);
  for (int i = 0; i < N2; i++) {
    double ar = fa[i], ai = fa[i + N2], br = fb[i], bi = fb[i + N2];
    fr[i] = (ar * br - ai * bi) * 0.5;
    fr[i + N2] = (ar * bi + ai * br) * 0.5;
  }
  orc_fft(fr, out);
}

/* fft/mod.rs:240-255 (the reference's own schoolbook oracle) */
void orc_poly_mul_exact(const uint32_t *a, const uint32_t *b, uint32_t *out) {
  uint32_t res[N];
  memset(res, 0, sizeof(res));
  for (int i = 0; i < N; i++) {
    uint32_t ai = a[i];
    if (ai == 0) continue;
    for (int j = 0; j < N - i; j++) res[i + j] += ai * b[j];
    for (int j = N - i; j < N; j++) res[i + j - N] -= ai * b[j];
  }
  memcpy(out, res, sizeof(res));
}

/* -------------------------------------------------------------------- keys */

/* key.rs:33-48: uniform binary keys */
void orc_secret_key(const orc_params *p, uint64_t seed, uint32_t *s0, uint32_t *s1) {
  orc_rng r;
  orc_rng_seed(&r, seed);
  for (uint32_t i = 0; i < p->n; i++) s0[i] = orc_rng_u32(&r) & 1u;
  for (uint32_t i = 0; i < N; i++) s1[i] = orc_rng_u32(&r) & 1u;
}

/* key.rs:91-100 */
void orc_gen_testvec(uint32_t *a, uint32_t *b) {
  uint32_t bt = orc_f64_to_torus(0.125);
  for (int i = 0; i < N; i++) { a[i] = 0; b[i] = bt; }
}

/* tlwe.rs:37-53 + utils.rs:22-38 */
void orc_lwe_encrypt_f64(const orc_params *p, double mu, double alpha,
                         const uint32_t *s0, orc_rng *rng, uint32_t *ct) {
  uint32_t inner = 0;
  for (uint32_t i = 0; i < p->n; i++) {
    uint32_t r = orc_rng_u32(rng);
    inner += s0[i] * r;
    ct[i] = r;
  }
  uint32_t b = orc_f64_to_torus(orc_rng_normal(rng, alpha)) + orc_f64_to_torus(mu);
  ct[p->n] = inner + b;
}
/* tlwe.rs:55-58 */
void orc_lwe_encrypt_bool(const orc_params *p, int bit, const uint32_t *s0,
                          orc_rng *rng, uint32_t *ct) {
  orc_lwe_encrypt_f64(p, bit ? 0.125 : -0.125, p->alpha_lv0, s0, rng, ct);
}
/* tlwe.rs:84-100 */
void orc_lwe_encrypt_message(const orc_params *p, uint32_t msg, uint32_t modulus,
                             const uint32_t *s0, orc_rng *rng, uint32_t *ct) {
  msg %= modulus;
  double scale = 1.0 / (2.0 * (double)modulus);
  orc_lwe_encrypt_f64(p, (double)msg * scale, p->alpha_lv0, s0, rng, ct);
}
uint32_t orc_lwe_phase(const uint32_t *ct, const uint32_t *key, uint32_t n) {
  uint32_t inner = 0;
  for (uint32_t i = 0; i < n; i++) inner += ct[i] * key[i];
  return ct[n] - inner;
}
/* tlwe.rs:60-68 */
int orc_lwe_decrypt_bool(const uint32_t *ct, const uint32_t *key, uint32_t n) {
  return (int32_t)orc_lwe_phase(ct, key, n) >= 0;
}
/* tlwe.rs:111-126 */
uint32_t orc_lwe_decrypt_message(const uint32_t *ct, const uint32_t *key,
                                 uint32_t n, uint32_t modulus) {
  double f = orc_torus_to_f64(orc_lwe_phase(ct, key, n));
  double scale = 1.0 / (2.0 * (double)modulus);
  uint64_t m = (uint64_t)(f / scale + 0.5);
  return (uint32_t)(m % modulus);
}

void orc_lwe_encrypt_batch(const orc_params *p, const double *mu, size_t count, double alpha,
                           const uint32_t *s0, uint64_t seed, uint32_t *cts) {
  const size_t w = p->n + 1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)count; i++) {
    orc_rng r;
    orc_rng_seed(&r, seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
    orc_lwe_encrypt_f64(p, mu[i], alpha, s0, &r, cts + (size_t)i * w);
  }
}
void orc_lwe_phase_batch(const uint32_t *cts, size_t count, const uint32_t *key, uint32_t n,
                         uint32_t *phases) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)count; i++) phases[i] = orc_lwe_phase(cts + (size_t)i * (n + 1), key, n);
}

/* key.rs:102-122.  One RNG stream per i so the result is thread-count free. */
void orc_gen_ksk(const orc_params *p, const uint32_t *s0, const uint32_t *s1,
                 uint64_t seed, uint32_t *ksk) {
  const uint32_t base = 1u << p->basebit, t = p->iks_t, w = p->n + 1;
  memset(ksk, 0, orc_ksk_words(p) * sizeof(uint32_t));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; i++) {
    orc_rng r;
    orc_rng_seed(&r, seed ^ (0x4B534B00ull + (uint64_t)i * 0x9E3779B97F4A7C15ull));
    for (uint32_t j = 0; j < t; j++) {
      for (uint32_t k = 1; k < base; k++) {
        double mu = (double)(k * s1[i]) / (double)((uint64_t)1 << ((j + 1) * p->basebit));
        size_t idx = ((size_t)base * t * i) + (size_t)base * j + k;
        orc_lwe_encrypt_f64(p, mu, p->alpha_lv0, s0, &r, ksk + idx * w);
      }
    }
  }
}

/* trlwe.rs:30-52 with p == 0 */
static void trlwe_encrypt_zero(double alpha, const uint32_t *s1, orc_rng *r,
                               uint32_t *a, uint32_t *b) {
  uint32_t as1[N];
  for (int i = 0; i < N; i++) a[i] = orc_rng_u32(r);
  for (int i = 0; i < N; i++) b[i] = orc_f64_to_torus(orc_rng_normal(r, alpha)) + orc_f64_to_torus(0.0);
  orc_poly_mul(a, s1, as1);
  for (int i = 0; i < N; i++) b[i] += as1[i];
}

/* key.rs:128-156 + trgsw.rs:29-68 + trlwe.rs:91-96 */
void orc_gen_bsk(const orc_params *p, const uint32_t *s0, const uint32_t *s1,
                 uint64_t seed, double *bsk_fft, uint32_t *bsk_torus) {
  const uint32_t l = p->l;
  const double bg = (double)(1u << p->bgbit);
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < (int)p->n; i++) {
    orc_rng r;
    orc_rng_seed(&r, seed ^ (0x42534B00ull + (uint64_t)i * 0xD1B54A32D192ED03ull));
    uint32_t rows[2 * 8][2][N]; /* l <= 8 */
    for (uint32_t k = 0; k < 2 * l; k++)
      trlwe_encrypt_zero(p->alpha_lv1, s1, &r, rows[k][0], rows[k][1]);
    for (uint32_t k = 0; k < l; k++) {
      uint32_t pt = orc_f64_to_torus(pow(bg, -(double)(1 + k))); /* trgsw.rs:33-36 */
      rows[k][0][0] += s0[i] * pt;     /* trgsw.rs:45 */
      rows[k + l][1][0] += s0[i] * pt; /* trgsw.rs:46 */
    }
    for (uint32_t k = 0; k < 2 * l; k++) {
      size_t off = (((size_t)i * 2 * l + k) * 2) * N;
      orc_ifft(rows[k][0], bsk_fft + off);
      orc_ifft(rows[k][1], bsk_fft + off + N);
      if (bsk_torus) {
        memcpy(bsk_torus + off, rows[k][0], N * sizeof(uint32_t));
        memcpy(bsk_torus + off + N, rows[k][1], N * sizeof(uint32_t));
      }
    }
  }
}

/* ---------------------------------------------------------------- hot path */

/* gates.rs:54-150: out = ca*a + cb*b on all n+1 words, then b += offset */
static const struct { int32_t ca, cb; uint32_t off; } k_gate[10] = {
    /* NAND  */ {-1, -1, 0x20000000u}, /* gates.rs:54-58   */
    /* AND   */ {1, 1, 0xE0000000u},   /* gates.rs:70-74   */
    /* OR    */ {1, 1, 0x20000000u},   /* gates.rs:62-66   */
    /* XOR   */ {1, 2, 0x40000000u},   /* gates.rs:78-82   */
    /* XNOR  */ {1, -2, 0xC0000000u},  /* gates.rs:86-90   */
    /* NOR   */ {-1, -1, 0xE0000000u}, /* gates.rs:94-98   */
    /* ANDNY */ {-1, 1, 0xE0000000u},  /* gates.rs:102-111 */
    /* ANDYN */ {1, -1, 0xE0000000u},  /* gates.rs:115-124 */
    /* ORNY  */ {-1, 1, 0x20000000u},  /* gates.rs:128-137 */
    /* ORYN  */ {1, -1, 0x20000000u},  /* gates.rs:141-150 */
};
void orc_gate_prep(const orc_params *p, int op, const uint32_t *a,
                   const uint32_t *b, uint32_t *out) {
  uint32_t ca = (uint32_t)k_gate[op].ca, cb = (uint32_t)k_gate[op].cb;
  for (uint32_t i = 0; i <= p->n; i++) out[i] = ca * a[i] + cb * b[i];
  out[p->n] += k_gate[op].off;
}

/* trgsw.rs:307-330 (note Torus::MAX - x, i.e. ~x, not -x) */
void orc_poly_mul_with_x_k(const uint32_t *a, uint32_t k, uint32_t *out) {
  uint32_t res[N];
  memset(res, 0, sizeof(res));
  if (k < N) {
    for (uint32_t i = 0; i < N - k; i++) res[i + k] = a[i];
    for (uint32_t i = N - k; i < N; i++) res[i + k - N] = 0xFFFFFFFFu - a[i];
  } else {
    for (uint32_t i = 0; i < 2 * N - k; i++) res[i + k - N] = 0xFFFFFFFFu - a[i];
    for (uint32_t i = 2 * N - k; i < N; i++) res[i - (2 * N - k)] = a[i];
  }
  memcpy(out, res, sizeof(res));
}

/* trgsw.rs:144-171 */
void orc_decomposition(const orc_params *p, uint32_t offset, const uint32_t *a,
                       const uint32_t *b, uint32_t *dec) {
  const uint32_t l = p->l, bgbit = p->bgbit;
  const uint32_t mask = (1u << bgbit) - 1, half = 1u << (bgbit - 1);
  for (int j = 0; j < N; j++) {
    uint32_t t0 = a[j] + offset, t1 = b[j] + offset;
    for (uint32_t i = 0; i < l; i++) {
      dec[(size_t)i * N + j] = ((t0 >> (32 - (i + 1) * bgbit)) & mask) - half;
      dec[(size_t)(i + l) * N + j] = ((t1 >> (32 - (i + 1) * bgbit)) & mask) - half;
    }
  }
}

/* trgsw.rs:118-142, same evaluation order */
static void fma_in_fd(double *res, const double *a, const double *b) {
  for (int i = 0; i < N2; i++) {
    res[i] = (a[i + N2] * b[i + N2]) * 0.5 - res[i];
    res[i] = (a[i] * b[i]) * 0.5 - res[i];
    res[i + N2] += (a[i] * b[i + N2] + a[i + N2] * b[i]) * 0.5;
  }
}

static void external_product_impl(const orc_params *p, uint32_t offset,
                                  const double *bsk_row, const uint32_t *a,
                                  const uint32_t *b, uint32_t *out_a,
                                  uint32_t *out_b, double *max_frac) {
  const uint32_t l2 = 2 * p->l;
  uint32_t dec[16 * N];
  double dfft[N], oa[N], ob[N];
  orc_decomposition(p, offset, a, b, dec);
  memset(oa, 0, sizeof(oa));
  memset(ob, 0, sizeof(ob));
  for (uint32_t i = 0; i < l2; i++) {
    orc_ifft(dec + (size_t)i * N, dfft);            /* trgsw.rs:99 */
    fma_in_fd(oa, dfft, bsk_row + (size_t)i * 2 * N);     /* trgsw.rs:104 */
    fma_in_fd(ob, dfft, bsk_row + (size_t)i * 2 * N + N); /* trgsw.rs:105 */
  }
  fft_impl(oa, out_a, max_frac); /* trgsw.rs:113 */
  fft_impl(ob, out_b, max_frac); /* trgsw.rs:114 */
}
void orc_external_product(const orc_params *p, uint32_t offset, const double *bsk_row,
                          const uint32_t *a, const uint32_t *b, uint32_t *out_a,
                          uint32_t *out_b) {
  external_product_impl(p, offset, bsk_row, a, b, out_a, out_b, NULL);
}

/* ground truth: sum_r dec_r (*) bsk_r in Z_{2^32}[X]/(X^N+1) */
void orc_external_product_exact(const orc_params *p, uint32_t offset,
                                const uint32_t *bsk_torus_row, const uint32_t *a,
                                const uint32_t *b, uint32_t *out_a, uint32_t *out_b) {
  const uint32_t l2 = 2 * p->l;
  uint32_t dec[16 * N], ra[N], rb[N], t[N];
  orc_decomposition(p, offset, a, b, dec);
  memset(ra, 0, sizeof(ra));
  memset(rb, 0, sizeof(rb));
  for (uint32_t i = 0; i < l2; i++) {
    orc_poly_mul_exact(dec + (size_t)i * N, bsk_torus_row + (size_t)i * 2 * N, t);
    for (int j = 0; j < N; j++) ra[j] += t[j];
    orc_poly_mul_exact(dec + (size_t)i * N, bsk_torus_row + (size_t)i * 2 * N + N, t);
    for (int j = 0; j < N; j++) rb[j] += t[j];
  }
  memcpy(out_a, ra, sizeof(ra));
  memcpy(out_b, rb, sizeof(rb));
}

/* trgsw.rs:198-274 (blind_rotate / blind_rotate_with_testvec) + cmux :174-196 */
void orc_blind_rotate(const orc_params *p, uint32_t offset, const double *bsk_fft,
                      const uint32_t *bsk_torus, const uint32_t *tv_a,
                      const uint32_t *tv_b, const uint32_t *lwe, int steps,
                      uint32_t *acc_a, uint32_t *acc_b, double *max_frac) {
  const uint32_t n = p->n;
  const size_t row = (size_t)2 * p->l * 2 * N;
  /* trgsw.rs:202-203 (usize arithmetic: no wrap) */
  uint32_t b_tilda = (uint32_t)(2 * N - (((uint64_t)lwe[n] + (1u << 20)) >> 21));
  uint32_t ra[N], rb[N], ta[N], tb[N], ea[N], eb[N];
  orc_poly_mul_with_x_k(tv_a, b_tilda, acc_a);
  orc_poly_mul_with_x_k(tv_b, b_tilda, acc_b);
  uint32_t count = steps < 0 ? n : (uint32_t)steps;
  if (max_frac) *max_frac = 0.0;
  for (uint32_t i = 0; i < count; i++) {
    uint32_t a_tilda = (uint32_t)(lwe[i] + (1u << 20)) >> 21; /* trgsw.rs:210-211 */
    orc_poly_mul_with_x_k(acc_a, a_tilda, ra);
    orc_poly_mul_with_x_k(acc_b, a_tilda, rb);
    for (int j = 0; j < N; j++) { ta[j] = ra[j] - acc_a[j]; tb[j] = rb[j] - acc_b[j]; }
    if (bsk_torus)
      orc_external_product_exact(p, offset, bsk_torus + i * row, ta, tb, ea, eb);
    else
      external_product_impl(p, offset, bsk_fft + i * row, ta, tb, ea, eb, max_frac);
    for (int j = 0; j < N; j++) { acc_a[j] += ea[j]; acc_b[j] += eb[j]; }
  }
}

/* trlwe.rs:106-120 */
void orc_sample_extract_index(const uint32_t *a, const uint32_t *b, uint32_t k,
                              uint32_t *out) {
  for (uint32_t i = 0; i < N; i++)
    out[i] = (i <= k) ? a[k - i] : 0xFFFFFFFFu - a[N + k - i];
  out[N] = b[k];
}
/* trlwe.rs:122-136 (N := tlwe_lv0::N -- reproduced for data-flow parity only) */
void orc_sample_extract_index_2(const orc_params *p, const uint32_t *a,
                                const uint32_t *b, uint32_t k, uint32_t *out) {
  const uint32_t n = p->n;
  for (uint32_t i = 0; i < n; i++)
    out[i] = (i <= k) ? a[k - i] : 0xFFFFFFFFu - a[n + k - i];
  out[n] = b[k];
}

/* trgsw.rs:332-360 */
void orc_identity_key_switching(const orc_params *p, const uint32_t *ksk,
                                const uint32_t *src, uint32_t *out) {
  const uint32_t n = p->n, w = n + 1, basebit = p->basebit, t = p->iks_t;
  const uint32_t base = 1u << basebit;
  const uint32_t prec = orc_prec_offset(p);
  memset(out, 0, w * sizeof(uint32_t));
  out[n] = src[N];
  for (uint32_t i = 0; i < N; i++) {
    uint32_t a_bar = src[i] + prec;
    for (uint32_t j = 0; j < t; j++) {
      uint32_t k = (a_bar >> (32 - (j + 1) * basebit)) & (base - 1);
      if (k != 0) {
        const uint32_t *rowp = ksk + ((size_t)base * t * i + (size_t)base * j + k) * w;
        for (uint32_t x = 0; x < w; x++) out[x] -= rowp[x];
      }
    }
  }
}

/* bootstrap/vanilla.rs:40-63 */
void orc_bootstrap(const orc_params *p, uint32_t offset, const double *bsk_fft,
                   const uint32_t *ksk, const uint32_t *tv_a, const uint32_t *tv_b,
                   const uint32_t *lwe, int key_switch, uint32_t *out) {
  uint32_t acc_a[N], acc_b[N], ext[N + 1];
  orc_blind_rotate(p, offset, bsk_fft, NULL, tv_a, tv_b, lwe, -1, acc_a, acc_b, NULL);
  if (key_switch) {
    orc_sample_extract_index(acc_a, acc_b, 0, ext);
    orc_identity_key_switching(p, ksk, ext, out);
  } else {
    orc_sample_extract_index_2(p, acc_a, acc_b, 0, out);
  }
}

/* ------------------------------------------------------- proxy re-encryption */

/* proxy_reenc.rs:354-392 (new_symmetric_with_params) */
void orc_gen_reenc_key(const orc_params *p, const uint32_t *key_from, const uint32_t *key_to,
                       uint64_t seed, uint32_t basebit, uint32_t t, uint32_t *out) {
  const uint32_t base = 1u << basebit, n = p->n, w = n + 1;
  memset(out, 0, (size_t)base * t * n * w * sizeof(uint32_t));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < (int)n; i++) {
    orc_rng r;
    orc_rng_seed(&r, seed ^ (0x52454E00ull + (uint64_t)i * 0x9E3779B97F4A7C15ull));
    for (uint32_t j = 0; j < t; j++)
      for (uint32_t k = 1; k < base; k++) {
        double mu = (double)(k * key_from[i]) / (double)((uint64_t)1 << ((j + 1) * basebit));
        size_t idx = ((size_t)base * t * i) + (size_t)base * j + k;
        orc_lwe_encrypt_f64(p, mu, p->alpha_lv0, key_to, &r, out + idx * w);
      }
  }
}

/* proxy_reenc.rs:468-511 */
void orc_reencrypt(const orc_params *p, const uint32_t *reenc_key, uint32_t basebit, uint32_t t,
                   const uint32_t *ct_from, uint32_t *out) {
  const uint32_t n = p->n, w = n + 1, base = 1u << basebit;
  const uint32_t prec = 1u << (32 - (1 + basebit * t));
  memset(out, 0, w * sizeof(uint32_t));
  out[n] = ct_from[n];
  for (uint32_t i = 0; i < n; i++) {
    uint32_t a_bar = ct_from[i] + prec;
    for (uint32_t j = 0; j < t; j++) {
      uint32_t k = (a_bar >> (32 - (j + 1) * basebit)) & (base - 1);
      if (k != 0) {
        const uint32_t *rowp = reenc_key + ((size_t)base * t * i + (size_t)base * j + k) * w;
        for (uint32_t x = 0; x < w; x++) out[x] -= rowp[x];
      }
    }
  }
}

/* --------------------------------------------------------------------- LUT */

/* lut/generator.rs:264-266 */
uint32_t orc_div_round(uint32_t a, uint32_t b) { return (a + b / 2) / b; }
/* lut/encoder.rs:66-73 */
uint32_t orc_lut_encode(uint32_t msg, uint32_t modulus, double scale) {
  msg %= modulus;
  return orc_f64_to_torus((double)msg * scale);
}
/* lut/generator.rs:89-137; scale<=0 selects Encoder::new's 1/(2m) (encoder.rs:36) */
void orc_lut_generate(const uint32_t *f_table, uint32_t modulus, double scale,
                      uint32_t *lut_b) {
  uint32_t raw[N], rot[N];
  if (scale <= 0.0) scale = 1.0 / (2.0 * (double)modulus);
  memset(raw, 0, sizeof(raw));
  for (uint32_t x = 0; x < modulus; x++) {
    uint32_t start = orc_div_round(x * N, modulus);
    uint32_t end = orc_div_round((x + 1) * N, modulus);
    uint32_t enc = orc_lut_encode(f_table[x], modulus, scale);
    for (uint32_t i = start; i < end && i < N; i++) raw[i] = enc;
  }
  uint32_t offset = orc_div_round(N, 2 * modulus);
  for (uint32_t i = 0; i < N; i++) rot[i] = raw[(i + offset) % N];
  for (uint32_t i = N - offset; i < N; i++) rot[i] = 0u - rot[i];
  memcpy(lut_b, rot, sizeof(rot));
}

/* ------------------------------------------------------------ batch drivers */

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* gates.rs:352-547: prep all, par_map blind_rotate, par_map extract+KS */
void orc_batch_gate(const orc_cloud_key *ck, int op, const uint8_t *ops,
                    const uint32_t *in_pairs, uint32_t *out, size_t count,
                    int threads) {
  const uint32_t w = ck->p.n + 1;
  if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (long g = 0; g < (long)count; g++) {
    uint32_t prep[2048];
    int o = ops ? ops[g] : op;
    orc_gate_prep(&ck->p, o, in_pairs + (size_t)g * 2 * w, in_pairs + (size_t)g * 2 * w + w, prep);
    orc_bootstrap(&ck->p, ck->offset, ck->bsk_fft, ck->ksk, ck->tv_a, ck->tv_b, prep, 1,
                  out + (size_t)g * w);
  }
}

/* Bootstrap::bootstrap{,_without_key_switch} / LutBootstrap::bootstrap_lut over a batch */
void orc_batch_bootstrap(const orc_cloud_key *ck, const uint32_t *tv_b_override,
                         const uint32_t *in, uint32_t *out, size_t count,
                         int key_switch, int threads) {
  const uint32_t w = ck->p.n + 1;
  static const uint32_t zeros[N];
  const uint32_t *tva = tv_b_override ? zeros : ck->tv_a;
  const uint32_t *tvb = tv_b_override ? tv_b_override : ck->tv_b;
  if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (long g = 0; g < (long)count; g++)
    orc_bootstrap(&ck->p, ck->offset, ck->bsk_fft, ck->ksk, tva, tvb, in + (size_t)g * w,
                  key_switch, out + (size_t)g * w);
}

/* trgsw.rs:289-305 */
void orc_batch_blind_rotate(const orc_cloud_key *ck, const uint32_t *in,
                            uint32_t *out_trlwe, size_t count, int threads) {
  const uint32_t w = ck->p.n + 1;
  if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (long g = 0; g < (long)count; g++)
    orc_blind_rotate(&ck->p, ck->offset, ck->bsk_fft, NULL, ck->tv_a, ck->tv_b,
                     in + (size_t)g * w, -1, out_trlwe + (size_t)g * 2 * N,
                     out_trlwe + (size_t)g * 2 * N + N, NULL);
}
