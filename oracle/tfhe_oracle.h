/*
 * tfhe_oracle.h -- CPU restatement of rs-tfhe's bootstrapped-gate hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker and the CPU baseline
 * ("port" of the reference's Rayon path).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (rs_tfhe_b200/, libtfhe_b200.so) never links, imports or calls anything here.
 *
 * Parity pinning: the reference (thedonutfactory/rs-tfhe, Rust) cannot be built
 * in this image (no cargo/rustc) and holds no golden vectors (its RNG is
 * unseeded, src/key.rs:34, src/tlwe.rs:38).  The 512-point complex FFT lives in
 * the third-party crate rustfft ^6.1 (Cargo.toml:21, unpinned, no lockfile),
 * whose bits are not portable between CPUs; the reference's own tests pin the
 * transform only to +-1 LSB of the exact integer negacyclic product
 * (src/fft/mod.rs:135-159,240-255).  This oracle is therefore pinned to
 *   (1) every known-answer the reference's tests hold for the path
 *       (tests/test_oracle_known_answers.py), and
 *   (2) an exact-integer external product (orc_external_product_exact), which
 *       the f64 path must equal bit for bit at the l=3 parameter sets, and
 *   (3) the reference's dormant SPQLIOS C++/asm FFT compiled from
 *       /root/reference into oracle/_ref (negacyclic product cross-check).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_N 1024 /* TRLWE ring degree; fixed in the reference (params.rs:391) */

/* Runtime restatement of SecurityParams (src/params.rs:53-84). */
typedef struct {
  uint32_t n;       /* tlwe_lv0.n                      */
  uint32_t N;       /* trgsw_lv1.n, must be 1024       */
  uint32_t l;       /* trgsw_lv1.l                     */
  uint32_t bgbit;   /* trgsw_lv1.bgbit                 */
  uint32_t basebit; /* trgsw_lv1.basebit               */
  uint32_t iks_t;   /* trgsw_lv1.iks_t                 */
  double alpha_lv0; /* tlwe_lv0.alpha (= KSK_ALPHA)    */
  double alpha_lv1; /* tlwe_lv1.alpha (= BSK_ALPHA)    */
} orc_params;

/* Named sets: "80", "110", "128", "uint1".."uint8" (src/params.rs:91-404). */
int orc_params_by_name(const char *name, orc_params *out);

/* ---- seeded RNG (the reference uses unseeded thread_rng; nothing to match) */
typedef struct { uint64_t s[4]; int has_spare; double spare; } orc_rng;
void orc_rng_seed(orc_rng *r, uint64_t seed);
uint32_t orc_rng_u32(orc_rng *r);
double orc_rng_normal(orc_rng *r, double sigma);

/* ---- scalar helpers */
uint32_t orc_f64_to_torus(double d);                 /* utils.rs:9-12 */
double orc_torus_to_f64(uint32_t t);                 /* utils.rs:14-16 */
uint32_t orc_decomposition_offset(const orc_params *p); /* key.rs:78-89 */
uint32_t orc_prec_offset(const orc_params *p);       /* trgsw.rs:345 */

/* ---- sizes (in elements) */
size_t orc_ksk_words(const orc_params *p); /* N * t * 2^basebit * (n+1) u32 */
size_t orc_bsk_doubles(const orc_params *p); /* n * 2l * 2 * N f64 */

/* ---- negacyclic transforms, Klemsa convention (fft/klemsa.rs:88-174) */
void orc_ifft(const uint32_t *in /*N*/, double *out /*N: re|im*/);
void orc_fft(const double *in /*N: re|im*/, uint32_t *out /*N*/);
void orc_poly_mul(const uint32_t *a, const uint32_t *b, uint32_t *out);
/* exact O(N^2) wrapping negacyclic product (fft/mod.rs:240-255) */
void orc_poly_mul_exact(const uint32_t *a, const uint32_t *b, uint32_t *out);

/* ---- keys (key.rs:33-48, 91-156; trgsw.rs:29-68; trlwe.rs:30-52,91-96) */
void orc_secret_key(const orc_params *p, uint64_t seed, uint32_t *s0 /*n*/,
                    uint32_t *s1 /*N*/);
void orc_gen_testvec(uint32_t *a /*N*/, uint32_t *b /*N*/);
void orc_gen_ksk(const orc_params *p, const uint32_t *s0, const uint32_t *s1,
                 uint64_t seed, uint32_t *ksk);
/* bsk_torus may be NULL.  bsk_fft layout: f64[n][2l][2(a,b)][N re|im], x2-scaled
 * (the reference's TRGSWLv1FFT memory image, trgsw.rs:52-68).  bsk_torus:
 * u32[n][2l][2][N], the same TRGSW rows before the transform. */
void orc_gen_bsk(const orc_params *p, const uint32_t *s0, const uint32_t *s1,
                 uint64_t seed, double *bsk_fft, uint32_t *bsk_torus);

/* ---- LWE (tlwe.rs:37-68, 84-126) */
void orc_lwe_encrypt_f64(const orc_params *p, double mu, double alpha,
                         const uint32_t *s0, orc_rng *rng, uint32_t *ct);
void orc_lwe_encrypt_bool(const orc_params *p, int bit, const uint32_t *s0,
                          orc_rng *rng, uint32_t *ct);
void orc_lwe_encrypt_message(const orc_params *p, uint32_t msg, uint32_t modulus,
                             const uint32_t *s0, orc_rng *rng, uint32_t *ct);
uint32_t orc_lwe_phase(const uint32_t *ct, const uint32_t *key, uint32_t n);
int orc_lwe_decrypt_bool(const uint32_t *ct, const uint32_t *key, uint32_t n);
uint32_t orc_lwe_decrypt_message(const uint32_t *ct, const uint32_t *key,
                                 uint32_t n, uint32_t modulus);

/* batch helpers for large trial counts (OpenMP; one RNG stream per element) */
void orc_lwe_encrypt_batch(const orc_params *p, const double *mu, size_t count, double alpha,
                           const uint32_t *s0, uint64_t seed, uint32_t *cts);
void orc_lwe_phase_batch(const uint32_t *cts, size_t count, const uint32_t *key, uint32_t n,
                         uint32_t *phases);

/* ---- the hot path */
/* gates.rs:54-150 (prep only).  op codes == enum tfhe_gate in include/tfhe_b200.h */
void orc_gate_prep(const orc_params *p, int op, const uint32_t *a,
                   const uint32_t *b, uint32_t *out);
void orc_poly_mul_with_x_k(const uint32_t *a, uint32_t k, uint32_t *out); /* trgsw.rs:307-330 */
void orc_decomposition(const orc_params *p, uint32_t offset, const uint32_t *a,
                       const uint32_t *b, uint32_t *dec /*[2l][N]*/); /* trgsw.rs:144-171 */
/* trgsw.rs:77-142: one TRGSW (x) TRLWE; bsk_row = &bsk_fft[i*2l*2*N] */
void orc_external_product(const orc_params *p, uint32_t offset,
                          const double *bsk_row, const uint32_t *a,
                          const uint32_t *b, uint32_t *out_a, uint32_t *out_b);
/* ground truth: same contraction in exact wrapping integer arithmetic */
void orc_external_product_exact(const orc_params *p, uint32_t offset,
                                const uint32_t *bsk_torus_row, const uint32_t *a,
                                const uint32_t *b, uint32_t *out_a, uint32_t *out_b);
/* trgsw.rs:198-274.  steps<0 => all n; otherwise stop after `steps` CMUXes
 * (trajectory checks).  bsk_torus!=NULL selects the exact-integer product.
 * max_frac (may be NULL) receives the largest |y - round(y)| seen before
 * rounding in the f64 inverse transform. */
void orc_blind_rotate(const orc_params *p, uint32_t offset, const double *bsk_fft,
                      const uint32_t *bsk_torus, const uint32_t *tv_a,
                      const uint32_t *tv_b, const uint32_t *lwe, int steps,
                      uint32_t *acc_a, uint32_t *acc_b, double *max_frac);
void orc_sample_extract_index(const uint32_t *a, const uint32_t *b, uint32_t k,
                              uint32_t *out /*N+1*/); /* trlwe.rs:106-120 */
void orc_sample_extract_index_2(const orc_params *p, const uint32_t *a,
                                const uint32_t *b, uint32_t k,
                                uint32_t *out /*n+1*/); /* trlwe.rs:122-136 */
void orc_identity_key_switching(const orc_params *p, const uint32_t *ksk,
                                const uint32_t *src /*N+1*/, uint32_t *out /*n+1*/); /* trgsw.rs:332-360 */
/* bootstrap/vanilla.rs:40-63 (key_switch=0 => sample_extract_index_2) */
void orc_bootstrap(const orc_params *p, uint32_t offset, const double *bsk_fft,
                   const uint32_t *ksk, const uint32_t *tv_a, const uint32_t *tv_b,
                   const uint32_t *lwe, int key_switch, uint32_t *out);

/* ---- proxy re-encryption (src/proxy_reenc.rs:354-392 symmetric key, :468-511 reencrypt) */
void orc_gen_reenc_key(const orc_params *p, const uint32_t *key_from, const uint32_t *key_to,
                       uint64_t seed, uint32_t basebit, uint32_t t, uint32_t *out /*[2^basebit*t*n][n+1]*/);
void orc_reencrypt(const orc_params *p, const uint32_t *reenc_key, uint32_t basebit, uint32_t t,
                   const uint32_t *ct_from, uint32_t *out);

/* ---- LUT (lut/generator.rs:89-137,264-266; lut/encoder.rs:29-42,66-73) */
uint32_t orc_div_round(uint32_t a, uint32_t b);
uint32_t orc_lut_encode(uint32_t msg, uint32_t modulus, double scale);
void orc_lut_generate(const uint32_t *f_table, uint32_t modulus, double scale,
                      uint32_t *lut_b /*N*/);

/* ---- batch drivers (gates.rs:352-547; OpenMP plays Rayon's par_map) */
typedef struct {
  orc_params p;
  uint32_t offset;
  const uint32_t *tv_a, *tv_b;
  const uint32_t *ksk;
  const double *bsk_fft;
} orc_cloud_key;
/* ops==NULL => every element uses `op` */
void orc_batch_gate(const orc_cloud_key *ck, int op, const uint8_t *ops,
                    const uint32_t *in_pairs, uint32_t *out, size_t count,
                    int threads);
void orc_batch_bootstrap(const orc_cloud_key *ck, const uint32_t *tv_b_override,
                         const uint32_t *in, uint32_t *out, size_t count,
                         int key_switch, int threads);
void orc_batch_blind_rotate(const orc_cloud_key *ck, const uint32_t *in,
                            uint32_t *out_trlwe, size_t count, int threads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
