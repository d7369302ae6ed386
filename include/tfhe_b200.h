/*
 * tfhe_b200.h -- C ABI of the B200-native TFHE bootstrapping engine.
 *
 * This is the drop-in boundary for rs-tfhe's hot path (gate prep -> blind
 * rotation -> sample extract -> key switch, and its LUT variant).  Each entry
 * point names the reference interface it replaces (file:line under the
 * rs-tfhe tree).  Style follows the reference's only FFI precedent,
 * src/fft/spqlios/spqlios-wrapper.cpp:10-41: an opaque handle, raw pointers,
 * caller-allocated outputs, no exceptions/panics across the boundary.
 *
 * Conventions
 *  - every function returns TFHE_OK (0) or a negative tfhe_status; the text of
 *    the last failure on the calling thread is available from tfhe_last_error();
 *  - host pointers are borrowed for the duration of the call only;
 *  - ciphertexts use the reference's memory images verbatim:
 *      TLWELv0  = u32[n+1]            (src/tlwe.rs:12-14; b is the last word)
 *      TRLWELv1 = u32[2][N] (a then b) (src/trlwe.rs:11-14)
 *      &[(Ciphertext, Ciphertext)] = u32[count][2][n+1] (src/gates.rs:352)
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with TFHE_ERR_CUDA.
 *  - calls on one engine are serialised on its stream; an engine may be shared
 *    between host threads (the Rust wrapper is Send + Sync, bootstrap/mod.rs:23).
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFHE_B200_ABI_VERSION 2
#define TFHE_N 1024 /* params::trgsw_lv1::N -- fixed in the reference (params.rs:391) */

typedef enum {
  TFHE_OK = 0,
  TFHE_ERR_INVALID = -1,   /* bad argument / unsupported parameter set */
  TFHE_ERR_CUDA = -2,      /* CUDA runtime failure or no device */
  TFHE_ERR_NO_KEY = -3,    /* cloud key not loaded */
  TFHE_ERR_ALLOC = -4
} tfhe_status;

/* Runtime image of SecurityParams (src/params.rs:53-84); the reference fixes
 * these at compile time to the 128-bit set (params.rs:426-469). */
typedef struct {
  uint32_t n;       /* tlwe_lv0::N                */
  uint32_t N;       /* trgsw_lv1::N, must be 1024 */
  uint32_t l;       /* trgsw_lv1::L      (1..3)   */
  uint32_t bgbit;   /* trgsw_lv1::BGBIT           */
  uint32_t basebit; /* trgsw_lv1::BASEBIT         */
  uint32_t iks_t;   /* trgsw_lv1::IKS_T           */
} tfhe_params;

/* Gate selector; values are shared with the oracle.  Each is the linear
 * pre-combination of src/gates.rs:54-150 followed by a bootstrap. */
typedef enum {
  TFHE_GATE_NAND = 0,  /* gates.rs:54  */
  TFHE_GATE_AND = 1,   /* gates.rs:70  */
  TFHE_GATE_OR = 2,    /* gates.rs:62  */
  TFHE_GATE_XOR = 3,   /* gates.rs:78  */
  TFHE_GATE_XNOR = 4,  /* gates.rs:86  */
  TFHE_GATE_NOR = 5,   /* gates.rs:94  */
  TFHE_GATE_ANDNY = 6, /* gates.rs:102 */
  TFHE_GATE_ANDYN = 7, /* gates.rs:115 */
  TFHE_GATE_ORNY = 8,  /* gates.rs:128 */
  TFHE_GATE_ORYN = 9,  /* gates.rs:141 */
  TFHE_GATE_COUNT = 10
} tfhe_gate;

typedef struct tfhe_engine tfhe_engine; /* opaque; one per (process, GPU), or per GPU set (create_multi) */

/* ---- library ------------------------------------------------------------ */
int tfhe_abi_version(void);
const char *tfhe_last_error(void);
/* number of CUDA devices visible (0 => every compute call will fail) */
int tfhe_device_count(void);

/* ---- engine lifetime ---------------------------------------------------- */
/* Replaces: bootstrap::default_bootstrap() / VanillaBootstrap::new()
 * (bootstrap/mod.rs:41-43, vanilla.rs:28-31) -- the strategy object a Gates
 * instance owns (gates.rs:30-45).  device_id is the CUDA ordinal. */
int tfhe_engine_create(const tfhe_params *params, int device_id, tfhe_engine **out);
/* One engine over several GPUs of this process.  The reference's batch entry points
 * parallelise with `par_map` over ciphertexts (trgsw.rs:297-305, gates.rs:352-383); here every
 * host-buffer batch call on the returned engine shards its ciphertexts over the devices in
 * contiguous index ranges [r*count/G, (r+1)*count/G), one host thread and one stream set per GPU,
 * with no per-gate communication.  The single collective is the cloud key: tfhe_engine_load_cloud_key
 * / _generate_cloud_key / _import_cloud_key upload and re-lay out on device_ids[0] and ncclBroadcast
 * the re-laid-out blob to the others over NVLink (NCCL is resolved at run time; n_devices == 1 needs
 * none).  LUT tables are kept identical on all devices.  The *_dev entry points and
 * tfhe_engine_cloud_key_blob address device_ids[0] only. */
int tfhe_engine_create_multi(const tfhe_params *params, const int *device_ids, int n_devices,
                             tfhe_engine **out);
/* number of GPUs the engine spans */
int tfhe_engine_device_count(const tfhe_engine *e);
/* device time (ms, CUDA events) of the last cloud-key ncclBroadcast; 0 for a one-GPU engine */
int tfhe_engine_last_broadcast_ms(tfhe_engine *e, float *ms_out);
void tfhe_engine_destroy(tfhe_engine *e);
/* Run all engine work on `cuda_stream` (a cudaStream_t; NULL restores the
 * engine's own stream).  Lets a host runtime order engine calls with its own. */
int tfhe_engine_set_stream(tfhe_engine *e, void *cuda_stream);
/* Number of this library's kernels launched by the engine so far. */
uint64_t tfhe_engine_kernel_launches(const tfhe_engine *e);
/* Device time (ms) of the most recent call's kernels: [0]=blind rotation,
 * [1]=key switch, measured with CUDA events on the engine stream. */
int tfhe_engine_last_kernel_ms(tfhe_engine *e, float out_ms[2]);

/* Diagnostic (roofline denominator): sustained FP64 FMA rate of the engine's GPU,
 * measured now with a register-only DFMA kernel; TFLOP/s (2 flop per FMA). */
int tfhe_probe_fp64_tflops(tfhe_engine *e, double *tflops_out);
/* The same with three distinct register operands per FMA (d = a * b + d), which is what the blind
 * rotation issues (per-lane twiddles and key values).  B200 feeds its FP64 unit one 64-bit operand
 * per lane per cycle, so this is 2/3 of the figure above (profiles/r2_fp64_operand_probe.json). */
int tfhe_probe_fp64_3op_tflops(tfhe_engine *e, double *tflops_out);

/* ---- cloud key ----------------------------------------------------------- */
/* Replaces: the CloudKey a caller passes to every gate (key.rs:51-56).  Takes
 * the reference's memory images and re-lays them out on the device once:
 *   decomposition_offset  key.rs:52 (used verbatim)
 *   testvec_a/b  u32[N]   key.rs:53, 91-100
 *   ksk  u32[N][t][2^basebit][n+1]            key.rs:54, 102-122 (index :115)
 *   bsk  f64[n][2l][2][N], each N = re[0..512) | im[0..512), values = 2*FFT
 *        (the TRGSWLv1FFT image, trgsw.rs:52-68; klemsa.rs:110-114) */
int tfhe_engine_load_cloud_key(tfhe_engine *e, uint32_t decomposition_offset,
                               const uint32_t *testvec_a, const uint32_t *testvec_b,
                               const uint32_t *ksk, const double *bsk);
/* Replaces: key::CloudKey::new(&SecretKey) (src/key.rs:59-66: gen_key_switching_key :102-122 +
 * gen_bootstrapping_key :128-156, TRGSW encryption trgsw.rs:29-68) ON THE DEVICE.
 * s0 = SecretKey.key_lv0 (n words of 0/1), s1 = key_lv1 (N words); alpha_lv0 = KSK_ALPHA,
 * alpha_lv1 = BSK_ALPHA (params.rs:468-469).  The reference's RNG is unseeded (rand::thread_rng, an
 * OS-seeded ChaCha generator), so the key is not comparable bit for bit; masks and noise come from
 * the ChaCha20 block function under a 256-bit key, straight into the device layout.
 *   seed == 0  : the key is drawn from OS entropy (getentropy) -- the ONLY setting for real keys;
 *   seed != 0  : key expanded from `seed`, reproducible, for tests and benches; INSECURE (the
 *                published masks of the key-switching key would let a 2^64 search recover it).
 * The test vector and decomposition offset are the standard ones (key.rs:78-100). */
int tfhe_engine_generate_cloud_key(tfhe_engine *e, const uint32_t *s0, const uint32_t *s1,
                                   double alpha_lv0, double alpha_lv1, uint64_t seed);
/* Multi-GPU: the device-resident, re-laid-out key is one contiguous blob.  Rank
 * 0 loads it with tfhe_engine_load_cloud_key; the other ranks call
 * tfhe_engine_alloc_cloud_key, receive the blob (e.g. ncclBroadcast over
 * NVLink into tfhe_engine_cloud_key_blob's pointer) and then
 * tfhe_engine_commit_cloud_key.  No reference counterpart (single process). */
int tfhe_engine_alloc_cloud_key(tfhe_engine *e);
int tfhe_engine_cloud_key_blob(tfhe_engine *e, void **device_ptr, size_t *bytes);
int tfhe_engine_commit_cloud_key(tfhe_engine *e, uint32_t decomposition_offset);

/* Blob/wire format (the reference has no serialization at all): the device-resident
 * key as one self-describing buffer = 64-byte header {magic "TFHEB200", version, params,
 * decomposition_offset, lut slots in use, payload bytes} + the device blob verbatim.  Lets a
 * re-laid-out key be checkpointed or shipped to another host/rank without the reference layout. */
size_t tfhe_engine_cloud_key_export_bytes(tfhe_engine *e);
int tfhe_engine_export_cloud_key(tfhe_engine *e, void *host_buf, size_t bytes);
int tfhe_engine_import_cloud_key(tfhe_engine *e, const void *host_buf, size_t bytes);

/* ---- the hot path, host buffers (H2D + kernels + D2H inside the call) ---- */
/* Replaces: gates::batch_{nand,and,or,xor,nor,xnor}[_with_railgun]
 * (gates.rs:352-547) and, with count==1, Gates::{nand,...,or_yn} (gates.rs:54-150). */
int tfhe_batch_gate(tfhe_engine *e, tfhe_gate op, const uint32_t *in_pairs /*[count][2][n+1]*/,
                    uint32_t *out /*[count][n+1]*/, size_t count);
/* Same with a per-element gate (config "random boolean circuit batch"). */
int tfhe_batch_gate_mixed(tfhe_engine *e, const uint8_t *ops /*[count]*/,
                          const uint32_t *in_pairs, uint32_t *out, size_t count);
/* Replaces: Bootstrap::bootstrap (key_switch=1, vanilla.rs:40-52) and
 * Bootstrap::bootstrap_without_key_switch (key_switch=0, vanilla.rs:54-63,
 * i.e. sample_extract_index_2, trlwe.rs:122-136) over a batch. */
int tfhe_batch_bootstrap(tfhe_engine *e, const uint32_t *in /*[count][n+1]*/,
                         uint32_t *out /*[count][n+1]*/, size_t count, int key_switch);
/* Replaces: trgsw::batch_blind_rotate[_with_railgun] (trgsw.rs:289-305). */
int tfhe_batch_blind_rotate(tfhe_engine *e, const uint32_t *in /*[count][n+1]*/,
                            uint32_t *out_trlwe /*[count][2][N]*/, size_t count);
/* Replaces: Generator::generate_lookup_table (lut/generator.rs:66-137) with the
 * closure tabulated by the caller: f_table[x] = f(x), x < modulus.  scale <= 0
 * selects Encoder::new's 1/(2*modulus) (lut/encoder.rs:29-42).  The table is
 * generated on the device and kept there under *lut_id_out; lut_b_out (may be
 * NULL) receives LookupTable.poly.b (poly.a is identically 0). */
int tfhe_lut_generate(tfhe_engine *e, const uint32_t *f_table, uint32_t modulus, double scale,
                      uint32_t *lut_b_out /*[N] or NULL*/, int *lut_id_out);
/* Register a caller-made test vector (LookupTable::from_poly, lookup_table.rs:33). */
int tfhe_lut_register(tfhe_engine *e, const uint32_t *poly_a /*[N] or NULL => 0*/,
                      const uint32_t *poly_b /*[N]*/, int *lut_id_out);
/* Drop a table made by tfhe_lut_generate / tfhe_lut_register (dropping a lookup_table::LookupTable,
 * lookup_table.rs:16-19): its device slot is reused by later tables.  An engine holds at most 62
 * live tables.  Table ids are opaque; ids from before the latest cloud-key load are rejected
 * (TFHE_ERR_INVALID), never resolved to another table.  Id 0 is the cloud key's own test vector. */
int tfhe_lut_release(tfhe_engine *e, int lut_id);
/* Replaces: LutBootstrap::bootstrap_lut (bootstrap/lut.rs:79-99) over a batch;
 * bootstrap_func (lut.rs:49-65) = tfhe_lut_generate + this. */
int tfhe_batch_bootstrap_lut(tfhe_engine *e, int lut_id, const uint32_t *in /*[count][n+1]*/,
                             uint32_t *out /*[count][n+1]*/, size_t count);
/* Replaces: LutBootstrap::bootstrap_func (bootstrap/lut.rs:49-65) over a batch in ONE call: the
 * table of f (f_table[x] = f(x), x < modulus; scale <= 0 => 1/(2*modulus)) is generated on the
 * device into a scratch slot that every call overwrites, so -- like the reference, which builds
 * and drops a LookupTable per call -- it can be called without bound and holds no table id. */
int tfhe_batch_bootstrap_func(tfhe_engine *e, const uint32_t *f_table, uint32_t modulus, double scale,
                              const uint32_t *in /*[count][n+1]*/, uint32_t *out /*[count][n+1]*/,
                              size_t count);
/* Same with a per-ciphertext table (lut_ids[count]): independent LUT bootstraps of one circuit
 * level -- e.g. the sum and carry tables of examples/lut_add_two_numbers.rs:138,147, which read
 * the same input -- go out as ONE batch instead of one call per table. */
int tfhe_batch_bootstrap_lut_multi(tfhe_engine *e, const int32_t *lut_ids /*[count]*/,
                                   const uint32_t *in /*[count][n+1]*/, uint32_t *out, size_t count);
/* Replaces: trgsw::identity_key_switching after trlwe::sample_extract_index(.,0)
 * (trgsw.rs:332-360, trlwe.rs:106-120) on caller-supplied TRLWE samples. */
int tfhe_batch_extract_key_switch(tfhe_engine *e, const uint32_t *in_trlwe /*[count][2][N]*/,
                                  uint32_t *out /*[count][n+1]*/, size_t count);

/* ---- the FFT seam: trait FFTProcessor (src/fft/mod.rs:80-107; active implementation
 *      KlemsaProcessor, src/fft/klemsa.rs:88-174) over batches, on the same device passes the
 *      blind rotation uses.  Layouts are the reference's: a polynomial is u32[N]; a spectrum is
 *      f64[N] = re[0..N/2) | im[0..N/2), natural bin order, values = 2 x the twisted DFT. ---- */
/* Replaces: FFTProcessor::batch_ifft::<1024> / ifft (fft/mod.rs:84,99-101; klemsa.rs:88-117). */
int tfhe_batch_ifft(tfhe_engine *e, const uint32_t *in /*[count][N]*/, double *out /*[count][N]*/, size_t count);
/* Replaces: FFTProcessor::batch_fft::<1024> / fft (fft/mod.rs:89,104-106; klemsa.rs:119-150). */
int tfhe_batch_fft(tfhe_engine *e, const double *in /*[count][N]*/, uint32_t *out /*[count][N]*/, size_t count);
/* Replaces: FFTProcessor::poly_mul::<1024> (fft/mod.rs:93; klemsa.rs:152-174): a*b mod X^N+1. */
int tfhe_batch_poly_mul(tfhe_engine *e, const uint32_t *a /*[count][N]*/, const uint32_t *b /*[count][N]*/,
                        uint32_t *out /*[count][N]*/, size_t count);

/* ---- next to the path: LWE proxy re-encryption = the key-switch kernel with the
 *      level-0 dimension as input ------------------------------------------------------ */
typedef struct tfhe_reenc_key tfhe_reenc_key;
/* Replaces: holding a proxy_reenc::ProxyReencryptionKey (src/proxy_reenc.rs:224-233):
 * key_encryptions u32[base*t*n][n+1], index base*t*i + base*j + k (proxy_reenc.rs:316,381). */
int tfhe_reenc_key_load(tfhe_engine *e, const uint32_t *key_encryptions, uint32_t base, uint32_t t,
                        tfhe_reenc_key **out);
void tfhe_reenc_key_destroy(tfhe_reenc_key *k);
/* Replaces: proxy_reenc::reencrypt_tlwe_lv0 (src/proxy_reenc.rs:468-511) over a batch. */
int tfhe_batch_reencrypt(tfhe_engine *e, const tfhe_reenc_key *key, const uint32_t *in /*[count][n+1]*/,
                         uint32_t *out /*[count][n+1]*/, size_t count);

/* ---- next to the path: levelised boolean circuits -------------------------------------
 * The reference evaluates circuits gate by gate through gates::Gates (examples/add_two_numbers.rs:11-49,
 * src/gates.rs:157-199 mux / mux_naive, src/circuits.rs compare_bit).  A tfhe_circuit records the same
 * gates over integer wire ids; tfhe_circuit_run splits it into levels of independent bootstraps and runs
 * each level as ONE device batch over `batch` independent input sets, wires resident in HBM (gather
 * kernel -> blind rotation -> [mux combine] -> key switch).  NOT and constants are free (resolved while
 * gathering).  tfhe_circuit_mux is the SOUND fused multiplexer: AND(sel, a) and AND(!sel, b) are blind-
 * rotated, extracted at level 1 with the ring's real degree, added with the OR offset, and key-switched
 * once -- what Gates::mux (gates.rs:157-183) intends (its sample_extract_index_2 is wrong). */
typedef struct tfhe_circuit tfhe_circuit;
int tfhe_circuit_create(tfhe_engine *e, tfhe_circuit **out);
void tfhe_circuit_destroy(tfhe_circuit *c);
int tfhe_circuit_input(tfhe_circuit *c, uint32_t *wire_out);
int tfhe_circuit_constant(tfhe_circuit *c, int value, uint32_t *wire_out);            /* gates.rs:212-218 */
int tfhe_circuit_not(tfhe_circuit *c, uint32_t a, uint32_t *wire_out);               /* gates.rs:202-204 */
int tfhe_circuit_gate(tfhe_circuit *c, tfhe_gate op, uint32_t a, uint32_t b, uint32_t *wire_out);
int tfhe_circuit_mux(tfhe_circuit *c, uint32_t sel, uint32_t then_wire, uint32_t else_wire, uint32_t *wire_out);
int tfhe_circuit_output(tfhe_circuit *c, uint32_t wire);
/* depth in bootstrap levels, blind rotations and key switches per input set */
int tfhe_circuit_stats(tfhe_circuit *c, uint32_t *levels, uint32_t *bootstraps, uint32_t *key_switches);
/* inputs u32[n_inputs][batch][n+1] (input order = tfhe_circuit_input order), outputs
 * u32[n_outputs][batch][n+1] (tfhe_circuit_output order); host buffers. */
int tfhe_circuit_run(tfhe_circuit *c, const uint32_t *inputs, uint32_t *outputs, size_t batch);

/* ---- the hot path, device-resident buffers (no copies; asynchronous on the
 *      engine stream).  Pointers are CUDA device pointers on the engine's GPU. */
int tfhe_batch_gate_dev(tfhe_engine *e, tfhe_gate op, const uint8_t *d_ops /*NULL or [count]*/,
                        const uint32_t *d_in_pairs, uint32_t *d_out, size_t count);
int tfhe_batch_bootstrap_dev(tfhe_engine *e, int lut_id /*<0: cloud-key test vector*/,
                             const uint32_t *d_in, uint32_t *d_out, size_t count, int key_switch);
int tfhe_engine_synchronize(tfhe_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* TFHE_B200_H */
