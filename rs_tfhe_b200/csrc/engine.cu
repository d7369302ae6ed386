// engine.cu -- the C ABI (include/tfhe_b200.h): engine handle, cloud-key upload and
// re-layout, batch entry points.  No CPU fallback: every compute entry point
// needs a CUDA device and fails loudly (TFHE_ERR_CUDA) without one.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "kernels.h"

#ifndef TFHE_KS_DEFAULT
#define TFHE_KS_DEFAULT 0   // 0 = tcgen05 (umma), 1 = mma.sync, 2 = row walk
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                     \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess)                                                           \
      return fail(TFHE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                               \
  } while (0)

constexpr int kMaxLut = 64;          // test-vector slots (slot 0 = cloud-key test vector)
constexpr size_t kChunk = 1u << 17;  // ciphertexts per internal pass (bounds scratch memory)

struct Scratch {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace

struct tfhe_engine {
  tfhe_params p{};
  int dev = 0;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::mutex mu;
  // cloud key: one contiguous device blob = BSK | KSK | test-vector slots
  uint8_t *blob = nullptr;
  cplx *bsk2 = nullptr;   // BSK rows in the TMEM-exchange kernel's thread order (derived, not in the blob)
  uint8_t *kumma = nullptr;  // KSK as tcgen05 operand tiles (derived from the blob's KSK rows; gate sets)
  size_t blob_bytes = 0, off_ksk = 0, off_tv = 0;
  uint32_t *kmma = nullptr;  // KSK as mma.sync B fragments (derived; only under TFHE_KS_VARIANT=mma, basebit 2)
  bool key_loaded = false;
  uint32_t decomp_offset = 0;
  uint32_t ksk_rows = 0, ksk_stride = 0;
  int n_lut = 1;
  cplx *tw_a = nullptr, *tw_b = nullptr;
  Scratch s_misc;
  // two pipeline slots: H2D of chunk k+1 and D2H of chunk k-1 overlap the kernels of chunk k
  struct Slot {
    Scratch in, ext, out, ops, idx;
    cudaEvent_t h2d_done = nullptr, br_start = nullptr, br_end = nullptr, ks_end = nullptr,
                d2h_done = nullptr;
    bool busy = false;
  } slot[2];
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_ms[2] = {0.f, 0.f};
  uint64_t launches = 0;

  const cplx *bsk() const { return reinterpret_cast<const cplx *>(blob); }
  const uint32_t *ksk() const { return reinterpret_cast<const uint32_t *>(blob + off_ksk); }
  uint32_t *tv() const { return reinterpret_cast<uint32_t *>(blob + off_tv); }
};

// proxy_reenc::ProxyReencryptionKey (src/proxy_reenc.rs:224-233) resident on the device
struct tfhe_reenc_key {
  uint32_t *rows = nullptr;  // [base*t*n + 1][stride], last row zero
  uint32_t basebit = 0, t = 0, n_rows = 0;
  int dev = 0;
};

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

void blob_layout(tfhe_engine *e) {
  const tfhe_params &p = e->p;
  e->ksk_rows = TFHE_N * p.iks_t * (1u << p.basebit);
  e->ksk_stride = ks_stride(p.n);
  size_t bsk_bytes = (size_t)p.n * 2 * p.l * br::kChunkCplx * sizeof(cplx);
  size_t ksk_bytes = ((size_t)e->ksk_rows + 1) * e->ksk_stride * 4;
  e->off_ksk = align_up(bsk_bytes, 256);
  e->off_tv = align_up(e->off_ksk + ksk_bytes, 256);
  e->blob_bytes = e->off_tv + (size_t)kMaxLut * 2 * TFHE_N * 4;
}

// K4 kernel selection: TFHE_KS_VARIANT = umma (default: tcgen05.mma, TMEM accumulators) | mma
// (mma.sync, register accumulators) | rows (row-walk kernels; always used when basebit != 2).
enum { KS_UMMA = 0, KS_MMA = 1, KS_ROWS = 2 };
int ks_variant() {
  static const int v = [] {
    const char *s = getenv("TFHE_KS_VARIANT");
    if (!s || !s[0]) return (int)TFHE_KS_DEFAULT;
    return s[0] == 'r' ? (int)KS_ROWS : s[0] == 'm' ? (int)KS_MMA : (int)KS_UMMA;
  }();
  return v;
}

// Derived key material: rebuilt whenever the blob's BSK changes (upload, generation, commit, import).
int finalize_key(tfhe_engine *e) {
  if (br_uses_permuted_key()) {
    const size_t rows = (size_t)e->p.n * 2 * e->p.l;
    if (!e->bsk2) CU(cudaMalloc(reinterpret_cast<void **>(&e->bsk2), rows * br::kChunkCplx * sizeof(cplx)));
    CU(bsk_permute_launch(e->bsk(), e->bsk2, rows, e->stream));
    e->launches++;
  }
  if (ks_variant() == KS_UMMA && ks_umma_supported(e->p.basebit, e->p.iks_t)) {
    if (!e->kumma) CU(cudaMalloc(reinterpret_cast<void **>(&e->kumma), ks_umma_key_bytes(e->p.n, e->p.iks_t, e->p.basebit)));
    CU(ksk_umma_relayout_launch(e->ksk(), e->ksk_stride, e->kumma, e->p.n, e->p.iks_t, e->p.basebit, e->stream));
    e->launches++;
  }
  if (ks_variant() == KS_MMA && e->p.basebit == 2) {
    if (!e->kmma) CU(cudaMalloc(reinterpret_cast<void **>(&e->kmma), ks_mma_words(e->p.n, e->p.iks_t) * 4));
    CU(ksk_mma_relayout_launch(e->ksk(), e->ksk_stride, e->kmma, e->p.n, e->p.iks_t, e->stream));
    e->launches++;
  }
  CU(cudaStreamSynchronize(e->stream));
  return TFHE_OK;
}

int ensure_blob(tfhe_engine *e) {
  if (e->blob) return TFHE_OK;
  blob_layout(e);
  CU(cudaMalloc(reinterpret_cast<void **>(&e->blob), e->blob_bytes));
  return TFHE_OK;
}

// K4 dispatch: tensor-core GEMM for the gate sets, row-walk kernels otherwise
int key_switch(tfhe_engine *e, const uint32_t *d_ext, uint32_t *d_out, size_t count) {
  const int variant = ks_variant();
  if (variant == KS_UMMA && e->kumma) {
    KsUmmaArgs k{};
    k.key = e->kumma; k.ext = d_ext; k.out = d_out;
    k.n = e->p.n; k.iks_t = e->p.iks_t; k.basebit = e->p.basebit; k.count = count;
    CU(ks_umma_launch(k, e->stream));
  } else if (variant == KS_MMA && e->kmma) {
    KsMmaArgs k{};
    k.w = e->kmma; k.ext = d_ext; k.out = d_out;
    k.n = e->p.n; k.iks_t = e->p.iks_t; k.nxg = ks_mma_nxg(e->p.n); k.count = count;
    CU(ks_mma_launch(k, e->stream));
  } else {
    KsArgs k{};
    k.ksk = e->ksk(); k.ext = d_ext; k.out = d_out;
    k.n = e->p.n; k.basebit = e->p.basebit; k.iks_t = e->p.iks_t;
    k.stride = e->ksk_stride; k.zero_row = e->ksk_rows; k.n_in = TFHE_N; k.count = count;
    CU(ks_launch(k, e->stream));
  }
  e->launches++;
  return TFHE_OK;
}

// One pass over <= kChunk ciphertexts already on the device.
//   gate mode: op >= 0 or d_ops != NULL; plain mode: op < 0 and d_ops == NULL.
//   out_kind: 0 key-switched LWE [n+1]; 1 extract_2 [n+1]; 2 TRLWE [2][N]
int run_device(tfhe_engine *e, tfhe_engine::Slot &sl, int op, const uint8_t *d_ops, int lut_id,
               const uint32_t *d_in, uint32_t *d_out, size_t count, int out_kind,
               const int32_t *d_lut_ids = nullptr) {
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  if (lut_id >= e->n_lut) return fail(TFHE_ERR_INVALID, "unknown lut id %d", lut_id);
  BrArgs a{};
  a.bsk = e->bsk();
  a.bsk2 = e->bsk2;
  a.tw_a = e->tw_a; a.tw_b = e->tw_b;
  a.tv = e->tv(); a.tv_index = d_lut_ids; a.tv_default = lut_id < 0 ? 0 : lut_id;
  a.in = d_in; a.ops = d_ops; a.op = op;
  a.n = e->p.n; a.offset = e->decomp_offset; a.count = count;
  if (out_kind == 0) {
    CU(sl.ext.reserve(count * (TFHE_N + 1) * 4));
    a.out = static_cast<uint32_t *>(sl.ext.p);
    a.out_mode = BR_OUT_EXTRACT;
  } else {
    a.out = d_out;
    a.out_mode = out_kind == 1 ? BR_OUT_EXTRACT2 : BR_OUT_TRLWE;
  }
  CU(cudaEventRecord(sl.br_start, e->stream));
  CU(br_launch(e->p.l, e->p.bgbit, a, e->num_sms, e->stream));
  e->launches++;
  CU(cudaEventRecord(sl.br_end, e->stream));
  if (out_kind == 0) {
    int rc = key_switch(e, static_cast<const uint32_t *>(sl.ext.p), d_out, count);
    if (rc != TFHE_OK) return rc;
  }
  CU(cudaEventRecord(sl.ks_end, e->stream));
  return TFHE_OK;
}

// wait for a slot's D2H and add its kernel times to the running totals
int retire_slot(tfhe_engine::Slot &sl, float &ms0, float &ms1) {
  if (!sl.busy) return TFHE_OK;
  CU(cudaEventSynchronize(sl.d2h_done));
  float t0 = 0.f, t1 = 0.f;
  CU(cudaEventElapsedTime(&t0, sl.br_start, sl.br_end));
  CU(cudaEventElapsedTime(&t1, sl.br_end, sl.ks_end));
  ms0 += t0; ms1 += t1;
  sl.busy = false;
  return TFHE_OK;
}

// Host-buffer driver: chunks of a few dozen persistent-grid rounds flow through a two-slot
// pipeline (copy-in stream -> engine stream -> copy-out stream), so the host<->device copies of
// neighbouring chunks overlap the kernels.
int run_host(tfhe_engine *e, int op, const uint8_t *ops, int lut_id, const uint32_t *in,
             size_t in_words, uint32_t *out, size_t out_words, size_t count, int out_kind,
             const int32_t *lut_ids = nullptr) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (count == 0) return TFHE_OK;
  if (!in || !out) return fail(TFHE_ERR_INVALID, "null buffer");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  float ms0 = 0.f, ms1 = 0.f;
  // Chunks are whole rounds of the persistent grid (4 ciphertexts per SM): 28 rounds in steady
  // state, but a short first chunk (its upload is the only one no kernel hides) and a short last
  // one (likewise its download).
  const size_t round = (size_t)e->num_sms * 4;
  const size_t chunk = round * 28, edge = round * 4;
  // order after whatever the caller queued on the engine stream
  CU(cudaEventRecord(e->ev[3], e->stream));
  CU(cudaStreamWaitEvent(e->copy_in, e->ev[3], 0));
  int k = 0;
  for (size_t base = 0, c = 0; base < count; base += c, k++) {
    const size_t left = count - base;
    if (base == 0 && left > 2 * edge) c = edge;
    else if (left > chunk + edge) c = chunk;
    else if (left > 2 * edge) c = left - edge;
    else c = left;
    tfhe_engine::Slot &sl = e->slot[k & 1];
    int rc = retire_slot(sl, ms0, ms1);
    if (rc != TFHE_OK) return rc;
    CU(sl.in.reserve(c * in_words * 4));
    CU(sl.out.reserve(c * out_words * 4));
    CU(cudaMemcpyAsync(sl.in.p, in + base * in_words, c * in_words * 4, cudaMemcpyHostToDevice,
                       e->copy_in));
    const uint8_t *d_ops = nullptr;
    if (ops) {
      CU(sl.ops.reserve(c));
      CU(cudaMemcpyAsync(sl.ops.p, ops + base, c, cudaMemcpyHostToDevice, e->copy_in));
      d_ops = static_cast<const uint8_t *>(sl.ops.p);
    }
    const int32_t *d_ids = nullptr;
    if (lut_ids) {
      CU(sl.idx.reserve(c * 4));
      CU(cudaMemcpyAsync(sl.idx.p, lut_ids + base, c * 4, cudaMemcpyHostToDevice, e->copy_in));
      d_ids = static_cast<const int32_t *>(sl.idx.p);
    }
    CU(cudaEventRecord(sl.h2d_done, e->copy_in));
    CU(cudaStreamWaitEvent(e->stream, sl.h2d_done, 0));
    rc = run_device(e, sl, op, d_ops, lut_id, static_cast<const uint32_t *>(sl.in.p),
                    static_cast<uint32_t *>(sl.out.p), c, out_kind, d_ids);
    if (rc != TFHE_OK) return rc;
    CU(cudaStreamWaitEvent(e->copy_out, sl.ks_end, 0));
    CU(cudaMemcpyAsync(out + base * out_words, sl.out.p, c * out_words * 4, cudaMemcpyDeviceToHost,
                       e->copy_out));
    CU(cudaEventRecord(sl.d2h_done, e->copy_out));
    sl.busy = true;
  }
  for (auto &sl : e->slot) {
    int rc = retire_slot(sl, ms0, ms1);
    if (rc != TFHE_OK) return rc;
  }
  CU(cudaStreamSynchronize(e->stream));
  e->last_ms[0] = ms0; e->last_ms[1] = ms1;
  return TFHE_OK;
}

}  // namespace

static int engine_init(tfhe_engine *e, int device_id) {
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device_id));
  e->num_sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  CU(cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking));
  for (auto &ev : e->ev) CU(cudaEventCreate(&ev));
  for (auto &sl : e->slot)
    for (cudaEvent_t *pe : {&sl.h2d_done, &sl.br_start, &sl.br_end, &sl.ks_end, &sl.d2h_done})
      CU(cudaEventCreate(pe));
  // twiddles (br_core.cuh): ta[r][k0] = e^{i pi r(1-4k0)/1024}, tb[j][x] = e^{-2 pi i jx/64}
  std::vector<cplx> ta(64 * 8), tb(8 * 8);
  for (int r = 0; r < 64; r++)
    for (int k0 = 0; k0 < 8; k0++) {
      int idx = ((r * (1 - 4 * k0)) % 2048 + 2048) % 2048;
      double ang = M_PI * (double)idx / 1024.0;
      ta[r * 8 + k0] = br::mk(std::cos(ang), std::sin(ang));
    }
  for (int j = 0; j < 8; j++)
    for (int x = 0; x < 8; x++) {
      double ang = -2.0 * M_PI * (double)((j * x) % 64) / 64.0;
      tb[j * 8 + x] = br::mk(std::cos(ang), std::sin(ang));
    }
  CU(cudaMalloc(reinterpret_cast<void **>(&e->tw_a), ta.size() * sizeof(cplx)));
  CU(cudaMalloc(reinterpret_cast<void **>(&e->tw_b), tb.size() * sizeof(cplx)));
  CU(cudaMemcpy(e->tw_a, ta.data(), ta.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->tw_b, tb.data(), tb.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  blob_layout(e);
  return TFHE_OK;
}

extern "C" {
void tfhe_engine_destroy(tfhe_engine *e);

int tfhe_abi_version(void) { return TFHE_B200_ABI_VERSION; }
const char *tfhe_last_error(void) { return g_err; }

int tfhe_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int tfhe_engine_create(const tfhe_params *params, int device_id, tfhe_engine **out) {
  if (!params || !out) return fail(TFHE_ERR_INVALID, "null argument");
  *out = nullptr;
  const tfhe_params &p = *params;
  if (p.N != TFHE_N) return fail(TFHE_ERR_INVALID, "N must be %d (got %u)", TFHE_N, p.N);
  if (!br_supported(p.l, p.bgbit))
    return fail(TFHE_ERR_INVALID, "unsupported gadget (l=%u, bgbit=%u)", p.l, p.bgbit);
  if (p.n == 0 || p.n > 1216) return fail(TFHE_ERR_INVALID, "n out of range (%u)", p.n);
  if (p.basebit == 0 || p.iks_t == 0 || p.basebit * p.iks_t > 31)
    return fail(TFHE_ERR_INVALID, "bad key-switch parameters");
  if (ks_stride(p.n) / 4 > 320) return fail(TFHE_ERR_INVALID, "n too large for key switch");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(TFHE_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU path",
                ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  }
  if (device_id < 0 || device_id >= ndev)
    return fail(TFHE_ERR_INVALID, "device %d out of range (have %d)", device_id, ndev);
  CU(cudaSetDevice(device_id));
  tfhe_engine *e = new (std::nothrow) tfhe_engine();
  if (!e) return fail(TFHE_ERR_ALLOC, "out of host memory");
  e->p = p;
  e->dev = device_id;
  int rc = engine_init(e, device_id);
  if (rc != TFHE_OK) {  // release whatever was created; keep the error text
    char saved[sizeof(g_err)];
    memcpy(saved, g_err, sizeof(saved));
    tfhe_engine_destroy(e);
    memcpy(g_err, saved, sizeof(saved));
    return rc;
  }
  *out = e;
  return TFHE_OK;
}

void tfhe_engine_destroy(tfhe_engine *e) {
  if (!e) return;
  cudaSetDevice(e->dev);
  cudaDeviceSynchronize();
  if (e->blob) cudaFree(e->blob);
  if (e->bsk2) cudaFree(e->bsk2);
  if (e->kumma) cudaFree(e->kumma);
  if (e->kmma) cudaFree(e->kmma);
  if (e->tw_a) cudaFree(e->tw_a);
  if (e->tw_b) cudaFree(e->tw_b);
  e->s_misc.release();
  for (auto &sl : e->slot) {
    sl.in.release(); sl.ext.release(); sl.out.release(); sl.ops.release(); sl.idx.release();
    for (cudaEvent_t pe : {sl.h2d_done, sl.br_start, sl.br_end, sl.ks_end, sl.d2h_done})
      if (pe) cudaEventDestroy(pe);
  }
  if (e->copy_in) cudaStreamDestroy(e->copy_in);
  if (e->copy_out) cudaStreamDestroy(e->copy_out);
  for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

int tfhe_engine_set_stream(tfhe_engine *e, void *cuda_stream) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  return TFHE_OK;
}

uint64_t tfhe_engine_kernel_launches(const tfhe_engine *e) { return e ? e->launches : 0; }

int tfhe_engine_last_kernel_ms(tfhe_engine *e, float out_ms[2]) {
  if (!e || !out_ms) return fail(TFHE_ERR_INVALID, "null argument");
  out_ms[0] = e->last_ms[0];
  out_ms[1] = e->last_ms[1];
  return TFHE_OK;
}

int tfhe_probe_fp64_tflops(tfhe_engine *e, double *tflops_out) {
  if (!e || !tflops_out) return fail(TFHE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  CU(e->s_misc.reserve(64));
  const int blocks = e->num_sms * 8, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(e->ev[3], e->stream));
    CU(fp64_probe_launch(static_cast<double *>(e->s_misc.p), blocks, iters, e->stream));
    CU(cudaEventRecord(e->ev[2], e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e->ev[3], e->ev[2]));
    double flops = (double)blocks * 256.0 * (double)iters * 32.0 * 2.0;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  e->launches += 4;
  *tflops_out = best;
  return TFHE_OK;
}

int tfhe_engine_load_cloud_key(tfhe_engine *e, uint32_t decomposition_offset,
                               const uint32_t *testvec_a, const uint32_t *testvec_b,
                               const uint32_t *ksk, const double *bsk) {
  if (!e || !testvec_a || !testvec_b || !ksk || !bsk) return fail(TFHE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  int rc = ensure_blob(e);
  if (rc != TFHE_OK) return rc;
  const tfhe_params &p = e->p;
  const size_t bsk_ref_bytes = (size_t)p.n * 2 * p.l * 2 * TFHE_N * sizeof(double);
  const size_t ksk_ref_bytes = (size_t)e->ksk_rows * (p.n + 1) * 4;
  CU(e->s_misc.reserve(bsk_ref_bytes > ksk_ref_bytes ? bsk_ref_bytes : ksk_ref_bytes));
  CU(cudaMemcpyAsync(e->s_misc.p, bsk, bsk_ref_bytes, cudaMemcpyHostToDevice, e->stream));
  CU(bsk_relayout_launch(static_cast<const double *>(e->s_misc.p),
                         reinterpret_cast<cplx *>(e->blob), p.n, 2 * p.l, e->stream));
  CU(cudaMemcpyAsync(e->s_misc.p, ksk, ksk_ref_bytes, cudaMemcpyHostToDevice, e->stream));
  CU(ksk_relayout_launch(static_cast<const uint32_t *>(e->s_misc.p),
                         reinterpret_cast<uint32_t *>(e->blob + e->off_ksk), e->ksk_rows, p.n,
                         e->ksk_stride, e->stream));
  e->launches += 2;
  CU(cudaMemsetAsync(e->tv(), 0, (size_t)kMaxLut * 2 * TFHE_N * 4, e->stream));
  CU(cudaMemcpyAsync(e->tv(), testvec_a, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->tv() + TFHE_N, testvec_b, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->s_misc.release();
  e->decomp_offset = decomposition_offset;
  e->n_lut = 1;
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return TFHE_OK;
}

int tfhe_engine_generate_cloud_key(tfhe_engine *e, const uint32_t *s0, const uint32_t *s1,
                                   double alpha_lv0, double alpha_lv1, uint64_t seed) {
  if (!e || !s0 || !s1) return fail(TFHE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  int rc = ensure_blob(e);
  if (rc != TFHE_OK) return rc;
  const tfhe_params &p = e->p;
  // scratch: s0 | s1 | spectrum(s1) | KSK in the reference row layout
  const size_t off_s1 = align_up((size_t)p.n * 4, 256);
  const size_t off_spec = align_up(off_s1 + TFHE_N * 4, 256);
  const size_t off_ksk = align_up(off_spec + 512 * sizeof(cplx), 256);
  const size_t ksk_ref_bytes = (size_t)e->ksk_rows * (p.n + 1) * 4;
  CU(e->s_misc.reserve(off_ksk + ksk_ref_bytes));
  uint8_t *sc = static_cast<uint8_t *>(e->s_misc.p);
  CU(cudaMemcpyAsync(sc, s0, (size_t)p.n * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(sc + off_s1, s1, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  uint32_t *d_ksk_ref = reinterpret_cast<uint32_t *>(sc + off_ksk);
  CU(keygen_launch(e->tw_a, e->tw_b, reinterpret_cast<const uint32_t *>(sc),
                   reinterpret_cast<const uint32_t *>(sc + off_s1),
                   reinterpret_cast<cplx *>(sc + off_spec), reinterpret_cast<cplx *>(e->blob),
                   d_ksk_ref, p.n, p.l, p.bgbit, p.basebit, p.iks_t, alpha_lv0, alpha_lv1, seed,
                   e->stream));
  CU(ksk_relayout_launch(d_ksk_ref, reinterpret_cast<uint32_t *>(e->blob + e->off_ksk), e->ksk_rows,
                         p.n, e->ksk_stride, e->stream));
  e->launches += 4;
  // key.rs:91-100 (test vector) and :78-89 (decomposition offset)
  std::vector<uint32_t> tv(2 * TFHE_N, 0u);
  for (int x = 0; x < TFHE_N; x++) tv[TFHE_N + x] = 0x20000000u;
  CU(cudaMemsetAsync(e->tv(), 0, (size_t)kMaxLut * 2 * TFHE_N * 4, e->stream));
  CU(cudaMemcpyAsync(e->tv(), tv.data(), tv.size() * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->s_misc.release();
  uint32_t offset = 0;
  for (uint32_t i = 0; i < p.l; i++) offset += (1u << (p.bgbit - 1)) << (32 - (i + 1) * p.bgbit);
  e->decomp_offset = offset;
  e->n_lut = 1;
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return TFHE_OK;
}

int tfhe_engine_alloc_cloud_key(tfhe_engine *e) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  return ensure_blob(e);
}

int tfhe_engine_cloud_key_blob(tfhe_engine *e, void **device_ptr, size_t *bytes) {
  if (!e || !device_ptr || !bytes) return fail(TFHE_ERR_INVALID, "null argument");
  if (!e->blob) return fail(TFHE_ERR_NO_KEY, "cloud key blob not allocated");
  *device_ptr = e->blob;
  *bytes = e->blob_bytes;
  return TFHE_OK;
}

int tfhe_engine_commit_cloud_key(tfhe_engine *e, uint32_t decomposition_offset) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (!e->blob) return fail(TFHE_ERR_NO_KEY, "cloud key blob not allocated");
  std::lock_guard<std::mutex> lock(e->mu);
  e->decomp_offset = decomposition_offset;
  e->n_lut = 1;
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return TFHE_OK;
}

namespace {
// Blob format 2: BSK | KSK rows | test-vector slots (format 1 also carried the mma.sync fragment copy
// of the KSK, now derived on demand like the other kernel-specific key orders).
constexpr uint32_t kBlobFormat = 2;
struct BlobHeader {
  char magic[8];
  uint32_t version, n, N, l, bgbit, basebit, iks_t, decomposition_offset, n_lut, reserved;
  uint64_t payload_bytes;
  uint8_t pad[8];
};
static_assert(sizeof(BlobHeader) == 64, "header is 64 bytes");
}  // namespace

size_t tfhe_engine_cloud_key_export_bytes(tfhe_engine *e) {
  return e ? sizeof(BlobHeader) + e->blob_bytes : 0;
}

int tfhe_engine_export_cloud_key(tfhe_engine *e, void *host_buf, size_t bytes) {
  if (!e || !host_buf) return fail(TFHE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  if (bytes < sizeof(BlobHeader) + e->blob_bytes) return fail(TFHE_ERR_INVALID, "buffer too small");
  CU(cudaSetDevice(e->dev));
  BlobHeader hd{};
  memcpy(hd.magic, "TFHEB200", 8);
  hd.version = kBlobFormat;
  hd.n = e->p.n; hd.N = e->p.N; hd.l = e->p.l; hd.bgbit = e->p.bgbit; hd.basebit = e->p.basebit;
  hd.iks_t = e->p.iks_t; hd.decomposition_offset = e->decomp_offset; hd.n_lut = (uint32_t)e->n_lut;
  hd.payload_bytes = e->blob_bytes;
  memcpy(host_buf, &hd, sizeof(hd));
  CU(cudaMemcpyAsync(static_cast<uint8_t *>(host_buf) + sizeof(hd), e->blob, e->blob_bytes,
                     cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return TFHE_OK;
}

int tfhe_engine_import_cloud_key(tfhe_engine *e, const void *host_buf, size_t bytes) {
  if (!e || !host_buf) return fail(TFHE_ERR_INVALID, "null argument");
  if (bytes < sizeof(BlobHeader)) return fail(TFHE_ERR_INVALID, "truncated blob");
  BlobHeader hd;
  memcpy(&hd, host_buf, sizeof(hd));
  if (memcmp(hd.magic, "TFHEB200", 8) != 0 || hd.version != kBlobFormat)
    return fail(TFHE_ERR_INVALID, "not a tfhe_b200 key blob (or wrong version)");
  const tfhe_params &p = e->p;
  if (hd.n != p.n || hd.N != p.N || hd.l != p.l || hd.bgbit != p.bgbit || hd.basebit != p.basebit ||
      hd.iks_t != p.iks_t)
    return fail(TFHE_ERR_INVALID, "blob parameters differ from the engine's");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  int rc = ensure_blob(e);
  if (rc != TFHE_OK) return rc;
  if (hd.payload_bytes != e->blob_bytes || bytes < sizeof(hd) + hd.payload_bytes)
    return fail(TFHE_ERR_INVALID, "blob size mismatch");
  if (hd.n_lut < 1 || hd.n_lut > (uint32_t)kMaxLut) return fail(TFHE_ERR_INVALID, "bad lut count");
  CU(cudaMemcpyAsync(e->blob, static_cast<const uint8_t *>(host_buf) + sizeof(hd), e->blob_bytes,
                     cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->decomp_offset = hd.decomposition_offset;
  e->n_lut = (int)hd.n_lut;
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return TFHE_OK;
}

int tfhe_batch_gate(tfhe_engine *e, tfhe_gate op, const uint32_t *in_pairs, uint32_t *out,
                    size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if ((int)op < 0 || (int)op >= TFHE_GATE_COUNT) return fail(TFHE_ERR_INVALID, "bad gate %d", (int)op);
  const size_t w = e->p.n + 1;
  return run_host(e, (int)op, nullptr, -1, in_pairs, 2 * w, out, w, count, 0);
}

int tfhe_batch_gate_mixed(tfhe_engine *e, const uint8_t *ops, const uint32_t *in_pairs,
                          uint32_t *out, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (!ops && count) return fail(TFHE_ERR_INVALID, "null ops");
  for (size_t i = 0; i < count; i++)
    if (ops[i] >= TFHE_GATE_COUNT) return fail(TFHE_ERR_INVALID, "bad gate %d at %zu", ops[i], i);
  const size_t w = e->p.n + 1;
  return run_host(e, 0, ops, -1, in_pairs, 2 * w, out, w, count, 0);
}

int tfhe_batch_bootstrap(tfhe_engine *e, const uint32_t *in, uint32_t *out, size_t count,
                         int key_switch) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  const size_t w = e->p.n + 1;
  return run_host(e, -1, nullptr, -1, in, w, out, w, count, key_switch ? 0 : 1);
}

int tfhe_batch_blind_rotate(tfhe_engine *e, const uint32_t *in, uint32_t *out_trlwe, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  return run_host(e, -1, nullptr, -1, in, e->p.n + 1, out_trlwe, 2 * TFHE_N, count, 2);
}

int tfhe_lut_register(tfhe_engine *e, const uint32_t *poly_a, const uint32_t *poly_b,
                      int *lut_id_out) {
  if (!e || !poly_b || !lut_id_out) return fail(TFHE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->blob) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  if (e->n_lut >= kMaxLut) return fail(TFHE_ERR_ALLOC, "out of LUT slots (%d)", kMaxLut);
  CU(cudaSetDevice(e->dev));
  uint32_t *slot = e->tv() + (size_t)e->n_lut * 2 * TFHE_N;
  if (poly_a) CU(cudaMemcpyAsync(slot, poly_a, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  else CU(cudaMemsetAsync(slot, 0, TFHE_N * 4, e->stream));
  CU(cudaMemcpyAsync(slot + TFHE_N, poly_b, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *lut_id_out = e->n_lut++;
  return TFHE_OK;
}

int tfhe_lut_generate(tfhe_engine *e, const uint32_t *f_table, uint32_t modulus, double scale,
                      uint32_t *lut_b_out, int *lut_id_out) {
  if (!e || !f_table || !lut_id_out) return fail(TFHE_ERR_INVALID, "null argument");
  if (modulus == 0 || modulus > TFHE_N) return fail(TFHE_ERR_INVALID, "bad modulus %u", modulus);
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->blob) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  if (e->n_lut >= kMaxLut) return fail(TFHE_ERR_ALLOC, "out of LUT slots (%d)", kMaxLut);
  CU(cudaSetDevice(e->dev));
  if (scale <= 0.0) scale = 1.0 / (2.0 * (double)modulus);  // lut/encoder.rs:36
  CU(e->s_misc.reserve((size_t)modulus * 4));
  CU(cudaMemcpyAsync(e->s_misc.p, f_table, (size_t)modulus * 4, cudaMemcpyHostToDevice, e->stream));
  uint32_t *slot = e->tv() + (size_t)e->n_lut * 2 * TFHE_N;
  CU(lut_generate_launch(static_cast<const uint32_t *>(e->s_misc.p), modulus, scale, slot, e->stream));
  e->launches++;
  if (lut_b_out)
    CU(cudaMemcpyAsync(lut_b_out, slot + TFHE_N, TFHE_N * 4, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *lut_id_out = e->n_lut++;
  return TFHE_OK;
}

int tfhe_batch_bootstrap_lut(tfhe_engine *e, int lut_id, const uint32_t *in, uint32_t *out,
                             size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (lut_id < 0 || lut_id >= e->n_lut) return fail(TFHE_ERR_INVALID, "unknown lut id %d", lut_id);
  const size_t w = e->p.n + 1;
  return run_host(e, -1, nullptr, lut_id, in, w, out, w, count, 0);
}

int tfhe_batch_bootstrap_lut_multi(tfhe_engine *e, const int32_t *lut_ids, const uint32_t *in,
                                   uint32_t *out, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (!lut_ids && count) return fail(TFHE_ERR_INVALID, "null lut_ids");
  for (size_t i = 0; i < count; i++)
    if (lut_ids[i] < 0 || lut_ids[i] >= e->n_lut)
      return fail(TFHE_ERR_INVALID, "unknown lut id %d at %zu", lut_ids[i], i);
  const size_t w = e->p.n + 1;
  return run_host(e, -1, nullptr, 0, in, w, out, w, count, 0, lut_ids);
}

int tfhe_batch_extract_key_switch(tfhe_engine *e, const uint32_t *in_trlwe, uint32_t *out,
                                  size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (count == 0) return TFHE_OK;
  if (!in_trlwe || !out) return fail(TFHE_ERR_INVALID, "null buffer");
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  CU(cudaSetDevice(e->dev));
  const size_t w = e->p.n + 1;
  for (size_t base = 0; base < count; base += kChunk) {
    size_t c = count - base < kChunk ? count - base : kChunk;
    tfhe_engine::Slot &sl = e->slot[0];
    CU(sl.in.reserve(c * 2 * TFHE_N * 4));
    CU(sl.ext.reserve(c * (TFHE_N + 1) * 4));
    CU(sl.out.reserve(c * w * 4));
    CU(cudaMemcpyAsync(sl.in.p, in_trlwe + base * 2 * TFHE_N, c * 2 * TFHE_N * 4,
                       cudaMemcpyHostToDevice, e->stream));
    CU(extract_launch(static_cast<const uint32_t *>(sl.in.p), static_cast<uint32_t *>(sl.ext.p), c,
                      e->stream));
    int rc = key_switch(e, static_cast<const uint32_t *>(sl.ext.p), static_cast<uint32_t *>(sl.out.p), c);
    if (rc != TFHE_OK) return rc;
    e->launches += 1;
    CU(cudaMemcpyAsync(out + base * w, sl.out.p, c * w * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  return TFHE_OK;
}

int tfhe_reenc_key_load(tfhe_engine *e, const uint32_t *key_encryptions, uint32_t base, uint32_t t,
                        tfhe_reenc_key **out) {
  if (!e || !key_encryptions || !out) return fail(TFHE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (base < 2 || (base & (base - 1)) != 0) return fail(TFHE_ERR_INVALID, "base must be a power of two");
  uint32_t basebit = 0;
  while ((1u << basebit) < base) basebit++;
  if (t == 0 || basebit * t > 31) return fail(TFHE_ERR_INVALID, "bad decomposition (basebit*t > 31)");
  if ((size_t)8 * e->p.n * 4 > 48 * 1024) return fail(TFHE_ERR_INVALID, "n too large");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  tfhe_reenc_key *k = new (std::nothrow) tfhe_reenc_key();
  if (!k) return fail(TFHE_ERR_ALLOC, "out of host memory");
  k->basebit = basebit; k->t = t; k->n_rows = base * t * e->p.n; k->dev = e->dev;
  const size_t src_bytes = (size_t)k->n_rows * (e->p.n + 1) * 4;
  const size_t dst_bytes = ((size_t)k->n_rows + 1) * e->ksk_stride * 4;
  auto upload = [&]() -> int {
    CU(cudaMalloc(reinterpret_cast<void **>(&k->rows), dst_bytes));
    CU(e->s_misc.reserve(src_bytes));
    CU(cudaMemcpyAsync(e->s_misc.p, key_encryptions, src_bytes, cudaMemcpyHostToDevice, e->stream));
    CU(ksk_relayout_launch(static_cast<const uint32_t *>(e->s_misc.p), k->rows, k->n_rows, e->p.n,
                           e->ksk_stride, e->stream));
    e->launches++;
    CU(cudaStreamSynchronize(e->stream));
    return TFHE_OK;
  };
  int rc = upload();
  e->s_misc.release();
  if (rc != TFHE_OK) {
    if (k->rows) cudaFree(k->rows);
    delete k;
    return rc;
  }
  *out = k;
  return TFHE_OK;
}

void tfhe_reenc_key_destroy(tfhe_reenc_key *k) {
  if (!k) return;
  cudaSetDevice(k->dev);
  if (k->rows) cudaFree(k->rows);
  delete k;
}

int tfhe_batch_reencrypt(tfhe_engine *e, const tfhe_reenc_key *key, const uint32_t *in, uint32_t *out,
                         size_t count) {
  if (!e || !key) return fail(TFHE_ERR_INVALID, "null argument");
  if (count == 0) return TFHE_OK;
  if (!in || !out) return fail(TFHE_ERR_INVALID, "null buffer");
  if (key->dev != e->dev) return fail(TFHE_ERR_INVALID, "key lives on another device");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  const size_t w = e->p.n + 1;
  for (size_t base = 0; base < count; base += kChunk) {
    size_t c = count - base < kChunk ? count - base : kChunk;
    tfhe_engine::Slot &sl = e->slot[0];
    CU(sl.in.reserve(c * w * 4));
    CU(sl.out.reserve(c * w * 4));
    CU(cudaMemcpyAsync(sl.in.p, in + base * w, c * w * 4, cudaMemcpyHostToDevice, e->stream));
    KsArgs k{};
    k.ksk = key->rows; k.ext = static_cast<const uint32_t *>(sl.in.p);
    k.out = static_cast<uint32_t *>(sl.out.p);
    k.n = e->p.n; k.basebit = key->basebit; k.iks_t = key->t;
    k.stride = e->ksk_stride; k.zero_row = key->n_rows; k.n_in = e->p.n; k.count = c;
    CU(ks_launch(k, e->stream));
    e->launches++;
    CU(cudaMemcpyAsync(out + base * w, sl.out.p, c * w * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
  }
  return TFHE_OK;
}

int tfhe_batch_gate_dev(tfhe_engine *e, tfhe_gate op, const uint8_t *d_ops,
                        const uint32_t *d_in_pairs, uint32_t *d_out, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (!d_ops && ((int)op < 0 || (int)op >= TFHE_GATE_COUNT))
    return fail(TFHE_ERR_INVALID, "bad gate %d", (int)op);
  if (count == 0) return TFHE_OK;
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  const size_t w = e->p.n + 1;
  for (size_t base = 0; base < count; base += kChunk) {
    size_t c = count - base < kChunk ? count - base : kChunk;
    int rc = run_device(e, e->slot[0], d_ops ? 0 : (int)op, d_ops ? d_ops + base : nullptr, -1,
                        d_in_pairs + base * 2 * w, d_out + base * w, c, 0);
    if (rc != TFHE_OK) return rc;
  }
  return TFHE_OK;
}

int tfhe_batch_bootstrap_dev(tfhe_engine *e, int lut_id, const uint32_t *d_in, uint32_t *d_out,
                             size_t count, int key_switch) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (count == 0) return TFHE_OK;
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  const size_t w = e->p.n + 1;
  for (size_t base = 0; base < count; base += kChunk) {
    size_t c = count - base < kChunk ? count - base : kChunk;
    int rc = run_device(e, e->slot[0], -1, nullptr, lut_id, d_in + base * w, d_out + base * w, c,
                        key_switch ? 0 : 1);
    if (rc != TFHE_OK) return rc;
  }
  return TFHE_OK;
}

int tfhe_engine_synchronize(tfhe_engine *e) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  CU(cudaSetDevice(e->dev));
  CU(cudaStreamSynchronize(e->stream));
  float t0 = 0.f, t1 = 0.f;
  if (cudaEventElapsedTime(&t0, e->slot[0].br_start, e->slot[0].br_end) == cudaSuccess &&
      cudaEventElapsedTime(&t1, e->slot[0].br_end, e->slot[0].ks_end) == cudaSuccess) {
    e->last_ms[0] = t0; e->last_ms[1] = t1;
  } else {
    cudaGetLastError();
  }
  return TFHE_OK;
}

}  // extern "C"
