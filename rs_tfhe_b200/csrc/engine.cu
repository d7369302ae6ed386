// engine.cu -- the C ABI (include/tfhe_b200.h): engine handle, cloud-key upload and
// re-layout, batch entry points.  No CPU fallback: every compute entry point
// needs a CUDA device and fails loudly (TFHE_ERR_CUDA) without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>
#include <unistd.h>

#include "kernels.h"
#include "brs_core.cuh"

#ifndef TFHE_KS_DEFAULT
#define TFHE_KS_DEFAULT 0   // 0 = tcgen05 (umma), 2 = row walk
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

// NVTX ranges (header-only NVTX v3; no-ops unless a profiler is attached): key upload / broadcast,
// and per pipeline chunk H2D, blind rotation (K0+K3), key switch (K4), D2H -- the timeline rows an
// nsys capture of a batch call shows.
struct Nvtx {
  explicit Nvtx(const char *name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
};

constexpr int kMaxLut = 64;          // test-vector slots (slot 0 = cloud-key test vector,
                                     // slot kMaxLut-1 = scratch of the ephemeral bootstrap_func path)
constexpr int kScratchLut = kMaxLut - 1;
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

struct PinnedScratch {
  void *p = nullptr;
  size_t bytes = 0;
  cudaError_t reserve(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocDefault);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
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
  cplx *bsk3 = nullptr;   // BSK rows in the 128-thread kernel's order (derived, not in the blob)
  cplx *tw_s = nullptr;   // per-thread constants of the 128-thread kernel
  uint8_t *kumma = nullptr;  // KSK as tcgen05 operand tiles (derived from the blob's KSK rows; gate sets)
  size_t blob_bytes = 0, off_ksk = 0, off_tv = 0;
  bool key_loaded = false;
  uint32_t decomp_offset = 0;
  uint32_t ksk_rows = 0, ksk_stride = 0;
  // LUT ids handed out = (key epoch << 8) | slot, so ids from before a key (re)load are rejected;
  // id 0 is always the cloud-key test vector.
  bool lut_used[kMaxLut] = {true};
  uint32_t key_epoch = 0;
  cplx *tw_a = nullptr, *tw_b = nullptr;
  Scratch s_misc;
  // two pipeline slots: H2D of chunk k+1 and D2H of chunk k-1 overlap the kernels of chunk k
  struct Slot {
    Scratch in, ext, out, ops, idx;
    // pinned staging for callers whose buffers are pageable (a Rust Vec<Ciphertext> is): the host thread
    // copies chunk k+1 into hin while the GPU runs chunk k, and drains hout after the chunk's D2H
    PinnedScratch hin, hout;
    void *drain_dst = nullptr;       // pending hout -> caller copy
    size_t drain_bytes = 0;
    cudaEvent_t h2d_done = nullptr, br_start = nullptr, br_end = nullptr, ks_end = nullptr,
                d2h_done = nullptr;
    bool busy = false;
  } slot[2];
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_ms[2] = {0.f, 0.f};
  float dev_ms[2] = {0.f, 0.f};   // device-pointer path: kernel time of the chunks already retired in this call
  uint64_t launches = 0;
  // multi-GPU (tfhe_engine_create_multi): this engine is device_ids[0]; peers are the other devices'
  // engines, owned here.  The cloud key is broadcast root -> peers with NCCL at key-load time; batch
  // calls shard contiguous index ranges over all of them, one host thread per GPU.
  std::vector<tfhe_engine *> peers;
  void *nccl_comms = nullptr;   // ncclComm_t[1 + peers.size()]
  float last_broadcast_ms = 0.f;

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

// K4 kernel selection: TFHE_KS_VARIANT = umma (default: tcgen05.mma, TMEM accumulators) | rows
// (row-walk kernel; always used when the tcgen05 shape does not cover the set, e.g. basebit 7).
enum { KS_UMMA = 0, KS_ROWS = 2 };
int ks_variant() {
  static const int v = [] {
    const char *s = getenv("TFHE_KS_VARIANT");
    if (!s || !s[0]) return (int)TFHE_KS_DEFAULT;
    return s[0] == 'r' ? (int)KS_ROWS : (int)KS_UMMA;
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
  if (br_uses_s_key()) {
    const size_t rows = (size_t)e->p.n * 2 * e->p.l;
    if (!e->bsk3) CU(cudaMalloc(reinterpret_cast<void **>(&e->bsk3), rows * brs::kRowCplx * sizeof(cplx)));
    CU(bsk_permute_s_launch(e->bsk(), e->bsk3, rows, e->stream));
    e->launches++;
  }
  if (ks_variant() == KS_UMMA && ks_umma_supported(e->p.basebit, e->p.iks_t)) {
    if (!e->kumma) CU(cudaMalloc(reinterpret_cast<void **>(&e->kumma), ks_umma_key_bytes(e->p.n, e->p.iks_t, e->p.basebit)));
    CU(ksk_umma_relayout_launch(e->ksk(), e->ksk_stride, e->kumma, e->p.n, e->p.iks_t, e->p.basebit, e->stream));
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

// LUT ids: decode to a slot, or -1 when the id is stale / unknown
int lut_slot(const tfhe_engine *e, int lut_id) {
  if (lut_id == 0) return 0;
  if (lut_id < 0) return -1;
  const int slot = lut_id & 0xff;
  if ((uint32_t)(lut_id >> 8) != (e->key_epoch & 0x7fffffu)) return -1;
  if (slot <= 0 || slot >= kScratchLut || !e->lut_used[slot]) return -1;
  return slot;
}
int lut_make_id(const tfhe_engine *e, int slot) { return (int)((e->key_epoch & 0x7fffffu) << 8) | slot; }
int lut_take_slot(tfhe_engine *e) {
  for (int s = 1; s < kScratchLut; s++)
    if (!e->lut_used[s]) { e->lut_used[s] = true; return s; }
  return -1;
}
// a (re)loaded key invalidates every table id handed out before
void lut_reset(tfhe_engine *e, int n_used = 1) {
  for (int s = 0; s < kMaxLut; s++) e->lut_used[s] = s < n_used;
  e->key_epoch++;
}
int lut_count(const tfhe_engine *e) {
  int hi = 0;
  for (int s = 0; s < kScratchLut; s++) if (e->lut_used[s]) hi = s + 1;
  return hi;
}

// K4 dispatch: tensor-core GEMM for the gate sets, row-walk kernels otherwise
int key_switch(tfhe_engine *e, const uint32_t *d_ext, uint32_t *d_out, size_t count) {
  const int variant = ks_variant();
  // latency path: up to 48 ciphertexts (a dependent chain's level; measured crossover with the tensor-core kernel ~70)
  // split each ciphertext's sum over the whole GPU (TFHE_KS_SMALL_MAX overrides)
  static const size_t small_max = [] { const char *v = getenv("TFHE_KS_SMALL_MAX"); return v ? (size_t)atoi(v) : (size_t)48; }();
  // (parameter sets the tensor-core kernel does not cover -- basebit 7 -- would fall to the row walk, whose
  // latency is milliseconds at any batch size: for them the split-sum kernel stays ahead up to ~2000 ciphertexts)
  const bool no_umma = variant == KS_UMMA && !e->kumma;
  if ((count <= small_max || (no_umma && small_max > 0 && count <= 2048)) && variant == KS_UMMA) {
    KsArgs k{};
    k.ksk = e->ksk(); k.ext = d_ext; k.out = d_out;
    k.n = e->p.n; k.basebit = e->p.basebit; k.iks_t = e->p.iks_t;
    k.stride = e->ksk_stride; k.zero_row = e->ksk_rows; k.n_in = TFHE_N; k.count = count;
    CU(ks_small_launch(k, e->num_sms, e->stream));
  } else if (variant == KS_UMMA && e->kumma) {
    KsUmmaArgs k{};
    k.key = e->kumma; k.ext = d_ext; k.out = d_out;
    k.n = e->p.n; k.iks_t = e->p.iks_t; k.basebit = e->p.basebit; k.count = count;
    CU(ks_umma_launch(k, e->stream));
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
//   out_kind: 0 key-switched LWE [n+1]; 1 extract_2 [n+1]; 2 TRLWE [2][N]; 3 extracted level-1 LWE [N+1]
int run_device(tfhe_engine *e, tfhe_engine::Slot &sl, int op, const uint8_t *d_ops, int lut_id,
               const uint32_t *d_in, uint32_t *d_out, size_t count, int out_kind,
               const int32_t *d_lut_ids = nullptr) {
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  if (lut_id >= kMaxLut) return fail(TFHE_ERR_INVALID, "unknown lut slot %d", lut_id);
  BrArgs a{};
  a.bsk = e->bsk();
  a.bsk2 = e->bsk2;
  a.bsk3 = e->bsk3; a.tw_s = e->tw_s;
  a.tw_a = e->tw_a; a.tw_b = e->tw_b;
  a.tv = e->tv(); a.tv_index = d_lut_ids; a.tv_default = lut_id < 0 ? 0 : lut_id;
  a.in = d_in; a.ops = d_ops; a.op = op;
  a.n = e->p.n; a.offset = e->decomp_offset; a.count = count;
  if (out_kind == 0) {
    CU(sl.ext.reserve(count * (TFHE_N + 1) * 4));
    a.out = static_cast<uint32_t *>(sl.ext.p);
    a.out_mode = BR_OUT_EXTRACT;
  } else {
    a.out = d_out;   // 1: sample_extract_index_2 image [n+1]; 2: TRLWE; 3: level-1 sample [N+1], no key switch
    a.out_mode = out_kind == 1 ? BR_OUT_EXTRACT2 : out_kind == 3 ? BR_OUT_EXTRACT : BR_OUT_TRLWE;
  }
  CU(cudaEventRecord(sl.br_start, e->stream));
  {
    Nvtx r("tfhe:K0+K3 blind_rotate");
    int n_br = 0;
    CU(br_launch(e->p.l, e->p.bgbit, a, e->num_sms, e->stream, &n_br));
    e->launches += (uint64_t)n_br;
  }
  CU(cudaEventRecord(sl.br_end, e->stream));
  if (out_kind == 0) {
    Nvtx r("tfhe:K4 key_switch");
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
  if (sl.drain_dst) {
    memcpy(sl.drain_dst, sl.hout.p, sl.drain_bytes);
    sl.drain_dst = nullptr;
  }
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
int run_host_locked(tfhe_engine *e, int op, const uint8_t *ops, int lut_id, const uint32_t *in,
                    size_t in_words, uint32_t *out, size_t out_words, size_t count, int out_kind,
                    const int32_t *lut_ids = nullptr);
int run_host(tfhe_engine *e, int op, const uint8_t *ops, int lut_id, const uint32_t *in,
             size_t in_words, uint32_t *out, size_t out_words, size_t count, int out_kind,
             const int32_t *lut_ids = nullptr) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (count == 0) return TFHE_OK;
  if (!in || !out) return fail(TFHE_ERR_INVALID, "null buffer");
  std::lock_guard<std::mutex> lock(e->mu);
  return run_host_locked(e, op, ops, lut_id, in, in_words, out, out_words, count, out_kind, lut_ids);
}
int run_host_locked(tfhe_engine *e, int op, const uint8_t *ops, int lut_id, const uint32_t *in,
                    size_t in_words, uint32_t *out, size_t out_words, size_t count, int out_kind,
                    const int32_t *lut_ids) {
  if (count == 0) return TFHE_OK;
  if (!in || !out) return fail(TFHE_ERR_INVALID, "null buffer");
  CU(cudaSetDevice(e->dev));
  if (!e->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  float ms0 = 0.f, ms1 = 0.f;
  // Chunks are whole rounds of the persistent grid (4 ciphertexts per SM): 28 rounds in steady
  // state, but a short first chunk (its upload is the only one no kernel hides) and a short last
  // one (likewise its download).
  const size_t round = (size_t)e->num_sms * 4;
  const size_t chunk = round * 28, edge = round * 4;
  // Pageable caller buffers (the normal case for a drop-in caller) go through the pinned staging ring
  // when the call spans several chunks; a pinned / registered buffer is copied from directly.
  auto pageable = [](const void *ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
  };
  const bool stage_in = count > 2 * edge && pageable(in);
  const bool stage_out = count > 2 * edge && pageable(out);
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
    Nvtx r_chunk("tfhe:chunk");
    nvtxRangePushA("tfhe:H2D");
    const void *h_src = in + base * in_words;
    if (stage_in) {
      CU(sl.hin.reserve(chunk * in_words * 4));
      memcpy(sl.hin.p, h_src, c * in_words * 4);   // overlaps the previous chunk's kernels
      h_src = sl.hin.p;
    }
    CU(cudaMemcpyAsync(sl.in.p, h_src, c * in_words * 4, cudaMemcpyHostToDevice, e->copy_in));
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
    nvtxRangePop();
    CU(cudaStreamWaitEvent(e->stream, sl.h2d_done, 0));
    rc = run_device(e, sl, op, d_ops, lut_id, static_cast<const uint32_t *>(sl.in.p),
                    static_cast<uint32_t *>(sl.out.p), c, out_kind, d_ids);
    if (rc != TFHE_OK) return rc;
    Nvtx r_d2h("tfhe:D2H");
    CU(cudaStreamWaitEvent(e->copy_out, sl.ks_end, 0));
    void *h_dst = out + base * out_words;
    if (stage_out) {
      CU(sl.hout.reserve(chunk * out_words * 4));
      sl.drain_dst = h_dst;
      sl.drain_bytes = c * out_words * 4;
      h_dst = sl.hout.p;
    }
    CU(cudaMemcpyAsync(h_dst, sl.out.p, c * out_words * 4, cudaMemcpyDeviceToHost, e->copy_out));
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

// ---- NCCL, resolved at run time ----------------------------------------------------------------
// The library does not link libnccl: a one-GPU caller needs none, and a host process that already
// carries its own NCCL (e.g. PyTorch) must keep using that copy.  tfhe_engine_create_multi resolves
// the five entry points it needs from the copy already in the process, else from libnccl.so.2.
namespace {
struct NcclApi {
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
const NcclApi &nccl_api() {
  static const NcclApi api = [] {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(h, "ncclBroadcast"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.CommInitAll && a.CommDestroy && a.Broadcast && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}
#define NC(call)                                                                              \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess)                                                                    \
      return fail(TFHE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r_), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

// After the root engine's blob is final: one ncclBroadcast of the re-laid-out key (BSK | KSK | test
// vectors) to every peer over NVLink, then each peer derives its kernel-specific key orders.
int broadcast_key(tfhe_engine *e, int n_lut_used) {
  if (e->peers.empty()) return TFHE_OK;
  Nvtx r("tfhe:ncclBroadcast cloud key");
  const NcclApi &nc = nccl_api();
  ncclComm_t *comms = static_cast<ncclComm_t *>(e->nccl_comms);
  for (tfhe_engine *p : e->peers) {
    std::lock_guard<std::mutex> lock(p->mu);
    CU(cudaSetDevice(p->dev));
    int rc = ensure_blob(p);
    if (rc != TFHE_OK) return rc;
    if (p->blob_bytes != e->blob_bytes) return fail(TFHE_ERR_INVALID, "peer blob size differs");
  }
  CU(cudaSetDevice(e->dev));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventRecord(e->ev[0], e->stream));
  NC(nc.GroupStart());
  NC(nc.Broadcast(e->blob, e->blob, e->blob_bytes, ncclUint8, 0, comms[0], e->stream));
  for (size_t i = 0; i < e->peers.size(); i++) {
    tfhe_engine *p = e->peers[i];
    NC(nc.Broadcast(p->blob, p->blob, p->blob_bytes, ncclUint8, 0, comms[i + 1], p->stream));
  }
  NC(nc.GroupEnd());
  CU(cudaEventRecord(e->ev[1], e->stream));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventElapsedTime(&e->last_broadcast_ms, e->ev[0], e->ev[1]));
  for (tfhe_engine *p : e->peers) {
    std::lock_guard<std::mutex> lock(p->mu);
    CU(cudaSetDevice(p->dev));
    CU(cudaStreamSynchronize(p->stream));
    p->decomp_offset = e->decomp_offset;
    lut_reset(p, n_lut_used);
    p->key_epoch = e->key_epoch;          // table ids are valid on every device
    for (int sidx = 0; sidx < kMaxLut; sidx++) p->lut_used[sidx] = e->lut_used[sidx];
    int rc = finalize_key(p);
    if (rc != TFHE_OK) return rc;
    p->key_loaded = true;
  }
  CU(cudaSetDevice(e->dev));
  return TFHE_OK;
}

// Contiguous shards [r*count/G, (r+1)*count/G), one host thread per peer GPU; the root's
// shard runs on the calling thread.  fn(engine, base, n) -> status.
template <class F> int shard(tfhe_engine *e, size_t count, F fn) {
  if (e->peers.empty()) return fn(e, (size_t)0, count);
  const size_t G = e->peers.size() + 1;
  std::vector<int> rcs(G, TFHE_OK);
  std::vector<std::string> errs(G);
  std::vector<std::thread> th;
  auto range = [&](size_t r, size_t &b, size_t &n) { b = r * count / G; n = (r + 1) * count / G - b; };
  for (size_t r = 1; r < G; r++)
    th.emplace_back([&, r] {
      size_t b, n;
      range(r, b, n);
      rcs[r] = fn(e->peers[r - 1], b, n);
      if (rcs[r] != TFHE_OK) errs[r] = g_err;
    });
  size_t b0, n0;
  range(0, b0, n0);
  rcs[0] = fn(e, b0, n0);
  for (auto &t : th) t.join();
  float br = e->last_ms[0], ks = e->last_ms[1];
  for (tfhe_engine *p : e->peers) { br = br > p->last_ms[0] ? br : p->last_ms[0]; ks = ks > p->last_ms[1] ? ks : p->last_ms[1]; }
  e->last_ms[0] = br; e->last_ms[1] = ks;      // slowest device
  for (size_t r = 1; r < G; r++)
    if (rcs[r] != TFHE_OK) return fail(rcs[r], "device %d: %s", e->peers[r - 1]->dev, errs[r].c_str());
  return rcs[0];
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
  {  // per-thread constants of the 128-thread kernel (brs_core.cuh)
    std::vector<cplx> tws((size_t)brs::kT * brs::kTwPerThread);
    for (int t = 0; t < brs::kT; t++) {
      cplx tw[brs::kTwPerThread];
      brs::make_tw(t, tw);
      for (int k = 0; k < brs::kTwPerThread; k++) tws[(size_t)t * brs::kTwPerThread + k] = tw[k];
    }
    CU(cudaMalloc(reinterpret_cast<void **>(&e->tw_s), tws.size() * sizeof(cplx)));
    CU(cudaMemcpy(e->tw_s, tws.data(), tws.size() * sizeof(cplx), cudaMemcpyHostToDevice));
  }
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

int tfhe_engine_create_multi(const tfhe_params *params, const int *device_ids, int n_devices,
                             tfhe_engine **out) {
  if (!params || !device_ids || !out) return fail(TFHE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_devices < 1) return fail(TFHE_ERR_INVALID, "need at least one device");
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < i; j++)
      if (device_ids[i] == device_ids[j]) return fail(TFHE_ERR_INVALID, "device %d listed twice", device_ids[i]);
  tfhe_engine *root = nullptr;
  int rc = tfhe_engine_create(params, device_ids[0], &root);
  if (rc != TFHE_OK) return rc;
  if (n_devices == 1) { *out = root; return TFHE_OK; }
  auto bail = [&](int code) {
    char saved[sizeof(g_err)];
    memcpy(saved, g_err, sizeof(saved));
    tfhe_engine_destroy(root);
    memcpy(g_err, saved, sizeof(saved));
    return code;
  };
  if (!nccl_api().ok) {
    fail(TFHE_ERR_CUDA, "NCCL not found (libnccl.so.2): needed for the multi-GPU cloud-key broadcast");
    return bail(TFHE_ERR_CUDA);
  }
  for (int i = 1; i < n_devices; i++) {
    tfhe_engine *p = nullptr;
    rc = tfhe_engine_create(params, device_ids[i], &p);
    if (rc != TFHE_OK) return bail(rc);
    root->peers.push_back(p);
  }
  ncclComm_t *comms = new (std::nothrow) ncclComm_t[n_devices]();
  if (!comms) { fail(TFHE_ERR_ALLOC, "out of host memory"); return bail(TFHE_ERR_ALLOC); }
  root->nccl_comms = comms;
  ncclResult_t nr = nccl_api().CommInitAll(comms, n_devices, device_ids);
  if (nr != ncclSuccess) {
    fail(TFHE_ERR_CUDA, "ncclCommInitAll failed: %s", nccl_api().GetErrorString(nr));
    delete[] comms;
    root->nccl_comms = nullptr;
    return bail(TFHE_ERR_CUDA);
  }
  cudaSetDevice(device_ids[0]);
  *out = root;
  return TFHE_OK;
}

int tfhe_engine_device_count(const tfhe_engine *e) { return e ? (int)e->peers.size() + 1 : 0; }

int tfhe_engine_last_broadcast_ms(tfhe_engine *e, float *ms_out) {
  if (!e || !ms_out) return fail(TFHE_ERR_INVALID, "null argument");
  *ms_out = e->last_broadcast_ms;
  return TFHE_OK;
}

void tfhe_engine_destroy(tfhe_engine *e) {
  if (!e) return;
  if (e->nccl_comms) {
    ncclComm_t *comms = static_cast<ncclComm_t *>(e->nccl_comms);
    for (size_t i = 0; i < e->peers.size() + 1; i++)
      if (comms[i]) nccl_api().CommDestroy(comms[i]);
    delete[] comms;
    e->nccl_comms = nullptr;
  }
  for (tfhe_engine *p : e->peers) tfhe_engine_destroy(p);
  e->peers.clear();
  cudaSetDevice(e->dev);
  cudaDeviceSynchronize();
  if (e->blob) cudaFree(e->blob);
  if (e->bsk2) cudaFree(e->bsk2);
  if (e->bsk3) cudaFree(e->bsk3);
  if (e->tw_s) cudaFree(e->tw_s);
  if (e->kumma) cudaFree(e->kumma);
  if (e->tw_a) cudaFree(e->tw_a);
  if (e->tw_b) cudaFree(e->tw_b);
  e->s_misc.release();
  for (auto &sl : e->slot) {
    sl.in.release(); sl.ext.release(); sl.out.release(); sl.ops.release(); sl.idx.release();
    sl.hin.release(); sl.hout.release();
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

uint64_t tfhe_engine_kernel_launches(const tfhe_engine *e) {
  if (!e) return 0;
  uint64_t n = e->launches;
  for (const tfhe_engine *p : e->peers) n += p->launches;
  return n;
}

int tfhe_engine_last_kernel_ms(tfhe_engine *e, float out_ms[2]) {
  if (!e || !out_ms) return fail(TFHE_ERR_INVALID, "null argument");
  out_ms[0] = e->last_ms[0];
  out_ms[1] = e->last_ms[1];
  return TFHE_OK;
}

static int probe_fp64(tfhe_engine *e, bool three_operands, double *tflops_out) {
  if (!e || !tflops_out) return fail(TFHE_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  CU(e->s_misc.reserve(64));
  const int blocks = e->num_sms * 8, iters = 1 << 15;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(e->ev[3], e->stream));
    CU(fp64_probe_launch(static_cast<double *>(e->s_misc.p), blocks, iters, three_operands, e->stream));
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
int tfhe_probe_fp64_tflops(tfhe_engine *e, double *tflops_out) { return probe_fp64(e, false, tflops_out); }
int tfhe_probe_fp64_3op_tflops(tfhe_engine *e, double *tflops_out) { return probe_fp64(e, true, tflops_out); }

int tfhe_engine_load_cloud_key(tfhe_engine *e, uint32_t decomposition_offset,
                               const uint32_t *testvec_a, const uint32_t *testvec_b,
                               const uint32_t *ksk, const double *bsk) {
  if (!e || !testvec_a || !testvec_b || !ksk || !bsk) return fail(TFHE_ERR_INVALID, "null argument");
  Nvtx r("tfhe:load_cloud_key (upload + re-layout + broadcast)");
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
  lut_reset(e);
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return broadcast_key(e, 1);
}

int tfhe_engine_generate_cloud_key(tfhe_engine *e, const uint32_t *s0, const uint32_t *s1,
                                   double alpha_lv0, double alpha_lv1, uint64_t seed) {
  if (!e || !s0 || !s1) return fail(TFHE_ERR_INVALID, "null argument");
  // 256-bit generator key: OS entropy (seed == 0, the secure default -- the reference draws from
  // rand::thread_rng, an OS-seeded ChaCha generator), or expanded from `seed` for reproducible
  // TESTS ONLY (a 64-bit seed caps the key's security at 2^64 at best).
  uint32_t key256[8];
  if (seed == 0) {
    if (getentropy(key256, sizeof(key256)) != 0) return fail(TFHE_ERR_INVALID, "getentropy failed");
  } else {
    uint64_t z = seed;
    for (int i = 0; i < 4; i++) {   // SplitMix64
      z += 0x9E3779B97F4A7C15ull;
      uint64_t x = z;
      x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
      x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
      x ^= x >> 31;
      key256[2 * i] = (uint32_t)x; key256[2 * i + 1] = (uint32_t)(x >> 32);
    }
  }
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
                   d_ksk_ref, p.n, p.l, p.bgbit, p.basebit, p.iks_t, alpha_lv0, alpha_lv1, key256,
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
  lut_reset(e);
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return broadcast_key(e, 1);
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
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->blob) return fail(TFHE_ERR_NO_KEY, "cloud key blob not allocated");
  CU(cudaSetDevice(e->dev));   // several engines may live in one process: the re-layout kernels go to OUR device
  e->decomp_offset = decomposition_offset;
  lut_reset(e);
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
  hd.iks_t = e->p.iks_t; hd.decomposition_offset = e->decomp_offset; hd.n_lut = (uint32_t)lut_count(e);
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
  if (hd.n_lut < 1 || hd.n_lut > (uint32_t)kScratchLut) return fail(TFHE_ERR_INVALID, "bad lut count");
  CU(cudaMemcpyAsync(e->blob, static_cast<const uint8_t *>(host_buf) + sizeof(hd), e->blob_bytes,
                     cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  e->decomp_offset = hd.decomposition_offset;
  lut_reset(e, (int)hd.n_lut);   // slots 1..n_lut-1 of the blob stay addressable under the new epoch
  { int rc_ = finalize_key(e); if (rc_ != TFHE_OK) return rc_; }
  e->key_loaded = true;
  return broadcast_key(e, (int)hd.n_lut);
}

int tfhe_batch_gate(tfhe_engine *e, tfhe_gate op, const uint32_t *in_pairs, uint32_t *out,
                    size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if ((int)op < 0 || (int)op >= TFHE_GATE_COUNT) return fail(TFHE_ERR_INVALID, "bad gate %d", (int)op);
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) {
    return run_host(d, (int)op, nullptr, -1, in_pairs + b * 2 * w, 2 * w, out + b * w, w, n, 0);
  });
}

int tfhe_batch_gate_mixed(tfhe_engine *e, const uint8_t *ops, const uint32_t *in_pairs,
                          uint32_t *out, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (!ops && count) return fail(TFHE_ERR_INVALID, "null ops");
  for (size_t i = 0; i < count; i++)
    if (ops[i] >= TFHE_GATE_COUNT) return fail(TFHE_ERR_INVALID, "bad gate %d at %zu", ops[i], i);
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) {
    return run_host(d, 0, ops + b, -1, in_pairs + b * 2 * w, 2 * w, out + b * w, w, n, 0);
  });
}

int tfhe_batch_bootstrap(tfhe_engine *e, const uint32_t *in, uint32_t *out, size_t count,
                         int key_switch) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) {
    return run_host(d, -1, nullptr, -1, in + b * w, w, out + b * w, w, n, key_switch ? 0 : 1);
  });
}

int tfhe_batch_blind_rotate(tfhe_engine *e, const uint32_t *in, uint32_t *out_trlwe, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) {
    return run_host(d, -1, nullptr, -1, in + b * w, w, out_trlwe + b * 2 * TFHE_N, 2 * TFHE_N, n, 2);
  });
}

// lut/generator.rs:89-137 on the device into test-vector slot `sidx` (engine lock held)
static int generate_into_slot(tfhe_engine *e, int sidx, const uint32_t *f_table, uint32_t modulus,
                              double scale, uint32_t *lut_b_out) {
  if (scale <= 0.0) scale = 1.0 / (2.0 * (double)modulus);  // lut/encoder.rs:36
  CU(e->s_misc.reserve((size_t)modulus * 4));
  CU(cudaMemcpyAsync(e->s_misc.p, f_table, (size_t)modulus * 4, cudaMemcpyHostToDevice, e->stream));
  uint32_t *slot = e->tv() + (size_t)sidx * 2 * TFHE_N;
  CU(lut_generate_launch(static_cast<const uint32_t *>(e->s_misc.p), modulus, scale, slot, e->stream));
  e->launches++;
  if (lut_b_out)
    CU(cudaMemcpyAsync(lut_b_out, slot + TFHE_N, TFHE_N * 4, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return TFHE_OK;
}

// fill test-vector slot `sidx` of ONE device with a caller-made polynomial pair (engine lock held)
static int register_into_slot(tfhe_engine *e, int sidx, const uint32_t *poly_a, const uint32_t *poly_b) {
  CU(cudaSetDevice(e->dev));
  uint32_t *slot = e->tv() + (size_t)sidx * 2 * TFHE_N;
  if (poly_a) CU(cudaMemcpyAsync(slot, poly_a, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  else CU(cudaMemsetAsync(slot, 0, TFHE_N * 4, e->stream));
  CU(cudaMemcpyAsync(slot + TFHE_N, poly_b, TFHE_N * 4, cudaMemcpyHostToDevice, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  return TFHE_OK;
}
// Tables live in the same slot on every device of a multi-GPU engine: the root picks the slot, the
// peers mirror it.  fill(engine) writes the slot on one device.
extern "C++" {
template <class F> static int lut_new(tfhe_engine *e, F fill, int *lut_id_out) {
  std::lock_guard<std::mutex> lock(e->mu);
  if (!e->blob) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
  const int sidx = lut_take_slot(e);
  if (sidx < 0)
    return fail(TFHE_ERR_ALLOC, "out of LUT slots (%d); release tables with tfhe_lut_release", kScratchLut - 1);
  int rc = fill(e, sidx, true);
  for (size_t i = 0; rc == TFHE_OK && i < e->peers.size(); i++) {
    tfhe_engine *p = e->peers[i];
    std::lock_guard<std::mutex> plock(p->mu);
    rc = fill(p, sidx, false);
    if (rc == TFHE_OK) p->lut_used[sidx] = true;
  }
  if (rc != TFHE_OK) {
    e->lut_used[sidx] = false;
    for (tfhe_engine *p : e->peers) p->lut_used[sidx] = false;
    return rc;
  }
  CU(cudaSetDevice(e->dev));
  *lut_id_out = lut_make_id(e, sidx);
  return TFHE_OK;
}
}  // extern "C++"

int tfhe_lut_register(tfhe_engine *e, const uint32_t *poly_a, const uint32_t *poly_b,
                      int *lut_id_out) {
  if (!e || !poly_b || !lut_id_out) return fail(TFHE_ERR_INVALID, "null argument");
  return lut_new(e, [&](tfhe_engine *d, int sidx, bool) { return register_into_slot(d, sidx, poly_a, poly_b); },
                 lut_id_out);
}

int tfhe_lut_release(tfhe_engine *e, int lut_id) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  const int sidx = lut_slot(e, lut_id);
  if (sidx <= 0) return fail(TFHE_ERR_INVALID, "unknown or stale lut id %d", lut_id);
  CU(cudaSetDevice(e->dev));
  CU(cudaStreamSynchronize(e->stream));   // no queued kernel still reads the slot
  e->lut_used[sidx] = false;
  for (tfhe_engine *p : e->peers) {
    std::lock_guard<std::mutex> plock(p->mu);
    p->lut_used[sidx] = false;
  }
  return TFHE_OK;
}

int tfhe_lut_generate(tfhe_engine *e, const uint32_t *f_table, uint32_t modulus, double scale,
                      uint32_t *lut_b_out, int *lut_id_out) {
  if (!e || !f_table || !lut_id_out) return fail(TFHE_ERR_INVALID, "null argument");
  if (modulus == 0 || modulus > TFHE_N) return fail(TFHE_ERR_INVALID, "bad modulus %u", modulus);
  return lut_new(e, [&](tfhe_engine *d, int sidx, bool root) -> int {
    CU(cudaSetDevice(d->dev));
    return generate_into_slot(d, sidx, f_table, modulus, scale, root ? lut_b_out : nullptr);
  }, lut_id_out);
}

int tfhe_batch_bootstrap_func(tfhe_engine *e, const uint32_t *f_table, uint32_t modulus, double scale,
                              const uint32_t *in, uint32_t *out, size_t count) {
  if (!e || !f_table) return fail(TFHE_ERR_INVALID, "null argument");
  if (modulus == 0 || modulus > TFHE_N) return fail(TFHE_ERR_INVALID, "bad modulus %u", modulus);
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) -> int {
    std::lock_guard<std::mutex> lock(d->mu);   // table generation and its use are one critical section
    if (!d->key_loaded) return fail(TFHE_ERR_NO_KEY, "cloud key not loaded");
    CU(cudaSetDevice(d->dev));
    // the previous call's kernels have finished (run_host synchronises), so the scratch slot is free
    const int rc = generate_into_slot(d, kScratchLut, f_table, modulus, scale, nullptr);
    if (rc != TFHE_OK) return rc;
    return run_host_locked(d, -1, nullptr, kScratchLut, in + b * w, w, out + b * w, w, n, 0);
  });
}

int tfhe_batch_bootstrap_lut(tfhe_engine *e, int lut_id, const uint32_t *in, uint32_t *out,
                             size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  int sidx;
  {
    std::lock_guard<std::mutex> lock(e->mu);
    sidx = lut_slot(e, lut_id);
  }
  if (sidx < 0) return fail(TFHE_ERR_INVALID, "unknown or stale lut id %d", lut_id);
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) {
    return run_host(d, -1, nullptr, sidx, in + b * w, w, out + b * w, w, n, 0);
  });
}

int tfhe_batch_bootstrap_lut_multi(tfhe_engine *e, const int32_t *lut_ids, const uint32_t *in,
                                   uint32_t *out, size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (!lut_ids && count) return fail(TFHE_ERR_INVALID, "null lut_ids");
  std::vector<int32_t> slots;
  try { slots.resize(count); } catch (...) { return fail(TFHE_ERR_ALLOC, "out of host memory"); }
  {
    std::lock_guard<std::mutex> lock(e->mu);
    for (size_t i = 0; i < count; i++) {
      const int sidx = lut_slot(e, lut_ids[i]);
      if (sidx < 0) return fail(TFHE_ERR_INVALID, "unknown or stale lut id %d at %zu", lut_ids[i], i);
      slots[i] = sidx;
    }
  }
  const size_t w = e->p.n + 1;
  return shard(e, count, [&](tfhe_engine *d, size_t b, size_t n) {
    return run_host(d, -1, nullptr, 0, in + b * w, w, out + b * w, w, n, 0, slots.data() + b);
  });
}

// ---- FFTProcessor seam (fft/mod.rs:80-107) ------------------------------------------------------
static int run_seam(tfhe_engine *e, int mode, const void *in_a, const uint32_t *in_b, void *out,
                    size_t count) {
  if (!e) return fail(TFHE_ERR_INVALID, "null engine");
  if (count == 0) return TFHE_OK;
  if (!in_a || !out || (mode == MODE_POLYMUL && !in_b)) return fail(TFHE_ERR_INVALID, "null buffer");
  std::lock_guard<std::mutex> lock(e->mu);
  CU(cudaSetDevice(e->dev));
  const size_t in_bytes = (size_t)TFHE_N * (mode == MODE_FFT ? 8 : 4);
  const size_t out_bytes = (size_t)TFHE_N * (mode == MODE_IFFT ? 8 : 4);
  const size_t chunk = 1u << 15;
  float ms_total = 0.f;
  tfhe_engine::Slot &sl = e->slot[0];
  for (size_t base = 0; base < count; base += chunk) {
    const size_t c = count - base < chunk ? count - base : chunk;
    CU(sl.in.reserve(c * in_bytes));
    CU(sl.out.reserve(c * out_bytes));
    CU(cudaMemcpyAsync(sl.in.p, static_cast<const uint8_t *>(in_a) + base * in_bytes, c * in_bytes,
                       cudaMemcpyHostToDevice, e->stream));
    const uint32_t *d_b = nullptr;
    if (mode == MODE_POLYMUL) {
      CU(sl.ext.reserve(c * in_bytes));
      CU(cudaMemcpyAsync(sl.ext.p, in_b + base * TFHE_N, c * in_bytes, cudaMemcpyHostToDevice, e->stream));
      d_b = static_cast<const uint32_t *>(sl.ext.p);
    }
    CU(cudaEventRecord(sl.br_start, e->stream));
    CU(fft_seam_launch(mode, e->tw_s, sl.in.p, d_b, sl.out.p, c, e->num_sms, e->stream));
    e->launches++;
    CU(cudaEventRecord(sl.br_end, e->stream));
    CU(cudaMemcpyAsync(static_cast<uint8_t *>(out) + base * out_bytes, sl.out.p, c * out_bytes,
                       cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, sl.br_start, sl.br_end));
    ms_total += ms;
  }
  e->last_ms[0] = ms_total; e->last_ms[1] = 0.f;
  return TFHE_OK;
}
int tfhe_batch_ifft(tfhe_engine *e, const uint32_t *in, double *out, size_t count) {
  return run_seam(e, MODE_IFFT, in, nullptr, out, count);
}
int tfhe_batch_fft(tfhe_engine *e, const double *in, uint32_t *out, size_t count) {
  return run_seam(e, MODE_FFT, in, nullptr, out, count);
}
int tfhe_batch_poly_mul(tfhe_engine *e, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t count) {
  return run_seam(e, MODE_POLYMUL, a, b, out, count);
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

// The device path re-records slot 0's events per chunk: before the next chunk is queued the finished
// one's kernel times are added to the running totals (tfhe_engine_synchronize adds the last chunk).
static int dev_retire_chunk(tfhe_engine *e) {
  tfhe_engine::Slot &sl = e->slot[0];
  CU(cudaEventSynchronize(sl.ks_end));
  float t0 = 0.f, t1 = 0.f;
  CU(cudaEventElapsedTime(&t0, sl.br_start, sl.br_end));
  CU(cudaEventElapsedTime(&t1, sl.br_end, sl.ks_end));
  e->dev_ms[0] += t0; e->dev_ms[1] += t1;
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
  e->dev_ms[0] = e->dev_ms[1] = 0.f;
  for (size_t base = 0; base < count; base += kChunk) {
    size_t c = count - base < kChunk ? count - base : kChunk;
    if (base) { int rc = dev_retire_chunk(e); if (rc != TFHE_OK) return rc; }
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
  if (lut_id >= 0) {
    lut_id = lut_slot(e, lut_id);
    if (lut_id < 0) return fail(TFHE_ERR_INVALID, "unknown or stale lut id");
  }
  const size_t w = e->p.n + 1;
  e->dev_ms[0] = e->dev_ms[1] = 0.f;
  for (size_t base = 0; base < count; base += kChunk) {
    size_t c = count - base < kChunk ? count - base : kChunk;
    if (base) { int rc = dev_retire_chunk(e); if (rc != TFHE_OK) return rc; }
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
    e->last_ms[0] = e->dev_ms[0] + t0; e->last_ms[1] = e->dev_ms[1] + t1;
    e->dev_ms[0] = e->dev_ms[1] = 0.f;
  } else {
    cudaGetLastError();
  }
  return TFHE_OK;
}

}  // extern "C"

#include "circuit.cuh"
