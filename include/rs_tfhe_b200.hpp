// rs_tfhe_b200.hpp -- header-only C++17 host mirror of rs-tfhe's interface for the
// bootstrapped-gate path, over the C ABI in tfhe_b200.h.
//
// The reference is compiled code (Rust) and its toolchain is absent from this image, so
// this is the compiled-language host side: the same names, argument meaning and error
// behaviour (the reference panics; this throws std::runtime_error) as
//   key::CloudKey              src/key.rs:51-56
//   bootstrap::Bootstrap       src/bootstrap/mod.rs:23-38     (CudaBootstrap implements it)
//   gates::Gates + free fns    src/gates.rs:30-326
//   gates::batch_*             src/gates.rs:352-547
//   trgsw::batch_blind_rotate  src/trgsw.rs:289-305
//   lut::Generator, LookupTable, bootstrap::lut::LutBootstrap
//                              src/lut/generator.rs:16-137, src/bootstrap/lut.rs:28-126
// Types are the reference's memory images, so a Rust caller and this header agree byte for
// byte.  There is no CPU fallback: without a CUDA device construction throws.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tfhe_b200.h"

namespace rs_tfhe {

using Torus = uint32_t;  // params.rs:40

// params::SecurityParams (params.rs:53-84), runtime image
struct SecurityParams {
  const char *name;
  uint32_t n, N, l, bgbit, basebit, iks_t;
  double alpha_lv0, alpha_lv1;
};
inline constexpr SecurityParams SECURITY_80_BIT{"80", 550, 1024, 3, 6, 2, 7, 5.0e-5, 3.73e-8};
inline constexpr SecurityParams SECURITY_110_BIT{"110", 630, 1024, 3, 6, 2, 8, 3.0517578125e-05, 2.9802322387695313e-8};
inline constexpr SecurityParams SECURITY_128_BIT{"128", 700, 1024, 3, 6, 2, 9, 2.0e-5, 2.0e-8};
inline constexpr SecurityParams SECURITY_UINT4{"uint4", 820, 1024, 1, 22, 5, 3, 0.0000025167616095979554, 0.0000000000000002220446049250313};

inline Torus f64_to_torus(double d) {  // utils.rs:9-12
  return static_cast<Torus>(static_cast<int64_t>(std::fmod(d, 1.0) * 4294967296.0));
}

// tlwe::TLWELv0 (tlwe.rs:12-14): p[0..n) = a, p[n] = b.  Runtime-sized.
struct Ciphertext {
  std::vector<Torus> p;
  Ciphertext() = default;
  explicit Ciphertext(uint32_t n) : p(n + 1, 0) {}
  Torus b() const { return p.back(); }
  Torus &b_mut() { return p.back(); }
};
// trlwe::TRLWELv1 (trlwe.rs:11-14)
struct TRLWELv1 { Torus a[TFHE_N]; Torus b[TFHE_N]; };

// key::CloudKey (key.rs:51-56) in the reference's layout; the vectors are borrowed by value here
struct CloudKey {
  SecurityParams params;
  Torus decomposition_offset;
  TRLWELv1 blind_rotate_testvec;
  std::vector<Torus> key_switching_key;   // [N*t*2^basebit][n+1]
  std::vector<double> bootstrapping_key;  // [n][2l][2][N]  (TRGSWLv1FFT image)
};

inline void check(int rc) {
  if (rc != TFHE_OK) throw std::runtime_error(std::string("tfhe_b200: ") + tfhe_last_error());
}

// Identity of a CloudKey's CONTENT (not its address: a new key at a reused address, or a mutated
// key, must be re-uploaded).  FNV-1a over the offset, the sizes and 1024 strided samples of each of
// the three arrays -- microseconds per call, and any regenerated key differs in every sample.
inline uint64_t cloud_key_fingerprint(const CloudKey &ck) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&h](uint64_t v) { for (int i = 0; i < 8; i++) { h ^= (v >> (8 * i)) & 0xff; h *= 1099511628211ull; } };
  mix(ck.decomposition_offset); mix(ck.key_switching_key.size()); mix(ck.bootstrapping_key.size());
  for (int i = 0; i < TFHE_N; i += 8) { mix(ck.blind_rotate_testvec.a[i]); mix(ck.blind_rotate_testvec.b[i]); }
  const size_t ks = ck.key_switching_key.size(), bs = ck.bootstrapping_key.size();
  for (size_t i = 0; i < 1024 && ks; i++) mix(ck.key_switching_key[(i * 2654435761ull) % ks]);
  for (size_t i = 0; i < 1024 && bs; i++) {
    uint64_t bits;
    std::memcpy(&bits, &ck.bootstrapping_key[(i * 2654435761ull) % bs], 8);
    mix(bits);
  }
  return h;
}

enum class Gate : int { Nand = 0, And, Or, Xor, Xnor, Nor, AndNy, AndYn, OrNy, OrYn };

// bootstrap::Bootstrap (bootstrap/mod.rs:23-38)
class Bootstrap {
 public:
  virtual ~Bootstrap() = default;
  virtual Ciphertext bootstrap(const Ciphertext &ctxt, const CloudKey &ck) = 0;
  virtual Ciphertext bootstrap_without_key_switch(const Ciphertext &ctxt, const CloudKey &ck) = 0;
  virtual const char *name() const = 0;
};

// lut::LookupTable (lut/lookup_table.rs:16-19): poly.a == 0, poly.b = table; lives on the device.
// Move-only owner of its device slot: dropping it releases the slot, as dropping the reference's
// LookupTable frees its polynomial.
struct LookupTable {
  std::vector<Torus> poly_b;
  int lut_id = -1;
  tfhe_engine *engine = nullptr;
  LookupTable() = default;
  LookupTable(const LookupTable &) = delete;
  LookupTable &operator=(const LookupTable &) = delete;
  LookupTable(LookupTable &&o) noexcept : poly_b(std::move(o.poly_b)), lut_id(o.lut_id), engine(o.engine) { o.lut_id = -1; }
  LookupTable &operator=(LookupTable &&o) noexcept {
    if (this != &o) { release(); poly_b = std::move(o.poly_b); lut_id = o.lut_id; engine = o.engine; o.lut_id = -1; }
    return *this;
  }
  ~LookupTable() { release(); }
  void release() {
    if (lut_id > 0 && engine) tfhe_lut_release(engine, lut_id);   // a stale id (key reloaded) is refused, not fatal
    lut_id = -1;
  }
  bool is_empty() const {
    for (Torus v : poly_b) if (v) return false;
    return true;
  }
};

// The B200 strategy: one engine per (process, GPU); the cloud key is resident on the device
// and re-uploaded only when a different CloudKey object is passed.
class CudaBootstrap final : public Bootstrap {
 public:
  explicit CudaBootstrap(const SecurityParams &p = SECURITY_128_BIT, int device = 0) : params_(p) {
    tfhe_params cp{p.n, p.N, p.l, p.bgbit, p.basebit, p.iks_t};
    check(tfhe_engine_create(&cp, device, &e_));
  }
  ~CudaBootstrap() override { tfhe_engine_destroy(e_); }
  CudaBootstrap(const CudaBootstrap &) = delete;
  CudaBootstrap &operator=(const CudaBootstrap &) = delete;

  const char *name() const override { return "cuda-b200"; }
  const SecurityParams &params() const { return params_; }
  tfhe_engine *raw() { return e_; }

  void bind(const CloudKey &ck) {
    const uint64_t fp = cloud_key_fingerprint(ck);
    if (have_key_ && fp == bound_fp_) return;
    check(tfhe_engine_load_cloud_key(e_, ck.decomposition_offset, ck.blind_rotate_testvec.a,
                                     ck.blind_rotate_testvec.b, ck.key_switching_key.data(),
                                     ck.bootstrapping_key.data()));
    bound_fp_ = fp;
    have_key_ = true;
  }

  Ciphertext bootstrap(const Ciphertext &ctxt, const CloudKey &ck) override {  // vanilla.rs:40-52
    bind(ck);
    Ciphertext out(params_.n);
    check(tfhe_batch_bootstrap(e_, ctxt.p.data(), out.p.data(), 1, 1));
    return out;
  }
  Ciphertext bootstrap_without_key_switch(const Ciphertext &ctxt, const CloudKey &ck) override {
    bind(ck);  // vanilla.rs:54-63
    Ciphertext out(params_.n);
    check(tfhe_batch_bootstrap(e_, ctxt.p.data(), out.p.data(), 1, 0));
    return out;
  }

  // gates::batch_<op> (gates.rs:352-547)
  std::vector<Ciphertext> batch_gate(Gate op, const std::vector<std::pair<Ciphertext, Ciphertext>> &inputs,
                                     const CloudKey &ck) {
    bind(ck);
    const size_t w = params_.n + 1, count = inputs.size();
    std::vector<Torus> in(count * 2 * w), out(count * w);
    for (size_t i = 0; i < count; i++) {
      std::copy(inputs[i].first.p.begin(), inputs[i].first.p.end(), in.begin() + i * 2 * w);
      std::copy(inputs[i].second.p.begin(), inputs[i].second.p.end(), in.begin() + i * 2 * w + w);
    }
    check(tfhe_batch_gate(e_, static_cast<tfhe_gate>(op), in.data(), out.data(), count));
    return unpack(out, count);
  }
  // trgsw::batch_blind_rotate (trgsw.rs:289-305)
  std::vector<TRLWELv1> batch_blind_rotate(const std::vector<Ciphertext> &srcs, const CloudKey &ck) {
    bind(ck);
    const size_t w = params_.n + 1;
    std::vector<Torus> in(srcs.size() * w);
    for (size_t i = 0; i < srcs.size(); i++) std::copy(srcs[i].p.begin(), srcs[i].p.end(), in.begin() + i * w);
    std::vector<TRLWELv1> out(srcs.size());
    check(tfhe_batch_blind_rotate(e_, in.data(), reinterpret_cast<Torus *>(out.data()), srcs.size()));
    return out;
  }
  // lut::Generator::generate_lookup_table (lut/generator.rs:66-137): closure tabulated on the host
  LookupTable generate_lookup_table(const std::function<size_t(size_t)> &f, uint32_t message_modulus,
                                    const CloudKey &ck, double scale = 0.0) {
    bind(ck);
    std::vector<Torus> table(message_modulus);
    for (uint32_t x = 0; x < message_modulus; x++) table[x] = static_cast<Torus>(f(x) % message_modulus);
    LookupTable lut;
    lut.poly_b.resize(TFHE_N);
    check(tfhe_lut_generate(e_, table.data(), message_modulus, scale, lut.poly_b.data(), &lut.lut_id));
    lut.engine = e_;
    return lut;
  }
  // LutBootstrap::bootstrap_func over a batch (bootstrap/lut.rs:49-65): table generated into the
  // engine's scratch slot, nothing to release, callable without bound
  std::vector<Ciphertext> batch_bootstrap_func(const std::vector<Ciphertext> &cts,
                                               const std::function<size_t(size_t)> &f,
                                               uint32_t message_modulus, const CloudKey &ck) {
    bind(ck);
    std::vector<Torus> table(message_modulus);
    for (uint32_t x = 0; x < message_modulus; x++) table[x] = static_cast<Torus>(f(x) % message_modulus);
    const size_t w = params_.n + 1;
    std::vector<Torus> in(cts.size() * w), out(cts.size() * w);
    for (size_t i = 0; i < cts.size(); i++) std::copy(cts[i].p.begin(), cts[i].p.end(), in.begin() + i * w);
    check(tfhe_batch_bootstrap_func(e_, table.data(), message_modulus, 0.0, in.data(), out.data(), cts.size()));
    return unpack(out, cts.size());
  }
  // LutBootstrap::bootstrap_lut over a batch (bootstrap/lut.rs:79-99)
  std::vector<Ciphertext> batch_bootstrap_lut(const std::vector<Ciphertext> &cts, const LookupTable &lut,
                                              const CloudKey &ck) {
    bind(ck);
    const size_t w = params_.n + 1;
    std::vector<Torus> in(cts.size() * w), out(cts.size() * w);
    for (size_t i = 0; i < cts.size(); i++) std::copy(cts[i].p.begin(), cts[i].p.end(), in.begin() + i * w);
    check(tfhe_batch_bootstrap_lut(e_, lut.lut_id, in.data(), out.data(), cts.size()));
    return unpack(out, cts.size());
  }

 private:
  std::vector<Ciphertext> unpack(const std::vector<Torus> &flat, size_t count) const {
    const size_t w = params_.n + 1;
    std::vector<Ciphertext> res(count, Ciphertext(params_.n));
    for (size_t i = 0; i < count; i++) std::copy(flat.begin() + i * w, flat.begin() + (i + 1) * w, res[i].p.begin());
    return res;
  }
  SecurityParams params_;
  tfhe_engine *e_ = nullptr;
  uint64_t bound_fp_ = 0;
  bool have_key_ = false;
};

// gates::Gates (gates.rs:30-218)
class Gates {
 public:
  explicit Gates(std::shared_ptr<CudaBootstrap> b) : bootstrap_(std::move(b)) {}
  static Gates with_bootstrap(std::shared_ptr<CudaBootstrap> b) { return Gates(std::move(b)); }
  const char *bootstrap_strategy() const { return bootstrap_->name(); }

  Ciphertext nand(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::Nand, a, b, ck); }
  Ciphertext or_(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::Or, a, b, ck); }
  Ciphertext and_(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::And, a, b, ck); }
  Ciphertext xor_(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::Xor, a, b, ck); }
  Ciphertext xnor(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::Xnor, a, b, ck); }
  Ciphertext nor(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::Nor, a, b, ck); }
  Ciphertext and_ny(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::AndNy, a, b, ck); }
  Ciphertext and_yn(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::AndYn, a, b, ck); }
  Ciphertext or_ny(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::OrNy, a, b, ck); }
  Ciphertext or_yn(const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) { return gate(Gate::OrYn, a, b, ck); }

  // gates.rs:189-199
  Ciphertext mux_naive(const Ciphertext &a, const Ciphertext &b, const Ciphertext &c, const CloudKey &ck) {
    auto u = bootstrap_->batch_gate(Gate::And, {{a, b}, {not_(a), c}}, ck);
    return or_(u[0], u[1], ck);
  }
  Ciphertext not_(const Ciphertext &a) const {  // gates.rs:202-204
    Ciphertext r = a;
    for (auto &v : r.p) v = 0u - v;
    return r;
  }
  Ciphertext copy(const Ciphertext &a) const { return a; }  // gates.rs:207-209
  Ciphertext constant(bool value) const {                   // gates.rs:212-218 (release-mode wrap)
    Torus mu = f64_to_torus(0.125);
    Ciphertext r(bootstrap_->params().n);
    r.b_mut() = value ? mu : 1u - mu;
    return r;
  }

 private:
  Ciphertext gate(Gate op, const Ciphertext &a, const Ciphertext &b, const CloudKey &ck) {
    return bootstrap_->batch_gate(op, {{a, b}}, ck)[0];
  }
  std::shared_ptr<CudaBootstrap> bootstrap_;
};

// gates::batch_* free functions (gates.rs:352-547) on a caller-provided strategy
using GatePairs = std::vector<std::pair<Ciphertext, Ciphertext>>;
inline std::vector<Ciphertext> batch_nand(CudaBootstrap &e, const GatePairs &in, const CloudKey &ck) { return e.batch_gate(Gate::Nand, in, ck); }
inline std::vector<Ciphertext> batch_and(CudaBootstrap &e, const GatePairs &in, const CloudKey &ck) { return e.batch_gate(Gate::And, in, ck); }
inline std::vector<Ciphertext> batch_or(CudaBootstrap &e, const GatePairs &in, const CloudKey &ck) { return e.batch_gate(Gate::Or, in, ck); }
inline std::vector<Ciphertext> batch_xor(CudaBootstrap &e, const GatePairs &in, const CloudKey &ck) { return e.batch_gate(Gate::Xor, in, ck); }
inline std::vector<Ciphertext> batch_nor(CudaBootstrap &e, const GatePairs &in, const CloudKey &ck) { return e.batch_gate(Gate::Nor, in, ck); }
inline std::vector<Ciphertext> batch_xnor(CudaBootstrap &e, const GatePairs &in, const CloudKey &ck) { return e.batch_gate(Gate::Xnor, in, ck); }

// bootstrap::lut::LutBootstrap (bootstrap/lut.rs:28-126)
class LutBootstrap {
 public:
  explicit LutBootstrap(std::shared_ptr<CudaBootstrap> b) : e_(std::move(b)) {}
  const char *name() const { return "lut"; }
  Ciphertext bootstrap_func(const Ciphertext &ct, const std::function<size_t(size_t)> &f,
                            uint32_t message_modulus, const CloudKey &ck) {  // lut.rs:49-65
    return e_->batch_bootstrap_func({ct}, f, message_modulus, ck)[0];
  }
  Ciphertext bootstrap_lut(const Ciphertext &ct, const LookupTable &lut, const CloudKey &ck) {  // lut.rs:79-99
    return e_->batch_bootstrap_lut({ct}, lut, ck)[0];
  }

 private:
  std::shared_ptr<CudaBootstrap> e_;
};

}  // namespace rs_tfhe
