// C++ parity test of the compiled host mirror (include/rs_tfhe_b200.hpp); reads like the
// reference's own tests (src/gates.rs:558-681, src/bootstrap/lut.rs:141-254).  Keys,
// encryption and decryption come from the CPU oracle (test infrastructure); evaluation goes
// through the C ABI on the GPU.  Exit code 0 = all checks passed; 3 = no CUDA device.
#include <cstdio>
#include <cstring>
#include <memory>

#include "../../include/rs_tfhe_b200.hpp"
#include "../../oracle/tfhe_oracle.h"

using namespace rs_tfhe;

struct TestKeys {
  orc_params p;
  std::vector<uint32_t> s0, s1;
  CloudKey ck;
  orc_rng rng;
};

static std::unique_ptr<TestKeys> make_keys(const SecurityParams &sp, uint64_t seed = 0x5EED0001) {
  auto k = std::make_unique<TestKeys>();
  orc_params_by_name(sp.name, &k->p);
  k->s0.resize(sp.n); k->s1.resize(1024);
  orc_secret_key(&k->p, seed, k->s0.data(), k->s1.data());
  k->ck.params = sp;
  k->ck.decomposition_offset = orc_decomposition_offset(&k->p);
  orc_gen_testvec(k->ck.blind_rotate_testvec.a, k->ck.blind_rotate_testvec.b);
  k->ck.key_switching_key.resize(orc_ksk_words(&k->p));
  orc_gen_ksk(&k->p, k->s0.data(), k->s1.data(), seed + 1, k->ck.key_switching_key.data());
  k->ck.bootstrapping_key.resize(orc_bsk_doubles(&k->p));
  orc_gen_bsk(&k->p, k->s0.data(), k->s1.data(), seed + 2, k->ck.bootstrapping_key.data(), nullptr);
  orc_rng_seed(&k->rng, 42);
  return k;
}
static Ciphertext enc(TestKeys &k, bool bit) {
  Ciphertext c(k.p.n);
  orc_lwe_encrypt_bool(&k.p, bit, k.s0.data(), &k.rng, c.p.data());
  return c;
}
static bool dec(TestKeys &k, const Ciphertext &c) { return orc_lwe_decrypt_bool(c.p.data(), k.s0.data(), k.p.n); }

static int failures = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { failures++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

int main() {
  if (tfhe_device_count() == 0) {
    try { CudaBootstrap e; } catch (const std::runtime_error &ex) { printf("no device, fails loudly: %s\n", ex.what()); return 3; }
    printf("expected construction to throw without a device\n");
    return 1;
  }
  auto k = make_keys(SECURITY_128_BIT);
  auto engine = std::make_shared<CudaBootstrap>(SECURITY_128_BIT, 0);
  Gates gates = Gates::with_bootstrap(engine);
  EXPECT(std::strcmp(gates.bootstrap_strategy(), "cuda-b200") == 0, "strategy name");

  // test_gate over the truth table (gates.rs:558-653); xnor is pinned to b ^ a (gates.rs:575-579)
  struct { const char *name; bool (*f)(bool, bool); Ciphertext (Gates::*g)(const Ciphertext &, const Ciphertext &, const CloudKey &); } cases[] = {
      {"nand", [](bool a, bool b) { return !(a && b); }, &Gates::nand},
      {"or", [](bool a, bool b) { return a || b; }, &Gates::or_},
      {"and", [](bool a, bool b) { return a && b; }, &Gates::and_},
      {"xor", [](bool a, bool b) { return a != b; }, &Gates::xor_},
      {"xnor", [](bool a, bool b) { return a != b; }, &Gates::xnor},
      {"nor", [](bool a, bool b) { return !(a || b); }, &Gates::nor},
      {"and_ny", [](bool a, bool b) { return !a && b; }, &Gates::and_ny},
      {"and_yn", [](bool a, bool b) { return a && !b; }, &Gates::and_yn},
      {"or_ny", [](bool a, bool b) { return !a || b; }, &Gates::or_ny},
      {"or_yn", [](bool a, bool b) { return a || !b; }, &Gates::or_yn},
  };
  for (auto &c : cases)
    for (int ab = 0; ab < 4; ab++) {
      bool a = ab & 1, b = ab & 2;
      Ciphertext r = (gates.*c.g)(enc(*k, a), enc(*k, b), k->ck);
      EXPECT(dec(*k, r) == c.f(a, b), "%s(%d,%d)", c.name, a, b);
    }
  // not / copy / constant (gates.rs:202-218)
  EXPECT(dec(*k, gates.not_(enc(*k, true))) == false, "not");
  EXPECT(dec(*k, gates.copy(enc(*k, true))) == true, "copy");
  EXPECT(dec(*k, gates.constant(true)) && !dec(*k, gates.constant(false)), "constant");
  // mux_naive (gates.rs:656-681)
  for (int abc = 0; abc < 8; abc++) {
    bool a = abc & 1, b = abc & 2, c = abc & 4;
    EXPECT(dec(*k, gates.mux_naive(enc(*k, a), enc(*k, b), enc(*k, c), k->ck)) == (a ? b : c), "mux_naive %d", abc);
  }
  // batch == sequential word for word, and equals the oracle (gates.rs:752-762)
  GatePairs in;
  for (int i = 0; i < 5; i++) in.push_back({enc(*k, i % 2 == 0), enc(*k, i % 3 == 0)});
  auto batch = batch_nand(*engine, in, k->ck);
  for (int i = 0; i < 5; i++) {
    Ciphertext seq = gates.nand(in[i].first, in[i].second, k->ck);
    EXPECT(seq.p == batch[i].p, "batch != sequential at %d", i);
    EXPECT(dec(*k, batch[i]) == !((i % 2 == 0) && (i % 3 == 0)), "batch nand %d", i);
    std::vector<uint32_t> prep(k->p.n + 1), ref(k->p.n + 1);
    orc_gate_prep(&k->p, 0, in[i].first.p.data(), in[i].second.p.data(), prep.data());
    orc_bootstrap(&k->p, k->ck.decomposition_offset, k->ck.bootstrapping_key.data(), k->ck.key_switching_key.data(),
                  k->ck.blind_rotate_testvec.a, k->ck.blind_rotate_testvec.b, prep.data(), 1, ref.data());
    EXPECT(ref == batch[i].p, "GPU != oracle at %d", i);
  }
  // Bootstrap trait (vanilla.rs:78-142)
  Bootstrap &strategy = *engine;
  EXPECT(dec(*k, strategy.bootstrap(enc(*k, true), k->ck)) == true, "bootstrap(true)");
  EXPECT(dec(*k, strategy.bootstrap(enc(*k, false), k->ck)) == false, "bootstrap(false)");
  (void)strategy.bootstrap_without_key_switch(enc(*k, true), k->ck);  // "does not panic" (vanilla.rs:124-141)
  // LUT bootstrap at modulus 2 (bootstrap/lut.rs:141-254)
  LutBootstrap lb(engine);
  for (int msg = 0; msg < 2; msg++) {
    Ciphertext c(k->p.n);
    orc_lwe_encrypt_message(&k->p, msg, 2, k->s0.data(), &k->rng, c.p.data());
    Ciphertext id = lb.bootstrap_func(c, [](size_t x) { return x; }, 2, k->ck);
    Ciphertext nt = lb.bootstrap_func(c, [](size_t x) { return 1 - x; }, 2, k->ck);
    EXPECT(orc_lwe_decrypt_message(id.p.data(), k->s0.data(), k->p.n, 2) == (uint32_t)msg, "lut identity %d", msg);
    EXPECT(orc_lwe_decrypt_message(nt.p.data(), k->s0.data(), k->p.n, 2) == (uint32_t)(1 - msg), "lut not %d", msg);
  }
  // bootstrap_func can be called without bound (the reference builds and drops a table per call,
  // lut.rs:49-65); explicit tables give their device slot back when dropped
  {
    Ciphertext c(k->p.n);
    orc_lwe_encrypt_message(&k->p, 1, 2, k->s0.data(), &k->rng, c.p.data());
    int wrong = 0;
    for (int it = 0; it < 200; it++) {
      Ciphertext r = lb.bootstrap_func(c, [it](size_t x) { return (x + it) & 1; }, 2, k->ck);
      wrong += orc_lwe_decrypt_message(r.p.data(), k->s0.data(), k->p.n, 2) != (uint32_t)((1 + it) & 1);
    }
    EXPECT(wrong == 0, "%d of 200 bootstrap_func calls wrong", wrong);
    for (int it = 0; it < 200; it++) {
      LookupTable lut = engine->generate_lookup_table([](size_t x) { return x; }, 2, k->ck);
      EXPECT(lut.lut_id > 0, "table %d got no slot", it);
    }
  }
  // a DIFFERENT key at the SAME address must be re-uploaded (identity is by content, not pointer)
  {
    auto k2 = make_keys(SECURITY_128_BIT, 0x5EED0A01);
    std::swap(k->ck, k2->ck); std::swap(k->s0, k2->s0); std::swap(k->s1, k2->s1);   // k->ck keeps its address, new content
    EXPECT(dec(*k, strategy.bootstrap(enc(*k, true), k->ck)) == true, "rebind after key content change (true)");
    EXPECT(dec(*k, strategy.bootstrap(enc(*k, false), k->ck)) == false, "rebind after key content change (false)");
  }
  printf(failures ? "FAILED (%d)\n" : "ALL OK\n", failures);
  return failures ? 1 : 0;
}
