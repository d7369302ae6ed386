"""BASELINE.json configs C1-C5 measured on one GPU with real (oracle-generated) keys, each with a
decryption check.  These are the parity-test workloads, not bench.py's headline line; results
go to gpurun_out/configs.json (copied to profiles/ per round)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import rs_tfhe_b200 as T

res = {}


def engine(name, seed=0x5EED0001):
    K = O.Keys(name, seed=seed)
    P = T.PARAMS_BY_NAME[name]
    e = T.CudaBootstrap(P, 0)
    e.load_cloud_key(T.CloudKey(P, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk))
    return K, P, e


def time_gate_batch(K, e, count, reps=3):
    r = np.random.default_rng(count)
    a = r.integers(0, 2, count).astype(bool)
    b = r.integers(0, 2, count).astype(bool)
    pairs = np.stack([K.encrypt_bool_batch(a, 11), K.encrypt_bool_batch(b, 12)], axis=1)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        out = e.batch_gate("NAND", pairs)
        wall = time.perf_counter() - t
        br, ks = e.last_kernel_ms()
        if best is None or wall < best[0]:
            best = (wall, br, ks)
    ok = bool((K.decrypt_bool_batch(out) == ~(a & b)).all())
    wall, br, ks = best
    return {"count": count, "e2e_ms": wall * 1e3, "blind_rotate_ms": br, "key_switch_ms": ks,
            "gates_per_s_e2e": count / wall, "gates_per_s_kernels": count / ((br + ks) * 1e-3),
            "us_per_pbs_kernels": (br + ks) * 1e3 / count, "decrypt_ok": ok}


# C1: batch of 1024 hom_nand gates at SECURITY_128_BIT
K, P, e = engine("128")
res["C1_batch_nand_1024_128bit"] = time_gate_batch(K, e, 1024)
res["C1b_batch_nand_65536_128bit"] = time_gate_batch(K, e, 65536, reps=2)

# C5 (one-GPU slice): 131072 mixed gates
r = np.random.default_rng(5)
cnt = 131072
a = r.integers(0, 2, cnt).astype(bool); b = r.integers(0, 2, cnt).astype(bool)
ops = r.integers(0, 6, cnt).astype(np.uint8)
pairs = np.stack([K.encrypt_bool_batch(a, 21), K.encrypt_bool_batch(b, 22)], axis=1)
t = time.perf_counter(); out = e.batch_gate_mixed(ops, pairs); wall = time.perf_counter() - t
fn = [lambda x, y: ~(x & y), lambda x, y: x & y, lambda x, y: x | y, lambda x, y: x ^ y, lambda x, y: x ^ y,
      lambda x, y: ~(x | y)]
want = np.zeros(cnt, dtype=bool)
for o in range(6):
    m = ops == o
    want[m] = fn[o](a[m], b[m])
res["C5_mixed_gates_131072_per_gpu_128bit"] = {"count": cnt, "e2e_ms": wall * 1e3, "gates_per_s_e2e": cnt / wall,
                                               "decrypt_ok": bool((K.decrypt_bool_batch(out) == want).all())}

# C4: lut_add_two_numbers (examples/lut_add_two_numbers.rs:80-157): 128-bit params, modulus 32.
# The reference algorithm is noise-marginal here: judged on equality with the
# oracle and on latency, not on 42+137=179 always decoding.
m = 32
luts = {"low": [x % 16 for x in range(m)], "carry": [int(x >= 16) for x in range(m)], "high": [x % 16 for x in range(m)]}
ids = {k: e.lut_generate(v, m) for k, v in luts.items()}
rng = O.Rng(99)
A_, B_ = 42, 137
enc = lambda v: K.encrypt_message([v], m, rng)[0]
a_lo, a_hi, b_lo, b_hi = enc(A_ & 15), enc(A_ >> 4), enc(B_ & 15), enc(B_ >> 4)
lat = []
for _ in range(5):
    t = time.perf_counter()
    ct_low = (a_lo + b_lo).astype(np.uint32)
    # PBS1 and PBS2 share the input: one batch with per-ciphertext tables (level 1 of the chain)
    both = e.batch_bootstrap_lut([ids["low"][0], ids["carry"][0]], np.stack([ct_low, ct_low]))
    s_lo, carry = both[0], both[1]
    ct_hi = (a_hi + b_hi + carry).astype(np.uint32)
    s_hi = e.batch_bootstrap_lut(ids["high"][0], ct_hi)
    lat.append(time.perf_counter() - t)
ref_lo = K.batch_bootstrap(ct_low[None], lut_b=ids["low"][1])[0]
ref_c = K.batch_bootstrap(ct_low[None], lut_b=ids["carry"][1])[0]
ref_hi = K.batch_bootstrap(((a_hi + b_hi + ref_c).astype(np.uint32))[None], lut_b=ids["high"][1])[0]
t = time.perf_counter(); one = e.batch_bootstrap_lut(ids["low"][0], ct_low); one_pbs = time.perf_counter() - t
res["C4_lut_add_two_numbers_128bit_m32"] = {
    "chain_latency_ms_best": min(lat) * 1e3, "single_pbs_latency_ms": one_pbs * 1e3,
    "equal_to_oracle_words": bool((s_lo == ref_lo).all() and (carry == ref_c).all() and (s_hi == ref_hi).all()),
    "decoded": int(K.decrypt_message(s_lo[None], m)[0] + 16 * K.decrypt_message(s_hi[None], m)[0]),
    "expected": (A_ + B_) & 255}
e.close()

# C2: 80-bit and 110-bit sweeps
for name in ("80", "110"):
    K, P, e = engine(name)
    res[f"C2_sweep_{name}bit"] = [time_gate_batch(K, e, c, reps=2) for c in (1024, 2048, 4096, 8192, 16384, 32768, 65536)]
    e.close()

# C3: LutBootstrap::bootstrap_func, SECURITY_UINT4, messageModulus 16, batch 16384
K, P, e = engine("uint4", seed=0x5EED0003)
m = 16
cnt = 16384
msgs = np.random.default_rng(3).integers(0, m, cnt)
cts = K.encrypt_message_batch(msgs, m, 31)
out3 = {}
for fname, f in (("identity", lambda x: x), ("square", lambda x: (x * x) % 16)):
    lut_id, _ = e.lut_generate([f(x) for x in range(m)], m)
    best = None
    for _ in range(3):
        t = time.perf_counter(); out = e.batch_bootstrap_lut(lut_id, cts); wall = time.perf_counter() - t
        br, ks = e.last_kernel_ms()
        if best is None or wall < best[0]:
            best = (wall, br, ks)
    want = np.array([f(int(x)) for x in msgs])
    out3[fname] = {"count": cnt, "e2e_ms": best[0] * 1e3, "blind_rotate_ms": best[1], "key_switch_ms": best[2],
                   "pbs_per_s_e2e": cnt / best[0], "us_per_pbs_kernels": (best[1] + best[2]) * 1e3 / cnt,
                   "decrypt_ok": bool((K.decrypt_message_batch(out, m) == want).all())}
res["C3_lut_uint4_m16_batch16384"] = out3
e.close()

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)
print(json.dumps(res, indent=1))
