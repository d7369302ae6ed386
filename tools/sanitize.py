"""Small workload for compute-sanitizer through the C ABI.  It reaches every shipped kernel shape: the 2-CTA
cluster latency kernel (5 gates; TFHE_BR_CLUSTER=0 in the environment: the one-SM latency kernel instead), the
128-thread throughput kernel (a partial round of 160 gates), the 64-thread one (SANITIZE_COUNT > 592),
LUT bootstrap, extract + key switch, the FFT seam and a levelised circuit with a fused MUX."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import rs_tfhe_b200 as T
from rs_tfhe_b200 import circuit as CI
K = O.Keys("128", seed=1)
e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
e.load_cloud_key(T.CloudKey(T.SECURITY_128_BIT, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk))
rng = O.Rng(2)
for count in [5, int(os.environ.get("SANITIZE_TAIL", "160"))] + ([int(os.environ["SANITIZE_COUNT"])] if "SANITIZE_COUNT" in os.environ else []):
    a = np.resize(np.array([0, 1, 1, 0, 1], dtype=bool), count); b = np.resize(np.array([1, 1, 0, 0, 1], dtype=bool), count)
    pairs = np.stack([K.encrypt_bool(a, rng), K.encrypt_bool(b, rng)], axis=1)
    ops = np.resize(np.array([0, 1, 2, 3, 5], dtype=np.uint8), count)
    out = e.batch_gate_mixed(ops, pairs)
    print(f"gates equal ({count}):", np.array_equal(out, K.batch_gate(ops, pairs)))
lut_id, lut_b = e.lut_generate([1, 0], 2)
ct = K.encrypt_message([1, 0], 2, rng)
print("lut equal:", np.array_equal(e.batch_bootstrap_lut(lut_id, ct), K.batch_bootstrap(ct, lut_b=lut_b)))
tr = e.batch_blind_rotate(ct)
print("ks equal:", np.array_equal(e.batch_extract_key_switch(tr), K.batch_bootstrap(ct)))
# FFT seam: round trip and a product against the schoolbook
r = np.random.default_rng(3)
polys = r.integers(0, 2**32, (3, 1024), dtype=np.uint32)
print("fft round trip:", np.array_equal(e.batch_fft(e.batch_ifft(polys)), polys))
small = r.integers(0, 64, (3, 1024), dtype=np.uint32)
prod = e.batch_poly_mul(polys, small)
x, y = polys[0].astype(np.int64).astype(object), small[0].astype(object)
ref0 = [(sum(int(x[j]) * int(y[i - j]) for j in range(i + 1)) - sum(int(x[j]) * int(y[1024 + i - j]) for j in range(i + 1, 1024))) % 2**32 for i in (0, 1, 1023)]
print("poly_mul equal:", [int(prod[0][i]) for i in (0, 1, 1023)] == ref0)
# circuit: one full adder and a fused MUX over 3 input sets
c = CI.Circuit()
wa, wb, wc = c.input(), c.input(), c.input()
s, carry = c.full_adder(wa, wb, wc)
c.output(s); c.output(carry); c.output(c.mux(wa, wb, wc))
bits = np.array([[0, 1, 1], [1, 1, 0], [1, 0, 1]], dtype=bool)      # [input][set]
inp = np.stack([K.encrypt_bool(bits[i], rng) for i in range(3)])
res = CI.evaluate(c, e, inp)
dec = np.stack([K.decrypt_bool(res[i]) for i in range(3)])
want = np.stack([bits[0] ^ bits[1] ^ bits[2], (bits[0] & bits[1]) | ((bits[0] ^ bits[1]) & bits[2]), np.where(bits[0], bits[1], bits[2])])
print("circuit equal:", np.array_equal(dec, want))
e.close()
