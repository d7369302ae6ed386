"""examples/lut_add_two_numbers.rs:80-157 of the reference on the B200 engine: 8-bit addition by
nibbles with three programmable bootstraps (sum, carry, high sum).  The reference runs this on the
128-bit *gate* parameters with modulus 32, where the algorithm's own noise makes a nibble decode
wrongly in a large fraction of runs; SECURITY_UINT5 (messageModulus 32) is the
parameter set meant for it and decodes reliably, so that is the default here.
Usage: python examples/lut_add_two_numbers.py [a b] [--gate-params]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rs_tfhe_b200 as T
from rs_tfhe_b200.client import Client, SecretKey

args = [x for x in sys.argv[1:] if not x.startswith("--")]
a, b = (int(args[0]), int(args[1])) if len(args) >= 2 else (42, 137)   # lut_add_two_numbers.rs:52-54
P = T.SECURITY_128_BIT if "--gate-params" in sys.argv else T.SECURITY_UINT5
m = 32
sk = SecretKey.new(P)
engine = T.CudaBootstrap(P, 0)
engine.generate_cloud_key(sk.key_lv0, sk.key_lv1)
client = Client(sk)
gen = T.Generator(m, engine)
lut_low = gen.generate_lookup_table(lambda x: x % 16)                   # sum nibble
lut_carry = gen.generate_lookup_table(lambda x: 1 if x >= 16 else 0)    # carry
enc = lambda v: client.encrypt_lwe_message([v], m)[0]
a_lo, a_hi, b_lo, b_hi = enc(a & 15), enc(a >> 4), enc(b & 15), enc(b >> 4)

t = time.perf_counter()
ct_low = (a_lo + b_lo).astype(np.uint32)                                # homomorphic add (tlwe.rs:129-139)
both = engine.batch_bootstrap_lut([lut_low.lut_id, lut_carry.lut_id], np.stack([ct_low, ct_low]))
ct_hi = (a_hi + b_hi + both[1]).astype(np.uint32)
hi = engine.batch_bootstrap_lut(lut_low.lut_id, ct_hi)
dt = time.perf_counter() - t
res = int(client.decrypt_lwe_message(both[0], m)[0]) + 16 * int(client.decrypt_lwe_message(hi, m)[0])
print(f"{a} + {b} = {res} (mod 256)   [{dt * 1e3:.2f} ms, 3 PBS in 2 dependent levels, params {P.name}]")
assert res == (a + b) % 256 or "--gate-params" in sys.argv
