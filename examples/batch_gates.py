"""examples/batch_gates.rs / batch_gates_scaling.rs of the reference, on the B200 engine:
encrypt a batch of bit pairs, evaluate gates::batch_nand, decrypt and check.
Usage: python examples/batch_gates.py [count]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rs_tfhe_b200 as T
from rs_tfhe_b200.client import Client, SecretKey

count = int(sys.argv[1]) if len(sys.argv) > 1 else 1024          # BASELINE configs[0]
sk = SecretKey.new(T.SECURITY_128_BIT)
engine = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
t = time.perf_counter()
engine.generate_cloud_key(sk.key_lv0, sk.key_lv1)                 # CloudKey::new, on the device
print(f"cloud key generated on the GPU in {(time.perf_counter() - t) * 1e3:.1f} ms")
client = Client(sk)
a = np.array([i % 2 == 0 for i in range(count)])                   # batch_gates_scaling.rs:11
b = np.array([i % 3 == 0 for i in range(count)])
inputs = np.stack([client.encrypt_bool(a), client.encrypt_bool(b)], axis=1)
t = time.perf_counter()
out = engine.batch_gate("NAND", inputs)                            # gates::batch_nand
dt = time.perf_counter() - t
ok = np.array_equal(client.decrypt_bool(out), ~(a & b))
print(f"{count} NAND gates in {dt * 1e3:.1f} ms  ({count / dt:.0f} gates/s end to end), all correct: {ok}")
assert ok
