"""Tiny workload for compute-sanitizer: 5 mixed gates + LUT + extract/key-switch through the C ABI."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import rs_tfhe_b200 as T
K = O.Keys("128", seed=1)
e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
e.load_cloud_key(T.CloudKey(T.SECURITY_128_BIT, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk))
rng = O.Rng(2)
count = int(os.environ.get("SANITIZE_COUNT", "5"))   # > 592 reaches the full-round kernel as well
a = np.resize(np.array([0, 1, 1, 0, 1], dtype=bool), count); b = np.resize(np.array([1, 1, 0, 0, 1], dtype=bool), count)
pairs = np.stack([K.encrypt_bool(a, rng), K.encrypt_bool(b, rng)], axis=1)
ops = np.resize(np.array([0, 1, 2, 3, 5], dtype=np.uint8), count)
out = e.batch_gate_mixed(ops, pairs)
ref = K.batch_gate(ops, pairs)
print("gates equal:", np.array_equal(out, ref))
lut_id, lut_b = e.lut_generate([1, 0], 2)
ct = K.encrypt_message([1, 0], 2, rng)
print("lut equal:", np.array_equal(e.batch_bootstrap_lut(lut_id, ct), K.batch_bootstrap(ct, lut_b=lut_b)))
tr = e.batch_blind_rotate(ct)
print("ks equal:", np.array_equal(e.batch_extract_key_switch(tr), K.batch_bootstrap(ct)))
e.close()
