"""Residency scaling: one persistent round with k ciphertext groups per SM (counts 148*k), blind-rotate ms."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TFHE_BR_LATENCY_MAX", "0")
import numpy as np
import rs_tfhe_b200 as T
P = T.PARAMS_BY_NAME["128"]
r = np.random.default_rng(1)
eng = T.CudaBootstrap(P, 0)
eng.generate_cloud_key(r.integers(0, 2, P.n, dtype=np.uint32), r.integers(0, 2, 1024, dtype=np.uint32), seed=7)
res = {}
for k in [int(x) for x in os.environ.get("RESID_K", "1,2,3,4,8").split(",")]:
    count = 148 * k
    pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
    best = 1e9
    for rep in range(3):
        eng.batch_gate("NAND", pairs)
        best = min(best, eng.last_kernel_ms()[0])
    res[k] = round(best, 3)
print(json.dumps({"variant": os.environ.get("TFHE_BR_VARIANT", "default"), "ms_by_groups": res}))
