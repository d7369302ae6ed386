"""One batch of NAND gates under a device-generated key (for ncu captures and quick timings).
usage: prof_run.py [count] [reps] [params]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rs_tfhe_b200 as T

count = int(sys.argv[1]) if len(sys.argv) > 1 else 3552
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P = T.PARAMS_BY_NAME[sys.argv[3] if len(sys.argv) > 3 else "128"]
r = np.random.default_rng(1)
eng = T.CudaBootstrap(P, 0)
eng.generate_cloud_key(r.integers(0, 2, P.n, dtype=np.uint32), r.integers(0, 2, 1024, dtype=np.uint32), seed=7)
pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
best = 1e9
for _ in range(reps):
    eng.batch_gate("NAND", pairs)
    best = min(best, eng.last_kernel_ms()[0])
print(json.dumps({"count": count, "br_ms": best, "gates_per_s": count / best * 1e3}))
