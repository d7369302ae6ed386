"""Quick device-side timing of the batch-gate path (development aid, not bench.py)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import rs_tfhe_b200 as T

name = sys.argv[1] if len(sys.argv) > 1 else "128"
counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [592, 2368, 16384]
K = O.Keys(name, seed=0x5EED0001)
P = T.PARAMS_BY_NAME[name]
ck = T.CloudKey(P, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk)
eng = T.CudaBootstrap(P, 0)
t = time.time(); eng.load_cloud_key(ck); print("load_cloud_key s", time.time() - t)
r = np.random.default_rng(0)
res = []
for count in counts:
    pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
    for rep in range(3):
        t = time.time()
        out = eng.batch_gate("NAND", pairs)
        wall = time.time() - t
        br, ks = eng.last_kernel_ms()
        print(f"{name} count={count} rep={rep} wall={wall*1e3:.1f} ms  br={br:.2f} ms ks={ks:.2f} ms  "
              f"gates/s(kernels)={count/((br+ks)*1e-3):.0f}  us/PBS={(br+ks)*1e3/count:.2f}", flush=True)
    res.append(dict(name=name, count=count, wall_ms=wall * 1e3, br_ms=br, ks_ms=ks))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/quick_bench_{name}.json", "w"))
