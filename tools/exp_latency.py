"""Latency of small batches: one PBS, 148 PBS, and the nibble-adder chain, for the selected latency kernel."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle as O
import rs_tfhe_b200 as T
name = sys.argv[1] if len(sys.argv) > 1 else "128"
P = T.PARAMS_BY_NAME[name]
K = O.Keys(name, seed=0x5EED0001)
e = T.CudaBootstrap(P, 0)
e.load_cloud_key(T.CloudKey(P, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk))
r = np.random.default_rng(3)
res = {"params": name, "kernel": os.environ.get("TFHE_BR_LATENCY_KERNEL", "s")}
for count in (1, 2, 37, 148):
    cts = r.integers(0, 2**32, (count, P.n + 1), dtype=np.uint32)
    got = e.bootstrap(cts)
    if P.l == 3:
        ref = K.batch_bootstrap(cts[:2], key_switch=True)
        res[f"equal_{count}"] = bool(np.array_equal(got[:2], ref))
    best = 1e9
    for _ in range(10):
        t = time.perf_counter(); e.bootstrap(cts); best = min(best, time.perf_counter() - t)
    res[f"ms_{count}"] = round(best * 1e3, 3)
    res[f"br_ms_{count}"] = round(e.last_kernel_ms()[0], 3)
print(json.dumps(res))
