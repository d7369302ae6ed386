"""End-to-end rate of mid-size host calls (pageable buffers): best of 7 wall-clock timings per count, and the
output CRC (compare across library builds).  usage: exp_midsize.py [counts]"""
import json, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rs_tfhe_b200 as T
counts = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [592, 1024, 1184, 2048, 4096, 4736]
P = T.PARAMS_BY_NAME["128"]
r = np.random.default_rng(1)
eng = T.CudaBootstrap(P, 0)
eng.generate_cloud_key(r.integers(0, 2, P.n, dtype=np.uint32), r.integers(0, 2, 1024, dtype=np.uint32), seed=7)
res = {}
for count in counts:
    pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
    best = 1e9
    for rep in range(7):
        t = time.perf_counter()
        out = eng.batch_gate("NAND", pairs)
        best = min(best, time.perf_counter() - t)
    br, ks = eng.last_kernel_ms()
    res[count] = {"e2e_ms": round(best * 1e3, 3), "gates_per_s": round(count / best), "kernels_ms": round(br + ks, 3),
                  "crc": zlib.crc32(out.tobytes())}
    print(count, res[count], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_midsize.json"), "w"), indent=1)
