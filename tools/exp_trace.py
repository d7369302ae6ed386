"""Phase timeline of the 128-thread kernel on SM 0 (ablation build only): clock64 at the phase
boundaries of steps 100..103 for lane 0 of every consumer warp."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TFHE_S_ABLATE"] = "8"
os.environ.setdefault("TFHE_BR_LATENCY_MAX", "0")
import numpy as np
import rs_tfhe_b200 as T
P = T.PARAMS_BY_NAME["128"]
r = np.random.default_rng(1)
eng = T.CudaBootstrap(P, 0)
eng.generate_cloud_key(r.integers(0, 2, P.n, dtype=np.uint32), r.integers(0, 2, 1024, dtype=np.uint32), seed=7)
pairs = r.integers(0, 2**32, (592, 2, P.n + 1), dtype=np.uint32)
eng.batch_gate("NAND", pairs)
eng.batch_gate("NAND", pairs)
L = T._load()
buf = (C.c_ulonglong * (16 * 4 * 16))()
assert L.tfhe_debug_read_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(16, 4, 16)
t0 = t[:, 0, 0].min()
names = ["step", "A0done", "bar", "row0", "rows", "bar", "A1done", "bar", "row3", "rows", "bar", "inv1", "bar", "A'", "bar"]
for w in range(16):
    print("warp", w, "g", w // 4, "W", w % 4)
    for st in range(3):
        rel = t[w, st, :15] - t0
        print("   step", 100 + st, " ".join(f"{int(x):6d}" for x in rel))
print("step period (warp0):", int(t[0, 1, 0] - t[0, 0, 0]), int(t[0, 2, 0] - t[0, 1, 0]))
