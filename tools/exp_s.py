"""Parameter sweep of the 128-thread kernel: env settings -> blind-rotation ms (and bit-exactness on a
ragged batch vs the 64-thread kernel's output).  usage: exp_s.py "K=V,K=V;K=V" [count]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import rs_tfhe_b200 as T
    count = int(sys.argv[2])
    P = T.PARAMS_BY_NAME["128"]
    r = np.random.default_rng(1)
    eng = T.CudaBootstrap(P, 0)
    eng.generate_cloud_key(r.integers(0, 2, P.n, dtype=np.uint32), r.integers(0, 2, 1024, dtype=np.uint32), seed=7)
    pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
    best = 1e9
    out = None
    for rep in range(3):
        out = eng.batch_gate("NAND", pairs)
        best = min(best, eng.last_kernel_ms()[0])
    import zlib
    print("RESULT " + json.dumps({"br_ms": best, "crc": zlib.crc32(out.tobytes())}), flush=True)
    sys.exit(0)
configs = sys.argv[1].split(";")
count = sys.argv[2] if len(sys.argv) > 2 else "14208"
res = {}
for c in configs:
    env = dict(os.environ)
    for kv in filter(None, c.split(",")):
        k, v = kv.split("=")
        env[k] = v
    try:
        p = subprocess.run([sys.executable, __file__, "--child", count], env=env, capture_output=True, text=True, timeout=int(os.environ.get("EXP_TIMEOUT", "120")))
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        res[c] = json.loads(line[0][7:]) if line else {"error": (p.stderr or p.stdout)[-400:]}
    except subprocess.TimeoutExpired:
        res[c] = {"error": "timeout"}
    print(c, res[c], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_s.json"), "w"), indent=1)
