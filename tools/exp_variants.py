"""A/B of blind-rotation kernel variants: bit-exactness vs the oracle on a ragged batch, then
device time at a throughput batch.  usage: exp_variants.py 3,7,8 [count]   (development aid)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import oracle as O
    import rs_tfhe_b200 as T
    count = int(sys.argv[2])
    K = O.Keys("128", seed=0x5EED0001)
    P = T.PARAMS_BY_NAME["128"]
    eng = T.CudaBootstrap(P, 0)
    eng.load_cloud_key(T.CloudKey(P, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk))
    r = np.random.default_rng(7)
    res = {}
    for m in (601, 1813):   # ragged: partial last round, idle tail slots
        pairs = r.integers(0, 2**32, (m, 2, P.n + 1), dtype=np.uint32)
        ops = r.integers(0, 10, m).astype(np.uint8)
        res[f"equal_{m}"] = bool(np.array_equal(eng.batch_gate_mixed(ops, pairs), K.batch_gate(ops, pairs)))
    pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
    best = 1e9
    for rep in range(3):
        eng.batch_gate("NAND", pairs)
        br, ks = eng.last_kernel_ms()
        best = min(best, br)
    res["br_ms"] = best; res["count"] = count
    print("RESULT " + json.dumps(res), flush=True)
    sys.exit(0)
variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["3"]
count = sys.argv[2] if len(sys.argv) > 2 else "14208"
out = {}
for v in variants:
    env = dict(os.environ, TFHE_BR_VARIANT=v, TFHE_BR_LATENCY_MAX="0")
    try:
        p = subprocess.run([sys.executable, __file__, "--child", count], env=env, capture_output=True, text=True,
                           timeout=int(os.environ.get("EXP_TIMEOUT", "240")))
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        out[v] = json.loads(line[0][7:]) if line else {"error": (p.stderr or p.stdout)[-600:]}
    except subprocess.TimeoutExpired:
        out[v] = {"error": "timeout (hang?)"}
    print(v, out[v], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "exp_variants.json"), "w"), indent=1)
