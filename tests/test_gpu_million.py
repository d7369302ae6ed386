""">= 10^6 random trials through the C ABI (north star: "decrypted gate and LUT outputs are
bit-exact on >= 10^6 random trials"), in chunks to bound host memory; also reports the
measured output-noise statistics next to the gate margin."""
import json
import os

import numpy as np
import pytest

import oracle as O
import rs_tfhe_b200 as T
from common import GATE_FN, keys

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(name, obj):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", name), "w") as f:
        json.dump(obj, f, indent=1)
    print(name, json.dumps(obj))


def test_million_mixed_gates_128():
    K, ck = keys("128")
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    try:
        e.load_cloud_key(ck)
        total, chunk = 1 << 20, 1 << 18
        r = np.random.default_rng(0x5EED0005)
        wrong = 0
        err = []
        for c in range(total // chunk):
            a = r.integers(0, 2, chunk).astype(bool)
            b = r.integers(0, 2, chunk).astype(bool)
            ops = r.integers(0, 10, chunk).astype(np.uint8)
            pairs = np.stack([K.encrypt_bool_batch(a, 1000 + 2 * c),
                              K.encrypt_bool_batch(b, 1001 + 2 * c)], axis=1)
            out = e.batch_gate_mixed(ops, pairs)
            want = np.zeros(chunk, dtype=bool)
            for o, name in enumerate(T.GATES):
                m = ops == o
                want[m] = GATE_FN[name](a[m], b[m])
            ph = K.phase_batch(out)
            got = ph.view(np.int32) >= 0
            wrong += int((got != want).sum())
            ideal = np.where(want, 0x20000000, 0xE0000000).astype(np.int64)
            d = (ph.astype(np.int64) - ideal + 2**31) % 2**32 - 2**31
            err.append(d / 2.0**32)
            if c == 0:   # word-for-word against the oracle on a slice of the same batch
                idx = r.choice(chunk, 96, replace=False)
                assert np.array_equal(out[idx], K.batch_gate(ops[idx], pairs[idx]))
        err = np.concatenate(err)
        _report("noise_gates_128.json", {
            "trials": total, "wrong_decryptions": wrong, "phase_error_mean": float(err.mean()),
            "phase_error_std": float(err.std()), "phase_error_max_abs": float(np.abs(err).max()),
            "gate_margin": 0.125})
        assert wrong == 0
        assert np.abs(err).max() < 0.125
    finally:
        e.close()


def test_million_lut_uint4():
    K, ck = keys("uint4", seed=0x5EED0003)
    e = T.CudaBootstrap(T.SECURITY_UINT4, 0)
    try:
        e.load_cloud_key(ck)
        m = 16
        total, chunk = 1 << 20, 1 << 18
        r = np.random.default_rng(0x5EED0006)
        tables = {"identity": [x for x in range(m)], "square": [(x * x) % m for x in range(m)]}
        ids = {k: e.lut_generate(v, m)[0] for k, v in tables.items()}
        wrong = 0
        err = []
        for c in range(total // chunk):
            name = "identity" if c % 2 == 0 else "square"
            msgs = r.integers(0, m, chunk)
            cts = K.encrypt_message_batch(msgs, m, 2000 + c)
            out = e.batch_bootstrap_lut(ids[name], cts)
            want = np.array(tables[name])[msgs]
            ph = K.phase_batch(out)
            got = K.decrypt_message_batch(out, m)
            wrong += int((got != want).sum())
            ideal = (want.astype(np.int64) << 32) // (2 * m)
            d = (ph.astype(np.int64) - ideal + 2**31) % 2**32 - 2**31
            err.append(d / 2.0**32)
        err = np.concatenate(err)
        _report("noise_lut_uint4.json", {
            "trials": total, "wrong_decryptions": wrong, "phase_error_mean": float(err.mean()),
            "phase_error_std": float(err.std()), "phase_error_max_abs": float(np.abs(err).max()),
            "slot_half_width": 1.0 / (4 * m)})
        assert wrong == 0
    finally:
        e.close()
