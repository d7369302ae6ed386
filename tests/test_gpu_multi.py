"""One engine over several GPUs of one process (tfhe_engine_create_multi): NCCL key broadcast inside
load_cloud_key, contiguous sharding of every batch call, LUT tables mirrored on all devices.  Needs
>= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

import oracle as O
import rs_tfhe_b200 as T
from common import GATE_FN, bool_pairs, keys

pytestmark = pytest.mark.gpu


def _need2():
    if T.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_multi_engine_matches_oracle_on_every_shard():
    _need2()
    n_dev = min(T.device_count(), 8)
    K, ck = keys("128")
    e = T.CudaBootstrap(T.SECURITY_128_BIT, list(range(n_dev)))
    assert e.n_devices == n_dev
    e.load_cloud_key(ck)                                  # upload on device 0 + ncclBroadcast
    assert e.last_broadcast_ms() > 0.0
    r = np.random.default_rng(5)
    rng = O.Rng(5)
    count = 311 * n_dev + 3                               # ragged shards
    a = r.integers(0, 2, count).astype(bool)
    b = r.integers(0, 2, count).astype(bool)
    ops = r.integers(0, 10, count).astype(np.uint8)
    pairs = bool_pairs(K, a, b, rng)
    got = e.batch_gate_mixed(ops, pairs)
    ref = K.batch_gate(ops, pairs)
    bad = np.nonzero((got != ref).any(axis=1))[0]
    assert bad.size == 0, f"mismatching gates {bad[:8]} (shard = index * {n_dev} // {count})"
    want = np.array([GATE_FN[T.GATES[o]](x, y) for o, x, y in zip(ops, a, b)]).astype(bool)
    assert np.array_equal(K.decrypt_bool(got), want)
    # same call on a one-GPU engine: identical words
    e1 = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    e1.load_cloud_key(ck)
    assert np.array_equal(e1.batch_gate_mixed(ops, pairs), got)
    # count smaller than the device count, and an empty batch
    assert np.array_equal(e.batch_gate("NAND", pairs[:1]), K.batch_gate(0, pairs[:1]))
    assert e.batch_gate("NAND", pairs[:0]).shape == (0, 701)
    e1.close()
    e.close()


def test_multi_engine_luts_on_all_devices():
    _need2()
    n_dev = min(T.device_count(), 8)
    K, ck = keys("uint4")
    e = T.CudaBootstrap(T.PARAMS_BY_NAME["uint4"], list(range(n_dev)))
    e.load_cloud_key(ck)
    m = 16
    rng = O.Rng(9)
    count = 77 * n_dev + 1
    msgs = np.arange(count) % m
    cts = K.encrypt_message(msgs, m, rng)
    gen = T.Generator(m, e)
    lut = gen.generate_lookup_table(lambda x: (5 * x + 3) % m)
    out = e.batch_bootstrap_lut(lut.lut_id, cts)          # every shard finds the table in its own HBM
    assert np.array_equal(K.decrypt_message(out, m), (5 * msgs + 3) % m)
    out = T.LutBootstrap(e).bootstrap_func(cts, lambda x: (x * x) % m, m)
    assert np.array_equal(K.decrypt_message(out, m), (msgs * msgs) % m)
    e.close()
