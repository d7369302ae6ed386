"""GPU parity tests (run on a B200 via gpurun): the CUDA path, called through the
C ABI, against the oracle on identical seeded keys and ciphertexts.
Parity classes follow SURVEY.md 8c: P0 integer-exact, P1 f64-exact at l=3 sets,
P2 phase-level at l<3 sets, P3 semantic (decrypt == plaintext function)."""
import numpy as np
import pytest

import oracle as O
import rs_tfhe_b200 as T
from common import GATE_FN, bool_pairs, keys, torus_dist

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng128():
    K, ck = keys("128", with_torus_bsk=True)
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    e.load_cloud_key(ck)
    yield K, ck, e
    e.close()


def test_native_library_is_the_path(eng128):
    _, _, e = eng128
    assert T.device_count() >= 1
    before = e.kernel_launches
    K = eng128[0]
    pairs = bool_pairs(K, [1], [1], O.Rng(1))
    e.batch_gate("NAND", pairs)
    assert e.kernel_launches >= before + 2   # blind rotate + key switch


def test_lut_generation_p0(eng128):
    _, _, e = eng128
    for m, f in [(2, lambda x: x), (2, lambda x: 1 - x), (4, lambda x: (3 * x) % 4),
                 (16, lambda x: (x * x) % 16), (32, lambda x: x % 16), (32, lambda x: int(x >= 16)),
                 (3, lambda x: x), (256, lambda x: 255 - x)]:
        table = [f(x) for x in range(m)]
        _, got = e.lut_generate(table, m)
        assert np.array_equal(got, O.lut_generate(table, m)), m


def test_blind_rotate_p1_128(eng128):
    K, _, e = eng128
    rng = O.Rng(21)
    cts = K.encrypt_bool([1, 0, 1, 1, 0], rng)
    cts[3] = O.Rng(5).u32() * np.ones(701, dtype=np.uint32)  # degenerate: all words equal
    got = e.batch_blind_rotate(cts)
    ref = K.batch_blind_rotate(cts)
    assert np.array_equal(got, ref)
    ea, eb, _ = K.blind_rotate(cts[0], exact=True)       # exact-integer ground truth
    assert np.array_equal(got[0, 0], ea) and np.array_equal(got[0, 1], eb)


def test_extract_key_switch_p0(eng128):
    K, _, e = eng128
    r = np.random.default_rng(3)
    trlwe = r.integers(0, 2**32, (11, 2, 1024), dtype=np.uint32)
    trlwe[1] = 0
    trlwe[2] = 0xFFFFFFFF
    got = e.batch_extract_key_switch(trlwe)
    ref = np.stack([K.identity_key_switching(O.sample_extract_index(t[0], t[1], 0)) for t in trlwe])
    assert np.array_equal(got, ref)


def test_all_gates_bit_exact_and_truth_tables(eng128):
    K, _, e = eng128
    rng = O.Rng(31)
    a = np.array([0, 0, 1, 1], dtype=bool)
    b = np.array([0, 1, 0, 1], dtype=bool)
    pairs = bool_pairs(K, a, b, rng)
    for op in T.GATES:
        got = e.batch_gate(op, pairs)
        ref = K.batch_gate(O.GATE_CODE[op], pairs)
        assert np.array_equal(got, ref), op                      # P1: every LWE word
        assert np.array_equal(K.decrypt_bool(got), GATE_FN[op](a, b)), op   # P3


def test_mixed_gate_batch_ragged(eng128):
    K, _, e = eng128
    rng = O.Rng(41)
    r = np.random.default_rng(41)
    for count in (1, 3, 149, 601):
        a = r.integers(0, 2, count).astype(bool)
        b = r.integers(0, 2, count).astype(bool)
        ops = r.integers(0, 10, count).astype(np.uint8)
        pairs = bool_pairs(K, a, b, rng)
        got = e.batch_gate_mixed(ops, pairs)
        ref = K.batch_gate(ops, pairs)
        assert np.array_equal(got, ref), count
        want = np.array([GATE_FN[T.GATES[o]](x, y) for o, x, y in zip(ops, a, b)]).astype(bool)
        assert np.array_equal(K.decrypt_bool(got), want), count


def test_empty_batch(eng128):
    _, _, e = eng128
    out = e.batch_gate("NAND", np.empty((0, 2, 701), dtype=np.uint32))
    assert out.shape == (0, 701)


def test_bootstrap_trait(eng128):
    K, _, e = eng128
    rng = O.Rng(51)
    bits = np.array([1, 0, 1], dtype=bool)
    cts = K.encrypt_bool(bits, rng)
    got = e.bootstrap(cts)
    assert np.array_equal(got, K.batch_bootstrap(cts, key_switch=True))
    assert np.array_equal(K.decrypt_bool(got), bits)
    got2 = e.bootstrap_without_key_switch(cts)
    assert np.array_equal(got2, K.batch_bootstrap(cts, key_switch=False))
    one = e.bootstrap(cts[0])
    assert one.shape == (701,) and np.array_equal(one, got[0])


def test_gates_api_mux_and_helpers(eng128):
    K, ck, e = eng128
    g = T.Gates.with_bootstrap(e)
    assert g.bootstrap_strategy() == "cuda-b200"
    rng = O.Rng(61)
    for a, b, c in [(0, 1, 0), (1, 1, 0), (1, 0, 1), (0, 0, 1)]:
        ca, cb, cc = (K.encrypt_bool([x], rng)[0] for x in (a, b, c))
        out = g.mux_naive(ca, cb, cc)
        assert bool(K.decrypt_bool(out)[0]) == bool(b if a else c)
    x = K.encrypt_bool([1], rng)[0]
    assert not K.decrypt_bool(g.not_(x))[0]
    assert K.decrypt_bool(g.constant(True))[0] and not K.decrypt_bool(g.constant(False))[0]
    # data-flow parity of the reference's optimised mux (gates.rs:157-183)
    ca, cb, cc = (K.encrypt_bool([v], rng)[0] for v in (1, 0, 1))
    t_and = K.gate_prep(O.GATE_CODE["AND"], ca, cb)
    t_ny = K.gate_prep(O.GATE_CODE["ANDNY"], ca, cc)
    u = K.batch_bootstrap(np.stack([t_and, t_ny]), key_switch=False)
    t_or = K.gate_prep(O.GATE_CODE["OR"], u[0], u[1])
    ref = K.batch_bootstrap(t_or[None], key_switch=True)[0]
    assert np.array_equal(g.mux(ca, cb, cc), ref)


def test_lut_bootstrap_binary_on_gate_params(eng128):
    """bootstrap/lut.rs:141-254: identity / NOT / constant at modulus 2, bit-exact vs oracle."""
    K, _, e = eng128
    lb = T.LutBootstrap(e)
    rng = O.Rng(71)
    for f in (lambda x: x, lambda x: 1 - x, lambda x: 1):
        lut = T.Generator(2, e).generate_lookup_table(f)
        ref_lut = O.lut_generate([f(0) % 2, f(1) % 2], 2)
        assert np.array_equal(lut.poly_b, ref_lut)
        for msg in (0, 1):
            ct = K.encrypt_message([msg], 2, rng)
            got = lb.bootstrap_lut(ct, lut)
            assert np.array_equal(got, K.batch_bootstrap(ct, lut_b=ref_lut))
            assert K.decrypt_message(got, 2)[0] == f(msg) % 2
    ct = K.encrypt_message([1], 2, rng)[0]
    assert K.decrypt_message(lb.bootstrap_func(ct, lambda x: 1 - x, 2), 2)[0] == 0


def test_lut_multi_table_batch_and_nibble_adder(eng128):
    """examples/lut_add_two_numbers.rs:80-157 (modulus 32 on gate params): sum and carry tables
    read the same input, so they go out as one batch with per-ciphertext table ids.  Judged on
    word-for-word equality with the oracle (the algorithm itself is noise-marginal here)."""
    K, _, e = eng128
    m = 32
    tabs = {"low": [x % 16 for x in range(m)], "carry": [int(x >= 16) for x in range(m)]}
    ids = {k: e.lut_generate(v, m) for k, v in tabs.items()}
    rng = O.Rng(111)
    a, b = 42, 137
    enc = lambda v: K.encrypt_message([v], m, rng)[0]
    a_lo, a_hi, b_lo, b_hi = enc(a & 15), enc(a >> 4), enc(b & 15), enc(b >> 4)
    ct_low = (a_lo + b_lo).astype(np.uint32)
    both = e.batch_bootstrap_lut([ids["low"][0], ids["carry"][0]], np.stack([ct_low, ct_low]))
    ref_lo = K.batch_bootstrap(ct_low[None], lut_b=ids["low"][1])[0]
    ref_c = K.batch_bootstrap(ct_low[None], lut_b=ids["carry"][1])[0]
    assert np.array_equal(both[0], ref_lo) and np.array_equal(both[1], ref_c)
    ct_hi = (a_hi + b_hi + both[1]).astype(np.uint32)
    s_hi = e.batch_bootstrap_lut(ids["low"][0], ct_hi)
    assert np.array_equal(s_hi, K.batch_bootstrap(ct_hi[None], lut_b=ids["low"][1])[0])


@pytest.mark.parametrize("name", ["80", "110"])
def test_other_gate_sets_bit_exact(name):
    K, ck = keys(name)
    e = T.CudaBootstrap(T.PARAMS_BY_NAME[name], 0)
    try:
        e.load_cloud_key(ck)
        rng = O.Rng(81)
        a = np.array([0, 1, 1, 0, 1], dtype=bool)
        b = np.array([1, 1, 0, 0, 1], dtype=bool)
        pairs = bool_pairs(K, a, b, rng)
        got = e.batch_gate("NAND", pairs)
        assert np.array_equal(got, K.batch_gate(O.GATE_CODE["NAND"], pairs))
        assert np.array_equal(K.decrypt_bool(got), ~(a & b))
    finally:
        e.close()


@pytest.mark.parametrize("name", ["uint1", "uint2", "uint3", "uint4", "uint5", "uint7"])
def test_extract_key_switch_p0_other_bases(name):
    """Key switch is integer-exact on every parameter set: basebit 2/4/5/6 run on tcgen05.mma
    (one-hot over 2^basebit digit values), basebit 7 on the row-walk kernel (trgsw.rs:332-360)."""
    K, ck = keys(name, seed=0x5EED0007)
    e = T.CudaBootstrap(T.PARAMS_BY_NAME[name], 0)
    try:
        e.load_cloud_key(ck)
        r = np.random.default_rng(11)
        trlwe = r.integers(0, 2**32, (137, 2, 1024), dtype=np.uint32)   # ragged: one full + one partial M tile
        trlwe[1] = 0
        trlwe[2] = 0xFFFFFFFF
        got = e.batch_extract_key_switch(trlwe)
        ref = np.stack([K.identity_key_switching(O.sample_extract_index(t[0], t[1], 0)) for t in trlwe])
        assert np.array_equal(got, ref)
    finally:
        e.close()


def test_uint4_lut_p2():
    """l=1, Bg=2^22: f64 FFT is inexact and mask words decorrelate, so
    parity is phase-level: equal decryptions and |phase_gpu - phase_oracle| <= 4e-3
    (about 6 sigma of the measured PBS noise 6e-4; slot half-width is 1.56e-2)."""
    K, ck = keys("uint4", seed=0x5EED0003)
    e = T.CudaBootstrap(T.SECURITY_UINT4, 0)
    try:
        e.load_cloud_key(ck)
        m = 16
        rng = O.Rng(91)
        msgs = np.arange(48) % m
        cts = K.encrypt_message(msgs, m, rng)
        for f in (lambda x: x, lambda x: (x * x) % 16):
            table = [f(x) for x in range(m)]
            lut_id, lut_b = e.lut_generate(table, m)
            assert np.array_equal(lut_b, O.lut_generate(table, m))
            got = e.batch_bootstrap_lut(lut_id, cts)
            ref = K.batch_bootstrap(cts, lut_b=lut_b)
            want = np.array([f(int(x)) for x in msgs])
            assert np.array_equal(K.decrypt_message(got, m), want)
            assert np.array_equal(K.decrypt_message(ref, m), want)
            assert torus_dist(K.phase(got), K.phase(ref)).max() < 4e-3
    finally:
        e.close()


def test_large_batch_semantic_128(eng128):
    """Full-size batch (> one wave of the persistent grid): decrypt == plaintext NAND on
    every element, plus a spot check of 64 elements against the oracle, word for word."""
    K, _, e = eng128
    count = 148 * 4 * 3 + 37
    r = np.random.default_rng(7)
    a = r.integers(0, 2, count).astype(bool)
    b = r.integers(0, 2, count).astype(bool)
    pairs = bool_pairs(K, a, b, O.Rng(101))
    got = e.batch_gate("NAND", pairs)
    assert np.array_equal(K.decrypt_bool(got), ~(a & b))
    idx = r.choice(count, 64, replace=False)
    assert np.array_equal(got[idx], K.batch_gate(O.GATE_CODE["NAND"], pairs[idx]))


def test_proxy_reencryption_p0(eng128):
    """proxy_reenc::reencrypt_tlwe_lv0 (src/proxy_reenc.rs:468-511) is the key-switch
    kernel with the level-0 dimension as input -- bit-exact vs the oracle, and Bob decrypts
    (proxy_reenc.rs tests :519-703)."""
    alice, _, e = eng128
    bob = O.Keys("128", seed=0xB0B)
    rk = O.gen_reenc_key(alice, bob, seed=77)
    p = alice.params
    key = e.load_reenc_key(rk, 1 << p.basebit, p.iks_t)
    bits = np.array([1, 0, 1, 1, 0, 0, 1, 0, 1], dtype=bool)
    cts = alice.encrypt_bool(bits, O.Rng(5))
    got = e.batch_reencrypt(key, cts)
    ref = np.stack([O.reencrypt(p, rk, p.basebit, p.iks_t, c) for c in cts])
    assert np.array_equal(got, ref)
    assert np.array_equal(bob.decrypt_bool(got), bits)
    key.close()


def test_cloud_key_blob_export_import(eng128, tmp_path):
    """the re-laid-out key round-trips through the blob format (via a file) into a
    fresh engine, LUT slots included, and evaluates identically."""
    K, _, e = eng128
    lut_id, lut_b = e.lut_generate([1, 0], 2)
    blob = e.export_cloud_key()
    assert bytes(blob[:8]) == b"TFHEB200"
    path = tmp_path / "key.blob"
    blob.tofile(path)
    e2 = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    try:
        e2.import_cloud_key(np.fromfile(path, dtype=np.uint8))
        rng = O.Rng(7)
        pairs = bool_pairs(K, [1, 0, 1], [1, 1, 0], rng)
        assert np.array_equal(e2.batch_gate("XOR", pairs), e.batch_gate("XOR", pairs))
        ct = K.encrypt_message([0, 1], 2, rng)
        assert np.array_equal(e2.batch_bootstrap_lut(lut_id, ct), K.batch_bootstrap(ct, lut_b=lut_b))
        bad = blob.copy(); bad[0] ^= 1
        with pytest.raises(T.EngineError):
            e2.import_cloud_key(bad)
        e80 = T.CudaBootstrap(T.SECURITY_80_BIT, 0)
        with pytest.raises(T.EngineError, match="parameters differ"):
            e80.import_cloud_key(blob)
        e80.close()
    finally:
        e2.close()


def test_device_keygen_structure_and_function():
    """CloudKey::new on the device.  The reference's RNG is unseeded, so the check
    is structural and exact: every generated TRGSW row, pulled back from the device blob and
    inverse-transformed, satisfies b - a*s1 = noise + gadget term (exact integer product), KSK
    rows decrypt to k*s1_i/2^((j+1)basebit) within the noise, and gates evaluated with the
    device-made key decrypt correctly with the expected PBS noise level."""
    from test_fft_layout_model import inverse_model
    K, _ = keys("128")
    p = K.params
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    try:
        e.generate_cloud_key(K.s0, K.s1, seed=0xC0FFEE)
        blob = e.export_cloud_key()[64:]
        l2 = 2 * p.l
        bsk = blob[: p.n * l2 * 1024 * 16].view(np.float64).reshape(p.n, l2, 8, 2, 64, 2)
        s1 = K.s1
        for (i, r) in [(0, 0), (0, 3), (5, 2), (699, 5), (123, 1)]:
            polys = []
            for o in range(2):
                G = bsk[i, r, :, o, :, 0].T + 1j * bsk[i, r, :, o, :, 1].T     # [v][k2]
                x = inverse_model(G)
                assert np.abs(x - np.round(x)).max() < 1e-3
                polys.append((np.round(x).astype(np.int64) & 0xFFFFFFFF).astype(np.uint32))
            a, b = polys
            phase = (b.astype(np.int64) - O.poly_mul_exact(a, s1).astype(np.int64)) & 0xFFFFFFFF
            gad = int(K.s0[i]) << (32 - ((r % p.l) + 1) * p.bgbit)
            if r >= p.l:
                phase[0] = (phase[0] - gad) & 0xFFFFFFFF
            else:
                phase = (phase + gad * s1.astype(np.int64)) & 0xFFFFFFFF
            noise = ((phase + 2**31) % 2**32 - 2**31) / 2.0**32
            assert np.abs(noise).max() < 8 * p.alpha_lv1, (i, r)
            assert 0.5 * p.alpha_lv1 < noise.std() < 1.5 * p.alpha_lv1, (i, r)
            assert len(np.unique(a)) > 1000                                  # a is uniform, not degenerate
        # KSK rows (row layout region follows the BSK, 256-byte aligned)
        off = (p.n * l2 * 1024 * 16 + 255) // 256 * 256
        stride = (p.n + 1 + 3) // 4 * 4
        ksk = blob[off: off + p.ksk_rows * stride * 4].view(np.uint32).reshape(p.ksk_rows, stride)
        for (i, j, k) in [(0, 0, 1), (17, 3, 2), (1023, 8, 3), (500, 5, 0)]:
            row = ksk[(i * p.iks_t + j) * 4 + k, : p.n + 1]
            if k == 0:
                assert not row.any()
                continue
            ph = int(K.phase(row[None])[0])
            want = (k * int(K.s1[i])) << (32 - (j + 1) * p.basebit)
            d = ((ph - want + 2**31) % 2**32 - 2**31) / 2.0**32
            assert abs(d) < 8 * p.alpha_lv0
        # functional: truth tables + noise level of 4096 random gates
        r_ = np.random.default_rng(9)
        a_ = r_.integers(0, 2, 4096).astype(bool)
        b_ = r_.integers(0, 2, 4096).astype(bool)
        pairs = np.stack([K.encrypt_bool_batch(a_, 71), K.encrypt_bool_batch(b_, 72)], axis=1)
        out = e.batch_gate("NAND", pairs)
        assert np.array_equal(K.decrypt_bool_batch(out), ~(a_ & b_))
        ph = K.phase_batch(out).astype(np.int64)
        ideal = np.where(~(a_ & b_), 0x20000000, 0xE0000000)
        err = ((ph - ideal + 2**31) % 2**32 - 2**31) / 2.0**32
        assert 0.007 < err.std() < 0.015
    finally:
        e.close()
