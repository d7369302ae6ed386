"""Round-2 GPU tests through the C ABI: the FFTProcessor seam (reference tests src/fft/mod.rs:118-255
re-expressed), unbounded bootstrap_func / LUT slot release / stale ids, the alloc -> blob copy ->
commit key path with two engines in one process, throughput-kernel parity at 80/110-bit and the
other advertised gadgets."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
import rs_tfhe_b200 as T
from common import GATE_FN, bool_pairs, keys, torus_dist

pytestmark = pytest.mark.gpu
N = 1024


@pytest.fixture(scope="module")
def eng128():
    K, ck = keys("128", with_torus_bsk=True)
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    e.load_cloud_key(ck)
    yield K, ck, e
    e.close()


def signed(x):
    return np.asarray(x, dtype=np.uint32).astype(np.int32).astype(np.int64)


# ---- FFT seam ------------------------------------------------------------------------------------
def test_fft_ifft_round_trip(eng128):
    """src/fft/mod.rs:118-133, 179-210: fft(ifft(a)) == a within +-1 for uniform torus inputs."""
    _, _, e = eng128
    r = np.random.default_rng(1)
    a = r.integers(0, 2**32, (37, N), dtype=np.uint32)
    a[0] = 0
    a[1, :] = 0; a[1, 0] = 1000                      # the delta test, mod.rs:161-177
    spec = e.batch_ifft(a)
    back = e.batch_fft(spec)
    d = signed(a) - signed(back)
    assert np.abs(d).max() < 2
    assert np.array_equal(back[:2], a[:2])           # exactly representable cases come back exactly
    # same spectrum as the oracle's restatement of klemsa.rs:88-117 (relative to the spectrum's scale)
    ref = np.stack([O.ifft(x) for x in a[:5]])
    scale = np.abs(ref).max()
    assert np.abs(spec[:5] - ref).max() <= 1e-12 * max(scale, 1.0)


def test_fft_seam_layout_known_answer(eng128):
    """ifft of X^0 (a[0] = 1): every bin is 2 x e^{0} = 2 + 0i; of X^1: 2 e^{i pi (1-4k)/1024}."""
    _, _, e = eng128
    a = np.zeros((2, N), dtype=np.uint32)
    a[0, 0] = 1
    a[1, 1] = 1
    s = e.batch_ifft(a)
    assert np.allclose(s[0, :512], 2.0, atol=1e-13) and np.allclose(s[0, 512:], 0.0, atol=1e-13)
    k = np.arange(512)
    ang = np.pi * (1 - 4 * k) / 1024
    assert np.allclose(s[1, :512], 2 * np.cos(ang), atol=1e-12)
    assert np.allclose(s[1, 512:], 2 * np.sin(ang), atol=1e-12)


def test_fft_poly_mul_vs_schoolbook(eng128):
    """src/fft/mod.rs:135-159, 212-238: a uniform, b < Bg = 64; within +-1 of the exact O(N^2)
    negacyclic product (here: exactly equal, the f64 error is << 0.5 at these sizes), 100 trials."""
    _, _, e = eng128
    r = np.random.default_rng(2)
    a = r.integers(0, 2**32, (100, N), dtype=np.uint32)
    b = r.integers(0, 64, (100, N), dtype=np.uint32)
    got = e.batch_poly_mul(a, b)
    for i in range(100):
        want = O.poly_mul_exact(a[i], b[i])
        d = signed(got[i]) - signed(want)
        assert np.abs(d).max() < 2, i
    assert np.array_equal(got[0], O.poly_mul_exact(a[0], b[0]))


def test_fft_seam_ragged_counts(eng128):
    _, _, e = eng128
    r = np.random.default_rng(3)
    for count in (1, 2, 3, 5, 149, 593):
        a = r.integers(0, 2**32, (count, N), dtype=np.uint32)
        b = r.integers(0, 64, (count, N), dtype=np.uint32)
        got = e.batch_poly_mul(a, b)
        j = count - 1
        assert np.abs(signed(got[j]) - signed(O.poly_mul_exact(a[j], b[j]))).max() < 2, count
    assert e.batch_ifft(np.empty((0, N), dtype=np.uint32)).shape == (0, N)


# ---- LUT slots -------------------------------------------------------------------------------------
def test_bootstrap_func_1000_calls_and_slot_release():
    """bootstrap_func (bootstrap/lut.rs:49-65) builds and drops a table per call in the reference; it
    must be callable without bound here too.  Explicit tables release their slot when dropped."""
    K, ck = keys("uint4")
    e = T.CudaBootstrap(T.PARAMS_BY_NAME["uint4"], 0)
    e.load_cloud_key(ck)
    lb = T.LutBootstrap(e)
    m = 16
    rng = O.Rng(71)
    msgs = np.arange(16) % m
    cts = K.encrypt_message(msgs, m, rng)
    wrong = 0
    for it in range(1000):
        out = lb.bootstrap_func(cts[it % 16], lambda x, it=it: (x + it) % m, m)
        wrong += int(K.decrypt_message(out[None, :], m)[0] != (msgs[it % 16] + it) % m)
    assert wrong == 0
    # explicit tables: 200 generate/drop cycles never exhaust the 62 slots
    gen = T.Generator(m, e)
    for it in range(200):
        lut = gen.generate_lookup_table(lambda x: x)
        assert lut.lut_id > 0
        del lut
    # holding more than the engine has slots fails loudly, and releasing recovers
    held = []
    with pytest.raises(T.EngineError):
        for _ in range(100):
            held.append(gen.generate_lookup_table(lambda x: (x + 1) % m))
    assert 50 <= len(held) <= 62
    ref_id = held[0].lut_id
    out = e.batch_bootstrap_lut(ref_id, cts)
    assert np.array_equal(K.decrypt_message(out, m), (msgs + 1) % m)
    held.clear()
    lut = gen.generate_lookup_table(lambda x: x)
    # a key reload invalidates old ids: they are rejected, not resolved to another table
    stale = lut.lut_id
    e.load_cloud_key(ck)
    with pytest.raises(T.EngineError):
        e.batch_bootstrap_lut(stale, cts)
    lut.lut_id = -1
    e.close()


# ---- alloc -> copy -> commit (the non-root rank's key path), two engines in one process ------------
def test_commit_path_matches_oracle(eng128):
    K, ck, e0 = eng128
    cudart = C.CDLL("libcudart.so")
    e1 = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    e1.alloc_cloud_key()
    src, nbytes = e0.cloud_key_blob()
    dst, nbytes1 = e1.cloud_key_blob()
    assert nbytes == nbytes1
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert cudart.cudaMemcpy(C.c_void_p(dst), C.c_void_p(src), nbytes, 3) == 0   # device to device
    e1.commit_cloud_key(K.offset)
    r = np.random.default_rng(81)
    rng = O.Rng(81)
    count = 613                                   # > one SM round: the throughput kernel
    a = r.integers(0, 2, count).astype(bool)
    b = r.integers(0, 2, count).astype(bool)
    ops = r.integers(0, 10, count).astype(np.uint8)
    pairs = bool_pairs(K, a, b, rng)
    got = e1.batch_gate_mixed(ops, pairs)
    assert np.array_equal(got, K.batch_gate(ops, pairs))      # derived key orders rebuilt correctly
    assert np.array_equal(got, e0.batch_gate_mixed(ops, pairs))
    e1.close()


# ---- C2: the THROUGHPUT kernel at 80 / 110 bit, word for word ---------------------------------------
@pytest.mark.parametrize("name", ["80", "110"])
def test_throughput_kernel_other_gate_sets_bit_exact(name):
    K, ck = keys(name)
    P = T.PARAMS_BY_NAME[name]
    e = T.CudaBootstrap(P, 0)
    e.load_cloud_key(ck)
    r = np.random.default_rng(91)
    rng = O.Rng(91)
    count = 1813                                  # ragged: 3 full rounds + a partial one
    a = r.integers(0, 2, count).astype(bool)
    b = r.integers(0, 2, count).astype(bool)
    ops = r.integers(0, 10, count).astype(np.uint8)
    pairs = bool_pairs(K, a, b, rng)
    got = e.batch_gate_mixed(ops, pairs)
    ref = K.batch_gate(ops, pairs)
    assert np.array_equal(got, ref)
    want = np.array([GATE_FN[T.GATES[o]](x, y) for o, x, y in zip(ops, a, b)]).astype(bool)
    assert np.array_equal(K.decrypt_bool(got), want)
    e.close()


# ---- blind rotation at every advertised gadget (l, bgbit), both kernel shapes -------------------------
@pytest.mark.parametrize("name,m", [("uint1", 2), ("uint2", 4), ("uint3", 8), ("uint5", 32), ("uint7", 128)])
def test_blind_rotate_other_gadgets_p2(name, m):
    """l < 3 sets: two f64 implementations decorrelate on mask words, so parity is
    phase-level: equal decryptions and |phase_gpu - phase_oracle| <= 6 sigma of the PBS noise, for the
    latency kernel (count 5) and the throughput kernel (count 601)."""
    if name not in T.PARAMS_BY_NAME:
        pytest.skip("parameter set not defined")
    K, ck = keys(name)
    P = T.PARAMS_BY_NAME[name]
    e = T.CudaBootstrap(P, 0)
    e.load_cloud_key(ck)
    rng = O.Rng(101)
    table = [(3 * x + 1) % m for x in range(m)]
    lut_id, lut_b = e.lut_generate(table, m)
    assert np.array_equal(lut_b, O.lut_generate(table, m))
    bound = 0.25 / m                                # half of a message slot half-width
    for count in (5, 601):                          # 601 = one full round (64-thread kernel) + a tail of 9 (128-thread)
        msgs = np.arange(count) % m
        cts = K.encrypt_message(msgs, m, rng)
        got = e.batch_bootstrap_lut(lut_id, cts)
        ref = K.batch_bootstrap(cts, key_switch=True, lut_b=lut_b)
        d = torus_dist(K.phase(got), K.phase(ref))
        assert d.max() < bound, (name, count, d.max())             # same phase as the oracle, every ciphertext
        dec_g, dec_r = K.decrypt_message(got, m), K.decrypt_message(ref, m)
        # the algorithm itself mis-decodes a message now and then at the larger moduli (input modulus
        # switch noise vs slot width); GPU and oracle must agree, and both be mostly right
        assert (dec_g != dec_r).mean() <= 0.01, (name, count)
        if m <= 32:   # at m = 128 the input modulus switch alone (sigma 3.4e-3 vs slot half-width 2e-3) decodes
            assert (dec_g == (3 * msgs + 1) % m).mean() >= 0.97, (name, count)   # wrongly most of the time, oracle included
    e.close()


# ---- roofline denominators -------------------------------------------------------------------------
def test_fp64_probes_show_the_operand_path_limit(eng128):
    """B200 feeds its FP64 unit one 64-bit register operand per lane per cycle: a DFMA stream with three
    distinct register operands runs at ~2/3 of the constant-multiplicand stream (DESIGN.md section 4,
    profiles/r2_fp64_operand_probe.json).  bench.py reports both; keep the pair honest."""
    _, _, e = eng128
    const_mult = e.probe_fp64_tflops()
    three_reg = e.probe_fp64_3op_tflops()
    assert 20.0 < const_mult < 45.0
    assert 0.55 < three_reg / const_mult < 0.80, (three_reg, const_mult)
