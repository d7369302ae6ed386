"""Host-side client helpers (rs_tfhe_b200/client.py) against the oracle's restatement of
tlwe.rs / key.rs: what one encrypts the other decrypts, and the noise level is the parameter's."""
import numpy as np

import oracle as O
import rs_tfhe_b200 as T
from rs_tfhe_b200.client import Client, SecretKey, _f64_to_torus


def test_f64_to_torus_vectorised_matches_reference():
    vals = np.array([0.125, -0.125, 0.25, -0.25, 0.5, 1 / 64, 0.0, 1.75, -1.125])
    assert list(_f64_to_torus(vals)) == [O.f64_to_torus(float(v)) for v in vals]


def test_client_and_oracle_interoperate():
    K = O.Keys.__new__(O.Keys)                      # only the secret key + params are needed
    p = O.Params.by_name("128")
    sk = SecretKey.new(T.SECURITY_128_BIT, seed=5)
    c = Client(sk, seed=6)
    bits = np.random.default_rng(7).integers(0, 2, 500).astype(bool)
    cts = c.encrypt_bool(bits)
    assert cts.shape == (500, 701) and np.array_equal(c.decrypt_bool(cts), bits)
    # oracle decrypts client ciphertexts
    lib = O.lib()
    dec = np.array([lib.orc_lwe_decrypt_bool(ct.ctypes.data, sk.key_lv0.ctypes.data, 700) for ct in np.ascontiguousarray(cts)])
    assert np.array_equal(dec.astype(bool), bits)
    # noise level: phase - ideal ~ N(0, alpha)
    ideal = np.where(bits, 0x20000000, 0xE0000000).astype(np.int64)
    err = ((c.phase(cts).astype(np.int64) - ideal + 2**31) % 2**32 - 2**31) / 2.0**32
    assert 0.7 * p.alpha_lv0 < err.std() < 1.3 * p.alpha_lv0
    msgs = np.arange(64) % 16
    sk4 = SecretKey.new(T.SECURITY_UINT4, seed=8)
    c4 = Client(sk4, seed=9)
    assert np.array_equal(c4.decrypt_lwe_message(c4.encrypt_lwe_message(msgs, 16), 16), msgs)


def test_os_entropy_rng_round_trip():
    """seed=None: key bits, masks and noise come from os.urandom; encrypt -> decrypt still round-trips
    and two keys differ."""
    import rs_tfhe_b200 as T
    from rs_tfhe_b200.client import Client, SecretKey
    sk = SecretKey.new(T.SECURITY_128_BIT)
    sk2 = SecretKey.new(T.SECURITY_128_BIT)
    assert set(np.unique(sk.key_lv0)) <= {0, 1} and not np.array_equal(sk.key_lv0, sk2.key_lv0)
    c = Client(sk)
    bits = np.array([0, 1, 1, 0, 1], dtype=bool)
    assert np.array_equal(c.decrypt_bool(c.encrypt_bool(bits)), bits)
    m = np.arange(16)
    assert np.array_equal(c.decrypt_lwe_message(c.encrypt_lwe_message(m, 16), 16), m)
    noise = c.rng.normal(0.0, 1.0, 20000)
    assert abs(noise.mean()) < 0.05 and abs(noise.std() - 1.0) < 0.05
