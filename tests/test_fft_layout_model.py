"""numpy model of the thread mapping used by the CUDA negacyclic transforms
(rs_tfhe_b200/csrc/fft512.cuh): three radix-8 passes over 64 threads, twist
merged into the pass-A twiddles, decimation-in-frequency forward /
decimation-in-time inverse so that no reordering pass exists and the
bootstrapping key is permuted once at upload.  The model is checked against
numpy's FFT and the oracle; the CUDA code transcribes it line by line.
"""
import numpy as np

import oracle as O

W8 = np.exp(-2j * np.pi / 8)
M8 = np.array([[W8 ** (a * b) for a in range(8)] for b in range(8)])  # M8[k, m]
PRE = np.exp(1j * np.pi * np.arange(8) / 16)  # omega^(64 m)
R = np.arange(64)
TA = np.exp(1j * np.pi * R[:, None] * (1 - 4 * np.arange(8)[None, :]) / 1024)  # [r][k0]
TB = np.exp(-2j * np.pi * np.arange(8)[:, None] * np.arange(8)[None, :] / 64)  # [j0][k1]


def forward_model(x):
    """x: int array[1024] -> D[v][k2] with v = k0*8+k1 holding F_{k0+8k1+64k2}."""
    x = np.asarray(x, dtype=np.float64)
    c = x[:512] + 1j * x[512:]
    buf = np.zeros(512, dtype=complex)
    # pass A: thread r, points 64m+r
    for r in range(64):
        a_in = c[64 * np.arange(8) + r] * PRE
        a_out = (M8 @ a_in) * TA[r]
        buf[np.arange(8) * 64 + r] = a_out
    # pass B: thread u=(k0,j0), over j1
    nxt = np.zeros(512, dtype=complex)
    for k0 in range(8):
        for j0 in range(8):
            idx = k0 * 64 + np.arange(8) * 8 + j0
            nxt[idx] = (M8 @ buf[idx]) * TB[j0]
    buf = nxt
    # pass C: thread v=(k0,k1), over j0
    D = np.zeros((64, 8), dtype=complex)
    for k0 in range(8):
        for k1 in range(8):
            idx = k0 * 64 + k1 * 8 + np.arange(8)
            D[k0 * 8 + k1] = M8 @ buf[idx]
    return D


def inverse_model(G):
    """G[v][k2] -> real array[1024] (before rounding), unnormalised."""
    M8c = M8.conj()
    buf = np.zeros(512, dtype=complex)
    for k0 in range(8):
        for k1 in range(8):
            idx = k0 * 64 + k1 * 8 + np.arange(8)
            buf[idx] = (M8c @ G[k0 * 8 + k1]) * TB[k1].conj()
    nxt = np.zeros(512, dtype=complex)
    for k0 in range(8):
        for j0 in range(8):
            idx = k0 * 64 + np.arange(8) * 8 + j0
            nxt[idx] = M8c @ buf[idx]
    buf = nxt
    out = np.zeros(1024)
    for r in range(64):
        a_in = buf[np.arange(8) * 64 + r] * TA[r].conj()
        y = (M8c @ a_in) * PRE.conj()
        out[64 * np.arange(8) + r] = y.real
        out[64 * np.arange(8) + r + 512] = y.imag
    return out


def natural_to_device(spec):
    """spec[k], k<512 natural order -> [v][k2] (the upload permutation)."""
    v = np.arange(64)
    k0, k1 = v // 8, v % 8
    k = k0[:, None] + 8 * k1[:, None] + 64 * np.arange(8)[None, :]
    return spec[k]


def test_forward_matches_twisted_dft():
    rng = np.random.default_rng(0)
    x = rng.integers(-32, 32, 1024)
    c = x[:512] + 1j * x[512:]
    z = c * np.exp(1j * np.pi * np.arange(512) / 1024)
    F = np.fft.fft(z)
    D = forward_model(x)
    assert np.abs(D - natural_to_device(F)).max() < 1e-8


def test_forward_matches_oracle_ifft():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 2**32, 1024, dtype=np.uint32)
    ref = O.ifft(x)  # 2F, re|im split (klemsa.rs:110-114)
    F = (ref[:512] + 1j * ref[512:]) / 2
    D = forward_model(x.view(np.int32))
    assert np.abs(D - natural_to_device(F)).max() < 1e-13 * np.abs(F).max()


def test_inverse_roundtrip():
    rng = np.random.default_rng(2)
    x = rng.integers(-2**31, 2**31, 1024)
    y = inverse_model(forward_model(x)) / 512
    assert np.array_equal(np.round(y).astype(np.int64), x)


def test_product_via_model_is_exact():
    """digit poly (x) torus poly through the device layout == exact negacyclic
    product, with the reference-layout BSK scaled by 1/1024 at upload."""
    rng = np.random.default_rng(3)
    for _ in range(4):
        d = rng.integers(-32, 32, 1024)
        b = rng.integers(0, 2**32, 1024, dtype=np.uint32)
        bref = O.ifft(b)
        bdev = natural_to_device((bref[:512] + 1j * bref[512:]) / 1024)
        y = inverse_model(forward_model(d) * bdev)
        got = (np.round(y).astype(np.int64) & 0xFFFFFFFF).astype(np.uint32)
        exact = O.poly_mul_exact(d.astype(np.int64).astype(np.uint32), b)
        assert np.array_equal(got, exact)
        assert np.abs(y - np.round(y)).max() < 0.05
