"""CPU emulation of the CUDA kernel's per-thread phases (rs_tfhe_b200/csrc/emu.cpp
runs the same __host__ __device__ code as blind_rotate.cu) against the oracle.
Class P1: at l=3 every accumulator word must equal the oracle's f64
path AND the exact-integer ground truth."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rs_tfhe_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(CSRC, "libtfhe_emu.so")
    srcs = [os.path.join(CSRC, f) for f in ("emu.cpp", "br_core.cuh", "brs_core.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                               srcs[0], "-o", so])
    return C.CDLL(so)


def emu_blind_rotate(emu, K, lwe, steps=-1, tv=None):
    p = K.params
    tv = np.stack([K.tv_a, K.tv_b]) if tv is None else tv
    tv = np.ascontiguousarray(tv, dtype=np.uint32)
    lwe = np.ascontiguousarray(lwe, dtype=np.uint32)
    out = np.empty((2, 1024), dtype=np.uint32)
    rc = emu.emu_blind_rotate(p.n, p.l, p.bgbit, C.c_uint32(K.offset), K.bsk.ctypes.data_as(C.c_void_p),
                              tv.ctypes.data_as(C.c_void_p), lwe.ctypes.data_as(C.c_void_p),
                              steps, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


@pytest.fixture(scope="module")
def keys128():
    return O.Keys("128", seed=0x5EED0001, with_torus_bsk=True)


def test_emulated_kernel_trajectory_128(emu, keys128):
    K = keys128
    rng = O.Rng(11)
    a = K.encrypt_bool([1], rng)[0]
    b = K.encrypt_bool([0], rng)[0]
    lwe = K.gate_prep(O.GATE_CODE["NAND"], a, b)
    for steps in (0, 1, 2, 7, 40):
        got = emu_blind_rotate(emu, K, lwe, steps)
        ra, rb, mf = K.blind_rotate(lwe, steps=steps)
        ea, eb, _ = K.blind_rotate(lwe, steps=steps, exact=True)
        assert np.array_equal(got[0], ra) and np.array_equal(got[1], rb), steps
        assert np.array_equal(got[0], ea) and np.array_equal(got[1], eb), steps
        assert mf < 0.05


def test_emulated_kernel_full_gate_128(emu, keys128):
    K = keys128
    rng = O.Rng(12)
    a = K.encrypt_bool([1], rng)[0]
    b = K.encrypt_bool([1], rng)[0]
    lwe = K.gate_prep(O.GATE_CODE["NAND"], a, b)
    got = emu_blind_rotate(emu, K, lwe)
    ra, rb, _ = K.blind_rotate(lwe)
    assert np.array_equal(got[0], ra) and np.array_equal(got[1], rb)
    out = K.identity_key_switching(O.sample_extract_index(got[0], got[1], 0))
    assert not K.decrypt_bool(out)[0]  # NAND(1,1) = 0


def test_emulated_kernel_lut_uint4(emu):
    """l=1, Bg=2^22: f64 is inexact, so parity is phase-level (class P2)."""
    K = O.Keys("uint4", seed=0x5EED0003)
    rng = O.Rng(13)
    m = 16
    lut = O.lut_generate(np.arange(m, dtype=np.uint32), m)
    tv = np.stack([np.zeros(1024, dtype=np.uint32), lut])
    for msg in (0, 5, 15):
        ct = K.encrypt_message([msg], m, rng)[0]
        got = emu_blind_rotate(emu, K, ct, tv=tv)
        ra, rb, _ = K.blind_rotate(ct, tv_b=lut)
        out = K.identity_key_switching(O.sample_extract_index(got[0], got[1], 0))
        ref = K.identity_key_switching(O.sample_extract_index(ra, rb, 0))
        assert K.decrypt_message(out, m)[0] == msg
        d = np.int64(K.phase(out)[0]) - np.int64(K.phase(ref)[0])
        d = (d + 2**31) % 2**32 - 2**31
        assert abs(d) / 2**32 < 4e-3


# ---- 128-thread kernel (brs_core.cuh / blind_rotate_s.cu), tensor-memory exchanges modelled ----------
def emu_blind_rotate_s(emu, K, lwe, steps=-1, tv=None):
    p = K.params
    tv = np.stack([K.tv_a, K.tv_b]) if tv is None else tv
    tv = np.ascontiguousarray(tv, dtype=np.uint32)
    lwe = np.ascontiguousarray(lwe, dtype=np.uint32)
    out = np.empty((2, 1024), dtype=np.uint32)
    rc = emu.emu_blind_rotate_s(p.n, p.l, p.bgbit, C.c_uint32(K.offset), K.bsk.ctypes.data_as(C.c_void_p),
                                tv.ctypes.data_as(C.c_void_p), lwe.ctypes.data_as(C.c_void_p),
                                steps, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def test_emulated_s_kernel_trajectory_128(emu, keys128):
    K = keys128
    rng = O.Rng(21)
    a = K.encrypt_bool([1], rng)[0]
    b = K.encrypt_bool([0], rng)[0]
    lwe = K.gate_prep(O.GATE_CODE["NAND"], a, b)
    for steps in (0, 1, 2, 7, 40):
        got = emu_blind_rotate_s(emu, K, lwe, steps)
        ra, rb, mf = K.blind_rotate(lwe, steps=steps)
        ea, eb, _ = K.blind_rotate(lwe, steps=steps, exact=True)
        assert np.array_equal(got[0], ra) and np.array_equal(got[1], rb), steps
        assert np.array_equal(got[0], ea) and np.array_equal(got[1], eb), steps


def test_emulated_s_kernel_full_gate_128(emu, keys128):
    K = keys128
    rng = O.Rng(22)
    a = K.encrypt_bool([0], rng)[0]
    b = K.encrypt_bool([1], rng)[0]
    lwe = K.gate_prep(O.GATE_CODE["AND"], a, b)
    got = emu_blind_rotate_s(emu, K, lwe)
    ra, rb, _ = K.blind_rotate(lwe)
    assert np.array_equal(got[0], ra) and np.array_equal(got[1], rb)


def test_emulated_s_kernel_lut_uint4(emu):
    K = O.Keys("uint4", seed=0x5EED0003)
    rng = O.Rng(23)
    m = 16
    lut = O.lut_generate(np.arange(m, dtype=np.uint32), m)
    tv = np.stack([np.zeros(1024, dtype=np.uint32), lut])
    for msg in (3, 12):
        ct = K.encrypt_message([msg], m, rng)[0]
        got = emu_blind_rotate_s(emu, K, ct, tv=tv)
        ra, rb, _ = K.blind_rotate(ct, tv_b=lut)
        out = K.identity_key_switching(O.sample_extract_index(got[0], got[1], 0))
        ref = K.identity_key_switching(O.sample_extract_index(ra, rb, 0))
        assert K.decrypt_message(out, m)[0] == msg
        d = np.int64(K.phase(out)[0]) - np.int64(K.phase(ref)[0])
        d = (d + 2**31) % 2**32 - 2**31
        assert abs(d) / 2**32 < 4e-3
