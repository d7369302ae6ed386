"""Operand packing of the tcgen05 key switch (rs_tfhe_b200/csrc/keyswitch_umma.cu) checked on the
CPU: emu.cpp runs the same index arithmetic (ku_layout.h) -- key relayout into canonical K-major
operand tiles, one-hot A words, stage / K-step / accumulator-half walk, byte-plane recombination --
and must reproduce the oracle's identity_key_switching (reference src/trgsw.rs:332-360) on every
word (class P0).  Oracle and emulator are test infrastructure only."""
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
    srcs = [os.path.join(CSRC, f) for f in ("emu.cpp", "br_core.cuh", "ku_layout.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                               srcs[0], "-o", so])
    lib = C.CDLL(so)
    lib.emu_ku_key_words.restype = C.c_size_t
    lib.emu_ku_key_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    return lib


@pytest.mark.parametrize("name,count", [("128", 5), ("80", 3), ("110", 3), ("uint1", 2), ("uint2", 3),
                                        ("uint4", 3), ("uint3", 2), ("uint5", 2)])
def test_umma_key_switch_model_equals_oracle(emu, name, count):
    K = O.Keys(name, seed=0x5EED0001)
    p = K.params
    assert 2 <= p.basebit <= 6
    words = emu.emu_ku_key_words(p.n, p.iks_t, p.basebit)
    # [n tiles][1024 t 2^basebit / 64 stages][4 operand tiles][240 x 32 B]
    assert words * 4 == -(-(p.n + 1) // 120) * (16 * p.iks_t << p.basebit) * 4 * 7680
    key = np.empty(words, dtype=np.uint32)
    ksk = np.ascontiguousarray(K.ksk, dtype=np.uint32)
    emu.emu_ku_build_key(ksk.ctypes.data_as(C.c_void_p), p.n, p.iks_t, p.basebit,
                         key.ctypes.data_as(C.c_void_p))
    if p.basebit == 2:
        # the k = 0 byte of every word is zero (a zero digit selects nothing, trgsw.rs:351)
        assert not np.any(key & 0xFF)
    rng = np.random.default_rng(7)
    ext = rng.integers(0, 2**32, (count, 1025), dtype=np.uint32)
    ext[0, :1024] = 0                 # all digits from PREC_OFFSET alone
    ext[1, :1024] = 0xFFFFFFFF        # carries through every digit
    out = np.zeros((count, p.n + 1), dtype=np.uint32)
    rc = emu.emu_ku_key_switch(key.ctypes.data_as(C.c_void_p), ext.ctypes.data_as(C.c_void_p),
                               C.c_size_t(count), p.n, p.iks_t, p.basebit, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    ref = np.stack([K.identity_key_switching(ext[i]) for i in range(count)])
    assert np.array_equal(out, ref)
