"""ctypes face of the CPU oracle (oracle/tfhe_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(rs_tfhe_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtfhe_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libspqlios_ref.so")
N = 1024

GATES = ["NAND", "AND", "OR", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]
GATE_CODE = {g: i for i, g in enumerate(GATES)}


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "tfhe_oracle.c"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "libtfhe_oracle.so"], stdout=subprocess.DEVNULL)
    if force or not os.path.exists(_REF_PATH):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL,
                        stderr=subprocess.DEVNULL)


class _Params(C.Structure):
    _fields_ = [("n", C.c_uint32), ("N", C.c_uint32), ("l", C.c_uint32),
                ("bgbit", C.c_uint32), ("basebit", C.c_uint32), ("iks_t", C.c_uint32),
                ("alpha_lv0", C.c_double), ("alpha_lv1", C.c_double)]


class _Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4), ("has_spare", C.c_int), ("spare", C.c_double)]


class _CloudKey(C.Structure):
    _fields_ = [("p", _Params), ("offset", C.c_uint32), ("tv_a", C.c_void_p),
                ("tv_b", C.c_void_p), ("ksk", C.c_void_p), ("bsk_fft", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_f64_to_torus.restype = C.c_uint32
        _lib.orc_f64_to_torus.argtypes = [C.c_double]
        _lib.orc_torus_to_f64.restype = C.c_double
        _lib.orc_torus_to_f64.argtypes = [C.c_uint32]
        _lib.orc_decomposition_offset.restype = C.c_uint32
        _lib.orc_prec_offset.restype = C.c_uint32
        _lib.orc_ksk_words.restype = C.c_size_t
        _lib.orc_bsk_doubles.restype = C.c_size_t
        _lib.orc_lwe_phase.restype = C.c_uint32
        _lib.orc_lwe_decrypt_message.restype = C.c_uint32
        _lib.orc_div_round.restype = C.c_uint32
        _lib.orc_lut_encode.restype = C.c_uint32
        _lib.orc_lut_encode.argtypes = [C.c_uint32, C.c_uint32, C.c_double]
        _lib.orc_rng_normal.restype = C.c_double
        _lib.orc_rng_normal.argtypes = [C.c_void_p, C.c_double]
        _lib.orc_rng_u32.restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


@dataclass
class Params:
    name: str
    n: int
    N: int
    l: int
    bgbit: int
    basebit: int
    iks_t: int
    alpha_lv0: float
    alpha_lv1: float

    @staticmethod
    def by_name(name: str) -> "Params":
        c = _Params()
        if lib().orc_params_by_name(name.encode(), C.byref(c)) != 0:
            raise KeyError(name)
        return Params(name, c.n, c.N, c.l, c.bgbit, c.basebit, c.iks_t, c.alpha_lv0, c.alpha_lv1)

    def c(self) -> _Params:
        return _Params(self.n, self.N, self.l, self.bgbit, self.basebit, self.iks_t,
                       self.alpha_lv0, self.alpha_lv1)

    @property
    def ksk_rows(self) -> int:
        return N * self.iks_t * (1 << self.basebit)

    @property
    def decomposition_offset(self) -> int:
        c = self.c()
        return lib().orc_decomposition_offset(C.byref(c))

    @property
    def prec_offset(self) -> int:
        c = self.c()
        return lib().orc_prec_offset(C.byref(c))


class Rng:
    def __init__(self, seed: int):
        self.s = _Rng()
        lib().orc_rng_seed(C.byref(self.s), C.c_uint64(seed))

    def u32(self) -> int:
        return lib().orc_rng_u32(C.byref(self.s))

    def normal(self, sigma: float) -> float:
        return lib().orc_rng_normal(C.byref(self.s), sigma)


def f64_to_torus(d: float) -> int:
    return lib().orc_f64_to_torus(d)


def torus_to_f64(t: int) -> float:
    return lib().orc_torus_to_f64(int(t) & 0xFFFFFFFF)


def ifft(x):
    x = _u32(x)
    out = np.empty(N, dtype=np.float64)
    lib().orc_ifft(_p(x), _p(out))
    return out


def fft(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(N, dtype=np.uint32)
    lib().orc_fft(_p(x), _p(out))
    return out


def poly_mul(a, b):
    a, b = _u32(a), _u32(b)
    out = np.empty(N, dtype=np.uint32)
    lib().orc_poly_mul(_p(a), _p(b), _p(out))
    return out


def poly_mul_exact(a, b):
    a, b = _u32(a), _u32(b)
    out = np.empty(N, dtype=np.uint32)
    lib().orc_poly_mul_exact(_p(a), _p(b), _p(out))
    return out


def poly_mul_with_x_k(a, k: int):
    a = _u32(a)
    out = np.empty(N, dtype=np.uint32)
    lib().orc_poly_mul_with_x_k(_p(a), C.c_uint32(k), _p(out))
    return out


def div_round(a: int, b: int) -> int:
    return lib().orc_div_round(C.c_uint32(a), C.c_uint32(b))


def lut_encode(msg: int, modulus: int, scale: float | None = None) -> int:
    if scale is None:
        scale = 1.0 / (2.0 * modulus)
    return lib().orc_lut_encode(msg, modulus, scale)


def lut_generate(f_table, modulus: int, scale: float = 0.0):
    f = _u32(f_table)
    assert f.shape[0] == modulus
    out = np.empty(N, dtype=np.uint32)
    lib().orc_lut_generate(_p(f), C.c_uint32(modulus), C.c_double(scale), _p(out))
    return out


def sample_extract_index(a, b, k: int = 0):
    a, b = _u32(a), _u32(b)
    out = np.empty(N + 1, dtype=np.uint32)
    lib().orc_sample_extract_index(_p(a), _p(b), C.c_uint32(k), _p(out))
    return out


class Keys:
    """Seeded secret key + cloud key in the reference's memory layout
    (key.rs:51-56): ksk u32[N][t][2^basebit][n+1], bsk f64[n][2l][2][N]."""

    def __init__(self, params: Params | str, seed: int = 0x5EED0001, with_torus_bsk: bool = False):
        if isinstance(params, str):
            params = Params.by_name(params)
        self.params = params
        self.seed = seed
        L = lib()
        c = params.c()
        self._c = c
        self.s0 = np.empty(params.n, dtype=np.uint32)
        self.s1 = np.empty(N, dtype=np.uint32)
        L.orc_secret_key(C.byref(c), C.c_uint64(seed), _p(self.s0), _p(self.s1))
        self.tv_a = np.empty(N, dtype=np.uint32)
        self.tv_b = np.empty(N, dtype=np.uint32)
        L.orc_gen_testvec(_p(self.tv_a), _p(self.tv_b))
        self.offset = params.decomposition_offset
        self.ksk = np.empty((params.ksk_rows, params.n + 1), dtype=np.uint32)
        L.orc_gen_ksk(C.byref(c), _p(self.s0), _p(self.s1), C.c_uint64(seed + 1), _p(self.ksk))
        self.bsk = np.empty((params.n, 2 * params.l, 2, N), dtype=np.float64)
        self.bsk_torus = (np.empty((params.n, 2 * params.l, 2, N), dtype=np.uint32)
                          if with_torus_bsk else None)
        L.orc_gen_bsk(C.byref(c), _p(self.s0), _p(self.s1), C.c_uint64(seed + 2), _p(self.bsk),
                      _p(self.bsk_torus) if with_torus_bsk else None)

    # ---- client side
    def encrypt_bool(self, bits, rng: Rng):
        bits = np.atleast_1d(np.asarray(bits)).astype(bool)
        out = np.empty((bits.shape[0], self.params.n + 1), dtype=np.uint32)
        for i, b in enumerate(bits):
            lib().orc_lwe_encrypt_bool(C.byref(self._c), int(b), _p(self.s0), C.byref(rng.s),
                                       _p(out[i]))
        return out

    def encrypt_message(self, msgs, modulus: int, rng: Rng):
        msgs = np.atleast_1d(np.asarray(msgs))
        out = np.empty((msgs.shape[0], self.params.n + 1), dtype=np.uint32)
        for i, m in enumerate(msgs):
            lib().orc_lwe_encrypt_message(C.byref(self._c), int(m), modulus, _p(self.s0),
                                          C.byref(rng.s), _p(out[i]))
        return out

    def encrypt_f64_batch(self, mu, seed: int):
        """tlwe.rs:37-53 over a batch (OpenMP, one seeded stream per element)."""
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        out = np.empty((mu.shape[0], self.params.n + 1), dtype=np.uint32)
        lib().orc_lwe_encrypt_batch(C.byref(self._c), _p(mu), C.c_size_t(mu.shape[0]),
                                    C.c_double(self.params.alpha_lv0), _p(self.s0),
                                    C.c_uint64(seed), _p(out))
        return out

    def encrypt_bool_batch(self, bits, seed: int):
        return self.encrypt_f64_batch(np.where(np.asarray(bits).astype(bool), 0.125, -0.125), seed)

    def encrypt_message_batch(self, msgs, modulus: int, seed: int):
        m = np.asarray(msgs) % modulus
        return self.encrypt_f64_batch(m.astype(np.float64) * (1.0 / (2.0 * modulus)), seed)

    def phase_batch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint32)
        out = np.empty(cts.shape[0], dtype=np.uint32)
        lib().orc_lwe_phase_batch(_p(cts), C.c_size_t(cts.shape[0]), _p(self.s0), self.params.n, _p(out))
        return out

    def decrypt_bool_batch(self, cts):
        return self.phase_batch(cts).view(np.int32) >= 0

    def decrypt_message_batch(self, cts, modulus: int):
        f = self.phase_batch(cts).astype(np.float64) / 4294967296.0      # tlwe.rs:118-125
        return (f / (1.0 / (2.0 * modulus)) + 0.5).astype(np.int64) % modulus

    def phase(self, cts, level: int = 0):
        cts = np.atleast_2d(_u32(cts))
        key = self.s0 if level == 0 else self.s1
        n = key.shape[0]
        return np.array([lib().orc_lwe_phase(_p(np.ascontiguousarray(ct)), _p(key), n) for ct in cts],
                        dtype=np.uint32)

    def decrypt_bool(self, cts, level: int = 0):
        return self.phase(cts, level).astype(np.int32) >= 0

    def decrypt_message(self, cts, modulus: int):
        cts = np.atleast_2d(_u32(cts))
        return np.array([lib().orc_lwe_decrypt_message(_p(np.ascontiguousarray(ct)), _p(self.s0),
                                                       self.params.n, modulus) for ct in cts])

    # ---- evaluation side (the reference path)
    def _ck(self, tv_b=None):
        ck = _CloudKey(self._c, self.offset, _p(self.tv_a), _p(self.tv_b), _p(self.ksk), _p(self.bsk))
        return ck

    def gate_prep(self, op: int, a, b):
        a, b = _u32(a), _u32(b)
        out = np.empty(self.params.n + 1, dtype=np.uint32)
        lib().orc_gate_prep(C.byref(self._c), op, _p(a), _p(b), _p(out))
        return out

    def decomposition(self, a, b):
        a, b = _u32(a), _u32(b)
        out = np.empty((2 * self.params.l, N), dtype=np.uint32)
        lib().orc_decomposition(C.byref(self._c), C.c_uint32(self.offset), _p(a), _p(b), _p(out))
        return out

    def external_product(self, i: int, a, b, exact: bool = False):
        a, b = _u32(a), _u32(b)
        oa = np.empty(N, dtype=np.uint32)
        ob = np.empty(N, dtype=np.uint32)
        if exact:
            lib().orc_external_product_exact(C.byref(self._c), C.c_uint32(self.offset),
                                             _p(self.bsk_torus[i]), _p(a), _p(b), _p(oa), _p(ob))
        else:
            lib().orc_external_product(C.byref(self._c), C.c_uint32(self.offset),
                                       _p(self.bsk[i]), _p(a), _p(b), _p(oa), _p(ob))
        return oa, ob

    def blind_rotate(self, lwe, steps: int = -1, tv_b=None, exact: bool = False):
        """returns (acc_a, acc_b, max_frac)"""
        lwe = _u32(lwe)
        tva = self.tv_a if tv_b is None else np.zeros(N, dtype=np.uint32)
        tvb = self.tv_b if tv_b is None else _u32(tv_b)
        oa = np.empty(N, dtype=np.uint32)
        ob = np.empty(N, dtype=np.uint32)
        mf = C.c_double(0.0)
        lib().orc_blind_rotate(C.byref(self._c), C.c_uint32(self.offset), _p(self.bsk),
                               _p(self.bsk_torus) if exact else None, _p(tva), _p(tvb), _p(lwe),
                               C.c_int(steps), _p(oa), _p(ob), C.byref(mf))
        return oa, ob, mf.value

    def identity_key_switching(self, src):
        src = _u32(src)
        out = np.empty(self.params.n + 1, dtype=np.uint32)
        lib().orc_identity_key_switching(C.byref(self._c), _p(self.ksk), _p(src), _p(out))
        return out

    def batch_gate(self, op, in_pairs, threads: int = 0):
        """op: int or per-element uint8 array; in_pairs: [count][2][n+1]"""
        in_pairs = _u32(in_pairs)
        count = in_pairs.shape[0]
        out = np.empty((count, self.params.n + 1), dtype=np.uint32)
        ck = self._ck()
        if np.ndim(op) == 0:
            lib().orc_batch_gate(C.byref(ck), int(op), None, _p(in_pairs), _p(out),
                                 C.c_size_t(count), threads)
        else:
            ops = np.ascontiguousarray(op, dtype=np.uint8)
            lib().orc_batch_gate(C.byref(ck), 0, _p(ops), _p(in_pairs), _p(out),
                                 C.c_size_t(count), threads)
        return out

    def batch_bootstrap(self, cts, key_switch: bool = True, lut_b=None, threads: int = 0):
        cts = np.atleast_2d(_u32(cts))
        count = cts.shape[0]
        out = np.empty((count, self.params.n + 1), dtype=np.uint32)
        ck = self._ck()
        lut = _u32(lut_b) if lut_b is not None else None
        lib().orc_batch_bootstrap(C.byref(ck), _p(lut) if lut is not None else None, _p(cts),
                                  _p(out), C.c_size_t(count), int(bool(key_switch)), threads)
        return out

    def batch_blind_rotate(self, cts, threads: int = 0):
        cts = np.atleast_2d(_u32(cts))
        count = cts.shape[0]
        out = np.empty((count, 2, N), dtype=np.uint32)
        ck = self._ck()
        lib().orc_batch_blind_rotate(C.byref(ck), _p(cts), _p(out), C.c_size_t(count), threads)
        return out


def gen_reenc_key(K_from: "Keys", K_to: "Keys", seed: int, basebit: int = None, t: int = None):
    """ProxyReencryptionKey::new_symmetric (src/proxy_reenc.rs:336-392)."""
    p = K_from.params
    basebit = p.basebit if basebit is None else basebit
    t = p.iks_t if t is None else t
    out = np.empty(((1 << basebit) * t * p.n, p.n + 1), dtype=np.uint32)
    lib().orc_gen_reenc_key(C.byref(K_from._c), _p(K_from.s0), _p(K_to.s0), C.c_uint64(seed),
                            C.c_uint32(basebit), C.c_uint32(t), _p(out))
    return out


def reencrypt(params: Params, reenc_key, basebit: int, t: int, ct):
    """proxy_reenc::reencrypt_tlwe_lv0 (src/proxy_reenc.rs:468-511)."""
    ct = _u32(ct)
    key = _u32(reenc_key)
    out = np.empty(params.n + 1, dtype=np.uint32)
    c = params.c()
    lib().orc_reencrypt(C.byref(c), _p(key), C.c_uint32(basebit), C.c_uint32(t), _p(ct), _p(out))
    return out


def max_threads() -> int:
    return lib().orc_max_threads()


# ---- the reference's own (dormant) SPQLIOS FFT, compiled from /root/reference
# into oracle/_ref by oracle/Makefile.  Secondary cross-check only.
_ref = None


def spqlios_available() -> bool:
    return os.path.exists(_REF_PATH)


def spqlios_poly_mul(a, b):
    """Spqlios_poly_mul_1024 (src/fft/spqlios/spqlios-wrapper.cpp:27-40)."""
    global _ref
    if _ref is None:
        L = C.CDLL(_REF_PATH)
        L.Spqlios_new.restype = C.c_void_p
        L.Spqlios_new.argtypes = [C.c_int32]
        L.Spqlios_poly_mul_1024.argtypes = [C.c_void_p] * 4
        _ref = (L, L.Spqlios_new(N))
    L, h = _ref
    a, b = _u32(a), _u32(b)
    out = np.empty(N, dtype=np.uint32)
    L.Spqlios_poly_mul_1024(h, _p(out), _p(a), _p(b))
    return out
