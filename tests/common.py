"""Shared helpers for the parity tests: seeded keys from the oracle wrapped as the
product's CloudKey, so oracle and engine see identical key / ciphertext blobs."""
import functools

import numpy as np

import oracle as O
import rs_tfhe_b200 as T


@functools.lru_cache(maxsize=4)
def keys(name: str, seed: int = 0x5EED0001, with_torus_bsk: bool = False):
    K = O.Keys(name, seed=seed, with_torus_bsk=with_torus_bsk)
    ck = T.CloudKey(T.PARAMS_BY_NAME[name], K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk)
    return K, ck


def bool_pairs(K, bits_a, bits_b, rng):
    a = K.encrypt_bool(bits_a, rng)
    b = K.encrypt_bool(bits_b, rng)
    return np.stack([a, b], axis=1)


GATE_FN = {
    "NAND": lambda a, b: ~(a & b), "AND": lambda a, b: a & b, "OR": lambda a, b: a | b,
    "XOR": lambda a, b: a ^ b,
    # reference quirk kept for parity: Gates::xnor (gates.rs:86-90: a - 2b - 1/4) decrypts to
    # a XOR b, and the reference's own test pins exactly that (gates.rs:575-579: `false ^ (b ^ a)`)
    "XNOR": lambda a, b: a ^ b,
    "NOR": lambda a, b: ~(a | b),
    "ANDNY": lambda a, b: ~a & b, "ANDYN": lambda a, b: a & ~b,
    "ORNY": lambda a, b: ~a | b, "ORYN": lambda a, b: a | ~b,
}


def torus_dist(x, y):
    d = (np.asarray(x, dtype=np.int64) - np.asarray(y, dtype=np.int64) + 2**31) % 2**32 - 2**31
    return np.abs(d) / 2.0**32
