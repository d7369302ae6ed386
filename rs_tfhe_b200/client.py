"""Client-side operations of rs-tfhe that stay on the host (numpy): secret keys, LWE encryption
and decryption.  They are not on the accelerated path -- in the Rust crate they remain what they
are -- but a Python caller needs them to drive the engine end to end (see examples/).

  key::SecretKey::new                 src/key.rs:33-48
  TLWELv0::encrypt_f64 / encrypt_bool src/tlwe.rs:37-58   (+ utils::gaussian_f64, utils.rs:22-38)
  TLWELv0::decrypt_bool               src/tlwe.rs:60-68
  encrypt_lwe_message / decrypt_lwe_message   src/tlwe.rs:84-126
The reference draws from an unseeded thread_rng; here the generator is numpy's, seeded by the caller.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import N, SECURITY_128_BIT, SecurityParams


def _f64_to_torus(d: np.ndarray) -> np.ndarray:
    """utils::f64_to_torus (src/utils.rs:9-12), vectorised: fmod, scale, truncate, wrap."""
    t = np.fmod(np.asarray(d, dtype=np.float64), 1.0) * 4294967296.0
    return np.trunc(t).astype(np.int64).astype(np.uint32)


@dataclass
class SecretKey:
    """key::SecretKey (src/key.rs:21-48): uniform binary level-0 and level-1 keys."""
    params: SecurityParams
    key_lv0: np.ndarray
    key_lv1: np.ndarray

    @staticmethod
    def new(params: SecurityParams = SECURITY_128_BIT, seed: int | None = None) -> "SecretKey":
        r = np.random.default_rng(seed)
        return SecretKey(params, r.integers(0, 2, params.n, dtype=np.uint32),
                         r.integers(0, 2, N, dtype=np.uint32))


class Client:
    """Encrypts / decrypts batches under one SecretKey."""

    def __init__(self, sk: SecretKey, seed: int | None = None):
        self.sk = sk
        self.rng = np.random.default_rng(seed)

    def encrypt_f64(self, mu, alpha: float | None = None) -> np.ndarray:
        """tlwe.rs:37-53 over a batch: a uniform, b = <a,s> + f64_to_torus(N(0,alpha)) + f64_to_torus(mu)."""
        p = self.sk.params
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        alpha = p.alpha_lv0 if alpha is None else alpha
        a = self.rng.integers(0, 2**32, (mu.shape[0], p.n), dtype=np.uint32)
        inner = (a.astype(np.uint64) * self.sk.key_lv0.astype(np.uint64)).sum(axis=1).astype(np.uint32)
        noise = _f64_to_torus(self.rng.normal(0.0, alpha, mu.shape[0]))
        b = inner + noise + _f64_to_torus(mu)
        return np.concatenate([a, b[:, None].astype(np.uint32)], axis=1)

    def encrypt_bool(self, bits) -> np.ndarray:                                  # tlwe.rs:55-58
        return self.encrypt_f64(np.where(np.asarray(bits).astype(bool), 0.125, -0.125))

    def encrypt_lwe_message(self, messages, message_modulus: int) -> np.ndarray:  # tlwe.rs:84-100
        m = np.asarray(messages) % message_modulus
        return self.encrypt_f64(m.astype(np.float64) * (1.0 / (2.0 * message_modulus)))

    def phase(self, cts) -> np.ndarray:
        cts = np.atleast_2d(np.asarray(cts, dtype=np.uint32))
        n = self.sk.params.n
        inner = (cts[:, :n].astype(np.uint64) * self.sk.key_lv0.astype(np.uint64)).sum(axis=1).astype(np.uint32)
        return (cts[:, n] - inner).astype(np.uint32)

    def decrypt_bool(self, cts) -> np.ndarray:                                    # tlwe.rs:60-68
        return self.phase(cts).view(np.int32) >= 0

    def decrypt_lwe_message(self, cts, message_modulus: int) -> np.ndarray:        # tlwe.rs:111-126
        f = self.phase(cts).astype(np.float64) / 4294967296.0
        return (f / (1.0 / (2.0 * message_modulus)) + 0.5).astype(np.int64) % message_modulus
