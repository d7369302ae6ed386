"""Client-side operations of rs-tfhe that stay on the host (numpy): secret keys, LWE encryption
and decryption.  They are not on the accelerated path -- in the Rust crate they remain what they
are -- but a Python caller needs them to drive the engine end to end (see examples/).

  key::SecretKey::new                 src/key.rs:33-48
  TLWELv0::encrypt_f64 / encrypt_bool src/tlwe.rs:37-58   (+ utils::gaussian_f64, utils.rs:22-38)
  TLWELv0::decrypt_bool               src/tlwe.rs:60-68
  encrypt_lwe_message / decrypt_lwe_message   src/tlwe.rs:84-126
The reference draws from rand::thread_rng (an OS-seeded ChaCha generator).  Here the default
(seed=None) draws every key bit, mask word and noise sample from the operating system's CSPRNG
(os.urandom); passing a seed switches to numpy's PCG64, which is reproducible and NOT
cryptographically secure -- for tests and benches only (the mask words of a ciphertext are public,
and a non-cryptographic generator's state can be recovered from them).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import N, SECURITY_128_BIT, SecurityParams


def _f64_to_torus(d: np.ndarray) -> np.ndarray:
    """utils::f64_to_torus (src/utils.rs:9-12), vectorised: fmod, scale, truncate, wrap."""
    t = np.fmod(np.asarray(d, dtype=np.float64), 1.0) * 4294967296.0
    return np.trunc(t).astype(np.int64).astype(np.uint32)


class _OsRng:
    """The subset of numpy's Generator interface used below, backed by os.urandom."""

    def integers(self, low, high, size, dtype=np.uint32):
        shape = (size,) if np.isscalar(size) else tuple(size)
        count = int(np.prod(shape))
        raw = np.frombuffer(os.urandom(4 * count), dtype=np.uint32).reshape(shape)
        span = int(high) - int(low)
        if span == 2**32:
            return raw.astype(dtype)
        if span & (span - 1):
            raise ValueError("power-of-two ranges only")
        return ((raw & np.uint32(span - 1)) + np.uint32(low)).astype(dtype)

    def normal(self, mean, sigma, size):
        count = int(np.prod((size,) if np.isscalar(size) else tuple(size)))
        u = np.frombuffer(os.urandom(16 * count), dtype=np.uint64).reshape(2, count)
        a = ((u[0] >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        b = ((u[1] >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        return mean + sigma * np.sqrt(-2.0 * np.log(a)) * np.cos(2.0 * np.pi * b)   # Box-Muller


def _rng(seed):
    return _OsRng() if seed is None else np.random.default_rng(seed)


@dataclass
class SecretKey:
    """key::SecretKey (src/key.rs:21-48): uniform binary level-0 and level-1 keys."""
    params: SecurityParams
    key_lv0: np.ndarray
    key_lv1: np.ndarray

    @staticmethod
    def new(params: SecurityParams = SECURITY_128_BIT, seed: int | None = None) -> "SecretKey":
        r = _rng(seed)
        return SecretKey(params, r.integers(0, 2, params.n, dtype=np.uint32),
                         r.integers(0, 2, N, dtype=np.uint32))


class Client:
    """Encrypts / decrypts batches under one SecretKey."""

    def __init__(self, sk: SecretKey, seed: int | None = None):
        self.sk = sk
        self.rng = _rng(seed)

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
