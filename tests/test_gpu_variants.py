"""Every selectable kernel stays bit-exact: the two throughput blind-rotation shapes (TFHE_BR_VARIANT=8:
64 threads per ciphertext with a 128-thread partial last round -- the default; 9: 128 threads per ciphertext
on every round), the latency kernels forced on (TFHE_BR_LATENCY_MAX) and off, the 2-CTA cluster kernel (default for up
to one ciphertext per SM pair) switched off (TFHE_BR_CLUSTER=0: the one-SM latency kernel), and the
key-switch kernels (TFHE_KS_VARIANT=umma -- tcgen05, the default --, rows, TFHE_KS_GENERIC=1; the split-sum
latency kernel for small batches forced off and on with TFHE_KS_SMALL_MAX) run
tools/sanitize.py -- mixed gates at 5 / 160 (/ SANITIZE_COUNT) ciphertexts, LUT bootstrap, blind rotate +
extract/key switch, each compared word for word with the oracle, plus the FFT seam and a small circuit -- in
their own process (the selectors are read once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [
    {"TFHE_BR_VARIANT": "8", "TFHE_BR_LATENCY_MAX": "0", "SANITIZE_COUNT": "601"},
    {"TFHE_BR_VARIANT": "9", "TFHE_BR_LATENCY_MAX": "0", "SANITIZE_COUNT": "601"},
    {"TFHE_BR_VARIANT": "9"}, {"TFHE_BR_LATENCY_MAX": "1000"}, {"TFHE_BR_CLUSTER": "0"},
    {"TFHE_BR_LATENCY_KERNEL": "x"},
    {"TFHE_KS_VARIANT": "umma"}, {"TFHE_KS_VARIANT": "umma", "TFHE_KS_SMALL_MAX": "0"}, {"TFHE_KS_SMALL_MAX": "1000"},
    {"TFHE_KS_VARIANT": "rows"}, {"TFHE_KS_VARIANT": "rows", "TFHE_KS_GENERIC": "1"},
], ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_variant_bit_exact(env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanitize.py")],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    checks = [l for l in r.stdout.splitlines() if l.endswith(": True") or l.endswith(": False")]
    assert len(checks) >= 7 and all(l.endswith(": True") for l in checks), r.stdout
    for must in ("gates equal (5)", "gates equal (160)", "lut equal", "ks equal", "fft round trip", "poly_mul equal",
                 "circuit equal") + (("gates equal (%s)" % env["SANITIZE_COUNT"],) if "SANITIZE_COUNT" in env else ()):
        assert any(l.startswith(must) for l in checks), (must, r.stdout)
