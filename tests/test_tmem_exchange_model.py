"""The index/sign algebra of the TMEM-exchange blind-rotation kernel (blind_rotate_kernel_x,
rs_tfhe_b200/csrc/blind_rotate.cu) as modelled in tools/model/xchg_model.py: the forward transform
(shared-memory exchange 1, tcgen05.st 32x32b -> tcgen05.ld 16x256b + shfl.xor 16 exchange 2, signs
carried by the permuted key) must equal numpy's FFT, the inverse chain must return the input, and
pass B's 128-bit shared-memory loads must be bank-conflict free."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_model_matches_fft():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "model", "xchg_model.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    fwd = float(re.search(r"forward max err ([0-9.e+-]+)", r.stdout).group(1))
    inv = float(re.search(r"inverse max err ([0-9.e+-]+)", r.stdout).group(1))
    deg = int(re.search(r"pass B worst conflict degree (\d+)", r.stdout).group(1))
    assert fwd < 1e-10 and inv < 1e-12 and deg == 1


def test_key_permutation_is_a_signed_bijection():
    """bsk_permute_kernel's slot map (aux.cu): T -> standard slot 8 k0 + k1 is a bijection of 0..63."""
    seen = set()
    for T in range(64):
        k0 = 4 * (T >> 5) + ((T >> 2) & 3)
        k1 = 4 * ((T >> 4) & 1) + (T & 3)
        seen.add(8 * k0 + k1)
    assert seen == set(range(64))
