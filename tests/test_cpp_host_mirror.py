"""The compiled (C++) host mirror include/rs_tfhe_b200.hpp: builds against the C ABI on the
CPU box (and fails loudly there without a device); on the GPU box the reference's gate /
bootstrap / LUT tests re-expressed in C++ must all pass."""
import os
import subprocess

import pytest

import oracle as O
import rs_tfhe_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def build():
    O.build()
    T.build_native()
    src = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
    deps = [src, os.path.join(ROOT, "include", "rs_tfhe_b200.hpp"), os.path.join(ROOT, "include", "tfhe_b200.h")]
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(map(os.path.getmtime, deps)):
        return
    csrc = os.path.join(ROOT, "rs_tfhe_b200", "csrc")
    orc = os.path.join(ROOT, "oracle")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", src, "-o", EXE,
                           f"-L{csrc}", "-ltfhe_b200", f"-L{orc}", "-ltfhe_oracle",
                           f"-Wl,-rpath,{csrc}", f"-Wl,-rpath,{orc}", "-fopenmp"])


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu():
    build()
    if T.device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 3 and "fails loudly" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_gpu():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
