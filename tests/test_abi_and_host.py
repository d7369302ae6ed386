"""CPU-side checks of the drop-in boundary and host logic (no compute calls without a GPU):
the C-ABI library loads and exports every symbol include/tfhe_b200.h declares, fails loudly
without a device, the Python mirror keeps the reference's names, and the sharding helpers
work under a world_size-2 gloo group."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import rs_tfhe_b200 as T
from rs_tfhe_b200.dist import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tfhe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfhe_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    T.build_native()
    lib = C.CDLL(T.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.tfhe_abi_version() == 2


def test_library_is_sm100a_with_tma():
    out = subprocess.run(["cuobjdump", "-lelf", T.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", T.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass      # TMA bulk copy of the bootstrapping-key rows
    assert "DFMA" in sass and "USETMAXREG" in sass
    assert "UTCIMMA" in sass     # tcgen05.mma kind::i8: the key switch
    assert "LDTM" in sass and "STTM" in sass   # tcgen05.ld / .st: transform exchanges + constants in tensor memory
    assert "STAS" in sass and "UCGABAR" in sass   # st.async into the cluster peer's shared memory, cluster barrier


def test_no_gpu_fails_loudly():
    if T.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(T.EngineError, match="no CUDA device|no CPU"):
        T.CudaBootstrap(T.SECURITY_128_BIT, 0)


def test_invalid_params_rejected():
    lib = T._load()
    h = C.c_void_p()
    bad = T._CParams(700, 2048, 3, 6, 2, 9)
    assert lib.tfhe_engine_create(C.byref(bad), 0, C.byref(h)) == -1
    assert b"N must be" in lib.tfhe_last_error()
    bad = T._CParams(700, 1024, 4, 6, 2, 9)
    assert lib.tfhe_engine_create(C.byref(bad), 0, C.byref(h)) == -1
    assert lib.tfhe_engine_create(None, 0, C.byref(h)) == -1


def test_reference_surface_is_mirrored():
    for name in ["nand", "and_", "or_", "xor", "xnor", "nor", "and_ny", "and_yn", "or_ny", "or_yn",
                 "mux", "mux_naive", "not_", "copy", "constant", "with_bootstrap", "bootstrap_strategy"]:
        assert hasattr(T.Gates, name), name
    for name in ["batch_nand", "batch_and", "batch_or", "batch_xor", "batch_nor", "batch_xnor",
                 "batch_blind_rotate", "default_bootstrap"]:
        assert hasattr(T, name), name
    for name in ["bootstrap", "bootstrap_without_key_switch", "name"]:
        assert hasattr(T.CudaBootstrap, name) and hasattr(T.LutBootstrap, name)
    assert hasattr(T.LutBootstrap, "bootstrap_func") and hasattr(T.LutBootstrap, "bootstrap_lut")
    assert T.SECURITY_128_BIT.n == 700 and T.SECURITY_UINT4.bgbit == 22


def test_host_helpers_match_reference_constants():
    assert T.f64_to_torus(0.125) == 0x20000000 and T.f64_to_torus(-0.25) == 0xC0000000
    c = T.constant(False)
    assert c[-1] == 0xE0000001 and not c[:-1].any()      # gates.rs:212-218, release-mode wrap
    e = T.Encoder(4)
    assert [e.decode(e.encode(i)) for i in range(4)] == [0, 1, 2, 3]


def test_shard_range_partitions():
    for count in (0, 1, 7, 1024, 1048576):
        for world in (1, 2, 4, 8):
            spans = [shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from rs_tfhe_b200.dist import shard_range, gather_outputs
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
count = 13
lo, hi = shard_range(count, rank, 2)
local = np.arange(lo, hi, dtype=np.uint32)[:, None] * np.ones((1, 3), dtype=np.uint32)
out = gather_outputs(local, count, 2)
if rank == 0:
    assert out.shape == (13, 3) and (out[:, 0] == np.arange(13)).all()
    print("GLOO_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_gloo_world_size_2_shard_and_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0]
