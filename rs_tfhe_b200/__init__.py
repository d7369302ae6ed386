"""rs_tfhe_b200 -- host-side mirror of rs-tfhe's gate / bootstrap API over the
B200 engine's C ABI (include/tfhe_b200.h, built into csrc/libtfhe_b200.so).

The names follow the reference (thedonutfactory/rs-tfhe) so parity tests read
like its own:  params (SECURITY_128_BIT ...), key::CloudKey, gates::Gates and the
batch_* free functions, bootstrap::{Bootstrap, vanilla, lut::LutBootstrap},
lut::{Generator, LookupTable, Encoder}.  Ciphertexts are numpy uint32 arrays in
the reference's memory image (TLWELv0 = u32[n+1]).

There is no CPU fallback: constructing an engine without the CUDA library or
without a GPU raises EngineError.  Key generation and encryption are client-side
operations of the reference and are not part of this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("TFHE_B200_LIB") or os.path.join(_CSRC, "libtfhe_b200.so")
N = 1024

__all__ = [
    "EngineError", "SecurityParams", "SECURITY_80_BIT", "SECURITY_110_BIT", "SECURITY_128_BIT",
    "SECURITY_UINT1", "SECURITY_UINT2", "SECURITY_UINT3", "SECURITY_UINT4", "SECURITY_UINT5",
    "SECURITY_UINT6", "SECURITY_UINT7", "SECURITY_UINT8", "PARAMS_BY_NAME", "CloudKey",
    "CudaBootstrap", "default_bootstrap", "Gates", "LutBootstrap", "Generator", "LookupTable",
    "Encoder", "GATES", "f64_to_torus", "build_native",
    "batch_nand", "batch_and", "batch_or", "batch_xor", "batch_nor", "batch_xnor", "batch_gate",
    "batch_gate_mixed", "batch_blind_rotate", "nand", "and_", "or_", "xor", "xnor", "nor",
    "and_ny", "and_yn", "or_ny", "or_yn", "mux", "not_", "copy", "constant",
]


class EngineError(RuntimeError):
    """A C-ABI call failed (message from tfhe_last_error)."""


# --------------------------------------------------------------------------- params
@dataclass(frozen=True)
class SecurityParams:
    """Runtime image of params::SecurityParams (src/params.rs:53-84)."""
    name: str
    n: int          # tlwe_lv0.n
    N: int          # trgsw_lv1.n
    l: int
    bgbit: int
    basebit: int
    iks_t: int
    alpha_lv0: float
    alpha_lv1: float

    @property
    def bg(self) -> int:
        return 1 << self.bgbit

    @property
    def ksk_rows(self) -> int:
        return self.N * self.iks_t * (1 << self.basebit)


# src/params.rs:91-404
SECURITY_80_BIT = SecurityParams("80", 550, 1024, 3, 6, 2, 7, 5.0e-5, 3.73e-8)
SECURITY_110_BIT = SecurityParams("110", 630, 1024, 3, 6, 2, 8, 3.0517578125e-05, 2.9802322387695313e-8)
SECURITY_128_BIT = SecurityParams("128", 700, 1024, 3, 6, 2, 9, 2.0e-5, 2.0e-8)
SECURITY_UINT1 = SecurityParams("uint1", 700, 1024, 2, 10, 2, 8, 2.0e-05, 2.0e-08)
SECURITY_UINT2 = SecurityParams("uint2", 687, 1024, 1, 18, 4, 3, 0.00002120846893069972, 0.0000000000023184122752704995)
SECURITY_UINT3 = SecurityParams("uint3", 820, 1024, 1, 23, 6, 2, 0.0000025167616095979554, 0.0000000000000002220446049250313)
SECURITY_UINT4 = SecurityParams("uint4", 820, 1024, 1, 22, 5, 3, 0.0000025167616095979554, 0.0000000000000002220446049250313)
SECURITY_UINT5 = SecurityParams("uint5", 1071, 1024, 1, 22, 6, 3, 7.08822676541043e-8, 2.2204460492503131e-17)
SECURITY_UINT6 = SecurityParams("uint6", 1071, 1024, 1, 22, 6, 3, 7.08822676541043e-8, 2.2204460492503131e-17)
SECURITY_UINT7 = SecurityParams("uint7", 1160, 1024, 1, 22, 7, 3, 1.9662200074984027e-8, 2.2204460492503131e-17)
SECURITY_UINT8 = SecurityParams("uint8", 1160, 1024, 1, 22, 7, 3, 1.9662200074984027e-8, 2.2204460492503131e-17)
PARAMS_BY_NAME = {p.name: p for p in (
    SECURITY_80_BIT, SECURITY_110_BIT, SECURITY_128_BIT, SECURITY_UINT1, SECURITY_UINT2,
    SECURITY_UINT3, SECURITY_UINT4, SECURITY_UINT5, SECURITY_UINT6, SECURITY_UINT7, SECURITY_UINT8)}

# enum tfhe_gate (include/tfhe_b200.h); order shared with the oracle
GATES = ["NAND", "AND", "OR", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]
_GATE_CODE = {g: i for i, g in enumerate(GATES)}


def f64_to_torus(d: float) -> int:
    """utils::f64_to_torus (src/utils.rs:9-12)."""
    import math
    return int(math.fmod(d, 1.0) * 4294967296.0) & 0xFFFFFFFF


# --------------------------------------------------------------------------- native lib
class _CParams(C.Structure):
    _fields_ = [("n", C.c_uint32), ("N", C.c_uint32), ("l", C.c_uint32), ("bgbit", C.c_uint32),
                ("basebit", C.c_uint32), ("iks_t", C.c_uint32)]


_lib = None


def build_native(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libtfhe_b200.so (nvcc required)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "tfhe_b200.h"))
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs))
    if force or stale:
        subprocess.check_call(["make", "-C", _CSRC, "libtfhe_b200.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). rs_tfhe_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32p = C.c_void_p, C.c_void_p
    L.tfhe_abi_version.restype = C.c_int
    L.tfhe_last_error.restype = C.c_char_p
    L.tfhe_device_count.restype = C.c_int
    L.tfhe_engine_create.argtypes = [C.POINTER(_CParams), C.c_int, C.POINTER(vp)]
    L.tfhe_engine_create_multi.argtypes = [C.POINTER(_CParams), C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.tfhe_engine_device_count.argtypes = [vp]
    L.tfhe_engine_last_broadcast_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tfhe_engine_destroy.argtypes = [vp]
    L.tfhe_engine_destroy.restype = None
    L.tfhe_engine_set_stream.argtypes = [vp, vp]
    L.tfhe_engine_kernel_launches.argtypes = [vp]
    L.tfhe_engine_kernel_launches.restype = C.c_uint64
    L.tfhe_engine_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tfhe_engine_load_cloud_key.argtypes = [vp, C.c_uint32, u32p, u32p, u32p, vp]
    L.tfhe_engine_alloc_cloud_key.argtypes = [vp]
    L.tfhe_engine_cloud_key_blob.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.tfhe_engine_commit_cloud_key.argtypes = [vp, C.c_uint32]
    L.tfhe_batch_gate.argtypes = [vp, C.c_int, u32p, u32p, C.c_size_t]
    L.tfhe_batch_gate_mixed.argtypes = [vp, vp, u32p, u32p, C.c_size_t]
    L.tfhe_batch_bootstrap.argtypes = [vp, u32p, u32p, C.c_size_t, C.c_int]
    L.tfhe_batch_blind_rotate.argtypes = [vp, u32p, u32p, C.c_size_t]
    L.tfhe_lut_generate.argtypes = [vp, u32p, C.c_uint32, C.c_double, u32p, C.POINTER(C.c_int)]
    L.tfhe_lut_register.argtypes = [vp, u32p, u32p, C.POINTER(C.c_int)]
    L.tfhe_batch_ifft.argtypes = [vp, u32p, vp, C.c_size_t]
    L.tfhe_batch_fft.argtypes = [vp, vp, u32p, C.c_size_t]
    L.tfhe_batch_poly_mul.argtypes = [vp, u32p, u32p, u32p, C.c_size_t]
    L.tfhe_circuit_create.argtypes = [vp, C.POINTER(vp)]
    L.tfhe_circuit_destroy.argtypes = [vp]
    L.tfhe_circuit_destroy.restype = None
    u32o = C.POINTER(C.c_uint32)
    L.tfhe_circuit_input.argtypes = [vp, u32o]
    L.tfhe_circuit_constant.argtypes = [vp, C.c_int, u32o]
    L.tfhe_circuit_not.argtypes = [vp, C.c_uint32, u32o]
    L.tfhe_circuit_gate.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, u32o]
    L.tfhe_circuit_mux.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, u32o]
    L.tfhe_circuit_output.argtypes = [vp, C.c_uint32]
    L.tfhe_circuit_stats.argtypes = [vp, u32o, u32o, u32o]
    L.tfhe_circuit_run.argtypes = [vp, vp, vp, C.c_size_t]
    L.tfhe_lut_release.argtypes = [vp, C.c_int]
    L.tfhe_batch_bootstrap_func.argtypes = [vp, u32p, C.c_uint32, C.c_double, u32p, u32p, C.c_size_t]
    L.tfhe_batch_bootstrap_lut.argtypes = [vp, C.c_int, u32p, u32p, C.c_size_t]
    L.tfhe_batch_extract_key_switch.argtypes = [vp, u32p, u32p, C.c_size_t]
    L.tfhe_batch_bootstrap_lut_multi.argtypes = [vp, vp, u32p, u32p, C.c_size_t]
    L.tfhe_batch_gate_dev.argtypes = [vp, C.c_int, vp, vp, vp, C.c_size_t]
    L.tfhe_batch_bootstrap_dev.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, C.c_int]
    L.tfhe_engine_synchronize.argtypes = [vp]
    L.tfhe_reenc_key_load.argtypes = [vp, u32p, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.tfhe_reenc_key_destroy.argtypes = [vp]
    L.tfhe_reenc_key_destroy.restype = None
    L.tfhe_batch_reencrypt.argtypes = [vp, vp, u32p, u32p, C.c_size_t]
    L.tfhe_engine_cloud_key_export_bytes.argtypes = [vp]
    L.tfhe_engine_cloud_key_export_bytes.restype = C.c_size_t
    L.tfhe_engine_export_cloud_key.argtypes = [vp, vp, C.c_size_t]
    L.tfhe_engine_import_cloud_key.argtypes = [vp, vp, C.c_size_t]
    L.tfhe_engine_generate_cloud_key.argtypes = [vp, u32p, u32p, C.c_double, C.c_double, C.c_uint64]
    L.tfhe_probe_fp64_tflops.argtypes = [vp, C.POINTER(C.c_double)]
    L.tfhe_probe_fp64_3op_tflops.argtypes = [vp, C.POINTER(C.c_double)]
    if L.tfhe_abi_version() != 2:
        raise EngineError("libtfhe_b200.so ABI version mismatch")
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise EngineError(f"tfhe_b200 error {rc}: {_load().tfhe_last_error().decode()}")


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(a, shape_tail: Optional[tuple] = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if shape_tail is not None and tuple(a.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError(f"expected trailing shape {shape_tail}, got {a.shape}")
    return a


def device_count() -> int:
    return _load().tfhe_device_count()


# --------------------------------------------------------------------------- key
@dataclass
class CloudKey:
    """key::CloudKey (src/key.rs:51-56) in the reference's memory layout."""
    params: SecurityParams
    decomposition_offset: int
    blind_rotate_testvec_a: np.ndarray   # u32[N]
    blind_rotate_testvec_b: np.ndarray   # u32[N]
    key_switching_key: np.ndarray        # u32[N*t*2^basebit][n+1]
    bootstrapping_key: np.ndarray        # f64[n][2l][2][N]  (TRGSWLv1FFT image)

    def validate(self) -> None:
        p = self.params
        if self.key_switching_key.shape != (p.ksk_rows, p.n + 1):
            raise ValueError("key_switching_key shape")
        if self.bootstrapping_key.shape != (p.n, 2 * p.l, 2, N):
            raise ValueError("bootstrapping_key shape")


# --------------------------------------------------------------------------- engine
class CudaBootstrap:
    """The B200 `Bootstrap` strategy (trait at src/bootstrap/mod.rs:23-38; the slot the
    reference reserves for "gpu", examples/bootstrap_strategies.rs:94-97).  One per
    (process, GPU).  Unlike the CPU strategies it keeps the cloud key resident on the
    device, so the key is bound with `load_cloud_key` instead of passed per call."""

    def __init__(self, params: SecurityParams = SECURITY_128_BIT, device=0):
        """`device`: one CUDA ordinal, or a sequence of ordinals for one engine over several GPUs
        (batches are sharded over them; the cloud key is broadcast with NCCL at load time)."""
        L = _load()
        self.params = params
        self._h = C.c_void_p()
        cp = _CParams(params.n, params.N, params.l, params.bgbit, params.basebit, params.iks_t)
        if np.ndim(device) == 0:
            _check(L.tfhe_engine_create(C.byref(cp), int(device), C.byref(self._h)))
        else:
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            _check(L.tfhe_engine_create_multi(C.byref(cp), ids, len(device), C.byref(self._h)))
        self._key: Optional[CloudKey] = None
        self.device = device

    @property
    def n_devices(self) -> int:
        return int(_load().tfhe_engine_device_count(self._h))

    def last_broadcast_ms(self) -> float:
        ms = C.c_float(0.0)
        _check(_load().tfhe_engine_last_broadcast_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            _load().tfhe_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def name(self) -> str:
        return "cuda-b200"

    # ---- key management
    def load_cloud_key(self, ck: CloudKey) -> None:
        ck.validate()
        if ck.params != self.params:
            raise ValueError("cloud key parameters differ from the engine's")
        ksk = _u32(ck.key_switching_key)
        bsk = np.ascontiguousarray(ck.bootstrapping_key, dtype=np.float64)
        _check(_load().tfhe_engine_load_cloud_key(
            self._h, C.c_uint32(ck.decomposition_offset), _ptr(_u32(ck.blind_rotate_testvec_a)),
            _ptr(_u32(ck.blind_rotate_testvec_b)), _ptr(ksk), _ptr(bsk)))
        self._key = ck

    def generate_cloud_key(self, key_lv0, key_lv1, seed: int = 0) -> None:
        """key::CloudKey::new(&secret_key) (src/key.rs:59-66) on the device: KSK + Fourier BSK are
        generated straight into the device layout by a ChaCha20-based generator.  seed=0 (default)
        keys it from OS entropy; a non-zero seed gives a reproducible, INSECURE key (tests only)."""
        p = self.params
        s0, s1 = _u32(key_lv0, (p.n,)), _u32(key_lv1, (N,))
        _check(_load().tfhe_engine_generate_cloud_key(self._h, _ptr(s0), _ptr(s1), p.alpha_lv0,
                                                      p.alpha_lv1, C.c_uint64(seed)))
        self._key = None

    def alloc_cloud_key(self) -> None:
        _check(_load().tfhe_engine_alloc_cloud_key(self._h))

    def cloud_key_blob(self):
        """(device pointer, bytes) of the re-laid-out key, for NCCL broadcast."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(_load().tfhe_engine_cloud_key_blob(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def commit_cloud_key(self, decomposition_offset: int) -> None:
        _check(_load().tfhe_engine_commit_cloud_key(self._h, C.c_uint32(decomposition_offset)))

    def export_cloud_key(self) -> np.ndarray:
        """The device-resident key as one self-describing byte buffer (header + device blob)."""
        nbytes = _load().tfhe_engine_cloud_key_export_bytes(self._h)
        buf = np.empty(nbytes, dtype=np.uint8)
        _check(_load().tfhe_engine_export_cloud_key(self._h, _ptr(buf), nbytes))
        return buf

    def import_cloud_key(self, buf: np.ndarray) -> None:
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        _check(_load().tfhe_engine_import_cloud_key(self._h, _ptr(buf), buf.size))
        self._key = None

    def _bind(self, cloud_key: Optional[CloudKey]) -> None:
        if cloud_key is not None and cloud_key is not self._key:
            self.load_cloud_key(cloud_key)

    # ---- stream / counters
    def set_stream(self, cuda_stream: int) -> None:
        _check(_load().tfhe_engine_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        _check(_load().tfhe_engine_synchronize(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(_load().tfhe_engine_kernel_launches(self._h))

    def last_kernel_ms(self):
        ms = (C.c_float * 2)()
        _check(_load().tfhe_engine_last_kernel_ms(self._h, ms))
        return float(ms[0]), float(ms[1])

    def probe_fp64_tflops(self) -> float:
        """Measured DFMA rate of this GPU (roofline denominator for the FP64-bound kernel)."""
        v = C.c_double(0.0)
        _check(_load().tfhe_probe_fp64_tflops(self._h, C.byref(v)))
        return v.value

    def probe_fp64_3op_tflops(self) -> float:
        """The same with three distinct register operands per DFMA (the blind rotation's operand pattern)."""
        v = C.c_double(0.0)
        _check(_load().tfhe_probe_fp64_3op_tflops(self._h, C.byref(v)))
        return v.value

    # ---- Bootstrap trait (batched: accepts [n+1] or [count][n+1])
    def bootstrap(self, ctxt, cloud_key: Optional[CloudKey] = None):
        return self._bootstrap(ctxt, cloud_key, True)

    def bootstrap_without_key_switch(self, ctxt, cloud_key: Optional[CloudKey] = None):
        return self._bootstrap(ctxt, cloud_key, False)

    def _bootstrap(self, ctxt, cloud_key, key_switch: bool):
        self._bind(cloud_key)
        w = self.params.n + 1
        cts = _u32(ctxt, (w,))
        single = cts.ndim == 1
        cts = cts.reshape(-1, w)
        out = np.empty_like(cts)
        _check(_load().tfhe_batch_bootstrap(self._h, _ptr(cts), _ptr(out), cts.shape[0], int(key_switch)))
        return out[0] if single else out

    # ---- batch entry points
    def batch_gate(self, op, inputs, cloud_key: Optional[CloudKey] = None) -> np.ndarray:
        """gates::batch_<op> (src/gates.rs:352-547). inputs: [count][2][n+1] (= &[(Ct, Ct)])."""
        self._bind(cloud_key)
        w = self.params.n + 1
        pairs = _u32(inputs, (2, w)).reshape(-1, 2, w)
        out = np.empty((pairs.shape[0], w), dtype=np.uint32)
        code = _GATE_CODE[op] if isinstance(op, str) else int(op)
        _check(_load().tfhe_batch_gate(self._h, code, _ptr(pairs), _ptr(out), pairs.shape[0]))
        return out

    def batch_gate_mixed(self, ops, inputs, cloud_key: Optional[CloudKey] = None) -> np.ndarray:
        self._bind(cloud_key)
        w = self.params.n + 1
        pairs = _u32(inputs, (2, w)).reshape(-1, 2, w)
        ops = np.ascontiguousarray(ops, dtype=np.uint8)
        if ops.shape != (pairs.shape[0],):
            raise ValueError("ops must have one entry per input pair")
        out = np.empty((pairs.shape[0], w), dtype=np.uint32)
        _check(_load().tfhe_batch_gate_mixed(self._h, _ptr(ops), _ptr(pairs), _ptr(out), pairs.shape[0]))
        return out

    def batch_blind_rotate(self, srcs, cloud_key: Optional[CloudKey] = None) -> np.ndarray:
        """trgsw::batch_blind_rotate (src/trgsw.rs:289-305) -> [count][2][N]."""
        self._bind(cloud_key)
        w = self.params.n + 1
        cts = _u32(srcs, (w,)).reshape(-1, w)
        out = np.empty((cts.shape[0], 2, N), dtype=np.uint32)
        _check(_load().tfhe_batch_blind_rotate(self._h, _ptr(cts), _ptr(out), cts.shape[0]))
        return out

    def batch_extract_key_switch(self, trlwes, cloud_key: Optional[CloudKey] = None) -> np.ndarray:
        """sample_extract_index(.,0) + identity_key_switching (trlwe.rs:106, trgsw.rs:332)."""
        self._bind(cloud_key)
        t = _u32(trlwes, (2, N)).reshape(-1, 2, N)
        out = np.empty((t.shape[0], self.params.n + 1), dtype=np.uint32)
        _check(_load().tfhe_batch_extract_key_switch(self._h, _ptr(t), _ptr(out), t.shape[0]))
        return out

    # ---- FFTProcessor seam (src/fft/mod.rs:80-107)
    def batch_ifft(self, polys) -> np.ndarray:
        """FFTProcessor::batch_ifft::<1024>: u32[count][N] -> f64[count][N] (re | im, 2 x DFT)."""
        x = _u32(polys, (N,)).reshape(-1, N)
        out = np.empty((x.shape[0], N), dtype=np.float64)
        _check(_load().tfhe_batch_ifft(self._h, _ptr(x), _ptr(out), x.shape[0]))
        return out

    def batch_fft(self, spectra) -> np.ndarray:
        """FFTProcessor::batch_fft::<1024>: f64[count][N] -> u32[count][N]."""
        f = np.ascontiguousarray(spectra, dtype=np.float64).reshape(-1, N)
        out = np.empty((f.shape[0], N), dtype=np.uint32)
        _check(_load().tfhe_batch_fft(self._h, _ptr(f), _ptr(out), f.shape[0]))
        return out

    def batch_poly_mul(self, a, b) -> np.ndarray:
        """FFTProcessor::poly_mul::<1024> over a batch: a*b mod X^N+1."""
        a = _u32(a, (N,)).reshape(-1, N)
        b = _u32(b, (N,)).reshape(-1, N)
        if a.shape != b.shape:
            raise ValueError("a and b must have the same shape")
        out = np.empty_like(a)
        _check(_load().tfhe_batch_poly_mul(self._h, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out

    # ---- LUT
    def lut_generate(self, f_table: Sequence[int], modulus: int, scale: float = 0.0):
        f = _u32(list(f_table))
        if f.shape != (modulus,):
            raise ValueError("f_table must have `modulus` entries")
        b = np.empty(N, dtype=np.uint32)
        lut_id = C.c_int(-1)
        _check(_load().tfhe_lut_generate(self._h, _ptr(f), modulus, scale, _ptr(b), C.byref(lut_id)))
        return lut_id.value, b

    def lut_register(self, poly_b, poly_a=None) -> int:
        lut_id = C.c_int(-1)
        a = None if poly_a is None else _u32(poly_a, (N,))
        _check(_load().tfhe_lut_register(self._h, _ptr(a), _ptr(_u32(poly_b, (N,))), C.byref(lut_id)))
        return lut_id.value

    def lut_release(self, lut_id: int) -> None:
        """Give a table's device slot back (dropping a LookupTable)."""
        _check(_load().tfhe_lut_release(self._h, int(lut_id)))

    def batch_bootstrap_func(self, f_table: Sequence[int], modulus: int, ctxt, scale: float = 0.0,
                             cloud_key: Optional[CloudKey] = None):
        """LutBootstrap::bootstrap_func (bootstrap/lut.rs:49-65) over a batch in one call: the
        table is generated into a scratch slot, so no table id is held and the call can be
        repeated without bound."""
        self._bind(cloud_key)
        f = _u32(list(f_table))
        if f.shape != (modulus,):
            raise ValueError("f_table must have `modulus` entries")
        w = self.params.n + 1
        cts = _u32(ctxt, (w,))
        single = cts.ndim == 1
        cts = cts.reshape(-1, w)
        out = np.empty_like(cts)
        _check(_load().tfhe_batch_bootstrap_func(self._h, _ptr(f), modulus, scale, _ptr(cts), _ptr(out),
                                                 cts.shape[0]))
        return out[0] if single else out

    def batch_bootstrap_lut(self, lut_id, ctxt, cloud_key: Optional[CloudKey] = None):
        """LutBootstrap::bootstrap_lut over a batch; `lut_id` may be one id or one id per
        ciphertext (independent tables of one circuit level in a single batch)."""
        self._bind(cloud_key)
        w = self.params.n + 1
        cts = _u32(ctxt, (w,))
        single = cts.ndim == 1
        cts = cts.reshape(-1, w)
        out = np.empty_like(cts)
        if np.ndim(lut_id) == 0:
            _check(_load().tfhe_batch_bootstrap_lut(self._h, int(lut_id), _ptr(cts), _ptr(out), cts.shape[0]))
        else:
            ids = np.ascontiguousarray(lut_id, dtype=np.int32)
            if ids.shape != (cts.shape[0],):
                raise ValueError("need one lut id per ciphertext")
            _check(_load().tfhe_batch_bootstrap_lut_multi(self._h, _ptr(ids), _ptr(cts), _ptr(out), cts.shape[0]))
        return out[0] if single else out

    # ---- proxy re-encryption (src/proxy_reenc.rs), SURVEY 8(f3)
    def load_reenc_key(self, key_encryptions, base: int, t: int) -> "ReencKey":
        """Upload a proxy_reenc::ProxyReencryptionKey{key_encryptions, base, t} (:224-233)."""
        p = self.params
        k = _u32(key_encryptions, (p.n + 1,)).reshape(-1, p.n + 1)
        if k.shape[0] != base * t * p.n:
            raise ValueError("key_encryptions must have base*t*n rows")
        h = C.c_void_p()
        _check(_load().tfhe_reenc_key_load(self._h, _ptr(k), base, t, C.byref(h)))
        return ReencKey(h, base, t)

    def batch_reencrypt(self, key: "ReencKey", cts) -> np.ndarray:
        """proxy_reenc::reencrypt_tlwe_lv0 (src/proxy_reenc.rs:468-511) over a batch."""
        w = self.params.n + 1
        c = _u32(cts, (w,))
        single = c.ndim == 1
        c = c.reshape(-1, w)
        out = np.empty_like(c)
        _check(_load().tfhe_batch_reencrypt(self._h, key._h, _ptr(c), _ptr(out), c.shape[0]))
        return out[0] if single else out

    # ---- device-resident (raw CUDA pointers; asynchronous on the engine stream)
    def batch_gate_dev(self, op, d_in_pairs: int, d_out: int, count: int, d_ops: int = 0) -> None:
        code = _GATE_CODE[op] if isinstance(op, str) else int(op)
        _check(_load().tfhe_batch_gate_dev(self._h, code, C.c_void_p(d_ops or None),
                                           C.c_void_p(d_in_pairs), C.c_void_p(d_out), count))

    def batch_bootstrap_dev(self, d_in: int, d_out: int, count: int, lut_id: int = -1,
                            key_switch: bool = True) -> None:
        _check(_load().tfhe_batch_bootstrap_dev(self._h, lut_id, C.c_void_p(d_in), C.c_void_p(d_out),
                                                count, int(key_switch)))


class ReencKey:
    """Device-resident proxy re-encryption key."""

    def __init__(self, handle, base: int, t: int):
        self._h, self.base, self.t = handle, base, t

    def close(self) -> None:
        if self._h and self._h.value:
            _load().tfhe_reenc_key_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: dict = {}


def default_bootstrap(params: SecurityParams = SECURITY_128_BIT, device: int = 0) -> CudaBootstrap:
    """bootstrap::default_bootstrap (src/bootstrap/mod.rs:41-43): a cached engine per
    (params, device) so the free functions below do not re-upload keys per call."""
    key = (params.name, device)
    if key not in _default:
        _default[key] = CudaBootstrap(params, device)
    return _default[key]


# --------------------------------------------------------------------------- gates
def _neg(ct):
    return (np.uint32(0) - _u32(ct)).astype(np.uint32)


class Gates:
    """gates::Gates (src/gates.rs:30-218)."""

    def __init__(self, bootstrap: Optional[CudaBootstrap] = None):
        self.bootstrap = bootstrap or default_bootstrap()

    @staticmethod
    def with_bootstrap(bootstrap: CudaBootstrap) -> "Gates":
        return Gates(bootstrap)

    def bootstrap_strategy(self) -> str:
        return self.bootstrap.name()

    def _gate(self, op, a, b, ck):
        w = self.bootstrap.params.n + 1
        a, b = _u32(a, (w,)), _u32(b, (w,))
        single = a.ndim == 1
        pairs = np.stack([a.reshape(-1, w), b.reshape(-1, w)], axis=1)
        out = self.bootstrap.batch_gate(op, pairs, ck)
        return out[0] if single else out

    def nand(self, a, b, cloud_key=None): return self._gate("NAND", a, b, cloud_key)
    def or_(self, a, b, cloud_key=None): return self._gate("OR", a, b, cloud_key)
    def and_(self, a, b, cloud_key=None): return self._gate("AND", a, b, cloud_key)
    def xor(self, a, b, cloud_key=None): return self._gate("XOR", a, b, cloud_key)
    def xnor(self, a, b, cloud_key=None): return self._gate("XNOR", a, b, cloud_key)
    def nor(self, a, b, cloud_key=None): return self._gate("NOR", a, b, cloud_key)
    def and_ny(self, a, b, cloud_key=None): return self._gate("ANDNY", a, b, cloud_key)
    def and_yn(self, a, b, cloud_key=None): return self._gate("ANDYN", a, b, cloud_key)
    def or_ny(self, a, b, cloud_key=None): return self._gate("ORNY", a, b, cloud_key)
    def or_yn(self, a, b, cloud_key=None): return self._gate("ORYN", a, b, cloud_key)

    def mux_naive(self, a, b, c, cloud_key=None):
        """gates.rs:189-199: AND(a,b), AND(NOT a, c), OR -- the two ANDs share one batch."""
        w = self.bootstrap.params.n + 1
        a, b, c = _u32(a, (w,)), _u32(b, (w,)), _u32(c, (w,))
        pairs = np.stack([np.stack([a, b]), np.stack([_neg(a), c])])
        u = self.bootstrap.batch_gate("AND", pairs, cloud_key)
        return self.or_(u[0], u[1], cloud_key)

    def mux(self, a, b, c, cloud_key=None):
        """gates.rs:157-183, data-flow exact (incl. sample_extract_index_2; see SURVEY 0.9:
        the reference's optimised mux is not cryptographically sound -- use mux_naive)."""
        w = self.bootstrap.params.n + 1
        a, b, c = _u32(a, (w,)), _u32(b, (w,)), _u32(c, (w,))
        t_and = (a + b).astype(np.uint32)
        t_and[-1] = np.uint32((int(t_and[-1]) + f64_to_torus(-0.125)) & 0xFFFFFFFF)
        t_ny = (_neg(a) + c).astype(np.uint32)
        t_ny[-1] = np.uint32((int(t_ny[-1]) + f64_to_torus(-0.125)) & 0xFFFFFFFF)
        u = self.bootstrap.bootstrap_without_key_switch(np.stack([t_and, t_ny]), cloud_key)
        t_or = (u[0] + u[1]).astype(np.uint32)
        t_or[-1] = np.uint32((int(t_or[-1]) + f64_to_torus(0.125)) & 0xFFFFFFFF)
        return self.bootstrap.bootstrap(t_or, cloud_key)

    def not_(self, a): return _neg(a)                     # gates.rs:202-204
    def copy(self, a): return _u32(a).copy()              # gates.rs:207-209

    def constant(self, value: bool):                       # gates.rs:212-218 (release-mode wrap)
        mu = f64_to_torus(0.125)
        mu = mu if value else (1 - mu) & 0xFFFFFFFF
        res = np.zeros(self.bootstrap.params.n + 1, dtype=np.uint32)
        res[-1] = mu
        return res


# free functions (src/gates.rs:233-326, 352-547) on the default strategy
def _g(cloud_key: CloudKey) -> Gates:
    return Gates(default_bootstrap(cloud_key.params))


def nand(a, b, cloud_key): return _g(cloud_key).nand(a, b, cloud_key)
def and_(a, b, cloud_key): return _g(cloud_key).and_(a, b, cloud_key)
def or_(a, b, cloud_key): return _g(cloud_key).or_(a, b, cloud_key)
def xor(a, b, cloud_key): return _g(cloud_key).xor(a, b, cloud_key)
def xnor(a, b, cloud_key): return _g(cloud_key).xnor(a, b, cloud_key)
def nor(a, b, cloud_key): return _g(cloud_key).nor(a, b, cloud_key)
def and_ny(a, b, cloud_key): return _g(cloud_key).and_ny(a, b, cloud_key)
def and_yn(a, b, cloud_key): return _g(cloud_key).and_yn(a, b, cloud_key)
def or_ny(a, b, cloud_key): return _g(cloud_key).or_ny(a, b, cloud_key)
def or_yn(a, b, cloud_key): return _g(cloud_key).or_yn(a, b, cloud_key)
def mux(a, b, c, cloud_key): return _g(cloud_key).mux(a, b, c, cloud_key)
def not_(a): return _neg(a)
def copy(a): return _u32(a).copy()
def constant(value: bool, params: SecurityParams = SECURITY_128_BIT):
    """gates::constant (src/gates.rs:322-326)."""
    mu = f64_to_torus(0.125)
    res = np.zeros(params.n + 1, dtype=np.uint32)
    res[-1] = mu if value else (1 - mu) & 0xFFFFFFFF
    return res


def batch_gate(op, inputs, cloud_key: CloudKey):
    return default_bootstrap(cloud_key.params).batch_gate(op, inputs, cloud_key)


def batch_gate_mixed(ops, inputs, cloud_key: CloudKey):
    return default_bootstrap(cloud_key.params).batch_gate_mixed(ops, inputs, cloud_key)


def batch_nand(inputs, cloud_key): return batch_gate("NAND", inputs, cloud_key)   # gates.rs:352
def batch_and(inputs, cloud_key): return batch_gate("AND", inputs, cloud_key)     # gates.rs:388
def batch_or(inputs, cloud_key): return batch_gate("OR", inputs, cloud_key)       # gates.rs:420
def batch_xor(inputs, cloud_key): return batch_gate("XOR", inputs, cloud_key)     # gates.rs:452
def batch_nor(inputs, cloud_key): return batch_gate("NOR", inputs, cloud_key)     # gates.rs:484
def batch_xnor(inputs, cloud_key): return batch_gate("XNOR", inputs, cloud_key)   # gates.rs:516


def batch_blind_rotate(srcs, cloud_key: CloudKey):                                  # trgsw.rs:289
    return default_bootstrap(cloud_key.params).batch_blind_rotate(srcs, cloud_key)


# --------------------------------------------------------------------------- LUT
class Encoder:
    """lut::Encoder (src/lut/encoder.rs:14-110)."""

    def __init__(self, message_modulus: int, scale: Optional[float] = None):
        self.message_modulus = message_modulus
        self.scale = 1.0 / (2.0 * message_modulus) if scale is None else scale

    @staticmethod
    def with_scale(message_modulus: int, scale: float) -> "Encoder":
        return Encoder(message_modulus, scale)

    def encode(self, message: int) -> int:
        return f64_to_torus((message % self.message_modulus) * self.scale)

    def decode(self, value: int) -> int:
        f = (int(value) & 0xFFFFFFFF) / 4294967296.0
        return int(f / self.scale + 0.5) % self.message_modulus

    def decode_bool(self, value: int) -> bool:
        return self.decode(value) != 0


@dataclass
class LookupTable:
    """lut::LookupTable (src/lut/lookup_table.rs:16-19): poly.a == 0, poly.b = table."""
    poly_b: np.ndarray
    lut_id: int
    engine: CudaBootstrap

    def is_empty(self) -> bool:
        return not self.poly_b.any()

    def release(self) -> None:
        """Free the device slot (the reference's LookupTable is dropped by scope)."""
        if self.lut_id > 0:
            self.engine.lut_release(self.lut_id)
            self.lut_id = -1

    def __del__(self):
        try:
            if self.lut_id > 0 and getattr(self.engine, "_h", None) and self.engine._h.value:
                self.engine.lut_release(self.lut_id)
        except Exception:
            pass


class Generator:
    """lut::Generator (src/lut/generator.rs:16-262); tables are generated on the device."""

    def __init__(self, message_modulus: int, engine: Optional[CudaBootstrap] = None,
                 scale: Optional[float] = None):
        self.encoder = Encoder(message_modulus, scale)
        self.engine = engine or default_bootstrap()
        self.poly_degree = N
        self.lookup_table_size = N

    def message_modulus(self) -> int:
        return self.encoder.message_modulus

    def generate_lookup_table(self, f: Callable[[int], int]) -> LookupTable:
        m = self.encoder.message_modulus
        table = [int(f(x)) % m for x in range(m)]   # the closure is tabulated host-side
        lut_id, b = self.engine.lut_generate(table, m, self.encoder.scale)
        return LookupTable(b, lut_id, self.engine)


class LutBootstrap:
    """bootstrap::lut::LutBootstrap (src/bootstrap/lut.rs:28-126)."""

    def __init__(self, engine: Optional[CudaBootstrap] = None):
        self.engine = engine or default_bootstrap()

    def name(self) -> str:
        return "lut"

    def bootstrap_func(self, ct_in, f: Callable[[int], int], message_modulus: int,
                       cloud_key: Optional[CloudKey] = None):
        m = message_modulus
        table = [int(f(x)) % m for x in range(m)]   # the closure is tabulated host-side
        return self.engine.batch_bootstrap_func(table, m, ct_in, 0.0, cloud_key)

    def bootstrap_lut(self, ct_in, lut: LookupTable, cloud_key: Optional[CloudKey] = None):
        return self.engine.batch_bootstrap_lut(lut.lut_id, ct_in, cloud_key)

    def bootstrap(self, ctxt, cloud_key: Optional[CloudKey] = None):
        return self.bootstrap_func(ctxt, lambda x: x, 2, cloud_key)        # lut.rs:109-112

    def bootstrap_without_key_switch(self, ctxt, cloud_key: Optional[CloudKey] = None):
        return self.bootstrap(ctxt, cloud_key)                             # lut.rs:114-121
