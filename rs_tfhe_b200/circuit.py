"""Levelised boolean-circuit evaluation on top of the batch-gate engine.

The reference evaluates circuits gate by gate (examples/add_two_numbers.rs:11-49: a ripple-carry
adder built from `xor`, `and`, `or`).  Here a circuit is recorded once, split into levels of
mutually independent bootstrapped gates, and every level goes to the device as ONE mixed-gate
batch over all `batch` independent input sets; the wires stay resident in HBM between levels.
Recording and the adder / comparator builders live here; scheduling and execution are the library's
`tfhe_circuit_*` entry points (csrc/circuit.cuh): device-side gather, no torch.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

import ctypes as C

from . import GATES, CudaBootstrap, _check, _load

_CODE = {g: i for i, g in enumerate(GATES)}


@dataclass
class _Gate:
    op: str                  # a GATES name, or "NOT" / "COPY" / "CONST0" / "CONST1" / "INPUT" / "MUX"
    a: int = -1
    b: int = -1
    c: int = -1


@dataclass
class Circuit:
    """Records gates over integer wire ids; wire ids are returned by every method."""
    gates: List[_Gate] = field(default_factory=list)
    inputs: List[int] = field(default_factory=list)
    outputs: List[int] = field(default_factory=list)

    def _add(self, op: str, a: int = -1, b: int = -1, c: int = -1) -> int:
        self.gates.append(_Gate(op, a, b, c))
        return len(self.gates) - 1

    def input(self) -> int:
        w = self._add("INPUT")
        self.inputs.append(w)
        return w

    def output(self, w: int) -> int:
        self.outputs.append(w)
        return w

    # gates::Gates surface (src/gates.rs:54-218)
    def nand(self, a, b): return self._add("NAND", a, b)
    def and_(self, a, b): return self._add("AND", a, b)
    def or_(self, a, b): return self._add("OR", a, b)
    def xor(self, a, b): return self._add("XOR", a, b)
    def xnor(self, a, b): return self._add("XNOR", a, b)
    def nor(self, a, b): return self._add("NOR", a, b)
    def and_ny(self, a, b): return self._add("ANDNY", a, b)
    def and_yn(self, a, b): return self._add("ANDYN", a, b)
    def or_ny(self, a, b): return self._add("ORNY", a, b)
    def or_yn(self, a, b): return self._add("ORYN", a, b)
    def not_(self, a): return self._add("NOT", a)
    def copy(self, a): return self._add("COPY", a)
    def constant(self, v: bool): return self._add("CONST1" if v else "CONST0")

    def mux_naive(self, a, b, c):          # gates.rs:189-199
        return self.or_(self.and_(a, b), self.and_(self.not_(a), c))

    def mux(self, a, b, c):
        """a ? b : c as the sound fused multiplexer (what gates.rs:157-183 intends): AND(a, b) and
        AND(!a, c) are added at level 1 and key-switched once -- one level, 2 blind rotations."""
        return self._add("MUX", a, b, c)

    # src/circuits.rs:3-6 (compare_bit); the comparator built from it
    def compare_bit(self, a, b, lsb_carry):
        """carry out of one comparator cell: a == b ? lsb_carry : a."""
        return self.mux(self.xnor_true(a, b), lsb_carry, a)

    def xnor_true(self, a, b):
        """logical XNOR.  Gates::xnor decrypts to XOR in the reference (gates.rs:86-90, pinned by its own
        test, :575-579), so equality is NOT(xor)."""
        return self.not_(self.xor(a, b))

    def greater_than(self, a: Sequence[int], b: Sequence[int]) -> int:
        """a > b for little-endian bit vectors: fold compare_bit from the least significant bit."""
        assert len(a) == len(b)
        carry = self.constant(False)
        for x, y in zip(a, b):
            carry = self.compare_bit(x, y, carry)
        return carry

    def equals(self, a: Sequence[int], b: Sequence[int]) -> int:
        """a == b: AND-tree over the per-bit equalities (src/circuits.rs leaves `equals` empty)."""
        assert len(a) == len(b)
        eq = [self.xnor_true(x, y) for x, y in zip(a, b)]
        while len(eq) > 1:
            eq = [self.and_(eq[i], eq[i + 1]) if i + 1 < len(eq) else eq[i] for i in range(0, len(eq), 2)]
        return eq[0]

    # examples/add_two_numbers.rs:11-29
    def full_adder(self, a, b, c) -> Tuple[int, int]:
        a_xor_b = self.xor(a, b)
        a_and_b = self.and_(a, b)
        a_xor_b_and_c = self.and_(a_xor_b, c)
        s = self.xor(a_xor_b, c)
        carry = self.or_(a_and_b, a_xor_b_and_c)
        return s, carry

    # examples/add_two_numbers.rs:31-49
    def add(self, a: Sequence[int], b: Sequence[int], cin: int) -> Tuple[List[int], int]:
        assert len(a) == len(b)
        out, carry = [], cin
        for x, y in zip(a, b):
            s, carry = self.full_adder(x, y, carry)
            out.append(s)
        return out, carry

    def schedule(self) -> Tuple[List[List[int]], Dict[int, List[int]]]:
        """(levels, free): bootstrapped gates grouped by depth 1..D, and the free gates
        (NOT/COPY/constants, which take their operand's depth) grouped by depth 0..D in
        recording order."""
        depth: Dict[int, int] = {}
        lv: Dict[int, List[int]] = {}
        free: Dict[int, List[int]] = {}
        for w, g in enumerate(self.gates):
            if g.op == "INPUT":
                depth[w] = 0
            elif g.op in ("CONST0", "CONST1"):
                depth[w] = 0
                free.setdefault(0, []).append(w)
            elif g.op in ("NOT", "COPY"):
                depth[w] = depth[g.a]
                free.setdefault(depth[w], []).append(w)
            elif g.op == "MUX":
                depth[w] = 1 + max(depth[g.a], depth[g.b], depth[g.c])
                lv.setdefault(depth[w], []).append(w)
            else:
                depth[w] = 1 + max(depth[g.a], depth[g.b])
                lv.setdefault(depth[w], []).append(w)
        return [lv[d] for d in sorted(lv)], free

    def levels(self) -> List[List[int]]:
        return self.schedule()[0]

    def bootstrapped_gate_count(self) -> int:
        """blind rotations per input set (a fused MUX costs two)"""
        return sum(2 if self.gates[w].op == "MUX" else 1 for l in self.levels() for w in l)


def _handle_for(circuit: Circuit, engine: CudaBootstrap):
    """The library-side circuit of `circuit` on `engine`: recorded once and kept on the Circuit object, so
    repeated evaluations reuse the compiled schedule that already sits on the device.  Recording more gates
    or outputs, or switching engines, records afresh."""
    L = _load()
    key = (engine._h.value, len(circuit.gates), tuple(circuit.outputs))
    cached = getattr(circuit, "_native", None)
    if cached is not None and cached[0] == key and cached[2] is engine:   # the very same live engine object
        return cached[1]
    if cached is not None:
        L.tfhe_circuit_destroy(cached[1])
        circuit._native = None
    h = C.c_void_p()
    _check(L.tfhe_circuit_create(engine._h, C.byref(h)))
    try:
        wire = {}
        out = C.c_uint32()
        for wid, g in enumerate(circuit.gates):
            if g.op == "INPUT":
                _check(L.tfhe_circuit_input(h, C.byref(out)))
            elif g.op in ("CONST0", "CONST1"):
                _check(L.tfhe_circuit_constant(h, int(g.op == "CONST1"), C.byref(out)))
            elif g.op == "NOT":
                _check(L.tfhe_circuit_not(h, wire[g.a], C.byref(out)))
            elif g.op == "COPY":
                wire[wid] = wire[g.a]
                continue
            elif g.op == "MUX":
                _check(L.tfhe_circuit_mux(h, wire[g.a], wire[g.b], wire[g.c], C.byref(out)))
            else:
                _check(L.tfhe_circuit_gate(h, _CODE[g.op], wire[g.a], wire[g.b], C.byref(out)))
            wire[wid] = out.value
        for w in circuit.outputs:
            _check(L.tfhe_circuit_output(h, wire[w]))
    except Exception:
        L.tfhe_circuit_destroy(h)
        raise
    circuit._native = (key, h, engine)      # the engine reference keeps it alive as long as the handle
    return h


def release(circuit: Circuit) -> None:
    """Free the library-side circuit (before closing its engine)."""
    cached = getattr(circuit, "_native", None)
    if cached is not None:
        _load().tfhe_circuit_destroy(cached[1])
        circuit._native = None


def evaluate(circuit: Circuit, engine: CudaBootstrap, inputs: np.ndarray) -> np.ndarray:
    """Evaluate `circuit` on `inputs` u32[num_inputs][batch][n+1] (one ciphertext per input wire
    and batch element).  Returns u32[num_outputs][batch][n+1].  The circuit is handed to the library
    (tfhe_circuit_*), which levelises it and runs every level as one device batch with the wires
    resident in HBM; `engine.last_kernel_ms()` afterwards holds the summed kernel times."""
    L = _load()
    n1 = engine.params.n + 1
    inputs = np.ascontiguousarray(inputs, dtype=np.uint32)
    if inputs.ndim != 3 or inputs.shape[0] != len(circuit.inputs) or inputs.shape[2] != n1:
        raise ValueError("inputs must be [num_inputs][batch][n+1]")
    batch = inputs.shape[1]
    h = _handle_for(circuit, engine)
    res = np.empty((len(circuit.outputs), batch, n1), dtype=np.uint32)
    _check(L.tfhe_circuit_run(h, inputs.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p), batch))
    lv, pbs, ks = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(L.tfhe_circuit_stats(h, C.byref(lv), C.byref(pbs), C.byref(ks)))
    circuit.last_stats = {"levels": lv.value, "bootstraps": pbs.value, "key_switches": ks.value}
    return res
