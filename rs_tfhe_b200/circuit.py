"""Levelised boolean-circuit evaluation on top of the batch-gate engine.

The reference evaluates circuits gate by gate (examples/add_two_numbers.rs:11-49: a ripple-carry
adder built from `xor`, `and`, `or`).  Here a circuit is recorded once, split into levels of
mutually independent bootstrapped gates, and every level goes to the device as ONE mixed-gate
batch (`tfhe_batch_gate_dev` with per-element gate codes) over all `batch` independent input
sets; the wires stay resident in HBM between levels (torch is used only for the gather/scatter
plumbing and the free linear gates NOT/COPY).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import GATES, CudaBootstrap, f64_to_torus

_CODE = {g: i for i, g in enumerate(GATES)}


@dataclass
class _Gate:
    op: str                  # a GATES name, or "NOT" / "COPY" / "CONST0" / "CONST1" / "INPUT"
    a: int = -1
    b: int = -1


@dataclass
class Circuit:
    """Records gates over integer wire ids; wire ids are returned by every method."""
    gates: List[_Gate] = field(default_factory=list)
    inputs: List[int] = field(default_factory=list)
    outputs: List[int] = field(default_factory=list)

    def _add(self, op: str, a: int = -1, b: int = -1) -> int:
        self.gates.append(_Gate(op, a, b))
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
            else:
                depth[w] = 1 + max(depth[g.a], depth[g.b])
                lv.setdefault(depth[w], []).append(w)
        return [lv[d] for d in sorted(lv)], free

    def levels(self) -> List[List[int]]:
        return self.schedule()[0]

    def bootstrapped_gate_count(self) -> int:
        return sum(len(l) for l in self.levels())


def evaluate(circuit: Circuit, engine: CudaBootstrap, inputs: np.ndarray) -> np.ndarray:
    """Evaluate `circuit` on `inputs` u32[num_inputs][batch][n+1] (one ciphertext per input wire
    and batch element).  Returns u32[num_outputs][batch][n+1].  Level by level, device-resident."""
    import torch

    n1 = engine.params.n + 1
    inputs = np.ascontiguousarray(inputs, dtype=np.uint32)
    if inputs.ndim != 3 or inputs.shape[0] != len(circuit.inputs) or inputs.shape[2] != n1:
        raise ValueError("inputs must be [num_inputs][batch][n+1]")
    batch = inputs.shape[1]
    dev = torch.device("cuda", engine.device)
    # one explicit stream carries both torch's gather/scatter and the engine's kernels
    # (handle 0 -- the legacy default stream -- would mean "engine's own stream" to the ABI)
    stream = torch.cuda.Stream(device=dev)
    engine.set_stream(stream.cuda_stream)
    try:
        with torch.cuda.stream(stream):
            nw = len(circuit.gates)
            wires = torch.zeros((nw, batch, n1), dtype=torch.int32, device=dev)
            wires[torch.tensor(circuit.inputs, device=dev)] = torch.from_numpy(inputs.view(np.int32)).to(dev)
            mu = f64_to_torus(0.125)
            levels, free = circuit.schedule()

            def run_free(d: int) -> None:
                # linear gates need no bootstrap (gates.rs:202-218); recording order within a depth
                for w in free.get(d, []):
                    g = circuit.gates[w]
                    if g.op == "NOT":
                        wires[w] = -wires[g.a]
                    elif g.op == "COPY":
                        wires[w] = wires[g.a]
                    else:
                        v = mu if g.op == "CONST1" else (1 - mu) & 0xFFFFFFFF
                        wires[w, :, -1] = v - (1 << 32) if v >= (1 << 31) else v

            run_free(0)
            for d, level in enumerate(levels, start=1):
                a_idx = torch.tensor([circuit.gates[w].a for w in level], device=dev)
                b_idx = torch.tensor([circuit.gates[w].b for w in level], device=dev)
                pairs = torch.stack([wires[a_idx], wires[b_idx]], dim=2).reshape(-1, 2, n1).contiguous()
                ops = torch.tensor([_CODE[circuit.gates[w].op] for w in level], dtype=torch.uint8,
                                   device=dev).repeat_interleave(batch).contiguous()
                out = torch.empty((len(level) * batch, n1), dtype=torch.int32, device=dev)
                engine.batch_gate_dev(0, pairs.data_ptr(), out.data_ptr(), len(level) * batch,
                                      d_ops=ops.data_ptr())
                wires[torch.tensor(level, device=dev)] = out.view(len(level), batch, n1)
                run_free(d)
            res = wires[torch.tensor(circuit.outputs, device=dev)].cpu().numpy().view(np.uint32)
            stream.synchronize()
            return res
    finally:
        engine.set_stream(0)
