"""Levelised circuit evaluation: host scheduling logic on the CPU; on the GPU the
reference's ripple-carry adder (examples/add_two_numbers.rs:11-49) over a batch of encrypted
operands, word for word against the oracle evaluating the same gates one by one."""
import numpy as np
import pytest

import oracle as O
import rs_tfhe_b200 as T
from rs_tfhe_b200.circuit import Circuit


def adder(bits: int):
    c = Circuit()
    a = [c.input() for _ in range(bits)]
    b = [c.input() for _ in range(bits)]
    cin = c.constant(False)
    s, carry = c.add(a, b, cin)
    for w in s:
        c.output(w)
    c.output(carry)
    return c


def test_schedule_levels_and_free_gates():
    c = adder(8)
    levels, free = c.schedule()
    assert c.bootstrapped_gate_count() == 5 * 8
    assert len(levels) == 2 * 8 + 1            # xor/and, then a carry chain of and->or per bit
    assert free[0] == [16]                     # the constant carry-in
    seen = set(c.inputs) | {16}
    for lv in levels:                          # every operand is ready before its level
        for w in lv:
            g = c.gates[w]
            assert g.a in seen and g.b in seen
        seen |= set(lv)
    m = Circuit()
    x, y, z = m.input(), m.input(), m.input()
    o = m.mux_naive(x, y, z)
    lv, fr = m.schedule()
    assert [len(l) for l in lv] == [2, 1] and len(fr[0]) == 1 and m.gates[o].op == "OR"


@pytest.mark.gpu
def test_ripple_carry_adder_batch_matches_oracle():
    from common import keys
    K, ck = keys("128")
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    try:
        e.load_cloud_key(ck)
        from rs_tfhe_b200.circuit import evaluate
        bits, batch = 8, 24
        c = adder(bits)
        r = np.random.default_rng(4)
        xa = r.integers(0, 256, batch)
        xb = r.integers(0, 256, batch)
        xa[0], xb[0] = 42, 137                 # examples/lut_add_two_numbers.rs:52-54
        in_bits = np.array([[(v >> i) & 1 for v in xa] for i in range(bits)] +
                           [[(v >> i) & 1 for v in xb] for i in range(bits)], dtype=bool)
        inputs = np.stack([K.encrypt_bool_batch(row, 500 + i) for i, row in enumerate(in_bits)])
        out = evaluate(c, e, inputs)
        dec = np.stack([K.decrypt_bool_batch(o) for o in out])
        total = sum(dec[i].astype(np.int64) << i for i in range(bits + 1))
        assert np.array_equal(total, xa + xb)
        # oracle, gate by gate in recording order, first 2 batch elements
        nb = 2
        w = {}
        for wid, g in enumerate(c.gates):
            if g.op == "INPUT":
                w[wid] = inputs[c.inputs.index(wid), :nb]
            elif g.op in ("CONST0", "CONST1"):
                v = np.zeros((nb, 701), dtype=np.uint32)
                v[:, -1] = 0x20000000 if g.op == "CONST1" else 0xE0000001
                w[wid] = v
            elif g.op == "NOT":
                w[wid] = (0 - w[g.a]).astype(np.uint32)
            elif g.op == "COPY":
                w[wid] = w[g.a]
            else:
                w[wid] = K.batch_gate(O.GATE_CODE[g.op], np.stack([w[g.a], w[g.b]], axis=1))
        ref = np.stack([w[o] for o in c.outputs])
        assert np.array_equal(out[:, :nb], ref)
    finally:
        e.close()
