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


def test_schedule_fused_mux_and_comparator():
    c = Circuit()
    a = [c.input() for _ in range(4)]
    b = [c.input() for _ in range(4)]
    gt = c.greater_than(a, b)
    eq = c.equals(a, b)
    c.output(gt); c.output(eq)
    levels, _ = c.schedule()
    # xor (1 level) then a chain of 4 fused muxes; the AND tree of equals sits beside it
    assert len(levels) == 1 + 4
    assert c.bootstrapped_gate_count() == 2 * 4 + 2 * 4 + 3     # 2 x 4 xors, 4 fused muxes (2 rotations each), 3 ands


@pytest.mark.gpu
def test_fused_mux_is_sound_and_matches_oracle_composition():
    """Truth table of the fused MUX, and word-for-word equality with the oracle evaluating the same data
    flow: AND / ANDNY blind rotations, sample_extract_index(., 0) at N = 1024, add + 1/8, one key switch."""
    from common import keys
    from rs_tfhe_b200.circuit import evaluate
    K, ck = keys("128")
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    try:
        e.load_cloud_key(ck)
        c = Circuit()
        s, a, b = c.input(), c.input(), c.input()
        c.output(c.mux(s, a, b))
        c.output(c.mux(c.not_(s), a, c.constant(True)))      # free NOT / constant operands
        combos = np.array([[(v >> i) & 1 for v in range(8)] for i in range(3)], dtype=bool)   # s, a, b
        reps = 4
        bits = np.repeat(combos, reps, axis=1)
        inputs = np.stack([K.encrypt_bool_batch(row, 700 + i) for i, row in enumerate(bits)])
        out = evaluate(c, e, inputs)
        assert c.last_stats == {"levels": 1, "bootstraps": 4, "key_switches": 2}
        sv, av, bv = bits
        assert np.array_equal(K.decrypt_bool_batch(out[0]), np.where(sv, av, bv))
        assert np.array_equal(K.decrypt_bool_batch(out[1]), np.where(~sv, av, True))
        # oracle composition for the first output, every batch element
        AND, ANDNY = O.GATE_CODE["AND"], O.GATE_CODE["ANDNY"]
        for j in range(bits.shape[1]):
            u = []
            for op, x in ((AND, inputs[1, j]), (ANDNY, inputs[2, j])):
                ra, rb, _ = K.blind_rotate(K.gate_prep(op, inputs[0, j], x))
                u.append(O.sample_extract_index(ra, rb, 0))
            t = (u[0] + u[1]).astype(np.uint32)
            t[-1] = np.uint32((int(t[-1]) + 0x20000000) & 0xFFFFFFFF)
            assert np.array_equal(out[0, j], K.identity_key_switching(t)), j
    finally:
        e.close()


@pytest.mark.gpu
def test_comparator_batch():
    from common import keys
    from rs_tfhe_b200.circuit import evaluate
    K, ck = keys("128")
    e = T.CudaBootstrap(T.SECURITY_128_BIT, 0)
    try:
        e.load_cloud_key(ck)
        bits, batch = 8, 40
        c = Circuit()
        a = [c.input() for _ in range(bits)]
        b = [c.input() for _ in range(bits)]
        c.output(c.greater_than(a, b))
        c.output(c.equals(a, b))
        r = np.random.default_rng(8)
        xa = r.integers(0, 256, batch)
        xb = r.integers(0, 256, batch)
        xb[:8] = xa[:8]                                   # some equal pairs
        xb[8:12] = xa[8:12] ^ 1                           # differ in the lowest bit only
        in_bits = np.array([[(v >> i) & 1 for v in xa] for i in range(bits)] +
                           [[(v >> i) & 1 for v in xb] for i in range(bits)], dtype=bool)
        inputs = np.stack([K.encrypt_bool_batch(row, 800 + i) for i, row in enumerate(in_bits)])
        out = evaluate(c, e, inputs)
        assert np.array_equal(K.decrypt_bool_batch(out[0]), xa > xb)
        assert np.array_equal(K.decrypt_bool_batch(out[1]), xa == xb)
    finally:
        e.close()
