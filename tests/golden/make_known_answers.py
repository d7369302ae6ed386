"""Independent pure-Python restatement of the reference's *deterministic integer* code
(no oracle, no numpy): derives known answers the oracle and the GPU engine are pinned to.
Each item cites the reference file:line it follows.  Run:  python tests/golden/make_known_answers.py
Output: tests/golden/known_answers.json (committed)."""
import json
import math
import os

N = 1024
M32 = 0xFFFFFFFF

SETS = {  # src/params.rs:91-404: (n, l, bgbit, basebit, iks_t)
    "80": (550, 3, 6, 2, 7), "110": (630, 3, 6, 2, 8), "128": (700, 3, 6, 2, 9),
    "uint1": (700, 2, 10, 2, 8), "uint2": (687, 1, 18, 4, 3), "uint3": (820, 1, 23, 6, 2),
    "uint4": (820, 1, 22, 5, 3), "uint5": (1071, 1, 22, 6, 3), "uint7": (1160, 1, 22, 7, 3),
}


def f64_to_torus(d):  # src/utils.rs:9-12
    return int(math.fmod(d, 1.0) * 4294967296.0) & M32


def decomposition_offset(l, bgbit):  # src/key.rs:78-89
    off = 0
    for i in range(l):
        off = (off + (1 << bgbit) // 2 * (1 << (32 - (i + 1) * bgbit))) & M32
    return off


def prec_offset(basebit, t):  # src/trgsw.rs:345
    return 1 << (32 - (1 + basebit * t))


def div_round(a, b):  # src/lut/generator.rs:264-266
    return (a + b // 2) // b


def lut(table, m, scale=None):  # src/lut/generator.rs:89-137, src/lut/encoder.rs:66-73
    scale = 1.0 / (2.0 * m) if scale is None else scale
    raw = [0] * N
    for x in range(m):
        enc = f64_to_torus((table[x] % m) * scale)
        for i in range(div_round(x * N, m), min(div_round((x + 1) * N, m), N)):
            raw[i] = enc
    off = div_round(N, 2 * m)
    rot = [raw[(i + off) % N] for i in range(N)]
    for i in range(N - off, N):
        rot[i] = (-rot[i]) & M32
    return rot


def rle(v):
    out = []
    for x in v:
        if out and out[-1][0] == x:
            out[-1][1] += 1
        else:
            out.append([x, 1])
    return out


def x_k(a, k):  # src/trgsw.rs:307-330
    res = [0] * N
    if k < N:
        for i in range(N - k):
            res[i + k] = a[i]
        for i in range(N - k, N):
            res[i + k - N] = M32 - a[i]
    else:
        for i in range(2 * N - k):
            res[i + k - N] = M32 - a[i]
        for i in range(2 * N - k, N):
            res[i - (2 * N - k)] = a[i]
    return res


def main():
    ka = {"f64_to_torus": {str(d): f64_to_torus(d) for d in (0.125, -0.125, 0.25, -0.25, 0.5, 1.0 / 64, 0.0)},
          "params": {}, "div_round": [[5, 2, 3], [4, 2, 2], [3, 2, 2], [1, 2, 1], [0, 2, 0]],  # generator.rs:350-356
          "lut_rle": {}, "x_k": {}}
    for name, (n, l, bgbit, basebit, t) in SETS.items():
        ka["params"][name] = {"n": n, "l": l, "bgbit": bgbit, "basebit": basebit, "iks_t": t,
                              "decomposition_offset": decomposition_offset(l, bgbit),
                              "prec_offset": prec_offset(basebit, t),
                              "ksk_rows": N * t * (1 << basebit)}
    luts = {"m2_id": ([0, 1], 2), "m2_not": ([1, 0], 2), "m2_one": ([1, 1], 2),
            "m4_inc": ([1, 2, 3, 0], 4), "m16_sq": ([(x * x) % 16 for x in range(16)], 16),
            "m32_mod16": ([x % 16 for x in range(32)], 32), "m3_id": ([0, 1, 2], 3)}
    for key, (table, m) in luts.items():
        ka["lut_rle"][key] = {"table": table, "m": m, "rle": rle(lut(table, m))}
    ka["lut_rle"]["m2_id_scale_half"] = {"table": [0, 1], "m": 2, "scale": 0.5, "rle": rle(lut([0, 1], 2, 0.5))}
    base = [(i * 2654435761 + 12345) & M32 for i in range(N)]
    ka["x_k"]["base_formula"] = "(i*2654435761 + 12345) mod 2^32"
    for k in (0, 1, 511, 1023, 1024, 1025, 2047, 2048):
        r = x_k(base, k)
        ka["x_k"][str(k)] = {"first4": r[:4], "last4": r[-4:], "xor": __import__("functools").reduce(lambda a, b: a ^ b, r)}
    ka["gate_offsets"] = {"NAND": f64_to_torus(0.125), "AND": f64_to_torus(-0.125), "OR": f64_to_torus(0.125),
                          "XOR": f64_to_torus(0.25), "XNOR": f64_to_torus(-0.25), "NOR": f64_to_torus(-0.125),
                          "ANDNY": f64_to_torus(-0.125), "ANDYN": f64_to_torus(-0.125),
                          "ORNY": f64_to_torus(0.125), "ORYN": f64_to_torus(0.125)}  # gates.rs:54-150
    ka["testvec_b"] = f64_to_torus(0.125)  # key.rs:91-100
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    json.dump(ka, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
