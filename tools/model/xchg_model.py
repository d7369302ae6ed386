"""numpy model of the index/sign algebra of blind_rotate_kernel_x (rs_tfhe_b200/csrc/blind_rotate.cu).

One 512-point complex transform over a group of 64 threads (2 warps x 32 lanes), 8 values per
thread, three radix-8 passes.  Exchange 1 (pass A -> B) is a padded shared-memory transpose;
exchange 2 (pass B -> C) runs inside each warp: tcgen05.st.32x32b + tcgen05.ld.16x256b swaps
lane bits (4,3) with two register bits, one shfl.xor(16) trades the third bit.  All lane-dependent
selects are absorbed into signs: a radix-8 butterfly with a per-thread sigma = +-1 in its last stage
delivers its outputs with slots s and s^4 swapped; an input swapped that way yields odd outputs
negated, which the permuted bootstrapping key (forward) and the inverse pass-A twiddles (inverse)
carry.  Run: python tools/model/xchg_model.py  (checks forward and inverse against numpy.fft)."""
import numpy as np
N = 512
S = 73  # exchange-buffer row pitch


def w(n, e):
    return np.exp(-2j * np.pi * e / n)


def dft8(v, inv=False, swap=0):
    """radix-8 butterfly; swap=1: slot s holds X[s ^ 4] (the sigma = -1 form)"""
    k = np.arange(8)
    M = np.exp((2j if inv else -2j) * np.pi * np.outer(k, k) / 8)
    X = M @ v
    return X[k ^ (4 * swap)]


rng = np.random.default_rng(1)
x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
ref = np.fft.fft(x)


def lanes_b(l):  # pass B / B' lane -> (j0, k0 low bits)
    l4, l3, l2 = (l >> 4) & 1, (l >> 3) & 1, (l >> 2) & 1
    return 4 * l2 + 2 * l4 + l3, l & 3, l2


def st32_col(s, im, hi):
    return 16 * (s >> 2) + 8 * im + 4 * ((s >> 1) & 1) + 2 * (s & 1) + hi


# ---------------- forward ----------------
exch = np.zeros(8 * S, complex)
for T in range(64):                      # pass A: T = j0 + 8 j1, registers j2 -> k0
    j0, j1 = T & 7, T >> 3
    v = dft8(np.array([x[j0 + 8 * j1 + 64 * j2] for j2 in range(8)]))
    for k0 in range(8):
        exch[k0 * S + j1 * 9 + j0] = v[k0] * w(512, (j0 + 8 * j1) * k0)
tmem = np.zeros((2, 32, 32), dtype=object)
for T in range(64):                      # pass B
    W, l = T >> 5, T & 31
    j0, k0lo, l2 = lanes_b(l)
    k0 = 4 * W + k0lo
    v = dft8(np.array([exch[k0 * S + j1 * 9 + j0] for j1 in range(8)]), swap=l2)
    tbf = np.array([w(64, j0 * (s ^ (4 * l2))) for s in range(8)])   # per-thread table
    v = v * tbf
    for s in range(8):
        for im in range(2):
            for hi in range(2):
                tmem[W][l][st32_col(s, im, hi)] = (v[s], im, hi)
vC = {}
val = {}
for T in range(64):                      # tcgen05.ld.16x256b.x4, halves H
    W, l = T >> 5, T & 31
    for H in range(2):
        R = [None] * 16
        for rep in range(4):
            for h in range(2):
                for b in range(2):
                    R[4 * rep + 2 * h + b] = tmem[W][16 * H + (l >> 2) + 8 * h][8 * rep + 2 * (l & 3) + b]
        for h in range(2):
            for z in range(2):
                words = [R[4 * (2 * z + im) + 2 * h + hi] for im in range(2) for hi in range(2)]
                assert all(wd[0] == words[0][0] for wd in words)
                assert [(wd[1], wd[2]) for wd in words] == [(0, 0), (0, 1), (1, 0), (1, 1)]
                val[(T, H, h, z)] = words[0][0]
for T in range(64):                      # shuffle stage: keep z = 0, trade z = 1 with lane ^ 16 (no selects)
    for H in range(2):
        for h in range(2):
            vC[(T, 0 + 2 * H + h)] = val[(T, H, h, 0)]
            vC[(T, 4 + 2 * H + h)] = val[(T ^ 16, H, h, 1)]
out = np.zeros(N, complex)
bsk_sign = {}
for T in range(64):                      # pass C: W = k0[2], lane = (k1[2] k0[1] k0[0] k1[1] k1[0])
    W, l = T >> 5, T & 31
    l4 = l >> 4
    k0 = 4 * W + ((l >> 2) & 3)
    k1 = 4 * l4 + (l & 3)
    v = dft8(np.array([vC[(T, p)] for p in range(8)]))
    for k2 in range(8):
        sign = -1.0 if (l4 and (k2 & 1)) else 1.0     # carried by the permuted key in the kernel
        out[k0 + 8 * k1 + 64 * k2] = sign * v[k2]
print('forward max err', np.abs(out - ref).max())

# ---------------- inverse ----------------
spec = ref
uC = {}
for T in range(64):                      # pass C': true spectrum in registers k2
    W, l = T >> 5, T & 31
    l4 = l >> 4
    k0 = 4 * W + ((l >> 2) & 3)
    k1 = 4 * l4 + (l & 3)
    u = dft8(np.array([spec[k0 + 8 * k1 + 64 * k2] for k2 in range(8)]), inv=True, swap=l4)
    tbi = np.conj(np.array([w(64, ((p ^ (4 * l4))) * k1) for p in range(8)]))
    u = u * tbi
    for p in range(8):
        uC[(T, p)] = u[p]
tmem = np.zeros((2, 32, 32), dtype=object)
for T in range(64):                      # shuffle: keep slots 0..3, trade slots 4..7; tcgen05.st.16x256b.x4
    W, l = T >> 5, T & 31
    for H in range(2):
        for h in range(2):
            zval = [uC[(T, 2 * H + h)], uC[(T ^ 16, 4 + 2 * H + h)]]
            for z in range(2):
                for im in range(2):
                    for b in range(2):
                        rep = 2 * z + im
                        tmem[W][16 * H + (l >> 2) + 8 * h][8 * rep + 2 * (l & 3) + b] = (zval[z], im, b)
exch = np.zeros(8 * S, complex)
for T in range(64):                      # tcgen05.ld.32x32b.x32 + pass B'
    W, l = T >> 5, T & 31
    j0, k0lo, l2 = lanes_b(l)
    k0 = 4 * W + k0lo
    v = np.zeros(8, complex)
    for s in range(8):
        words = [tmem[W][l][st32_col(s, im, hi)] for im in range(2) for hi in range(2)]
        assert all(wd[0] == words[0][0] for wd in words)
        v[s] = words[0][0]
    v = dft8(v, inv=True)                # input slots swapped for l2 = 1 -> odd outputs negated
    for j1 in range(8):
        exch[k0 * S + j1 * 9 + j0] = v[j1]
y = np.zeros(N, complex)
for T in range(64):                      # pass A' with the signed twiddles
    j0, j1 = T & 7, T >> 3
    sigma = -1.0 if ((j1 & 1) and (j0 >> 2)) else 1.0
    v = np.array([exch[k0 * S + j1 * 9 + j0] * sigma * np.conj(w(512, (j0 + 8 * j1) * k0)) for k0 in range(8)])
    v = dft8(v, inv=True)
    for j2 in range(8):
        y[j0 + 8 * j1 + 64 * j2] = v[j2]
print('inverse max err', np.abs(y / 512 - x).max())

# bank check of pass B's 128-bit loads (quarter-warps of 8 lanes, 16-byte bank groups)
worst = 0
for W in range(2):
    for q in range(4):
        for j1 in range(8):
            groups = [((4 * W + lanes_b(l)[1]) * S + j1 * 9 + lanes_b(l)[0]) % 8 for l in range(8 * q, 8 * q + 8)]
            worst = max(worst, 8 - len(set(groups)) + 1)
print('pass B worst conflict degree', worst)
