"""Gate programs vanish on honestly generated witness rows (CPU).  The accepting fixtures multiply every gate
constraint by a zero filter, and the two restatements were written from the same Haskell, so
this is the independent semantic check: each gate's witness is computed from what the gate MEANS (a Poseidon
permutation, an interpolation, a mux, ...) and both restatements must return an all-zero constraint vector on it
and a non-zero one after a single wire is changed."""
import json
import os
import sys

import numpy as np
import pytest

import fixtures

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyref  # noqa: E402

P = pyref.P
E = pyref.E
rng = np.random.default_rng(2024)


def rnd():
    return int(rng.integers(0, P, dtype=np.uint64))


def rnd_ext():
    return E(rnd(), rnd())


def poseidon_gate_row(swap):
    """Wires of PoseidonGate (Gate/Custom/Poseidon.hs:144-150) for a random input: the s-box inputs of the
    fast-partial schedule, out = permutation(in') with in' = swapped input."""
    w = [0] * 135
    inp = [rnd() for _ in range(12)]
    w[0:12] = inp
    w[24] = swap
    delta = [swap * (inp[i + 4] - inp[i]) % P for i in range(4)]
    w[25:29] = delta
    st = [(inp[i] + delta[i]) % P for i in range(4)] + [(inp[i] - delta[i - 4]) % P for i in range(4, 8)] + inp[8:]
    perm_in = list(st)
    mds = lambda s: [sum(pyref.MDS[i][j] * s[j] for j in range(12)) % P for i in range(12)]
    for r in range(4):
        st = [(x + pyref.RC[12 * r + i]) % P for i, x in enumerate(st)]
        if r:
            w[29 + 12 * (r - 1): 29 + 12 * r] = st
        st = mds([pow(x, 7, P) for x in st])
    st = [(x + c) % P for x, c in zip(st, pyref.FIRST_RC)]
    st = [st[0]] + [sum(pyref.INIT_MAT[11 * j + i] * st[j + 1] for j in range(11)) % P for i in range(11)]
    for r in range(22):
        w[65 + r] = st[0]
        y = (pow(st[0], 7, P) + (pyref.PARTIAL_RC[r] if r < 21 else 0)) % P
        d = (pyref.MDS[0][0] * y + sum(pyref.W_HATS[11 * r + i] * st[i + 1] for i in range(11))) % P
        st = [d] + [(st[i + 1] + y * pyref.VS[11 * r + i]) % P for i in range(11)]
    for r in range(4):
        st = [(x + pyref.RC[12 * (26 + r) + i]) % P for i, x in enumerate(st)]
        w[87 + 12 * r: 87 + 12 * (r + 1)] = st
        st = mds([pow(x, 7, P) for x in st])
    w[12:24] = st
    assert st == pyref.permutation(perm_in)  # the fast schedule inside the gate IS the permutation
    return [E(x) for x in w]


def honest_row(gate, consts):
    """-> list of 135 FExt wires on which the gate's constraints must vanish."""
    name = gate[0]
    w = [rnd_ext() for _ in range(135)]
    c0, c1 = consts
    ee = lambda i: (w[i], w[i + 1])
    put = lambda i, v: w.__setitem__(slice(i, i + 2), [v[0], v[1]])
    if name == "ArithmeticGate":
        for i in range(gate[1]):
            j = 4 * i
            w[j + 3] = c0 * w[j] * w[j + 1] + c1 * w[j + 2]
    elif name == "ArithmeticExtensionGate":
        for i in range(gate[1]):
            j = 8 * i
            put(j + 6, pyref.ee_add(pyref.ee_mul(pyref.ee_scale(c0, ee(j)), ee(j + 2)), pyref.ee_scale(c1, ee(j + 4))))
    elif name == "MulExtensionGate":
        for i in range(gate[1]):
            j = 6 * i
            put(j + 4, pyref.ee_mul(pyref.ee_scale(c0, ee(j)), ee(j + 2)))
    elif name == "BaseSumGate":
        L, B = gate[1], gate[2]
        limbs = [int(rng.integers(0, B)) for _ in range(L)]
        for i, l in enumerate(limbs):
            w[1 + i] = E(l)
        w[0] = E(sum(l * B ** i for i, l in enumerate(limbs)))
    elif name == "ConstantGate":
        for i in range(gate[1]):
            w[i] = consts[i]
    elif name == "PublicInputGate":
        pass  # handled by the caller: wires 0..3 = pi hash
    elif name == "ExponentiationGate":
        n = gate[1]
        bits = [int(rng.integers(0, 2)) for _ in range(n)]  # little endian in the wires, processed MSB first
        base = rnd_ext()
        w[0] = base
        for i, b in enumerate(bits):
            w[1 + i] = E(b)
        acc = pyref.E1
        for i in range(n):
            cur = bits[n - 1 - i]
            acc = (acc * acc if i else pyref.E1) * (base if cur else pyref.E1)
            w[n + 2 + i] = acc
        w[n + 1] = acc
        e = sum(b << i for i, b in enumerate(bits))
        assert acc == base.pow(e)  # the gate computes base^(sum 2^i e_i)
    elif name in ("ReducingGate", "ReducingExtensionGate"):
        n, ext = gate[1], name == "ReducingExtensionGate"
        alpha, acc = ee(2), ee(4)
        for i in range(n):
            coeff = ee(6 + 2 * i) if ext else (w[6 + i], pyref.E0)
            acc = pyref.ee_add(pyref.ee_mul(acc, alpha), coeff)
            put(6 + (2 * n if ext else n) + 2 * i if i < n - 1 else 0, acc)
    elif name == "RandomAccessGate":
        nb, copies, extra = gate[1], gate[2], gate[3]
        width = 2 + (1 << nb)
        start = width * copies + extra
        for k in range(copies):
            idx = int(rng.integers(0, 1 << nb))
            w[k * width] = E(idx)
            w[k * width + 1] = w[k * width + 2 + idx]
            for j in range(nb):
                w[start + k * nb + j] = E((idx >> j) & 1)
        for j in range(extra):
            w[copies * width + j] = consts[j]
    elif name == "CosetInterpolationGate":
        bits, degree, weights = gate[1], gate[2], gate[3]
        npts, nint = 1 << bits, ((1 << bits) - 2) // (degree - 1)
        gen = pyref.subgroup_generator(bits)
        dom = [pow(gen, k, P) for k in range(npts)]
        shift = E(rnd())   # coset shift is a base-field value in real use; an ext value also works for the identity
        w[0] = shift
        base = 1 + 2 * (npts + 2)
        x0 = (rnd_ext(), rnd_ext())        # shifted evaluation point, as an ext-of-ext value
        put(base + 4 * nint, x0)
        put(1 + 2 * npts, pyref.ee_scale(shift, x0))
        ev, pr = (pyref.E0, pyref.E0), (pyref.E1, pyref.E0)
        chunks = [list(range(0, degree))] + [list(range(s, min(s + degree - 1, npts))) for s in range(degree, npts, degree - 1)]
        for ci, chunk in enumerate(chunks):
            for idx in chunk:
                val = pyref.ee_scale(E(weights[idx]), ee(1 + 2 * idx))
                term = (x0[0] - E(dom[idx]), x0[1])
                ev, pr = pyref.ee_add(pyref.ee_mul(term, ev), pyref.ee_mul(val, pr)), pyref.ee_mul(term, pr)
            if ci < len(chunks) - 1:
                put(base + 2 * ci, ev)
                put(base + 2 * (nint + ci), pr)
        put(1 + 2 * npts + 2, ev)
    elif name == "PoseidonMdsGate":
        for i in range(12):
            acc = (pyref.E0, pyref.E0)
            for j in range(12):
                acc = pyref.ee_add(acc, pyref.ee_scale(E(pyref.MDS[i][j]), ee(2 * j)))
            put(2 * (i + 12), acc)
    return w


@pytest.mark.parametrize("swap", [0, 1])
def test_poseidon_gate_on_a_real_permutation(orc, swap):
    shape, *_ = fixtures.load("mid5")
    k = [i for i in range(shape.num_gates) if shape.gates[i].kind == 11][0]
    w = poseidon_gate_row(swap)
    wires = np.array([e.pair() for e in w], dtype=np.uint64)
    zero2 = np.zeros((2, 2), dtype=np.uint64)
    got = orc.gate_constraints(shape, k, wires, zero2, np.zeros(4, dtype=np.uint64))
    assert len(got) == 123 and not got.any()
    assert all(e == pyref.E0 for e in pyref.gate_constraints(("PoseidonGate", 12), w, [pyref.E0, pyref.E0], [0] * 4))
    wires[40, 0] = (int(wires[40, 0]) + 1) % P  # one s-box input off by one
    assert orc.gate_constraints(shape, k, wires, zero2, np.zeros(4, dtype=np.uint64)).any()


def test_every_other_gate_vanishes_on_honest_rows(orc):
    shape, *_ = fixtures.load("mid5")
    common = json.loads(fixtures.read("mid5", "common"))
    seen = set()
    for k, text in enumerate(common["gates"]):
        g = pyref.parse_gate(text)
        if g[0] in ("PoseidonGate", "NoopGate"):
            continue
        consts = [rnd_ext(), rnd_ext()]
        pih = [rnd() for _ in range(4)]
        w = honest_row(g, consts)
        if g[0] == "PublicInputGate":
            for i in range(4):
                w[i] = E(pih[i])
        wires = np.array([e.pair() for e in w], dtype=np.uint64)
        cs = np.array([e.pair() for e in consts], dtype=np.uint64)
        got = orc.gate_constraints(shape, k, wires, cs, np.array(pih, dtype=np.uint64))
        assert len(got) == shape.gates[k].num_constraints > 0, g[0]
        assert not got.any(), g[0]
        assert all(e == pyref.E0 for e in pyref.gate_constraints(g, w, consts, pih)), g[0]
        # break one wire the gate reads: some constraint must fire
        wires[0, 0] = (int(wires[0, 0]) + 1) % P
        assert orc.gate_constraints(shape, k, wires, cs, np.array(pih, dtype=np.uint64)).any(), g[0]
        seen.add(g[0])
    assert len(seen) == 12


def test_coset_interpolation_gate_interpolates(orc):
    """eval_result of the honest CosetInterpolationGate row equals the Lagrange interpolant through
    (shift*w^k, value_k) evaluated at eval_loc (base-field instance, checked numerically)."""
    bits, degree = 4, 6
    npts = 1 << bits
    gen = pyref.subgroup_generator(bits)
    dom = [pow(gen, k, P) for k in range(npts)]
    weights = []
    for i in range(npts):
        pr = 1
        for j in range(npts):
            if j != i:
                pr = pr * (dom[i] - dom[j]) % P
        weights.append(pyref.inv(pr))
    shift, x = rnd(), rnd()
    vals = [rnd() for _ in range(npts)]
    x0 = x * pyref.inv(shift) % P
    ev, pr = 0, 1
    for k in range(npts):  # the gate's running barycentric recurrence, unchunked
        ev, pr = ((x0 - dom[k]) * ev + weights[k] * vals[k] % P * pr) % P, (x0 - dom[k]) * pr % P
    lag = 0
    for i in range(npts):
        num = den = 1
        for j in range(npts):
            if j != i:
                num = num * (x - shift * dom[j]) % P
                den = den * (shift * dom[i] - shift * dom[j]) % P
        lag = (lag + vals[i] * num % P * pyref.inv(den)) % P
    assert ev == lag
