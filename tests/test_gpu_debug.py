"""GPU parity through the test hooks of include/p2v.h: device field routines on the lazy-representation edge values
(SURVEY 8 rows a1/a2), recomputed Merkle roots, and verifier branches that no honest transcript reaches (zeta = 1 in
evalLagrange0, x = zeta / x = omega*zeta in combineInitial with `inv 0 = 0`), each against the oracle (which decodes the
JSON with its OWN reader, oracle/json_reader.hpp)."""
import itertools
import os
import sys

import numpy as np
import pytest

import fixtures
import oracle_lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyref  # noqa: E402

pytestmark = pytest.mark.gpu
P = fixtures.P
EDGE = [0, 1, 2, 7, P - 2, P - 1, P, P + 1, 2**64 - 1, 2**64 - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**63, 0xFFFFFFFF00000000, 0xFFFFFFFE00000002]


def operands(seed=0, extra=3000):
    rng = np.random.default_rng(seed)
    pairs = list(itertools.product(EDGE, EDGE))
    a = np.array([x for x, _ in pairs] + [int(v) for v in rng.integers(0, 2**64, extra, dtype=np.uint64)], dtype=np.uint64)
    b = np.array([y for _, y in pairs] + [int(v) for v in rng.integers(0, 2**64, extra, dtype=np.uint64)], dtype=np.uint64)
    return a, b


def test_base_field_ops_on_edge_values(ctx):
    """Algebra/Goldilocks.hs:140-175 with Python integers as the reference (the reference itself computes on Integer)."""
    a, b = operands()
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    want = {
        0: [(x + y) % P for x, y in zip(ai, bi)],
        1: [(x - y) % P for x, y in zip(ai, bi)],
        2: [(x * y) % P for x, y in zip(ai, bi)],
        3: [pow(x % P, P - 2, P) for x in ai],          # inv 0 = 0 (Goldilocks.hs:155-156)
        4: [(-x) % P for x in ai],
        5: [(x * (y & 0xFFFFFFFF)) % P for x, y in zip(ai, bi)],
        6: [(x + (y << 64)) % P for x, y in zip(ai, bi)],
        8: [x % P for x in ai],
        9: [pow(x % P, 7, P) for x in ai],
    }
    for op, w in want.items():
        got = ctx.field_op(op, a, b)
        assert [int(x) for x in got] == w, "base op %d" % op
    e = np.array([int(v) % 1000 for v in b], dtype=np.uint64)
    got = ctx.field_op(7, a, e)
    assert [int(x) for x in got] == [pow(x % P, int(k), P) for x, k in zip(ai, e)]


def test_extension_field_ops_on_edge_values(ctx):
    """Algebra/GoldilocksExt.hs:54-99: mul, inv (incl. inv 0 = 0 and zero-norm-free edge pairs), add, sub, sqr, scale, *X, pow."""
    a0, a1 = operands(1, 1500)
    b0, b1 = operands(2, 1500)
    n = len(a0)
    # rotate so that (re, im) pairs mix different edge values
    a = np.stack([a0, np.roll(a1, 5)])
    b = np.stack([np.roll(b0, 3), b1])
    E = pyref.E
    ea = [E(int(x) % P, int(y) % P) for x, y in zip(a[0], a[1])]
    eb = [E(int(x) % P, int(y) % P) for x, y in zip(b[0], b[1])]
    pairs = lambda es: ([e.pair()[0] % P for e in es], [e.pair()[1] % P for e in es])
    def check(op, es, bb=b):
        got = ctx.field_op(op, a, bb)
        wr, wi = pairs(es)
        assert [int(x) for x in got[0]] == wr and [int(x) for x in got[1]] == wi, "ext op %d" % op
    check(16, [x * y for x, y in zip(ea, eb)])
    check(18, [x + y for x, y in zip(ea, eb)])
    check(19, [x - y for x, y in zip(ea, eb)])
    check(20, [x * x for x in ea])
    check(21, [x.scale(y.pair()[0]) for x, y in zip(ea, eb)])
    check(22, [x * E(0, 1) for x in ea])
    check(24, [E(0) - x for x in ea])
    # invExt (GoldilocksExt.hs:75-80): (a - bX) / (a^2 - 7 b^2) with inv 0 = 0
    def inv_e(x):
        ar, ai_ = x.pair()
        d = pow((ar * ar - 7 * ai_ * ai_) % P, P - 2, P)
        return E(ar * d % P, (-ai_) * d % P)
    check(17, [inv_e(x) for x in ea])
    assert [int(v) for v in ctx.field_op(17, np.zeros((2, 4), dtype=np.uint64))[:, 0]] == [0, 0]
    ex = np.stack([np.array([int(v) % 300 for v in b[0]], dtype=np.uint64), np.zeros(n, dtype=np.uint64)])
    check(23, [x.pow(int(k)) for x, k in zip(ea, ex[0])], ex)


def _circuit(name):
    common = fixtures.REJECTING.get(name, name)
    return oracle_lib.circuit_from_json(fixtures.read(common, "common"), fixtures.read(name, "vkey"))


@pytest.mark.parametrize("name,n", [("small6", 96), ("fixed4", 64), ("arity5", 40), ("lookup6", 64), ("real5", 64), ("mid5", 33)])
def test_recomputed_merkle_roots_and_all_intermediates(p2v, ctx, name, n):
    """north_star: challenges, recomputed Merkle roots, folded evaluations, vanishing check and verdicts, bit-exact on
    honest and tampered proofs — one call (p2v_verify_intermediates), oracle side decoded by the oracle's own JSON reader."""
    shape, lay, vkey, blob = fixtures.load(name)
    oc = _circuit(name)
    assert np.array_equal(oc.proof_blob(fixtures.read(name, "proof")), blob)
    blobs, words, _ = fixtures.tampered_batch(blob, lay, shape, n, seed=11)
    cir = p2v.Circuit(ctx, shape, oc.vkey_words())
    got = cir.verifyIntermediates(blobs)
    want = oc.verify_batch(blobs, threads=8, fast=True)
    for k in ("challenges", "combined", "eqmask", "status"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["accept"], want["status"] == 0)
    roots = oc.fri_roots(blobs)
    assert np.array_equal(got["roots"], roots)
    # folded evaluations / per-query status are defined where the reference gets that far (eqs hold, earlier queries pass)
    st = want["status"]
    Q = shape.num_queries
    for p in range(n):
        if (int(st[p]) & 0xFF) in (0, 3):
            upto = Q if int(st[p]) == 0 else ((int(st[p]) >> 8) & 0xFF) + 1
            assert np.array_equal(got["folded"][:, p * Q: p * Q + upto], want["folded"][:, p * Q: p * Q + upto]), p
    assert np.array_equal(got["qstatus"], want["qstatus"])
    # the roots of an untouched copy are the caps it is checked against
    assert (st[words < 0] == 0).all()


@pytest.mark.parametrize("name", ["small6", "real5", "reallu6"])
def test_zeta_equal_one(p2v, ctx, name):
    """evalLagrange0 at zeta = 1 returns 1 (Algebra/Poly.hs:14-17) — unreachable through Fiat-Shamir, forced here."""
    shape, lay, vkey, blob = fixtures.load(name)
    oc = _circuit(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 8
    blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, n, seed=2)
    ch = oc.verify_batch(blobs, threads=4)["challenges"].copy()
    r = shape.num_challenges
    zi = 3 * r + (4 * r if shape.num_lookup_polys > 0 else 0)
    ch[zi, :] = 1
    ch[zi + 1, :] = 0
    ch[zi, 1] = P + 1  # non-canonical spelling of 1
    got = cir.verifyIntermediates(blobs, challenges_in=ch, want_roots=False)
    want = oc.verify_with_challenges(blobs, ch)
    assert np.array_equal(got["combined"], want["combined"])
    assert np.array_equal(got["eqmask"], want["eqmask"])
    assert np.array_equal(got["status"], want["status"])


@pytest.mark.parametrize("name", ["small6", "fixed4", "real5"])
def test_point_equal_zeta_in_combine_initial(p2v, ctx, name):
    """combineInitial divides by (x - zeta) and (x - omega*zeta) (Plonk/FRI.hs:151-207); with zeta = x of a query, or
    zeta = x/omega, the denominator is 0 and `inv 0 = 0` (Algebra/Goldilocks.hs:155, GoldilocksExt.hs:75-80) decides the value."""
    shape, lay, vkey, blob = fixtures.load(name)
    oc = _circuit(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 6
    blobs = np.tile(blob, (n, 1))
    ch = oc.verify_batch(blobs, threads=4)["challenges"].copy()
    r, Q = shape.num_challenges, shape.num_queries
    zi = 3 * r + (4 * r if shape.num_lookup_polys > 0 else 0)
    lde = shape.degree_bits + shape.rate_bits
    eta, omega = pyref.subgroup_generator(lde), pyref.subgroup_generator(shape.degree_bits)
    for p in range(n):
        q = p % Q
        idx = int(ch[-Q + q, p])
        x = pyref.MUL_GEN * pow(eta, pyref.reverse_bits(lde, idx), P) % P
        if p % 2 == 0:
            ch[zi, p], ch[zi + 1, p] = x, 0                                   # x - zeta = 0
        else:
            ch[zi, p], ch[zi + 1, p] = x * pow(omega, P - 2, P) % P, 0          # x - omega*zeta = 0
    got = cir.verifyIntermediates(blobs, challenges_in=ch, want_roots=False)
    want = oc.verify_with_challenges(blobs, ch)
    assert np.array_equal(got["combined"], want["combined"])
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["qstatus"], want["qstatus"])
    # the query whose point hits zeta: compare the folded value wherever the reference computes it
    for p in range(n):
        q = p % Q
        if int(want["qstatus"][p, q]) & 0xFF in (0, 3):
            assert got["folded"][0, p * Q + q] == want["folded"][0, p * Q + q] and got["folded"][1, p * Q + q] == want["folded"][1, p * Q + q]


def test_stage_planes_layout(p2v, ctx):
    """K0 alone (p2v_stage): pp[w][n] for the per-proof part, qp[wq*Q + q][n] for the per-query parts."""
    shape, lay, vkey, blob = fixtures.load("small6")
    cir = p2v.Circuit(ctx, shape, vkey)
    rng = np.random.default_rng(3)
    for n in (1, 31, 77):
        blobs = rng.integers(0, 2**64, size=(n, lay.blob_words), dtype=np.uint64)
        planes = cir.stage(blobs)
        pw, qw, Q = lay.proof_words, lay.query_words, shape.num_queries
        assert np.array_equal(planes[:pw], blobs[:, :pw].T)
        qpart = blobs[:, pw:].reshape(n, Q, qw)           # [p][q][wq]
        want = qpart.transpose(2, 1, 0).reshape(qw * Q, n)  # [wq*Q + q][p]
        assert np.array_equal(planes[pw:], want)
