"""The two independent restatements of the reference verifier — C++ (oracle/*.hpp) and pure Python
(oracle/pyref.py, which reads the JSON itself) — must agree on challenges, combined constraints and verdicts, on
the bundled fixtures and on every class of tampered proof.  (CPU only; small shapes: pure Python is slow.)"""
import copy
import json
import os
import sys

import numpy as np
import pytest

import fixtures

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyref  # noqa: E402

P = pyref.P


def refill(proof_json, blob):
    """Inverse of the flattening: write the words of `blob` back into a copy of the proof JSON, walking the fields
    in the order of Types.hs:251-279 (the p2v_layout order)."""
    it = iter(int(x) for x in blob)
    out = copy.deepcopy(proof_json)
    def dig(d): d["elements"] = [next(it) for _ in d["elements"]]
    def cap(c):
        for d in c: dig(d)
    def ext(xs):
        for i in range(len(xs)): xs[i] = [next(it), next(it)]
    pr = out["proof"]
    cap(pr["wires_cap"]); cap(pr["plonk_zs_partial_products_cap"]); cap(pr["quotient_polys_cap"])
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys",
              "lookup_zs", "lookup_zs_next"):
        ext(pr["openings"][k])
    fri = pr["opening_proof"]
    for c in fri["commit_phase_merkle_caps"]: cap(c)
    ext(fri["final_poly"]["coeffs"])
    fri["pow_witness"] = next(it)
    out["public_inputs"] = [next(it) for _ in out["public_inputs"]]
    for rd in fri["query_round_proofs"]:
        for ep in rd["initial_trees_proof"]["evals_proofs"]:
            ep[0] = [next(it) for _ in ep[0]]
            cap(ep[1]["siblings"])
        for st in rd["steps"]:
            ext(st["evals"])
            cap(st["merkle_proof"]["siblings"])
    assert next(it, None) is None
    return out


def test_python_twin_kat():
    assert pyref.permutation(list(range(12)))[0] == 0xd64e1e3efc5b8e9e
    assert pyref.sponge(list(range(1, 10)))[0] == 0x5a90f7c562413c2b


@pytest.mark.parametrize("name", ["small6", "fixed4", "arity5", "lookup6", "mid5", "small6_badfinal", "small6_badlayer0", "small6_badlayer1",
                                  "real5", "real5_badwitness", "real5_badcopy", "reallu6", "reallu6_badlookup"])
def test_twins_agree_on_fixtures(orc, name):
    shape, lay, vkey, blob = fixtures.load(name)
    common, vk, proof = pyref.load_fixture(fixtures.GOLDEN, name, fixtures.REJECTING.get(name))
    tr = {}
    st_py = pyref.verify(common, vk, proof, tr)
    res = orc.verify_batch(shape, vkey, blob, threads=1, fast=False)
    assert st_py == int(res["status"][0])
    assert tr["challenges"] == [int(x) for x in res["challenges"][:, 0]]
    assert tr["combined"] == [int(x) for x in res["combined"][:, 0]]


@pytest.mark.parametrize("name", ["small6", "lookup6", "fixed4"])
def test_twins_agree_on_tamper_matrix(orc, name):
    shape, lay, vkey, blob = fixtures.load(name)
    common, vk, proof = pyref.load_fixture(fixtures.GOLDEN, name)
    words = fixtures.tamper_words(lay, shape)
    names = sorted(words)
    rng = np.random.default_rng(4)
    extra = [int(x) for x in rng.integers(0, lay.blob_words, 6)]
    targets = [(nm, words[nm]) for nm in names] + [("rand%d" % i, wd) for i, wd in enumerate(extra)]
    blobs = np.tile(blob, (len(targets), 1))
    for i, (_, wd) in enumerate(targets):
        blobs[i, wd] = (int(blobs[i, wd]) + 1 + i) % P
    res = orc.verify_batch(shape, vkey, blobs, threads=4, fast=False)
    for i, (nm, wd) in enumerate(targets):
        tr = {}
        st_py = pyref.verify(common, vk, refill(proof, blobs[i]), tr)
        assert st_py == int(res["status"][i]), (nm, hex(st_py), hex(int(res["status"][i])))
        assert tr["challenges"] == [int(x) for x in res["challenges"][:, i]], nm
        assert tr["combined"] == [int(x) for x in res["combined"][:, i]], nm


def test_gate_programs_agree_on_random_rows(orc):
    """Every gate kind of the standard recursion gate set, unfiltered constraint vectors on random openings
    (inputs in the spirit of Gate/Computation.hs:187-198), C++ vs Python."""
    shape, lay, vkey, blob = fixtures.load("mid5")
    common = json.loads(fixtures.read("mid5", "common"))
    rng = np.random.default_rng(9)
    wires = rng.integers(0, P, size=(shape.num_wires, 2), dtype=np.uint64)
    consts = rng.integers(0, P, size=(2, 2), dtype=np.uint64)
    pih = rng.integers(0, P, size=4, dtype=np.uint64)
    w = [pyref.E(int(a), int(b)) for a, b in wires]
    c = [pyref.E(int(a), int(b)) for a, b in consts]
    kinds = set()
    for k, text in enumerate(common["gates"]):
        g = pyref.parse_gate(text)
        got = orc.gate_constraints(shape, k, wires, consts, pih)
        want = pyref.gate_constraints(g, w, c, [int(x) for x in pih])
        assert len(got) == len(want) == shape.gates[k].num_constraints, text[:30]
        assert [tuple(int(x) for x in row) for row in got] == [e.pair() for e in want], text[:30]
        kinds.add(g[0])
    assert len(kinds) == 14
