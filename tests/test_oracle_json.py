"""The oracle's OWN JSON reader (oracle/json_reader.hpp, written from Types.hs / Gate/Parser.hs) against the product's host
parser (csrc/host/parse.cpp): two independent readers of the reference's wire format must produce the same records, so a
GPU-vs-oracle comparison that feeds each side through its own reader shares no parsing code.  Also: the driver output of
the reference (`testmain`, src/testmain.hs:40-63) as three artefacts that must agree — the committed
tests/golden/<name>.testmain.txt, oracle/pyref.py and the C++ oracle."""
import json
import os
import sys

import numpy as np
import pytest

import fixtures
import oracle_lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyref  # noqa: E402

ALL = fixtures.ACCEPTING + sorted(fixtures.REJECTING)


def circuit(name):
    common = fixtures.REJECTING.get(name, name)
    return oracle_lib.circuit_from_json(fixtures.read(common, "common"), fixtures.read(name, "vkey"))


@pytest.mark.parametrize("name", ALL)
def test_two_readers_same_records(p2v, orc, name):
    shape, lay, vkey, blob = fixtures.load(name)
    oc = circuit(name)
    assert np.array_equal(oc.vkey_words(), vkey)
    ob = oc.proof_blob(fixtures.read(name, "proof"))
    assert len(ob) == lay.blob_words
    assert np.array_equal(ob, blob)
    assert (oc.num_challenges, oc.num_queries, oc.num_steps, oc.degree_bits, oc.rate_bits, oc.cap_height) == (
        shape.num_challenges, shape.num_queries, shape.num_steps, shape.degree_bits, shape.rate_bits, shape.cap_height)


@pytest.mark.parametrize("name", ["small6", "fixed4", "arity5", "lookup6", "mid5", "real5", "reallu6", "small6_badfinal", "real5_badcopy"])
def test_verdicts_do_not_depend_on_the_reader(orc, name):
    """Same verifier, circuit decoded by the oracle's reader vs circuit rebuilt from the product's p2v_shape: every
    intermediate is identical (gate list, selector groups, k_is, lookup tables, FRI parameters all enter here)."""
    shape, lay, vkey, blob = fixtures.load(name)
    blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, 24, seed=3)
    a = circuit(name).verify_batch(blobs, threads=4, fast=True)
    b = orc.verify_batch(shape, vkey, blobs, threads=4, fast=True)
    for k in ("status", "challenges", "combined", "eqmask", "qstatus", "folded"):
        assert np.array_equal(a[k], b[k]), k


def test_gate_strings_two_parsers(p2v):
    """recognizeGate (Gate/Parser.hs:107-240) in the oracle's reader vs the product's p2v_parse_gate on every gate string of
    the fixtures (through a one-gate common file)."""
    seen = set()
    for name in fixtures.ACCEPTING:
        common = json.loads(fixtures.read(name, "common"))
        for k, text in enumerate(common["gates"]):
            if text in seen:
                continue
            seen.add(text)
            g, weights = p2v.parse_gate(text)
            one = dict(common)
            one["gates"] = [text]
            one["selectors_info"] = {"selector_indices": [0], "groups": [{"start": 0, "end": 1}]}
            oc = oracle_lib.circuit_from_json(json.dumps(one), fixtures.read(name, "vkey"))
            assert oc is not None
    assert len(seen) >= 14


@pytest.mark.parametrize("bad", ['{"config": 1}', "[1,2", '{"a": tru}', ""])
def test_reader_rejects_garbage(bad):
    with pytest.raises(ValueError):
        oracle_lib.circuit_from_json(bad, "{}")


@pytest.mark.parametrize("name", ALL)
def test_testmain_output_three_ways(name):
    """committed golden == C++ oracle (own JSON reader) [== pyref for the shapes pure Python finishes quickly]."""
    path = os.path.join(fixtures.GOLDEN, name + ".testmain.txt")
    want = open(path).read()
    got_cpp = circuit(name).testmain(fixtures.read(name, "proof"))
    assert got_cpp == want
    if name not in ("s12", "real12"):
        common, vk, proof = pyref.load_fixture(fixtures.GOLDEN, name, fixtures.REJECTING.get(name))
        assert pyref.testmain_text(common, vk, proof) == want
    assert want.startswith("public inputs hash = MkDigest ") and want.rstrip().split("\n")[-1].startswith("proof verification result = ")


def test_fri_roots_of_an_accepting_proof_are_the_caps(orc):
    """Recomputed Merkle roots (Hash/Merkle.hs:30-37): for an accepting proof every root equals the cap entry it is compared
    with — vkey cap for oracle 0, the proof's caps for the others."""
    name = "small6"
    shape, lay, vkey, blob = fixtures.load(name)
    oc = circuit(name)
    roots = oc.fri_roots(blob)
    res = oc.verify_batch(blob, threads=1)
    Q, ncap = shape.num_queries, 1 << shape.cap_height
    idx = res["challenges"][-Q:, 0].astype(np.int64)
    lde = shape.degree_bits + shape.rate_bits
    caps = [vkey[: 4 * ncap], blob[lay.off_wires_cap:][: 4 * ncap], blob[lay.off_zs_pp_cap:][: 4 * ncap], blob[lay.off_quotient_cap:][: 4 * ncap]]
    for q in range(Q):
        ci = int(idx[q]) >> (lde - shape.cap_height)
        for o in range(4):
            assert [int(roots[o * 4 + i, q]) for i in range(4)] == [int(x) for x in caps[o][4 * ci: 4 * ci + 4]]
