"""Differential fuzz of the two INDEPENDENT readers of the reference's wire format: the product's host parser
(csrc/host/parse.cpp: forward-scan fast path + tape reader) and the oracle's reader (oracle/json_reader.hpp, written from
Types.hs).  On every mutated proof text they must agree on accept-vs-reject of the text, and on every word when both accept.
Mutations cover what aeson tolerates (key order, whitespace, numbers >= p and >= 2^64 that `mkGoldilocks` reduces) and what it
does not (missing fields, wrong list lengths, fractions, truncation)."""
import json
import random

import numpy as np
import pytest

import fixtures
import oracle_lib

P = fixtures.P


def _shuffle_keys(o, rng):
    if isinstance(o, dict):
        items = [(k, _shuffle_keys(v, rng)) for k, v in o.items()]
        rng.shuffle(items)
        return dict(items)
    if isinstance(o, list):
        return [_shuffle_keys(v, rng) for v in o]
    return o


def _paths(o, pre=()):
    """paths to every leaf number"""
    if isinstance(o, dict):
        for k, v in o.items():
            yield from _paths(v, pre + (k,))
    elif isinstance(o, list):
        for i, v in enumerate(o):
            yield from _paths(v, pre + (i,))
    elif isinstance(o, int) and not isinstance(o, bool):
        yield pre


def _set(o, path, val):
    for k in path[:-1]:
        o = o[k]
    o[path[-1]] = val


def _get(o, path):
    for k in path:
        o = o[k]
    return o


@pytest.mark.parametrize("name", ["small6", "lookup6", "real5"])
def test_two_readers_agree_on_mutated_proofs(p2v, name):
    shape, lay, vkey, blob = fixtures.load(name)
    oc = oracle_lib.circuit_from_json(fixtures.read(name, "common"), fixtures.read(name, "vkey"))
    base = json.loads(fixtures.read(name, "proof"))
    leaves = list(_paths(base))
    rng = random.Random(1234)
    agree_ok = agree_bad = 0
    for trial in range(120):
        doc = json.loads(json.dumps(base))
        kind = trial % 8
        if kind == 0:      # key order + whitespace: same records
            text = json.dumps(_shuffle_keys(doc, rng), indent=rng.choice([None, 1, 3]))
        elif kind == 1:    # non-canonical field elements: reduced like mkGoldilocks
            for _ in range(5):
                pth = rng.choice(leaves)
                _set(doc, pth, _get(doc, pth) + P * rng.choice([1, 2, 5, 2**64, 10**12]))  # same residue: >= p, >= 2^64, >= 2^128
            text = json.dumps(doc)
        elif kind == 2:    # one value changed: both accept, same words, differs from the original
            pth = rng.choice(leaves)
            _set(doc, pth, rng.randrange(P))
            text = json.dumps(doc)
        elif kind == 3:    # a field missing
            pr = doc["proof"]
            victim = rng.choice(["wires_cap", "openings", "opening_proof"])
            del pr[victim]
            text = json.dumps(doc)
        elif kind == 4:    # a fraction where an integer belongs
            pth = rng.choice(leaves)
            text = json.dumps(doc).replace(str(_get(doc, pth)), str(_get(doc, pth)) + ".5", 1)
        elif kind == 5:    # truncated text
            t = json.dumps(doc)
            text = t[: rng.randrange(len(t) // 2, len(t) - 1)]
        elif kind == 6:    # extension element with three components
            doc["proof"]["openings"]["wires"][0] = [1, 2, 3]
            text = json.dumps(doc)
        else:              # digest with five elements
            doc["proof"]["wires_cap"][0]["elements"].append(7)
            text = json.dumps(doc)
        # product: fast path and tape reader
        got = []
        for env_slow in (False, True):
            try:
                got.append(p2v.parse_proof(text, shape))
            except p2v.P2VError:
                got.append(None)
            break  # the slow path is selected per process (P2V_NO_FAST_PARSE); tests/test_host.py covers fast == tape
        try:
            want = oc.proof_blob(text)
            if len(want) != lay.blob_words:
                want = None       # the oracle reader decodes any list lengths; the fixed-shape batch cannot hold it
        except ValueError:
            want = None
        if want is None:
            assert got[0] is None, "product accepted a text the oracle's reader rejects (mutation kind %d)" % kind
            agree_bad += 1
        else:
            assert got[0] is not None, "product rejected a text the oracle's reader accepts (mutation kind %d)" % kind
            assert np.array_equal(got[0], want), kind
            agree_ok += 1
            if kind in (0, 1):
                assert np.array_equal(got[0], blob)
    assert agree_ok >= 40 and agree_bad >= 40
