"""Property test over random circuit shapes (CPU): for every shape the product's host parser (p2v_parse_*) and the oracle's own
JSON reader must produce the same blob, layout sizes and verifier key; the C++ oracle must accept the prover's proof and tell
tampered copies apart; and on the small shapes the pure-Python twin (oracle/pyref.py, which reads the JSON itself) must agree with
the C++ oracle on verdict, challenges and combined constraint values.  The GPU half is tests/test_gpu_random_shapes.py."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib
import random_shapes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyref  # noqa: E402

PRESETS = ["rand%d" % k for k in range(1, 17)] + ["rreal1", "rreal2"]


@pytest.mark.parametrize("preset", PRESETS)
def test_readers_agree_and_oracle_accepts(p2v, tmp_path, preset):
    fx = random_shapes.generate(tmp_path, preset)
    shape = p2v.parse_common(fx["common"])
    lay = p2v.shape_layout(shape)
    vkey = p2v.parse_vkey(fx["vkey"], shape)
    blob = p2v.parse_proof(fx["proof"], shape)
    oc = oracle_lib.circuit_from_json(fx["common"], fx["vkey"])
    assert np.array_equal(oc.proof_blob(fx["proof"]), blob), "the two JSON readers disagree on the proof"
    assert np.array_equal(oc.vkey_words(), vkey)
    common = json.loads(fx["common"])
    assert shape.num_queries == common["config"]["fri_config"]["num_query_rounds"] == oc.num_queries
    assert shape.num_steps == oc.num_steps and shape.num_challenges == oc.num_challenges
    assert p2v.challenges_words(shape) == oc.challenges_words()
    assert lay.blob_words == len(blob) == lay.proof_words + shape.num_queries * lay.query_words
    blobs, words = random_shapes.tampered(blob, lay.proof_words, 24, seed=3)
    res = oc.verify_batch(blobs, threads=4, fast=True)
    assert (res["status"][words < 0] == 0).all(), "intact copies must be accepted"
    assert (res["status"][words >= 0] != 0).all(), "a changed word must be rejected"
    slow = oc.verify_batch(blobs[:6], threads=4, fast=False)  # the literal restatement (foldCosetWith with its inversions)
    for k in ("status", "challenges", "combined", "folded"):
        assert np.array_equal(slow[k], res[k][..., : slow[k].shape[-1]] if k != "status" else res[k][:6]), k


@pytest.mark.parametrize("preset", ["rand4", "rand8", "rand9", "rand12"])
def test_python_twin_agrees(tmp_path, preset):
    fx = random_shapes.generate(tmp_path, preset)
    oc = oracle_lib.circuit_from_json(fx["common"], fx["vkey"])
    res = oc.verify_batch(oc.proof_blob(fx["proof"]), threads=1, fast=False)
    tr = {}
    st = pyref.verify(json.loads(fx["common"]), json.loads(fx["vkey"]), json.loads(fx["proof"]), tr)
    assert st == int(res["status"][0]) == 0
    assert tr["challenges"] == [int(x) for x in res["challenges"][:, 0]]
    assert tr["combined"] == [int(x) for x in res["combined"][:, 0]]
