"""Bundled JSON fixtures (tests/golden, made by tests/golden/make_fixtures.sh) -> (shape, vkey, blob),
plus the word-level tamper schedule of SURVEY.md App. E.  The JSON is parsed by the PRODUCT's host
parser (libp2v, no GPU needed); tests/test_host.py checks that parser against Python's json module."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P = 0xFFFFFFFF00000001
ACCEPTING = ["s12", "mid5", "small6", "fixed4", "lookup6", "real5", "real12", "reallu6", "arity5"]
REJECTING = {"small6_badfinal": "small6", "small6_badlayer0": "small6", "small6_badlayer1": "small6",
             "real5_badwitness": "real5", "real5_badcopy": "real5", "reallu6_badlookup": "reallu6"}


def read(name, kind):
    with open(os.path.join(GOLDEN, "%s_%s.json" % (name, kind))) as fh:
        return fh.read()


_cache = {}


def load(name):
    """-> (Shape, Layout, vkey u64[vkey_words], blob u64[blob_words])"""
    if name in _cache:
        return _cache[name]
    import plonky2_verifier_b200 as p2v

    common = REJECTING.get(name, name)
    shape = p2v.parse_common(read(common, "common"))
    lay = p2v.shape_layout(shape)
    vkey = p2v.parse_vkey(read(name, "vkey"), shape)
    blob = p2v.parse_proof(read(name, "proof"), shape)
    _cache[name] = (shape, lay, vkey, blob)
    return _cache[name]


def tamper_words(lay, shape):
    """Named word offsets for the tamper matrix (one representative word per proof field)."""
    q0 = lay.proof_words  # query 0
    last_q = lay.proof_words + (shape.num_queries - 1) * lay.query_words
    t = {
        "wires_cap": lay.off_wires_cap + 5,
        "zs_pp_cap": lay.off_zs_pp_cap + 1,
        "quotient_cap": lay.off_quotient_cap + 2,
        "open_constant": lay.off_open_constants + 1,
        "open_sigma": lay.off_open_sigmas + 3,
        "open_wire_routed": lay.off_open_wires + 2,
        "open_wire_advice": lay.off_open_wires + 2 * (shape.num_wires - 1),
        "open_zs": lay.off_open_zs,
        "open_zs_next": lay.off_open_zs_next + 1,
        "open_pp": lay.off_open_pp + 1,
        "open_quotient": lay.off_open_quotient + 3,
        "commit_cap": lay.off_commit_caps + 2,
        "final_poly": lay.off_final_poly + 1,
        "pow_witness": lay.off_pow_witness,
        "q0_leaf0": q0 + lay.q_off_leaf[0] + 1,
        "q0_leaf1": q0 + lay.q_off_leaf[1] + lay.oracle_width[1] - 1,
        "q0_leaf2": q0 + lay.q_off_leaf[2],
        "q0_leaf3": q0 + lay.q_off_leaf[3] + 1,
        "q0_sib1": q0 + lay.q_off_sibs[1] + 2 if lay.init_path_len else q0 + lay.q_off_leaf[1],
        "q0_sib3_last": q0 + lay.q_off_sibs[3] + 4 * lay.init_path_len - 1 if lay.init_path_len else q0 + lay.q_off_leaf[3],
        "q0_step0_eval": q0 + lay.q_off_step_evals[0] + 3,
        "qlast_step_last_eval": last_q + lay.q_off_step_evals[shape.num_steps - 1],
        "qlast_leaf1": last_q + lay.q_off_leaf[1] + 7,
    }
    if lay.step_path_len[0] > 0:
        t["q0_step0_sib"] = q0 + lay.q_off_step_sibs[0] + 1
    if shape.num_public_inputs > 0:
        t["public_input"] = lay.off_public_inputs
    if shape.num_lookup_polys > 0:
        t["open_lookup_zs"] = lay.off_open_lookup_zs + 2
        t["open_lookup_zs_next"] = lay.off_open_lookup_zs_next + 1
    return t


def tamper_table(name):
    """The tamper matrix of a fixture as committed data (tests/golden/<name>_tamper.json, written by
    tests/golden/make_tamper_tables.py from `tamper_words`): lets a process that must not load the product (bench.py's
    reference arm) build the SAME seeded batch."""
    import json

    with open(os.path.join(GOLDEN, "%s_tamper.json" % name)) as fh:
        return json.load(fh)


def tampered_batch_from_table(blob, table, n, seed=0, accept_every=4):
    """n copies of `blob`; copy i is left intact iff i % accept_every == 0, else one word (cycled through the tamper
    matrix `table` = {field name: word offset}, then random words) gets +delta mod p.
    Returns (blobs [n][W], tamper_word int32[n], delta u64[n])."""
    blob = np.asarray(blob, dtype=np.uint64)
    rng = np.random.default_rng(seed)
    names = sorted(table.items())
    words = np.full(n, -1, dtype=np.int32)
    deltas = np.zeros(n, dtype=np.uint64)
    k = 0
    for i in range(n):
        if i % accept_every == 0:
            continue
        if k < len(names):
            words[i] = names[k][1]
            deltas[i] = 1
        else:
            words[i] = rng.integers(0, len(blob))
            deltas[i] = rng.integers(1, P, dtype=np.uint64)
        k += 1
    blobs = np.tile(blob, (n, 1))
    for i in range(n):
        if words[i] >= 0:
            v = (int(blobs[i, words[i]]) % P + int(deltas[i])) % P
            blobs[i, words[i]] = v
    return blobs, words, deltas


def tampered_batch(blob, lay, shape, n, seed=0, accept_every=4):
    """`tampered_batch_from_table` with the tamper matrix computed from the product's layout."""
    return tampered_batch_from_table(blob, tamper_words(lay, shape), n, seed, accept_every)
