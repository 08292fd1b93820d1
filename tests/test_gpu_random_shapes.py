"""Property test over random circuit shapes on the GPU: for 30 random all-Noop shapes (gate set, selector groups, wires, number
of challenges, public inputs, quotient degree factor, rate, cap height, PoW bits, queries, ConstantArityBits / Fixed reduction
strategies incl. no folding step at all) and 6 real circuits under random FRI parameters, the verdicts of p2v_verify_batch on a
tampered batch and EVERY intermediate of p2v_verify_intermediates (challenges, combined constraint values, quotient-identity
verdicts, per-query status, folded evaluations, recomputed Merkle roots) must equal the CPU oracle's, bit for bit.  The oracle
reads the JSON with its own reader; the shapes come from oracle/p2v_prover at test time (tests/random_shapes.py)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
import random_shapes

pytestmark = pytest.mark.gpu
PRESETS = ["rand%d" % k for k in range(1, 31)] + ["rreal%d" % k for k in range(1, 7)]


@pytest.mark.parametrize("preset", PRESETS)
def test_gpu_equals_oracle_on_random_shape(p2v, ctx, tmp_path, preset):
    fx = random_shapes.generate(tmp_path, preset)
    shape = p2v.parse_common(fx["common"])
    lay = p2v.shape_layout(shape)
    vkey = p2v.parse_vkey(fx["vkey"], shape)
    blob = p2v.parse_proof(fx["proof"], shape)
    oc = oracle_lib.circuit_from_json(fx["common"], fx["vkey"])
    assert np.array_equal(oc.proof_blob(fx["proof"]), blob)
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 96
    blobs, words = random_shapes.tampered(blob, lay.proof_words, n, seed=11)
    want = oc.verify_batch(blobs, threads=8, fast=True)
    accept, status = cir.verifyProof(blobs)
    assert np.array_equal(status, want["status"]), [(int(w), hex(int(a)), hex(int(b))) for w, a, b in zip(words, status, want["status"]) if a != b][:8]
    assert np.array_equal(accept, want["status"] == 0) and accept[words < 0].all() and not accept[words >= 0].any()
    m = 32
    got = cir.verifyIntermediates(blobs[:m])
    assert np.array_equal(got["challenges"], want["challenges"][:, :m])
    assert np.array_equal(got["combined"], want["combined"][:, :m])
    assert np.array_equal(got["eqmask"], want["eqmask"][:m])
    assert np.array_equal(got["status"], want["status"][:m])
    assert np.array_equal(got["qstatus"], want["qstatus"][:m])
    # folded evaluations are defined where the reference gets that far (the oracle raises at the first failing check)
    Q = shape.num_queries
    for p in range(m):
        st = int(want["status"][p])
        if (st & 0xFF) in (0, 3):
            upto = Q if st == 0 else ((st >> 8) & 0xFF) + 1
            assert np.array_equal(got["folded"][:, p * Q: p * Q + upto], want["folded"][:, p * Q: p * Q + upto]), p
    assert np.array_equal(got["roots"], oc.fri_roots(blobs[:m]))
    cir.close()
    # the reference's driver on this shape (src/testmain.hs:40-63): the C++ host program above the C ABI must print what the
    # oracle's rendering of `testmain` prints (0 public inputs, 1 or 3 challenge rounds, no folding step ... included)
    exe = os.path.join(random_shapes.ROOT, "plonky2-verifier_b200", "p2v_testmain")
    r = subprocess.run([exe, str(tmp_path), preset], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == oc.testmain(fx["proof"])
