"""GPU parity: K0 + K4-K7 through the C ABI vs the CPU oracle on the bundled fixtures and on
tampered copies (bit-exact challenges, combined constraints, folded evaluations, verdict codes)."""
import numpy as np
import pytest

import fixtures

pytestmark = pytest.mark.gpu


def _circuit(p2v, ctx, name):
    shape, lay, vkey, blob = fixtures.load(name)
    return p2v.Circuit(ctx, shape, vkey), shape, lay, vkey, blob


@pytest.mark.parametrize("name", fixtures.ACCEPTING)
def test_fixture_accepts_and_intermediates_match(p2v, ctx, orc, name):
    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, name)
    blobs = blob.reshape(1, -1)
    want = orc.verify_batch(shape, vkey, blobs, threads=1, fast=False)
    assert want["status"][0] == 0, "oracle must accept the bundled fixture"
    ch = cir.proofChallenges(blobs)
    assert np.array_equal(ch, want["challenges"])
    comb, mask = cir.evalCombinedPlonkConstraints(blobs)
    assert np.array_equal(comb, want["combined"])
    assert mask[0] == want["eqmask"][0] == (1 << shape.num_challenges) - 1
    st, qs, folded = cir.checkFRIProof(blobs, want_debug=True)
    assert st[0] == 0 and (qs == 0).all()
    assert np.array_equal(folded, want["folded"])
    acc, status = cir.verifyProof(blobs)
    assert acc[0] and status[0] == 0


@pytest.mark.parametrize("name,code", [("small6_badfinal", 3), ("small6_badlayer0", 18), ("small6_badlayer1", 18 | (1 << 16)),
                                       ("real5_badwitness", 1 | (3 << 16)), ("real5_badcopy", 1 | (3 << 16)),
                                       ("reallu6_badlookup", 1 | (3 << 16))])
def test_regrinded_rejections(p2v, ctx, orc, name, code):
    """Proofs that reach the deep checks (need prover-side re-grinding): FALSE_FINAL, ERR_STEP_EVAL."""
    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, name)
    blobs = blob.reshape(1, -1)
    want = orc.verify_batch(shape, vkey, blobs, threads=1, fast=False)
    acc, status = cir.verifyProof(blobs)
    assert status[0] == want["status"][0] == code
    assert not acc[0]
    st, qs, folded = cir.checkFRIProof(blobs, want_debug=True)
    assert np.array_equal(qs, want["qstatus"])
    # the oracle raises at the first failing check of a round, so it has folded values only for
    # rounds that ran to the final-polynomial comparison
    done = ((want["qstatus"] == 0) | ((want["qstatus"] & 0xFF) == 3)).reshape(-1)
    assert np.array_equal(folded[:, done], want["folded"][:, done])


@pytest.mark.parametrize("name,n", [("small6", 96), ("fixed4", 80), ("lookup6", 96), ("mid5", 64), ("s12", 48),
                                    ("real5", 64), ("reallu6", 96), ("real12", 48), ("arity5", 64)])
def test_tamper_matrix(p2v, ctx, orc, name, n):
    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, name)
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, n, seed=7)
    want = orc.verify_batch(shape, vkey, blobs, threads=8, fast=True)
    assert (want["status"] != 0xEE).all()
    ch = cir.proofChallenges(blobs)
    assert np.array_equal(ch, want["challenges"])
    comb, mask = cir.evalCombinedPlonkConstraints(blobs)
    assert np.array_equal(comb, want["combined"])
    assert np.array_equal(mask, want["eqmask"])
    st, qs, folded = cir.checkFRIProof(blobs, want_debug=True)
    assert np.array_equal(qs, want["qstatus"])
    assert np.array_equal(st, want["fri_status"])
    # folded evaluations only compared where the oracle got that far (it raises at the first failure)
    ok_rounds = (want["qstatus"] == 0) | ((want["qstatus"] & 0xFF) == 3)
    f_g = folded.reshape(2, n, shape.num_queries)
    f_o = want["folded"].reshape(2, n, shape.num_queries)
    assert np.array_equal(f_g[:, ok_rounds], f_o[:, ok_rounds])
    acc, status = cir.verifyProof(blobs)
    assert np.array_equal(status, want["status"])
    assert np.array_equal(acc, want["status"] == 0)
    # untouched copies accept, every class of rejection shows up
    assert acc[words < 0].all()
    codes = set(int(s) & 0xFF for s in status)
    assert {0, 1, 2, 16, 17} <= codes, codes


def test_chunked_and_device_resident(p2v, ctx, orc):
    """Same verdicts when the batch is split into chunks and when it already lives in HBM."""
    import torch

    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, "small6")
    n = 200
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, n, seed=3)
    acc0, st0 = cir.verifyProof(blobs)
    ctx.set_chunk(64)
    acc1, st1 = cir.verifyProof(blobs)       # 4 chunks on the multi-lane pipeline (default depth 4)
    for depth in (1, 2, 3):                  # strictly serial, two lanes, three lanes
        ctx.set_pipeline(depth)
        acc3, st3 = cir.verifyProof(blobs)
        assert np.array_equal(st1, st3) and np.array_equal(acc1, acc3), depth
    ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
    ch_p = cir.proofChallenges(blobs)        # debug outputs gathered across pipelined chunks
    st_p, qs_p, fold_p = cir.checkFRIProof(blobs, want_debug=True)
    d_blobs = torch.from_numpy(blobs.view(np.int64)).cuda()
    d_out = torch.empty((n, lay.blob_words), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()  # torch's stream -> the context's own stream
    acc2, st2 = cir.verifyProof(d_blobs, n=n)
    ctx.set_chunk(0)
    st_1, qs_1, fold_1 = cir.checkFRIProof(blobs, want_debug=True)
    assert np.array_equal(ch_p, cir.proofChallenges(blobs))
    assert np.array_equal(st_p, st_1) and np.array_equal(qs_p, qs_1) and np.array_equal(fold_p, fold_1)
    assert np.array_equal(st0, st1) and np.array_equal(st0, st2)
    assert np.array_equal(acc0, acc1) and np.array_equal(acc0, acc2)
    # device-side synthesis gives the same batch
    cir.synth_batch(blob, n, words, deltas, d_out)
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), blobs)


@pytest.mark.parametrize("n", [0, 1, 31, 33])
def test_ragged_batch_sizes(p2v, ctx, orc, n):
    """Empty, single and non-multiple-of-32 batches: the packed accept bitmap has ceil(n/32) words, unused bits 0."""
    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, "real5")
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, max(n, 1), seed=9, accept_every=3)
    blobs = blobs[:n]
    bits = np.full((n + 31) // 32 + 1, 0xDEADBEEF, dtype=np.uint32)  # one guard word behind the bitmap
    status = np.empty(n, dtype=np.uint32)
    cir.verifyProof(blobs if n else np.zeros((1, lay.blob_words), dtype=np.uint64), n=n, accept_bits=bits, status=status)
    assert bits[-1] == 0xDEADBEEF
    if n == 0:
        return
    want = orc.verify_batch(shape, vkey, blobs, threads=2, fast=True)
    assert np.array_equal(status, want["status"])
    acc = p2v.unpack_bits(bits[:-1], n)
    assert np.array_equal(acc, want["status"] == 0)
    tail = int(bits[(n - 1) // 32]) >> (n % 32) if n % 32 else 0
    assert tail == 0


def test_host_transcript_first_windows(p2v, ctx, orc, monkeypatch):
    """Host input on the pipeline copies the per-proof parts of a whole window of chunks first (strided), then the query parts
    chunk by chunk.  Small windows (test hook P2V_PPFIRST_BYTES) force several of them, with a ragged last chunk; verdicts,
    challenges and per-query results must equal the serial single-chunk run and the oracle."""
    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, "real5")
    n = 64 * 11 + 19
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, n, seed=33, accept_every=3)
    ctx.set_pipeline(1)
    acc0, st0 = cir.verifyProof(blobs)
    ch0 = cir.proofChallenges(blobs)
    st_f0, qs0, fold0 = cir.checkFRIProof(blobs, want_debug=True)
    ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
    ctx.set_chunk(64)
    try:
        for win_chunks in (1, 3, 1000):
            monkeypatch.setenv("P2V_PPFIRST_BYTES", str(lay.proof_words * 8 * 64 * win_chunks))
            for depth in (2, 4):
                ctx.set_pipeline(depth)
                acc1, st1 = cir.verifyProof(blobs)
                assert np.array_equal(st0, st1) and np.array_equal(acc0, acc1), (win_chunks, depth)
            assert np.array_equal(ch0, cir.proofChallenges(blobs))
            st_f1, qs1, fold1 = cir.checkFRIProof(blobs, want_debug=True)
            assert np.array_equal(st_f0, st_f1) and np.array_equal(qs0, qs1) and np.array_equal(fold0, fold1)
    finally:
        ctx.set_chunk(0)
        ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
    want = orc.verify_batch(shape, vkey, blobs, threads=4, fast=True)
    assert np.array_equal(st0, want["status"])


def test_host_ramped_pipeline_equals_device_resident(p2v, ctx, orc):
    """Host input with chunks >= 8192 takes the ramped multi-lane pipeline (chunk/8 growing x9/8): every proof must
    land in exactly one chunk.  Compared with the device-resident run of the same batch and, on a sample, the oracle."""
    import torch

    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, "small6")
    n = 40000 + 17
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, n, seed=21, accept_every=5)
    ctx.set_chunk(8192)
    try:
        acc_h, st_h = cir.verifyProof(blobs)
        d_blobs = torch.from_numpy(blobs.view(np.int64)).cuda()
        torch.cuda.synchronize()
        acc_d, st_d = cir.verifyProof(d_blobs, n=n)
        st_d = np.asarray(st_d.cpu().numpy() if hasattr(st_d, "cpu") else st_d).view(np.uint32)
    finally:
        ctx.set_chunk(0)
    assert np.array_equal(st_h, st_d)
    sample = np.r_[0:64, 8192 - 32:8192 + 32, n - 64:n]
    want = orc.verify_batch(shape, vkey, blobs[sample], threads=4, fast=True)
    assert np.array_equal(st_h[sample], want["status"])
    assert acc_h[words < 0].all()


def test_json_in_verdict_out(p2v, ctx, orc):
    """The reference's testmain flow for a batch: JSON texts -> threaded decode into pinned memory -> GPU verdicts.
    A text that does not decode is reported and rejected without stopping the others."""
    import json

    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, "small6")
    good = fixtures.read("small6", "proof")
    texts = [good, fixtures.read("small6_badfinal", "proof"), good, "{ not json", fixtures.read("small6_badlayer1", "proof")]
    bad = json.loads(good)
    bad["proof"]["openings"]["wires"].pop()
    texts.append(json.dumps(bad))
    accept, status, rcs = cir.verifyProofJson(texts, threads=3)
    assert list(rcs) == [0, 0, 0, -4, 0, -5]
    assert list(accept) == [True, False, True, False, False, False]
    assert [int(s) for s in status[[0, 1, 2, 4]]] == [0, 3, 0, 18 | (1 << 16)]
    a0, s0, r0 = cir.verifyProofJson([])
    assert a0.size == 0 and s0.size == 0 and r0.size == 0


def test_heterogeneous_groups(p2v, ctx, orc):
    """Proofs of different circuits (other degree_bits, gate sets, reduction strategies, lookups) in one call: the chunks of
    all groups share the pipeline lanes; ten groups (some tiny, one empty) against the oracle."""
    groups, wants = [], []
    plan = (("small6", 40), ("fixed4", 33), ("reallu6", 21), ("real5", 17), ("arity5", 5), ("mid5", 1), ("lookup6", 64),
            ("small6", 0), ("real5", 3), ("fixed4", 96))
    for name, n in plan:
        cir, shape, lay, vkey, blob = _circuit(p2v, ctx, name)
        blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, max(n, 1), seed=5)
        blobs = blobs[:n]
        groups.append((cir, blobs))
        wants.append(orc.verify_batch(shape, vkey, blobs, threads=4, fast=True)["status"] if n else np.zeros(0, dtype=np.uint32))
    for depth in (p2v.DEFAULT_PIPELINE, 1):
        ctx.set_pipeline(depth)
        got = p2v.verify_groups(ctx, groups)
        for (acc, status), want in zip(got, wants):
            assert np.array_equal(status, want)
            assert np.array_equal(acc, want == 0)
    ctx.set_pipeline(p2v.DEFAULT_PIPELINE)


def test_a_group_that_cannot_run_does_not_stop_the_others(p2v, ctx, orc):
    """p2v_verify_groups keeps going past a bad group and reports per-group codes."""
    cir, shape, lay, vkey, blob = _circuit(p2v, ctx, "small6")
    other = p2v.Context(0)
    try:
        foreign = p2v.Circuit(other, shape, vkey)  # belongs to another context: invalid in this call
        blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, 20, seed=8)
        want = orc.verify_batch(shape, vkey, blobs, threads=4, fast=True)["status"]
        out, rcs = p2v.verify_groups(ctx, [(cir, blobs), (foreign, blobs), (cir, blobs[:7])], return_codes=True)
        assert list(rcs) == [0, -1, 0]
        assert np.array_equal(out[0][1], want) and np.array_equal(out[2][1], want[:7])
        assert not out[1][0].any()
        with pytest.raises(p2v.P2VError):
            p2v.verify_groups(ctx, [(cir, blobs), (foreign, blobs)])
        foreign.close()
    finally:
        other.close()


@pytest.mark.parametrize("gate", [
    "RandomAccessGate { bits: 7, num_copies: 1073741824, num_extra_constants: 0, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
    "RandomAccessGate { bits: 1, num_copies: 4294967297, num_extra_constants: 0, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
    "BaseSumGate { num_limbs: 3 } + Base: 1000000",
    "ArithmeticGate { num_ops: 4294967300 }",
    "ExponentiationGate { num_power_bits: 10 }",
])
def test_hostile_gate_parameters_are_refused(p2v, ctx, gate):
    """A circuit description whose gate parameters would wrap 32-bit index arithmetic (num_copies = 2^30 makes the wire count of a
    RandomAccessGate 130 * 2^30), truncate (2^32 + 1 read as 1), run a per-limb product over a huge base, or simply need more wires
    than the circuit has, is refused when the circuit is created — never handed to a kernel."""
    import json

    common = json.loads(fixtures.read("small6", "common"))
    common["gates"][2] = gate
    try:
        shape = p2v.parse_common(json.dumps(common))
    except p2v.P2VError as e:
        assert e.code in (-5, -6)
        return
    lay = p2v.shape_layout(shape)
    with pytest.raises(p2v.P2VError) as ei:
        p2v.Circuit(ctx, shape, np.zeros(lay.vkey_words, dtype=np.uint64))
    assert ei.value.code in (-5, -6), ei.value
