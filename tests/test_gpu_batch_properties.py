"""Size-independent properties of a batch call (Plonk/Verifier.hs:56-65 mapped over a list is what the call replaces, so a
verdict may depend on NOTHING but its own proof): a proof's status is the same wherever it stands in the batch, whatever
its neighbours are, however the batch is cut into chunks and lanes, and whether the batch is in host or device memory."""
import numpy as np
import pytest

import fixtures

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["real5", "lookup6"])
def test_verdicts_follow_their_proofs(p2v, ctx, orc, name):
    import torch

    shape, lay, vkey, blob = fixtures.load(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 149
    blobs, words, _ = fixtures.tampered_batch(blob, lay, shape, n, seed=17)
    want = orc.verify_batch(shape, vkey, blobs, threads=8, fast=True)["status"]
    rng = np.random.default_rng(5)
    try:
        for depth, chunk in ((4, 0), (1, 0), (4, 32), (3, 64), (2, 32)):
            ctx.set_pipeline(depth)
            ctx.set_chunk(chunk)
            perm = rng.permutation(n)
            acc, st = cir.verifyProof(np.ascontiguousarray(blobs[perm]))
            assert np.array_equal(st, want[perm]), (depth, chunk)
            assert np.array_equal(acc, want[perm] == 0)
            # any sub-batch, in any order, with repetitions
            pick = rng.integers(0, n, size=int(rng.integers(1, 100)))
            d = torch.from_numpy(np.ascontiguousarray(blobs[pick]).view(np.int64)).cuda()
            torch.cuda.synchronize()
            _, st2 = cir.verifyProof(d, n=len(pick))
            assert np.array_equal(st2, want[pick]), (depth, chunk, len(pick))
            got = cir.verifyIntermediates(np.ascontiguousarray(blobs[pick]))
            assert np.array_equal(got["status"], want[pick])
    finally:
        ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
        ctx.set_chunk(0)
        cir.close()


def test_two_contexts_do_not_share_state(p2v, ctx, orc):
    """Different contexts may be used side by side (include/p2v.h): interleaved calls on two contexts, each with its own circuit
    and chunking, give what each gives alone."""
    shape_a, lay_a, vkey_a, blob_a = fixtures.load("small6")
    shape_b, lay_b, vkey_b, blob_b = fixtures.load("fixed4")
    other = p2v.Context(0)
    try:
        ca, cb = p2v.Circuit(ctx, shape_a, vkey_a), p2v.Circuit(other, shape_b, vkey_b)
        ba, _, _ = fixtures.tampered_batch(blob_a, lay_a, shape_a, 70, seed=1)
        bb, _, _ = fixtures.tampered_batch(blob_b, lay_b, shape_b, 45, seed=2)
        wa = orc.verify_batch(shape_a, vkey_a, ba, threads=4, fast=True)["status"]
        wb = orc.verify_batch(shape_b, vkey_b, bb, threads=4, fast=True)["status"]
        other.set_chunk(32)
        for _ in range(3):
            _, sa = ca.verifyProof(ba)
            _, sb = cb.verifyProof(bb)
            assert np.array_equal(sa, wa) and np.array_equal(sb, wb)
        ca.close()
        cb.close()
    finally:
        other.close()


def test_two_contexts_from_two_threads(p2v, ctx, orc):
    """"different contexts may be used from different threads" (include/p2v.h): two host threads, one context each, hammering
    the verifier at the same time (ctypes drops the GIL during the calls) — every call must give the oracle's verdicts."""
    import threading

    shape, lay, vkey, blob = fixtures.load("real5")
    blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, 90, seed=4)
    want = orc.verify_batch(shape, vkey, blobs, threads=8, fast=True)["status"]
    errors = []

    def worker(k):
        try:
            c = p2v.Context(0)
            c.set_chunk(32 if k else 0)
            cir = p2v.Circuit(c, shape, vkey)
            for it in range(6):
                sub = blobs[k::2] if it % 2 else blobs
                exp = want[k::2] if it % 2 else want
                _, st = cir.verifyProof(np.ascontiguousarray(sub))
                if not np.array_equal(st, exp):
                    errors.append((k, it, "status differs"))
            cir.close()
            c.close()
        except Exception as e:  # noqa: BLE001 - reported to the main thread
            errors.append((k, repr(e)))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "a worker thread hangs"
    assert not errors, errors
