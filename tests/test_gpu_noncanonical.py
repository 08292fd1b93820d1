"""include/p2v.h: "Inputs may be any u64 (they are reduced mod p exactly like `mkGoldilocks`, Algebra/Goldilocks.hs:132)".
The JSON parsers canonicalise, but a caller may write blobs itself (the blob is the compact wire format), so the kernels
must give the SAME bits for every spelling of a field element.  Only values below 2^32 - 1 have a second spelling in 64
bits (v and v + p), and an honest proof holds none, so every named field of the tamper matrix in turn (and random words)
is overwritten with a small value: batch A spells it v, batch B spells it v + p.  Both batches are rejected proofs; every
intermediate of the two must be identical, and equal to the oracle's on the canonical spelling.  The same for the vkey."""
import numpy as np
import pytest

import fixtures

pytestmark = pytest.mark.gpu
P = fixtures.P


def _two_spellings(blob, lay, shape, seed):
    rng = np.random.default_rng(seed)
    table = sorted(fixtures.tamper_words(lay, shape).items())
    words = [w for _, w in table] + [int(w) for w in rng.integers(0, len(blob), 24)]
    n = len(words) + 1
    a = np.tile(np.asarray(blob, dtype=np.uint64), (n, 1))
    b = a.copy()
    for i, w in enumerate(words):
        v = int(rng.integers(0, 2**32 - 1)) if i % 3 else i % 2  # 0 and 1 among them: p and p + 1 are the classic traps
        a[i, w] = v
        b[i, w] = v + P
    # last proof: SEVERAL respelled words in one honest proof's zero-valued... there are none; respell three fields at once
    for w in words[:3]:
        a[n - 1, w] = 7
        b[n - 1, w] = 7 + P
    assert (b >= a).all() and (b != a).any()
    return a, b


@pytest.mark.parametrize("name", ["small6", "lookup6", "real5", "fixed4"])
def test_every_spelling_of_a_field_element_gives_the_same_bits(p2v, ctx, orc, name):
    shape, lay, vkey, blob = fixtures.load(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    a, b = _two_spellings(blob, lay, shape, seed=len(name))
    ga, gb = cir.verifyIntermediates(a), cir.verifyIntermediates(b)
    for key in ("challenges", "combined", "eqmask", "status", "accept", "qstatus", "roots"):
        assert np.array_equal(ga[key], gb[key]), key
    done = ((ga["qstatus"] == 0) | ((ga["qstatus"] & 0xFF) == 3)).reshape(-1)
    assert np.array_equal(ga["folded"][:, done], gb["folded"][:, done])
    want = orc.verify_batch(shape, vkey, a, threads=4, fast=True)
    assert np.array_equal(ga["status"], want["status"]) and np.array_equal(ga["challenges"], want["challenges"])
    assert np.array_equal(ga["combined"], want["combined"]) and np.array_equal(ga["qstatus"], want["qstatus"])
    # the plain entry point and the device-resident path see the same thing
    import torch

    acc_b, st_b = cir.verifyProof(b)
    assert np.array_equal(st_b, want["status"])
    d = torch.from_numpy(b.view(np.int64)).cuda()
    torch.cuda.synchronize()
    _, st_d = cir.verifyProof(d, n=len(b))
    assert np.array_equal(st_d, want["status"])
    cir.close()


def test_respelled_verifier_key(p2v, ctx, orc):
    """The verifier key (constants_sigmas_cap, circuit_digest) is a caller-written array too."""
    shape, lay, vkey, blob = fixtures.load("small6")
    vk_a = np.asarray(vkey, dtype=np.uint64).copy()
    vk_a[1] = 3          # cap word: the circuit no longer matches (every proof fails the constants tree), both spellings alike
    vk_a[len(vk_a) - 1] = 0  # circuit digest word: changes every challenge
    vk_b = vk_a.copy()
    vk_b[1] += np.uint64(P)
    vk_b[len(vk_b) - 1] += np.uint64(P)
    blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, 12, seed=2)
    ca, cb = p2v.Circuit(ctx, shape, vk_a), p2v.Circuit(ctx, shape, vk_b)
    ga, gb = ca.verifyIntermediates(blobs), cb.verifyIntermediates(blobs)
    for key in ("challenges", "combined", "eqmask", "status", "qstatus", "roots"):
        assert np.array_equal(ga[key], gb[key]), key
    want = orc.verify_batch(shape, vk_a, blobs, threads=2, fast=True)
    assert np.array_equal(ga["status"], want["status"]) and np.array_equal(ga["challenges"], want["challenges"])
    ca.close()
    cb.close()
