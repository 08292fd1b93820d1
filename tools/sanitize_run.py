"""Every kernel of libp2v.so once, on small and ragged inputs, for `compute-sanitizer` (tools/sanitize.sh).

No oracle here (parity is tests/' business): the point is that memcheck / initcheck / racecheck see every kernel with
batch sizes that are not multiples of a warp, a block or a bitmap word, with host AND device buffers, through the serial
mode, the chunk pipeline (several chunks per call) and the heterogeneous-group call.  Verdicts are still compared with
what the fixture's tamper schedule implies (accepting copies accept, tampered ones do not), so a run that "passes"
because nothing executed is not possible.   python tools/sanitize_run.py [quick]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import fixtures
import plonky2_verifier_b200 as p2v

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
rng = np.random.default_rng(7)
P = fixtures.P
ctx = p2v.Context(0)
launched = {}


def note(tag):
    launched[tag] = ctx.launch_count


# ---- hash API: K1, K2, k_compress, K3, tree builder / opener (host buffers) ----
for n in (1, 33, 257):
    st = rng.integers(0, P, size=(12, n), dtype=np.uint64)
    out = ctx.permutation(st)
    assert out.shape == (12, n)
    for w in (0, 1, 7, 8, 9, 20, 135):
        leaves = rng.integers(0, P, size=(w, n), dtype=np.uint64)
        d = ctx.sponge(leaves)
        assert d.shape == (4, n)
    l4, r4 = rng.integers(0, P, size=(4, n), dtype=np.uint64), rng.integers(0, P, size=(4, n), dtype=np.uint64)
    ctx.compress(l4, r4)
note("hash")
for (w, log_n, cap_h, n) in ((5, 6, 2, 37), (135, 8, 4, 100), (9, 3, 3, 5), (16, 4, 0, 19)):
    leaves = rng.integers(0, P, size=(w, 1 << log_n), dtype=np.uint64)
    dig = ctx.merkle_build(leaves, log_n, cap_h)
    idx = rng.integers(0, 1 << log_n, size=n, dtype=np.uint32)
    lo, so, cap = ctx.merkle_open(leaves, log_n, cap_h, dig, idx)
    ok, roots = ctx.checkMerkleProof(cap, idx, lo, so, want_roots=True)
    assert ok.all(), "honest Merkle openings must verify"
    if so.shape[0]:
        so2 = so.copy()
        so2[0, 0] = (int(so2[0, 0]) + 1) % P
        ok2 = ctx.checkMerkleProof(cap, idx, lo, so2)
        assert not ok2[0] and ok2[1:].all()
note("merkle")

# ---- field-op hook ----
edge = np.array([0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32 + 1], dtype=np.uint64)
a, b = np.repeat(edge, len(edge)), np.tile(edge, len(edge))
for op in (0, 1, 2, 3, 4, 5, 6, 8, 9):
    ctx.field_op(op, a, b)
ctx.field_op(7, a, b & np.uint64(0xFFFF))
a2, b2 = np.stack([a, b]), np.stack([b, a])
for op in (16, 17, 18, 19, 20, 21, 22, 24):
    ctx.field_op(op, a2, b2)
note("field")

# ---- the verifier on every bundled shape: host buffers, device buffers, serial mode and multi-chunk pipeline ----
names = ["small6", "fixed4", "lookup6", "real5", "reallu6", "arity5", "mid5"] + ([] if quick else ["real12"])
circuits = {}
for name in names:
    shape, lay, vkey, blob = fixtures.load(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    circuits[name] = (cir, shape, lay, blob)
    for n in ((1, 33) if name != "real12" else (35,)):
        blobs, words, _ = fixtures.tampered_batch(blob, lay, shape, n, seed=11 + n)
        honest = words < 0
        for depth, chunk in ((1, 0), (4, 8)):
            ctx.set_pipeline(depth)
            if chunk:
                ctx.set_chunk(chunk)
            acc, status = cir.verifyProof(blobs)  # host buffers (pageable numpy memory: staged by the library)
            assert acc[honest].all() and (status[honest] == 0).all(), (name, n, depth)
            assert honest.all() or not acc[~honest].all(), (name, n, depth)
            d_blobs = torch.from_numpy(blobs.view(np.int64)).cuda()
            d_bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
            d_status = torch.full((n,), -1, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            cir.verifyProof(d_blobs, n=n, accept_bits=d_bits, status=d_status)
            ctx.sync()
            assert np.array_equal(d_status.cpu().numpy().view(np.uint32), status), (name, n, depth)
        ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
        ctx.set_chunk(0)
        got = cir.verifyIntermediates(blobs)
        assert np.array_equal(got["status"], status), (name, n)
        ch = cir.proofChallenges(blobs)
        assert np.array_equal(ch, got["challenges"])
        cir.evalCombinedPlonkConstraints(blobs)
        cir.checkFRIProof(blobs, want_debug=True)
        cir.stage(blobs)
    note(name)

# ---- heterogeneous groups through the shared lanes ----
groups = []
for name in ("small6", "lookup6", "real5", "fixed4", "small6"):
    cir, shape, lay, blob = circuits[name]
    blobs, words, _ = fixtures.tampered_batch(blob, lay, shape, 9 + len(groups) * 7, seed=3 + len(groups))
    groups.append((cir, blobs, words))
out = p2v.verify_groups(ctx, [(c, b) for c, b, _ in groups])
for (acc, status), (_, _, words) in zip(out, groups):
    assert acc[words < 0].all() and not acc[words >= 0].all()
note("groups")

# ---- JSON in, verdict out ----
cir, shape, lay, blob = circuits["small6"]
acc, _, rcs = cir.verifyProofJson([fixtures.read("small6", "proof"), fixtures.read("small6_badfinal", "proof")] * 3)
assert not rcs.any()
assert list(np.asarray(acc).astype(bool)) == [True, False] * 3
note("json")

for c, _, _, _ in circuits.values():
    c.close()
ctx.close()
print("sanitize_run ok:", launched)
