"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle only checks a
sample): config 2 (2^20 Merkle openings, tree height 20, cap height 4), config 3 (FRI check on 10^4
copies/tamperings of the S12 fixture), config 4 (full verifier on a large S12 batch)."""
import numpy as np
import pytest

import fixtures
from oracle_lib import P

pytestmark = pytest.mark.gpu


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


@pytest.mark.parametrize("width", [8, 135])
def test_config2_merkle_2pow20_openings(ctx, orc, width):
    """2^20 openings against ONE synthetic tree of 2^20 leaves, 16 siblings per path, 1/64 tampered:
    build -> open -> verify is the identity on untampered openings, every tampered opening is rejected,
    and a random sample of paths is re-checked by the oracle."""
    import torch

    log_n, cap_height, n = 20, 4, 1 << 20
    if width == 135:
        log_n, n = 20, 1 << 20
    nl = 1 << log_n
    with np.errstate(over="ignore"):
        idx64 = np.arange(nl, dtype=np.uint64)
        leaves = torch.empty((width, nl), dtype=torch.int64, device="cuda")
        for j in range(width):  # leaf(i, j) = splitmix64(seed ^ (i*width + j)) mod p, generated plane by plane
            v = splitmix64(np.uint64(0x9E3779B97F4A7C15) ^ (idx64 * np.uint64(width) + np.uint64(j))) % np.uint64(P)
            leaves[j] = torch.from_numpy(v.view(np.int64)).cuda()
        q = (splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(1)) % np.uint64(nl)).astype(np.uint32)
    total = 4 * ((2 << log_n) - (1 << cap_height))
    digests = torch.empty(total, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()  # torch filled the planes on ITS stream; the context has its own (non-blocking) stream
    ctx.merkle_build(leaves, log_n, cap_height, out=digests)
    d_idx = torch.from_numpy(q.view(np.int32)).cuda()
    lo = torch.empty((width, n), dtype=torch.int64, device="cuda")
    so = torch.empty(((log_n - cap_height) * 4, n), dtype=torch.int64, device="cuda")
    cap = torch.empty((1 << cap_height, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.merkle_open(leaves, log_n, cap_height, digests, d_idx, leaves_out=lo, sibs_out=so, cap_out=cap)
    ctx.sync()
    # tamper every 64th opening: one sibling word + 1
    bad = torch.arange(0, n, 64, device="cuda")
    so[5, bad] += 1
    bits = torch.zeros(n // 32, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.checkMerkleProof(cap, d_idx, lo, so, ok_bits=bits)
    ctx.sync()
    ok = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little").astype(bool)
    want = np.ones(n, dtype=bool)
    want[::64] = False
    assert np.array_equal(ok, want)
    # oracle on a sample (incl. tampered ones)
    sample = np.concatenate([np.arange(0, 256), np.random.default_rng(0).integers(0, n, 256)])
    s_t = torch.from_numpy(sample).cuda()
    ok_o, _ = orc.checkMerkleProof(cap.cpu().numpy().view(np.uint64), q[sample], lo[:, s_t].cpu().numpy().view(np.uint64),
                                   so[:, s_t].cpu().numpy().view(np.uint64))
    assert np.array_equal(ok_o == 1, want[sample])


def _device_batch(p2v, ctx, cir, blob, lay, shape, n, seed):
    import torch

    sched, words, deltas = fixtures.tampered_batch(blob, lay, shape, min(n, 2048), seed=seed)
    reps = (n + len(words) - 1) // len(words)
    words_n, deltas_n = np.tile(words, reps)[:n].copy(), np.tile(deltas, reps)[:n].copy()
    d_blobs = torch.empty((n, lay.blob_words), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    cir.synth_batch(blob, n, words_n, deltas_n, d_blobs)
    return d_blobs, sched, words_n


def test_config3_fri_batch_10k(p2v, ctx, orc):
    """FRI check alone on 10^4 copies/tamperings of the S12 fixture: the schedule has period 2048, so the statuses
    must be periodic; the first period is compared with the oracle on a sample; untouched copies pass."""
    shape, lay, vkey, blob = fixtures.load("s12")
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 10000
    d_blobs, sched, words_n = _device_batch(p2v, ctx, cir, blob, lay, shape, n, seed=21)
    status = cir.checkFRIProof(d_blobs, n=n)
    assert (status[words_n < 0] == 0).all()
    assert np.array_equal(status[2048:4096], status[:2048]) and np.array_equal(status[8192:], status[: n - 8192])
    sample = np.arange(0, 96)
    want = orc.verify_batch(shape, vkey, sched[sample], threads=8, fast=True)
    assert np.array_equal(status[sample], want["fri_status"])
    codes = set(int(s) & 0xFF for s in status)
    assert {0, 2, 16, 17} <= codes


@pytest.mark.parametrize("name", ["s12", "real12"])
def test_config4_full_verifier_large_batch(p2v, ctx, orc, name):
    """Full verifier on 20 000 standard-recursion-configuration proofs resident in HBM (the all-Noop fixture and the one
    with real rows), in one chunk and in 3 chunks: identical verdicts, accept bit == (status == 0), periodic in the
    tamper schedule, sample checked by the oracle."""
    import torch

    shape, lay, vkey, blob = fixtures.load(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 20000
    d_blobs, sched, words_n = _device_batch(p2v, ctx, cir, blob, lay, shape, n, seed=33)
    bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    st2 = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    cir.verifyProof(d_blobs, n=n, accept_bits=bits, status=st)
    ctx.sync()
    status = st.cpu().numpy().view(np.uint32)
    accept = p2v.unpack_bits(bits.cpu().numpy().view(np.uint32), n)
    assert np.array_equal(accept, status == 0)
    assert accept[words_n < 0].all() and accept.sum() == (words_n < 0).sum()
    assert np.array_equal(status[2048:4096], status[:2048])
    ctx.set_chunk(8192)
    cir.verifyProof(d_blobs, n=n, accept_bits=bits, status=st2)
    ctx.sync()
    ctx.set_chunk(0)
    assert np.array_equal(st2.cpu().numpy().view(np.uint32), status)
    want = orc.verify_batch(shape, vkey, sched[:64], threads=8, fast=True)
    assert np.array_equal(status[:64], want["status"])
