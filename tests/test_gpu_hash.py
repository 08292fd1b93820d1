"""GPU parity: K1-K3 through the C ABI vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest

from oracle_lib import P, rand_felts

pytestmark = pytest.mark.gpu

KAT = [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2,
       0x414754bfebd051f0, 0xcb1f8980294a023f, 0x6eb2a9e4d54a9d0f, 0x1902bc3af467e056,
       0xf045d5eafdc6021f, 0xe4150f77caaa3be5, 0xc9bfd01d39b50cce, 0x5c0a27fcb0e1459b]


def test_poseidon_kat(ctx):
    """Reference KAT, Hash/Poseidon.hs:27-35: permutation [0..11]."""
    s = np.arange(12, dtype=np.uint64).reshape(12, 1)
    out = ctx.permutation(s)
    assert [int(x) for x in out[:, 0]] == KAT


def test_poseidon_random_vs_oracle(ctx, orc):
    rng = np.random.default_rng(1)
    for n in (1, 31, 33, 1000, 20000):
        s = rand_felts(rng, (12, n))
        got = ctx.permutation(s)
        want = orc.permutation(s, which=0 if n <= 1000 else 2)
        assert np.array_equal(got, want), n
        assert (got < np.uint64(P)).all()


def test_poseidon_edge_states(ctx, orc):
    """All-equal edge states: 0, p-1, p, 2^64-1 (non-canonical inputs are reduced like mkGoldilocks)."""
    vals = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 2**63]
    s = np.array([[v] * 12 for v in vals], dtype=np.uint64).T.copy()
    assert np.array_equal(ctx.permutation(s), orc.permutation(s, which=0))


@pytest.mark.parametrize("w", [0, 1, 4, 7, 8, 9, 16, 20, 85, 135])
def test_sponge_widths(ctx, orc, w):
    rng = np.random.default_rng(100 + w)
    n = 257
    leaves = rand_felts(rng, (w, n))
    assert np.array_equal(ctx.sponge(leaves), orc.sponge(leaves))


def test_compress(ctx, orc):
    rng = np.random.default_rng(7)
    a, b = rand_felts(rng, (4, 500)), rand_felts(rng, (4, 500))
    assert np.array_equal(ctx.compress(a, b), orc.compress(a, b))


@pytest.mark.parametrize("w,log_n,cap_height", [(5, 6, 0), (8, 8, 4), (135, 10, 4), (3, 4, 4), (20, 5, 2)])
def test_merkle_build_open_verify(ctx, orc, w, log_n, cap_height):
    rng = np.random.default_rng(log_n * 100 + w)
    nl = 1 << log_n
    leaves = rand_felts(rng, (w, nl))
    digests = ctx.merkle_build(leaves, log_n, cap_height)
    # level 0 of the tree = leaf sponges
    assert np.array_equal(digests[: 4 * nl].reshape(4, nl), orc.sponge(leaves))
    n = 300
    idx = rng.integers(0, nl, size=n, dtype=np.uint32)
    lo, so, cap = ctx.merkle_open(leaves, log_n, cap_height, digests, idx)
    assert np.array_equal(lo, leaves[:, idx])
    ok, roots = ctx.checkMerkleProof(cap, idx, lo, so, want_roots=True)
    ok_o, roots_o = orc.checkMerkleProof(cap, idx, lo, so)
    assert ok.all() and (ok_o == 1).all()
    assert np.array_equal(roots, roots_o)
    # tamper: one sibling word, one leaf word, one cap word
    so2, lo2 = so.copy(), lo.copy()
    bad = np.arange(0, n, 3)
    if so2.shape[0]:
        so2[rng.integers(0, so2.shape[0]), bad] += np.uint64(1)
    else:
        lo2[0, bad] += np.uint64(1)
    lo2[rng.integers(0, w), bad[::2]] += np.uint64(1)
    ok2 = ctx.checkMerkleProof(cap, idx, lo2, so2)
    ok2_o, _ = orc.checkMerkleProof(cap, idx, lo2, so2)
    assert np.array_equal(ok2, ok2_o == 1)
    assert not ok2[bad].any() and ok2[np.setdiff1d(np.arange(n), bad)].all()


def test_merkle_device_pointers(ctx, orc):
    """Same call with device-resident buffers (torch CUDA tensors)."""
    import torch

    rng = np.random.default_rng(5)
    s = rand_felts(rng, (12, 4096))
    d_in = torch.from_numpy(s.view(np.int64)).cuda()
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()  # the H2D copy ran on torch's stream, the context has its own
    ctx.permutation(d_in, out=d_out)
    ctx.sync()
    got = d_out.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, orc.permutation(s, which=2))
