"""Guard-band tests: the pool's GPU boxes do not allow compute-sanitizer, so out-of-bounds accesses of the kernels and
of the library's copies are looked for the slow way.  Every caller-visible buffer sits in the middle of a larger
allocation whose margins carry a canary:
  * a write outside the buffer changes a canary word (checked after every call, device and host buffers alike);
  * a read outside an input that reaches a result makes the result depend on the canary, so every call runs with two
    different canaries and must produce identical bits — which must also equal the result of the plain host-buffer call.
Batch sizes are ragged on purpose (1, 33, 70: not a warp, not a bitmap word, not a transpose tile, several chunks with a
short last one)."""
import numpy as np
import pytest

import fixtures

pytestmark = pytest.mark.gpu

PAD = 4096  # elements of margin on either side
CANARIES = (0x5A5A5A5A5A5A5A5A, 0xFFFFFFFF00000000)  # the second one is p - 1: a valid field element, the worst case for a silent read


def _dev(n, dtype, canary):
    """-> (whole allocation, the n-element window in its middle), margins filled with `canary` (truncated to dtype)."""
    import torch

    item = torch.empty(0, dtype=dtype).element_size()
    val = canary >> (64 - 8 * item)  # the TOP bytes: 0xFFFFFFFF for the second canary in a 32-bit buffer, not 0
    if val >= 1 << (8 * item - 1):
        val -= 1 << (8 * item)
    whole = torch.full((PAD + n + PAD,), val, dtype=dtype, device="cuda")
    return whole, whole[PAD:PAD + n]


def _intact(whole, n, canary):
    import torch

    item = whole.element_size()
    val = canary >> (64 - 8 * item)
    if val >= 1 << (8 * item - 1):
        val -= 1 << (8 * item)
    return bool((whole[:PAD] == val).all().item()) and bool((whole[PAD + n:] == val).all().item())


def _host(n, dtype, canary):
    whole = np.full(PAD + n + PAD, canary >> (64 - 8 * np.dtype(dtype).itemsize), dtype=dtype)
    return whole, whole[PAD:PAD + n]


def _host_intact(whole, n, canary):
    val = canary >> (64 - 8 * whole.dtype.itemsize)
    return bool((whole[:PAD] == val).all()) and bool((whole[PAD + n:] == val).all())


@pytest.mark.parametrize("name", ["small6", "lookup6", "real5", "arity5", "fixed4"])
@pytest.mark.parametrize("n", [1, 33, 70])
def test_verifier_stays_inside_its_buffers(p2v, ctx, name, n):
    import torch

    shape, lay, vkey, blob = fixtures.load(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    blobs, words, _ = fixtures.tampered_batch(blob, lay, shape, n, seed=21 + n)
    ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
    ctx.set_chunk(0)
    acc0, status0 = cir.verifyProof(blobs)  # plain host call: the result every guarded variant must reproduce
    assert acc0[words < 0].all() and (n == 1 or not acc0.all())
    W = lay.blob_words
    nb = (n + 31) // 32
    try:
        for depth, chunk in ((1, 0), (4, 16)):
            ctx.set_pipeline(depth)
            ctx.set_chunk(chunk)
            for canary in CANARIES:
                # device buffers
                wb, b = _dev(n * W, torch.int64, canary)
                b.copy_(torch.from_numpy(blobs.reshape(-1).view(np.int64)))
                wbits, bits = _dev(nb, torch.int32, canary)
                wst, st = _dev(n, torch.int32, canary)
                bits.zero_()
                st.fill_(-1)
                torch.cuda.synchronize()
                cir.verifyProof(b, n=n, accept_bits=bits, status=st)
                ctx.sync()
                assert np.array_equal(st.cpu().numpy().view(np.uint32), status0), (depth, hex(canary))
                assert np.array_equal(p2v.unpack_bits(bits.cpu().numpy().view(np.uint32), n), acc0)
                assert _intact(wb, n * W, canary) and _intact(wbits, nb, canary) and _intact(wst, n, canary), "write outside a device buffer"
                # host buffers (the library stages them itself and copies the results back)
                hwb, hb = _host(n * W, np.uint64, canary)
                hb[:] = blobs.reshape(-1)
                hwbits, hbits = _host(nb, np.uint32, canary)
                hwst, hst = _host(n, np.uint32, canary)
                hbits[:] = 0
                cir.verifyProof(hb, n=n, accept_bits=hbits, status=hst)
                ctx.sync()
                assert np.array_equal(hst, status0) and np.array_equal(p2v.unpack_bits(hbits, n), acc0)
                assert _host_intact(hwb, n * W, canary) and _host_intact(hwbits, nb, canary) and _host_intact(hwst, n, canary), "write outside a host buffer"
    finally:
        ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
        ctx.set_chunk(0)
        cir.close()


@pytest.mark.parametrize("name", ["small6", "reallu6"])
def test_intermediate_entry_points_stay_inside_their_buffers(p2v, ctx, name):
    """p2v_challenges and p2v_stage write caller-sized planes ([words][n]); K0 alone moves the whole blob."""
    shape, lay, vkey, blob = fixtures.load(name)
    cir = p2v.Circuit(ctx, shape, vkey)
    n = 37
    blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, n, seed=5)
    cw = p2v.challenges_words(shape)
    ch0 = cir.proofChallenges(blobs)
    planes0 = cir.stage(blobs)
    assert np.array_equal(planes0[:lay.proof_words], blobs[:, :lay.proof_words].T)
    for canary in CANARIES:
        hwb, hb = _host(n * lay.blob_words, np.uint64, canary)
        hb[:] = blobs.reshape(-1)
        wch, ch = _host(cw * n, np.uint64, canary)
        cir.proofChallenges(hb.reshape(n, -1), out=ch.reshape(cw, n))
        assert np.array_equal(ch.reshape(cw, n), ch0) and _host_intact(wch, cw * n, canary)
        wpl, pl = _host(lay.blob_words * n, np.uint64, canary)
        cir.stage(hb.reshape(n, -1), out=pl.reshape(lay.blob_words, n))
        assert np.array_equal(pl.reshape(lay.blob_words, n), planes0) and _host_intact(wpl, lay.blob_words * n, canary)
    cir.close()


@pytest.mark.parametrize("n", [1, 33, 257])
def test_hash_entry_points_stay_inside_their_buffers(p2v, ctx, n):
    import torch

    rng = np.random.default_rng(n)
    states = rng.integers(0, fixtures.P, size=(12, n), dtype=np.uint64)
    want_perm = ctx.permutation(states)
    for w in (1, 9, 135):
        leaves = rng.integers(0, fixtures.P, size=(w, n), dtype=np.uint64)
        want_dig = ctx.sponge(leaves)
        for canary in CANARIES:
            wl, l = _dev(w * n, torch.int64, canary)
            l.copy_(torch.from_numpy(leaves.reshape(-1).view(np.int64)))
            wd, d = _dev(4 * n, torch.int64, canary)
            torch.cuda.synchronize()
            ctx.sponge(l.view(w, n), out=d.view(4, n))
            ctx.sync()
            assert np.array_equal(d.cpu().numpy().view(np.uint64).reshape(4, n), want_dig)
            assert _intact(wl, w * n, canary) and _intact(wd, 4 * n, canary)
    for canary in CANARIES:
        ws, s = _dev(12 * n, torch.int64, canary)
        s.copy_(torch.from_numpy(states.reshape(-1).view(np.int64)))
        wo, o = _dev(12 * n, torch.int64, canary)
        torch.cuda.synchronize()
        ctx.permutation(s.view(12, n), out=o.view(12, n))
        ctx.sync()
        assert np.array_equal(o.cpu().numpy().view(np.uint64).reshape(12, n), want_perm)
        assert _intact(ws, 12 * n, canary) and _intact(wo, 12 * n, canary)


@pytest.mark.parametrize("w,log_n,cap_h,n", [(5, 6, 2, 37), (135, 8, 4, 100), (9, 3, 3, 5), (16, 4, 0, 19)])
def test_merkle_entry_points_stay_inside_their_buffers(p2v, ctx, w, log_n, cap_h, n):
    import torch

    rng = np.random.default_rng(w + n)
    nl = 1 << log_n
    leaves = rng.integers(0, fixtures.P, size=(w, nl), dtype=np.uint64)
    idx = rng.integers(0, nl, size=n, dtype=np.uint32)
    dig0 = ctx.merkle_build(leaves, log_n, cap_h)
    lo0, so0, cap0 = ctx.merkle_open(leaves, log_n, cap_h, dig0, idx)
    plen = log_n - cap_h
    nd = dig0.size
    for canary in CANARIES:
        wl, l = _dev(w * nl, torch.int64, canary)
        l.copy_(torch.from_numpy(leaves.reshape(-1).view(np.int64)))
        wdg, dg = _dev(nd, torch.int64, canary)
        wi, ix = _dev(n, torch.int32, canary)
        ix.copy_(torch.from_numpy(idx.view(np.int32)))
        wlo, lo = _dev(w * n, torch.int64, canary)
        wso, so = _dev(max(plen * 4 * n, 1), torch.int64, canary)
        wcap, cap = _dev(4 << cap_h, torch.int64, canary)
        wok, ok = _dev((n + 31) // 32, torch.int32, canary)
        wro, ro = _dev(4 * n, torch.int64, canary)
        ok.zero_()
        torch.cuda.synchronize()
        ctx.merkle_build(l.view(w, nl), log_n, cap_h, out=dg)
        ctx.merkle_open(l.view(w, nl), log_n, cap_h, dg, ix, leaves_out=lo.view(w, n), sibs_out=so[:plen * 4 * n].view(plen * 4, n),
                        cap_out=cap.view(1 << cap_h, 4))
        ctx.checkMerkleProof(cap.view(1 << cap_h, 4), ix, lo.view(w, n), so[:plen * 4 * n].view(plen * 4, n), want_roots=True, ok_bits=ok,
                             roots=ro.view(4, n))
        ctx.sync()
        assert np.array_equal(dg.cpu().numpy().view(np.uint64), dig0)
        assert np.array_equal(lo.cpu().numpy().view(np.uint64).reshape(w, n), lo0)
        assert np.array_equal(so[:plen * 4 * n].cpu().numpy().view(np.uint64).reshape(plen * 4, n), so0)
        assert np.array_equal(cap.cpu().numpy().view(np.uint64).reshape(-1, 4), cap0)
        assert p2v.unpack_bits(ok.cpu().numpy().view(np.uint32), n).all()
        for whole, m in ((wl, w * nl), (wdg, nd), (wi, n), (wlo, w * n), (wso, max(plen * 4 * n, 1)), (wcap, 4 << cap_h), (wok, (n + 31) // 32),
                         (wro, 4 * n)):
            assert _intact(whole, m, canary)
