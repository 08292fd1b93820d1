"""Worker of tests/test_gpu_sharded.py (one process per GPU, launched by torch.distributed.run): BASELINE config 5 in small —
the SAME seeded batch on every rank, sharded with p2v_shard_bounds, verified through the C export p2v_verify_batch_sharded
(NCCL all-gather of the accept bitmap).  Every rank requires: gathered bitmap == its own single-GPU verification of the
whole batch == the CPU oracle on the whole batch; local status words == the matching slice.  Ragged sizes leave the last
rank with a short or EMPTY slice."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import fixtures
import oracle_lib
import plonky2_verifier_b200 as p2v
from plonky2_verifier_b200 import sharding


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = p2v.Context(local)
    orc = oracle_lib.load()
    for name, sizes in (("small6", (100, 64, 31, 33, 1)), ("real5", (200,)), ("s12", (96,))):
        shape, lay, vkey, blob = fixtures.load(name)
        cir = p2v.Circuit(ctx, shape, vkey)
        for n_total in sizes:
            blobs, words, _ = fixtures.tampered_batch(blob, lay, shape, n_total, seed=77)
            want = orc.verify_batch(shape, vkey, blobs, threads=8, fast=True)["status"]
            acc1, st1 = cir.verifyProof(blobs)                      # this rank alone, whole batch
            assert np.array_equal(st1, want), (name, n_total, "single-GPU run differs from the oracle")
            start, stop = p2v.shard_bounds(n_total, rank, world)
            assert (start, stop) == sharding.shard_bounds(n_total, rank, world)
            # (a) host buffers straight through the C export
            sharding.init_comm(ctx, dist)
            acc, st = cir.verifyProofSharded(blobs[start:stop], n_total, rank, world)
            assert np.array_equal(acc, want == 0), (name, n_total, rank, "gathered bitmap differs")
            assert np.array_equal(st, want[start:stop]), (name, n_total, rank, "local status differs")
            # (b) device tensors through the torch-facing wrapper (stream ordering against torch's current stream)
            d_local = torch.from_numpy(blobs[start:stop].view(np.int64).copy()).cuda()
            full, st_d = sharding.verify_batch_sharded(cir, d_local, n_total, rank, world, dist)
            bits = p2v.unpack_bits(full.cpu().numpy().view(np.uint32), n_total)
            assert np.array_equal(bits, want == 0), (name, n_total, rank, "device path: gathered bitmap differs")
            assert np.array_equal(st_d.cpu().numpy().view(np.uint32), want[start:stop])
            # padding bits of the gathered words are zero
            words_full = p2v.shard_slice_len(n_total, world) // 32 * world
            raw = np.zeros(words_full, dtype=np.uint32)
            cir.verifyProofSharded(blobs[start:stop], n_total, rank, world, accept_bits_full=raw, status=np.empty(max(stop - start, 1), dtype=np.uint32))
            allbits = p2v.unpack_bits(raw, words_full * 32)
            per = p2v.shard_slice_len(n_total, world)
            for r in range(world):
                a, b = p2v.shard_bounds(n_total, r, world)
                assert not allbits[r * per + (b - a):(r + 1) * per].any(), "padding bits set"
                assert np.array_equal(allbits[r * per: r * per + (b - a)], want[a:b] == 0), "rank %d's slice is not at words [%d, %d)" % (r, r * per // 32, (r + 1) * per // 32)
        cir.close()
    # ---- the call is a collective: a rank whose slice cannot be verified still takes part in the gather (all-zero slice),
    # the other ranks neither hang nor see accepted proofs there, and only the failing rank reports an error ----
    def failing_rank_case(tag):
        ctx_other = p2v.Context(local)
        shape, lay, vkey, blob = fixtures.load("small6")
        cir, foreign = p2v.Circuit(ctx, shape, vkey), p2v.Circuit(ctx_other, shape, vkey)  # `foreign` belongs to another context
        n_total, bad = 48 * world + 5, 0
        blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, n_total, seed=5)
        want = orc.verify_batch(shape, vkey, blobs, threads=8, fast=True)["status"]
        start, stop = p2v.shard_bounds(n_total, rank, world)
        raw = np.zeros(p2v.shard_slice_len(n_total, world) // 32 * world, dtype=np.uint32)
        st = np.empty(max(stop - start, 1), dtype=np.uint32)
        local_blobs = np.ascontiguousarray(blobs[start:stop])
        rc = p2v.lib().p2v_verify_batch_sharded(ctx._h, (foreign if rank == bad else cir)._h, p2v._ptr(local_blobs), n_total, rank, world,
                                                p2v._ptr(raw), p2v._ptr(st))
        if rank == bad:
            assert rc == -1 and b"all-zero slice" in p2v.lib().p2v_last_error(ctx._h), (tag, rc)
        else:
            assert rc == 0, (tag, rank, rc)
        expect = want == 0
        a, b = p2v.shard_bounds(n_total, bad, world)
        assert expect[a:b].any(), "the failing rank's slice must hold proofs that would have been accepted"
        expect[a:b] = False
        assert np.array_equal(p2v.unpack_bits(raw, n_total), expect), (tag, rank, "gather after a local failure")
        # and the communicator is still usable afterwards
        acc, _ = cir.verifyProofSharded(local_blobs, n_total, rank, world)
        assert np.array_equal(acc, want == 0), (tag, rank, "call after a local failure")
        cir.close(); foreign.close(); ctx_other.close()

    failing_rank_case("nccl")
    # ---- the gather without NCCL: direct stores into the peers' buffers + flags (p2v_peer_enable) ----
    ctx.peer_enable()
    failing_rank_case("peer")
    shape, lay, vkey, blob = fixtures.load("small6")
    cir = p2v.Circuit(ctx, shape, vkey)
    for rep, n_total in enumerate((100, 33, 64, 100, 7, 256, 100)):  # several epochs back to back: both halves of the buffers
        blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, n_total, seed=100 + rep)
        want = orc.verify_batch(shape, vkey, blobs, threads=8, fast=True)["status"]
        start, stop = p2v.shard_bounds(n_total, rank, world)
        acc, st = cir.verifyProofSharded(blobs[start:stop], n_total, rank, world)
        assert np.array_equal(acc, want == 0), ("peer gather", n_total, rank)
        assert np.array_equal(st, want[start:stop])
        d_local = torch.from_numpy(blobs[start:stop].view(np.int64).copy()).cuda()
        full, _ = sharding.verify_batch_sharded(cir, d_local, n_total, rank, world, dist)
        assert np.array_equal(p2v.unpack_bits(full.cpu().numpy().view(np.uint32), n_total), want == 0), ("peer gather, device buffers", n_total, rank)
    cir.close()
    ctx.peer_disable()
    r_, w_, ver = ctx.nccl_info()
    assert (r_, w_) == (rank, world) and ver > 0
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK world=%d nccl=%d (NCCL all-gather and peer-store gather)" % (world, ver))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
