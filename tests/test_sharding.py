"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo groups.  Each rank "verifies" its slice with the CPU
oracle (the GPU verifier cannot run here), packs its accept bits, and the gathered bitmap must equal the
single-process verdicts."""
import os
import socket

import numpy as np
import pytest

import fixtures


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from plonky2_verifier_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape, lay, vkey, blob = fixtures.load("small6")
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, n_total, seed=5)
    start, stop = sharding.shard_bounds(n_total, rank, world)
    words_per_rank = sharding.slice_len(n_total, world) // 32
    local = np.zeros(words_per_rank, dtype=np.uint32)
    if stop > start:
        res = oracle_lib.load().verify_batch(shape, vkey, blobs[start:stop], threads=2, fast=True)
        packed = sharding.pack_bits(res["status"] == 0)
        local[: len(packed)] = packed
    full = sharding.gather_accept_bitmap(torch.from_numpy(local.view(np.int32)), n_total, dist)
    np.save(os.path.join(out_dir, "bitmap_%d.npy" % rank), full.numpy().view(np.uint32))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 100), (2, 64), (3, 70)])
def test_sharded_bitmap_equals_single_process(tmp_path, orc, p2v, world, n_total):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    shape, lay, vkey, blob = fixtures.load("small6")
    blobs, words, deltas = fixtures.tampered_batch(blob, lay, shape, n_total, seed=5)
    want = orc.verify_batch(shape, vkey, blobs, threads=4, fast=True)["status"] == 0
    for rank in range(world):
        got = np.load(os.path.join(str(tmp_path), "bitmap_%d.npy" % rank))
        assert np.array_equal(p2v.unpack_bits(got, n_total), want), rank


def test_shard_bounds_cover_the_batch(p2v):
    from plonky2_verifier_b200 import sharding

    for n in (1, 31, 32, 33, 1000, 100000):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(s[0] % 32 == 0 for s in spans if s[1] > s[0])
    flags = np.random.default_rng(0).random(77) < 0.5
    assert np.array_equal(p2v.unpack_bits(sharding.pack_bits(flags), 77), flags)


def test_c_slicing_rule_equals_python(p2v):
    """p2v_shard_slice_len / p2v_shard_bounds (csrc/sharded_api.cu) == sharding.slice_len / shard_bounds."""
    from plonky2_verifier_b200 import sharding

    for n in (0, 1, 31, 32, 33, 64, 1000, 32808, 100000, 10**6 + 7):
        for world in (1, 2, 3, 4, 8):
            assert p2v.shard_slice_len(n, world) == sharding.slice_len(n, world)
            for r in range(world):
                assert p2v.shard_bounds(n, r, world) == sharding.shard_bounds(n, r, world)
    with pytest.raises(p2v.P2VError):
        p2v.shard_bounds(10, 2, 2)


def test_nccl_bootstrap_symbols_without_gpu(p2v):
    """The communicator bootstrap is bound at run time: without libnccl the call fails cleanly, with it rank 0 gets 128 bytes."""
    try:
        uid = p2v.nccl_unique_id()
    except p2v.P2VError as e:
        assert e.code in (-6, -2)
    else:
        assert len(uid) == 128 and any(uid)

