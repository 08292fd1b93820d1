"""Multi-GPU parity on hardware (BASELINE config 5): spawns one process per GPU with torch.distributed.run when at least
two GPUs are visible (skipped otherwise; the gloo tests in tests/test_sharding.py cover the host logic on CPU)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import fixtures

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_sharded_world1_is_verify_batch(p2v, ctx, orc):
    """world == 1: p2v_verify_batch_sharded needs no communicator and equals p2v_verify_batch."""
    shape, lay, vkey, blob = fixtures.load("small6")
    cir = p2v.Circuit(ctx, shape, vkey)
    blobs, _, _ = fixtures.tampered_batch(blob, lay, shape, 77, seed=1)
    a, s = cir.verifyProof(blobs)
    a2, s2 = cir.verifyProofSharded(blobs, 77, 0, 1)
    assert np.array_equal(a, a2) and np.array_equal(s, s2)
    with pytest.raises(p2v.P2VError):
        cir.verifyProofSharded(blobs[:64], 77, 0, 2)  # world 2 without a communicator


def test_two_gpus_same_batch_bitmap_identical():
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4 if ngpu < 8 else 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED_OK world=%d" % world in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ---- the same from a host without Python or torch: plonky2-verifier_b200/apps/shard_demo.cpp (include/p2v.h only) forks one
# process per GPU, bootstraps NCCL with p2v_nccl_unique_id / p2v_nccl_init (the id travels through a file) and requires on
# every rank: gathered bitmap == that rank's single-GPU verification of the whole batch (NCCL gather and peer-store gather)
def _demo(world, n_total, name="small6"):
    import plonky2_verifier_b200 as p2v_mod

    p2v_mod.build()
    exe = os.path.join(os.path.dirname(HERE), "plonky2-verifier_b200", "p2v_shard_demo")
    g = lambda k: os.path.join(fixtures.GOLDEN, "%s_%s.json" % (name, k))
    env = dict(os.environ, P2V_PEER_TIMEOUT_S="20")
    return subprocess.run([exe, g("common"), g("vkey"), g("proof"), str(world), str(n_total)], capture_output=True, text=True, timeout=600, env=env)


def test_c_host_world1():
    r = _demo(1, 333)
    assert r.returncode == 0 and "SHARD_DEMO_OK world=1 n=333 accepted=84" in r.stdout, r.stdout + r.stderr


def test_c_host_one_process_per_gpu():
    import torch

    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4 if ngpu < 8 else 8
    for n_total in (1000, 37):  # 37: the last ranks hold a short or EMPTY slice
        r = _demo(world, n_total)
        assert r.returncode == 0 and "SHARD_DEMO_OK world=%d n=%d accepted=%d" % (world, n_total, (n_total + 3) // 4) in r.stdout, r.stdout + r.stderr
    r = _demo(world, 200, "real5")
    assert r.returncode == 0 and "SHARD_DEMO_OK world=%d n=200 accepted=50" % world in r.stdout, r.stdout + r.stderr
