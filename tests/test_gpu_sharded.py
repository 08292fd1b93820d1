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
