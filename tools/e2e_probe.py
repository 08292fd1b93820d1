"""Host-buffer path timing, step by step (tuning helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import fixtures, plonky2_verifier_b200 as p2v
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
shape, lay, vkey, blob = fixtures.load("s12")
ctx = p2v.Context(0); cir = p2v.Circuit(ctx, shape, vkey)
if len(sys.argv) > 2:
    ctx.set_chunk(int(sys.argv[2]))
if len(sys.argv) > 3:
    ctx.set_pipeline(int(sys.argv[3]))
W = lay.blob_words
h = torch.empty((n, W), dtype=torch.int64, pin_memory=True)
hb = h.numpy().view(np.uint64); hb[:] = blob
bits = torch.zeros((n + 31) // 32, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
st = torch.zeros(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
for i in range(10):
    t0 = time.perf_counter(); cir.verifyProof(hb, n=n, accept_bits=bits, status=st); dt = time.perf_counter() - t0
    print("step %d: %.1f ms  %.0f proofs/s" % (i, dt * 1e3, n / dt))
d = torch.empty((n, W), dtype=torch.int64, device="cuda"); torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("raw H2D: %.1f GB/s" % (n * W * 8 / dt / 1e9))
