"""Pipeline depth 2 vs 3 on the host-buffer and device-resident paths (tuning helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import fixtures, plonky2_verifier_b200 as p2v
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
shape, lay, vkey, blob = fixtures.load("s12")
ctx = p2v.Context(0); cir = p2v.Circuit(ctx, shape, vkey)
W = lay.blob_words
h = torch.empty((n, W), dtype=torch.int64, pin_memory=True)
hb = h.numpy().view(np.uint64); hb[:] = blob
d = h.cuda(); torch.cuda.synchronize()
bits = torch.zeros((n + 31) // 32, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
st = torch.zeros(n, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
dbits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda"); dst = torch.zeros(n, dtype=torch.int32, device="cuda")
for depth in (2, 3, 2, 3):
    ctx.set_pipeline(depth)
    for name, src, ob, os_ in (("host", hb, bits, st), ("device", d, dbits, dst)):
        ts = []
        for i in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            cir.verifyProof(src, n=n, accept_bits=ob, status=os_); ctx.sync()
            ts.append(time.perf_counter() - t0)
        print("depth %d %-6s: %.1f ms  %.0f proofs/s" % (depth, name, min(ts[1:]) * 1e3, n / min(ts[1:])))
