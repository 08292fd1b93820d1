"""Device-resident throughput as a function of the chunk size and pipeline depth (tuning helper)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import fixtures, plonky2_verifier_b200 as p2v
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
shape, lay, vkey, blob = fixtures.load("s12")
ctx = p2v.Context(0); cir = p2v.Circuit(ctx, shape, vkey)
W = lay.blob_words
d = torch.from_numpy(np.tile(blob, (n, 1)).view(np.int64)).cuda(); torch.cuda.synchronize()
dbits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda"); dst = torch.zeros(n, dtype=torch.int32, device="cuda")
chunks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 2048, 3072, 4096, 6144, 8192, 16896]
depths = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3]
for chunk in chunks:
    for depth in depths:
        ctx.set_chunk(chunk); ctx.set_pipeline(depth)
        ts = []
        for i in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            cir.verifyProof(d, n=n, accept_bits=dbits, status=dst); ctx.sync()
            ts.append(time.perf_counter() - t0)
        print("chunk %5d depth %d: %.1f ms  %.0f proofs/s" % (chunk, depth, min(ts[1:]) * 1e3, n / min(ts[1:])))
