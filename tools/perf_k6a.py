"""Lean tuning probe: per-kernel times of one serial pass + pipelined throughput of the full verifier, for the library
selected by P2V_LIB_PATH.  One JSON line.   python tools/perf_k6a.py [proofs] [fixture]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import fixtures, plonky2_verifier_b200 as p2v

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
fx = sys.argv[2] if len(sys.argv) > 2 else "real12"
ctx = p2v.Context(0)
shape, lay, vkey, blob = fixtures.load(fx)
cir = p2v.Circuit(ctx, shape, vkey)
_, words, deltas = fixtures.tampered_batch(blob, lay, shape, 2048, seed=5)
reps = (n + len(words) - 1) // len(words)
words_n = np.tile(words, reps)[:n].copy(); deltas_n = np.tile(deltas, reps)[:n].copy()
d_blobs = torch.empty((n, lay.blob_words), dtype=torch.int64, device="cuda")
cir.synth_batch(blob, n, words_n, deltas_n, d_blobs)
bits = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda"); status = torch.zeros(n, dtype=torch.int32, device="cuda")
ctx.sync()
peak = ctx.int_pipe_peak(0)
ctx.set_pipeline(1)
for _ in range(2):
    cir.verifyProof(d_blobs, n=n, accept_bits=bits, status=status)
ctx.sync()
ms = {k: ctx.last_ms(k) for k in ("stage", "challenges", "constraints", "fri", "fri_merkle", "verdict")}
c8 = lambda w: (w + 7) // 8
ppq = sum(c8(lay.oracle_width[o]) + lay.init_path_len for o in range(4)) + sum(c8(2 << shape.step_arity_bits[s]) + lay.step_path_len[s] for s in range(shape.num_steps))
perms = n * shape.num_queries * ppq
pps = perms / (ms["fri_merkle"] * 1e-3)
ctx.set_pipeline(p2v.DEFAULT_PIPELINE)
if os.environ.get("P2V_PERF_CHUNK"):  # several chunks even for a small batch, so that the pipelined instantiation is what launches (profiles)
    ctx.set_chunk(int(os.environ["P2V_PERF_CHUNK"]))
st = torch.cuda.ExternalStream(ctx.stream)
for _ in range(2):
    cir.verifyProof(d_blobs, n=n, accept_bits=bits, status=status)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record()
K = 4
for _ in range(K):
    cir.verifyProof(d_blobs, n=n, accept_bits=bits, status=status)
with torch.cuda.stream(st):
    e1.record()
e1.synchronize()
step = e0.elapsed_time(e1) / K
h = int(torch.sum(status.to(torch.int64) * torch.arange(1, n + 1, device="cuda") % 1000003).item())
print(json.dumps({"lib": os.path.basename(p2v.LIB_PATH), "n": n, "fixture": fx, "frac": pps * 6376 / peak, "perms_per_s": pps, "kernel_ms": ms, "step_ms": step,
                  "proofs_per_s": n / (step * 1e-3), "status_hash": h, "accepted": int((status == 0).sum().item())}))
