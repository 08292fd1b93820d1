"""Throughput of BASELINE.json configs 2 and 3 (parity-test cases, timed here for the record):
config 2 = 2^20 Merkle openings against one tree of 2^20 leaves (width 135, cap 2^4); config 3 = p2v_fri on 10^4 S12 proofs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import fixtures, plonky2_verifier_b200 as p2v

ctx = p2v.Context(0)
width, log_n, cap_height, n = 135, 20, 4, 1 << 20
nl = 1 << log_n
leaves = torch.randint(0, 2**62, (width, nl), dtype=torch.int64, device="cuda")
digests = torch.empty(4 * ((2 << log_n) - (1 << cap_height)), dtype=torch.int64, device="cuda")
d_idx = torch.randint(0, nl, (n,), dtype=torch.int32, device="cuda")
lo = torch.empty((width, n), dtype=torch.int64, device="cuda")
so = torch.empty(((log_n - cap_height) * 4, n), dtype=torch.int64, device="cuda")
cap = torch.empty((1 << cap_height, 4), dtype=torch.int64, device="cuda")
bits = torch.zeros(n // 32, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()


def timed(f, reps=5):
    f(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    ctx.sync()
    return (time.perf_counter() - t0) / reps


t_build = timed(lambda: ctx.merkle_build(leaves, log_n, cap_height, out=digests))
ctx.merkle_open(leaves, log_n, cap_height, digests, d_idx, leaves_out=lo, sibs_out=so, cap_out=cap); ctx.sync()
t_verify = timed(lambda: ctx.checkMerkleProof(cap, d_idx, lo, so, ok_bits=bits))
ok = int(np.unpackbits(bits.cpu().numpy().view(np.uint8)).sum())
perms_build = nl * 17 + (nl - (1 << cap_height))
perms_verify = n * (17 + log_n - cap_height)
print("config 2: tree build 2^20 x 135: %.1f ms (%.3e perms/s); 2^20 openings verified: %.1f ms (%.3e openings/s, %.3e perms/s, %.0f GB/s), %d accepted"
      % (t_build * 1e3, perms_build / t_build, t_verify * 1e3, n / t_verify, perms_verify / t_verify, n * (width * 8 + 16 * 32 + 4) / t_verify / 1e9, ok))

shape, lay, vkey, blob = fixtures.load("s12")
cir = p2v.Circuit(ctx, shape, vkey)
m = 10000
sched, words, deltas = fixtures.tampered_batch(blob, lay, shape, 2048, seed=3)
words_n = np.tile(words, 5)[:m].copy(); deltas_n = np.tile(deltas, 5)[:m].copy()
d_blobs = torch.empty((m, lay.blob_words), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
cir.synth_batch(blob, m, words_n, deltas_n, d_blobs); ctx.sync()
t_fri = timed(lambda: cir.checkFRIProof(d_blobs, n=m))
print("config 3: p2v_fri on 10^4 S12 proofs (device-resident): %.1f ms (%.3e proofs/s)" % (t_fri * 1e3, m / t_fri))
