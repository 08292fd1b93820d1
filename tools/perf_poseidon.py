"""Quick device-side throughput probe for K1 (used while tuning; bench.py is the contract)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import plonky2_verifier_b200 as p2v

ctx = p2v.Context(0)
peaks = {m: ctx.int_pipe_peak(m) for m in range(5)}
print("int-pipe groups/s: lop3+imad.wide %.3e | 2lop3+imad.wide %.3e | lop3+imad32 %.3e | lop3+iadd3 %.3e | 2imad.wide+lop3 %.3e" % tuple(peaks[m] for m in range(5)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 2048 * 8
x = torch.randint(0, 2**62, (12, n), dtype=torch.int64, device="cuda")
y = torch.empty_like(x)
stream = torch.cuda.ExternalStream(ctx.stream)
for _ in range(3):
    ctx.permutation(x, out=y)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
    for _ in range(5):
        ctx.permutation(x, out=y)
    e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 5
perms = n / (ms * 1e-3)
print(json.dumps({"n": n, "ms": ms, "perms_per_s": perms, "imad_frac_6376": perms * 6376 / peaks[0], "peaks": peaks}))
