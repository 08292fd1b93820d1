#!/usr/bin/env python3
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table: kernel, block, launches, total ms, share.
    python tools/launch_summary.py gpurun_out/r02c_launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
k_name, k_block, k_unit, k_val = hdr.index("Kernel Name"), hdr.index("Block Size"), hdr.index("Metric Unit"), hdr.index("Metric Value")
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= k_val:
        continue
    name = re.sub(r"\(.*", "", r[k_name])
    ns = float(r[k_val].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[k_unit], 1)
    key = (name, r[k_block])
    n, t = acc.get(key, (0, 0.0))
    acc[key] = (n + 1, t + ns)
total = sum(t for _, t in acc.values())
print("| kernel | block | launches | total ms | share |\n|---|---|---|---|---|")
for (name, block), (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %s | %d | %.2f | %.1f%% |" % (name, block, n, t / 1e6, 100 * t / total))
merkle = sum(t for (name, _), (_, t) in acc.items() if "k_fri_merkle" in name)
verifier = sum(t for (name, _), (_, t) in acc.items() if not any(x in name for x in ("k_int_pipe", "k_synth", "at::")))
print("\nk_fri_merkle (all instantiations) = %.1f%% of the verifier's kernel time in this list (%.2f of %.2f ms)" % (100 * merkle / verifier, merkle / 1e6, verifier / 1e6))
