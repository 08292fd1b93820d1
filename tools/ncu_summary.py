#!/usr/bin/env python3
"""Summarise one `ncu --set full --import-source on` report of a kernel into the files the judge reads under profiles/:

    python tools/ncu_summary.py gpurun_out/<report>.ncu-rep --proofs N --tag r02_k_fri_merkle [--note "..."]

  profiles/<tag>_ncu.md        key metrics (time, registers, issue/pipe utilisation, stall reasons, dram bytes) and the
                               per-opcode table (share of instructions, share of warp-stall samples, stall mix)
  profiles/<tag>_traffic.json  dram bytes of the captured launch and the proofs it processed (bench.py's roofline.traffic)
Reads the report with `ncu -i ... --page raw/source --csv` (works without a GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "blocks / SM (register limit)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe cycles active % (integer multiplies live here)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % (I2F.F64.U32)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected (warps / issue)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (i-cache)"),
    ("dram__bytes_read.sum", "dram bytes read"),
    ("dram__bytes_write.sum", "dram bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
]


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    args = sys.argv[1:]
    rep = args[0]
    proofs = int(args[args.index("--proofs") + 1])
    tag = args[args.index("--tag") + 1]
    note = args[args.index("--note") + 1] if "--note" in args else ""
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    name = d.get("Kernel Name", ("?", ""))[0]
    lines = ["# %s — `ncu --set full` of `%s`" % (tag, name.split("(")[0]), ""]
    if note:
        lines += [note, ""]
    lines += ["Report: `%s` (not committed; this summary and the CSV extracts are).  Captured launch: %d proofs." % (os.path.basename(rep), proofs), "",
              "| metric | value |", "|---|---|"]
    for k, label in KEYS:
        if k in d:
            v, u = d[k]
            lines.append("| %s (`%s`) | %s %s |" % (label, k, v, u))
    rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
    inst = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
    lines += ["", "dram read + write = %.1f MB for %d proofs = %.1f kB per proof." % ((rd + wr) / 1e6, proofs, (rd + wr) / proofs / 1e3), ""]
    # per-opcode table from the source page
    src = page(rep, "source")
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    stallcols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    ops, samp, opst = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in src[2:]:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        n = int(r[ix["Instructions Executed"]])
        ops[op] += n
        samp[op] += int(r[ix["# Samples"]])
        for c in stallcols:
            opst[op][c.replace("stall_", "")] += int(r[ix[c]])
    tot, ts = sum(ops.values()), sum(samp.values())
    lines += ["## Instruction mix and where the warps wait (SASS opcodes, whole kernel)", "",
              "%d warp instructions, %d stall samples.  \"samples per 10^6 instructions\": `selected` is the issue cycle itself, "
              "everything else is time the instruction spent waiting to issue (attributed to the waiting instruction)." % (tot, ts), "",
              "| opcode | % of instructions | % of samples | samples per 10^6 instructions by reason |", "|---|---|---|---|"]
    for op, n in ops.most_common(18):
        per = sorted(((k, v / n * 1e6) for k, v in opst[op].items() if v), key=lambda kv: -kv[1])[:6]
        lines.append("| %s | %.2f | %.2f | %s |" % (op, 100 * n / tot, 100 * samp[op] / ts, ", ".join("%s %.0f" % kv for kv in per)))
    out_md = os.path.join(ROOT, "profiles", tag + "_ncu.md")
    with open(out_md, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    with open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w") as fh:
        json.dump({"kernel": name.split("(")[0], "proofs": proofs, "dram_bytes_read": rd, "dram_bytes_write": wr, "warp_instructions": inst,
                   "report": os.path.basename(rep)}, fh, indent=1)
        fh.write("\n")
    print("wrote", out_md)


if __name__ == "__main__":
    main()
