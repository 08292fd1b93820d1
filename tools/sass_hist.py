#!/usr/bin/env python3
"""SASS opcode histogram of one kernel, per basic block, with the pipe cycles each block costs one warp.

    python tools/sass_hist.py <file.o|file.so> <kernel-name-substring> [--min N] [--dump]

Pipe model (measured with p2v_int_pipe_peak on B200), cycles a warp instruction occupies its pipe on
one SM sub-partition: IMAD.WIDE / IMAD.HI 4 (FMA pipe), other IMAD.* 2 (FMA pipe), DFMA/DADD/DMUL 2 (FP64 pipe),
everything else 2 (ALU pipe; loads/branches are counted there too, they are few).  Issue: 1 instruction per cycle.
Blocks are split at branch targets and after branches; only blocks with at least --min instructions are printed.
"""
import collections
import re
import subprocess
import sys

FMA32 = {"IMAD", "IMAD.IADD", "IMAD.X", "IMAD.MOV.U32", "IMAD.MOV", "IMAD.U32", "IMAD.SHL.U32", "IMAD.HI.U32.X"}
WIDE = {"IMAD.WIDE.U32", "IMAD.WIDE.U32.X", "IMAD.HI.U32", "IMAD.WIDE", "IMAD.HI"}
FP64 = {"DFMA", "DADD", "DMUL"}


def kernel_sass(path, pattern):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    chunks = re.split(r"\n\s*Function : ", out)
    for ch in chunks[1:]:
        name = ch.split("\n", 1)[0].strip()
        if pattern in name:
            return name, ch
    raise SystemExit("no kernel matching %r in %s" % (pattern, path))


def parse(text):
    ins = []
    for l in text.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", l)
        if m:
            t = m.group(2).strip()
            toks = t.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            ins.append((int(m.group(1), 16), op, t))
    return ins


def classify(c):
    f = sum(v for k, v in c.items() if k in FMA32)
    w = sum(v for k, v in c.items() if k in WIDE)
    d = sum(v for k, v in c.items() if k in FP64)
    n = sum(c.values())
    return n, f, w, d, n - f - w - d


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    min_n = 30
    if "--min" in sys.argv:
        min_n = int(sys.argv[sys.argv.index("--min") + 1])
        args = [a for a in args if a != str(min_n)]
    path, pattern = args[0], args[1]
    name, text = kernel_sass(path, pattern)
    ins = parse(text)
    if "--dump" in sys.argv:
        for a, op, t in ins:
            print("%05x  %s" % (a, t))
        return
    cuts = {ins[0][0]}
    for i, (a, op, t) in enumerate(ins):
        if op.startswith(("BRA", "EXIT", "RET", "CALL", "BSYNC", "BRX")):
            if i + 1 < len(ins):
                cuts.add(ins[i + 1][0])
            m = re.search(r"0x([0-9a-f]+)\s*$", t)
            if op.startswith("BRA") and m:
                cuts.add(int(m.group(1), 16))
    cuts = sorted(cuts)
    regs = re.search(r"REG:(\d+)", text)
    print("kernel %s: %d instructions%s" % (name, len(ins), ", %s registers" % regs.group(1) if regs else ""))
    tot = collections.Counter(op for _, op, _ in ins)
    print("static total: n %d fma32 %d wide %d fp64 %d alu/other %d" % classify(tot))
    for k, lo in enumerate(cuts):
        hi = cuts[k + 1] if k + 1 < len(cuts) else ins[-1][0] + 16
        blk = [(a, op, t) for a, op, t in ins if lo <= a < hi]
        if len(blk) < min_n:
            continue
        c = collections.Counter(op for _, op, _ in blk)
        n, f, w, d, o = classify(c)
        last = blk[-1][2]
        print("block %05x-%05x  n %4d | FMA-pipe %4d cyc (imad32 %d, wide %d) | ALU %4d cyc (%d) | FP64 %4d cyc (%d) | ends: %s"
              % (lo, hi, n, 2 * f + 4 * w, f, w, 2 * o, o, 2 * d, d, last))
        print("      " + ", ".join("%s %d" % kv for kv in c.most_common(16)))


if __name__ == "__main__":
    main()
