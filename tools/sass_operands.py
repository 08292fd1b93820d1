#!/usr/bin/env python3
"""Register-operand words read per basic block of one kernel (the register-file side of the issue model).

    python tools/sass_operands.py <file.so> <kernel-name-substring> [--min N]

For every instruction the register SOURCE operands are counted in 32-bit words: 2 per operand of DFMA / DADD / DMUL and for
the 64-bit addend of IMAD.WIDE, 1 otherwise; RZ, immediates, constant-bank and uniform-register operands are free; an
operand that hits the reuse cache (same register in the same slot as the previous instruction, which marked it `.reuse`) is
counted separately.  Measured on B200 (p2v_int_pipe_peak modes 16-23): a scheduler sustains about 2 such words per cycle."""
import collections
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from sass_hist import kernel_sass, parse  # noqa: E402

FP64 = ("DFMA", "DADD", "DMUL")


def words(op, operands, prev):
    w = hits = 0
    cur = {}
    slot = 0
    for o in operands[1:]:
        if re.match(r"^!?U?P[T0-9]", o):
            continue  # predicate / carry operands
        m = re.match(r"^[-~|!]?(R\d+)(\.reuse)?(\.64)?\|?$", o)
        if m:
            wd = 2 if op in FP64 or (op.startswith("IMAD.WIDE") and slot == 2) else 1
            if prev.get(slot) == m.group(1):
                hits += wd
            else:
                w += wd
            if m.group(2):
                cur[slot] = m.group(1)
        slot += 1
    return w, hits, cur


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    min_n = 200
    if "--min" in sys.argv:
        min_n = int(sys.argv[sys.argv.index("--min") + 1])
        args = [a for a in args if a != str(min_n)]
    name, text = kernel_sass(args[0], args[1])
    ins = parse(text)
    cuts = {ins[0][0]}
    for i, (a, op, t) in enumerate(ins):
        if op.startswith(("BRA", "EXIT", "RET", "CALL", "BSYNC", "BRX")):
            if i + 1 < len(ins):
                cuts.add(ins[i + 1][0])
            m = re.search(r"0x([0-9a-f]+)\s*$", t)
            if op.startswith("BRA") and m:
                cuts.add(int(m.group(1), 16))
    cuts = sorted(cuts)
    print("kernel %s" % name)
    for k, lo in enumerate(cuts):
        hi = cuts[k + 1] if k + 1 < len(cuts) else ins[-1][0] + 16
        blk = [t for a, op, t in ins if lo <= a < hi]
        if len(blk) < min_n:
            continue
        tot = hits = 0
        prev = {}
        per = collections.Counter()
        for t in blk:
            toks = t.split(None, 1)
            if toks[0].startswith("@"):
                toks = toks[1].split(None, 1)
            op = toks[0]
            operands = [o.strip() for o in (toks[1] if len(toks) > 1 else "").split(",")]
            w, h, prev = words(op, operands, prev)
            tot += w
            hits += h
            per[op] += w
        print("block %05x-%05x  n %4d | %5d operand words (%.2f per instruction, %d more served by the reuse cache) = %d cycles at 2 words/clk"
              % (lo, hi, len(blk), tot, tot / len(blk), hits, tot // 2))
        print("      " + ", ".join("%s %d" % kv for kv in per.most_common(8)))


if __name__ == "__main__":
    main()
