#!/usr/bin/env python3
"""Write tests/golden/<name>.testmain.txt for every bundled fixture: exactly what the reference's driver
(/root/reference/src/testmain.hs:40-63) prints for that (common, vkey, proof) triple — public-input hash, opening counts,
combined constraint values, quotient-identity verdicts, final verdict — produced by oracle/pyref.py FROM THE JSON ALONE
(no product code, no C++ oracle).  tests/test_testmain_golden.py keeps the committed files, pyref and the C++ oracle in
agreement; tools/ghc_crosscheck.sh is the one-command check for someone who has GHC.

    python tools/testmain_expected.py [--check]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REJECTING = {"small6_badfinal": "small6", "small6_badlayer0": "small6", "small6_badlayer1": "small6",
             "real5_badwitness": "real5", "real5_badcopy": "real5", "reallu6_badlookup": "reallu6"}
# the two 2^12-row fixtures take minutes in pure Python (literal foldCosetWith): written once, checked by the C++ oracle
SLOW = {"s12", "real12"}


def names():
    out = sorted(f[: -len("_proof.json")] for f in os.listdir(GOLDEN) if f.endswith("_proof.json"))
    return out


def main():
    check = "--check" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    bad = 0
    for name in names():
        if only and name not in only:
            continue
        path = os.path.join(GOLDEN, name + ".testmain.txt")
        if name in SLOW and os.path.exists(path) and not only:
            continue
        common, vkey, proof = pyref.load_fixture(GOLDEN, name, REJECTING.get(name))
        text = pyref.testmain_text(common, vkey, proof)
        if check:
            ok = os.path.exists(path) and open(path).read() == text
            print("%-20s %s" % (name, "ok" if ok else "DIFFERS"))
            bad += not ok
        else:
            with open(path, "w") as fh:
                fh.write(text)
            print("wrote", os.path.relpath(path, ROOT))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
