#!/usr/bin/env python3
"""Write tests/golden/<name>_tamper.json = the tamper matrix (field name -> word offset in the flat proof) of every
accepting fixture, from tests/fixtures.py::tamper_words.  tests/test_host.py checks the files against the live layout."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fixtures  # noqa: E402

for name in fixtures.ACCEPTING:
    shape, lay, vkey, blob = fixtures.load(name)
    table = {k: int(v) for k, v in sorted(fixtures.tamper_words(lay, shape).items())}
    with open(os.path.join(HERE, "%s_tamper.json" % name), "w") as fh:
        json.dump(table, fh, indent=0, sort_keys=True)
        fh.write("\n")
    print(name, len(table))
