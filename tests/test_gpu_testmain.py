"""Config 1 of BASELINE.json as the reference runs it: `testmain` (src/testmain.hs:24-63) on one bundled (common, vkey, proof)
triple.  plonky2-verifier_b200/p2v_testmain is that driver written in C++ above the C ABI (include/p2v.h only): it must print
byte for byte what the Haskell prints — tests/golden/<name>.testmain.txt, rendered from the JSON alone by oracle/pyref.py
and held equal to the C++ oracle's rendering by tests/test_oracle_json.py; tools/ghc_crosscheck.sh diffs the same files
against the real `testmain` for whoever has GHC."""
import os
import subprocess

import pytest

import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "plonky2-verifier_b200", "p2v_testmain")
NAMES = fixtures.ACCEPTING + sorted(fixtures.REJECTING)


def _paths(name):
    g = lambda n, k: os.path.join(fixtures.GOLDEN, "%s_%s.json" % (n, k))
    return [g(fixtures.REJECTING.get(name, name), "common"), g(name, "vkey"), g(name, "proof")]


def _binary():
    import plonky2_verifier_b200 as p2v

    p2v.build()
    assert os.access(BIN, os.X_OK), "p2v_testmain was not built"
    return BIN


def test_testmain_builds_and_reports_usage():
    r = subprocess.run([_binary()], capture_output=True, text=True)
    assert r.returncode == 64 and "usage" in r.stderr


def test_testmain_decode_errors_need_no_gpu(tmp_path):
    """`decode` failing (testmain.hs:35-37: an irrefutable `Just` pattern) is reported before the GPU is touched."""
    bad = tmp_path / "bad_proof.json"
    bad.write_text(fixtures.read("small6", "proof")[:-20])
    c, v, _ = _paths("small6")
    r = subprocess.run([_binary(), c, v, str(bad)], capture_output=True, text=True)
    assert r.returncode == 2 and "proof" in r.stderr and r.stdout == ""
    r = subprocess.run([_binary(), str(tmp_path), "missing"], capture_output=True, text=True)
    assert r.returncode == 66


def test_testmain_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([_binary()] + _paths("small6"), capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr and r.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_testmain_matches_reference_output(name):
    want = open(os.path.join(fixtures.GOLDEN, name + ".testmain.txt")).read()
    r = subprocess.run([_binary()] + _paths(name), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want


@pytest.mark.gpu
def test_testmain_dir_prefix_form():
    """testmain.hs:31-33 reads <dir>/<prefix>_{common,vkey,proof}.json"""
    want = open(os.path.join(fixtures.GOLDEN, "real5.testmain.txt")).read()
    r = subprocess.run([_binary(), fixtures.GOLDEN, "real5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout == want
