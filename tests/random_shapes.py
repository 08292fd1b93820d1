"""Random VALID circuit shapes for the property tests: oracle/p2v_prover's `rand<K>` (all-Noop rows, random gate set, selector
groups, wires, challenges, public inputs, quotient degree factor, FRI parameters and reduction strategy) and `rreal<K>` (the real
circuit of real5 — active gates, copy constraints, real quotient — under random FRI parameters) presets, generated at test time
into a temporary directory.  TEST INFRASTRUCTURE (the prover shares the oracle's headers)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROVER = os.path.join(ROOT, "oracle", "p2v_prover")
P = 0xFFFFFFFF00000001


def generate(tmpdir, preset, seed=1):
    """-> {"common": text, "vkey": text, "proof": text}; the prover's own self-check must ACCEPT."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "p2v_prover"], check=True)
    prefix = os.path.join(str(tmpdir), preset)
    r = subprocess.run([PROVER, preset, prefix, "--seed", str(seed), "--threads", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1] == "0", (preset, r.stdout[-300:], r.stderr[-600:])
    return {k: open("%s_%s.json" % (prefix, k)).read() for k in ("common", "vkey", "proof")}


def tampered(blob, proof_words, n, seed):
    """n copies of the blob; copy i is intact iff i % 4 == 0, otherwise ONE word gets +delta (mod p): the first half of the
    tampered copies hit the per-proof part (caps, openings, final polynomial, PoW witness, public inputs), the rest any word."""
    rng = np.random.default_rng(seed)
    blob = np.asarray(blob, dtype=np.uint64)
    blobs = np.tile(blob, (n, 1))
    words = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        if i % 4 == 0:
            continue
        w = int(rng.integers(0, proof_words)) if i % 8 < 4 else int(rng.integers(0, len(blob)))
        d = 1 if i % 3 else int(rng.integers(1, P, dtype=np.uint64))
        blobs[i, w] = (int(blobs[i, w]) % P + d) % P
        words[i] = w
    return blobs, words
