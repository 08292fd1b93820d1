"""Helper of tests/test_host.py::test_differential_fuzz_of_the_two_proof_decoders: decodes randomly mutated proof texts
and prints a digest of every outcome (blob bytes or error code).  Run once normally (forward-scan fast path in front of
the tape decoder) and once with P2V_NO_FAST_PARSE=1 (tape decoder only): the digests must be equal."""
import hashlib
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import fixtures  # noqa: E402
import plonky2_verifier_b200 as p2v  # noqa: E402


def main(rounds):
    shape, lay, vkey, blob = fixtures.load("small6")
    text = fixtures.read("small6", "proof").encode()
    rng = random.Random(7)
    h = hashlib.sha256()
    ok = bad = 0
    for _ in range(rounds):
        b = bytearray(text)
        for _ in range(rng.choice([1, 1, 2, 5])):
            if len(b) < 2:
                break
            op, pos = rng.randrange(4), rng.randrange(len(b))
            if op == 0:
                b[pos] = rng.choice(b'0123456789[]{},:" -e.')
            elif op == 1:
                del b[pos:pos + rng.randrange(1, 8)]
            elif op == 2:
                b[pos:pos] = bytes(rng.choice(b'0123456789[]{},:" ') for _ in range(rng.randrange(1, 6)))
            else:
                b = b[:pos]
        try:
            h.update(p2v.parse_proof(bytes(b), shape).tobytes())
            ok += 1
        except p2v.P2VError as e:
            h.update(b"err%d" % e.code)
            bad += 1
    print(h.hexdigest(), ok, bad)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1500)
