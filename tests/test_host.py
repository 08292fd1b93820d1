"""CPU tests of the product's host side (no GPU): the C-ABI library loads and exports every symbol that
include/p2v.h declares; the JSON / gate-string parsers agree with an independent Python reading of the same
files (json module = exact integers); shape errors are reported, not crashed on; without a GPU the context
refuses to start (there is no CPU fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 0xFFFFFFFF00000001


def test_library_exports_every_declared_symbol(p2v):
    hdr = open(os.path.join(ROOT, "include", "p2v.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(p2v_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(p2v.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), "libp2v.so does not export %s" % sym
    assert declared == set(p2v.EXPORTED_SYMBOLS)
    assert p2v.lib().p2v_abi_version() == 2


def test_no_cpu_fallback(p2v):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(p2v.P2VError) as e:
        p2v.Context(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


# ---- independent Python reading of the wire format ----------------------------------

def py_blob(common, proof_json):
    """Flatten a *_proof.json in the field order of Types.hs:251-279 (the p2v_layout order)."""
    out = []
    f = lambda x: out.append(int(x) % P)
    dig = lambda d: [f(x) for x in d["elements"]]
    cap = lambda c: [dig(d) for d in c]
    ext = lambda xs: [(f(a), f(b)) for a, b in xs]
    pr = proof_json["proof"]
    cap(pr["wires_cap"]); cap(pr["plonk_zs_partial_products_cap"]); cap(pr["quotient_polys_cap"])
    op = pr["openings"]
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys",
              "lookup_zs", "lookup_zs_next"):
        ext(op[k])
    fri = pr["opening_proof"]
    for c in fri["commit_phase_merkle_caps"]:
        cap(c)
    ext(fri["final_poly"]["coeffs"])
    f(fri["pow_witness"])
    for x in proof_json["public_inputs"]:
        f(x)
    for rd in fri["query_round_proofs"]:
        for leaf, path in rd["initial_trees_proof"]["evals_proofs"]:
            for x in leaf:
                f(x)
            cap(path["siblings"])
        for st in rd["steps"]:
            ext(st["evals"])
            cap(st["merkle_proof"]["siblings"])
    return np.array(out, dtype=np.uint64)


GATE_RE = [
    (r"^ArithmeticGate \{ num_ops: (\d+) \}$", 0), (r"^ArithmeticExtensionGate \{ num_ops: (\d+) \}$", 1),
    (r"^BaseSumGate \{ num_limbs: (\d+) \} \+ Base: (\d+)$", 2),
    (r"^CosetInterpolationGate \{ subgroup_bits: (\d+), degree: (\d+), barycentric_weights: \[([0-9, ]*)\], _phantom: .*\}<D=2>$", 3),
    (r"^ConstantGate \{ num_consts: (\d+) \}", 4), (r"^ExponentiationGate \{ num_power_bits: (\d+) \}", 5),
    (r"^LookupGate \{ num_slots: (\d+), lut_hash: \[[0-9, ]*\] \}", 6),
    (r"^LookupTableGate \{ num_slots: (\d+), lut_hash: \[[0-9, ]*\], last_lut_row: (\d+) \}", 7),
    (r"^MulExtensionGate \{ num_ops: (\d+) \}", 8), (r"^NoopGate", 9), (r"^PublicInputGate", 10),
    (r"^PoseidonGate\(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>\)<WIDTH=(\d+)>$", 11),
    (r"^PoseidonMdsGate\(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>\)<WIDTH=(\d+)>$", 12),
    (r"^RandomAccessGate \{ bits: (\d+), num_copies: (\d+), num_extra_constants: (\d+), _phantom: .*\}<D=2>", 13),
    (r"^ReducingGate \{ num_coeffs: (\d+) \}", 14), (r"^ReducingExtensionGate \{ num_coeffs: (\d+) \}", 15),
]


def py_gate(s):
    for rx, kind in GATE_RE:
        m = re.match(rx, s)
        if m:
            g = [x for x in m.groups()]
            if kind == 3:
                return kind, [int(g[0]), int(g[1])], [int(x) % P for x in g[2].split(",") if x.strip()]
            return kind, [int(x) for x in g], []
    return 16, [], []


@pytest.mark.parametrize("name", fixtures.ACCEPTING)
def test_parsers_match_python_reading(p2v, name):
    common = json.loads(fixtures.read(name, "common"))
    shape, lay, vkey, blob = fixtures.load(name)
    cfg, fc = common["config"], common["config"]["fri_config"]
    assert (shape.num_wires, shape.num_routed_wires, shape.num_gate_constants, shape.num_challenges) == (
        cfg["num_wires"], cfg["num_routed_wires"], cfg["num_constants"], cfg["num_challenges"])
    assert (shape.rate_bits, shape.cap_height, shape.pow_bits, shape.num_queries) == (
        fc["rate_bits"], fc["cap_height"], fc["proof_of_work_bits"], fc["num_query_rounds"])
    assert shape.degree_bits == common["fri_params"]["degree_bits"]
    # expandReductionStrategy, Plonk/FRI.hs:337-354
    strat = fc["reduction_strategy"]
    if "ConstantArityBits" in strat:
        a, fbits = strat["ConstantArityBits"]
        steps, logn = [], shape.degree_bits
        while logn > fbits:
            steps.append(a)
            logn -= a
    else:
        steps = strat["Fixed"]
    assert list(shape.step_arity_bits)[: shape.num_steps] == steps
    assert shape.final_poly_len == 1 << (shape.degree_bits - sum(steps))
    for k in ("quotient_degree_factor", "num_constants", "num_public_inputs", "num_partial_products", "num_lookup_polys",
              "num_lookup_selectors"):
        assert getattr(shape, k) == common[k], k
    assert [int(x) for x in shape.k_is[: shape.num_routed_wires]] == [int(x) % P for x in common["k_is"]]
    sel = common["selectors_info"]
    assert [shape.gates[i].group for i in range(shape.num_gates)] == sel["selector_indices"]
    assert [(shape.group_start[i], shape.group_end[i]) for i in range(shape.num_groups)] == [(g["start"], g["end"]) for g in sel["groups"]]
    assert shape.num_gates == len(common["gates"])
    for i, s in enumerate(common["gates"]):
        kind, params, weights = py_gate(s)
        g = shape.gates[i]
        assert g.kind == kind, s
        assert [g.p0, g.p1, g.p2][: len(params)] == params, s
        assert [int(x) for x in shape.weights[g.weights_off: g.weights_off + g.weights_len]] == weights
    luts = common["luts"]
    assert shape.num_luts == len(luts)
    for k, t in enumerate(luts):
        got = [(int(shape.lut_pairs[2 * j]), int(shape.lut_pairs[2 * j + 1])) for j in range(shape.lut_off[k], shape.lut_off[k + 1])]
        assert got == [(a % P, b % P) for a, b in t]
    # proof and vkey blobs
    want = py_blob(common, json.loads(fixtures.read(name, "proof")))
    assert lay.blob_words == len(want)
    assert np.array_equal(blob, want)
    vk = json.loads(fixtures.read(name, "vkey"))
    want_vk = [int(x) % P for d in vk["constants_sigmas_cap"] for x in d["elements"]] + [int(x) % P for x in vk["circuit_digest"]["elements"]]
    assert [int(x) for x in vkey] == want_vk


def test_layout_offsets_are_consistent(p2v):
    shape, lay, vkey, blob = fixtures.load("s12")
    assert lay.blob_words == lay.proof_words + shape.num_queries * lay.query_words == 15881
    assert list(lay.oracle_width) == [85, 135, 20, 16]  # commentary/FRI.md:256
    assert lay.init_path_len == 11 and list(lay.step_path_len)[:2] == [7, 3]
    assert p2v.challenges_words(shape) == 3 * 2 + 2 + 2 + 4 + 1 + 28


GATE_CASES = [
    ("ArithmeticGate { num_ops: 20 }", 0, [20], 20),
    ("ArithmeticGate { num_ops: 20 } trailing", 16, [], 0),  # withEOF (Parser.hs:134-136)
    ("ArithmeticExtensionGate { num_ops: 10 }", 1, [10], 20),
    ("BaseSumGate { num_limbs: 63 } + Base: 2", 2, [63, 2], 64),
    ("ConstantGate { num_consts: 2 }", 4, [2], 2),
    ("ConstantGate { num_consts: 2 } and then some", 4, [2], 2),  # no EOF check in constantGateP (Parser.hs:165-167)
    ("ConstantGate{num_consts:2}", 4, [2], 2),  # `spaces` accepts zero blanks
    ("ExponentiationGate { num_power_bits: 66 }", 5, [66], 67),
    ("MulExtensionGate { num_ops: 13 }", 8, [13], 26),
    ("NoopGate", 9, [], 0), ("NoopGateWithSuffix", 9, [], 0),  # `string "NoopGate"` only
    ("PublicInputGate", 10, [], 4),
    ("PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>", 11, [12], 123),
    ("PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>", 12, [12], 24),
    ("RandomAccessGate { bits: 4, num_copies: 4, num_extra_constants: 2, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>", 13, [4, 4, 2], 26),
    ("ReducingGate { num_coeffs: 43 }", 14, [43], 86), ("ReducingGate { num_coeffs: 43 }<D=2>", 14, [43], 86),
    ("ReducingExtensionGate { num_coeffs: 32 }", 15, [32], 64),
    ("LookupGate { num_slots: 40, lut_hash: [1, 2, 3] }", 6, [40], 0),
    ("LookupTableGate { num_slots: 26, lut_hash: [9, 8], last_lut_row: 3 }", 7, [26, 3], 0),
    ("SomeFutureGate { x: 1 }", 16, [], 0), ("", 16, [], 0), ("arithmeticgate { num_ops: 1 }", 16, [], 0),
]


@pytest.mark.parametrize("text,kind,params,ncons", GATE_CASES)
def test_gate_string_parser(p2v, text, kind, params, ncons):
    g, w = p2v.parse_gate(text)
    assert g.kind == kind
    assert [g.p0, g.p1, g.p2][: len(params)] == params
    assert g.num_constraints == ncons


def test_coset_gate_string(p2v):
    shape, *_ = fixtures.load("s12")
    text = json.loads(fixtures.read("s12", "common"))["gates"][11]
    g, w = p2v.parse_gate(text)
    assert g.kind == 3 and (g.p0, g.p1) == (4, 6) and len(w) == 16 and g.num_constraints == 12
    assert p2v.parse_gate(text.replace("<D=2>", ""))[0].kind == 16  # the suffix is mandatory (Parser.hs:162)


def test_shape_and_parse_errors(p2v):
    shape, lay, vkey, blob = fixtures.load("small6")
    proof = json.loads(fixtures.read("small6", "proof"))
    def expect(obj, code):
        with pytest.raises(p2v.P2VError) as e:
            p2v.parse_proof(json.dumps(obj), shape)
        assert e.value.code == code, str(e.value)
    bad = json.loads(json.dumps(proof)); bad["proof"]["wires_cap"].pop()
    expect(bad, -5)  # validateMerkleCapLength (Plonk/FRI.hs:79-85)
    bad = json.loads(json.dumps(proof)); bad["proof"]["opening_proof"]["query_round_proofs"].pop()
    expect(bad, -5)  # safeZipWith (Plonk/FRI.hs:372)
    bad = json.loads(json.dumps(proof)); bad["proof"]["opening_proof"]["query_round_proofs"][0]["initial_trees_proof"]["evals_proofs"].pop()
    expect(bad, -5)  # expecting 4 Merkle proofs (Plonk/FRI.hs:107)
    bad = json.loads(json.dumps(proof)); bad["proof"]["opening_proof"]["query_round_proofs"][1]["steps"][0]["evals"].pop()
    expect(bad, -5)  # reduction strategy incompatibility (Plonk/FRI.hs:312)
    bad = json.loads(json.dumps(proof)); bad["proof"]["openings"]["wires"][0][0] = 1.5
    expect(bad, -4)
    bad = json.loads(json.dumps(proof)); del bad["public_inputs"]
    expect(bad, -4)
    with pytest.raises(p2v.P2VError):
        p2v.parse_proof("{ not json", shape)
    # field elements >= p and >= 2^64 are reduced like mkGoldilocks (Goldilocks.hs:132)
    big = json.loads(json.dumps(proof)); big["proof"]["opening_proof"]["pow_witness"] = int(proof["proof"]["opening_proof"]["pow_witness"]) + 3 * P
    assert np.array_equal(p2v.parse_proof(json.dumps(big), shape), blob)
    common = json.loads(fixtures.read("small6", "common"))
    c2 = json.loads(json.dumps(common)); c2["config"]["fri_config"]["reduction_strategy"] = {"MinSize": None}
    with pytest.raises(p2v.P2VError) as e:
        p2v.parse_common(json.dumps(c2))
    assert e.value.code == -6  # "reduction strategy not implemented" (Plonk/FRI.hs:342)
    c2 = json.loads(json.dumps(common)); c2["num_constants"] += 1
    with pytest.raises(p2v.P2VError) as e:
        p2v.parse_common(json.dumps(c2))
    assert e.value.code == -5  # getSelectorConfig tally (Selector.hs:33-36)


def test_number_tokens_of_any_length(p2v):
    """`mkGoldilocks <$> parseJSON` (Goldilocks.hs:101-102,132-133): any non-negative integer, reduced mod p exactly —
    19-, 20-, 38-, 39- and 60-digit tokens, values around p, 2^64 and 2^128, leading zeros are not JSON."""
    shape, lay, vkey, blob = fixtures.load("small6")
    proof = json.loads(fixtures.read("small6", "proof"))
    rng = np.random.default_rng(5)
    vals = [0, 1, P - 1, P, P + 1, 2**64 - 1, 2**64, 2**64 + 1, 10**19 - 1, 10**19, 10**38 - 1, 10**38, 2**128 - 1, 2**128, 2**128 + 12345,
            10**59 + 7, 3 * P * P + 5]
    vals += [int(rng.integers(0, 2**63)) ** k + int(rng.integers(0, 2**63)) for k in (1, 2, 3) for _ in range(20)]
    txt = json.dumps(proof)
    for v in vals:
        p2 = json.loads(txt)
        p2["proof"]["opening_proof"]["pow_witness"] = v
        got = p2v.parse_proof(json.dumps(p2), shape)
        assert int(got[lay.off_pow_witness]) == v % P, v


def test_batch_parse_on_threads(p2v):
    """p2v_parse_proofs == p2v_parse_proof on every text, for any thread count; a text that fails to decode is
    reported per proof, leaves a zero blob and does not stop the others."""
    shape, lay, vkey, blob = fixtures.load("small6")
    good = fixtures.read("small6", "proof")
    variants = [good, fixtures.read("small6_badfinal", "proof"), fixtures.read("small6_badlayer0", "proof")]
    texts = [variants[i % 3] for i in range(37)]
    want = np.stack([p2v.parse_proof(t, shape) for t in texts])
    for threads in (1, 3, 0):
        assert np.array_equal(p2v.parse_proofs(texts, shape, threads=threads), want)
    bad = json.loads(good)
    bad["proof"]["wires_cap"].pop()
    texts[5] = json.dumps(bad)
    texts[20] = "{ not json"
    out, rcs = p2v.parse_proofs(texts, shape, threads=4, return_codes=True)
    assert rcs[5] == -5 and rcs[20] == -4 and (np.delete(rcs, [5, 20]) == 0).all()
    assert not out[5].any() and not out[20].any()
    keep = [i for i in range(37) if i not in (5, 20)]
    assert np.array_equal(out[keep], want[keep])
    with pytest.raises(p2v.P2VError) as e:
        p2v.parse_proofs(texts, shape, threads=4)
    assert e.value.code == -5 and "proof 5" in str(e.value)
    assert p2v.parse_proofs([], shape).shape == (0, lay.blob_words)


@pytest.mark.parametrize("name", ["small6", "lookup6", "real5", "s12"])
def test_fast_scan_and_tape_decoders_agree(p2v, name):
    """The forward-scan decoder (keys in serde's declaration order) and the tape decoder (any key order, like aeson)
    give the same blob; re-ordered keys, indentation and exotic-but-legal number tokens fall back to the tape."""
    shape, lay, vkey, blob = fixtures.load(name)
    text = fixtures.read(name, "proof")
    doc = json.loads(text)
    assert np.array_equal(p2v.parse_proof(text, shape), blob)
    assert np.array_equal(p2v.parse_proof(json.dumps(doc, sort_keys=True), shape), blob)      # other key order -> tape
    assert np.array_equal(p2v.parse_proof(json.dumps(doc, indent=2), shape), blob)            # whitespace -> still fast
    doc2 = json.loads(text)
    pw = int(doc2["proof"]["opening_proof"]["pow_witness"])
    big = json.dumps(doc2).replace('"pow_witness": %d' % pw, '"pow_witness": %d' % (pw + 5 * P * P * P))  # 58+ digits -> tape
    assert np.array_equal(p2v.parse_proof(big, shape), blob)
    texts = [text, json.dumps(doc, sort_keys=True), json.dumps(doc, indent=1)] * 3
    assert np.array_equal(p2v.parse_proofs(texts, shape, threads=3), np.tile(blob, (9, 1)))


def test_differential_fuzz_of_the_two_proof_decoders(p2v):
    """1500 randomly mutated proof texts (byte flips, deletions, insertions, truncations): the forward-scan fast path
    + tape fallback and the tape decoder alone must agree on every outcome — same blob or same error code — and
    neither may crash."""
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz_parse.py")
    outs = []
    for extra in ({}, {"P2V_NO_FAST_PARSE": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, script, "1500"], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.split())
    assert outs[0] == outs[1]
    assert int(outs[0][1]) > 50 and int(outs[0][2]) > 500  # both accepted and rejected mutants occurred


def test_committed_tamper_tables_match_the_layout(p2v):
    """tests/golden/<name>_tamper.json (used by bench.py's reference arm, which must not load the product) == the tamper
    matrix computed from the live layout."""
    for name in fixtures.ACCEPTING:
        shape, lay, vkey, blob = fixtures.load(name)
        assert fixtures.tamper_table(name) == {k: int(v) for k, v in fixtures.tamper_words(lay, shape).items()}, name
        a = fixtures.tampered_batch(blob, lay, shape, 40, seed=9)
        b = fixtures.tampered_batch_from_table(blob, fixtures.tamper_table(name), 40, seed=9)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_shape_fields_are_range_checked(p2v):
    """Every integer of a shape is checked before it enters an offset, a shift or a loop bound: values that
    would overflow the layout, divide by zero on the device or wrap in a narrowing cast are refused, both for JSON input and
    for a p2v_shape built by the caller."""
    common = json.loads(fixtures.read("small6", "common"))
    def refuse(mutate, codes=(-6, -5, -4)):
        c = json.loads(json.dumps(common))
        mutate(c)
        with pytest.raises(p2v.P2VError) as e:
            p2v.parse_common(json.dumps(c))
        assert e.value.code in codes, str(e.value)
    def both_cfg(key, val):
        def f(c):
            c["config"]["fri_config"][key] = val
            c["fri_params"]["config"][key] = val
        return f
    refuse(both_cfg("cap_height", 31))          # 4 << 31 overflows an int
    refuse(both_cfg("cap_height", 25))
    refuse(both_cfg("num_query_rounds", 2**31 + 5))   # would wrap to a small positive int
    refuse(both_cfg("num_query_rounds", 0))
    refuse(lambda c: c["config"].__setitem__("num_wires", 2**32 + 20))
    refuse(lambda c: c.__setitem__("num_public_inputs", 2**40))
    refuse(lambda c: c["fri_params"].__setitem__("degree_bits", 40))
    refuse(lambda c: c["selectors_info"]["groups"][0].__setitem__("end", 99))
    # lookup tables with quotient_degree_factor 1: lookup_deg = 0 chunks (Plonk/Lookups.hs:57) -> refused, not divided by
    lu = json.loads(fixtures.read("lookup6", "common"))
    lu["quotient_degree_factor"] = 1
    with pytest.raises(p2v.P2VError) as e:
        p2v.parse_common(json.dumps(lu))
    assert e.value.code == -6
    # a shape handed over as a struct goes through the same checks
    shape, lay, vkey, blob = fixtures.load("small6")
    import copy, ctypes
    bad = p2v.Shape.from_buffer_copy(bytes(shape))
    bad.cap_height = 30
    with pytest.raises(p2v.P2VError):
        p2v.shape_layout(bad)
    bad = p2v.Shape.from_buffer_copy(bytes(shape))
    bad.num_gates = 1000
    with pytest.raises(p2v.P2VError):
        p2v.shape_layout(bad)


def test_fast_scanner_leaves_leading_zeros_to_the_tape_reader(p2v):
    """`007` is not a JSON number (aeson refuses it): the forward-scan fast path must not accept what the definition rejects."""
    shape, lay, vkey, blob = fixtures.load("small6")
    text = fixtures.read("small6", "proof")
    pw = str(int(json.loads(text)["proof"]["opening_proof"]["pow_witness"]))
    assert text.count('"pow_witness":' + pw) + text.count('"pow_witness": ' + pw) >= 1
    bad = text.replace('"pow_witness":' + pw, '"pow_witness":00' + pw).replace('"pow_witness": ' + pw, '"pow_witness": 00' + pw)
    with pytest.raises(p2v.P2VError) as e:
        p2v.parse_proof(bad, shape)
    assert e.value.code == -4



def test_host_parsers_under_asan_and_ubsan(tmp_path):
    """tools/fuzz_host_asan.cpp: the C++ host parsers compiled with -fsanitize=address,undefined and fed mutated common / vkey /
    proof / gate texts of seven fixtures.  Found (fourth session): signed overflow in the constraint count of a gate with a huge
    parameter, and parameters above INT_MAX truncated instead of refused."""
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "fuzz_host")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I", os.path.join(root, "include"),
                         os.path.join(root, "tools", "fuzz_host_asan.cpp"), os.path.join(root, "plonky2-verifier_b200", "csrc", "host", "parse.cpp"),
                         "-o", exe, "-lpthread"], capture_output=True, text=True)
    if cc.returncode != 0 and "sanitize" in cc.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([exe, os.path.join(root, "tests", "golden"), "600"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "proof ok/bad" in run.stdout


def test_hostile_circuit_descriptions_under_asan(tmp_path):
    """oracle/fuzz_shapes.cpp: mutated `*_common.json` texts (structural numbers replaced by boundary values, gate strings dropped or
    duplicated) that the host parser accepts AND p2v_shape_check passes are verified by the CPU oracle compiled with ASAN, UBSAN and
    _GLIBCXX_ASSERTIONS, with a blob of the shape's size.  The oracle indexes where the reference indexes, so a report, an oracle
    exception or a run-away loop means the shape check lets through a circuit a kernel would misbehave on."""
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "fuzz_shapes")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-D_GLIBCXX_ASSERTIONS",
                         "-I", os.path.join(root, "include"), "-I", os.path.join(root, "oracle"), os.path.join(root, "oracle", "fuzz_shapes.cpp"),
                         os.path.join(root, "plonky2-verifier_b200", "csrc", "host", "parse.cpp"), "-o", exe, "-lpthread"], capture_output=True, text=True)
    if cc.returncode != 0 and "sanitize" in cc.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([exe, os.path.join(root, "tests", "golden"), "1200"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "oracle exceptions 0" in run.stdout, run.stdout + run.stderr[-2000:]


def test_shape_check_without_a_gpu():
    """p2v_shape_check vets a circuit description on the host: every bundled circuit passes, hostile gate parameters do not."""
    import plonky2_verifier_b200 as p2v

    for name in fixtures.ACCEPTING:
        p2v.shape_check(fixtures.load(name)[0])
    hostile = [
        "RandomAccessGate { bits: 7, num_copies: 1073741824, num_extra_constants: 0, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
        "RandomAccessGate { bits: 1, num_copies: 4294967297, num_extra_constants: 0, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
        "BaseSumGate { num_limbs: 3 } + Base: 1000000",
        "ArithmeticGate { num_ops: 4294967300 }",
        "ExponentiationGate { num_power_bits: 10 }",   # needs 22 wires, the circuit has 20
        "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>",  # needs 135 wires
        "SomeFutureGate { x: 1 }",                      # UnknownGate: `error` in gateConstraints (Gate/Constraints.hs:108)
    ]
    for gate in hostile:
        common = json.loads(fixtures.read("small6", "common"))
        common["gates"][2] = gate
        with pytest.raises(p2v.P2VError) as ei:
            p2v.shape_check(p2v.parse_common(json.dumps(common)))
        assert ei.value.code in (-5, -6), (gate, ei.value)
