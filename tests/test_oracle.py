"""CPU tests of the oracle itself (no GPU): the reference's only known-answer test (Poseidon KAT,
Hash/Poseidon.hs:27-35), the second-order vectors of SURVEY.md App. I, internal consistency of the three
permutation forms, and the verdicts on the bundled fixtures / the tamper matrix."""
import ctypes as C

import numpy as np
import pytest

import fixtures
from oracle_lib import P, rand_felts

KAT = [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2,
       0x414754bfebd051f0, 0xcb1f8980294a023f, 0x6eb2a9e4d54a9d0f, 0x1902bc3af467e056,
       0xf045d5eafdc6021f, 0xe4150f77caaa3be5, 0xc9bfd01d39b50cce, 0x5c0a27fcb0e1459b]


def col(vals):
    return np.array(vals, dtype=np.uint64).reshape(-1, 1)


def test_reference_kat_all_three_forms(orc):
    s = col(range(12))
    for which in (0, 1, 2):
        assert [int(x) for x in orc.permutation(s, which)[:, 0]] == KAT


def test_permutation_forms_agree(orc):
    rng = np.random.default_rng(0)
    s = rand_felts(rng, (12, 300))
    a, b, c = orc.permutation(s, 0), orc.permutation(s, 1), orc.permutation(s, 2)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert (a < np.uint64(P)).all()


def test_app_i_hash_vectors(orc):
    z = orc.permutation(np.zeros((12, 1), dtype=np.uint64))
    assert [int(x) for x in z[:4, 0]] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca]
    s8 = orc.sponge(col(range(1, 9)))
    want8 = [0xd110aa6a46373941, 0x8f238fcceb658894, 0x9cd4f8353866fb4f, 0x274913f0007aa232]
    assert [int(x) for x in s8[:, 0]] == want8
    assert [int(x) for x in orc.compress(col([1, 2, 3, 4]), col([5, 6, 7, 8]))[:, 0]] == want8
    s9 = orc.sponge(col(range(1, 10)))
    assert [int(x) for x in s9[:, 0]] == [0x5a90f7c562413c2b, 0xa1874b91e26076d4, 0x37b5cd4fe1fb94da, 0x3db54acf2fa3b131]
    assert not orc.sponge(np.zeros((0, 1), dtype=np.uint64)).any()  # sponge [] = zero digest, no permutation


def test_app_i_merkle_vector(orc):
    leaves = np.array([[10 * i + j for i in range(4)] for j in range(5)], dtype=np.uint64)  # SoA [5][4]
    dig = orc.sponge(leaves)
    n01 = orc.compress(dig[:, 0:1], dig[:, 1:2])
    n23 = orc.compress(dig[:, 2:3], dig[:, 3:4])
    root = orc.compress(n01, n23)
    assert [int(x) for x in root[:, 0]] == [0x4a10eed2d9416570, 0x158e1d2536247c53, 0xc960fc6ae726b93d, 0xcd1a719ae833f4e9]
    sib = np.concatenate([dig[:, 3:4], n01], axis=0)  # proof for idx 2 = [sponge leaf_3, compress(d0,d1)]
    ok, roots = orc.checkMerkleProof(root.T.copy(), np.array([2], dtype=np.uint32), leaves[:, 2:3], sib)
    assert ok[0] == 1 and np.array_equal(roots, root)


def _duplex(orc, ops, inputs):
    ops = np.array(ops, dtype=np.int32)
    inp = np.array(inputs, dtype=np.uint64)
    out = np.zeros(64, dtype=np.uint64)
    perms = C.c_ulonglong(0)
    fn = orc.lib.orc_duplex_script
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n = fn(ops.ctypes.data, len(ops), inp.ctypes.data, out.ctypes.data, C.addressof(perms))
    return [int(x) for x in out[:n]], perms.value


def test_app_i_duplex_vectors(orc):
    out, perms = _duplex(orc, [3, -3, 1, -2], [1, 2, 3, 4])
    assert out[:3] == [0xac10d793d6f2d150, 0xdb027851a321db07, 0xec4f99114b4e6723]
    assert out[3:] == [0x8ef2bffeb9ec72fc, 0xf09971c2335660a2]
    assert perms == 2
    # squeezing 9 elements from a fresh state takes two permutations (Pure.hs:64-69)
    out, perms = _duplex(orc, [-9], [])
    assert perms == 2 and len(out) == 9
    # a full buffer of 8 is only permuted when the 9th element arrives (lazy absorb, Pure.hs:54-56)
    assert _duplex(orc, [8], list(range(8)))[1] == 0
    assert _duplex(orc, [9], list(range(9)))[1] == 1


def test_app_i_algebra_vectors(orc):
    f = orc.lib.orc_field_op
    f.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    f.restype = C.c_uint64
    assert f(5, 4, 0) == 0x1000
    assert f(5, 15, 0) == 0x1a0037386d98ca5e
    assert f(3, 16, 0) == 0xefffffff10000001
    assert f(3, 0, 0) == 0  # inv 0 = 0 (Goldilocks.hs:155)
    assert f(0, f(3, 12345, 0), 12345) == 1
    assert f(4, 7, (-3) & (2**64 - 1)) == f(3, f(4, 7, 3), 0)  # negative exponents invert first
    e = orc.lib.orc_ext_op
    e.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    x, y, out = np.array([3, 5], dtype=np.uint64), np.array([7, 11], dtype=np.uint64), np.zeros(2, dtype=np.uint64)
    e(0, x.ctypes.data, y.ctypes.data, out.ctypes.data)
    assert [int(v) for v in out] == [406, 68]
    e(1, x.ctypes.data, None, out.ctypes.data)
    assert [int(v) for v in out] == [0x831597216a68de13, 0xd03159714ea68de2]
    rb = orc.lib.orc_reverse_bits
    rb.argtypes = [C.c_int, C.c_uint64]
    rb.restype = C.c_uint64
    assert rb(4, 0b0011) == 0b1100 and rb(15, 1) == 1 << 14


@pytest.mark.parametrize("name", fixtures.ACCEPTING)
def test_bundled_fixtures_accept(orc, name):
    shape, lay, vkey, blob = fixtures.load(name)
    res = orc.verify_batch(shape, vkey, blob, threads=1, fast=name.endswith("12"))
    assert res["status"][0] == 0
    assert res["eqmask"][0] == (1 << shape.num_challenges) - 1
    if name.startswith("real"):
        assert res["combined"].any()      # active gates: C_j(zeta) != 0 and equals Z_H(zeta) * quotient(zeta)
    else:
        assert not res["combined"].any()  # the trivial circuit: every combined constraint value is 0
    assert (res["qstatus"] == 0).all()
    assert orc.blob_words(shape, blob) == lay.blob_words


def test_s12_permutation_count(orc):
    """commentary/FRI.md:263-267: 2774 = 114 + 28*(77+11+7) at the standard recursion shape (+2: the PI hash is
    computed by proofChallenges and again by evalAllPlonkConstraints, 1 permutation each for 4 public inputs)."""
    shape, lay, vkey, blob = fixtures.load("s12")
    res = orc.verify_batch(shape, vkey, blob, threads=1, fast=True)
    assert res["perms"] == 2774 + 2


@pytest.mark.parametrize("name,code", [("small6_badfinal", 3), ("small6_badlayer0", 18), ("small6_badlayer1", 18 | (1 << 16)),
                                       ("real5_badwitness", 1 | (3 << 16)), ("real5_badcopy", 1 | (3 << 16)),
                                       ("reallu6_badlookup", 1 | (3 << 16))])
def test_regrinded_rejections(orc, name, code):
    shape, lay, vkey, blob = fixtures.load(name)
    res = orc.verify_batch(shape, vkey, blob, threads=1, fast=False)
    assert res["status"][0] == code


def test_tamper_matrix_verdicts(orc):
    """which first failure each class of tampered word produces."""
    shape, lay, vkey, blob = fixtures.load("small6")
    words = fixtures.tamper_words(lay, shape)
    names = sorted(words)
    blobs = np.tile(blob, (len(names), 1))
    for i, nm in enumerate(names):
        blobs[i, words[nm]] = (int(blobs[i, words[nm]]) + 1) % P
    res = orc.verify_batch(shape, vkey, blobs, threads=4, fast=False)
    got = {nm: int(res["status"][i]) & 0xFF for i, nm in enumerate(names)}
    eqs = {"wires_cap", "zs_pp_cap", "quotient_cap", "public_input", "open_zs", "open_zs_next", "open_pp", "open_quotient",
           "open_sigma", "open_constant"}
    powf = {"pow_witness", "final_poly", "commit_cap", "open_wire_routed", "open_wire_advice"}
    init = {"q0_leaf0", "q0_leaf1", "q0_leaf2", "q0_leaf3", "q0_sib1", "q0_sib3_last", "qlast_leaf1"}
    step = {"q0_step0_eval", "q0_step0_sib", "qlast_step_last_eval"}
    for nm in names:
        want = 1 if nm in eqs else 2 if nm in powf else 16 if nm in init else 17 if nm in step else None
        assert want is not None, nm
        assert got[nm] == want, (nm, hex(got[nm]))
    # detail bits: query index and oracle mask / step
    st = {nm: int(res["status"][i]) for i, nm in enumerate(names)}
    assert st["q0_leaf2"] == 16 | (0 << 8) | (0b0100 << 16)
    assert st["qlast_leaf1"] == 16 | ((shape.num_queries - 1) << 8) | (0b0010 << 16)
    assert st["qlast_step_last_eval"] == 17 | ((shape.num_queries - 1) << 8) | ((shape.num_steps - 1) << 16)


def test_vkey_cap_tamper_hits_only_matching_queries(orc):
    """Only circuit_digest is absorbed (Challenge/Verifier.hs:73): a vkey cap tamper fails INIT_MERKLE(oracle 0)
    at the first query that lands under that cap entry, and is invisible otherwise."""
    shape, lay, vkey, blob = fixtures.load("mid5")
    base = orc.verify_batch(shape, vkey, blob, threads=1, fast=True)
    idx = base["challenges"][-shape.num_queries:, 0].astype(np.int64)
    entries = idx >> lay.init_path_len
    for entry in range(1 << shape.cap_height):
        vk2 = vkey.copy()
        vk2[4 * entry] = (int(vk2[4 * entry]) + 1) % P
        st = int(orc.verify_batch(shape, vk2, blob, threads=1, fast=True)["status"][0])
        hits = np.nonzero(entries == entry)[0]
        if len(hits) == 0:
            assert st == 0
        else:
            assert st == 16 | (int(hits[0]) << 8) | (1 << 16)
