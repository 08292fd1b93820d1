"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

P = 0xFFFFFFFF00000001


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s", "liboracle.so"], check=True)
    return LIB


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp, sz = C.c_void_p, C.c_size_t
        lib.orc_poseidon_permute.argtypes = [vp, vp, sz, C.c_int]
        lib.orc_hash_leaves.argtypes = [vp, C.c_uint32, sz, vp]
        lib.orc_compress.argtypes = [vp, vp, vp, sz]
        lib.orc_merkle_verify.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint32, vp, C.c_uint32, sz, vp, vp]
        lib.orc_perm_loop.argtypes = [sz, C.c_int, C.c_int]
        lib.orc_perm_loop.restype = C.c_uint64

    def permutation(self, states, which=0):
        states = np.ascontiguousarray(states, dtype=np.uint64)
        out = np.empty_like(states)
        self.lib.orc_poseidon_permute(states.ctypes.data, out.ctypes.data, states.shape[1], which)
        return out

    def sponge(self, leaves):
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
        w, n = leaves.shape
        out = np.empty((4, n), dtype=np.uint64)
        self.lib.orc_hash_leaves(leaves.ctypes.data if w else None, w, n, out.ctypes.data)
        return out

    def compress(self, left, right):
        left = np.ascontiguousarray(left, dtype=np.uint64)
        right = np.ascontiguousarray(right, dtype=np.uint64)
        out = np.empty_like(left)
        self.lib.orc_compress(left.ctypes.data, right.ctypes.data, out.ctypes.data, left.shape[1])
        return out

    def checkMerkleProof(self, cap, idx, leaves, siblings):
        cap = np.ascontiguousarray(cap, dtype=np.uint64)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
        siblings = np.ascontiguousarray(siblings, dtype=np.uint64)
        w, n = leaves.shape
        path_len = siblings.shape[0] // 4
        cap_height = int(cap.shape[0]).bit_length() - 1
        ok = np.zeros(n, dtype=np.uint8)
        roots = np.empty((4, n), dtype=np.uint64)
        self.lib.orc_merkle_verify(leaves.ctypes.data if w else None, w, idx.ctypes.data,
                                   siblings.ctypes.data if path_len else None, path_len, cap.ctypes.data,
                                   cap_height, n, ok.ctypes.data, roots.ctypes.data)
        return ok, roots

    def perm_loop(self, iters, threads, which=2):
        return self.lib.orc_perm_loop(iters, threads, which)

    def verify_batch(self, shape, vkey, blobs, threads=8, fast=True):
        """Restatement of verifyProof (+ all intermediates) on a batch of flat blobs.
        Returns a dict: challenges [cw][n], combined [2r][n], eqmask, status, fri_status, qstatus [n][Q],
        folded [2][n*Q], perms."""
        import plonky2_verifier_b200 as p2v

        blobs = np.ascontiguousarray(blobs, dtype=np.uint64)
        if blobs.ndim == 1:
            blobs = blobs.reshape(1, -1)
        n = blobs.shape[0]
        vkey = np.ascontiguousarray(vkey, dtype=np.uint64)
        cw = challenges_words_py(shape)
        Q, r = shape.num_queries, shape.num_challenges
        out = dict(
            challenges=np.zeros((cw, n), dtype=np.uint64), combined=np.zeros((2 * r, n), dtype=np.uint64),
            eqmask=np.zeros(n, dtype=np.uint8), status=np.zeros(n, dtype=np.uint32),
            fri_status=np.zeros(n, dtype=np.uint32), qstatus=np.zeros((n, Q), dtype=np.uint32),
            folded=np.zeros((2, n * Q), dtype=np.uint64))
        perms = C.c_ulonglong(0)
        fn = self.lib.orc_verify_batch
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]
        fn(C.addressof(shape), vkey.ctypes.data, blobs.ctypes.data, n, threads, 1 if fast else 0,
           out["challenges"].ctypes.data, out["combined"].ctypes.data, out["eqmask"].ctypes.data,
           out["status"].ctypes.data, out["fri_status"].ctypes.data, out["qstatus"].ctypes.data,
           out["folded"].ctypes.data, C.addressof(perms))
        out["perms"] = perms.value
        return out

    def gate_constraints(self, shape, gate_index, wires, consts, pih, max_out=256):
        wires = np.ascontiguousarray(wires, dtype=np.uint64)
        consts = np.ascontiguousarray(consts, dtype=np.uint64)
        pih = np.ascontiguousarray(pih, dtype=np.uint64)
        out = np.zeros((max_out, 2), dtype=np.uint64)
        fn = self.lib.orc_gate_constraints
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        k = fn(C.addressof(shape), gate_index, wires.ctypes.data, wires.shape[0], consts.ctypes.data, consts.shape[0],
               pih.ctypes.data, out.ctypes.data, max_out)
        if k < 0:
            raise RuntimeError("oracle raised while evaluating the gate")
        return out[:k]

    def blob_words(self, shape, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint64)
        fn = self.lib.orc_blob_words
        fn.argtypes = [C.c_void_p, C.c_void_p]
        fn.restype = C.c_size_t
        return fn(C.addressof(shape), blob.ctypes.data)


class OracleCircuit:
    """The oracle's OWN way in (oracle/json_reader.hpp): the reference's JSON files -> the oracle's records, no product
    parser involved.  `proof_blob` flattens a proof in declaration order (what the product calls a blob)."""

    def __init__(self, lib, common_json, vkey_json):
        self.lib = lib
        to_b = lambda t: t.encode() if isinstance(t, str) else t
        cj, vj = to_b(common_json), to_b(vkey_json)
        lib.orc_circuit_from_json.restype = C.c_void_p
        lib.orc_circuit_from_json.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        lib.orc_last_error.restype = C.c_char_p
        self.h = lib.orc_circuit_from_json(cj, len(cj), vj, len(vj))
        if not self.h:
            raise ValueError("oracle: " + lib.orc_last_error().decode())
        lib.orc_proof_from_json.restype = C.c_longlong
        lib.orc_proof_from_json.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.orc_vkey_words.restype = C.c_longlong
        lib.orc_vkey_words.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orc_circuit_info.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_testmain.restype = C.c_longlong
        lib.orc_testmain.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        lib.orc_fri_roots.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orc_circuit_free.argtypes = [C.c_void_p]
        info = (C.c_int * 8)()
        lib.orc_circuit_info(self.h, info)
        (self.num_challenges, self.num_queries, self.num_steps, self.num_lookup_polys, self.degree_bits, self.rate_bits,
         self.cap_height, self.num_public_inputs) = list(info)

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_circuit_free(self.h)
                self.h = None
        except Exception:
            pass

    def vkey_words(self):
        n = self.lib.orc_vkey_words(self.h, None, 0)
        out = np.zeros(n, dtype=np.uint64)
        self.lib.orc_vkey_words(self.h, out.ctypes.data, n)
        return out

    def proof_blob(self, proof_json):
        pj = proof_json.encode() if isinstance(proof_json, str) else proof_json
        n = self.lib.orc_proof_from_json(self.h, pj, len(pj), None, 0)
        if n < 0:
            raise ValueError("oracle: " + self.lib.orc_last_error().decode())
        out = np.zeros(n, dtype=np.uint64)
        self.lib.orc_proof_from_json(self.h, pj, len(pj), out.ctypes.data, n)
        return out

    def challenges_words(self):
        r = self.num_challenges
        return 3 * r + (4 * r if self.num_lookup_polys > 0 else 0) + 4 + 2 * self.num_steps + 1 + self.num_queries

    def verify_batch(self, blobs, threads=8, fast=True):
        blobs = np.ascontiguousarray(blobs, dtype=np.uint64)
        if blobs.ndim == 1:
            blobs = blobs.reshape(1, -1)
        n = blobs.shape[0]
        cw, Q, r = self.challenges_words(), self.num_queries, self.num_challenges
        out = dict(
            challenges=np.zeros((cw, n), dtype=np.uint64), combined=np.zeros((2 * r, n), dtype=np.uint64),
            eqmask=np.zeros(n, dtype=np.uint8), status=np.zeros(n, dtype=np.uint32),
            fri_status=np.zeros(n, dtype=np.uint32), qstatus=np.zeros((n, Q), dtype=np.uint32),
            folded=np.zeros((2, n * Q), dtype=np.uint64))
        perms = C.c_ulonglong(0)
        fn = self.lib.orc_verify_batch_h
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_void_p]
        fn(self.h, blobs.ctypes.data, n, threads, 1 if fast else 0, out["challenges"].ctypes.data, out["combined"].ctypes.data,
           out["eqmask"].ctypes.data, out["status"].ctypes.data, out["fri_status"].ctypes.data, out["qstatus"].ctypes.data,
           out["folded"].ctypes.data, C.addressof(perms))
        out["perms"] = perms.value
        return out

    def verify_with_challenges(self, blobs, challenges_in):
        """verifyProof against GIVEN challenges (SoA [cw][n], include/p2v.h order) -> combined, eqmask, status, qstatus, folded."""
        blobs = np.ascontiguousarray(blobs, dtype=np.uint64)
        if blobs.ndim == 1:
            blobs = blobs.reshape(1, -1)
        n = blobs.shape[0]
        ch = np.ascontiguousarray(challenges_in, dtype=np.uint64)
        Q, r = self.num_queries, self.num_challenges
        out = dict(combined=np.zeros((2 * r, n), dtype=np.uint64), eqmask=np.zeros(n, dtype=np.uint8), status=np.zeros(n, dtype=np.uint32),
                   qstatus=np.zeros((n, Q), dtype=np.uint32), folded=np.zeros((2, n * Q), dtype=np.uint64))
        fn = self.lib.orc_verify_with_challenges_h
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 6
        fn(self.h, blobs.ctypes.data, n, ch.ctypes.data, out["combined"].ctypes.data, out["eqmask"].ctypes.data, out["status"].ctypes.data,
           out["qstatus"].ctypes.data, out["folded"].ctypes.data)
        return out

    def fri_roots(self, blobs):
        """Recomputed Merkle roots of every opening: SoA [(4+steps)*4][n*Q] (index p*Q+q)."""
        blobs = np.ascontiguousarray(blobs, dtype=np.uint64)
        if blobs.ndim == 1:
            blobs = blobs.reshape(1, -1)
        n = blobs.shape[0]
        out = np.zeros(((4 + self.num_steps) * 4, n * self.num_queries), dtype=np.uint64)
        self.lib.orc_fri_roots(self.h, blobs.ctypes.data, n, out.ctypes.data)
        return out

    def testmain(self, proof_json):
        pj = proof_json.encode() if isinstance(proof_json, str) else proof_json
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.orc_testmain(self.h, pj, len(pj), buf, len(buf))
        if n < 0:
            raise ValueError("oracle: " + self.lib.orc_last_error().decode())
        return buf.value.decode()


def circuit_from_json(common_json, vkey_json):
    return OracleCircuit(load().lib, common_json, vkey_json)


def challenges_words_py(shape):
    r = shape.num_challenges
    return 3 * r + (4 * r if shape.num_lookup_polys > 0 else 0) + 4 + 2 * shape.num_steps + 1 + shape.num_queries


_cached = None


def load():
    global _cached
    if _cached is None:
        build()
        _cached = Oracle(C.CDLL(LIB))
    return _cached


def rand_felts(rng, shape, mode="mixed"):
    """Random u64 test inputs; 'mixed' sprinkles in the edge values of the lazy representation."""
    a = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    if mode == "mixed":
        edge = np.array([0, 1, P - 1, P, P + 1, 2**64 - 1, 2**32 - 1, 2**32, 0xFFFFFFFF00000000], dtype=np.uint64)
        mask = rng.random(size=shape) < 0.1
        a[mask] = rng.choice(edge, size=int(mask.sum()))
    return a
