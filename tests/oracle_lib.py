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
