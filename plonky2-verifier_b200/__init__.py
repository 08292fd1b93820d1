"""plonky2-verifier_b200 — B200-native batch Plonky2 verifier (host-side Python mirror).

The product is `libp2v.so` (hand-written sm_100a CUDA + C++ host code behind the C ABI in
`include/p2v.h`).  This module is the thin ctypes layer over it, named after the reference's
module surface so parity tests read like the reference:

    permutation / sponge / compress / checkMerkleProof     src/Hash/*.hs
    proofChallenges                                         src/Challenge/Verifier.hs:58
    evalCombinedPlonkConstraints / checkCombinedPlonkEquations   src/Plonk/Vanishing.hs:48, Plonk/Verifier.hs:31
    checkFRIProof                                           src/Plonk/FRI.hs:358
    verifyProof                                             src/Plonk/Verifier.hs:56

There is NO CPU fallback: if the CUDA library is missing or no sm_100 GPU is present, calls raise.
Arrays are numpy (host) or anything exposing `data_ptr()` (torch CUDA tensors = device pointers).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# P2V_LIB_PATH: tuning aid only (tools/variants_prebuilt.sh loads compile-time variants of the same library)
LIB_PATH = os.environ.get("P2V_LIB_PATH") or os.path.join(_HERE, "libp2v.so")

P2V_MAX_GATES = 64
P2V_MAX_GROUPS = 16
P2V_MAX_ROUTED = 256
P2V_MAX_STEPS = 16
P2V_MAX_LUTS = 16
P2V_MAX_WEIGHTS = 256

GATE_KINDS = [
    "ArithmeticGate", "ArithmeticExtensionGate", "BaseSumGate", "CosetInterpolationGate", "ConstantGate",
    "ExponentiationGate", "LookupGate", "LookupTableGate", "MulExtensionGate", "NoopGate", "PublicInputGate",
    "PoseidonGate", "PoseidonMdsGate", "RandomAccessGate", "ReducingGate", "ReducingExtensionGate", "UnknownGate",
]

ST_ACCEPT, ST_FALSE_EQS, ST_FALSE_POW, ST_FALSE_FINAL = 0, 1, 2, 3
ST_ERR_INIT_MERKLE, ST_ERR_STEP_MERKLE, ST_ERR_STEP_EVAL = 16, 17, 18


class P2VError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libp2v error %d: %s" % (code, msg))
        self.code = code


class Gate(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "p0", "p1", "p2", "group", "num_constraints", "weights_off", "weights_len")]


class Shape(C.Structure):
    _fields_ = [
        ("num_wires", C.c_int32), ("num_routed_wires", C.c_int32), ("num_gate_constants", C.c_int32),
        ("num_challenges", C.c_int32),
        ("degree_bits", C.c_int32), ("rate_bits", C.c_int32), ("cap_height", C.c_int32), ("pow_bits", C.c_int32),
        ("num_queries", C.c_int32), ("num_steps", C.c_int32), ("step_arity_bits", C.c_int32 * P2V_MAX_STEPS),
        ("final_poly_len", C.c_int32),
        ("quotient_degree_factor", C.c_int32), ("num_constants", C.c_int32), ("num_public_inputs", C.c_int32),
        ("num_partial_products", C.c_int32), ("num_lookup_polys", C.c_int32), ("num_lookup_selectors", C.c_int32),
        ("num_gates", C.c_int32), ("num_groups", C.c_int32),
        ("group_start", C.c_int32 * P2V_MAX_GROUPS), ("group_end", C.c_int32 * P2V_MAX_GROUPS),
        ("gates", Gate * P2V_MAX_GATES),
        ("num_weights", C.c_int32), ("weights", C.c_uint64 * P2V_MAX_WEIGHTS),
        ("k_is", C.c_uint64 * P2V_MAX_ROUTED),
        ("num_luts", C.c_int32), ("lut_off", C.c_int32 * (P2V_MAX_LUTS + 1)),
        ("lut_pairs", C.POINTER(C.c_uint64)),
    ]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "cap_words", "off_wires_cap", "off_zs_pp_cap", "off_quotient_cap",
        "off_open_constants", "off_open_sigmas", "off_open_wires", "off_open_zs", "off_open_zs_next",
        "off_open_pp", "off_open_quotient", "off_open_lookup_zs", "off_open_lookup_zs_next",
        "n_open_constants", "n_open_sigmas", "n_open_wires", "n_open_zs", "n_open_zs_next",
        "n_open_pp", "n_open_quotient", "n_open_lookup_zs", "n_open_lookup_zs_next",
        "off_commit_caps", "off_final_poly", "off_pow_witness", "off_public_inputs", "proof_words")] + [
        ("oracle_width", C.c_int32 * 4), ("init_path_len", C.c_int32),
        ("q_off_leaf", C.c_int32 * 4), ("q_off_sibs", C.c_int32 * 4),
        ("q_off_step_evals", C.c_int32 * P2V_MAX_STEPS), ("q_off_step_sibs", C.c_int32 * P2V_MAX_STEPS),
        ("step_path_len", C.c_int32 * P2V_MAX_STEPS),
        ("query_words", C.c_int32), ("blob_words", C.c_int32), ("vkey_words", C.c_int32),
    ]


class Intermediates(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("challenges_in", "challenges", "combined", "eq_ok_mask", "status", "accept_bits",
                                           "query_status", "folded", "roots")]


_lib = None


def build(force=False, verbose=False):
    """Compile libp2v.so in-tree (nvcc, sm_100a)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_p2v_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise P2VError(-3, "libp2v.so is not built (run `python plonky2-verifier_b200/build.py`); "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64p, u32p, u8p, sz = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t
    sig = {
        "p2v_abi_version": (C.c_int, []),
        "p2v_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "p2v_ctx_destroy": (None, [vp]),
        "p2v_last_error": (C.c_char_p, [vp]),
        "p2v_ctx_stream": (vp, [vp]),
        "p2v_ctx_sync": (C.c_int, [vp]),
        "p2v_ctx_launch_count": (C.c_uint64, [vp]),
        "p2v_host_alloc": (C.c_int, [sz, C.POINTER(vp)]),
        "p2v_host_free": (None, [vp]),
        "p2v_poseidon_permute": (C.c_int, [vp, u64p, u64p, sz]),
        "p2v_hash_leaves": (C.c_int, [vp, u64p, C.c_uint32, sz, u64p]),
        "p2v_compress": (C.c_int, [vp, u64p, u64p, u64p, sz]),
        "p2v_merkle_verify": (C.c_int, [vp, u64p, C.c_uint32, u32p, u64p, C.c_uint32, u64p, C.c_uint32, sz, u32p, u64p]),
        "p2v_merkle_build": (C.c_int, [vp, u64p, C.c_uint32, C.c_uint32, C.c_uint32, u64p]),
        "p2v_merkle_open": (C.c_int, [vp, u64p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, sz, u64p, u64p, u64p]),
        "p2v_parse_common": (C.c_int, [C.c_char_p, sz, C.POINTER(Shape)]),
        "p2v_shape_free": (None, [C.POINTER(Shape)]),
        "p2v_parse_gate": (C.c_int, [C.c_char_p, sz, C.POINTER(Gate), u64p]),
        "p2v_shape_layout": (C.c_int, [C.POINTER(Shape), C.POINTER(Layout)]),
        "p2v_shape_check": (C.c_int, [C.POINTER(Shape)]),
        "p2v_challenges_words": (C.c_int, [C.POINTER(Shape)]),
        "p2v_parse_vkey": (C.c_int, [C.c_char_p, sz, C.POINTER(Shape), u64p]),
        "p2v_parse_proof": (C.c_int, [C.c_char_p, sz, C.POINTER(Shape), u64p]),
        "p2v_parse_proofs": (C.c_int, [C.POINTER(C.c_char_p), C.POINTER(sz), sz, C.POINTER(Shape), u64p, C.c_int, C.c_void_p]),
        "p2v_circuit_create": (C.c_int, [vp, C.POINTER(Shape), u64p, C.POINTER(vp)]),
        "p2v_circuit_destroy": (None, [vp]),
        "p2v_challenges": (C.c_int, [vp, vp, u64p, sz, u64p]),
        "p2v_constraints": (C.c_int, [vp, vp, u64p, sz, u64p, u8p]),
        "p2v_fri": (C.c_int, [vp, vp, u64p, sz, u32p, u32p, u64p]),
        "p2v_verify_batch": (C.c_int, [vp, vp, u64p, sz, u32p, u32p]),
        "p2v_verify_groups": (C.c_int, [vp, sz, C.POINTER(vp), C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(vp), C.c_void_p]),
        "p2v_ctx_set_chunk": (C.c_int, [vp, sz]),
        "p2v_ctx_set_pipeline": (C.c_int, [vp, C.c_int]),
        "p2v_synth_batch": (C.c_int, [vp, vp, u64p, sz, C.c_void_p, u64p, u64p]),
        "p2v_shard_slice_len": (sz, [sz, C.c_int]),
        "p2v_shard_bounds": (C.c_int, [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]),
        "p2v_nccl_unique_id": (C.c_int, [vp]),
        "p2v_nccl_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "p2v_nccl_attach": (C.c_int, [vp, vp]),
        "p2v_nccl_finalize": (C.c_int, [vp]),
        "p2v_nccl_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "p2v_verify_batch_sharded": (C.c_int, [vp, vp, u64p, sz, C.c_int, C.c_int, u32p, u32p]),
        "p2v_stage": (C.c_int, [vp, vp, u64p, sz, u64p]),
        "p2v_peer_enable": (C.c_int, [vp]),
        "p2v_peer_disable": (C.c_int, [vp]),
        "p2v_verify_intermediates": (C.c_int, [vp, vp, u64p, sz, C.POINTER(Intermediates)]),
        "p2v_debug_field_op": (C.c_int, [vp, C.c_int, u64p, u64p, u64p, sz]),
        "p2v_int_pipe_peak": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double)]),
        "p2v_ctx_last_ms": (C.c_int, [vp, C.c_char_p, C.POINTER(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if include/p2v.h and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "p2v_abi_version", "p2v_ctx_create", "p2v_ctx_destroy", "p2v_last_error", "p2v_ctx_stream", "p2v_ctx_sync",
    "p2v_ctx_launch_count", "p2v_host_alloc", "p2v_host_free", "p2v_poseidon_permute", "p2v_hash_leaves",
    "p2v_compress", "p2v_merkle_verify", "p2v_merkle_build", "p2v_merkle_open", "p2v_parse_common",
    "p2v_shape_free", "p2v_parse_gate", "p2v_shape_layout", "p2v_shape_check", "p2v_challenges_words", "p2v_parse_vkey",
    "p2v_parse_proof", "p2v_parse_proofs", "p2v_circuit_create", "p2v_circuit_destroy", "p2v_challenges", "p2v_constraints",
    "p2v_fri", "p2v_verify_batch", "p2v_verify_groups", "p2v_ctx_set_chunk", "p2v_ctx_set_pipeline", "p2v_synth_batch", "p2v_int_pipe_peak", "p2v_ctx_last_ms",
    "p2v_shard_slice_len", "p2v_shard_bounds", "p2v_nccl_unique_id", "p2v_nccl_init", "p2v_nccl_attach", "p2v_nccl_finalize",
    "p2v_nccl_info", "p2v_verify_batch_sharded", "p2v_stage", "p2v_verify_intermediates", "p2v_debug_field_op",
    "p2v_peer_enable", "p2v_peer_disable",
]


def _ptr(a):
    """Raw address of a numpy array (host) or a torch tensor (host or device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError("expected numpy array, torch tensor or raw address, got %r" % type(a))


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ---- host-only parsers (work without a GPU) --------------------------------------------------

def parse_common(json_text):
    """`FromJSON CommonCircuitData` (Types.hs:70) -> Shape."""
    if isinstance(json_text, str):
        json_text = json_text.encode()
    sh = Shape()
    rc = lib().p2v_parse_common(json_text, len(json_text), C.byref(sh))
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return sh


def parse_gate(s):
    """`recognizeGate` (Gate/Parser.hs:107) -> (Gate, weights list)."""
    if isinstance(s, str):
        s = s.encode()
    g = Gate()
    w = np.zeros(P2V_MAX_WEIGHTS, dtype=np.uint64)
    rc = lib().p2v_parse_gate(s, len(s), C.byref(g), w.ctypes.data)
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return g, [int(x) for x in w[: g.weights_len]]


def shape_layout(shape):
    lay = Layout()
    rc = lib().p2v_shape_layout(C.byref(shape), C.byref(lay))
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return lay


def shape_check(shape):
    """p2v_shape_check: vet a circuit description without a GPU (raises P2VError with the reason)."""
    rc = lib().p2v_shape_check(C.byref(shape))
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())


def challenges_words(shape):
    return lib().p2v_challenges_words(C.byref(shape))


def parse_vkey(json_text, shape):
    if isinstance(json_text, str):
        json_text = json_text.encode()
    lay = shape_layout(shape)
    out = np.zeros(lay.vkey_words, dtype=np.uint64)
    rc = lib().p2v_parse_vkey(json_text, len(json_text), C.byref(shape), out.ctypes.data)
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return out


def parse_proof(json_text, shape):
    if isinstance(json_text, str):
        json_text = json_text.encode()
    lay = shape_layout(shape)
    out = np.zeros(lay.blob_words, dtype=np.uint64)
    rc = lib().p2v_parse_proof(json_text, len(json_text), C.byref(shape), out.ctypes.data)
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return out


def parse_proofs(json_texts, shape, threads=0, out=None, return_codes=False):
    """Decode a batch of `*_proof.json` texts on `threads` host threads (0 = all) -> u64 [n][blob_words].
    `out` may be a (pinned) array to fill.  With return_codes=True a proof that fails to decode leaves a zero blob
    and its error code in the second result instead of raising."""
    texts = [t.encode() if isinstance(t, str) else t for t in json_texts]
    n = len(texts)
    lay = shape_layout(shape)
    if out is None:
        out = np.zeros((n, lay.blob_words), dtype=np.uint64)
    assert out.dtype == np.uint64 and out.size >= n * lay.blob_words and out.flags.c_contiguous
    ptrs = (C.c_char_p * n)(*texts)
    lens = (C.c_size_t * n)(*[len(t) for t in texts])
    rcs = np.zeros(n, dtype=np.int32)
    rc = lib().p2v_parse_proofs(ptrs, lens, n, C.byref(shape), out.ctypes.data, int(threads), rcs.ctypes.data)
    if return_codes:
        return out, rcs
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return out


class pinned_u64:
    """Context manager: a page-locked u64 host array from p2v_host_alloc (what the chunk pipeline copies from at
    full PCIe rate), released on exit."""

    def __init__(self, words):
        self.words = int(words)
        self._p = C.c_void_p()

    def __enter__(self):
        rc = lib().p2v_host_alloc(max(self.words, 1) * 8, C.byref(self._p))
        if rc:
            raise P2VError(rc, lib().p2v_last_error(None).decode())
        return np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint64)), shape=(self.words,))

    def __exit__(self, *exc):
        lib().p2v_host_free(self._p)
        self._p = C.c_void_p()
        return False


# ---- GPU context -------------------------------------------------------------------------------

DEFAULT_PIPELINE = 4  # lanes of the chunk pipeline a new context starts with (p2v_ctx_set_pipeline)


class Context:
    """One GPU, one stream.  All methods run on the GPU through the C ABI."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().p2v_ctx_create(int(device), C.byref(self._h))
        if rc:
            raise P2VError(rc, lib().p2v_last_error(None).decode())
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().p2v_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise P2VError(rc, lib().p2v_last_error(self._h).decode())

    @property
    def stream(self):
        return lib().p2v_ctx_stream(self._h)

    def sync(self):
        self._check(lib().p2v_ctx_sync(self._h))

    @property
    def launch_count(self):
        return int(lib().p2v_ctx_launch_count(self._h))

    def set_chunk(self, n):
        self._check(lib().p2v_ctx_set_chunk(self._h, int(n)))

    def set_pipeline(self, depth):
        self._check(lib().p2v_ctx_set_pipeline(self._h, int(depth)))

    # -- multi-GPU (sharded_api.cu) --
    def nccl_init(self, unique_id, rank, world):
        """Collective: join the communicator named by `unique_id` (128 bytes from `nccl_unique_id()` on rank 0)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(lib().p2v_nccl_init(self._h, C.cast(buf, C.c_void_p), int(rank), int(world)))

    def nccl_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        self._check(lib().p2v_nccl_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def peer_enable(self):
        """Collective: gather the bitmap with direct peer stores + flags instead of ncclAllGather (p2v_peer_enable)."""
        self._check(lib().p2v_peer_enable(self._h))

    def peer_disable(self):
        self._check(lib().p2v_peer_disable(self._h))

    def nccl_finalize(self):
        self._check(lib().p2v_nccl_finalize(self._h))

    def last_ms(self, section):
        v = C.c_float()
        self._check(lib().p2v_ctx_last_ms(self._h, section.encode(), C.byref(v)))
        return v.value

    def int_pipe_peak(self, mode=0):
        v = C.c_double()
        self._check(lib().p2v_int_pipe_peak(self._h, int(mode), C.byref(v)))
        return v.value

    def field_op(self, op, a, b=None):
        """Test hook p2v_debug_field_op: one device field routine on arrays of operands (see include/p2v.h)."""
        a = _u64(a)
        b = None if b is None else _u64(b)
        out = np.empty_like(a)
        n = a.shape[-1]
        self._check(lib().p2v_debug_field_op(self._h, int(op), _ptr(a), _ptr(b), _ptr(out), n))
        return out

    # -- L2 Hash (names follow src/Hash/*.hs) --
    def permutation(self, states, out=None):
        """states: SoA [12][n] u64.  `permutation`, Hash/Poseidon.hs:42."""
        n = states.shape[1]
        if out is None:
            out = np.empty((12, n), dtype=np.uint64)
        self._check(lib().p2v_poseidon_permute(self._h, _ptr(states), _ptr(out), n))
        return out

    def sponge(self, leaves, out=None):
        """leaves: SoA [w][n].  `sponge`, Hash/Sponge.hs:26 -> digests SoA [4][n]."""
        w, n = leaves.shape
        if out is None:
            out = np.empty((4, n), dtype=np.uint64)
        self._check(lib().p2v_hash_leaves(self._h, _ptr(leaves) if w else None, w, n, _ptr(out)))
        return out

    def compress(self, left, right, out=None):
        n = left.shape[1]
        if out is None:
            out = np.empty((4, n), dtype=np.uint64)
        self._check(lib().p2v_compress(self._h, _ptr(left), _ptr(right), _ptr(out), n))
        return out

    def checkMerkleProof(self, cap, idx, leaves, siblings, want_roots=False, ok_bits=None, roots=None):
        """`checkMerkleProof`, Hash/Merkle.hs:39, n openings against one cap.
        cap [2^h][4]; idx [n] u32; leaves SoA [w][n]; siblings SoA [len*4][n] -> (ok bool[n], roots?)"""
        w, n = leaves.shape
        path_len = siblings.shape[0] // 4
        cap_height = int(cap.shape[0]).bit_length() - 1
        host_bits = ok_bits is None
        if ok_bits is None:
            ok_bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        if want_roots and roots is None:
            roots = np.empty((4, n), dtype=np.uint64)
        self._check(lib().p2v_merkle_verify(self._h, _ptr(leaves) if w else None, w, _ptr(idx),
                                            _ptr(siblings) if path_len else None, path_len, _ptr(cap), cap_height,
                                            n, _ptr(ok_bits), _ptr(roots)))
        ok = unpack_bits(ok_bits, n) if host_bits else ok_bits
        return (ok, roots) if want_roots else ok

    def merkle_build(self, leaves, log_n, cap_height, out=None):
        w = leaves.shape[0]
        total = 4 * ((2 << log_n) - (1 << cap_height))
        if out is None:
            out = np.empty(total, dtype=np.uint64)
        self._check(lib().p2v_merkle_build(self._h, _ptr(leaves), w, log_n, cap_height, _ptr(out)))
        return out

    def merkle_open(self, leaves, log_n, cap_height, digests, idx, leaves_out=None, sibs_out=None, cap_out=None):
        w = leaves.shape[0]
        n = idx.shape[0]
        if leaves_out is None:
            leaves_out = np.empty((w, n), dtype=np.uint64)
            sibs_out = np.empty(((log_n - cap_height) * 4, n), dtype=np.uint64)
            cap_out = np.empty((1 << cap_height, 4), dtype=np.uint64)
        self._check(lib().p2v_merkle_open(self._h, _ptr(leaves), w, log_n, cap_height, _ptr(digests), _ptr(idx), n,
                                          _ptr(leaves_out), _ptr(sibs_out), _ptr(cap_out)))
        return leaves_out, sibs_out, cap_out


def nccl_unique_id():
    """128 opaque bytes naming a new communicator (rank 0 creates it and ships it to the other ranks)."""
    buf = C.create_string_buffer(128)
    rc = lib().p2v_nccl_unique_id(C.cast(buf, C.c_void_p))
    if rc:
        raise P2VError(rc, lib().p2v_last_error(None).decode())
    return buf.raw


def shard_slice_len(n_total, world):
    return int(lib().p2v_shard_slice_len(int(n_total), int(world)))


def shard_bounds(n_total, rank, world):
    a, b = C.c_size_t(), C.c_size_t()
    rc = lib().p2v_shard_bounds(int(n_total), int(rank), int(world), C.byref(a), C.byref(b))
    if rc:
        raise P2VError(rc, "p2v_shard_bounds: bad rank/world")
    return a.value, b.value


def verify_groups(ctx, groups, return_codes=False):
    """Heterogeneous batch: groups = [(Circuit, blobs), ...] with a different circuit per group ->
    [(accept bool[n_g], status u32[n_g]), ...]  (p2v_verify_groups; the groups' chunks share the pipeline lanes).
    With return_codes=True a group that cannot run is reported in the second result (per-group error codes) instead of
    raising; its verdicts are all-reject."""
    k = len(groups)
    ns = [c._n(b, None) if c is not None and b is not None else 0 for c, b in groups]
    bits = [np.zeros((n + 31) // 32, dtype=np.uint32) for n in ns]
    status = [np.full(n, 0xFFFFFFFF, dtype=np.uint32) for n in ns]
    rcs = np.zeros(k, dtype=np.int32)
    arr = lambda ptrs: (C.c_void_p * k)(*ptrs)
    rc = lib().p2v_verify_groups(ctx._h, k, arr([c._h if c is not None else None for c, _ in groups]),
                                 arr([_ptr(b) if b is not None else None for _, b in groups]), (C.c_size_t * k)(*ns),
                                 arr([_ptr(x) for x in bits]), arr([_ptr(x) for x in status]), rcs.ctypes.data)
    out = [(unpack_bits(bits[g], ns[g]), status[g]) for g in range(k)]
    if return_codes:
        return out, rcs
    ctx._check(rc)
    return out


def unpack_bits(words, n):
    """ceil(n/32) u32 words -> bool[n] (bit i%32 of word i/32)."""
    w = np.asarray(words, dtype=np.uint32)
    bits = ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)
    return bits[:n]


class Circuit:
    """`VerifierCircuitData` (common + verifier-only data) resident on a Context's GPU."""

    def __init__(self, ctx, shape, vkey):
        self.ctx = ctx
        self.shape = shape
        self.layout = shape_layout(shape)
        self.vkey = _u64(vkey)
        if self.vkey.size != self.layout.vkey_words:
            raise ValueError("vkey has %d words, shape expects %d" % (self.vkey.size, self.layout.vkey_words))
        self._h = C.c_void_p()
        ctx._check(lib().p2v_circuit_create(ctx._h, C.byref(shape), _ptr(self.vkey), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().p2v_circuit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _n(self, blobs, n):
        if n is not None:
            return int(n)
        return int(blobs.shape[0]) if getattr(blobs, "ndim", 1) == 2 else int(blobs.size // self.layout.blob_words)

    def proofChallenges(self, blobs, n=None, out=None):
        """`proofChallenges`, Challenge/Verifier.hs:58.  blobs AoS [n][blob_words] -> SoA [words][n]."""
        n = self._n(blobs, n)
        cw = challenges_words(self.shape)
        if out is None:
            out = np.empty((cw, n), dtype=np.uint64)
        self.ctx._check(lib().p2v_challenges(self.ctx._h, self._h, _ptr(blobs), n, _ptr(out)))
        return out

    def evalCombinedPlonkConstraints(self, blobs, n=None):
        """Plonk/Vanishing.hs:48 + Plonk/Verifier.hs:35 -> (combined SoA [2r][n], eq_ok_mask u8[n])."""
        n = self._n(blobs, n)
        r = self.shape.num_challenges
        comb = np.empty((2 * r, n), dtype=np.uint64)
        mask = np.empty(n, dtype=np.uint8)
        self.ctx._check(lib().p2v_constraints(self.ctx._h, self._h, _ptr(blobs), n, _ptr(comb), _ptr(mask)))
        return comb, mask

    def checkFRIProof(self, blobs, n=None, want_debug=False):
        """`checkFRIProof`, Plonk/FRI.hs:358 -> status u32[n] (+ per-query status, folded evals)."""
        n = self._n(blobs, n)
        q = self.shape.num_queries
        status = np.empty(n, dtype=np.uint32)
        qs = np.empty((n, q), dtype=np.uint32) if want_debug else None
        folded = np.empty((2, n * q), dtype=np.uint64) if want_debug else None
        self.ctx._check(lib().p2v_fri(self.ctx._h, self._h, _ptr(blobs), n, _ptr(status), _ptr(qs), _ptr(folded)))
        return (status, qs, folded) if want_debug else status

    def verifyProof(self, blobs, n=None, accept_bits=None, status=None):
        """`verifyProof`, Plonk/Verifier.hs:56, for a batch -> (accept bool[n], status u32[n])."""
        n = self._n(blobs, n)
        host = accept_bits is None
        if accept_bits is None:
            accept_bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        if status is None:
            status = np.empty(n, dtype=np.uint32)
        self.ctx._check(lib().p2v_verify_batch(self.ctx._h, self._h, _ptr(blobs), n, _ptr(accept_bits), _ptr(status)))
        return (unpack_bits(accept_bits, n), status) if host else (accept_bits, status)

    def verifyIntermediates(self, blobs, n=None, challenges_in=None, want_roots=True):
        """`verifyProof` with every intermediate (p2v_verify_intermediates) -> dict of host arrays: challenges [cw][n],
        combined [2r][n], eqmask [n], status [n], accept bool[n], qstatus [n][Q], folded [2][n*Q], roots [(4+steps)*4][n*Q].
        challenges_in (SoA [cw][n]) replaces the transcript's challenges (test hook)."""
        n = self._n(blobs, n)
        sh = self.shape
        cw, Q, r = challenges_words(sh), sh.num_queries, sh.num_challenges
        out = dict(challenges=np.empty((cw, n), dtype=np.uint64), combined=np.empty((2 * r, n), dtype=np.uint64),
                   eqmask=np.empty(n, dtype=np.uint8), status=np.empty(n, dtype=np.uint32),
                   qstatus=np.empty((n, Q), dtype=np.uint32), folded=np.empty((2, n * Q), dtype=np.uint64))
        bits = np.zeros((n + 31) // 32, dtype=np.uint32)
        if want_roots:
            out["roots"] = np.empty(((4 + sh.num_steps) * 4, n * Q), dtype=np.uint64)
        io = Intermediates()
        cin = None
        if challenges_in is not None:
            cin = _u64(challenges_in)
            assert cin.shape == (cw, n)
            io.challenges_in = _ptr(cin)
        io.challenges, io.combined, io.eq_ok_mask = _ptr(out["challenges"]), _ptr(out["combined"]), _ptr(out["eqmask"])
        io.status, io.accept_bits, io.query_status, io.folded = _ptr(out["status"]), _ptr(bits), _ptr(out["qstatus"]), _ptr(out["folded"])
        io.roots = _ptr(out.get("roots"))
        self.ctx._check(lib().p2v_verify_intermediates(self.ctx._h, self._h, _ptr(blobs), n, C.byref(io)))
        out["accept"] = unpack_bits(bits, n)
        return out

    def verifyProofSharded(self, blobs_local, n_total, rank, world, accept_bits_full=None, status=None):
        """`verifyProof` over a batch sharded across `world` GPUs (p2v_verify_batch_sharded): verifies this rank's slice and
        all-gathers the accept bitmap -> (bitmap of the WHOLE batch, status of the local slice).  With the default host
        outputs the call is synchronous and returns (accept bool[n_total], status u32[n_local]); with caller-supplied
        device tensors it is asynchronous on the context's stream."""
        start, stop = shard_bounds(n_total, rank, world)
        n_local = stop - start
        words_full = shard_slice_len(n_total, world) // 32 * world
        host = accept_bits_full is None
        if accept_bits_full is None:
            accept_bits_full = np.zeros(max(words_full, 1), dtype=np.uint32)
        if status is None:
            status = np.empty(max(n_local, 1), dtype=np.uint32)
        self.ctx._check(lib().p2v_verify_batch_sharded(self.ctx._h, self._h, _ptr(blobs_local) if n_local else None, int(n_total), int(rank),
                                                       int(world), _ptr(accept_bits_full), _ptr(status)))
        return (unpack_bits(accept_bits_full, n_total), status[:n_local]) if host else (accept_bits_full, status)

    def stage(self, blobs, n=None, out=None):
        """K0 alone (p2v_stage): AoS blobs -> word planes [blob_words][n]."""
        n = self._n(blobs, n)
        if out is None:
            out = np.empty((self.layout.blob_words, n), dtype=np.uint64)
        self.ctx._check(lib().p2v_stage(self.ctx._h, self._h, _ptr(blobs), n, _ptr(out)))
        return out

    def verifyProofJson(self, json_texts, threads=0):
        """What `testmain` does for one proof (src/testmain.hs:31-63), for a batch: decode the `*_proof.json` texts on
        `threads` host threads straight into pinned memory, verify on the GPU -> (accept bool[n], status u32[n],
        decode_rc i32[n]).  A text that does not decode (or does not match the circuit's shape) is reported in decode_rc and
        counted as rejected; it does not stop the batch."""
        n = len(json_texts)
        lay = self.layout
        if n == 0:
            return np.zeros(0, dtype=bool), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.int32)
        with pinned_u64(n * lay.blob_words) as buf:
            blobs, rcs = parse_proofs(json_texts, self.shape, threads=threads, out=buf.reshape(n, lay.blob_words), return_codes=True)
            accept, status = self.verifyProof(blobs, n=n)
        accept = accept & (rcs == 0)
        return accept, status, rcs

    def synth_batch(self, template_blob, n, tamper_word, tamper_delta, blobs_out):
        """Replicate + tamper a template into a device AoS batch (synthetic inputs)."""
        tw = np.ascontiguousarray(tamper_word, dtype=np.int32)
        td = _u64(tamper_delta)
        self.ctx._check(lib().p2v_synth_batch(self.ctx._h, self._h, _ptr(_u64(template_blob)), int(n), _ptr(tw),
                                              _ptr(td), _ptr(blobs_out)))
        return blobs_out
