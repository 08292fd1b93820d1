/*
 * p2v.h — C ABI of the B200-native batch Plonky2 verifier (libp2v.so).
 *
 * This is the drop-in boundary for the ONE hot path of bkomuves/plonky2-verifier:
 * batch proof verification (Poseidon/Merkle -> Fiat-Shamir challenger -> constraint
 * check at zeta -> FRI check).  The reference has no FFI of its own (it is pure
 * Haskell); every entry point below names the pure Haskell function it replaces
 * (file:line relative to the reference's src/), i.e. the function a
 * `foreign import ccall` shim would wrap.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary;
 *  - field elements are uint64_t.  Inputs may be any u64 (they are reduced mod
 *    p = 2^64-2^32+1 exactly like `mkGoldilocks`, Algebra/Goldilocks.hs:132);
 *    every output is the canonical representative in [0,p);
 *  - every data pointer may be a HOST pointer or a DEVICE pointer of the context's
 *    GPU; the library detects which (cudaPointerGetAttributes) and stages host
 *    buffers through the context's stream;
 *  - "SoA [k][n]" means k planes of n contiguous elements: element (j, i) is at
 *    ptr[j*n + i];
 *  - return value: 0 on success, a negative P2V_E_* code otherwise;
 *    p2v_last_error() gives the message.  No exception crosses the boundary;
 *  - one CUDA stream per context (non-blocking: it does NOT synchronise with the legacy
 *    default stream); calls on one context are serialised, different contexts may be used
 *    from different threads.  Device buffers handed to the library must be complete with
 *    respect to that stream: a caller that filled them on another stream synchronises
 *    first (or records an event and makes p2v_ctx_stream wait on it);
 *  - there is NO CPU fallback: without a usable sm_100 GPU p2v_ctx_create fails.
 */
#ifndef P2V_H
#define P2V_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2V_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------------ */
#define P2V_OK 0
#define P2V_E_INVALID (-1)  /* bad argument                                       */
#define P2V_E_CUDA (-2)     /* CUDA runtime error (message in p2v_last_error)      */
#define P2V_E_NOGPU (-3)    /* no usable GPU: there is no CPU fallback             */
#define P2V_E_PARSE (-4)    /* malformed JSON / gate string                        */
#define P2V_E_SHAPE (-5)    /* input does not have the circuit's shape             */
#define P2V_E_UNSUPPORTED (-6) /* reference `error` sites for unsupported features */
#define P2V_E_NOMEM (-7)

/* ---- limits of the fixed-size shape record (ABI 2: twice to four times the sizes of ABI 1; Types.hs:47-70 has unbounded
 * lists — a circuit beyond these limits is refused with P2V_E_UNSUPPORTED, never truncated) ------------------------- */
#define P2V_MAX_GATES 64
#define P2V_MAX_GROUPS 16
#define P2V_MAX_ROUTED 256
#define P2V_MAX_STEPS 16
#define P2V_MAX_LUTS 16
#define P2V_MAX_WEIGHTS 256
#define P2V_MAX_CHALLENGES 4 /* num_challenges the kernels keep in registers (Plonky2's configurations use 2 or 3) */

/* Gate kinds: the constructors of `data Gate`, Gate/Base.hs:27-45. */
enum p2v_gate_kind {
  P2V_GATE_ARITHMETIC = 0,      /* p0 = num_ops                                   */
  P2V_GATE_ARITHMETIC_EXT = 1,  /* p0 = num_ops                                   */
  P2V_GATE_BASE_SUM = 2,        /* p0 = num_limbs, p1 = base                      */
  P2V_GATE_COSET_INTERP = 3,    /* p0 = subgroup_bits, p1 = degree, weights       */
  P2V_GATE_CONSTANT = 4,        /* p0 = num_consts                                */
  P2V_GATE_EXPONENTIATION = 5,  /* p0 = num_power_bits                            */
  P2V_GATE_LOOKUP = 6,          /* p0 = num_slots                                 */
  P2V_GATE_LOOKUP_TABLE = 7,    /* p0 = num_slots, p1 = last_lut_row              */
  P2V_GATE_MUL_EXT = 8,         /* p0 = num_ops                                   */
  P2V_GATE_NOOP = 9,
  P2V_GATE_PUBLIC_INPUT = 10,
  P2V_GATE_POSEIDON = 11,       /* p0 = width (12)                                */
  P2V_GATE_POSEIDON_MDS = 12,   /* p0 = width (12)                                */
  P2V_GATE_RANDOM_ACCESS = 13,  /* p0 = bits, p1 = num_copies, p2 = num_extra_constants */
  P2V_GATE_REDUCING = 14,       /* p0 = num_coeffs                                */
  P2V_GATE_REDUCING_EXT = 15,   /* p0 = num_coeffs                                */
  P2V_GATE_UNKNOWN = 16
};

typedef struct p2v_gate {
  int32_t kind;            /* enum p2v_gate_kind                                  */
  int32_t p0, p1, p2;      /* parameters, see above                               */
  int32_t group;           /* selector_indices[k]      (Types.hs:97-101)          */
  int32_t num_constraints; /* length of the gate's committed constraint list      */
  int32_t weights_off;     /* CosetInterpolationGate: first weight in `weights`   */
  int32_t weights_len;
} p2v_gate;

/*
 * p2v_shape — everything `CommonCircuitData` (Types.hs:47-70) says about a circuit,
 * flattened to a POD.  A batch is shape-homogeneous.
 */
typedef struct p2v_shape {
  /* CircuitConfig, Types.hs:73-87 */
  int32_t num_wires;
  int32_t num_routed_wires;
  int32_t num_gate_constants; /* config_num_constants                            */
  int32_t num_challenges;
  /* FriConfig / FriParams, Types.hs:116-173 */
  int32_t degree_bits;
  int32_t rate_bits;
  int32_t cap_height;
  int32_t pow_bits;
  int32_t num_queries;
  int32_t num_steps;                       /* expandReductionStrategy, Plonk/FRI.hs:337-354 */
  int32_t step_arity_bits[P2V_MAX_STEPS];
  int32_t final_poly_len;                  /* 2^(degree_bits - sum(arity_bits))   */
  /* CommonCircuitData */
  int32_t quotient_degree_factor;
  int32_t num_constants;                   /* all constant columns                */
  int32_t num_public_inputs;
  int32_t num_partial_products;
  int32_t num_lookup_polys;
  int32_t num_lookup_selectors;
  int32_t num_gates;
  int32_t num_groups;
  int32_t group_start[P2V_MAX_GROUPS];
  int32_t group_end[P2V_MAX_GROUPS];
  p2v_gate gates[P2V_MAX_GATES];
  int32_t num_weights;
  uint64_t weights[P2V_MAX_WEIGHTS];       /* barycentric weights of all CosetInterpolation gates */
  uint64_t k_is[P2V_MAX_ROUTED];
  /* lookup tables, Types.hs:28-36: lut k = pairs lut_pairs[2*lut_off[k] .. 2*lut_off[k+1]) as (inp,out) */
  int32_t num_luts;
  int32_t lut_off[P2V_MAX_LUTS + 1];
  const uint64_t *lut_pairs;               /* owned by whoever built the shape    */
} p2v_shape;

/*
 * p2v_layout — the flat u64 "proof blob" a shape implies: the fields of
 * ProofWithPublicInputs (Types.hs:251-279) in declaration order.
 *
 *   proof part (per proof), word offsets:
 *     wires_cap | zs_pp_cap | quotient_cap                    each 4*2^cap_height
 *     openings: constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next,
 *               partial_products, quotient_polys, lookup_zs, lookup_zs_next   (2 words per FExt)
 *     commit_phase_merkle_caps[num_steps]                      each 4*2^cap_height
 *     final_poly (2*final_poly_len) | pow_witness (1) | public_inputs
 *   then num_queries query parts of `query_words` words each:
 *     for oracle o in constants,witness,zs_pp_lookup,quotient:  leaf[width_o] | siblings[4*init_path_len]
 *     for step s: evals[2*2^arity_s] | siblings[4*step_path_len[s]]
 */
typedef struct p2v_layout {
  int32_t cap_words;
  int32_t off_wires_cap, off_zs_pp_cap, off_quotient_cap;
  int32_t off_open_constants, off_open_sigmas, off_open_wires, off_open_zs, off_open_zs_next;
  int32_t off_open_pp, off_open_quotient, off_open_lookup_zs, off_open_lookup_zs_next;
  int32_t n_open_constants, n_open_sigmas, n_open_wires, n_open_zs, n_open_zs_next;
  int32_t n_open_pp, n_open_quotient, n_open_lookup_zs, n_open_lookup_zs_next;
  int32_t off_commit_caps, off_final_poly, off_pow_witness, off_public_inputs;
  int32_t proof_words;                 /* size of the per-proof part                */
  int32_t oracle_width[4];             /* oracleWidths, Plonk/FRI.hs:56-65          */
  int32_t init_path_len;               /* degree_bits + rate_bits - cap_height      */
  int32_t q_off_leaf[4], q_off_sibs[4];
  int32_t q_off_step_evals[P2V_MAX_STEPS], q_off_step_sibs[P2V_MAX_STEPS];
  int32_t step_path_len[P2V_MAX_STEPS];
  int32_t query_words;                 /* size of one query part                    */
  int32_t blob_words;                  /* proof_words + num_queries*query_words     */
  int32_t vkey_words;                  /* 4*2^cap_height + 4 (cap, circuit_digest)  */
} p2v_layout;

/*
 * Per-proof verdict.  status word = code | (query << 8) | (detail << 16).
 * The accept bit of a proof is (code == P2V_ST_ACCEPT).  Codes >= 16 are the reference's
 * `error` sites (imprecise exceptions); codes 1..3 are `False`.
 */
enum p2v_status_code {
  P2V_ST_ACCEPT = 0,
  P2V_ST_FALSE_EQS = 1,    /* checkCombinedPlonkEquations False (Plonk/Verifier.hs:62-64); detail = mask of failing rounds */
  P2V_ST_FALSE_POW = 2,    /* checkProofOfWork False (Plonk/FRI.hs:212-216)        */
  P2V_ST_FALSE_FINAL = 3,  /* final_poly_eval /= folded (Plonk/FRI.hs:407); query  */
  P2V_ST_ERR_INIT_MERKLE = 16, /* Plonk/FRI.hs:108; query, detail = mask of failing oracles */
  P2V_ST_ERR_STEP_MERKLE = 17, /* Plonk/FRI.hs:310; query, detail = step           */
  P2V_ST_ERR_STEP_EVAL = 18    /* Plonk/FRI.hs:311; query, detail = step           */
};

/* Challenges of one proof, SoA [p2v_challenges_words(shape)][n]; order (Challenge/Verifier.hs:45-53,
 * Challenge/FRI.hs:24-30): betas[r], gammas[r], alphas[r], deltas[4r or 0], zeta[2], fri_alpha[2],
 * fri_betas[2*num_steps], pow_response[1], query_indices[num_queries]. */

typedef struct p2v_ctx p2v_ctx;
typedef struct p2v_circuit p2v_circuit;

/* ---- context ----------------------------------------------------------------- */
int p2v_abi_version(void);
int p2v_ctx_create(int device, p2v_ctx **out);
void p2v_ctx_destroy(p2v_ctx *ctx);
const char *p2v_last_error(const p2v_ctx *ctx); /* ctx may be NULL: last error of this thread */
void *p2v_ctx_stream(p2v_ctx *ctx);              /* the context's cudaStream_t                 */
int p2v_ctx_sync(p2v_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t p2v_ctx_launch_count(const p2v_ctx *ctx);
/* pinned host memory for callers that want asynchronous staging */
int p2v_host_alloc(size_t bytes, void **out);
void p2v_host_free(void *p);

/* ---- L2 Hash ------------------------------------------------------------------ */
/* `permutation :: State -> State`, Hash/Poseidon.hs:42-46.  in/out SoA [12][n]. */
int p2v_poseidon_permute(p2v_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n);
/* `sponge :: [F] -> Digest`, Hash/Sponge.hs:26-31.  leaves SoA [w][n] -> digests SoA [4][n]. */
int p2v_hash_leaves(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, size_t n, uint64_t *digests);
/* `compress :: Digest -> Digest -> Digest`, Hash/Merkle.hs:21-23.  SoA [4][n] each. */
int p2v_compress(p2v_ctx *ctx, const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n);
/* `checkMerkleProof cap idx leaf proof`, Hash/Merkle.hs:27-42, for n openings against ONE cap.
 *   leaves SoA [w][n]; idx [n]; siblings SoA [path_len*4][n] (sibling l, word k at plane l*4+k);
 *   cap [2^cap_height][4] row-major; ok_bits: ceil(n/32) words, bit i%32 of word i/32;
 *   roots_out (optional, may be NULL) SoA [4][n] = the reconstructed cap entry. */
int p2v_merkle_verify(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, const uint32_t *idx,
                      const uint64_t *siblings, uint32_t path_len, const uint64_t *cap,
                      uint32_t cap_height, size_t n, uint32_t *ok_bits, uint64_t *roots_out);
/* Build a Poseidon Merkle tree over 2^log_n leaves (SoA [w][2^log_n]); the prover-side dual of
 * Hash/Merkle.hs used to make synthetic trees.  digests_out: SoA levels, level 0 = leaf digests
 * [4][2^log_n], level l = [4][2^(log_n-l)], concatenated up to and including the cap level
 * (log_n - cap_height); total 4*(2^(log_n+1) - 2^cap_height) words. */
int p2v_merkle_build(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, uint32_t log_n,
                     uint32_t cap_height, uint64_t *digests_out);
/* Gather n openings from a tree built by p2v_merkle_build: leaves_out SoA [w][n],
 * siblings_out SoA [(log_n-cap_height)*4][n], cap_out [2^cap_height][4].  A buffer of zero size (w == 0, or
 * log_n == cap_height: the cap is the leaf level) may be NULL. */
int p2v_merkle_open(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, uint32_t log_n,
                    uint32_t cap_height, const uint64_t *digests, const uint32_t *idx, size_t n,
                    uint64_t *leaves_out, uint64_t *siblings_out, uint64_t *cap_out);

/* ---- host side: JSON wire format (Types.hs:47-279, Gate/Parser.hs:107-240) ---- */
/* `FromJSON CommonCircuitData`.  On success *out is filled; free with p2v_shape_free. */
int p2v_parse_common(const char *json, size_t len, p2v_shape *out);
void p2v_shape_free(p2v_shape *shape);
/* `recognizeGate`, Gate/Parser.hs:107.  weights: out array of P2V_MAX_WEIGHTS. */
int p2v_parse_gate(const char *str, size_t len, p2v_gate *out, uint64_t *weights);
int p2v_shape_layout(const p2v_shape *shape, p2v_layout *out);
/* Vet a circuit description without a GPU: every integer that enters an offset, a loop bound or a shift is in range, and every
 * gate is one the kernels implement with parameters that fit the circuit (the reference's `error` / array-index sites:
 * Gate/Constraints.hs:93-108, Gate/Selector.hs:85).  P2V_OK, or P2V_E_SHAPE / P2V_E_UNSUPPORTED with the reason in
 * p2v_last_error(NULL).  p2v_circuit_create runs the same check. */
int p2v_shape_check(const p2v_shape *shape);
int p2v_challenges_words(const p2v_shape *shape);
/* `FromJSON VerifierOnlyCircuitData`: out[vkey_words] = cap ++ circuit_digest. */
int p2v_parse_vkey(const char *json, size_t len, const p2v_shape *shape, uint64_t *out);
/* `FromJSON ProofWithPublicInputs`: out[blob_words]; P2V_E_SHAPE when a list length differs
 * from the circuit's shape (the `error`s of safeZip, buildListOracle and validateMerkleCapLength in the reference). */
int p2v_parse_proof(const char *json, size_t len, const p2v_shape *shape, uint64_t *out);
/* The same for a batch (`mapM decodeProof` over the files testmain would read one by one, src/testmain.hs:33):
 * proof i = jsons[i][0..lens[i]) -> blobs[i * blob_words ..], decoded on `threads` host threads (<= 0: all
 * hardware threads).  rcs (may be NULL) receives each proof's P2V_OK / P2V_E_PARSE / P2V_E_SHAPE; a proof that
 * fails to decode leaves its blob zeroed and does NOT stop the others.  Returns P2V_OK when every proof decoded,
 * else the error code of the first failing proof (its message in p2v_last_error(NULL)).  blobs may be pinned
 * memory from p2v_host_alloc, so that p2v_verify_batch copies straight from it. */
int p2v_parse_proofs(const char *const *jsons, const size_t *lens, size_t n, const p2v_shape *shape, uint64_t *blobs,
                     int threads, int32_t *rcs);

/* ---- circuits and batches ------------------------------------------------------ */
/* VerifierCircuitData (Types.hs: verifier_only + verifier_common) resident on the GPU. */
int p2v_circuit_create(p2v_ctx *ctx, const p2v_shape *shape, const uint64_t *vkey, p2v_circuit **out);
void p2v_circuit_destroy(p2v_circuit *c);

/* `proofChallenges`, Challenge/Verifier.hs:58-103 (+ `friChallenges`, Challenge/FRI.hs:65-104).
 *   blobs: AoS [n][blob_words]; challenges_out: SoA [p2v_challenges_words][n]. */
int p2v_challenges(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n,
                   uint64_t *challenges_out);
/* `evalCombinedPlonkConstraints` (Plonk/Vanishing.hs:48-51) + `checkCombinedPlonkEquations'`
 * (Plonk/Verifier.hs:35-52).  combined_out: SoA [num_challenges*2][n] (may be NULL);
 * eq_ok_mask: [n] bytes, bit i = round i holds (may be NULL). */
int p2v_constraints(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n,
                    uint64_t *combined_out, uint8_t *eq_ok_mask);
/* `checkFRIProof`, Plonk/FRI.hs:358-407.  status [n] uses the P2V_ST_* encoding restricted to the
 * FRI codes; query_status (optional) [n][num_queries] per-query codes; debug outputs optional:
 * folded_out SoA [2][n*num_queries] = final folding_upstream_eval of every query. */
int p2v_fri(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint32_t *status,
            uint32_t *query_status, uint64_t *folded_out);
/* `verifyProof`, Plonk/Verifier.hs:56-65, for n proofs.
 *   accept_bits: ceil(n/32) words; status: [n] (may be NULL). */
int p2v_verify_batch(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n,
                     uint32_t *accept_bits, uint32_t *status);
/* `verifyProof` with every intermediate the north star names — challenges, combined constraint values, quotient-identity
 * verdicts, per-query status, folded evaluations, RECOMPUTED MERKLE ROOTS — in one call; every pointer is optional.
 *   challenges_in  test hook: SoA [p2v_challenges_words][n] challenges that REPLACE the Fiat-Shamir transcript's, so that
 *                  branches no honest transcript reaches can be compared with the oracle (zeta = 1 in `evalLagrange0`,
 *                  Algebra/Poly.hs:14-17; x = zeta in `combineInitial`, Plonk/FRI.hs:151-207 with `inv 0 = 0`,
 *                  Algebra/Goldilocks.hs:155); query indices are taken mod 2^lde_bits;
 *   roots          SoA [(4+num_steps)*4][n*num_queries]: word i of the root `reconstructMerkleRoot'` (Hash/Merkle.hs:30-37)
 *                  yields for tree t (0..3 the initial oracles, 4.. the folding steps) of (proof p, query q) at
 *                  plane t*4+i, index p*num_queries+q — computed whatever the verdict is. */
typedef struct p2v_intermediates {
  const uint64_t *challenges_in;
  uint64_t *challenges;   /* SoA [p2v_challenges_words][n]                    */
  uint64_t *combined;     /* SoA [2*num_challenges][n]                        */
  uint8_t *eq_ok_mask;    /* [n]                                              */
  uint32_t *status;       /* [n] verifyProof verdict                          */
  uint32_t *accept_bits;  /* ceil(n/32) words                                 */
  uint32_t *query_status; /* [n][num_queries]                                 */
  uint64_t *folded;       /* SoA [2][n*num_queries]                           */
  uint64_t *roots;        /* SoA [(4+num_steps)*4][n*num_queries]             */
} p2v_intermediates;
int p2v_verify_intermediates(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, const p2v_intermediates *io);

/* Heterogeneous batches: proofs of DIFFERENT circuits (other degree_bits, gate sets, FRI parameters,
 * lookups ...) grouped by circuit; group g gets exactly the verdicts of p2v_verify_batch(ctx, circuits[g], blobs[g],
 * counts[g], accept_bits[g], status[g]) — `map (uncurry verifyProof)` over (vkey, proof) pairs that do not share a vkey.
 * The chunks of ALL groups share the lanes of the context's pipeline, so many small groups run side by side (a kernel
 * launch is shape-homogeneous because the circuit is a kernel parameter; the call is not).  A group that cannot run (NULL
 * pointer, circuit of another context, oversized blob) does not stop the others: rcs[g] (may be NULL) receives each
 * group's P2V_OK / error code and the call returns the first group error (P2V_OK if none). */
int p2v_verify_groups(p2v_ctx *ctx, size_t n_groups, const p2v_circuit *const *circuits, const uint64_t *const *blobs,
                      const size_t *counts, uint32_t *const *accept_bits, uint32_t *const *status, int32_t *rcs);

/* ---- multi-GPU: one process per GPU, contiguous slices, ONE all-gather of the accept bitmap ------ */
/* The slicing rule.  slice_len = ceil(n_total / world) rounded up to a multiple of 32 (bitmap words never straddle two
 * ranks); rank r owns proofs [min(n_total, r*slice_len), min(n_total, (r+1)*slice_len)) — the last ranks may be short
 * or empty. */
size_t p2v_shard_slice_len(size_t n_total, int world);
int p2v_shard_bounds(size_t n_total, int rank, int world, size_t *start, size_t *stop);
/* Communicator bootstrap for hosts that have none (NCCL, bound at run time: libnccl.so.2).  Rank 0 calls
 * p2v_nccl_unique_id (128 bytes), ships the id to the other ranks by any means (a file, a socket, MPI, torch's store),
 * then every rank calls p2v_nccl_init (collective).  The communicator belongs to the context and dies with it. */
#define P2V_NCCL_UNIQUE_ID_BYTES 128
int p2v_nccl_unique_id(void *out128);
int p2v_nccl_init(p2v_ctx *ctx, const void *id128, int rank, int world);
/* ... or use an ncclComm_t the caller owns (rank and world are read from it; it is not destroyed by the library). */
int p2v_nccl_attach(p2v_ctx *ctx, void *nccl_comm);
int p2v_nccl_finalize(p2v_ctx *ctx);
/* rank / world of the attached communicator (0 / 1 without one) and the NCCL version in use (0: library not found) */
int p2v_nccl_info(p2v_ctx *ctx, int *rank, int *world, int *nccl_version);
/* `verifyProof` (Plonk/Verifier.hs:56-65) mapped over a batch of n_total proofs that is sharded over `world` GPUs:
 * this rank verifies its slice (blobs_local: AoS [stop-start][blob_words], host or device) and the packed accept bits
 * are all-gathered with ncclAllGather on the context's stream (in place, no host synchronisation in between).
 *   accept_bits_full: world * slice_len / 32 words, host or device; on return its first ceil(n_total/32) words are the
 *                     accept bitmap of the WHOLE batch — identical on every rank and identical to what one GPU
 *                     computes for the whole batch; padding bits are 0;
 *   status_local:     [stop-start] status words of this rank's proofs (may be NULL).
 * Collective: every rank of the communicator calls it with the same n_total.  world == 1 needs no communicator. */
int p2v_verify_batch_sharded(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs_local, size_t n_total, int rank,
                             int world, uint32_t *accept_bits_full, uint32_t *status_local);

/* Opt-in B200-native form of the gather (collective; needs the communicator once, for the handle exchange): every rank
 * exports a small gather buffer with cudaIpc, and from then on p2v_verify_batch_sharded publishes its slice with direct
 * stores into every peer's buffer over NVLink/NVSwitch plus a system-scope release flag, and waits for the peers' flags —
 * no NCCL kernel on the path.  All ranks on one node with peer access; at most 16 ranks and 2^18 bitmap words per call
 * (larger calls fall back to ncclAllGather).  Results are identical (tests/sharded_worker.py). */
int p2v_peer_enable(p2v_ctx *ctx);
int p2v_peer_disable(p2v_ctx *ctx);

/* K0 alone: AoS blobs [n][blob_words] -> structure-of-arrays word planes, the layout every kernel reads
 * (replaces the Haskell lists of Types.hs:251-279).  planes_out (device or host): [blob_words][n] with plane order
 * pp[w] for w < proof_words, then qp[wq * num_queries + q] for wq < query_words. */
int p2v_stage(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint64_t *planes_out);

/* Proofs per chunk (SoA workspace per pipeline lane ~ 1.1 * chunk * blob_words * 8 bytes).  0 = default: 3 GiB of
 * blobs for device-resident input, 0.5 GiB for host input (16 GiB / 2 GiB with p2v_ctx_set_pipeline(ctx, 1)). */
int p2v_ctx_set_chunk(p2v_ctx *ctx, size_t proofs_per_chunk);
/* Chunk pipelining: depth 2..4 (default 4) runs consecutive chunks round-robin on that many streams and
 * workspaces, so the latency-bound per-proof kernels (K0, K4, K5) of the next chunk(s) overlap the Merkle kernel of
 * the current one; depth 1 is strictly serial and is the mode in which p2v_ctx_last_ms is meaningful.
 * Results are identical. */
int p2v_ctx_set_pipeline(p2v_ctx *ctx, int depth);

/* Synthetic batches (north_star: "synthetic batches built from the bundled JSON proofs"):
 * replicate `template_blob` n times into blobs_out (device AoS [n][blob_words]) and add
 * tamper_delta[i] (mod p) to word tamper_word[i] of copy i (tamper_word[i] < 0: untouched). */
int p2v_synth_batch(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *template_blob, size_t n,
                    const int32_t *tamper_word, const uint64_t *tamper_delta, uint64_t *blobs_out);

/* ---- test hooks ------------------------------------------------------------------ */
/* The device field routines one by one (Algebra/Goldilocks.hs:140-175, GoldilocksExt.hs:54-99) on caller-chosen operands;
 * inputs may be any u64 (lazy representation), outputs are canonical.  Base field, arrays [n]:
 *   0 a+b, 1 a-b, 2 a*b, 3 inv a (inv 0 = 0), 4 -a, 5 a * (u32)b, 6 (b:a) mod p as a 128-bit value, 7 a^b, 8 canonical a,
 *   9 a^7 (s-box).  Extension field, arrays SoA [2][n]: 16 a*b, 17 inv a, 18 a+b, 19 a-b, 20 a^2, 21 b.re * a, 22 a*X,
 *   23 a^(b.re), 24 -a.  b may be NULL for unary operations. */
int p2v_debug_field_op(p2v_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);

/* ---- measurement helpers ------------------------------------------------------- */
/* Integer-pipe peak microbenchmark: dependent chains whose every step is a GROUP of instructions:
 * mode 0: LOP3 + IMAD.WIDE.U32, 1: 2 LOP3 + IMAD.WIDE.U32, 2: LOP3 + IMAD (32-bit), 3: LOP3 + IADD3,
 * 4: 2 IMAD.WIDE.U32 + LOP3, 5: 2 IMAD (32-bit), 6: 2 LOP3, 7: IMAD.WIDE.U32 alone, 8: 2 SHF, 9: IADD3 + IADD3.X,
 * 10: DFMA, 11: DFMA + LOP3 + IMAD, 12: DADD, 13: I2F.F64.U32 + LOP3, 14: mode 13 next to a DFMA chain,
 * 15: mode 13 next to an IMAD.WIDE.U32 chain.  16..23: operand-bandwidth probes with three distinct register sources per
 * instruction (hash_kernels.cuh k_rf_probe): 16 IADD3, 17 LOP3, 18 DFMA (3 registers), 19 DFMA (2 registers + immediate),
 * 20 = 18 + 16, 21 = 19 + 16, 22 IMAD.WIDE with a 64-bit accumulator, 23 = 22 + 19 + 16.
 * *ops_per_s = groups per second summed over all threads. */
int p2v_int_pipe_peak(p2v_ctx *ctx, int mode, double *ops_per_s);
/* Device time (ms) of a section of the most recent batch call (last chunk): "stage" (K0), "challenges" (K4),
 * "constraints" (K5), "fri" (K6a+K6b), "fri_merkle" (K6a alone), "verdict" (K7). */
int p2v_ctx_last_ms(p2v_ctx *ctx, const char *section, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* P2V_H */
