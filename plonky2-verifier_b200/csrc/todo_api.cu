// TEMPORARY: entry points not implemented yet (removed as they land).
#include "ctx.hpp"
#define TODO(name) return p2v_fail(nullptr, P2V_E_UNSUPPORTED, name ": not implemented yet")
extern "C" {
int p2v_parse_common(const char *, size_t, p2v_shape *) { TODO("p2v_parse_common"); }
void p2v_shape_free(p2v_shape *) {}
int p2v_parse_gate(const char *, size_t, p2v_gate *, uint64_t *) { TODO("p2v_parse_gate"); }
int p2v_shape_layout(const p2v_shape *, p2v_layout *) { TODO("p2v_shape_layout"); }
int p2v_challenges_words(const p2v_shape *) { return 0; }
int p2v_parse_vkey(const char *, size_t, const p2v_shape *, uint64_t *) { TODO("p2v_parse_vkey"); }
int p2v_parse_proof(const char *, size_t, const p2v_shape *, uint64_t *) { TODO("p2v_parse_proof"); }
int p2v_circuit_create(p2v_ctx *, const p2v_shape *, const uint64_t *, p2v_circuit **) { TODO("p2v_circuit_create"); }
void p2v_circuit_destroy(p2v_circuit *) {}
int p2v_challenges(p2v_ctx *, const p2v_circuit *, const uint64_t *, size_t, uint64_t *) { TODO("p2v_challenges"); }
int p2v_constraints(p2v_ctx *, const p2v_circuit *, const uint64_t *, size_t, uint64_t *, uint8_t *) { TODO("p2v_constraints"); }
int p2v_fri(p2v_ctx *, const p2v_circuit *, const uint64_t *, size_t, uint32_t *, uint32_t *, uint64_t *) { TODO("p2v_fri"); }
int p2v_verify_batch(p2v_ctx *, const p2v_circuit *, const uint64_t *, size_t, uint32_t *, uint32_t *) { TODO("p2v_verify_batch"); }
int p2v_synth_batch(p2v_ctx *, const p2v_circuit *, const uint64_t *, size_t, const int32_t *, const uint64_t *, uint64_t *) { TODO("p2v_synth_batch"); }
}
