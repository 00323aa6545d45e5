// Internal launcher declarations shared between the translation units of libreef_b200.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace reef {

// ---- poseidon.cu
void io_pattern_tag_le32(const uint32_t* ops, uint32_t n_ops, uint32_t domain_separator, uint8_t out[32]);
Fq fq_mont_from_le32(const uint8_t* b);
Fq fq_canon_from_le32(const uint8_t* b);
void poseidon_tables_host(PoseidonTables* t);
void poseidon_permute_host(Fq* s);
int poseidon_upload_constants(reef_ctx* c);
int launch_hash_batch(reef_ctx* c, const void* d_in, int arity, uint64_t n, void* d_out);
int launch_merkle(reef_ctx* c, const uint64_t* d_doc, uint64_t n_doc, void* d_levels, uint64_t* level_sizes,
                  uint32_t* n_levels_out, uint64_t idx_offset = 0);
// levels above `d_prev` (n_prev hashed nodes, canonical): ceil(n_prev/2), ... 1 nodes written to d_levels
int launch_merkle_inner(reef_ctx* c, const void* d_prev, uint64_t n_prev, void* d_levels, uint64_t* level_sizes,
                        uint32_t* n_levels_out);
int launch_p2p_allgather(reef_ctx* c, const void* src_dev, uint32_t nwords, void* dst_dev);
int launch_sponge_run(reef_ctx* c, const uint32_t* d_ops, uint32_t n_ops, const void* d_in, const uint8_t tag_le[32],
                      void* d_out);

int sponge_state_bytes();
int launch_sponge_step(reef_ctx* c, void* d_state, int op, const void* d_in, uint32_t n, const uint8_t* tag_le,
                       void* d_out);

// ---- poseidon_ro.cu (width-25 sponge over either field; field 0 = Fq, 1 = Fp)
int launch_poseidon_ro(reef_ctx* c, int field, const void* d_in, uint64_t n, int triples, void* d_out);
void poseidon_ro_host(int field, const uint8_t* elems, uint64_t n, uint8_t out[32]);
void poseidon_ro_constants_host(int field, uint8_t* rc_out, uint8_t* mds_out);
int poseidon_ro_fast_ok_host(int field);

// ---- mle.cu
struct NlookupArgs {
  int tag;                 // 0 = nl, 1 = nldoc, 2 = nlhybrid
  const void* d_table;     // device: Fq canonical (32 B) or u32 codes
  int table_is_u32;
  uint64_t n;              // padded length, power of two
  uint32_t ell;            // log2(n) (logmn semantics applied by the caller)
  uint32_t m;              // number of lookups
  const uint64_t* h_q;     // host: m indices
  const uint8_t* h_query;  // host: canonical 32-byte elements absorbed first (already ordered)
  uint32_t n_query;
  const uint8_t* h_prev_q; // host: ell x 32 bytes (prev_running_q, index 0 <-> MSB)
  uint8_t tag_le[32];      // IOPattern tag
  // outputs (host)
  uint8_t* out_claim_r;    // 32
  uint8_t* out_rounds;     // ell x 4 x 32: (sc_r, xsq, x, const)
  uint8_t* out_last_claim; // 32
  uint8_t* out_next_v;     // 32
  cudaEvent_t table_ready = nullptr;   // optional: event behind an asynchronous upload of d_table
  // (f1) optional: the same outputs scattered on the device into an index-addressed witness buffer (canonical
  // elements) right behind the last kernel; slot = UINT64_MAX skips an output
  void* d_wit = nullptr;
  uint64_t wit_len = 0;
  uint64_t slot_claim_r = ~0ull, slot_rounds = ~0ull, slot_last_claim = ~0ull, slot_next_v = ~0ull;
};
int nlookup_run(reef_ctx* c, const NlookupArgs& a);
int launch_hybrid_table(reef_ctx* c, const void* d_pub, uint64_t n_pub, const uint8_t* fill_le, uint64_t half_len, const uint32_t* d_codes,
                        uint64_t n_doc, void* d_out);
int launch_gen_eq_table(reef_ctx* c, const uint8_t* h_rs, const uint64_t* h_qs, uint32_t m, const uint8_t* h_last_q,
                        uint32_t ell, void* d_out);
int launch_mle_eval(reef_ctx* c, const void* d_table, int is_u32, uint64_t n, const uint8_t* h_x, uint32_t ell,
                    uint8_t* h_out);
int launch_mle_round_coeffs(reef_ctx* c, const void* d_t, const void* d_eq, uint32_t ell, uint32_t i, uint8_t* h_out3);
int launch_mle_round_fold(reef_ctx* c, void* d_t, void* d_eq, uint32_t ell, uint32_t i, const uint8_t* h_r);

}  // namespace reef
struct reef_nl_session;
namespace reef {
int nl_shard_begin(reef_ctx* c, const NlookupArgs& a, uint32_t rank, uint32_t world, reef_nl_session** out);
int nl_shard_round_local(reef_nl_session* s, void* d_out3);
int nl_shard_round_finish(reef_nl_session* s, const void* d_triples);
int nl_shard_export(reef_nl_session* s, void* d_out2);
int nl_shard_finish(reef_nl_session* s, const void* d_pairs, uint8_t* out_claim_r, uint8_t* out_rounds, uint8_t* out_last_claim,
                    uint8_t* out_next_v);
void nl_shard_free(reef_nl_session* s);   // caller holds the context lock and drops the session's context reference
reef_ctx* nl_shard_ctx(reef_nl_session* s);
int nl_shard_preload();
int nl_shard_round_p2p(reef_nl_session* s);
int nl_shard_finish_p2p(reef_nl_session* s, uint8_t* out_claim_r, uint8_t* out_rounds, uint8_t* out_last_claim, uint8_t* out_next_v);
int launch_lz(reef_ctx* c, const void* d_matrix, int is_u32, uint64_t rows, uint64_t cols, const uint8_t* h_L,
              uint8_t* h_out);

// ---- msm.cu
struct MsmPlanPublic {
  uint32_t c, W, L, G, B;
};
MsmPlanPublic msm_make_plan(uint64_t n, uint32_t scalar_bits, uint64_t max_level_bytes);
int msm_bases_register(reef_ctx* c, int curve, const uint8_t* h_bases, uint64_t n, const MsmPlanPublic& pl, void** d_levels);
struct MsmRunArgs {
  MsmPlanPublic plan;
  const void* d_levels;      // Affine[L][n_bases], Montgomery
  uint64_t n_bases;
  const void* d_scalars;     // device: n x 32 B canonical, or n x u32
  int scalars_u32;
  uint64_t n;                // number of terms (<= n_bases)
  uint32_t w_begin, w_end;   // window range handled by this call (multi-GPU split)
  uint8_t* h_out_affine;     // 64 B canonical or NULL
  uint8_t* h_out_xyzz;       // 128 B canonical XYZZ or NULL
  const void* h_extra_xyzz_mont;
  uint32_t n_extra;
  int p2p_combine = 0;       // 1: all-gather the partial over the context's mailboxes and combine on the device
};
int msm_run(reef_ctx* c, int curve, const MsmRunArgs& a);
int msm_combine(reef_ctx* c, int curve, const uint8_t* h_pts, uint32_t k, uint8_t* h_out);
int msm_preload();
int msm_levels_from_dev(reef_ctx* c, int curve, const void* d_canon, uint64_t n, const MsmPlanPublic& pl, void* d_levels, int* d_bad);
struct MsmRowsArgs {
  MsmPlanPublic plan;
  const void* d_levels;
  uint64_t n_bases;
  const void* d_scalars;     // device: rows x cols, u32 or 32 B canonical, row-major
  int scalars_u32;
  uint32_t scalar_bits;      // upper bound on the matrix entries' bit length
  uint64_t rows, cols;
  const void* d_blinds;      // device: rows x 32 B or NULL
  uint64_t blind_base;       // index of the blinding generator among the registered bases
  uint8_t* h_out;            // rows x 64 B
  void** d_rows_out = nullptr; // optional: receives the device address of the affine rows (in the context's scratch)
};
int msm_rows_run(reef_ctx* c, int curve, const MsmRowsArgs& a);

}  // namespace reef

// registered generators (reef_bases_register): shared by api.cu and the IPA session of sumcheck.cu
struct reef_bases {
  reef_ctx* ctx;
  int curve;
  uint64_t n;
  uint32_t scalar_bits;
  reef::MsmPlanPublic plan;
  void* d_levels;
};
