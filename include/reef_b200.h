/* reef_b200.h -- C ABI of libreef_b200.so: the B200-native prover hot path for eniac/Reef.
 *
 * The reference (pure Rust) has no FFI seam on this path; these entry points are the seams a
 * maintainer binds with `extern "C"` from src/backend.  Each one names
 * the reference interface it replaces.  Citations are relative to /root/reference/.
 *
 * Conventions
 *   - Field elements (Fq = Pallas scalar field, modulus at src/backend/r1cs_helper.rs:37-38):
 *     32-byte little-endian canonical integers, exactly `PrimeField::to_repr()` /
 *     `Integer::from_digits(.., Order::Lsf)` (r1cs_helper.rs:488, commitment.rs:528).
 *   - Curve points: affine, x || y, 64 bytes, each coordinate little-endian canonical in the
 *     curve's base field; the point at infinity is 64 zero bytes.
 *   - All buffers are caller-owned HOST memory unless the name ends in `_dev`.
 *   - Return value: 0 = OK; non-zero = failure, message from reef_last_error() (thread-local).
 *     The reference panics/asserts on this path (framework.rs:394-396, commitment.rs:269); the
 *     Rust shim turns a non-zero return into `panic!` to keep that behaviour.  REEF_EASSERT is
 *     returned exactly where the reference would have hit an `assert!`/index panic.
 *   - One context per calling thread (the reference calls this path from two threads,
 *     framework.rs:98-110); a context serialises its own calls and owns one CUDA stream.
 *   - No CPU fallback: every compute entry fails with REEF_ECUDA when no sm_100 device is usable.
 */
#ifndef REEF_B200_H
#define REEF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REEF_OK 0
#define REEF_EINVAL 1  /* malformed arguments */
#define REEF_ECUDA 2   /* CUDA runtime / no device */
#define REEF_EASSERT 3 /* the reference would have panicked (assert!, index out of bounds) */
#define REEF_ENOMEM 4

typedef struct reef_ctx reef_ctx;
typedef struct reef_table reef_table;   /* device-resident lookup table (T or the document) */
typedef struct reef_sponge reef_sponge; /* device-resident SAFE sponge session */
typedef struct reef_sumcheck reef_sumcheck; /* device-resident Spartan sum-check session */

/* ------------------------------------------------------------------ lifecycle */
int reef_abi_version(void);
/* number of CUDA kernels this library has launched in this process (all contexts) */
uint64_t reef_launch_count(void);
const char* reef_last_error(void);
int reef_init(int device, reef_ctx** out);
/* Same, with the context's stream at the highest (latency_critical = 1), lowest (0) or a middle (2) CUDA stream priority:
 * the contexts that run the Fiat-Shamir chains (sum-checks) are latency-critical, the ones that run the fold commitments
 * are not; 2 puts the longer of two concurrent commitment chains (the primary curve) ahead of the other. */
int reef_init_prio(int device, int latency_critical, reef_ctx** out);
/* SMs the context's kernels may run on.  With REEF_RESERVE_SMS=k in the environment (opt-in; 12 is the natural value on
 * B200) a background context (latency_critical = 0) lives in a CUDA green context that owns all but k SMs, so that the
 * single-CTA Fiat-Shamir kernels of the latency-critical contexts always find an empty SM; otherwise, or when the driver
 * cannot partition the device, this is the whole chip. */
uint32_t reef_ctx_sm_count(const reef_ctx* ctx);
void reef_shutdown(reef_ctx* ctx);
int reef_sync(reef_ctx* ctx);
/* The CUDA stream (cudaStream_t) every launch of this context goes to; for event timing. */
void* reef_stream(reef_ctx* ctx);

/* Per-kernel-class device timing with CUDA events on the context's stream (for bench.py's
 * roofline figures).  Classes: 0 sweep (round 1), 1 sweep fold+accumulate, 2 transcript round,
 * 3 tail, 4 nlookup setup, 5 MSM sort, 6 MSM bucket accumulation, 7 MSM bucket reduction,
 * 8 Poseidon batch/Merkle.  reef_profile_read sums launches-groups, work units and
 * milliseconds per class since the last read and clears the records. */
#define REEF_PROF_NCLASS 9
int reef_profile_enable(reef_ctx* ctx, int on);
int reef_profile_read(reef_ctx* ctx, uint32_t n_classes, uint64_t* counts, uint64_t* units, double* ms);

/* ------------------------------------------------------------------ host-side helpers
 * Pure host logic of the path, kept behind the same ABI so the Rust shim and the tests use
 * one definition. */

/* costs.rs:10-15 `logmn`: (mn as f32).log2().ceil(), logmn(1) == 1. */
uint32_t reef_logmn(uint64_t mn);

/* framework.rs:978-1011 `doc_transform`.  `ab`/`doc` are Unicode scalar values.  Writes
 * 2^logmn(doc_len+2) codes.  REEF_EASSERT if a character is not in the alphabet. */
int reef_doc_transform(const uint32_t* ab, uint32_t ab_len, const uint32_t* doc, uint64_t doc_len,
                       uint64_t* out_udoc, uint64_t out_cap, uint64_t* out_len);

/* r1cs.rs:2208-2243: packs the lookup-index bits into `num_cqs = ceil(m*sc_l/254)` field
 * elements (loop quirks reproduced).  `out` must hold num_cqs*32 bytes. */
int reef_combined_q(const uint64_t* q, uint32_t m, uint32_t sc_l, uint8_t* out, uint32_t out_cap_elems,
                    uint32_t* num_cqs);

/* neptune sponge::api IOPattern tag.  ops[k]: bit 31 set = Absorb(n), clear = Squeeze(n). */
int reef_io_pattern_tag(const uint32_t* ops, uint32_t n_ops, uint32_t domain_separator, uint8_t out[32]);

/* ------------------------------------------------------------------ B2: Poseidon (neptune 8.1.0, U4, Standard) */

/* n independent one-shot hashes, IOPattern [Absorb(arity), Squeeze(1)], arity in {2,4}:
 * merkle_tree.rs:80-114 `new_parent`, commitment.rs:495-510 `calc_d`.
 * in: n*arity elements; out: n elements. */
int reef_poseidon_hash(reef_ctx* ctx, const uint8_t* in, uint32_t arity, uint64_t n, uint8_t* out);

/* commitment.rs:495-510 `calc_d(&[v, salt], pc)`. */
int reef_calc_d(reef_ctx* ctx, const uint8_t v[32], const uint8_t salt[32], uint8_t out[32]);

/* A whole SAFE sponge session in one launch: start(pattern, domain separator), the absorbs
 * and squeezes of `ops` in order, finish.  `in` = all absorbed elements, `out` = all squeezed. */
int reef_poseidon_sponge(reef_ctx* ctx, const uint8_t* in, uint32_t n_in, const uint32_t* ops, uint32_t n_ops,
                         uint32_t domain_separator, uint8_t* out, uint32_t n_out);

/* Incremental session (SpongeAPI::start/absorb/squeeze/finish as used at r1cs.rs:2260-2311,
 * r1cs_helper.rs:485-488).  reef_sponge_finish frees the session and returns REEF_EASSERT on
 * an IOPattern mismatch (neptune's ParameterUsageMismatch). */
int reef_sponge_start(reef_ctx* ctx, const uint32_t* ops, uint32_t n_ops, uint32_t domain_separator,
                      reef_sponge** out);
int reef_sponge_absorb(reef_sponge* sp, const uint8_t* elems, uint32_t n);
int reef_sponge_squeeze(reef_sponge* sp, uint32_t n, uint8_t* out);
int reef_sponge_finish(reef_sponge* sp);

/* merkle_tree.rs:25-78 `MerkleCommitment::new`: all levels (leaf parents first), concatenated.
 * Level l has ceil(prev/2) nodes; reef_merkle_tree_elems(n) is the total.  `level_sizes` must
 * hold 64 entries. */
uint64_t reef_merkle_tree_elems(uint64_t n_doc);
int reef_merkle_build(reef_ctx* ctx, const uint64_t* doc, uint64_t n_doc, uint8_t* out_levels,
                      uint64_t* level_sizes, uint32_t* n_levels, uint8_t out_root[32]);
/* Same, document and tree resident in device memory (tree stays on device). */
int reef_merkle_build_dev(reef_ctx* ctx, const uint64_t* doc_dev, uint64_t n_doc, void* levels_dev,
                          uint64_t* level_sizes, uint32_t* n_levels, uint8_t out_root[32]);

/* Multi-GPU split of MerkleCommitment::new: rank g builds the subtree over the contiguous
 * leaves [idx_offset, idx_offset + n_local) -- the leaf hash takes the GLOBAL index, merkle_tree.rs:85-96 --
 * n_local a power of two >= 2, idx_offset a multiple of n_local.  out_levels (nullable) receives the
 * n_local - 1 elements of the subtree's levels, leaf parents first; out_root its root.  The G roots are
 * all-gathered by the caller and reef_merkle_top hashes them up: (G - 1) elements, levels of G/2, G/4, .. 1. */
int reef_merkle_subtree(reef_ctx* ctx, const uint64_t* doc_local, uint64_t n_local, uint64_t idx_offset,
                        uint8_t* out_levels, uint8_t out_root[32]);
int reef_merkle_top(reef_ctx* ctx, const uint8_t* subtree_roots, uint32_t g, uint8_t* out_levels, uint8_t out_root[32]);

/* merkle_tree.rs:128-191 `path_wits(idx)` on a host copy of the tree.  Writes n_levels
 * entries: l_or_r[k], has_idx[k] (1 only for the leaf entry), opposite_idx[k] (u64),
 * opposite[k] (32 B). */
int reef_merkle_path_wits(const uint64_t* doc, uint64_t n_doc, const uint8_t* levels, const uint64_t* level_sizes,
                          uint32_t n_levels, uint64_t idx, uint8_t* l_or_r, uint8_t* has_idx,
                          uint64_t* opposite_idx, uint8_t* opposite);

/* ------------------------------------------------------------------ B1: nlookup sum-check (MLE sweeps) */

#define REEF_TAG_NL 0
#define REEF_TAG_NLDOC 1
#define REEF_TAG_NLHYBRID 2

/* Upload a table once (the reference clones it per step, r1cs.rs:2321-2330).  Length is
 * zero-padded to the next power of two (r1cs.rs:2322-2329). */
int reef_table_upload(reef_ctx* ctx, const uint8_t* table, uint64_t n, reef_table** out);
/* Document codes (framework.rs:978-1011) as u32: 8x less HBM traffic for the first two passes. */
int reef_table_upload_u32(reef_ctx* ctx, const uint32_t* codes, uint64_t n, reef_table** out);
/* The same, but returns as soon as the copy is queued (on the context's copy stream): the first absorb of the next
 * reef_nlookup_prove on this table (5-7 Poseidon permutations that never read the table) overlaps the upload; every
 * consumer orders itself behind it.  `codes` must stay valid and unmodified until a call that consumes the table has
 * returned (or reef_sync); use page-locked memory, with pageable memory the copy is simply not asynchronous. */
int reef_table_upload_u32_async(reef_ctx* ctx, const uint32_t* codes, uint64_t n, reef_table** out);
/* The merged table of `--hybrid` (r1cs.rs:481-487 and 2101-2112), built ON THE DEVICE from its two small
 * inputs: [ pub_table (n_pub elements), `fill` up to half_len | then, until the length is 2 * half_len:
 * the document codes followed by zeros up to the next power of two ].  half_len a power of two >= n_pub.
 * (The reference re-materialises this 2 * half_len vector of big integers on the host for every step.) */
int reef_table_hybrid_u32(reef_ctx* ctx, const uint8_t* pub_table, uint64_t n_pub, const uint8_t fill[32], uint64_t half_len,
                          const uint32_t* doc_codes, uint64_t n_doc, reef_table** out);
/* Wrap memory that is already on the device (n must be a power of two; not freed by reef_table_free). */
int reef_table_wrap_dev(reef_ctx* ctx, void* dev_ptr, uint64_t n, int is_u32, reef_table** out);
int reef_table_download(const reef_table* t, uint8_t* out, uint64_t n);
uint64_t reef_table_len(const reef_table* t); /* original (unpadded) length */
void reef_table_free(reef_table* t);

typedef struct reef_nlookup_out {
  uint8_t* prev_running_claim; /* 32      {id}_prev_running_claim */
  uint8_t* combined_q;         /* cap*32  {id}_combined_q_{k} */
  uint32_t combined_q_cap;     /* in: capacity in elements */
  uint32_t num_cqs;            /* out */
  uint8_t* claim_r;            /* 32      {id}_claim_r */
  uint8_t* rounds;             /* ell*4*32: per round i=1..ell {id}_sc_r_{i}, {id}_sc_g_{i}_xsq, _x, _const */
  uint32_t rounds_cap;         /* in: capacity in rounds */
  uint32_t ell;                /* out: number of sum-check rounds (= next_running_q length) */
  uint8_t* sc_last_claim;      /* 32      {id}_sc_last_claim */
  uint8_t* next_running_claim; /* 32      {id}_next_running_claim */
} reef_nlookup_out;

/* r1cs.rs:2177-2393 `wit_nlookup_gadget`, everything after the named-wire bookkeeping:
 * transcript [doc_hash?] ++ combined_q ++ v ++ prev_q ++ [prev_v] -> claim_r, eq table,
 * ell sum-check rounds, last claim, next running claim.
 * prev_q/prev_v NULL = first step defaults (zeros / table[0], r1cs.rs:2194-2201).
 * doc_hash is required for NLDOC / NLHYBRID and ignored for NL.
 * next_running_q is rounds[i][0], i = 0..ell-1. */
int reef_nlookup_prove(reef_ctx* ctx, int tag, const reef_table* table, const uint64_t* q, const uint8_t* v,
                       uint32_t m, const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash,
                       reef_nlookup_out* out);

/* Multi-GPU sharded sum-check.  The table is sharded by LOW index bits: rank g of
 * `world` (a power of two) uploads T_g[j] = T[j*world + g] as its local table.  Every rank calls
 * reef_nl_shard_begin with the SAME q, v, prev_q, prev_v (prev_v is mandatory when world > 1),
 * then for each of the ell - log2(world) local rounds:
 *     reef_nl_shard_round_local  -> this rank's (const, g(1), xsq), 96 B, into a DEVICE buffer
 *     [all-gather of the 96-byte triples across ranks, e.g. ncclAllGather]
 *     reef_nl_shard_round_finish <- the world x 96 B gathered triples (DEVICE, rank-major)
 * then reef_nl_shard_export -> this rank's folded (T, EQ) pair (64 B, DEVICE), one more
 * all-gather, and reef_nl_shard_finish runs the last log2(world) rounds redundantly on every
 * rank and writes the same outputs as reef_nlookup_prove.  None of the per-round calls
 * synchronises the stream (order your collective after reef_stream(ctx)). */
typedef struct reef_nl_session reef_nl_session;
int reef_nl_shard_begin(reef_ctx* ctx, int tag, const reef_table* local_table, uint32_t rank, uint32_t world,
                        const uint64_t* q, const uint8_t* v, uint32_t m, const uint8_t* prev_q, const uint8_t* prev_v,
                        const uint8_t* doc_hash, reef_nlookup_out* out, reef_nl_session** session);
int reef_nl_shard_round_local(reef_nl_session* s, void* out_triple_dev);
int reef_nl_shard_round_finish(reef_nl_session* s, const void* all_triples_dev);
int reef_nl_shard_export(reef_nl_session* s, void* out_pair_dev);
int reef_nl_shard_finish(reef_nl_session* s, const void* all_pairs_dev, reef_nlookup_out* out);
/* The same protocol with the exchange FUSED into the round kernels (needs reef_mailbox_connect on the
 * session's context, world <= 32): the kernel that sums this rank's round polynomial also stores it
 * into every peer's mailbox over NVLink, the kernel that runs the transcript first acquires the
 * peers' triples.  reef_nl_shard_round_p2p = one whole round, stream-ordered, no host wait;
 * call it ell - log2(world) times, then reef_nl_shard_finish_p2p (export + last rounds + results). */
int reef_nl_shard_round_p2p(reef_nl_session* s);
int reef_nl_shard_finish_p2p(reef_nl_session* s, reef_nlookup_out* out);
void reef_nl_shard_free(reef_nl_session* s);

/* Reference-shaped building blocks (materialised tables), for drop-in use and for the parity
 * tests that mirror the reference's own unit tests. */

/* r1cs_helper.rs:508-544 `gen_eq_table(rs, qs, last_q)`; out: 2^ell elements. */
int reef_gen_eq_table(reef_ctx* ctx, const uint8_t* rs, const uint64_t* qs, uint32_t m, const uint8_t* last_q,
                      uint32_t ell, uint8_t* out);
/* r1cs_helper.rs:441-506 `linear_mle_product(table_t, table_eq, ell, i, sponge)`: both tables are
 * folded in place on the device.  out = (r_i, xsq, x, con), 4*32 bytes. */
int reef_linear_mle_product(reef_ctx* ctx, reef_table* table_t, reef_table* table_eq, uint32_t ell, uint32_t i,
                            reef_sponge* sponge, uint8_t out[128]);
/* r1cs_helper.rs:637-641 `verifier_mle_eval(table, q)` (q[0] <-> top index bit). */
int reef_verifier_mle_eval(reef_ctx* ctx, const reef_table* table, const uint8_t* q, uint32_t ell, uint8_t out[32]);
/* r1cs_helper.rs:551-634 `prover_mle_partial_eval(prods, x, 0..n, true, None)`.
 * hole = index of the x entry that is -1 in the reference, or -1 for none.
 * out_coeff / out_const as returned by the reference ((crap, value) without a hole). */
int reef_prover_mle_partial_eval(reef_ctx* ctx, const reef_table* table, const uint8_t* x, uint32_t ell, int32_t hole,
                                 uint8_t out_coeff[32], uint8_t out_const[32]);

/* Hyrax `prove_eval`'s vector-matrix product (commitment.rs:371-393 -> nova hyrax_pc):
 * the table seen as a rows x cols row-major matrix M (rows * cols == padded table length),
 *   out[j] = sum_i L[i] * M[i][j],  L = eq(q_left) (rows elements), out: cols elements.
 * (L itself is reef_gen_eq_table with rs = [1], no lookups, last_q = reversed q_left.) */
int reef_hyrax_lz(reef_ctx* ctx, const reef_table* table, uint64_t rows, uint64_t cols, const uint8_t* L, uint8_t* out);

/* ------------------------------------------------------------------ B3: multi-scalar multiplication
 * Replaces nova-snark's `vartime_multiscalar_mul` / Pedersen `CE::commit` reached from
 * framework.rs:668-675 (prove_step: commit(W), commit(T)), framework.rs:695-698 (IPA inside
 * CompressedSNARK::prove) and commitment.rs:187, 350-393 (Hyrax).  Generators are static per
 * PublicParams / per document commitment, so they are registered once and stay resident. */

#define REEF_CURVE_PALLAS 0 /* coordinates in Fp, scalars in Fq */
#define REEF_CURVE_VESTA 1  /* coordinates in Fq, scalars in Fp */

typedef struct reef_bases reef_bases;

/* bases: n affine points (64 B each).  scalar_bits: upper bound on the scalars' bit length
 * (0 = 255).  Precomputes the window levels 2^(c*w) * P_i.  REEF_EINVAL if a point is not on
 * y^2 = x^3 + 5. */
int reef_bases_register(reef_ctx* ctx, int curve, const uint8_t* bases, uint64_t n, uint32_t scalar_bits,
                        reef_bases** out);
void reef_bases_free(reef_bases* b);
/* number of Pippenger windows of this registration (the unit of the multi-GPU split) */
uint32_t reef_bases_windows(const reef_bases* b);
uint32_t reef_bases_window_bits(const reef_bases* b);

/* out = sum_{i<n} scalars[i] * bases[i]   (n <= registered n; scalars canonical, < 2^scalar_bits) */
/* (f2) Registered levels are shared: a process-wide, reference-counted cache keyed by (device, curve, scalar width,
 * 128-bit content hash of the points) makes a second registration of the same `CommitmentGens` -- another stream of the
 * prover, the next proof, the verifier's own setup (framework.rs:297-303, 770, 910-976) -- free: no k_precompute, no
 * second copy in HBM.  reef_bases_cache_stats: hits since process start, live entries.  (Deriving the generators
 * themselves -- nova-snark's hash-to-curve from a label -- stays with the Rust side.) */
int reef_bases_cache_stats(uint64_t* hits, uint64_t* entries);
int reef_msm(reef_ctx* ctx, const reef_bases* b, const uint8_t* scalars, uint64_t n, uint8_t out[64]);
/* scalars already resident in device memory */
int reef_msm_dev(reef_ctx* ctx, const reef_bases* b, const void* scalars_dev, uint64_t n, uint8_t out[64]);
/* 32-bit scalars (document codes, small witness values) */
int reef_msm_u32(reef_ctx* ctx, const reef_bases* b, const uint32_t* scalars, uint64_t n, uint8_t out[64]);

/* Row-batched MSM = Hyrax polynomial commitment `hyrax_gen.commit(&poly)` (commitment.rs:187):
 * the 2^left x 2^right matrix view of the document polynomial, row-major;
 *   out[r] = sum_j M[r][j] * bases[j]  (+ blinds[r] * bases[cols] when blinds != NULL).
 * `entry_bits` bounds the matrix entries (document codes: 8 for ascii/dna, 21 for utf8; 255 for
 * field elements).  The generators must have been registered with scalar_bits = 255 when
 * blinds are used.  out: rows x 64 B. */
int reef_msm_rows_u32(reef_ctx* ctx, const reef_bases* b, const uint32_t* matrix, uint64_t rows, uint64_t cols,
                      uint32_t entry_bits, const uint8_t* blinds, uint8_t* out);
int reef_msm_rows(reef_ctx* ctx, const reef_bases* b, const uint8_t* matrix, uint64_t rows, uint64_t cols,
                  const uint8_t* blinds, uint8_t* out);
/* The same with the matrix (rows x cols canonical 32-byte scalars) resident in device memory.  With rows = 2 this is
 * commit(W) and commit(T) of one fold (framework.rs:668-675: both over the same commitment key) as ONE bucket sort /
 * accumulation / reduction chain instead of two latency pipelines. */
int reef_msm_rows_dev(reef_ctx* ctx, const reef_bases* b, const void* matrix_dev, uint64_t rows, uint64_t cols, uint8_t* out);

/* Multi-GPU: windows [w_begin, w_end) only; out_xyzz = partial sum as (X, Y, ZZ, ZZZ), 4 x 32 B
 * canonical.  The host all-gathers the partials (NCCL has no EC-add reduce op) and every rank
 * finishes with reef_msm_combine. */
int reef_msm_partial_dev(reef_ctx* ctx, const reef_bases* b, const void* scalars_dev, uint64_t n, uint32_t w_begin,
                         uint32_t w_end, uint8_t out_xyzz[128]);
int reef_msm_combine(reef_ctx* ctx, int curve, const uint8_t* partials_xyzz, uint32_t k, uint8_t out[64]);
/* The whole window-sharded MSM in one stream-ordered call (needs reef_mailbox_connect on `ctx`): this
 * rank's windows [W*g/G, W*(g+1)/G), one 128-byte all-gather of the partial points over NVLink peer
 * memory done by a kernel of this library, the combine on every rank.  Every rank of the world must call
 * it with the same scalars (resident on its own device) and gets the same affine result. */
int reef_msm_sharded_dev(reef_ctx* ctx, const reef_bases* b, const void* scalars_dev, uint64_t n, uint8_t out[64]);

/* ------------------------------------------------------------------ multi-GPU: small-message exchange over NVLink peer memory
 * The per-round exchange of the sharded sum-check (96 bytes per rank) done by a kernel of this
 * library instead of a collective call: each rank owns a mailbox in its HBM that every peer maps
 * with CUDA IPC and writes with P2P stores; sequence numbers are published / acquired with
 * system-scope release / acquire accesses (reef_b200/csrc/p2p.cu).
 *   reef_mailbox_create   allocate this rank's mailbox, return its 64-byte IPC handle
 *   reef_mailbox_connect  open the peers' mailboxes (handles of all ranks, rank-major, own slot ignored)
 *   reef_mailbox_connect_local  same-process variant (tests): device pointers from reef_mailbox_ptr
 *   reef_p2p_allgather    out_dev[g*nbytes ..] = rank g's mine_dev[0 .. nbytes), nbytes <= 248, multiple of 4;
 *                         stream-ordered, no host synchronisation; every rank must issue the same sequence
 *   reef_p2p_status       synchronises and reports a peer that never posted (bounded wait, ~2 s) */
int reef_mailbox_create(reef_ctx* ctx, uint32_t world, uint8_t out_handle[64]);
void* reef_mailbox_ptr(reef_ctx* ctx);
int reef_mailbox_connect(reef_ctx* ctx, uint32_t rank, uint32_t world, const uint8_t* handles);
int reef_mailbox_connect_local(reef_ctx* ctx, uint32_t rank, uint32_t world, void* const* mailboxes);
int reef_p2p_allgather(reef_ctx* ctx, const void* mine_dev, uint32_t nbytes, void* out_dev);
int reef_p2p_status(reef_ctx* ctx);

/* ------------------------------------------------------------------ B4: Spartan sweeps behind CompressedSNARK::prove
 * Replaces the per-round work of nova-snark's `SumcheckProof::prove_quad` and
 * `prove_cubic_with_additive_term` (reached from framework.rs:695-698 with
 * S = spartan::RelaxedR1CSSNARK<G, ipa_pc::EvaluationEngine<G>>, framework.rs:5-8, and from
 * commitment.rs:261-268 `cap_prove`).  nova-snark is a git dependency without a pinned revision
 * (Cargo.toml:12) and is not under /root/reference: the convention implemented is the published
 * upstream one (top variable bound first; round polynomial sent as its evaluations) and is
 * PARITY-UNPINNED against Reef's fork.  The transcript stays with the caller.
 *   field: 0 = Fq (Pallas scalar field, primary), 1 = Fp (Vesta scalar field, secondary).
 *   kind 2: tables {A, B},        round polynomial g(X) = sum_x A(X,x) B(X,x);        evals: g(0), g(2)
 *   kind 4: tables {A, B, C, D},  g(X) = sum_x A(X,x) (B(X,x) C(X,x) - D(X,x));     evals: g(0), g(2), g(3)
 * (g(1) = claim - g(0) is the caller's, as upstream.)  Tables: n canonical elements each, n = 2^k.
 * Round 1: reef_sumcheck_round(s, NULL, evals).  Round i+1: reef_sumcheck_round(s, r_i, evals) binds
 * every table with r_i (`bound_poly_var_top`: Z[j] += r (Z[j + len/2] - Z[j])) and accumulates the
 * next evaluations in the same sweep.  After the last round reef_sumcheck_final(s, r_k, claims)
 * returns the `kind` bound values A(r), B(r), ... */
int reef_sumcheck_begin(reef_ctx* ctx, int field, int kind, const uint8_t* const* tables, uint64_t n, reef_sumcheck** out);
int reef_sumcheck_round(reef_sumcheck* s, const uint8_t* r_prev /* NULL on the first round */, uint8_t* out_evals);
int reef_sumcheck_final(reef_sumcheck* s, const uint8_t r_last[32], uint8_t* out_claims);
void reef_sumcheck_free(reef_sumcheck* s);

/* Sparse R1CS matrix times vector over the field (the A z, B z, C z products of the folding step,
 * framework.rs:668-675 -> nova-snark R1CSShape::multiply_vec).  CSR: row_ptr has n_rows + 1 entries,
 * col_idx / vals have row_ptr[n_rows] entries (vals canonical).  out: n_rows elements. */
int reef_r1cs_spmv(reef_ctx* ctx, int field, const uint64_t* row_ptr, const uint32_t* col_idx, const uint8_t* vals,
                   uint64_t n_rows, uint64_t n_cols, const uint8_t* z, uint8_t* out);

/* IPA generator folding (commitment.rs:371-393 -> nova-snark ipa_pc `ck.fold(&r_inverse, &r)`):
 * out[i] = s_lo * bases[i] + s_hi * bases[i + n/2], i < n/2; scalars canonical in the curve's scalar
 * field.  curve: 0 Pallas, 1 Vesta.  out: (n/2) x 64 B affine. */
int reef_ipa_fold_bases(reef_ctx* ctx, int curve, const uint8_t* bases, uint64_t n, const uint8_t s_lo[32],
                        const uint8_t s_hi[32], uint8_t* out);

/* Vector helpers of the composed provers (reef_b200/snark.py), both Pasta fields (0 = Fq, 1 = Fp):
 *   reef_eq_table  out[i] = eq(r, bits(i)), i < 2^k, r[0] <-> TOP index bit (nova `EqPolynomial::evals`; the order
 *                  `bound_poly_var_top` consumes the variables in)
 *   reef_vec_axpy  out = a * x + y  (y may be NULL) */
int reef_eq_table(reef_ctx* ctx, int field, const uint8_t* r, uint32_t k, uint8_t* out);
/* a3: the NIFS cross term of prove_step (nova-snark `R1CSShape::commit_T`, reached from framework.rs:668-675):
 *   T = Az1 o Bz2 + Az2 o Bz1 - u1 Cz2 - u2 Cz1,  abc1 = Az1 | Bz1 | Cz1 (3n elements, from reef_r1cs_spmv), abc2 likewise.
 * commit(T) is reef_msm over it; the folds W1 + r W2, E1 + r T are reef_vec_axpy; the challenge r is reef_poseidon_ro. */
int reef_nova_cross_term(reef_ctx* ctx, int field, const uint8_t* abc1, const uint8_t* abc2, const uint8_t u1[32], const uint8_t u2[32], uint64_t n,
                         uint8_t* out);
int reef_vec_axpy(reef_ctx* ctx, int field, const uint8_t a[32], const uint8_t* x, const uint8_t* y, uint64_t n, uint8_t* out);

/* Inner-product argument, prover side (commitment.rs:371-393 `hyrax_gen.prove_eval` -> nova-snark ipa_pc; also the
 * polynomial openings inside CompressedSNARK::prove, framework.rs:695-698).  nova-snark is not under /root/reference
 * and is not pinned (Cargo.toml:12): the convention is the published upstream one and is PARITY-UNPINNED --
 *     c_L = <a_lo, b_hi>,  c_R = <a_hi, b_lo>,
 *     L = <a_lo, G_hi> + c_L gen_c,  R = <a_hi, G_lo> + c_R gen_c,
 *     a' = a_lo r + a_hi r^-1,  b' = b_lo r^-1 + b_hi r,  G' = G_lo r^-1 + G_hi r        (`ck.fold(&r_inverse, &r)`).
 * The vectors and the generators stay on the device for the log2(n) rounds; the transcript (which turns L, R into
 * r) stays with the caller:  begin, then log2(n) x { round -> (L, R); fold(r, r^-1) }, then finish -> a_hat, b_hat, G_hat.
 * gen_c = the generator the inner-product value is committed with (already scaled by the caller's challenge).
 * curve 0 = Pallas (scalars in Fq), 1 = Vesta (scalars in Fp).  Points 64-byte affine, infinity = zeros. */
typedef struct reef_ipa reef_ipa;
int reef_ipa_begin(reef_ctx* ctx, int curve, const uint8_t* gens, const uint8_t gen_c[64], const uint8_t* a, const uint8_t* b, uint64_t n,
                   reef_ipa** out);
/* The same session over REGISTERED generators (the static commitment key: reef_bases_register with scalar_bits = 255;
 * the first n are used): the generators are never folded -- after rounds r_0..r_(i-1) the folded generator G'_j is
 * sum_{k = j mod m} w[k] G_k with w[k] = prod_t (bit_t(k) ? r_t : r_t^-1), so every round is ONE two-row MSM (L, R) over
 * the original generators, whose window levels were precomputed at registration, and reef_ipa_fold only updates the
 * weights.  Same proofs, ~20x less time per proof at n = 2^15 (profiles/r02_summary.md). */
int reef_ipa_begin_bases(reef_ctx* ctx, const reef_bases* gens, const uint8_t gen_c[64], const uint8_t* a, const uint8_t* b, uint64_t n,
                         reef_ipa** out);
int reef_ipa_round(reef_ipa* s, uint8_t out_L[64], uint8_t out_R[64]);
int reef_ipa_fold(reef_ipa* s, const uint8_t r[32], const uint8_t r_inv[32]);
int reef_ipa_finish(reef_ipa* s, uint8_t out_a[32], uint8_t out_b[32], uint8_t out_g[64]);
void reef_ipa_free(reef_ipa* s);

/* ------------------------------------------------------------------ a16: nova-snark PoseidonRO (arity 24, width 25)
 * `PoseidonRO<Base, Scalar>::new(constants, n)`, n x absorb, squeeze(num_bits) -- commitment.rs:190-198 (the
 * doc_commit_hash of NLDocCommitment::new over the decompressed Hyrax row commitments, num_bits = 256) and the NIFS
 * challenge of every prove_step (inside nova-snark, framework.rs:668-675).  One neptune sponge over the field the
 * absorbed elements live in (base_field: 0 = Fq, 1 = Fp), (R_F, R_P) = (8, 59), IOPattern [Absorb(n), Squeeze(1)];
 * out = the digest's low num_bits bits as a canonical element of the OTHER Pasta field.  nova-snark is not under
 * /root/reference and not pinned (Cargo.toml:12): published upstream construction, PARITY-UNPINNED.
 *   reef_poseidon_ro         elems: n canonical 32-byte elements of base_field
 *   reef_poseidon_ro_points  absorbs (x, y, is_infinity) of each of n_points affine points of `curve` (64 B, zeros =
 *                            identity) -- `absorb_in_ro`; base field = the curve's coordinate field */
int reef_poseidon_ro(reef_ctx* ctx, int base_field, const uint8_t* elems, uint64_t n, uint32_t num_bits, uint8_t out[32]);
int reef_poseidon_ro_points(reef_ctx* ctx, int curve, const uint8_t* points, uint64_t n_points, uint32_t num_bits, uint8_t out[32]);

/* The arithmetic of `NLDocCommitment::new` (commitment.rs:133-212) with the blinds INJECTED by the caller (the reference
 * draws them from OsRng inside hyrax_gen.commit, commitment.rs:187): Hyrax row commitments of the rows x cols matrix view
 * of the padded document codes over gens[0..cols) with blinding generator gens[cols], then doc_commit_hash = PoseidonRO
 * over the row commitments (commitment.rs:190-198); the rows never leave the device in between.
 * out_rows: rows x 64 B affine; out_hash: canonical Fq. */
int reef_doc_commit_u32(reef_ctx* ctx, const reef_bases* gens, const uint32_t* doc, uint64_t rows, uint64_t cols, uint32_t entry_bits,
                        const uint8_t* blinds, uint8_t* out_rows, uint8_t out_hash[32]);

/* ------------------------------------------------------------------ (f3) the .cmt wire format
 * bincode 1.3 (default options: fixed-width little-endian integers, u64 lengths, u8 Option tags) of `ReefCommitment`
 * (commitment.rs:44-52), as main.rs:37-51 writes it after --commit and reads it back for --prove / --verify.  Host-only.
 *   ReefCommitment   { nldoc: Option<NLDocCommitment>, merkle: Option<MerkleCommitment<Fq>>, orig_doc_len: usize, udoc_len: usize }
 *   MerkleCommitment { commitment: F, tree: Vec<Vec<F>>, doc: Vec<F> }                            (merkle_tree.rs:10-15)
 * F = 32-byte little-endian canonical repr, no length prefix.  `levels` / `level_sizes` are exactly what reef_merkle_build
 * returns (all levels concatenated, leaves' parents first); `doc` = the padded document codes. */
uint64_t reef_cmt_merkle_size(const uint64_t* level_sizes, uint32_t n_levels, uint64_t doc_len);
int reef_cmt_merkle_write(const uint8_t commitment[32], const uint8_t* levels, const uint64_t* level_sizes, uint32_t n_levels,
                          const uint64_t* doc, uint64_t doc_len, uint64_t orig_doc_len, uint8_t* out, uint64_t out_cap, uint64_t* out_len);
/* kind: 0 = nldoc, 1 = merkle */
int reef_cmt_probe(const uint8_t* data, uint64_t len, int* kind);
/* Sizes are always returned (n_levels, n_nodes, doc_len); levels / level_sizes / doc may be NULL on a first sizing call.
 * Malformed input -> REEF_EASSERT (the reference's `expect("Could not deserialize")`, main.rs:49-50). */
int reef_cmt_merkle_read(const uint8_t* data, uint64_t len, uint8_t commitment[32], uint8_t* levels, uint64_t levels_cap, uint64_t* level_sizes,
                         uint32_t level_cap, uint32_t* n_levels, uint64_t* n_nodes, uint64_t* doc, uint64_t doc_cap, uint64_t* doc_len,
                         uint64_t* orig_doc_len, uint64_t* udoc_len);
/* Compressed point as serde writes a commitment: x little-endian, parity of y in bit 255, identity = zeros. */
int reef_point_compress(const uint8_t affine[64], uint8_t out[32]);
/* NLDocCommitment (commitment.rs:54-68).  single_gens, hyrax_gen, cap_pk, cap_vk are nova-snark types whose serde layout
 * lives in the un-pinned fork (Cargo.toml:12): OPAQUE byte strings serialised by the Rust side; everything Reef's own
 * code computes (doc_poly, doc_commit, doc_decommit, doc_commit_hash, hash_salt, q_len) is encoded here. */
typedef struct reef_cmt_nldoc {
  const uint8_t* single_gens; uint64_t single_gens_len;
  const uint8_t* hyrax_gen;   uint64_t hyrax_gen_len;
  uint32_t num_vars;                       /* doc_poly.num_vars; Z is zero-padded to 2^num_vars */
  const uint32_t* doc_codes;  uint64_t doc_len;
  const uint8_t* row_commitments;          /* rows x 64 B affine (reef_msm_rows / reef_doc_commit_u32 output) */
  const uint8_t* blinds;                   /* rows x 32 B */
  uint64_t rows;
  const uint8_t* doc_commit_hash;          /* 32 B */
  const uint8_t* hash_salt;                /* 32 B */
  const uint8_t* cap_pk;      uint64_t cap_pk_len;
  const uint8_t* cap_vk;      uint64_t cap_vk_len;
  uint64_t q_len, orig_doc_len, udoc_len;
} reef_cmt_nldoc;
uint64_t reef_cmt_nldoc_size(const reef_cmt_nldoc* f);
int reef_cmt_nldoc_write(const reef_cmt_nldoc* f, uint8_t* out, uint64_t out_cap, uint64_t* out_len);

/* ------------------------------------------------------------------ (f1) index-addressed witness buffer
 * Replaces the per-step string-keyed wire map of `NFAStepCircuit::synthesize` (nova.rs:868-1399: `FxHashMap<String, Value>`
 * wires matched with format!()-built names, int_to_ff per variable, nova.rs:31-40, 937-946; handed over at
 * framework.rs:561-572): the host resolves every wire NAME to an INDEX once per circuit shape; per step it writes plain
 * values by index, the sum-check writes its own outputs (claim_r, the round polynomials and challenges, last claim,
 * next running claim) into their slots ON THE DEVICE, and commit(W) (reef_msm_dev over reef_witness_dev) reads the
 * buffer where it lies -- no per-wire host work, no host round trip of the sum-check outputs.
 * Elements are canonical Fq values (the primary circuit's field), zero-initialised. */
typedef struct reef_witness reef_witness;
typedef struct reef_nlookup_slots {
  uint64_t claim_r;      /* UINT64_MAX = do not write */
  uint64_t rounds;       /* base of ell x 4 consecutive slots: (sc_r, xsq, x, const) per round, as reef_nlookup_out.rounds */
  uint64_t last_claim;
  uint64_t next_claim;
} reef_nlookup_slots;
int reef_witness_create(reef_ctx* ctx, uint64_t n, reef_witness** out);
int reef_witness_set(reef_witness* w, const uint64_t* idx, const uint8_t* vals, uint64_t k);      /* k canonical 32-byte values */
int reef_witness_set_u64(reef_witness* w, const uint64_t* idx, const uint64_t* vals, uint64_t k);  /* bits / small integers */
int reef_witness_read(reef_witness* w, uint64_t first, uint64_t k, uint8_t* out);
void* reef_witness_dev(reef_witness* w);          /* device address: n x 32 B, usable as reef_msm_dev scalars */
uint64_t reef_witness_len(reef_witness* w);
void reef_witness_free(reef_witness* w);
/* reef_nlookup_prove that ALSO scatters its outputs into the witness buffer (same stream, right behind the last kernel) */
int reef_nlookup_prove_w(reef_ctx* ctx, int tag, const reef_table* table, const uint64_t* q, const uint8_t* v, uint32_t m,
                         const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash, reef_nlookup_out* out,
                         reef_witness* w, const reef_nlookup_slots* slots);

#ifdef __cplusplus
}
#endif
#endif /* REEF_B200_H */
