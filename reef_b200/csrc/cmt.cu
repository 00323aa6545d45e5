// (f3) The `.cmt` wire format: bincode 1.3 (default options: fixed-width little-endian integers, u64 lengths,
// u8 Option tags) of `ReefCommitment`, written by --commit and read back by --prove / --verify.
//
// Reference: /root/reference/src/backend/commitment.rs:44-68 (ReefCommitment, NLDocCommitment),
//            /root/reference/src/backend/merkle_tree.rs:10-15 (MerkleCommitment), /root/reference/src/main.rs:37-51.
//
//   ReefCommitment   { nldoc: Option<NLDocCommitment>, merkle: Option<MerkleCommitment<Fq>>, orig_doc_len: usize, udoc_len: usize }
//   MerkleCommitment { commitment: F, tree: Vec<Vec<F>>, doc: Vec<F> }
//   NLDocCommitment  { single_gens, hyrax_gen, doc_poly { num_vars: usize, Z: Vec<F> }, doc_commit { comm: Vec<compressed point> },
//                      doc_decommit { blinds: Vec<F> }, doc_commit_hash: F, hash_salt: F, cap_pk, cap_vk, q_len: usize }
//
// A field element is its 32-byte little-endian canonical repr (pasta_curves' serde impl serialises `to_repr()` as a
// fixed-size byte array in non-human-readable formats: no length prefix); a compressed point is 32 bytes (x with
// the parity of y in bit 255, identity = zeros).  The members single_gens, hyrax_gen, cap_pk, cap_vk are nova-snark
// types whose serde layout lives in the un-pinned fork (Cargo.toml:12): they are carried as OPAQUE byte strings
// that the Rust side serialises itself.  Host-only code: nothing here touches a GPU.
#include <cstring>
#include <string>

#include "common.cuh"

using namespace reef;

namespace {
struct Writer {
  uint8_t* p;
  uint64_t cap, n = 0;
  bool ok = true;
  void raw(const void* src, uint64_t k) {
    if (n + k > cap) ok = false;
    if (ok && k) memcpy(p + n, src, k);
    n += k;
  }
  void u8(uint8_t v) { raw(&v, 1); }
  void u64(uint64_t v) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
    raw(b, 8);
  }
  void fe_u64(uint64_t v) {   // F::from(v as u64)
    uint8_t b[32] = {0};
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
    raw(b, 32);
  }
};
struct Reader {
  const uint8_t* p;
  uint64_t len, n = 0;
  bool ok = true;
  const uint8_t* take(uint64_t k) {
    if (!ok || k > len - n) {
      ok = false;
      return nullptr;
    }
    const uint8_t* r = p + n;
    n += k;
    return r;
  }
  uint8_t u8() {
    const uint8_t* r = take(1);
    return r ? *r : 0;
  }
  uint64_t u64() {
    const uint8_t* r = take(8);
    uint64_t v = 0;
    if (r)
      for (int i = 0; i < 8; i++) v |= (uint64_t)r[i] << (8 * i);
    return v;
  }
};
}  // namespace

extern "C" {

uint64_t reef_cmt_merkle_size(const uint64_t* level_sizes, uint32_t n_levels, uint64_t doc_len) {
  uint64_t s = 1 + 1 + 32 + 8;
  for (uint32_t i = 0; i < n_levels; i++) s += 8 + 32 * level_sizes[i];
  return s + 8 + 32 * doc_len + 16;
}

int reef_cmt_merkle_write(const uint8_t commitment[32], const uint8_t* levels, const uint64_t* level_sizes, uint32_t n_levels,
                          const uint64_t* doc, uint64_t doc_len, uint64_t orig_doc_len, uint8_t* out, uint64_t out_cap, uint64_t* out_len) {
  REEF_REQUIRE(commitment && levels && level_sizes && doc && out && out_len, REEF_EINVAL, "reef_cmt_merkle_write: NULL argument");
  Writer w{out, out_cap};
  w.u8(0);                                   // nldoc: None
  w.u8(1);                                   // merkle: Some
  w.raw(commitment, 32);
  w.u64(n_levels);
  uint64_t off = 0;
  for (uint32_t i = 0; i < n_levels; i++) {
    w.u64(level_sizes[i]);
    w.raw(levels + 32 * off, 32 * level_sizes[i]);
    off += level_sizes[i];
  }
  w.u64(doc_len);
  for (uint64_t i = 0; i < doc_len; i++) w.fe_u64(doc[i]);
  w.u64(orig_doc_len);
  w.u64(doc_len);                            // udoc_len = doc.len() (commitment.rs:76)
  *out_len = w.n;
  REEF_REQUIRE(w.ok, REEF_EINVAL, "reef_cmt_merkle_write: output buffer too small (see reef_cmt_merkle_size)");
  return REEF_OK;
}

/* Sizes first (levels / doc may be NULL), then the data.  kind: 0 = nldoc, 1 = merkle. */
int reef_cmt_probe(const uint8_t* data, uint64_t len, int* kind) {
  REEF_REQUIRE(data && kind && len >= 2, REEF_EINVAL, "reef_cmt_probe: NULL / short input");
  REEF_REQUIRE(data[0] <= 1, REEF_EASSERT, "reef_cmt_probe: invalid Option tag (bincode: Could not deserialize)");
  *kind = data[0] == 1 ? 0 : 1;
  return REEF_OK;
}

int reef_cmt_merkle_read(const uint8_t* data, uint64_t len, uint8_t commitment[32], uint8_t* levels, uint64_t levels_cap, uint64_t* level_sizes,
                         uint32_t level_cap, uint32_t* n_levels, uint64_t* n_nodes, uint64_t* doc, uint64_t doc_cap, uint64_t* doc_len,
                         uint64_t* orig_doc_len, uint64_t* udoc_len) {
  REEF_REQUIRE(data && n_levels && n_nodes && doc_len && orig_doc_len && udoc_len, REEF_EINVAL, "reef_cmt_merkle_read: NULL argument");
  Reader r{data, len};
  const char* bad = "reef_cmt_merkle_read: malformed input (bincode: Could not deserialize)";
  REEF_REQUIRE(r.u8() == 0 && r.u8() == 1 && r.ok, REEF_EASSERT, "reef_cmt_merkle_read: not a Merkle commitment");
  const uint8_t* c = r.take(32);
  REEF_REQUIRE(r.ok, REEF_EASSERT, bad);
  if (commitment) memcpy(commitment, c, 32);
  const uint64_t nl = r.u64();
  REEF_REQUIRE(r.ok && nl <= 64, REEF_EASSERT, bad);
  uint64_t total = 0;
  for (uint64_t i = 0; i < nl; i++) {
    const uint64_t k = r.u64();
    REEF_REQUIRE(r.ok && k <= (len - r.n) / 32, REEF_EASSERT, bad);
    const uint8_t* lv = r.take(32 * k);
    if (level_sizes && i < level_cap) level_sizes[i] = k;
    if (levels && total + k <= levels_cap) memcpy(levels + 32 * total, lv, 32 * k);
    total += k;
  }
  const uint64_t dl = r.u64();
  REEF_REQUIRE(r.ok && dl <= (len - r.n) / 32, REEF_EASSERT, bad);
  for (uint64_t i = 0; i < dl; i++) {
    const uint8_t* e = r.take(32);
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) v |= (uint64_t)e[k] << (8 * k);
    for (int k = 8; k < 32; k++) REEF_REQUIRE(e[k] == 0, REEF_EASSERT, "reef_cmt_merkle_read: document entry does not fit 64 bits");
    if (doc && i < doc_cap) doc[i] = v;
  }
  *orig_doc_len = r.u64();
  *udoc_len = r.u64();
  REEF_REQUIRE(r.ok && r.n == len, REEF_EASSERT, bad);
  *n_levels = (uint32_t)nl;
  *n_nodes = total;
  *doc_len = dl;
  const bool fits = (!levels || total <= levels_cap) && (!level_sizes || nl <= level_cap) && (!doc || dl <= doc_cap);
  REEF_REQUIRE(fits, REEF_EINVAL, "reef_cmt_merkle_read: output buffers too small (sizes returned)");
  return REEF_OK;
}

/* pasta_curves `to_bytes`: x little-endian with the parity of y in bit 255; the identity is all zeros */
int reef_point_compress(const uint8_t affine[64], uint8_t out[32]) {
  REEF_REQUIRE(affine && out, REEF_EINVAL, "reef_point_compress: NULL argument");
  memcpy(out, affine, 32);
  out[31] |= (uint8_t)((affine[32] & 1u) << 7);
  return REEF_OK;
}

uint64_t reef_cmt_nldoc_size(const reef_cmt_nldoc* f) {
  if (!f) return 0;
  return 1 + f->single_gens_len + f->hyrax_gen_len + 8 + 8 + 32 * ((uint64_t)1 << f->num_vars) + 8 + 32 * f->rows + 8 + 32 * f->rows + 32 + 32 +
         f->cap_pk_len + f->cap_vk_len + 8 + 1 + 16;
}

int reef_cmt_nldoc_write(const reef_cmt_nldoc* f, uint8_t* out, uint64_t out_cap, uint64_t* out_len) {
  REEF_REQUIRE(f && out && out_len, REEF_EINVAL, "reef_cmt_nldoc_write: NULL argument");
  REEF_REQUIRE(f->doc_codes && f->row_commitments && f->blinds && f->doc_commit_hash && f->hash_salt, REEF_EINVAL, "reef_cmt_nldoc_write: NULL member");
  REEF_REQUIRE(f->num_vars < 40 && f->doc_len <= ((uint64_t)1 << f->num_vars), REEF_EINVAL, "reef_cmt_nldoc_write: document longer than 2^num_vars");
  Writer w{out, out_cap};
  w.u8(1);                                                         // nldoc: Some
  w.raw(f->single_gens, f->single_gens_len);                       // opaque (nova-snark CommitmentGens<G1>)
  w.raw(f->hyrax_gen, f->hyrax_gen_len);                           // opaque (nova-snark HyraxPC<G1>)
  w.u64(f->num_vars);                                              // doc_poly.num_vars
  const uint64_t n = (uint64_t)1 << f->num_vars;
  w.u64(n);                                                        // doc_poly.Z (zero-padded, commitment.rs:161-166)
  for (uint64_t i = 0; i < n; i++) w.fe_u64(i < f->doc_len ? f->doc_codes[i] : 0);
  w.u64(f->rows);                                                  // doc_commit.comm: compressed row commitments
  for (uint64_t r = 0; r < f->rows; r++) {
    uint8_t cp[32];
    reef_point_compress(f->row_commitments + 64 * r, cp);
    w.raw(cp, 32);
  }
  w.u64(f->rows);                                                  // doc_decommit.blinds
  w.raw(f->blinds, 32 * f->rows);
  w.raw(f->doc_commit_hash, 32);
  w.raw(f->hash_salt, 32);
  w.raw(f->cap_pk, f->cap_pk_len);                                 // opaque (SpartanProverKey / SpartanVerifierKey)
  w.raw(f->cap_vk, f->cap_vk_len);
  w.u64(f->q_len);
  w.u8(0);                                                         // merkle: None
  w.u64(f->orig_doc_len);
  w.u64(f->udoc_len);
  *out_len = w.n;
  REEF_REQUIRE(w.ok, REEF_EINVAL, "reef_cmt_nldoc_write: output buffer too small (see reef_cmt_nldoc_size)");
  return REEF_OK;
}

}  // extern "C"
