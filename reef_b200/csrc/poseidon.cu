// Poseidon kernels: batched one-shot hashes, Merkle levels, and a generic SAFE-sponge runner.
//
// Reference behaviour reproduced (neptune 8.1.0 through these call sites):
//   MerkleCommitment::new / new_parent   /root/reference/src/backend/merkle_tree.rs:25-114
//   calc_d                               /root/reference/src/backend/commitment.rs:495-510
//   SpongeAPI start/absorb/squeeze       /root/reference/src/backend/r1cs.rs:2260-2311
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace reef {

#include "poseidon_consts.inc"

// ---------------------------------------------------------------------------------------
// host: IOPattern tag (SAFE) -- u128 polynomial hash, wrapping arithmetic
// ---------------------------------------------------------------------------------------
void io_pattern_tag_le32(const uint32_t* ops, uint32_t n_ops, uint32_t domain_separator, uint8_t out[32]) {
  // ops[k]: bit 31 set = Absorb(n), clear = Squeeze(n); n in the low 31 bits.
  typedef unsigned __int128 u128;
  const u128 base = ~(u128)0 - 158;  // 2^128 - 159
  u128 x_i = 1, state = 0;
  bool cur_absorb = true;
  uint64_t cur_n = 0;
  auto update = [&](uint32_t a) {
    x_i = x_i * base;
    state = state + x_i * (u128)a;
  };
  auto finish_op = [&]() {
    if (cur_n == 0) return;
    uint32_t val = cur_absorb ? (uint32_t)(cur_n + (1u << 31)) : (uint32_t)cur_n;
    update(val);
  };
  for (uint32_t k = 0; k < n_ops; k++) {
    bool is_absorb = (ops[k] >> 31) != 0;
    uint32_t n = ops[k] & 0x7fffffffu;
    if (is_absorb == cur_absorb) {
      cur_n += n;
    } else {
      finish_op();
      cur_absorb = is_absorb;
      cur_n = n;
    }
  }
  finish_op();
  update(domain_separator);
  for (int i = 0; i < 32; i++) out[i] = 0;
  for (int i = 0; i < 16; i++) out[i] = (uint8_t)(state >> (8 * i));
}

static Fq fq_from_limbs(const uint32_t* l) {
  Fq x;
  for (int i = 0; i < 8; i++) x.v[i] = l[i];
  return x;
}

Fq fq_canon_from_le32(const uint8_t* b) {
  Fq x;
  for (int i = 0; i < 8; i++)
    x.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
             ((uint32_t)b[4 * i + 3] << 24);
  return x;
}

Fq fq_mont_from_le32(const uint8_t* b) {
  Fq x;
  for (int i = 0; i < 8; i++)
    x.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) |
             ((uint32_t)b[4 * i + 3] << 24);
  return to_mont<FqCfg>(x);
}

void poseidon_tables_host(PoseidonTables* t) {
  auto cv = [](const uint32_t* l) { return to_mont<FqCfg>(fq_from_limbs(l)); };
  for (int r = 0; r < 8; r++)
    for (int i = 0; i < 5; i++) t->rc_full[r][i] = cv(REEF_POSEIDON_RC_FULL[r * 5 + i]);
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) t->mds[i][j] = cv(REEF_POSEIDON_MDS[i * 5 + j]);
  for (int r = 0; r < 57; r++) t->kp[r] = cv(REEF_POSEIDON_KP[r]);
  for (int r = 0; r < 56; r++)
    for (int i = 0; i < 4; i++) t->beta[r][i] = cv(REEF_POSEIDON_BETA[r * 4 + i]);
  for (int i = 0; i < 4; i++) t->dshift[0][i] = fe_zero<FqCfg>();
  for (int r = 0; r < 56; r++)
    for (int i = 0; i < 4; i++) t->dshift[r + 1][i] = cv(REEF_POSEIDON_D[r * 4 + i]);
  t->lam_end = cv(REEF_POSEIDON_LAM_END[0]);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) t->post[i][j] = cv(REEF_POSEIDON_POST[i * 4 + j]);
  // 29-bit-limb copies: Montgomery-256 -> Montgomery-261 is a multiplication by 2^5
  auto c29 = [](const Fq& m256) {
    Fq x = m256;
    for (int k = 0; k < 5; k++) x = fe_dbl<FqCfg>(x);
    F29 f = f29_from_words(x.v);
    F29s o;
    for (int k = 0; k < 12; k++) o.l[k] = k < 9 ? f.l[k] : 0u;
    return o;
  };
  Poseidon29Tables& q = t->t29;
  for (int r = 0; r < 8; r++)
    for (int i = 0; i < 5; i++) q.rc_full[r][i] = c29(t->rc_full[r][i]);
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) q.mds[i][j] = c29(t->mds[i][j]);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) q.post[i][j] = c29(t->post[i][j]);
  for (int r = 0; r < 57; r++) q.kp[r] = c29(t->kp[r]);
  for (int r = 0; r < 56; r++)
    for (int i = 0; i < 4; i++) {
      q.beta[r][i] = c29(t->beta[r][i]);
      q.emat[r][i] = c29(mont_mul<FqCfg>(t->beta[r][i], t->dshift[r][i]));
    }
  for (int r = 0; r < 57; r++)
    for (int i = 0; i < 4; i++) q.dshift[r][i] = c29(t->dshift[r][i]);
  q.lam_end = c29(t->lam_end);
  {
    const F29 k = f29_const_2_266<FqCfg>();     // plain integer, not a Montgomery-form element
    for (int i = 0; i < 12; i++) q.k266.l[i] = i < 9 ? k.l[i] : 0u;
  }
}

// Tables of the lane-parallel transcript permutation (poseidon_lp.cuh), derived from the same
// constants: Gamma[r][t] = sum_i beta[r][i] D[t][i],  PD[j][t] = sum_i post[j][i] D[t][i].
void poseidon_lp_tables_host(PoseidonLpTables* o) {
  PoseidonTables* t = new PoseidonTables;
  poseidon_tables_host(t);
  memset(o, 0, sizeof(*o));
  auto plain = [](const Fq& m256, u32* out12) {          // Montgomery-256 -> canonical 29-bit limbs
    const Fq x = from_mont<FqCfg>(m256);
    const F29 f = f29_from_words(x.v);
    for (int k = 0; k < 12; k++) out12[k] = k < 9 ? f.l[k] : 0u;
  };
  auto padded = [&](const Fq& m256, LpPad* out) {
    u32 l[12];
    plain(m256, l);
    for (int k = 0; k < LP_PAD_WORDS; k++) out->w[k] = 0;
    for (int k = 0; k < 9; k++) out->w[LP_OFF_WORDS + k] = l[k];
  };
  auto m261 = [](const Fq& m256) {                       // Montgomery-256 -> Montgomery-261 (x 2^5)
    Fq x = m256;
    for (int k = 0; k < 5; k++) x = fe_dbl<FqCfg>(x);
    const F29 f = f29_from_words(x.v);
    F29s r;
    for (int k = 0; k < 12; k++) r.l[k] = k < 9 ? f.l[k] : 0u;
    return r;
  };
  for (int r = 0; r < 8; r++)
    for (int i = 0; i < 5; i++) plain(t->rc_full[r][i], o->rcf[r][i]);
  plain(t->kp[0], o->kp0);
  for (int r = 0; r < 57; r++) plain(t->kp[r], o->kp[r]);
  for (int j = 0; j < 4; j++) plain(t->rc_full[4][j + 1], o->rc4[j]);
  for (int j = 0; j < 5; j++)
    for (int i = 0; i < 5; i++) padded(t->mds[j][i], &o->mds[j][i]);
  padded(t->lam_end, &o->lam);
  {
    Fq r256;                                             // 2^256 mod p as a Montgomery-256 element = R^2 mod p as an integer
    for (int i = 0; i < 8; i++) r256.v[i] = FqCfg::r2(i);
    padded(r256, &o->r256);
  }
  for (int r = 0; r < 56; r++)
    for (int i = 0; i < 4; i++) o->beta[r][i] = m261(t->beta[r][i]);
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) o->post[j][i] = m261(t->post[j][i]);
  for (int r = 0; r < 56; r++)
    for (int tt = 0; tt < r; tt++) {
      Fq g = fe_zero<FqCfg>();
      for (int i = 0; i < 4; i++) g = fe_add<FqCfg>(g, mont_mul<FqCfg>(t->beta[r][i], t->dshift[tt + 1][i]));
      o->gam[r][tt] = m261(g);
      if (tt == r - 1) padded(g, &o->gam1[r]);
    }
  for (int j = 0; j < 4; j++)
    for (int tt = 0; tt < 56; tt++) {
      Fq g = fe_zero<FqCfg>();
      for (int i = 0; i < 4; i++) g = fe_add<FqCfg>(g, mont_mul<FqCfg>(t->post[j][i], t->dshift[tt + 1][i]));
      o->pd[j][tt] = m261(g);
      if (tt == 55) padded(g, &o->pd55[j]);
    }
  delete t;
}

int poseidon_upload_constants(reef_ctx* c) {
  PoseidonTables* h = new PoseidonTables;
  poseidon_tables_host(h);
  cudaError_t e = cudaMalloc(&c->d_pos, sizeof(PoseidonTables));
  if (e == cudaSuccess) e = cudaMemcpy(c->d_pos, h, sizeof(PoseidonTables), cudaMemcpyHostToDevice);
  delete h;
  if (e == cudaSuccess) {
    PoseidonLpTables* hl = new PoseidonLpTables;
    poseidon_lp_tables_host(hl);
    e = cudaMalloc(&c->d_lp, sizeof(PoseidonLpTables));
    if (e == cudaSuccess) e = cudaMemcpy(c->d_lp, hl, sizeof(PoseidonLpTables), cudaMemcpyHostToDevice);
    delete hl;
  }
  if (e != cudaSuccess) return fail(REEF_ECUDA, std::string("poseidon constants: ") + cudaGetErrorString(e));
  uint8_t tag[32];
  uint32_t p2[2] = {(1u << 31) | 2u, 1u}, p4[2] = {(1u << 31) | 4u, 1u};
  io_pattern_tag_le32(p2, 2, 0, tag);
  c->tags.a2s1 = fq_mont_from_le32(tag);
  io_pattern_tag_le32(p4, 2, 0, tag);
  c->tags.a4s1 = fq_mont_from_le32(tag);
  return REEF_OK;
}

// Host evaluation of the permutation through the SAME shared code path as the device
// (test hook only; see include/reef_b200_testing.h).
void poseidon_permute_host(Fq* s) {
  static PoseidonTables* T = nullptr;
  if (!T) {
    T = new PoseidonTables;
    poseidon_tables_host(T);
  }
  poseidon_permute(s, *T);
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------

// one thread per hash; in: n x arity canonical elements; IOPattern [Absorb(arity), Squeeze(1)]
__global__ void __launch_bounds__(128) k_hash_batch(const Fq* __restrict__ in, int arity, uint64_t n, Fq tag,
                                                    const PoseidonTables* __restrict__ K, Fq* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq s[5];
  s[0] = tag;
#pragma unroll 1
  for (int k = 0; k < 4; k++) s[1 + k] = (k < arity) ? to_mont<FqCfg>(ld256(in + i * arity + k)) : fe_zero<FqCfg>();
  poseidon_permute(s, *K);
  st256(out + i, from_mont<FqCfg>(s[1]));
}

// same contract, one WARP per hash (latency path: calc_d, small batches)
__global__ void __launch_bounds__(128) k_hash_batch_warp(const Fq* __restrict__ in, int arity, uint64_t n, Fq tag,
                                                         const PoseidonTables* __restrict__ K, Fq* __restrict__ out) {
  uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;  // warp-uniform
  Fq s = fe_zero<FqCfg>();
  if (lane >= 1 && lane <= arity) s = ld256(in + i * arity + (lane - 1));
  Fq sm = to_mont<FqCfg>(s);
  s = lane == 0 ? tag : sm;
  poseidon_permute_warp5(s, K);
  Fq o = from_mont<FqCfg>(s);
  if (lane == 1) st256(out + i, o);
}

// same contract, one two-warp CTA per hash (latency path: calc_d, a handful of hashes)
__global__ void __launch_bounds__(64) k_hash_batch_pair(const Fq* __restrict__ in, int arity, uint64_t n, Fq tag,
                                                        const PoseidonTables* __restrict__ K, Fq* __restrict__ out) {
  const uint64_t i = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const bool A = threadIdx.x < 32;
  Fq s = fe_zero<FqCfg>();
  if (A) {
    if (lane >= 1 && lane <= arity) s = ld256(in + i * arity + (lane - 1));
    const Fq sm = to_mont<FqCfg>(s);
    s = lane == 0 ? tag : sm;
  }
  poseidon_permute_pair(s, K);
  if (A) {
    const Fq o = from_mont<FqCfg>(s);
    if (lane == 1) st256(out + i, o);
  }
}

// leaf level: parent k = H4(2k, doc[2k], 2k+1, doc[2k+1]); missing right => (.., 0, 0)
__global__ void __launch_bounds__(128) k_merkle_leaves(const uint64_t* __restrict__ doc, uint64_t n_doc, Fq tag4,
                                                       const PoseidonTables* __restrict__ K, Fq* __restrict__ out,
                                                       uint64_t idx_offset) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t n_out = (n_doc + 1) / 2;
  if (k >= n_out) return;
  Fq s[5];
  s[0] = tag4;
  s[1] = fe_from_u64<FqCfg>(idx_offset + 2 * k);
  s[2] = fe_from_u64<FqCfg>(doc[2 * k]);
  bool has_r = 2 * k + 1 < n_doc;
  s[3] = has_r ? fe_from_u64<FqCfg>(idx_offset + 2 * k + 1) : fe_zero<FqCfg>();
  s[4] = has_r ? fe_from_u64<FqCfg>(doc[2 * k + 1]) : fe_zero<FqCfg>();
  poseidon_permute(s, *K);
  st256(out + k, from_mont<FqCfg>(s[1]));
}

// inner level: parent k = H2(prev[2k], prev[2k+1]); missing right => (l, 0)
__global__ void __launch_bounds__(128) k_merkle_level(const Fq* __restrict__ prev, uint64_t n_prev, Fq tag2,
                                                      const PoseidonTables* __restrict__ K, Fq* __restrict__ out) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t n_out = (n_prev + 1) / 2;
  if (k >= n_out) return;
  Fq s[5];
  s[0] = tag2;
  s[1] = to_mont<FqCfg>(ld256(prev + 2 * k));
  s[2] = (2 * k + 1 < n_prev) ? to_mont<FqCfg>(ld256(prev + 2 * k + 1)) : fe_zero<FqCfg>();
  s[3] = fe_zero<FqCfg>();
  s[4] = fe_zero<FqCfg>();
  poseidon_permute(s, *K);
  st256(out + k, from_mont<FqCfg>(s[1]));
}

// inner level, one WARP per parent (latency path for the small top levels of the tree)
__global__ void __launch_bounds__(128) k_merkle_level_warp(const Fq* __restrict__ prev, uint64_t n_prev, Fq tag2,
                                                           const PoseidonTables* __restrict__ K,
                                                           Fq* __restrict__ out) {
  uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  uint64_t n_out = (n_prev + 1) / 2;
  if (k >= n_out) return;  // warp-uniform
  Fq s = fe_zero<FqCfg>();
  if (lane == 0) s = tag2;
  if (lane == 1) s = ld256(prev + 2 * k);
  if (lane == 2 && 2 * k + 1 < n_prev) s = ld256(prev + 2 * k + 1);
  Fq sm = to_mont<FqCfg>(s);
  s = (lane == 1 || lane == 2) ? sm : s;
  poseidon_permute_warp5(s, K);
  Fq o = from_mont<FqCfg>(s);
  if (lane == 1) st256(out + k, o);
}

// SAFE sponge state machine shared by the one-shot runner and the incremental session.
struct SpongeDev {
  Fq s[5];          // Montgomery form
  uint32_t apos, spos;
};

__device__ __forceinline__ void sponge_absorb_warp(Fq& s, uint32_t& apos, uint32_t& spos, const Fq* in, uint32_t n,
                                                   const PoseidonTables* K) {
  const int lane = threadIdx.x & 31;
  for (uint32_t e = 0; e < n; e++) {
    if (apos == 4) {
      poseidon_permute_warp5(s, K);
      apos = 0;
    }
    Fq x = to_mont<FqCfg>(ld256(in + e));
    Fq sum = fe_add<FqCfg>(s, x);
    if (lane == (int)(1 + apos)) s = sum;
    apos++;
  }
  spos = 4;
}

__device__ __forceinline__ void sponge_squeeze_warp(Fq& s, uint32_t& apos, uint32_t& spos, Fq* out, uint32_t n,
                                                    const PoseidonTables* K) {
  const int lane = threadIdx.x & 31;
  for (uint32_t e = 0; e < n; e++) {
    if (spos == 4) {
      poseidon_permute_warp5(s, K);
      spos = 0;
      apos = 0;
    }
    Fq o = from_mont<FqCfg>(s);
    if (lane == (int)(1 + spos)) st256(out + e, o);
    spos++;
  }
}

// Whole session in one launch.  ops[k]: bit31 = absorb, low bits = count.
__global__ void __launch_bounds__(32) k_sponge_run(const uint32_t* __restrict__ ops, uint32_t n_ops,
                                                   const Fq* __restrict__ in, Fq tag,
                                                   const PoseidonTables* __restrict__ K, Fq* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  Fq s = fe_zero<FqCfg>();
  if (lane == 0) s = tag;
  uint32_t apos = 0, spos = 0, iin = 0, iout = 0;
  for (uint32_t k = 0; k < n_ops; k++) {
    uint32_t n = ops[k] & 0x7fffffffu;
    if (ops[k] >> 31) {
      sponge_absorb_warp(s, apos, spos, in + iin, n, K);
      iin += n;
    } else {
      sponge_squeeze_warp(s, apos, spos, out + iout, n, K);
      iout += n;
    }
  }
}

// Incremental session: op = 0 init(tag), 1 absorb(n from in), 2 squeeze(n to out)
__global__ void __launch_bounds__(32) k_sponge_step(SpongeDev* st, int op, const Fq* __restrict__ in, uint32_t n, Fq tag,
                                                    const PoseidonTables* __restrict__ K, Fq* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  if (op == 0) {
    if (lane < 5) st->s[lane] = lane == 0 ? tag : fe_zero<FqCfg>();
    if (lane == 0) { st->apos = 0; st->spos = 0; }
    return;
  }
  Fq s = lane < 5 ? st->s[lane] : fe_zero<FqCfg>();
  uint32_t apos = st->apos, spos = st->spos;
  __syncwarp();
  if (op == 1) sponge_absorb_warp(s, apos, spos, in, n, K);
  else sponge_squeeze_warp(s, apos, spos, out, n, K);
  if (lane < 5) st->s[lane] = s;
  if (lane == 0) { st->apos = apos; st->spos = spos; }
}

// ---------------------------------------------------------------------------------------
// launchers (device pointers)
// ---------------------------------------------------------------------------------------
int launch_hash_batch(reef_ctx* c, const void* d_in, int arity, uint64_t n, void* d_out) {
  if (n == 0) return REEF_OK;
  REEF_REQUIRE(arity == 2 || arity == 4, REEF_EINVAL, "poseidon hash arity must be 2 or 4");
  Fq tag = arity == 2 ? c->tags.a2s1 : c->tags.a4s1;
  ProfScope ps(c, PROF_POSEIDON, n);
  if (n <= (uint64_t)c->sm_count) {        // at most one hash per SM: two warps per hash (lowest latency)
    k_hash_batch_pair<<<(unsigned)n, 64, 0, c->stream>>>((const Fq*)d_in, arity, n, tag, c->d_pos, (Fq*)d_out);
  } else if (n < (uint64_t)c->sm_count * 256) {   // too few hashes to fill the machine: one warp each
    k_hash_batch_warp<<<(unsigned)((n * 32 + 127) / 128), 128, 0, c->stream>>>((const Fq*)d_in, arity, n, tag, c->d_pos, (Fq*)d_out);
  } else {
    k_hash_batch<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>((const Fq*)d_in, arity, n, tag, c->d_pos, (Fq*)d_out);
  }
  REEF_LAUNCHED();
  return REEF_OK;
}

// d_levels: concatenated levels, leaf-parents first; sizes ceil(n/2), ceil(ceil(n/2)/2), ... 1
int launch_merkle(reef_ctx* c, const uint64_t* d_doc, uint64_t n_doc, void* d_levels, uint64_t* level_sizes,
                  uint32_t* n_levels_out, uint64_t idx_offset) {
  REEF_REQUIRE(n_doc >= 1, REEF_EINVAL, "merkle: empty document");
  Fq* lv = (Fq*)d_levels;
  uint64_t n_out = (n_doc + 1) / 2;
  ProfScope ps(c, PROF_POSEIDON, n_doc);
  k_merkle_leaves<<<(unsigned)((n_out + 127) / 128), 128, 0, c->stream>>>(d_doc, n_doc, c->tags.a4s1, c->d_pos, lv, idx_offset);
  REEF_LAUNCHED();
  if (level_sizes) level_sizes[0] = n_out;
  uint32_t nl_in = 0;
  int rc = launch_merkle_inner(c, lv, n_out, lv + n_out, level_sizes ? level_sizes + 1 : nullptr, &nl_in);
  if (rc) return rc;
  if (n_levels_out) *n_levels_out = nl_in + 1;
  return REEF_OK;
}

int launch_merkle_inner(reef_ctx* c, const void* d_prev, uint64_t n_prev_in, void* d_levels, uint64_t* level_sizes,
                        uint32_t* n_levels_out) {
  uint32_t nl = 0;
  const Fq* prev = (const Fq*)d_prev;
  Fq* next_out = (Fq*)d_levels;
  uint64_t n_prev = n_prev_in;
  uint64_t n_out;
  while (n_prev > 1) {
    Fq* cur = next_out;
    n_out = (n_prev + 1) / 2;
    // thread-per-hash while the level still fills the machine, warp-per-hash for the top
    if (n_out >= (uint64_t)c->sm_count * 256) {
      k_merkle_level<<<(unsigned)((n_out + 127) / 128), 128, 0, c->stream>>>(prev, n_prev, c->tags.a2s1, c->d_pos, cur);
    } else {
      k_merkle_level_warp<<<(unsigned)((n_out * 32 + 127) / 128), 128, 0, c->stream>>>(prev, n_prev, c->tags.a2s1,
                                                                                         c->d_pos, cur);
    }
    REEF_LAUNCHED();
    if (level_sizes) level_sizes[nl] = n_out;
    nl++;
    prev = cur;
    next_out = cur + n_out;
    n_prev = n_out;
  }
  if (n_levels_out) *n_levels_out = nl;
  return REEF_OK;
}

int launch_sponge_run(reef_ctx* c, const uint32_t* d_ops, uint32_t n_ops, const void* d_in, const uint8_t tag_le[32],
                      void* d_out) {
  Fq tag = fq_mont_from_le32(tag_le);
  k_sponge_run<<<1, 32, 0, c->stream>>>(d_ops, n_ops, (const Fq*)d_in, tag, c->d_pos, (Fq*)d_out);
  REEF_LAUNCHED();
  return REEF_OK;
}

int sponge_state_bytes() { return (int)sizeof(SpongeDev); }

int launch_sponge_step(reef_ctx* c, void* d_state, int op, const void* d_in, uint32_t n, const uint8_t* tag_le,
                       void* d_out) {
  Fq tag = tag_le ? fq_mont_from_le32(tag_le) : fe_zero<FqCfg>();
  k_sponge_step<<<1, 32, 0, c->stream>>>((SpongeDev*)d_state, op, (const Fq*)d_in, n, tag, c->d_pos, (Fq*)d_out);
  REEF_LAUNCHED();
  return REEF_OK;
}

}  // namespace reef

// ---------------------------------------------------------------------------------------
// test hook: the lane-parallel transcript permutation alone (include/reef_b200_testing.h)
// ---------------------------------------------------------------------------------------
namespace reef {
__global__ void __launch_bounds__(LP_PERM_THREADS) k_lp_perm_test(const Fq* __restrict__ in, Fq* __restrict__ out, uint32_t n,
                                                                  const PoseidonLpTables* __restrict__ T, long long* cycles) {
  __shared__ LpPermShared sh;
  lp_perm_init(&sh, T);
  if (threadIdx.x < 45) sh.S[threadIdx.x / 9][threadIdx.x % 9] = lp_limb_of(in[threadIdx.x / 9].v, threadIdx.x % 9);
  __syncthreads();
  const long long t0 = clock64();
  for (uint32_t k = 0; k < n; k++) poseidon_permute_lp(&sh, T, k + 1);
  const long long t1 = clock64();
  if (threadIdx.x < 5) out[threadIdx.x] = lp_squeeze(&sh, threadIdx.x);
  if (threadIdx.x == 0) {
    cycles[0] = n ? (t1 - t0) / n : 0;
    // phases of the LAST permutation: first full rounds, partial rounds, end, last full rounds; waits inside the chain
    cycles[1] = sh.dbg[1] - sh.dbg[0];
    cycles[2] = sh.dbg[2] - sh.dbg[1];
    cycles[3] = sh.dbg[3] - sh.dbg[2];
    cycles[4] = sh.dbg[4] - sh.dbg[3];
    cycles[5] = sh.dbg[5];
    cycles[6] = sh.dbg[6];
    cycles[7] = sh.dbg[7];
    cycles[8] = sh.dbg[8];
  }
}
}  // namespace reef

extern "C" int reef_gputest_poseidon_permute_lp(void* ctx, const uint8_t in[160], uint32_t n_perms, uint8_t out[160],
                                                uint64_t* cycles_per_perm) {
  using namespace reef;
  reef_ctx* c = (reef_ctx*)ctx;
  REEF_REQUIRE(c && in && out, REEF_EINVAL, "reef_gputest_poseidon_permute_lp: NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* base;
  int rc = ctx_scratch(c, 1024, &base);
  if (rc) return rc;
  char* d = (char*)base;
  REEF_CUDA(cudaMemcpyAsync(d, in, 160, cudaMemcpyHostToDevice, c->stream));
  k_lp_perm_test<<<1, LP_PERM_THREADS, 0, c->stream>>>((const Fq*)d, (Fq*)(d + 160), n_perms, c->d_lp, (long long*)(d + 320));
  REEF_LAUNCHED();
  uint8_t h[160 + 80];
  REEF_CUDA(cudaMemcpyAsync(h, d + 160, 160 + 72, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(out, h, 160);
  if (cycles_per_perm) memcpy(cycles_per_perm, h + 160, 72);   // [0] per permutation; [1..4] phases, [5..6] chain waits of the last one
  return REEF_OK;
}
