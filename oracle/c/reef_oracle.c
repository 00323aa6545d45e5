/* reef_oracle.c -- CPU restatement of Reef's prover hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libreef_b200.so) never does.
 *
 * It follows the REFERENCE'S OWN ALGORITHM SHAPE (not the GPU schedule), so that it is both an
 * independent checker and an honest CPU baseline ("kind": "port"):
 *   gen_eq_table              O(N*ell)            /root/reference/src/backend/r1cs_helper.rs:508-544
 *   linear_mle_product        2 passes per round  /root/reference/src/backend/r1cs_helper.rs:441-506
 *   prover_mle_partial_eval   O(N*ell)            /root/reference/src/backend/r1cs_helper.rs:551-634
 *   wit_nlookup_gadget                            /root/reference/src/backend/r1cs.rs:2177-2393
 *   MerkleCommitment::new                         /root/reference/src/backend/merkle_tree.rs:25-114
 *   calc_d                                        /root/reference/src/backend/commitment.rs:495-510
 *   Poseidon / SAFE sponge   neptune 8.1.0 (un-vendored; textbook rounds, Grain constants)
 *   MSM                      Pippenger over Jacobian points, threads over windows (the reference
 *                            reaches a rayon-parallel MSM inside nova-snark)
 * The reference uses GMP big integers (`rug`) with `rem_floor` after every operation and is
 * single-threaded in the sweeps; this port uses 4x64-bit Montgomery arithmetic, which is
 * FASTER than GMP on the same core, so the reported CPU baseline is conservative.
 *
 * Pinned by tests/test_oracle_c.py against the Python oracle, which in turn is pinned by the
 * reference's KATs (tests/test_oracle_kats.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct { u64 v[4]; } fe;

typedef struct {
  u64 p[4];
  u64 r2[4];   /* 2^512 mod p */
  u64 one[4];  /* 2^256 mod p */
  u64 inv;     /* -p^-1 mod 2^64 */
} field_t;

static field_t FQ, FP;
static int g_init = 0;

/* ------------------------------------------------------------------ field arithmetic */
static inline int ge(const u64* a, const u64* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline u64 sub4(u64* r, const u64* a, const u64* b) {
  u64 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a[i] - b[i] - br;
    r[i] = (u64)d;
    br = (u64)(d >> 64) & 1;
  }
  return br;
}
static inline u64 add4(u64* r, const u64* a, const u64* b) {
  u64 c = 0;
  for (int i = 0; i < 4; i++) {
    u128 s = (u128)a[i] + b[i] + c;
    r[i] = (u64)s;
    c = (u64)(s >> 64);
  }
  return c;
}
static inline fe f_add(const field_t* F, fe a, fe b) {
  fe r;
  add4(r.v, a.v, b.v);
  if (ge(r.v, F->p)) sub4(r.v, r.v, F->p);
  return r;
}
static inline fe f_sub(const field_t* F, fe a, fe b) {
  fe r;
  if (sub4(r.v, a.v, b.v)) add4(r.v, r.v, F->p);
  return r;
}
/* CIOS Montgomery multiplication, unrolled; both Pasta primes have p[2] == 0. */
static inline __attribute__((always_inline)) fe f_mul(const field_t* F, fe a, fe b) {
  const u64 p0 = F->p[0], p1 = F->p[1], p3 = F->p[3], inv = F->inv;
  u64 t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#pragma GCC unroll 4
  for (int i = 0; i < 4; i++) {
    u128 s;
    u64 c, t5;
    s = (u128)a.v[0] * b.v[i] + t0; t0 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a.v[1] * b.v[i] + t1 + c; t1 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a.v[2] * b.v[i] + t2 + c; t2 = (u64)s; c = (u64)(s >> 64);
    s = (u128)a.v[3] * b.v[i] + t3 + c; t3 = (u64)s; c = (u64)(s >> 64);
    s = (u128)t4 + c; t4 = (u64)s; t5 = (u64)(s >> 64);
    const u64 m = t0 * inv;
    s = (u128)m * p0 + t0; c = (u64)(s >> 64);
    s = (u128)m * p1 + t1 + c; t0 = (u64)s; c = (u64)(s >> 64);
    s = (u128)t2 + c; t1 = (u64)s; c = (u64)(s >> 64);
    s = (u128)m * p3 + t3 + c; t2 = (u64)s; c = (u64)(s >> 64);
    s = (u128)t4 + c; t3 = (u64)s; t4 = t5 + (u64)(s >> 64);
  }
  fe r = {{t0, t1, t2, t3}};
  if (t4 || ge(r.v, F->p)) sub4(r.v, r.v, F->p);
  return r;
}
static inline fe f_zero(void) { fe z = {{0, 0, 0, 0}}; return z; }
static inline fe f_one(const field_t* F) { fe o; memcpy(o.v, F->one, 32); return o; }
static inline int f_is_zero(fe a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
static inline int f_eq(fe a, fe b) { return ((a.v[0] ^ b.v[0]) | (a.v[1] ^ b.v[1]) | (a.v[2] ^ b.v[2]) | (a.v[3] ^ b.v[3])) == 0; }
static inline fe f_to_mont(const field_t* F, fe a) { fe r2; memcpy(r2.v, F->r2, 32); return f_mul(F, a, r2); }
static inline fe f_from_mont(const field_t* F, fe a) { fe o = {{1, 0, 0, 0}}; return f_mul(F, a, o); }
static inline fe f_from_u64(const field_t* F, u64 x) { fe a = {{x, 0, 0, 0}}; return f_to_mont(F, a); }
static fe f_pow(const field_t* F, fe a, const u64* e) {
  fe acc = f_one(F);
  for (int i = 3; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      acc = f_mul(F, acc, acc);
      if ((e[i] >> b) & 1) acc = f_mul(F, acc, a);
    }
  return acc;
}
static fe f_inv(const field_t* F, fe a) {
  u64 e[4], two[4] = {2, 0, 0, 0};
  sub4(e, F->p, two);
  return f_pow(F, a, e);
}
static inline fe f_neg(const field_t* F, fe a) { return f_sub(F, f_zero(), a); }
static inline fe load_le(const uint8_t* b) { fe r; memcpy(r.v, b, 32); return r; }
static inline void store_le(uint8_t* b, fe a) { memcpy(b, a.v, 32); }

static void field_setup(field_t* F, const u64 p[4]) {
  memcpy(F->p, p, 32);
  u64 inv = 1;
  for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv;  /* p^-1 mod 2^64 */
  F->inv = (u64)0 - inv;
  /* one = 2^256 mod p by 256 doublings of 1; r2 by 256 more */
  u64 x[4] = {1, 0, 0, 0};
  for (int k = 0; k < 512; k++) {
    u64 c = add4(x, x, x);
    if (c || ge(x, p)) sub4(x, x, p);
    if (k == 255) memcpy(F->one, x, 32);
  }
  memcpy(F->r2, x, 32);
}

/* ------------------------------------------------------------------ Poseidon (textbook) */
#define PT 5
#define PRF 8
#define PRP 56
static fe P_RC[(PRF + PRP) * PT];
static fe P_MDS[PT][PT];

static int grain_state[80];
static int grain_clock(void) {
  int* s = grain_state;
  int b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0];
  memmove(s, s + 1, 79 * sizeof(int));
  s[79] = b;
  return b;
}
static int grain_bit(void) {
  int b = grain_clock();
  while (!b) {
    grain_clock();
    b = grain_clock();
  }
  return grain_clock();
}
static void poseidon_setup(void) {
  const int widths[7] = {2, 4, 12, 12, 10, 10, 30};
  const unsigned vals[7] = {1, 1, 255, PT, PRF, PRP, (1u << 30) - 1};
  int k = 0;
  for (int f = 0; f < 7; f++)
    for (int i = 0; i < widths[f]; i++) grain_state[k++] = (vals[f] >> (widths[f] - 1 - i)) & 1;
  for (int i = 0; i < 160; i++) grain_clock();
  int got = 0;
  while (got < (PRF + PRP) * PT) {
    u64 x[4] = {0, 0, 0, 0};
    for (int b = 0; b < 255; b++) {  /* big-endian bit stream */
      x[3] = (x[3] << 1) | (x[2] >> 63);
      x[2] = (x[2] << 1) | (x[1] >> 63);
      x[1] = (x[1] << 1) | (x[0] >> 63);
      x[0] = (x[0] << 1) | (u64)grain_bit();
    }
    if (!ge(x, FQ.p)) {
      fe c;
      memcpy(c.v, x, 32);
      P_RC[got++] = f_to_mont(&FQ, c);
    }
  }
  for (int i = 0; i < PT; i++)
    for (int j = 0; j < PT; j++) P_MDS[i][j] = f_inv(&FQ, f_from_u64(&FQ, (u64)(i + j + PT)));
}
static inline fe quintic(fe x) {
  fe x2 = f_mul(&FQ, x, x), x4 = f_mul(&FQ, x2, x2);
  return f_mul(&FQ, x4, x);
}
static void poseidon_permute(fe* s) {
  int k = 0;
  for (int r = 0; r < PRF + PRP; r++) {
    int full = r < PRF / 2 || r >= PRF / 2 + PRP;
    for (int i = 0; i < PT; i++) s[i] = f_add(&FQ, s[i], P_RC[k + i]);
    k += PT;
    if (full) for (int i = 0; i < PT; i++) s[i] = quintic(s[i]);
    else s[0] = quintic(s[0]);
    fe n[PT];
    for (int j = 0; j < PT; j++) {
      fe acc = f_zero();
      for (int i = 0; i < PT; i++) acc = f_add(&FQ, acc, f_mul(&FQ, s[i], P_MDS[i][j]));
      n[j] = acc;
    }
    memcpy(s, n, sizeof(n));
  }
}

/* Optimised schedule (same function; ~1000 instead of ~1900 multiplications), used so that
 * the CPU BASELINE is not handicapped relative to neptune, which hashes with its optimised
 * constants.  Tables come from the generator that also feeds the GPU library; equality with
 * the textbook permutation above is asserted by tests/test_oracle_c.py. */
#include "../../reef_b200/csrc/poseidon_consts.inc"
static fe O_RCF[8][5], O_RCP[56], O_ROW[56][5], O_COL[56][4], O_POST[4][4];
static int g_fast_poseidon = 0;
static fe limbs32(const uint32_t* l) {
  fe x;
  for (int i = 0; i < 4; i++) x.v[i] = (u64)l[2 * i] | ((u64)l[2 * i + 1] << 32);
  return f_to_mont(&FQ, x);
}
static void poseidon_fast_setup(void) {
  for (int r = 0; r < 8; r++) for (int i = 0; i < 5; i++) O_RCF[r][i] = limbs32(REEF_POSEIDON_RC_FULL[r * 5 + i]);
  for (int r = 0; r < 56; r++) O_RCP[r] = limbs32(REEF_POSEIDON_RC_PART[r]);
  for (int r = 0; r < 56; r++) for (int i = 0; i < 5; i++) O_ROW[r][i] = limbs32(REEF_POSEIDON_SP_ROW[r * 5 + i]);
  for (int r = 0; r < 56; r++) for (int i = 0; i < 4; i++) O_COL[r][i] = limbs32(REEF_POSEIDON_SP_COL[r * 4 + i]);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) O_POST[i][j] = limbs32(REEF_POSEIDON_POST[i * 4 + j]);
}
static void poseidon_permute_fast(fe* s) {
  const field_t* F = &FQ;
  for (int r = 0; r < 8; r++) {
    if (r == 4) {
      for (int pr = 0; pr < 56; pr++) {
        fe z0 = quintic(f_add(F, s[0], O_RCP[pr]));
        fe n0 = f_mul(F, O_ROW[pr][0], z0);
        for (int i = 1; i < 5; i++) n0 = f_add(F, n0, f_mul(F, O_ROW[pr][i], s[i]));
        for (int i = 1; i < 5; i++) s[i] = f_add(F, s[i], f_mul(F, O_COL[pr][i - 1], z0));
        s[0] = n0;
      }
      fe u[4];
      for (int j = 0; j < 4; j++) {
        fe acc = f_zero();
        for (int i = 0; i < 4; i++) acc = f_add(F, acc, f_mul(F, O_POST[j][i], s[1 + i]));
        u[j] = acc;
      }
      memcpy(s + 1, u, sizeof(u));
    }
    fe t[5], n[5];
    for (int i = 0; i < 5; i++) t[i] = quintic(f_add(F, s[i], O_RCF[r][i]));
    for (int j = 0; j < 5; j++) {
      fe acc = f_zero();
      for (int i = 0; i < 5; i++) acc = f_add(F, acc, f_mul(F, t[i], P_MDS[i][j]));
      n[j] = acc;
    }
    memcpy(s, n, sizeof(n));
  }
}
static inline void permute(fe* s) {
  if (g_fast_poseidon) poseidon_permute_fast(s);
  else poseidon_permute(s);
}
/* 0 = textbook rounds (checker default), 1 = optimised schedule (baseline timing) */
void oracle_set_fast_poseidon(int on) { g_fast_poseidon = on; }

/* SAFE sponge (Montgomery state) */
typedef struct { fe s[PT]; int apos, spos; } sponge_t;
static void tag_u128(const uint32_t* ops, uint32_t n_ops, uint32_t ds, uint8_t out[32]) {
  const u128 base = ~(u128)0 - 158;
  u128 x_i = 1, state = 0;
  int cur_abs = 1;
  u64 cur_n = 0;
  for (uint32_t k = 0; k <= n_ops; k++) {
    int is_abs = k < n_ops ? (int)(ops[k] >> 31) : !cur_abs;
    uint32_t n = k < n_ops ? (ops[k] & 0x7fffffffu) : 0;
    if (k < n_ops && is_abs == cur_abs) {
      cur_n += n;
      continue;
    }
    if (cur_n) {
      uint32_t val = cur_abs ? (uint32_t)(cur_n + (1u << 31)) : (uint32_t)cur_n;
      x_i *= base;
      state += x_i * (u128)val;
    }
    cur_abs = is_abs;
    cur_n = n;
  }
  x_i *= base;
  state += x_i * (u128)ds;
  memset(out, 0, 32);
  for (int i = 0; i < 16; i++) out[i] = (uint8_t)(state >> (8 * i));
}
static void sponge_start(sponge_t* sp, const uint32_t* ops, uint32_t n_ops) {
  uint8_t t[32];
  tag_u128(ops, n_ops, 0, t);
  sp->s[0] = f_to_mont(&FQ, load_le(t));
  for (int i = 1; i < PT; i++) sp->s[i] = f_zero();
  sp->apos = sp->spos = 0;
}
static void sponge_absorb(sponge_t* sp, const fe* e_mont, int n) {
  for (int i = 0; i < n; i++) {
    if (sp->apos == 4) {
      permute(sp->s);
      sp->apos = 0;
    }
    sp->s[1 + sp->apos] = f_add(&FQ, sp->s[1 + sp->apos], e_mont[i]);
    sp->apos++;
  }
  sp->spos = 4;
}
static fe sponge_squeeze1(sponge_t* sp) {
  if (sp->spos == 4) {
    permute(sp->s);
    sp->spos = 0;
    sp->apos = 0;
  }
  return sp->s[1 + sp->spos++];
}

static void ensure_init(void) {
  if (g_init) return;
  const u64 q[4] = {0x8c46eb2100000001ull, 0x224698fc0994a8ddull, 0x0ull, 0x4000000000000000ull};
  const u64 p[4] = {0x992d30ed00000001ull, 0x224698fc094cf91bull, 0x0ull, 0x4000000000000000ull};
  field_setup(&FQ, q);
  field_setup(&FP, p);
  poseidon_setup();
  poseidon_fast_setup();
  g_init = 1;
}

/* ------------------------------------------------------------------ exported: Poseidon */
int oracle_poseidon_hash(const uint8_t* in, uint32_t arity, uint64_t n, uint8_t* out) {
  ensure_init();
  uint32_t ops[2] = {(1u << 31) | arity, 1};
  for (uint64_t i = 0; i < n; i++) {
    sponge_t sp;
    sponge_start(&sp, ops, 2);
    fe e[4];
    for (uint32_t k = 0; k < arity; k++) e[k] = f_to_mont(&FQ, load_le(in + (i * arity + k) * 32));
    sponge_absorb(&sp, e, (int)arity);
    store_le(out + i * 32, f_from_mont(&FQ, sponge_squeeze1(&sp)));
  }
  return 0;
}

static fe hash2(fe a, fe b) {
  uint32_t ops[2] = {(1u << 31) | 2u, 1};
  sponge_t sp;
  sponge_start(&sp, ops, 2);
  fe e[2] = {a, b};
  sponge_absorb(&sp, e, 2);
  return sponge_squeeze1(&sp);
}
static fe hash4(fe a, fe b, fe c, fe d) {
  uint32_t ops[2] = {(1u << 31) | 4u, 1};
  sponge_t sp;
  sponge_start(&sp, ops, 2);
  fe e[4] = {a, b, c, d};
  sponge_absorb(&sp, e, 4);
  return sponge_squeeze1(&sp);
}

/* merkle_tree.rs:25-78.  out_levels: concatenated levels (canonical); returns #levels.
 * The reference builds the tree serially; `threads` > 1 lets the baseline use more cores. */
int oracle_merkle(const uint64_t* doc, uint64_t n, uint8_t* out_levels, uint64_t* level_sizes, int threads) {
  ensure_init();
  if (n == 0) return -1;
  uint64_t n_out = (n + 1) / 2;
  fe* cur = (fe*)malloc(sizeof(fe) * n_out);
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t k = 0; k < (int64_t)n_out; k++) {
    fe li = f_from_u64(&FQ, 2 * (u64)k), lc = f_from_u64(&FQ, doc[2 * k]);
    int has = 2 * (u64)k + 1 < n;
    fe ri = has ? f_from_u64(&FQ, 2 * (u64)k + 1) : f_zero();
    fe rc = has ? f_from_u64(&FQ, doc[2 * k + 1]) : f_zero();
    cur[k] = hash4(li, lc, ri, rc);
  }
  int nl = 0;
  uint8_t* o = out_levels;
  for (uint64_t k = 0; k < n_out; k++) store_le(o + k * 32, f_from_mont(&FQ, cur[k]));
  o += n_out * 32;
  level_sizes[nl++] = n_out;
  uint64_t n_prev = n_out;
  while (n_prev > 1) {
    n_out = (n_prev + 1) / 2;
    fe* nxt = (fe*)malloc(sizeof(fe) * n_out);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t k = 0; k < (int64_t)n_out; k++) {
      fe r = 2 * (u64)k + 1 < n_prev ? cur[2 * k + 1] : f_zero();
      nxt[k] = hash2(cur[2 * k], r);
    }
    for (uint64_t k = 0; k < n_out; k++) store_le(o + k * 32, f_from_mont(&FQ, nxt[k]));
    o += n_out * 32;
    level_sizes[nl++] = n_out;
    free(cur);
    cur = nxt;
    n_prev = n_out;
  }
  free(cur);
  return nl;
}

/* ------------------------------------------------------------------ exported: nlookup */
static uint32_t logmn(uint64_t mn) { return mn == 1 ? 1u : (uint32_t)ceilf(log2f((float)mn)); }

/* table: n canonical elements (32 B) or, if is_u32, n u32 codes.  query: the first absorb,
 * canonical, already ordered; ops/n_ops: the IOPattern.  prev_q: ell canonical elements.
 * Outputs canonical.  Reference shape: see the file header. */
int oracle_nlookup(const void* table, int is_u32, uint64_t n, const uint64_t* q, uint32_t m, const uint8_t* query,
                   uint32_t n_query, const uint32_t* ops, uint32_t n_ops, const uint8_t* prev_q, uint8_t* out_claim_r,
                   uint8_t* out_rounds, uint8_t* out_last_claim, uint8_t* out_next_v) {
  ensure_init();
  const field_t* F = &FQ;
  const uint32_t ell = logmn(n);
  const uint64_t N = (uint64_t)1 << ell;
  if (N < n) return -1;
  /* sc_table = table.to_vec() (+ zero padding)  r1cs.rs:2320-2332 */
  fe* T = (fe*)malloc(sizeof(fe) * N);
  fe* T0 = (fe*)malloc(sizeof(fe) * N);   /* pristine copy for prover_mle_partial_eval */
  fe* E = (fe*)malloc(sizeof(fe) * N);
  for (uint64_t i = 0; i < N; i++) {
    fe x = f_zero();
    if (i < n) {
      if (is_u32) x.v[0] = ((const uint32_t*)table)[i];
      else x = load_le((const uint8_t*)table + i * 32);
    }
    T[i] = f_to_mont(F, x);
    T0[i] = T[i];
  }
  sponge_t sp;
  sponge_start(&sp, ops, n_ops);
  fe* qm = (fe*)malloc(sizeof(fe) * (n_query + 1));
  for (uint32_t i = 0; i < n_query; i++) qm[i] = f_to_mont(F, load_le(query + (size_t)i * 32));
  sponge_absorb(&sp, qm, (int)n_query);
  free(qm);
  fe claim = sponge_squeeze1(&sp);
  store_le(out_claim_r, f_from_mont(F, claim));
  /* rs = claim_r^1..^(m+1)   r1cs.rs:2313-2316 */
  fe* rs = (fe*)malloc(sizeof(fe) * (m + 1));
  rs[0] = claim;
  for (uint32_t i = 1; i <= m; i++) rs[i] = f_mul(F, rs[i - 1], claim);
  /* gen_eq_table(rs, q, reversed(prev_q))   r1cs_helper.rs:508-544: O(N * ell) */
  fe* lq = (fe*)malloc(sizeof(fe) * ell);
  fe* lq1 = (fe*)malloc(sizeof(fe) * ell);
  const fe one = f_one(F);
  for (uint32_t j = 0; j < ell; j++) {
    lq[j] = f_to_mont(F, load_le(prev_q + (size_t)(ell - 1 - j) * 32));
    lq1[j] = f_sub(F, one, lq[j]);
  }
  for (uint64_t i = 0; i < N; i++) E[i] = f_zero();
  for (uint32_t k = 0; k < m; k++) E[q[k]] = f_add(F, E[q[k]], rs[k]);
  for (uint64_t i = 0; i < N; i++) {
    fe term = rs[m];
    for (int j = (int)ell - 1; j >= 0; j--) term = f_mul(F, term, ((i >> j) & 1) ? lq[j] : lq1[j]);
    E[i] = f_add(F, E[i], term);
  }
  /* ell rounds of linear_mle_product   r1cs_helper.rs:441-506 */
  fe* sc_rs = (fe*)malloc(sizeof(fe) * ell);
  fe g_xsq = f_zero(), g_x = f_zero(), g_c = f_zero(), r_i = f_zero();
  for (uint32_t i = 1; i <= ell; i++) {
    const uint64_t pw = (uint64_t)1 << (ell - i);
    fe xsq = f_zero(), x = f_zero(), con = f_zero();
    for (uint64_t b = 0; b < pw; b++) {
      fe t0 = T[b], t1 = T[b + pw], e0 = E[b], e1 = E[b + pw];
      fe ts = f_sub(F, t1, t0), es = f_sub(F, e1, e0);
      xsq = f_add(F, xsq, f_mul(F, ts, es));
      x = f_add(F, x, f_mul(F, es, t0));
      x = f_add(F, x, f_mul(F, ts, e0));
      con = f_add(F, con, f_mul(F, t0, e0));
    }
    fe ab[3] = {con, x, xsq};
    sponge_absorb(&sp, ab, 3);
    r_i = sponge_squeeze1(&sp);
    fe omr = f_sub(F, one, r_i);
    for (uint64_t b = 0; b < pw; b++) {
      T[b] = f_add(F, f_mul(F, T[b], omr), f_mul(F, T[b + pw], r_i));
      E[b] = f_add(F, f_mul(F, E[b], omr), f_mul(F, E[b + pw], r_i));
    }
    sc_rs[i - 1] = r_i;
    g_xsq = xsq; g_x = x; g_c = con;
    uint8_t* o = out_rounds + (size_t)(i - 1) * 128;
    store_le(o, f_from_mont(F, r_i));
    store_le(o + 32, f_from_mont(F, xsq));
    store_le(o + 64, f_from_mont(F, x));
    store_le(o + 96, f_from_mont(F, con));
  }
  /* last_claim   r1cs.rs:2373-2376 */
  fe lc = f_add(F, f_add(F, f_mul(F, f_mul(F, g_xsq, r_i), r_i), f_mul(F, g_x, r_i)), g_c);
  store_le(out_last_claim, f_from_mont(F, lc));
  /* prover_mle_partial_eval(table, sc_rs, 0..len, true, None)   r1cs_helper.rs:551-634: O(N*ell) */
  fe* omr = (fe*)malloc(sizeof(fe) * ell);
  for (uint32_t j = 0; j < ell; j++) omr[j] = f_sub(F, one, sc_rs[j]);
  fe minus = f_zero();
  for (uint64_t i = 0; i < n; i++) {
    fe prod = T0[i];
    for (int j = (int)ell - 1; j >= 0; j--) {
      int ej = (int)((i >> j) & 1);
      prod = f_mul(F, prod, ej ? sc_rs[ell - j - 1] : omr[ell - j - 1]);
    }
    minus = f_add(F, minus, prod);
  }
  store_le(out_next_v, f_from_mont(F, minus));
  free(T); free(T0); free(E); free(rs); free(lq); free(lq1); free(sc_rs); free(omr);
  return (int)ell;
}

/* ------------------------------------------------------------------ curves + MSM */
typedef struct { fe x, y, z; } jac;   /* Jacobian; infinity <=> z == 0 */

static jac j_inf(void) { jac r; r.x = f_zero(); r.y = f_zero(); r.z = f_zero(); return r; }
static jac j_dbl(const field_t* F, jac p) {
  if (f_is_zero(p.z) || f_is_zero(p.y)) return j_inf();
  fe a = f_mul(F, p.x, p.x), b = f_mul(F, p.y, p.y), c = f_mul(F, b, b);
  fe t = f_add(F, p.x, b);
  fe d = f_sub(F, f_sub(F, f_mul(F, t, t), a), c);
  d = f_add(F, d, d);
  fe e = f_add(F, f_add(F, a, a), a), f = f_mul(F, e, e);
  jac r;
  r.x = f_sub(F, f, f_add(F, d, d));
  fe c8 = f_add(F, c, c); c8 = f_add(F, c8, c8); c8 = f_add(F, c8, c8);
  r.y = f_sub(F, f_mul(F, e, f_sub(F, d, r.x)), c8);
  fe yz = f_mul(F, p.y, p.z);
  r.z = f_add(F, yz, yz);
  return r;
}
static jac j_add(const field_t* F, jac p, jac q) {
  if (f_is_zero(p.z)) return q;
  if (f_is_zero(q.z)) return p;
  fe z1z1 = f_mul(F, p.z, p.z), z2z2 = f_mul(F, q.z, q.z);
  fe u1 = f_mul(F, p.x, z2z2), u2 = f_mul(F, q.x, z1z1);
  fe s1 = f_mul(F, f_mul(F, p.y, q.z), z2z2), s2 = f_mul(F, f_mul(F, q.y, p.z), z1z1);
  if (f_eq(u1, u2)) return f_eq(s1, s2) ? j_dbl(F, p) : j_inf();
  fe h = f_sub(F, u2, u1), r = f_sub(F, s2, s1);
  fe hh = f_mul(F, h, h), hhh = f_mul(F, hh, h), v = f_mul(F, u1, hh);
  jac o;
  o.x = f_sub(F, f_sub(F, f_mul(F, r, r), hhh), f_add(F, v, v));
  o.y = f_sub(F, f_mul(F, r, f_sub(F, v, o.x)), f_mul(F, s1, hhh));
  o.z = f_mul(F, f_mul(F, p.z, q.z), h);
  return o;
}
/* p + (affine q, Montgomery) */
static jac j_add_affine(const field_t* F, jac p, fe qx, fe qy) {
  jac q; q.x = qx; q.y = qy; q.z = f_one(F);
  if (f_is_zero(qx) && f_is_zero(qy)) return p;
  return j_add(F, p, q);
}

/* bases: n x 64 B affine canonical; scalars: n x 32 B canonical; out: 64 B affine canonical.
 * Pippenger, unsigned windows of c bits, one thread per window. */
int oracle_msm(int curve, const uint8_t* bases, const uint8_t* scalars, uint64_t n, uint8_t* out, int threads) {
  ensure_init();
  const field_t* F = curve == 0 ? &FP : &FQ;
  int c = 1;
  while (((uint64_t)1 << (c + 5)) < n && c < 16) c++;
  if (c < 4) c = 4;
  const int W = (256 + c - 1) / c;
  fe* bx = (fe*)malloc(sizeof(fe) * (n ? n : 1));
  fe* by = (fe*)malloc(sizeof(fe) * (n ? n : 1));
#pragma omp parallel for num_threads(threads)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    bx[i] = f_to_mont(F, load_le(bases + i * 64));
    by[i] = f_to_mont(F, load_le(bases + i * 64 + 32));
  }
  jac* wsum = (jac*)malloc(sizeof(jac) * W);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
  for (int w = 0; w < W; w++) {
    const uint64_t nb = ((uint64_t)1 << c) - 1;
    jac* bucket = (jac*)malloc(sizeof(jac) * nb);
    for (uint64_t b = 0; b < nb; b++) bucket[b] = j_inf();
    for (uint64_t i = 0; i < n; i++) {
      const u64* s = (const u64*)(scalars + i * 32);
      const int bit = w * c, limb = bit >> 6, sh = bit & 63;
      u64 d = limb < 4 ? s[limb] >> sh : 0;
      if (sh + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
      d &= ((u64)1 << c) - 1;
      if (d) bucket[d - 1] = j_add_affine(F, bucket[d - 1], bx[i], by[i]);
    }
    jac run = j_inf(), acc = j_inf();
    for (int64_t b = (int64_t)nb - 1; b >= 0; b--) {
      run = j_add(F, run, bucket[b]);
      acc = j_add(F, acc, run);
    }
    wsum[w] = acc;
    free(bucket);
  }
  jac tot = j_inf();
  for (int w = W - 1; w >= 0; w--) {
    for (int k = 0; k < c; k++) tot = j_dbl(F, tot);
    tot = j_add(F, tot, wsum[w]);
  }
  memset(out, 0, 64);
  if (!f_is_zero(tot.z)) {
    fe zi = f_inv(F, tot.z), zi2 = f_mul(F, zi, zi), zi3 = f_mul(F, zi2, zi);
    store_le(out, f_from_mont(F, f_mul(F, tot.x, zi2)));
    store_le(out + 32, f_from_mont(F, f_mul(F, tot.y, zi3)));
  }
  free(bx); free(by); free(wsum);
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* field multiplication throughput probe (canonical in/out), used by the tests */
int oracle_field_mul(int field, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  ensure_init();
  const field_t* F = field == 0 ? &FQ : &FP;
  store_le(out, f_from_mont(F, f_mul(F, f_to_mont(F, load_le(a)), f_to_mont(F, load_le(b)))));
  return 0;
}

/* ------------------------------------------------------------------ PoseidonRO (width 25)
 * nova-snark's `PoseidonRO<Base, Scalar>` as the reference drives it at commitment.rs:190-198
 * (doc_commit_hash over the Hyrax row commitments): neptune sponge with arity 24 (t = 25),
 * Strength::Standard => (R_F, R_P) = (8, 59), IOPattern [Absorb(n), Squeeze(1)], Simplex mode; the
 * digest is state[1] after the last permutation.  Textbook rounds; constants by the same Grain
 * LFSR / Cauchy construction as above, over the field the absorbed elements live in.
 * nova-snark is not under /root/reference (Cargo.toml:12): parity unpinned, see oracle/poseidon.py. */
#define WT 25
#define WRF 8
#define WRP 59
typedef struct { fe rc[(WRF + WRP) * WT]; fe mds[WT][WT]; int ready; } wide_t;
static wide_t WIDE[2];
static void wide_setup(int field) {
  wide_t* W = &WIDE[field];
  if (W->ready) return;
  const field_t* F = field == 0 ? &FQ : &FP;
  const int widths[7] = {2, 4, 12, 12, 10, 10, 30};
  const unsigned vals[7] = {1, 1, 255, WT, WRF, WRP, (1u << 30) - 1};
  int k = 0;
  for (int f = 0; f < 7; f++)
    for (int i = 0; i < widths[f]; i++) grain_state[k++] = (vals[f] >> (widths[f] - 1 - i)) & 1;
  for (int i = 0; i < 160; i++) grain_clock();
  int got = 0;
  while (got < (WRF + WRP) * WT) {
    u64 x[4] = {0, 0, 0, 0};
    for (int b = 0; b < 255; b++) {
      x[3] = (x[3] << 1) | (x[2] >> 63);
      x[2] = (x[2] << 1) | (x[1] >> 63);
      x[1] = (x[1] << 1) | (x[0] >> 63);
      x[0] = (x[0] << 1) | (u64)grain_bit();
    }
    if (!ge(x, F->p)) {
      fe c;
      memcpy(c.v, x, 32);
      W->rc[got++] = f_to_mont(F, c);
    }
  }
  for (int i = 0; i < WT; i++)
    for (int j = 0; j < WT; j++) W->mds[i][j] = f_inv(F, f_from_u64(F, (u64)(i + j + WT)));
  W->ready = 1;
}
static void wide_permute(int field, fe* s) {
  const wide_t* W = &WIDE[field];
  const field_t* F = field == 0 ? &FQ : &FP;
  int k = 0;
  for (int r = 0; r < WRF + WRP; r++) {
    int full = r < WRF / 2 || r >= WRF / 2 + WRP;
    for (int i = 0; i < WT; i++) s[i] = f_add(F, s[i], W->rc[k + i]);
    k += WT;
    for (int i = 0; i < (full ? WT : 1); i++) {
      fe x2 = f_mul(F, s[i], s[i]), x4 = f_mul(F, x2, x2);
      s[i] = f_mul(F, x4, s[i]);
    }
    fe n[WT];
    for (int j = 0; j < WT; j++) {
      fe acc = f_zero();
      for (int i = 0; i < WT; i++) acc = f_add(F, acc, f_mul(F, s[i], W->mds[i][j]));
      n[j] = acc;
    }
    memcpy(s, n, sizeof(n));
  }
}
/* field: 0 = Fq, 1 = Fp (the field of the absorbed elements); elems: n canonical 32-byte values;
 * out: state[1] after the squeeze, canonical (the caller truncates to num_bits / maps to the other field) */
int oracle_poseidon_ro(int field, const uint8_t* elems, uint64_t n, uint8_t* out) {
  ensure_init();
  if (n == 0 || n >> 31) return -1;
  wide_setup(field);
  const field_t* F = field == 0 ? &FQ : &FP;
  uint32_t ops[2] = {(1u << 31) | (uint32_t)n, 1u};
  uint8_t t[32];
  tag_u128(ops, 2, 0, t);
  fe s[WT];
  s[0] = f_to_mont(F, load_le(t));
  for (int i = 1; i < WT; i++) s[i] = f_zero();
  int apos = 0;
  for (uint64_t e = 0; e < n; e++) {
    if (apos == WT - 1) {
      wide_permute(field, s);
      apos = 0;
    }
    s[1 + apos] = f_add(F, s[1 + apos], f_to_mont(F, load_le(elems + 32 * e)));
    apos++;
  }
  wide_permute(field, s);
  store_le(out, f_from_mont(F, s[1]));
  return 0;
}
