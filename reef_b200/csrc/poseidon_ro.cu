// Wide Poseidon sponge (arity 24, width 25) over either Pasta field: nova-snark's `PoseidonRO`.
//
// Reference call sites:
//   doc_commit_hash: RO over (x, y, is_infinity) of every Hyrax row commitment, squeeze(256)
//                    /root/reference/src/backend/commitment.rs:190-198
//   the NIFS challenge of every prove_step (inside nova-snark, reached from framework.rs:668-675)
// nova-snark is a git dependency without a pinned revision and is not under /root/reference
// (Cargo.toml:12); the construction is the published upstream provider/poseidon.rs: ONE neptune sponge,
// `Sponge::<Base, U24>::api_constants(Strength::Standard)` => (R_F, R_P) = (8, 59), Simplex mode,
// IOPattern [Absorb(n), Squeeze(1)], digest = state[1], its low num_bits bits re-read in the other field.
// Parity unpinned against the reference binary (tests pin it against oracle/poseidon.py).
//
// Kernel: a sponge is a serial chain of ceil(n / 24) permutations, each 67 rounds deep, so it is a
// LATENCY problem.  One CTA of 25 warps: warp w owns state element w (replicated over its lanes), lane i
// of warp w keeps MDS[i][w] in registers.  A round = round constant + S-box (every warp in a full round,
// warp 0 only in a partial round) -> one shared-memory exchange + barrier -> 25 x 25 products, one per
// thread -> a 5-level shuffle reduction inside each warp.  Constants (Grain LFSR + Cauchy matrix) are
// derived on the host at first use, over the field the absorbed elements live in.
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace reef {

constexpr int RO_T = 25;
constexpr int RO_RF = 8;
constexpr int RO_RP = 59;
constexpr int RO_ROUNDS = RO_RF + RO_RP;

template <class C>
struct RoTables {
  Fe<C> rc[RO_ROUNDS * RO_T];   // Montgomery form
  Fe<C> mds[RO_T * RO_T];       // mds[i * T + j] = 1 / (i + j + T)
  // Optimised-but-equivalent schedule (the derivation of tools/gen_poseidon_consts.py, done on the host at first use for
  // width 25): partial-round constants pushed into lane 0, sparse partial-round matrices, lane 0 rescaled so that the
  // chain of a partial round is  u = w^5,  w' = u + sum_i beta[r][i] rest_i + kp[r+1],  rest_i' = rest_i + D[r][i] u.
  Fe<C> rc_full[RO_RF * RO_T];            // round constants of the 8 full rounds (the leftovers folded into round 4)
  Fe<C> kp[RO_RP + 1];
  Fe<C> beta[RO_RP * (RO_T - 1)];
  Fe<C> dcol[RO_RP * (RO_T - 1)];
  Fe<C> post[(RO_T - 1) * (RO_T - 1)];    // dense block applied to lanes 1.. after the last partial round
  Fe<C> lam_end;
  int fast_ok;                            // the host cross-check  optimised == textbook  passed
};

// ---------------------------------------------------------------------------------------
// host: constants
// ---------------------------------------------------------------------------------------
namespace {
struct Grain {
  int s[80];
  Grain(unsigned field, unsigned sbox, unsigned n, unsigned t, unsigned rf, unsigned rp) {
    const int widths[7] = {2, 4, 12, 12, 10, 10, 30};
    const unsigned vals[7] = {field, sbox, n, t, rf, rp, (1u << 30) - 1};
    int k = 0;
    for (int f = 0; f < 7; f++)
      for (int i = 0; i < widths[f]; i++) s[k++] = (vals[f] >> (widths[f] - 1 - i)) & 1;
    for (int i = 0; i < 160; i++) clock();
  }
  int clock() {
    int b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0];
    memmove(s, s + 1, 79 * sizeof(int));
    s[79] = b;
    return b;
  }
  int bit() {
    int b = clock();
    while (!b) {
      clock();
      b = clock();
    }
    return clock();
  }
};
}  // namespace

template <class C>
static bool below_modulus(const u32* x) {
  for (int i = 7; i >= 0; i--) {
    if (x[i] < modulus_limb<C>(i)) return true;
    if (x[i] > modulus_limb<C>(i)) return false;
  }
  return false;
}

// ---- optimised schedule: host derivation (Montgomery-form field elements throughout)
template <class C>
using Mat = std::vector<std::vector<Fe<C>>>;

template <class C>
static std::vector<Fe<C>> ro_matvec(const Mat<C>& A, const std::vector<Fe<C>>& v) {
  std::vector<Fe<C>> out(A.size(), fe_zero<C>());
  for (size_t i = 0; i < A.size(); i++)
    for (size_t k = 0; k < v.size(); k++) out[i] = fe_add<C>(out[i], mont_mul<C>(A[i][k], v[k]));
  return out;
}
template <class C>
static Mat<C> ro_matmul(const Mat<C>& A, const Mat<C>& B) {
  Mat<C> out(A.size(), std::vector<Fe<C>>(B[0].size(), fe_zero<C>()));
  for (size_t i = 0; i < A.size(); i++)
    for (size_t k = 0; k < B.size(); k++) {
      if (fe_is_zero<C>(A[i][k])) continue;
      for (size_t j = 0; j < B[0].size(); j++) out[i][j] = fe_add<C>(out[i][j], mont_mul<C>(A[i][k], B[k][j]));
    }
  return out;
}
template <class C>
static Mat<C> ro_inverse(const Mat<C>& A) {   // Gauss-Jordan
  const size_t n = A.size();
  Mat<C> M(n, std::vector<Fe<C>>(2 * n, fe_zero<C>()));
  for (size_t i = 0; i < n; i++) {
    for (size_t j = 0; j < n; j++) M[i][j] = A[i][j];
    M[i][n + i] = fe_one<C>();
  }
  for (size_t c = 0; c < n; c++) {
    size_t piv = c;
    while (piv < n && fe_is_zero<C>(M[piv][c])) piv++;
    std::swap(M[c], M[piv]);
    const Fe<C> inv = fe_inv<C>(M[c][c]);
    for (auto& x : M[c]) x = mont_mul<C>(x, inv);
    for (size_t r = 0; r < n; r++) {
      if (r == c || fe_is_zero<C>(M[r][c])) continue;
      const Fe<C> f = M[r][c];
      for (size_t j = 0; j < 2 * n; j++) M[r][j] = fe_sub<C>(M[r][j], mont_mul<C>(f, M[c][j]));
    }
  }
  Mat<C> out(n, std::vector<Fe<C>>(n));
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < n; j++) out[i][j] = M[i][n + j];
  return out;
}
template <class C>
static Fe<C> ro_pow5(const Fe<C>& x) {
  const Fe<C> x2 = mont_sqr<C>(x), x4 = mont_sqr<C>(x2);
  return mont_mul<C>(x4, x);
}

template <class C>
struct RoTables;
template <class C>
static void ro_permute_textbook(const RoTables<C>* K, Fe<C>* s);
template <class C>
static void ro_permute_fast_host(const RoTables<C>* K, Fe<C>* s);

template <class C>
static void ro_derive_fast(RoTables<C>* t) {
  const int T = RO_T, half = RO_RF / 2;
  Mat<C> mds(T, std::vector<Fe<C>>(T));
  for (int i = 0; i < T; i++)
    for (int j = 0; j < T; j++) mds[i][j] = t->mds[i * T + j];
  auto c = [&](int r, int i) { return t->rc[r * T + i]; };
  // constants: push the lanes 1.. of the partial-round constants forward through the matrix
  std::vector<Fe<C>> k(RO_RP), g(T, fe_zero<C>());
  k[0] = c(half, 0);
  for (int i = 1; i < T; i++) g[i] = c(half, i);
  for (int r = 1; r < RO_RP; r++) {
    std::vector<Fe<C>> v = ro_matvec<C>(mds, g);
    for (int i = 0; i < T; i++) v[i] = fe_add<C>(v[i], c(half + r, i));
    k[r] = v[0];
    g = v;
    g[0] = fe_zero<C>();
  }
  const std::vector<Fe<C>> tail = ro_matvec<C>(mds, g);
  // matrices: forward factorisation  M B_r = B_(r+1) Sp_r
  Mat<C> B(T, std::vector<Fe<C>>(T, fe_zero<C>()));
  for (int i = 0; i < T; i++) B[i][i] = fe_one<C>();
  Mat<C> sp_row(RO_RP), sp_col(RO_RP);
  for (int r = 0; r < RO_RP; r++) {
    const Mat<C> N = ro_matmul<C>(mds, B);
    Mat<C> Nh(T - 1, std::vector<Fe<C>>(T - 1));
    std::vector<Fe<C>> w(T - 1);
    for (int i = 1; i < T; i++) {
      w[i - 1] = N[i][0];
      for (int j = 1; j < T; j++) Nh[i - 1][j - 1] = N[i][j];
    }
    sp_col[r] = ro_matvec<C>(ro_inverse<C>(Nh), w);
    sp_row[r] = N[0];
    for (int i = 0; i < T; i++)
      for (int j = 0; j < T; j++) B[i][j] = (i == 0 || j == 0) ? (i == j ? fe_one<C>() : fe_zero<C>()) : Nh[i - 1][j - 1];
  }
  for (int i = 1; i < T; i++)
    for (int j = 1; j < T; j++) t->post[(i - 1) * (T - 1) + (j - 1)] = B[i][j];
  for (int r = 0; r < half; r++)
    for (int i = 0; i < T; i++) {
      t->rc_full[r * T + i] = c(r, i);
      t->rc_full[(half + r) * T + i] = c(half + RO_RP + r, i);
    }
  for (int i = 0; i < T; i++) t->rc_full[half * T + i] = fe_add<C>(t->rc_full[half * T + i], tail[i]);
  // rescaled lane 0:  lam_0 = 1, lam_(r+1) = a_r lam_r^5
  std::vector<Fe<C>> lam(RO_RP + 1);
  lam[0] = fe_one<C>();
  for (int r = 0; r < RO_RP; r++) lam[r + 1] = mont_mul<C>(sp_row[r][0], ro_pow5<C>(lam[r]));
  for (int r = 0; r < RO_RP; r++) {
    t->kp[r] = mont_mul<C>(k[r], fe_inv<C>(lam[r]));
    const Fe<C> il = fe_inv<C>(lam[r + 1]), l5 = ro_pow5<C>(lam[r]);
    for (int i = 0; i < T - 1; i++) {
      t->beta[r * (T - 1) + i] = mont_mul<C>(sp_row[r][i + 1], il);
      t->dcol[r * (T - 1) + i] = mont_mul<C>(sp_col[r][i], l5);
    }
  }
  t->kp[RO_RP] = fe_zero<C>();
  t->lam_end = lam[RO_RP];
  // cross-check on the host: the optimised schedule is the same function as the textbook rounds
  t->fast_ok = 1;
  for (int trial = 0; trial < 2 && t->fast_ok; trial++) {
    Fe<C> a[RO_T], b[RO_T];
    for (int i = 0; i < T; i++) a[i] = b[i] = mont_mul<C>(t->rc[(7 * i + 13 * trial + 3) % (RO_ROUNDS * T)], t->mds[(5 * i + trial) % (T * T)]);
    ro_permute_textbook<C>(t, a);
    ro_permute_fast_host<C>(t, b);
    for (int i = 0; i < T; i++)
      if (!fe_eq<C>(a[i], b[i])) t->fast_ok = 0;
  }
}

template <class C>
static void ro_tables_host(RoTables<C>* t) {
  Grain g(1, 1, 255, RO_T, RO_RF, RO_RP);
  int got = 0;
  while (got < RO_ROUNDS * RO_T) {
    Fe<C> x = fe_zero<C>();
    for (int b = 0; b < 255; b++) {   // big-endian bit stream
      for (int i = 7; i > 0; i--) x.v[i] = (x.v[i] << 1) | (x.v[i - 1] >> 31);
      x.v[0] = (x.v[0] << 1) | (u32)g.bit();
    }
    if (below_modulus<C>(x.v)) t->rc[got++] = to_mont<C>(x);
  }
  for (int i = 0; i < RO_T; i++)
    for (int j = 0; j < RO_T; j++) t->mds[i * RO_T + j] = fe_inv<C>(fe_from_u64<C>((u64)(i + j + RO_T)));
  ro_derive_fast<C>(t);
}

template <class C>
static const RoTables<C>* ro_tables_cached() {
  static RoTables<C>* T = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    T = new RoTables<C>;
    ro_tables_host<C>(T);
  });
  return T;
}

// textbook permutation on the host through the shared __host__ __device__ field code (test hook)
template <class C>
static void ro_permute_textbook(const RoTables<C>* K, Fe<C>* s) {
  for (int r = 0; r < RO_ROUNDS; r++) {
    const bool full = r < RO_RF / 2 || r >= RO_RF / 2 + RO_RP;
    for (int i = 0; i < RO_T; i++) s[i] = fe_add<C>(s[i], K->rc[r * RO_T + i]);
    for (int i = 0; i < (full ? RO_T : 1); i++) {
      Fe<C> x2 = mont_sqr<C>(s[i]), x4 = mont_sqr<C>(x2);
      s[i] = mont_mul<C>(x4, s[i]);
    }
    Fe<C> n[RO_T];
    for (int j = 0; j < RO_T; j++) {
      Fe<C> acc = fe_zero<C>();
      for (int i = 0; i < RO_T; i++) acc = fe_add<C>(acc, mont_mul<C>(s[i], K->mds[i * RO_T + j]));
      n[j] = acc;
    }
    for (int j = 0; j < RO_T; j++) s[j] = n[j];
  }
}
template <class C>
static void ro_permute_host(Fe<C>* s) { ro_permute_textbook<C>(ro_tables_cached<C>(), s); }

// the optimised schedule on the host (what k_poseidon_ro_fast runs)
template <class C>
static void ro_permute_fast_host(const RoTables<C>* K, Fe<C>* s) {
  const int T = RO_T, half = RO_RF / 2;
  auto full = [&](int r) {
    Fe<C> x[RO_T], n[RO_T];
    for (int i = 0; i < T; i++) x[i] = ro_pow5<C>(fe_add<C>(s[i], K->rc_full[r * T + i]));
    for (int j = 0; j < T; j++) {
      Fe<C> acc = fe_zero<C>();
      for (int i = 0; i < T; i++) acc = fe_add<C>(acc, mont_mul<C>(x[i], K->mds[i * T + j]));
      n[j] = acc;
    }
    for (int j = 0; j < T; j++) s[j] = n[j];
  };
  for (int r = 0; r < half; r++) full(r);
  Fe<C> w = fe_add<C>(s[0], K->kp[0]);
  for (int r = 0; r < RO_RP; r++) {
    const Fe<C> u = ro_pow5<C>(w);
    Fe<C> cs = K->kp[r + 1];
    for (int i = 0; i < T - 1; i++) cs = fe_add<C>(cs, mont_mul<C>(K->beta[r * (T - 1) + i], s[1 + i]));
    w = fe_add<C>(u, cs);
    for (int i = 0; i < T - 1; i++) s[1 + i] = fe_add<C>(s[1 + i], mont_mul<C>(K->dcol[r * (T - 1) + i], u));
  }
  Fe<C> rest[RO_T - 1];
  for (int i = 0; i < T - 1; i++) rest[i] = s[1 + i];
  s[0] = mont_mul<C>(K->lam_end, w);
  for (int j = 0; j < T - 1; j++) {
    Fe<C> acc = fe_zero<C>();
    for (int i = 0; i < T - 1; i++) acc = fe_add<C>(acc, mont_mul<C>(K->post[j * (T - 1) + i], rest[i]));
    s[1 + j] = acc;
  }
  for (int r = half; r < RO_RF; r++) full(r);
}

static void ro_tag(uint64_t n, uint8_t tag[32]) {
  uint32_t ops[2] = {(1u << 31) | (uint32_t)n, 1u};
  io_pattern_tag_le32(ops, 2, 0, tag);
}

template <class C>
static Fe<C> fe_from_le32(const uint8_t* b) {
  Fe<C> x;
  for (int i = 0; i < 8; i++)
    x.v[i] = (u32)b[4 * i] | ((u32)b[4 * i + 1] << 8) | ((u32)b[4 * i + 2] << 16) | ((u32)b[4 * i + 3] << 24);
  return x;
}

template <class C>
static void ro_host(const uint8_t* elems, uint64_t n, uint8_t out[32]) {
  uint8_t tag[32];
  ro_tag(n, tag);
  Fe<C> s[RO_T];
  s[0] = to_mont<C>(fe_from_le32<C>(tag));
  for (int i = 1; i < RO_T; i++) s[i] = fe_zero<C>();
  int apos = 0;
  for (uint64_t e = 0; e < n; e++) {
    if (apos == RO_T - 1) {
      ro_permute_host<C>(s);
      apos = 0;
    }
    s[1 + apos] = fe_add<C>(s[1 + apos], to_mont<C>(fe_from_le32<C>(elems + 32 * e)));
    apos++;
  }
  ro_permute_host<C>(s);
  Fe<C> o = from_mont<C>(s[1]);
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) out[4 * i + k] = (uint8_t)(o.v[i] >> (8 * k));
}

void poseidon_ro_host(int field, const uint8_t* elems, uint64_t n, uint8_t out[32]) {
  if (field == 0) ro_host<FqCfg>(elems, n, out);
  else ro_host<FpCfg>(elems, n, out);
}

int poseidon_ro_fast_ok_host(int field) { return field == 0 ? ro_tables_cached<FqCfg>()->fast_ok : ro_tables_cached<FpCfg>()->fast_ok; }

void poseidon_ro_constants_host(int field, uint8_t* rc_out, uint8_t* mds_out) {
  auto dump = [](const u32* v, uint8_t* b) {
    for (int i = 0; i < 8; i++)
      for (int k = 0; k < 4; k++) b[4 * i + k] = (uint8_t)(v[i] >> (8 * k));
  };
  if (field == 0) {
    const RoTables<FqCfg>* K = ro_tables_cached<FqCfg>();
    for (int i = 0; i < RO_ROUNDS * RO_T; i++) dump(from_mont<FqCfg>(K->rc[i]).v, rc_out + 32 * i);
    for (int i = 0; i < RO_T * RO_T; i++) dump(from_mont<FqCfg>(K->mds[i]).v, mds_out + 32 * i);
  } else {
    const RoTables<FpCfg>* K = ro_tables_cached<FpCfg>();
    for (int i = 0; i < RO_ROUNDS * RO_T; i++) dump(from_mont<FpCfg>(K->rc[i]).v, rc_out + 32 * i);
    for (int i = 0; i < RO_T * RO_T; i++) dump(from_mont<FpCfg>(K->mds[i]).v, mds_out + 32 * i);
  }
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
// in: canonical elements.  triples = 0: n elements of 32 B.  triples = 1: n / 3 affine points of 64 B
// (x || y, all-zero = identity); element 3k, 3k+1, 3k+2 = x, y, is_infinity of point k (`to_coordinates`).
template <class C>
__global__ void __launch_bounds__(RO_T * 32) k_poseidon_ro(const Fe<C>* __restrict__ in, uint64_t n, int triples, Fe<C> tag_mont,
                                                          const RoTables<C>* __restrict__ K, Fe<C>* __restrict__ out) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ Fe<C> sb[2][RO_T];
  const Fe<C> zero = fe_zero<C>();
  const Fe<C> m = lane < RO_T ? ld256(&K->mds[lane * RO_T + w]) : zero;
  Fe<C> s = w == 0 ? tag_mont : zero;
  const uint64_t n_perm = (n + RO_T - 2) / (RO_T - 1);
#pragma unroll 1
  for (uint64_t p = 0; p < n_perm; p++) {
    if (w >= 1) {                                     // absorb: rate element w - 1 of block p
      const uint64_t e = p * (RO_T - 1) + (uint64_t)(w - 1);
      if (e < n) {
        Fe<C> x;
        if (!triples) {
          x = ld256(in + e);
        } else {
          const uint64_t k = e / 3;
          const int comp = (int)(e - 3 * k);
          const Fe<C> px = ld256(in + 2 * k), py = ld256(in + 2 * k + 1);
          const bool inf = fe_is_zero<C>(px) && fe_is_zero<C>(py);
          x = comp == 0 ? px : py;
          if (comp == 2) {
            x = zero;
            x.v[0] = inf ? 1u : 0u;
          }
        }
        s = fe_add<C>(s, to_mont<C>(x));
      }
    }
#pragma unroll 1
    for (int r = 0; r < RO_ROUNDS; r++) {
      s = fe_add<C>(s, ld256(&K->rc[r * RO_T + w]));
      const bool full = r < RO_RF / 2 || r >= RO_RF / 2 + RO_RP;
      if (full || w == 0) {                           // warp-uniform
        const Fe<C> x2 = mont_sqr<C>(s), x4 = mont_sqr<C>(x2);
        s = mont_mul<C>(x4, s);
      }
      if (lane == 0) sb[r & 1][w] = s;
      __syncthreads();
      const Fe<C> x = lane < RO_T ? sb[r & 1][lane] : zero;
      s = warp_sum_fe<C>(mont_mul<C>(x, m));          // new state element w, in every lane
    }
  }
  if (w == 1 && lane == 0) st256(out, from_mont<C>(s));
}

// The same sponge through the optimised schedule (RoTables: rc_full / kp / beta / dcol / post / lam_end).  Full rounds
// as above.  The 59 partial rounds run on TWO warps that meet at one named barrier per round: warp 0 is the chain
// (u = w^5, three dependent multiplications; w' = u + c_r + kp[r+1]), warp 1 keeps rest_1..24 in its lanes and supplies
// c_r = sum_i beta[r][i] rest_i (one multiplication + a shuffle reduction, computed WHILE warp 0 raises w to the fifth
// power) and then takes u for rest_i += D[r][i] u.  49 products per partial round instead of 625, no CTA-wide barrier.
template <class C>
__device__ __forceinline__ Fe<C> ro_pow5_dev(const Fe<C>& x) {
  const Fe<C> x2 = mont_sqr<C>(x), x4 = mont_sqr<C>(x2);
  return mont_mul<C>(x4, x);
}

template <class C>
__global__ void __launch_bounds__(RO_T * 32) k_poseidon_ro_fast(const Fe<C>* __restrict__ in, uint64_t n, int triples, Fe<C> tag_mont,
                                                               const RoTables<C>* __restrict__ K, Fe<C>* __restrict__ out) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ Fe<C> sb[2][RO_T];
  __shared__ Fe<C> rest_sh[RO_T - 1];
  __shared__ Fe<C> gsh[RO_T];
  __shared__ Fe<C> xch_u[2], xch_c[2];
  const Fe<C> zero = fe_zero<C>();
  const Fe<C> m = lane < RO_T ? ld256(&K->mds[lane * RO_T + w]) : zero;
  const Fe<C> pm = (w >= 1 && lane < RO_T - 1) ? ld256(&K->post[(w - 1) * (RO_T - 1) + lane]) : zero;
  Fe<C> s = w == 0 ? tag_mont : zero;
  const uint64_t n_perm = (n + RO_T - 2) / (RO_T - 1);
#pragma unroll 1
  for (uint64_t p = 0; p < n_perm; p++) {
    if (w >= 1) {                                     // absorb: rate element w - 1 of block p
      const uint64_t e = p * (RO_T - 1) + (uint64_t)(w - 1);
      if (e < n) {
        Fe<C> x;
        if (!triples) {
          x = ld256(in + e);
        } else {
          const uint64_t k = e / 3;
          const int comp = (int)(e - 3 * k);
          const Fe<C> px = ld256(in + 2 * k), py = ld256(in + 2 * k + 1);
          const bool inf = fe_is_zero<C>(px) && fe_is_zero<C>(py);
          x = comp == 0 ? px : py;
          if (comp == 2) {
            x = zero;
            x.v[0] = inf ? 1u : 0u;
          }
        }
        s = fe_add<C>(s, to_mont<C>(x));
      }
    }
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
#pragma unroll 1
      for (int r = half * (RO_RF / 2); r < (half + 1) * (RO_RF / 2); r++) {      // four full rounds
        // the 25 S-boxes of a round on the 25 lanes of ONE warp (3 warp-wide multiplications instead of 3 x 25: every
        // warp raising its own element to the fifth power on all of its lanes kept the four schedulers busy with copies)
        if (lane == 0) gsh[w] = s;
        __syncthreads();
        if (w == 0 && lane < RO_T) sb[r & 1][lane] = ro_pow5_dev<C>(fe_add<C>(gsh[lane], ld256(&K->rc_full[r * RO_T + lane])));
        __syncthreads();
        const Fe<C> x = lane < RO_T ? sb[r & 1][lane] : zero;
        s = warp_sum_fe<C>(mont_mul<C>(x, m));
      }
      if (half == 1) break;
      // ---- 59 partial rounds on warps 0 (chain) and 1 (side)
      if (w >= 1 && lane == 0) rest_sh[w - 1] = s;
      __syncthreads();
      if (w == 0) {
        Fe<C> wv = fe_add<C>(s, ld256(&K->kp[0]));
        Fe<C> kn = ld256(&K->kp[1]);
#pragma unroll 1
        for (int r = 0; r < RO_RP; r++) {
          const Fe<C> k_next = ld256(&K->kp[r + 2 <= RO_RP ? r + 2 : RO_RP]);     // prefetch
          const Fe<C> u = ro_pow5_dev<C>(wv);
          if (lane == 0) xch_u[r & 1] = u;
          asm volatile("bar.sync 1, 64;" ::: "memory");
          wv = fe_add<C>(fe_add<C>(u, xch_c[r & 1]), kn);
          kn = k_next;
        }
        s = mont_mul<C>(ld256(&K->lam_end), wv);
      } else if (w == 1) {
        const int i = lane < RO_T - 1 ? lane : 0;
        Fe<C> rest = lane < RO_T - 1 ? rest_sh[lane] : zero;
        Fe<C> be = ld256(&K->beta[i]), dc = ld256(&K->dcol[i]);
#pragma unroll 1
        for (int r = 0; r < RO_RP; r++) {
          const int rn = r + 1 < RO_RP ? r + 1 : r;
          const Fe<C> be_n = ld256(&K->beta[rn * (RO_T - 1) + i]), dc_n = ld256(&K->dcol[rn * (RO_T - 1) + i]);   // prefetch
          const Fe<C> cs = warp_sum_fe<C>(lane < RO_T - 1 ? mont_mul<C>(be, rest) : zero);
          if (lane == 0) xch_c[r & 1] = cs;
          asm volatile("bar.sync 1, 64;" ::: "memory");
          const Fe<C> u = xch_u[r & 1];
          rest = fe_add<C>(rest, mont_mul<C>(dc, u));
          be = be_n;
          dc = dc_n;
        }
        if (lane < RO_T - 1) rest_sh[lane] = rest;
      }
      __syncthreads();
      // dense block on lanes 1..: s_j = sum_i post[j-1][i] rest_i
      if (w >= 1) s = warp_sum_fe<C>(lane < RO_T - 1 ? mont_mul<C>(pm, rest_sh[lane]) : zero);
      __syncthreads();                                 // rest_sh / sb are reused by the next rounds
    }
  }
  if (w == 1 && lane == 0) st256(out, from_mont<C>(s));
}

// ---------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------
template <class C>
static int ro_device_tables(reef_ctx* c, const RoTables<C>** out) {
  void*& slot = std::is_same<C, FqCfg>::value ? c->d_ro_fq : c->d_ro_fp;
  if (!slot) {
    const RoTables<C>* h = ro_tables_cached<C>();
    REEF_CUDA(cudaMalloc(&slot, sizeof(RoTables<C>)));
    REEF_CUDA(cudaMemcpyAsync(slot, h, sizeof(RoTables<C>), cudaMemcpyHostToDevice, c->stream));
  }
  *out = (const RoTables<C>*)slot;
  return REEF_OK;
}

template <class C>
static int ro_launch(reef_ctx* c, const void* d_in, uint64_t n, int triples, void* d_out) {
  const RoTables<C>* K;
  int rc = ro_device_tables<C>(c, &K);
  if (rc) return rc;
  uint8_t tag[32];
  ro_tag(n, tag);
  const Fe<C> tag_mont = to_mont<C>(fe_from_le32<C>(tag));
  ProfScope ps(c, PROF_POSEIDON, n);
  // REEF_RO_TEXTBOOK=1 keeps the textbook rounds (A/B, and the fallback if the host cross-check of the derived tables failed)
  static const bool textbook = getenv("REEF_RO_TEXTBOOK") && atoi(getenv("REEF_RO_TEXTBOOK")) != 0;
  if (textbook || !ro_tables_cached<C>()->fast_ok)
    k_poseidon_ro<C><<<1, RO_T * 32, 0, c->stream>>>((const Fe<C>*)d_in, n, triples, tag_mont, K, (Fe<C>*)d_out);
  else
    k_poseidon_ro_fast<C><<<1, RO_T * 32, 0, c->stream>>>((const Fe<C>*)d_in, n, triples, tag_mont, K, (Fe<C>*)d_out);
  REEF_LAUNCHED();
  return REEF_OK;
}

int launch_poseidon_ro(reef_ctx* c, int field, const void* d_in, uint64_t n, int triples, void* d_out) {
  return field == 0 ? ro_launch<FqCfg>(c, d_in, n, triples, d_out) : ro_launch<FpCfg>(c, d_in, n, triples, d_out);
}

}  // namespace reef
