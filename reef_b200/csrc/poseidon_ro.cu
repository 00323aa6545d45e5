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
static void ro_permute_host(Fe<C>* s) {
  const RoTables<C>* K = ro_tables_cached<C>();
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
  k_poseidon_ro<C><<<1, RO_T * 32, 0, c->stream>>>((const Fe<C>*)d_in, n, triples, tag_mont, K, (Fe<C>*)d_out);
  REEF_LAUNCHED();
  return REEF_OK;
}

int launch_poseidon_ro(reef_ctx* c, int field, const void* d_in, uint64_t n, int triples, void* d_out) {
  return field == 0 ? ro_launch<FqCfg>(c, d_in, n, triples, d_out) : ro_launch<FpCfg>(c, d_in, n, triples, d_out);
}

}  // namespace reef
