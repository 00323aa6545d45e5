// extern "C" surface of libreef_b200.so + the host-side logic of the path (logmn,
// doc_transform, combined_q packing, transcript assembly, Merkle path witnesses).
// See include/reef_b200.h for the contract and the reference interfaces each entry replaces.
#include <cuda.h>
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/reef_b200.h"
#include "../../include/reef_b200_testing.h"
#include "common.cuh"
#include "kernels.h"

using namespace reef;

// ---------------------------------------------------------------------------------------
// error plumbing / context scratch
// ---------------------------------------------------------------------------------------
namespace reef {
static thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

static int grow(void** p, size_t* cap, size_t bytes, bool pinned) {
  if (*cap >= bytes) return REEF_OK;
  size_t want = bytes + bytes / 4 + 4096;
  if (*p) {
    if (pinned) cudaFreeHost(*p);
    else cudaFree(*p);
    *p = nullptr;
    *cap = 0;
  }
  cudaError_t e = pinned ? cudaMallocHost(p, want) : cudaMalloc(p, want);
  if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("allocation of ") + std::to_string(want) + " bytes: " + cudaGetErrorString(e));
  *cap = want;
  return REEF_OK;
}
int ctx_scratch(reef_ctx* c, size_t bytes, void** out) {
  // a grow may free memory still in use by queued kernels: drain first
  if (c->scratch_bytes < bytes) cudaStreamSynchronize(c->stream);
  int rc = grow(&c->scratch, &c->scratch_bytes, bytes, false);
  *out = c->scratch;
  return rc;
}
int ctx_scratch2(reef_ctx* c, size_t bytes, void** out) {
  if (c->scratch2_bytes < bytes) cudaStreamSynchronize(c->stream);
  int rc = grow(&c->scratch2, &c->scratch2_bytes, bytes, false);
  *out = c->scratch2;
  return rc;
}
int ctx_stage(reef_ctx* c, size_t bytes, void** out) {
  if (c->h_stage_bytes < bytes) cudaStreamSynchronize(c->stream);
  int rc = grow(&c->h_stage, &c->h_stage_bytes, bytes, true);
  *out = c->h_stage;
  return rc;
}

static void ctx_destroy(reef_ctx* c) {
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->scratch) cudaFree(c->scratch);
  if (c->scratch2) cudaFree(c->scratch2);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->d_pos) cudaFree(c->d_pos);
  if (c->d_lp) cudaFree(c->d_lp);
  if (c->d_ro_fq) cudaFree(c->d_ro_fq);
  if (c->d_ro_fp) cudaFree(c->d_ro_fp);
  for (auto& kv : c->table_cache) cudaFree(kv.second);
  if (c->shard_cache) cudaFree(c->shard_cache);
  for (void* p : c->mb_ipc_opened) cudaIpcCloseMemHandle(p);
  if (c->mb_mine) cudaFree(c->mb_mine);
  if (c->mb_peers_dev) cudaFree(c->mb_peers_dev);
  if (c->mb_err_dev) cudaFree(c->mb_err_dev);
  for (auto& r : c->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}
void ctx_retain(reef_ctx* c) { c->refs.fetch_add(1, std::memory_order_relaxed); }
void ctx_release(reef_ctx* c) {
  if (c->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) ctx_destroy(c);
}
}  // namespace reef

struct reef_table {
  reef_ctx* ctx;
  void* d;
  uint64_t n_pad, n_orig;
  int is_u32;
  int owns;
  uint8_t first[32];  // table[0], canonical (default prev_running_v, r1cs.rs:2198-2201)
  cudaEvent_t ready = nullptr;   // reef_table_upload_u32_async: recorded on the copy stream behind the upload
};

// every consumer of a table orders its stream behind a pending asynchronous upload
static void table_wait(reef_ctx* c, const reef_table* t) {
  if (t && t->ready) cudaStreamWaitEvent(c->stream, t->ready, 0);
}

struct reef_sponge {
  reef_ctx* ctx;
  void* d_state;
  void* d_buf;
  size_t buf_bytes;
  std::vector<uint32_t> ops;
  size_t io;
};

static const uint8_t FQ_LE[32] = {0x01, 0x00, 0x00, 0x00, 0x21, 0xeb, 0x46, 0x8c, 0xdd, 0xa8, 0x94, 0x09, 0xfc, 0x98, 0x46, 0x22,
                                  0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x00, 0x00, 0x00, 0x40};

static bool is_canonical_fq(const uint8_t* b) {
  for (int i = 31; i >= 0; i--) {
    if (b[i] < FQ_LE[i]) return true;
    if (b[i] > FQ_LE[i]) return false;
  }
  return false;  // == q
}

static int check_canonical(const uint8_t* b, size_t n, const char* what) {
  for (size_t i = 0; i < n; i++)
    if (!is_canonical_fq(b + i * 32)) return fail(REEF_EINVAL, std::string(what) + ": element " + std::to_string(i) + " is not a canonical Fq value");
  return REEF_OK;
}

static void load_le(uint32_t* l, const uint8_t* b) {
  for (int i = 0; i < 8; i++)
    l[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
}
static void store_le(uint8_t* b, const uint32_t* l) {
  for (int i = 0; i < 8; i++)
    for (int k = 0; k < 4; k++) b[4 * i + k] = (uint8_t)(l[i] >> (8 * k));
}

template <class C>
static void hosttest_field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Fe<C> x, y, r;
  load_le(x.v, a);
  load_le(y.v, b);
  switch (op) {
    case 0: r = from_mont<C>(mont_mul<C>(to_mont<C>(x), to_mont<C>(y))); break;  // a*b
    case 1: r = fe_add<C>(x, y); break;
    case 2: r = fe_sub<C>(x, y); break;
    case 3: r = from_mont<C>(fe_inv<C>(to_mont<C>(x))); break;                    // 1/a
    case 4: {                                                                     // lazy: a*b + a*a + b*b, canonical
      Wide17 w;
      wide_zero(w);
      wide_mac(w, x.v, y.v);
      wide_mac(w, x.v, x.v);
      wide_mac(w, y.v, y.v);
      r = wide_reduce_canonical<C>(w);
      break;
    }
    case 5: {                                                                     // stress the 17th limb: 40000 * (a*b)
      Wide17 w;
      wide_zero(w);
      for (int k = 0; k < 40000; k++) wide_mac(w, x.v, y.v);
      r = wide_reduce_canonical<C>(w);
      break;
    }
    case 6: {                                                                     // small mac: a[0] * b
      Wide17 w;
      wide_zero(w);
      wide_mac_small(w, x.v[0], y.v);
      r = wide_reduce_canonical<C>(w);
      break;
    }
    case 7: r = from_mont<C>(mont_sqr<C>(to_mont<C>(x))); break;                  // a*a
    case 8: r = f29_to_canonical<C>(mul29<C>(f29_from_canonical<C>(x), f29_from_canonical<C>(y))); break;  // 29-bit limbs
    case 9: {                                                                     // 29-bit limbs through Montgomery-256
      const F29 k = f29_const_2_266<C>();
      F29 a = f29_from_mont256<C>(to_mont<C>(x), k), b = f29_from_mont256<C>(to_mont<C>(y), k);
      F29 s = f29_relax(f29_add_lazy(f29_add_lazy(a, b), a));                    // 2a + b, relaxed
      F29 t = mul29<C>(f29_add_lazy(a, b), s);                                   // (a+b)(2a+b), lazy operand
      const F29 q = sqr29<C>(f29_add_lazy(a, b));                                // (a+b)^2, lazy operand
      t = f29_relax(f29_add_lazy(t, q));                                         // (a+b)(2a+b) + (a+b)^2
      r = from_mont<C>(f29_to_mont256<C>(t));
      break;
    }
    default: r = fe_zero<C>();
  }
  store_le(out, r.v);
}


extern "C" {

// ---------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------
int reef_abi_version(void) { return 1; }
uint64_t reef_launch_count(void) { return (uint64_t)g_launches.load(); }
const char* reef_last_error(void) { return g_last_error.c_str(); }

static int init_impl(int device, int latency_critical, reef_ctx** out);
int reef_init(int device, reef_ctx** out) { return init_impl(device, 0, out); }
// latency_critical: 1 = highest stream priority, 0 = background (lowest), 2 = background but ahead of the other
// background contexts (the longer of two concurrent commitment chains)
int reef_init_prio(int device, int latency_critical, reef_ctx** out) {
  int rc = init_impl(device, latency_critical == 1 ? 1 : (latency_critical == 2 ? 3 : 2), out);
  if (rc == REEF_OK && latency_critical == 0) (*out)->polite = getenv("REEF_MSM_POLITE") && atoi(getenv("REEF_MSM_POLITE")) != 0;
  return rc;
}

// ---------------------------------------------------------------------------------------
// SM partition for BACKGROUND contexts (reef_init_prio(.., 0, ..)).  The Fiat-Shamir kernels of a latency-critical
// context are single CTAs that want an SM to themselves; while the grids of a background MSM fill the chip such a
// CTA waits for a whole SM to drain -- measured: 0.6 ms of a 5.8 ms pass (profiles/r02_summary.md).  Background
// streams can therefore be created inside a CUDA green context that owns all but REEF_RESERVE_SMS SMs (12 = what is
// left when 136 of 148 SMs are split off in groups of 8): their kernels can never occupy the remaining SMs, which the
// latency-critical kernels (primary context: every SM) then find empty.  OPT-IN (REEF_RESERVE_SMS unset or 0 = an
// ordinary low-priority stream): it recovers about half of the interference (5.72 -> 5.65 ms per pass), and Nsight
// Compute's launch-list mode fails on kernels of a green context (LaunchFailed), so profiling runs keep it off.
// Driver API through dlopen (the library must still load, for the symbol checks, where no driver is installed); any
// failure falls back to the ordinary stream.
// ---------------------------------------------------------------------------------------
namespace {
struct GreenPartition {
  bool tried = false;
  CUgreenCtx ctx = nullptr;
  unsigned sm_count = 0;
};
std::mutex g_green_mu;
GreenPartition g_green[16];

#define REEF_DL(h, name, fn) (((fn) = reinterpret_cast<decltype(fn)>(dlsym((h), (name)))) != nullptr)

// stream of the device's background partition, or nullptr
cudaStream_t green_stream(int device, int sm_total, int priority, unsigned* sms_out) {
  if (device < 0 || device >= 16) return nullptr;
  const char* e = getenv("REEF_RESERVE_SMS");
  const int reserve = e ? atoi(e) : 0;
  if (reserve <= 0 || reserve >= sm_total) return nullptr;
  static void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return nullptr;
  static CUresult (*p_cuDeviceGet)(CUdevice*, int);
  static CUresult (*p_cuDeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
  static CUresult (*p_cuDevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  static CUresult (*p_cuDevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  static CUresult (*p_cuGreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  static CUresult (*p_cuGreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);
  static const bool ok = REEF_DL(lib, "cuDeviceGet", p_cuDeviceGet) && REEF_DL(lib, "cuDeviceGetDevResource", p_cuDeviceGetDevResource) &&
                         REEF_DL(lib, "cuDevSmResourceSplitByCount", p_cuDevSmResourceSplitByCount) &&
                         REEF_DL(lib, "cuDevResourceGenerateDesc", p_cuDevResourceGenerateDesc) &&
                         REEF_DL(lib, "cuGreenCtxCreate", p_cuGreenCtxCreate) && REEF_DL(lib, "cuGreenCtxStreamCreate", p_cuGreenCtxStreamCreate);
  if (!ok) return nullptr;
  std::lock_guard<std::mutex> lk(g_green_mu);
  GreenPartition& G = g_green[device];
  if (!G.tried) {
    G.tried = true;
    cudaFree(0);                                                   // the primary context exists and is current
    CUdevice dev;
    CUdevResource all, part, rest;
    CUdevResourceDesc desc;
    unsigned groups = 1;
    const unsigned want = (unsigned)(sm_total - reserve);
    if (p_cuDeviceGet(&dev, device) == CUDA_SUCCESS && p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
        p_cuDevSmResourceSplitByCount(&part, &groups, &all, &rest, 0, want) == CUDA_SUCCESS && groups == 1 &&
        part.sm.smCount < (unsigned)sm_total && p_cuDevResourceGenerateDesc(&desc, &part, 1) == CUDA_SUCCESS &&
        p_cuGreenCtxCreate(&G.ctx, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS)
      G.sm_count = part.sm.smCount;
    else
      G.ctx = nullptr;
  }
  if (!G.ctx) return nullptr;
  CUstream s = nullptr;
  if (p_cuGreenCtxStreamCreate(&s, G.ctx, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) return nullptr;
  *sms_out = G.sm_count;
  return (cudaStream_t)s;
}
}  // namespace

static int init_impl(int device, int latency_critical, reef_ctx** out) {
  REEF_REQUIRE(out != nullptr, REEF_EINVAL, "reef_init: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(REEF_ECUDA, std::string("reef_init: no CUDA device (") + cudaGetErrorString(e) + "); libreef_b200 has no CPU fallback");
  REEF_REQUIRE(device >= 0 && device < count, REEF_EINVAL, "reef_init: device index out of range");
  REEF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  REEF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(REEF_ECUDA, std::string("reef_init: device '") + prop.name + "' is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                "; this library is built for sm_100a only");
  reef_ctx* c = new reef_ctx;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  // latency-critical contexts (the Fiat-Shamir chains) get the highest stream priority: when an SM frees up, their
  // pending CTAs are placed before those of the throughput streams (the fold commitments)
  // (latency_critical: 0 = plain reef_init, 1 = latency-critical, 2 = background)
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  cudaError_t se = cudaSuccess;
  if (latency_critical == 2 || latency_critical == 3) {
    unsigned sms = 0;
    c->stream = green_stream(device, c->sm_count, prio_lo, &sms);
    if (c->stream) c->partition_sms = sms;
  }
  if (!c->stream) {
    // CUDA priorities: numerically lower = higher priority; prio_hi <= prio_lo
    const int prio = latency_critical == 1 ? prio_hi : (latency_critical == 3 ? (prio_lo + prio_hi) / 2 : prio_lo);
    se = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio);
  }
  if (se != cudaSuccess) {
    c->stream = nullptr;
    ctx_release(c);
    return fail(REEF_ECUDA, std::string("reef_init: cudaStreamCreateWithFlags: ") + cudaGetErrorString(se));
  }
  int rc = poseidon_upload_constants(c);
  if (rc) {
    ctx_release(c);
    return rc;
  }
  *out = c;
  return REEF_OK;
}

/* SMs this context's kernels may run on: the whole chip, or the background partition (reef_init_prio(.., 0, ..)) */
uint32_t reef_ctx_sm_count(const reef_ctx* c) { return c ? (c->partition_sms ? c->partition_sms : (uint32_t)c->sm_count) : 0; }

// Drops the caller's reference.  Child handles (tables, bases, sponges, sessions) created from this
// context stay valid for their *_free call and keep the context's resources alive until the last
// of them is freed: any destruction order is safe (e.g. Rust `Drop` order, Python finalisers).
void reef_shutdown(reef_ctx* c) {
  if (!c) return;
  if (c->closed.exchange(true)) return;   // double shutdown while children keep it alive: ignore
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
  }
  ctx_release(c);
}

int reef_sync(reef_ctx* c) {
  REEF_REQUIRE(c != nullptr, REEF_EINVAL, "reef_sync: ctx is NULL");
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

void* reef_stream(reef_ctx* c) { return c ? (void*)c->stream : nullptr; }

int reef_profile_enable(reef_ctx* c, int on) {
  REEF_REQUIRE(c != nullptr, REEF_EINVAL, "reef_profile_enable: ctx is NULL");
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  c->prof.clear();
  c->profile = on != 0;
  return REEF_OK;
}

int reef_profile_read(reef_ctx* c, uint32_t n_classes, uint64_t* counts, uint64_t* units, double* ms) {
  REEF_REQUIRE(c && counts && units && ms, REEF_EINVAL, "reef_profile_read: NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  for (uint32_t k = 0; k < n_classes; k++) {
    counts[k] = 0;
    units[k] = 0;
    ms[k] = 0.0;
  }
  for (auto& r : c->prof) {
    float t = 0.f;
    REEF_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    if ((uint32_t)r.cls < n_classes) {
      counts[r.cls] += 1;
      units[r.cls] += r.units;
      ms[r.cls] += (double)t;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  c->prof.clear();
  return REEF_OK;
}

// ---------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------
uint32_t reef_logmn(uint64_t mn) {
  if (mn == 1) return 1;
  float f = (float)mn;  // `mn as f32`: round-to-nearest-even like Rust's cast
  return (uint32_t)std::ceil(std::log2(f));
}

int reef_doc_transform(const uint32_t* ab, uint32_t ab_len, const uint32_t* doc, uint64_t doc_len, uint64_t* out_udoc,
                       uint64_t out_cap, uint64_t* out_len) {
  REEF_REQUIRE(ab && (doc || doc_len == 0) && out_len, REEF_EINVAL, "reef_doc_transform: NULL argument");
  // FxHashMap<Option<char>, usize>: later inserts overwrite (framework.rs:979-986)
  std::unordered_map<uint32_t, uint64_t> num_ab;
  uint64_t i = 0;
  for (uint32_t k = 0; k < ab_len; k++) {
    num_ab[ab[k]] = i;
    i += 1;
  }
  const uint64_t epsilon = i + 1;
  num_ab[26u] = i + 2;  // EOF
  const uint64_t total = doc_len + 2;
  const uint32_t lg = reef_logmn(total);
  REEF_REQUIRE(lg < 63, REEF_EINVAL, "reef_doc_transform: document too long");
  const uint64_t padded = (uint64_t)1 << lg;
  if (padded < total) return fail(REEF_EASSERT, "reef_doc_transform: attempt to subtract with overflow (f32 logmn)");
  *out_len = padded;
  REEF_REQUIRE(out_udoc != nullptr && out_cap >= padded, REEF_EINVAL, "reef_doc_transform: output buffer too small");
  for (uint64_t k = 0; k < doc_len; k++) {
    auto it = num_ab.find(doc[k]);
    if (it == num_ab.end()) return fail(REEF_EASSERT, "Character in document that's not in alphabet");
    out_udoc[k] = it->second;
  }
  out_udoc[doc_len] = num_ab[26u];
  out_udoc[doc_len + 1] = epsilon;
  for (uint64_t k = total; k < padded; k++) out_udoc[k] = 0;
  return REEF_OK;
}

// 256-bit little-endian accumulator for the bit packing below
static void set_bit_le(uint8_t* x, uint32_t bit) { x[bit >> 3] |= (uint8_t)(1u << (bit & 7)); }

int reef_combined_q(const uint64_t* q, uint32_t m, uint32_t sc_l, uint8_t* out, uint32_t out_cap_elems, uint32_t* num_cqs_out) {
  REEF_REQUIRE((q || m == 0) && num_cqs_out, REEF_EINVAL, "reef_combined_q: NULL argument");
  const uint32_t num_cqs = (uint32_t)std::ceil(((double)((uint64_t)m * sc_l)) / 254.0);
  *num_cqs_out = num_cqs;
  REEF_REQUIRE(num_cqs == 0 || (out && out_cap_elems >= num_cqs), REEF_EINVAL, "reef_combined_q: output buffer too small");
  std::vector<uint8_t> res;
  uint32_t cq = 0;
  uint8_t cur[32];
  while (cq < num_cqs) {
    memset(cur, 0, 32);
    uint32_t next_slot = 0;  // exponent of the next power of two
    for (uint32_t i = 0; i < m; i++) {
      uint32_t j = 0;
      for (int32_t bit = (int32_t)sc_l - 1; bit >= 0; bit--) {  // qjs reversed: MSB first
        const uint32_t qj = bit < 64 ? (uint32_t)((q[i] >> bit) & 1) : 0;
        if ((uint64_t)i * sc_l + j >= (uint64_t)254 * (cq + 1) || (i == m - 1 && j == sc_l - 1)) {
          cq += 1;
          res.insert(res.end(), cur, cur + 32);
          memset(cur, 0, 32);
          next_slot = 0;
        } else {
          if (qj) {
            REEF_REQUIRE(next_slot < 254, REEF_EASSERT, "reef_combined_q: slot overflow");
            set_bit_le(cur, next_slot);
          }
          next_slot += 1;
        }
        j += 1;
      }
    }
    if (m == 0 || sc_l == 0) break;
  }
  if (res.size() / 32 != num_cqs) return fail(REEF_EASSERT, "assertion failed: num_cqs == combined_qs.len()");
  if (num_cqs) memcpy(out, res.data(), res.size());
  return REEF_OK;
}

int reef_io_pattern_tag(const uint32_t* ops, uint32_t n_ops, uint32_t domain_separator, uint8_t out[32]) {
  REEF_REQUIRE((ops || n_ops == 0) && out, REEF_EINVAL, "reef_io_pattern_tag: NULL argument");
  io_pattern_tag_le32(ops, n_ops, domain_separator, out);
  return REEF_OK;
}

// ---------------------------------------------------------------------------------------
// B2: Poseidon
// ---------------------------------------------------------------------------------------
int reef_poseidon_hash(reef_ctx* c, const uint8_t* in, uint32_t arity, uint64_t n, uint8_t* out) {
  REEF_REQUIRE(c && (n == 0 || (in && out)), REEF_EINVAL, "reef_poseidon_hash: NULL argument");
  REEF_REQUIRE(arity == 2 || arity == 4, REEF_EINVAL, "reef_poseidon_hash: arity must be 2 or 4");
  if (n == 0) return REEF_OK;
  int rc = check_canonical(in, (size_t)n * arity, "reef_poseidon_hash");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* base;
  const size_t in_bytes = (size_t)n * arity * 32, out_bytes = (size_t)n * 32;
  rc = ctx_scratch(c, in_bytes + out_bytes, &base);
  if (rc) return rc;
  char* d_in = (char*)base;
  char* d_out = d_in + in_bytes;
  REEF_CUDA(cudaMemcpyAsync(d_in, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
  rc = launch_hash_batch(c, d_in, (int)arity, n, d_out);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

int reef_calc_d(reef_ctx* c, const uint8_t v[32], const uint8_t salt[32], uint8_t out[32]) {
  REEF_REQUIRE(c && v && salt && out, REEF_EINVAL, "reef_calc_d: NULL argument");
  uint8_t in[64];
  memcpy(in, v, 32);
  memcpy(in + 32, salt, 32);
  return reef_poseidon_hash(c, in, 2, 1, out);
}

int reef_poseidon_sponge(reef_ctx* c, const uint8_t* in, uint32_t n_in, const uint32_t* ops, uint32_t n_ops,
                         uint32_t domain_separator, uint8_t* out, uint32_t n_out) {
  REEF_REQUIRE(c && ops && n_ops > 0, REEF_EINVAL, "reef_poseidon_sponge: NULL/empty pattern");
  uint64_t tot_in = 0, tot_out = 0;
  for (uint32_t k = 0; k < n_ops; k++) {
    if (ops[k] >> 31) tot_in += ops[k] & 0x7fffffffu;
    else tot_out += ops[k];
  }
  // neptune asserts that the calls match the declared pattern
  if (tot_in != n_in || tot_out != n_out) return fail(REEF_EASSERT, "reef_poseidon_sponge: IOPattern does not match the supplied element counts");
  REEF_REQUIRE((n_in == 0 || in) && (n_out == 0 || out), REEF_EINVAL, "reef_poseidon_sponge: NULL buffer");
  int rc = check_canonical(in, n_in, "reef_poseidon_sponge");
  if (rc) return rc;
  uint8_t tag[32];
  io_pattern_tag_le32(ops, n_ops, domain_separator, tag);
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* base;
  const size_t ops_bytes = ((size_t)n_ops * 4 + 31) & ~(size_t)31;
  rc = ctx_scratch(c, ops_bytes + (size_t)(n_in + n_out + 1) * 32, &base);
  if (rc) return rc;
  char* d_ops = (char*)base;
  char* d_in = d_ops + ops_bytes;
  char* d_out = d_in + (size_t)n_in * 32;
  REEF_CUDA(cudaMemcpyAsync(d_ops, ops, (size_t)n_ops * 4, cudaMemcpyHostToDevice, c->stream));
  if (n_in) REEF_CUDA(cudaMemcpyAsync(d_in, in, (size_t)n_in * 32, cudaMemcpyHostToDevice, c->stream));
  rc = launch_sponge_run(c, (const uint32_t*)d_ops, n_ops, d_in, tag, d_out);
  if (rc) return rc;
  if (n_out) REEF_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_out * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

int reef_sponge_start(reef_ctx* c, const uint32_t* ops, uint32_t n_ops, uint32_t domain_separator, reef_sponge** out) {
  REEF_REQUIRE(c && ops && n_ops > 0 && out, REEF_EINVAL, "reef_sponge_start: NULL/empty argument");
  uint8_t tag[32];
  io_pattern_tag_le32(ops, n_ops, domain_separator, tag);
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  reef_sponge* sp = new reef_sponge;
  sp->ctx = c;
  sp->ops.assign(ops, ops + n_ops);
  sp->io = 0;
  sp->d_state = nullptr;
  sp->d_buf = nullptr;
  sp->buf_bytes = 0;
  cudaError_t e = cudaMalloc(&sp->d_state, (size_t)sponge_state_bytes());
  if (e != cudaSuccess) {
    delete sp;
    return fail(REEF_ENOMEM, std::string("reef_sponge_start: ") + cudaGetErrorString(e));
  }
  int rc = launch_sponge_step(c, sp->d_state, 0, nullptr, 0, tag, nullptr);
  if (rc) {
    cudaFree(sp->d_state);
    delete sp;
    return rc;
  }
  ctx_retain(c);
  *out = sp;
  return REEF_OK;
}

static int sponge_buf(reef_sponge* sp, size_t bytes) {
  if (sp->buf_bytes >= bytes) return REEF_OK;
  cudaStreamSynchronize(sp->ctx->stream);
  if (sp->d_buf) cudaFree(sp->d_buf);
  sp->d_buf = nullptr;
  sp->buf_bytes = 0;
  cudaError_t e = cudaMalloc(&sp->d_buf, bytes + 1024);
  if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("reef_sponge: ") + cudaGetErrorString(e));
  sp->buf_bytes = bytes + 1024;
  return REEF_OK;
}

int reef_sponge_absorb(reef_sponge* sp, const uint8_t* elems, uint32_t n) {
  REEF_REQUIRE(sp && (elems || n == 0), REEF_EINVAL, "reef_sponge_absorb: NULL argument");
  if (sp->io >= sp->ops.size() || sp->ops[sp->io] != ((1u << 31) | n))
    return fail(REEF_EASSERT, "reef_sponge_absorb: call does not match the declared IOPattern");
  int rc = check_canonical(elems, n, "reef_sponge_absorb");
  if (rc) return rc;
  reef_ctx* c = sp->ctx;
  REEF_CTX_LIVE(c, "reef_sponge_absorb");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  rc = sponge_buf(sp, (size_t)n * 32);
  if (rc) return rc;
  if (n) REEF_CUDA(cudaMemcpyAsync(sp->d_buf, elems, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
  rc = launch_sponge_step(c, sp->d_state, 1, sp->d_buf, n, nullptr, nullptr);
  if (rc) return rc;
  REEF_CUDA(cudaStreamSynchronize(c->stream));  // caller may reuse `elems`
  sp->io++;
  return REEF_OK;
}

int reef_sponge_squeeze(reef_sponge* sp, uint32_t n, uint8_t* out) {
  REEF_REQUIRE(sp && (out || n == 0), REEF_EINVAL, "reef_sponge_squeeze: NULL argument");
  if (sp->io >= sp->ops.size() || sp->ops[sp->io] != n)
    return fail(REEF_EASSERT, "reef_sponge_squeeze: call does not match the declared IOPattern");
  reef_ctx* c = sp->ctx;
  REEF_CTX_LIVE(c, "reef_sponge_squeeze");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  int rc = sponge_buf(sp, (size_t)n * 32);
  if (rc) return rc;
  rc = launch_sponge_step(c, sp->d_state, 2, nullptr, n, nullptr, sp->d_buf);
  if (rc) return rc;
  if (n) REEF_CUDA(cudaMemcpyAsync(out, sp->d_buf, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  sp->io++;
  return REEF_OK;
}

int reef_sponge_finish(reef_sponge* sp) {
  REEF_REQUIRE(sp != nullptr, REEF_EINVAL, "reef_sponge_finish: NULL argument");
  reef_ctx* c = sp->ctx;
  bool ok = sp->io == sp->ops.size();
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (sp->d_state) cudaFree(sp->d_state);
    if (sp->d_buf) cudaFree(sp->d_buf);
  }
  delete sp;
  ctx_release(c);
  if (!ok) return fail(REEF_EASSERT, "reef_sponge_finish: ParameterUsageMismatch");
  return REEF_OK;
}

uint64_t reef_merkle_tree_elems(uint64_t n_doc) {
  uint64_t total = 0, n = n_doc;
  if (n == 0) return 0;
  n = (n + 1) / 2;
  total += n;
  while (n > 1) {
    n = (n + 1) / 2;
    total += n;
  }
  return total;
}

int reef_merkle_build_dev(reef_ctx* c, const uint64_t* doc_dev, uint64_t n_doc, void* levels_dev, uint64_t* level_sizes,
                          uint32_t* n_levels, uint8_t out_root[32]) {
  REEF_REQUIRE(c && doc_dev && levels_dev && out_root, REEF_EINVAL, "reef_merkle_build_dev: NULL argument");
  REEF_REQUIRE(n_doc >= 1, REEF_EASSERT, "reef_merkle_build: empty document (index out of bounds)");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  uint64_t sizes[64];
  uint32_t nl = 0;
  int rc = launch_merkle(c, doc_dev, n_doc, levels_dev, sizes, &nl);
  if (rc) return rc;
  const uint64_t total = reef_merkle_tree_elems(n_doc);
  void* hs;
  rc = ctx_stage(c, 32, &hs);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(hs, (const char*)levels_dev + (total - 1) * 32, 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(out_root, hs, 32);
  if (level_sizes) memcpy(level_sizes, sizes, nl * sizeof(uint64_t));
  if (n_levels) *n_levels = nl;
  return REEF_OK;
}

int reef_merkle_build(reef_ctx* c, const uint64_t* doc, uint64_t n_doc, uint8_t* out_levels, uint64_t* level_sizes,
                      uint32_t* n_levels, uint8_t out_root[32]) {
  REEF_REQUIRE(c && doc && out_levels && out_root, REEF_EINVAL, "reef_merkle_build: NULL argument");
  REEF_REQUIRE(n_doc >= 1, REEF_EASSERT, "reef_merkle_build: empty document (index out of bounds)");
  const uint64_t total = reef_merkle_tree_elems(n_doc);
  void* base;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    REEF_CUDA(cudaSetDevice(c->device));
    const size_t doc_bytes = ((size_t)n_doc * 8 + 255) & ~(size_t)255;
    int rc = ctx_scratch2(c, doc_bytes + (size_t)total * 32, &base);
    if (rc) return rc;
    REEF_CUDA(cudaMemcpyAsync(base, doc, (size_t)n_doc * 8, cudaMemcpyHostToDevice, c->stream));
  }
  const size_t doc_bytes = ((size_t)n_doc * 8 + 255) & ~(size_t)255;
  char* d_levels = (char*)base + doc_bytes;
  int rc = reef_merkle_build_dev(c, (const uint64_t*)base, n_doc, d_levels, level_sizes, n_levels, out_root);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaMemcpyAsync(out_levels, d_levels, (size_t)total * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

int reef_merkle_subtree(reef_ctx* c, const uint64_t* doc_local, uint64_t n_local, uint64_t idx_offset, uint8_t* out_levels,
                        uint8_t out_root[32]) {
  REEF_REQUIRE(c && doc_local && out_root, REEF_EINVAL, "reef_merkle_subtree: NULL argument");
  REEF_REQUIRE(n_local >= 2 && (n_local & (n_local - 1)) == 0, REEF_EINVAL, "reef_merkle_subtree: n_local must be a power of two >= 2");
  REEF_REQUIRE(idx_offset % n_local == 0, REEF_EINVAL, "reef_merkle_subtree: idx_offset must be a multiple of n_local");
  const uint64_t total = n_local - 1;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* base;
  const size_t doc_bytes = ((size_t)n_local * 8 + 255) & ~(size_t)255;
  int rc = ctx_scratch2(c, doc_bytes + (size_t)total * 32, &base);
  if (rc) return rc;
  char* d_levels = (char*)base + doc_bytes;
  REEF_CUDA(cudaMemcpyAsync(base, doc_local, (size_t)n_local * 8, cudaMemcpyHostToDevice, c->stream));
  rc = launch_merkle(c, (const uint64_t*)base, n_local, d_levels, nullptr, nullptr, idx_offset);
  if (rc) return rc;
  if (out_levels) REEF_CUDA(cudaMemcpyAsync(out_levels, d_levels, (size_t)total * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaMemcpyAsync(out_root, d_levels + (total - 1) * 32, 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

int reef_merkle_top(reef_ctx* c, const uint8_t* roots, uint32_t g, uint8_t* out_levels, uint8_t out_root[32]) {
  REEF_REQUIRE(c && roots && out_root, REEF_EINVAL, "reef_merkle_top: NULL argument");
  REEF_REQUIRE(g >= 1 && (g & (g - 1)) == 0, REEF_EINVAL, "reef_merkle_top: the number of subtrees must be a power of two");
  int rc = check_canonical(roots, g, "reef_merkle_top");
  if (rc) return rc;
  if (g == 1) {
    memcpy(out_root, roots, 32);
    return REEF_OK;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* base;
  rc = ctx_scratch2(c, (size_t)(2 * g) * 32, &base);
  if (rc) return rc;
  char* d_top = (char*)base + (size_t)g * 32;
  REEF_CUDA(cudaMemcpyAsync(base, roots, (size_t)g * 32, cudaMemcpyHostToDevice, c->stream));
  rc = launch_merkle_inner(c, base, g, d_top, nullptr, nullptr);
  if (rc) return rc;
  if (out_levels) REEF_CUDA(cudaMemcpyAsync(out_levels, d_top, (size_t)(g - 1) * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaMemcpyAsync(out_root, d_top + (size_t)(g - 2) * 32, 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

int reef_merkle_path_wits(const uint64_t* doc, uint64_t n_doc, const uint8_t* levels, const uint64_t* level_sizes,
                          uint32_t n_levels, uint64_t idx, uint8_t* l_or_r, uint8_t* has_idx, uint64_t* opposite_idx,
                          uint8_t* opposite) {
  REEF_REQUIRE(doc && levels && level_sizes && l_or_r && has_idx && opposite_idx && opposite, REEF_EINVAL,
               "reef_merkle_path_wits: NULL argument");
  if (idx >= n_doc) return fail(REEF_EASSERT, "assertion failed: idx < self.doc.len()");
  auto put_u64 = [](uint8_t* o, uint64_t x) {
    memset(o, 0, 32);
    for (int i = 0; i < 8; i++) o[i] = (uint8_t)(x >> (8 * i));
  };
  // leaf entry (merkle_tree.rs:132-157)
  has_idx[0] = 1;
  if (idx % 2 == 0) {
    l_or_r[0] = 1;
    if (idx + 1 >= n_doc) {
      opposite_idx[0] = 0;
      put_u64(opposite, 0);
    } else {
      opposite_idx[0] = idx + 1;
      put_u64(opposite, doc[idx + 1]);
    }
  } else {
    l_or_r[0] = 0;
    opposite_idx[0] = idx - 1;
    put_u64(opposite, doc[idx - 1]);
  }
  // inner entries (merkle_tree.rs:159-188)
  uint64_t quo = idx / 2;
  const uint8_t* lvl = levels;
  for (uint32_t h = 0; h + 1 < n_levels; h++) {
    uint8_t* o = opposite + (size_t)(h + 1) * 32;
    has_idx[h + 1] = 0;
    opposite_idx[h + 1] = 0;
    if (quo % 2 == 0) {
      l_or_r[h + 1] = 1;
      if (quo + 1 >= level_sizes[h]) memset(o, 0, 32);
      else memcpy(o, lvl + (size_t)(quo + 1) * 32, 32);
    } else {
      l_or_r[h + 1] = 0;
      memcpy(o, lvl + (size_t)(quo - 1) * 32, 32);
    }
    quo /= 2;
    lvl += (size_t)level_sizes[h] * 32;
  }
  return REEF_OK;
}

// ---------------------------------------------------------------------------------------
// B1: tables + nlookup
// ---------------------------------------------------------------------------------------
static uint64_t next_pow2(uint64_t n) {
  uint64_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

static int table_new(reef_ctx* c, const void* host, uint64_t n, int is_u32, reef_table** out, bool async_upload = false) {
  REEF_REQUIRE(c && host && out, REEF_EINVAL, "reef_table_upload: NULL argument");
  REEF_REQUIRE(n >= 1, REEF_EASSERT, "reef_table_upload: empty table (index out of bounds: table[0])");
  const size_t esz = is_u32 ? 4 : 32;
  if (!is_u32) {
    int rc = check_canonical((const uint8_t*)host, n, "reef_table_upload");
    if (rc) return rc;
  }
  uint64_t n_pad = next_pow2(n);
  if (n_pad < 2) n_pad = 2;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* d = nullptr;
  cudaError_t e = cudaSuccess;
  for (size_t k = 0; k < c->table_cache.size(); k++) {
    if (c->table_cache[k].first == (size_t)n_pad * esz) {
      d = c->table_cache[k].second;
      c->table_cache.erase(c->table_cache.begin() + k);
      break;
    }
  }
  if (!d) e = cudaMalloc(&d, (size_t)n_pad * esz);
  if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("reef_table_upload: ") + cudaGetErrorString(e));
  cudaEvent_t ready = nullptr;
  if (async_upload) {
    // upload on the context's copy stream, ordered behind whatever the compute stream still has queued (the buffer may
    // come from the cache of freed tables); consumers wait on `ready`, the caller returns at once
    if (!c->copy_stream) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    cudaEvent_t before = nullptr;
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&before, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(before, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, before, 0);
    if (before) cudaEventDestroy(before);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, host, (size_t)n * esz, cudaMemcpyHostToDevice, c->copy_stream);
    if (e == cudaSuccess && n_pad > n) e = cudaMemsetAsync((char*)d + (size_t)n * esz, 0, (size_t)(n_pad - n) * esz, c->copy_stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ready, c->copy_stream);
  } else {
    e = cudaMemcpyAsync(d, host, (size_t)n * esz, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_pad > n) e = cudaMemsetAsync((char*)d + (size_t)n * esz, 0, (size_t)(n_pad - n) * esz, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  }
  if (e != cudaSuccess) {
    if (ready) cudaEventDestroy(ready);
    cudaFree(d);
    return fail(REEF_ECUDA, std::string("reef_table_upload: ") + cudaGetErrorString(e));
  }
  reef_table* t = new reef_table;
  t->ready = ready;
  t->ctx = c;
  t->d = d;
  t->n_pad = n_pad;
  t->n_orig = n;
  t->is_u32 = is_u32;
  t->owns = 1;
  memset(t->first, 0, 32);
  memcpy(t->first, host, is_u32 ? 4 : 32);
  ctx_retain(c);
  *out = t;
  return REEF_OK;
}

int reef_table_upload(reef_ctx* c, const uint8_t* table, uint64_t n, reef_table** out) { return table_new(c, table, n, 0, out); }
int reef_table_upload_u32(reef_ctx* c, const uint32_t* codes, uint64_t n, reef_table** out) { return table_new(c, codes, n, 1, out); }
int reef_table_upload_u32_async(reef_ctx* c, const uint32_t* codes, uint64_t n, reef_table** out) { return table_new(c, codes, n, 1, out, true); }

int reef_table_hybrid_u32(reef_ctx* c, const uint8_t* pub_table, uint64_t n_pub, const uint8_t fill[32], uint64_t half_len,
                          const uint32_t* doc_codes, uint64_t n_doc, reef_table** out) {
  REEF_REQUIRE(c && pub_table && fill && doc_codes && out, REEF_EINVAL, "reef_table_hybrid_u32: NULL argument");
  REEF_REQUIRE(n_pub >= 1 && n_doc >= 1, REEF_EASSERT, "reef_table_hybrid_u32: empty table (index out of bounds)");
  REEF_REQUIRE(half_len >= 2 && (half_len & (half_len - 1)) == 0 && half_len >= n_pub, REEF_EINVAL,
               "reef_table_hybrid_u32: half_len must be a power of two >= the public table length");
  REEF_REQUIRE(next_pow2(n_doc) <= half_len, REEF_EASSERT, "reef_table_hybrid_u32: padded document longer than half_len");
  int rc = check_canonical(pub_table, n_pub, "reef_table_hybrid_u32: pub_table");
  if (!rc) rc = check_canonical(fill, 1, "reef_table_hybrid_u32: fill");
  if (rc) return rc;
  const uint64_t n = 2 * half_len;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  void* d = nullptr;
  for (size_t k = 0; k < c->table_cache.size(); k++) {
    if (c->table_cache[k].first == (size_t)n * 32) {
      d = c->table_cache[k].second;
      c->table_cache.erase(c->table_cache.begin() + k);
      break;
    }
  }
  if (!d) {
    cudaError_t e = cudaMalloc(&d, (size_t)n * 32);
    if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("reef_table_hybrid_u32: ") + cudaGetErrorString(e));
  }
  void* stage;
  const size_t pub_bytes = ((size_t)n_pub * 32 + 255) & ~(size_t)255;
  rc = ctx_scratch2(c, pub_bytes + (size_t)n_doc * 4, &stage);
  if (rc) {
    cudaFree(d);
    return rc;
  }
  uint32_t* d_codes = (uint32_t*)((char*)stage + pub_bytes);
  cudaError_t e = cudaMemcpyAsync(stage, pub_table, (size_t)n_pub * 32, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_codes, doc_codes, (size_t)n_doc * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    rc = launch_hybrid_table(c, stage, n_pub, fill, half_len, d_codes, n_doc, d);
    if (!rc) e = cudaStreamSynchronize(c->stream);
  }
  if (rc || e != cudaSuccess) {
    cudaFree(d);
    return rc ? rc : fail(REEF_ECUDA, std::string("reef_table_hybrid_u32: ") + cudaGetErrorString(e));
  }
  reef_table* t = new reef_table;
  t->ctx = c;
  t->d = d;
  t->n_pad = t->n_orig = n;
  t->is_u32 = 0;
  t->owns = 1;
  memcpy(t->first, pub_table, 32);
  ctx_retain(c);
  *out = t;
  return REEF_OK;
}

int reef_table_wrap_dev(reef_ctx* c, void* dev_ptr, uint64_t n, int is_u32, reef_table** out) {
  REEF_REQUIRE(c && dev_ptr && out, REEF_EINVAL, "reef_table_wrap_dev: NULL argument");
  REEF_REQUIRE(n >= 2 && (n & (n - 1)) == 0, REEF_EINVAL, "reef_table_wrap_dev: length must be a power of two >= 2");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  reef_table* t = new reef_table;
  t->ctx = c;
  t->d = dev_ptr;
  t->n_pad = t->n_orig = n;
  t->is_u32 = is_u32 ? 1 : 0;
  t->owns = 0;
  memset(t->first, 0, 32);
  cudaError_t e = cudaMemcpy(t->first, dev_ptr, is_u32 ? 4 : 32, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    delete t;
    return fail(REEF_ECUDA, std::string("reef_table_wrap_dev: ") + cudaGetErrorString(e));
  }
  ctx_retain(c);
  *out = t;
  return REEF_OK;
}

int reef_table_download(const reef_table* t, uint8_t* out, uint64_t n) {
  REEF_REQUIRE(t && out, REEF_EINVAL, "reef_table_download: NULL argument");
  REEF_REQUIRE(!t->is_u32, REEF_EINVAL, "reef_table_download: u32 tables are not downloadable as field elements");
  REEF_REQUIRE(n <= t->n_pad, REEF_EINVAL, "reef_table_download: n exceeds the table length");
  reef_ctx* c = t->ctx;
  REEF_CTX_LIVE(c, "reef_table_download");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  table_wait(c, t);
  REEF_CUDA(cudaMemcpyAsync(out, t->d, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

uint64_t reef_table_len(const reef_table* t) { return t ? t->n_orig : 0; }

void reef_table_free(reef_table* t) {
  if (!t) return;
  reef_ctx* c = t->ctx;
  if (t->owns) {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    if (t->ready) {
      cudaEventSynchronize(t->ready);
      cudaEventDestroy(t->ready);
    }
    cudaStreamSynchronize(c->stream);
    if (!c->closed.load()) {
      // keep the most recently freed buffers (a prover re-uploads same-sized tables proof after proof); when the
      // cache is full the OLDEST entry goes, so a new working set takes the cache over after one round
      if (c->table_cache.size() >= 4) {
        cudaFree(c->table_cache.front().second);
        c->table_cache.erase(c->table_cache.begin());
      }
      c->table_cache.emplace_back((size_t)t->n_pad * (t->is_u32 ? 4 : 32), t->d);
    } else {
      cudaFree(t->d);
    }
  }
  delete t;
  ctx_release(c);
}

// Host half of wit_nlookup_gadget shared by the single-GPU and the sharded entry points:
// defaults, combined_q, IOPattern, first absorb (r1cs.rs:2187-2306).
struct NlPrep {
  uint32_t ell;
  std::vector<uint8_t> prevq, cqs, query;
  std::vector<uint32_t> ops;
  uint8_t prevv[32];
  uint32_t num_cqs, first;
};

static int nl_prepare(int tag, uint64_t n_orig, uint64_t n_pad, const uint8_t* first_elem, const uint64_t* q, const uint8_t* v,
                      uint32_t m, const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash, reef_nlookup_out* out,
                      NlPrep& P) {
  REEF_REQUIRE(tag == REEF_TAG_NL || tag == REEF_TAG_NLDOC || tag == REEF_TAG_NLHYBRID, REEF_EASSERT, "weird tag");
  REEF_REQUIRE(m == 0 || (q && v), REEF_EINVAL, "reef_nlookup_prove: NULL q/v");
  REEF_REQUIRE(tag == REEF_TAG_NL || doc_hash, REEF_EINVAL, "reef_nlookup_prove: doc_hash required for nldoc/nlhybrid");
  REEF_REQUIRE(out->claim_r && out->rounds && out->sc_last_claim && out->next_running_claim, REEF_EINVAL,
               "reef_nlookup_prove: NULL output buffer");
  // sc_l = logmn(table.len())  (r1cs.rs:2187)
  const uint32_t ell = reef_logmn(n_orig);
  const uint64_t n_full = ell < 63 ? ((uint64_t)1 << ell) : 0;
  if (tag == REEF_TAG_NLDOC) {
    // zero-padded to 2^logmn(len) (r1cs.rs:2322-2329)
    if (n_full < n_orig) return fail(REEF_EASSERT, "attempt to subtract with overflow (f32 logmn)");
  } else {
    // linear_mle_product asserts table_t.len() == 2^ell (r1cs_helper.rs:450)
    if (n_full != n_orig) return fail(REEF_EASSERT, "assertion failed: table_t.len() == base.pow(ell)");
  }
  REEF_REQUIRE(n_full == n_pad, REEF_EINVAL, "reef_nlookup_prove: padded table length does not match 2^logmn(len)");
  REEF_REQUIRE(out->rounds_cap >= ell, REEF_EINVAL, "reef_nlookup_prove: rounds buffer too small");
  for (uint32_t k = 0; k < m; k++)
    REEF_REQUIRE(q[k] < n_pad, REEF_EASSERT, "nlookup: lookup index out of range (index out of bounds)");
  out->ell = ell;
  P.ell = ell;
  int rc;
  if (m) {
    rc = check_canonical(v, m, "reef_nlookup_prove: v");
    if (rc) return rc;
  }
  P.prevq.assign((size_t)ell * 32, 0);
  if (prev_q) {
    rc = check_canonical(prev_q, ell, "reef_nlookup_prove: prev_q");
    if (rc) return rc;
    memcpy(P.prevq.data(), prev_q, (size_t)ell * 32);
  }
  if (prev_v) {
    rc = check_canonical(prev_v, 1, "reef_nlookup_prove: prev_v");
    if (rc) return rc;
    memcpy(P.prevv, prev_v, 32);
  } else {
    REEF_REQUIRE(first_elem != nullptr, REEF_EINVAL, "reef_nlookup_prove: prev_v required (table[0] is not on this rank)");
    memcpy(P.prevv, first_elem, 32);
  }
  if (out->prev_running_claim) memcpy(out->prev_running_claim, P.prevv, 32);
  // combined_q (r1cs.rs:2208-2249)
  P.num_cqs = 0;
  P.cqs.assign((size_t)((uint64_t)m * ell / 254 + 2) * 32, 0);
  rc = reef_combined_q(q, m, ell, P.cqs.data(), (uint32_t)(P.cqs.size() / 32), &P.num_cqs);
  if (rc) return rc;
  out->num_cqs = P.num_cqs;
  if (out->combined_q) {
    REEF_REQUIRE(out->combined_q_cap >= P.num_cqs, REEF_EINVAL, "reef_nlookup_prove: combined_q buffer too small");
    memcpy(out->combined_q, P.cqs.data(), (size_t)P.num_cqs * 32);
  }
  // IOPattern (r1cs.rs:2263-2282) and first absorb (r1cs.rs:2285-2306)
  const bool with_hash = tag != REEF_TAG_NL;
  P.first = m + ell + 1 + P.num_cqs + (with_hash ? 1 : 0);
  P.ops.clear();
  P.ops.push_back((1u << 31) | P.first);
  P.ops.push_back(1);
  for (uint32_t i = 0; i < ell; i++) {
    P.ops.push_back((1u << 31) | 3u);
    P.ops.push_back(1);
  }
  P.query.clear();
  P.query.reserve((size_t)P.first * 32);
  if (with_hash) {
    rc = check_canonical(doc_hash, 1, "reef_nlookup_prove: doc_hash");
    if (rc) return rc;
    P.query.insert(P.query.end(), doc_hash, doc_hash + 32);
  }
  P.query.insert(P.query.end(), P.cqs.begin(), P.cqs.begin() + (size_t)P.num_cqs * 32);
  if (m) P.query.insert(P.query.end(), v, v + (size_t)m * 32);
  P.query.insert(P.query.end(), P.prevq.begin(), P.prevq.end());
  P.query.insert(P.query.end(), P.prevv, P.prevv + 32);
  return REEF_OK;
}

struct reef_witness {
  reef_ctx* ctx;
  void* d;
  uint64_t n;
};

static int nlookup_prove_impl(reef_ctx* c, int tag, const reef_table* table, const uint64_t* q, const uint8_t* v, uint32_t m,
                              const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash, reef_nlookup_out* out,
                              reef_witness* w, const reef_nlookup_slots* slots) {
  REEF_REQUIRE(c && table && out, REEF_EINVAL, "reef_nlookup_prove: NULL argument");
  REEF_REQUIRE(table->ctx == c, REEF_EINVAL, "reef_nlookup_prove: table belongs to another context");
  NlPrep P;
  int rc = nl_prepare(tag, table->n_orig, table->n_pad, table->first, q, v, m, prev_q, prev_v, doc_hash, out, P);
  if (rc) return rc;
  NlookupArgs a;
  a.tag = tag;
  a.d_table = table->d;
  a.table_is_u32 = table->is_u32;
  a.n = table->n_pad;
  a.ell = P.ell;
  a.m = m;
  a.h_q = q;
  a.h_query = P.query.data();
  a.n_query = P.first;
  a.h_prev_q = P.prevq.data();
  io_pattern_tag_le32(P.ops.data(), (uint32_t)P.ops.size(), 0, a.tag_le);
  a.out_claim_r = out->claim_r;
  a.out_rounds = out->rounds;
  a.out_last_claim = out->sc_last_claim;
  a.out_next_v = out->next_running_claim;
  a.table_ready = table->ready;            // pending asynchronous upload: awaited right before the first kernel that reads the table
  if (w) {
    REEF_REQUIRE(slots, REEF_EINVAL, "reef_nlookup_prove_w: NULL slots");
    REEF_REQUIRE(w->ctx == c, REEF_EINVAL, "reef_nlookup_prove_w: witness buffer belongs to another context");
    const uint64_t none = ~0ull;
    auto fits = [&](uint64_t s, uint64_t k) { return s == none || (s < w->n && k <= w->n - s); };
    REEF_REQUIRE(fits(slots->claim_r, 1) && fits(slots->rounds, 4ull * P.ell) && fits(slots->last_claim, 1) && fits(slots->next_claim, 1),
                 REEF_EASSERT, "reef_nlookup_prove_w: witness slot out of range (index out of bounds)");
    a.d_wit = w->d;
    a.wit_len = w->n;
    a.slot_claim_r = slots->claim_r;
    a.slot_rounds = slots->rounds;
    a.slot_last_claim = slots->last_claim;
    a.slot_next_v = slots->next_claim;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return nlookup_run(c, a);
}

int reef_nlookup_prove(reef_ctx* c, int tag, const reef_table* table, const uint64_t* q, const uint8_t* v, uint32_t m,
                       const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash, reef_nlookup_out* out) {
  return nlookup_prove_impl(c, tag, table, q, v, m, prev_q, prev_v, doc_hash, out, nullptr, nullptr);
}

int reef_nlookup_prove_w(reef_ctx* c, int tag, const reef_table* table, const uint64_t* q, const uint8_t* v, uint32_t m,
                         const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash, reef_nlookup_out* out,
                         reef_witness* w, const reef_nlookup_slots* slots) {
  REEF_REQUIRE(w && slots, REEF_EINVAL, "reef_nlookup_prove_w: NULL witness / slots");
  return nlookup_prove_impl(c, tag, table, q, v, m, prev_q, prev_v, doc_hash, out, w, slots);
}

// ---- (f1) index-addressed witness buffer
int reef_witness_create(reef_ctx* c, uint64_t n, reef_witness** out) {
  REEF_REQUIRE(c && out && n >= 1, REEF_EINVAL, "reef_witness_create: NULL / empty argument");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_witness_create");
  REEF_CUDA(cudaSetDevice(c->device));
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, (size_t)n * 32);
  if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("reef_witness_create: ") + cudaGetErrorString(e));
  REEF_CUDA(cudaMemsetAsync(d, 0, (size_t)n * 32, c->stream));
  reef_witness* w = new reef_witness{c, d, n};
  ctx_retain(c);
  *out = w;
  return REEF_OK;
}

static int witness_set(reef_witness* w, const uint64_t* idx, const void* vals, uint64_t k, bool small) {
  REEF_REQUIRE(w && (k == 0 || (idx && vals)), REEF_EINVAL, "reef_witness_set: NULL argument");
  if (k == 0) return REEF_OK;
  for (uint64_t i = 0; i < k; i++) REEF_REQUIRE(idx[i] < w->n, REEF_EASSERT, "reef_witness_set: index out of bounds");
  if (!small) {
    int rc = check_canonical((const uint8_t*)vals, k, "reef_witness_set");
    if (rc) return rc;
  }
  reef_ctx* c = w->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_witness_set");
  REEF_CUDA(cudaSetDevice(c->device));
  // runs of consecutive indices become one copy each (the layout a circuit hands out is mostly contiguous)
  void* hs;
  int rc = ctx_stage(c, (size_t)k * 32, &hs);
  if (rc) return rc;
  uint8_t* h = (uint8_t*)hs;
  for (uint64_t i = 0; i < k; i++) {
    if (small) {
      memset(h + 32 * i, 0, 32);
      const uint64_t x = ((const uint64_t*)vals)[i];
      for (int b = 0; b < 8; b++) h[32 * i + b] = (uint8_t)(x >> (8 * b));
    } else {
      memcpy(h + 32 * i, (const uint8_t*)vals + 32 * i, 32);
    }
  }
  uint64_t i = 0;
  while (i < k) {
    uint64_t j = i + 1;
    while (j < k && idx[j] == idx[j - 1] + 1) j++;
    REEF_CUDA(cudaMemcpyAsync((char*)w->d + 32 * idx[i], h + 32 * i, (size_t)(j - i) * 32, cudaMemcpyHostToDevice, c->stream));
    i = j;
  }
  REEF_CUDA(cudaStreamSynchronize(c->stream));   // the staging buffer is reused by the next call
  return REEF_OK;
}

int reef_witness_set(reef_witness* w, const uint64_t* idx, const uint8_t* vals, uint64_t k) { return witness_set(w, idx, vals, k, false); }
int reef_witness_set_u64(reef_witness* w, const uint64_t* idx, const uint64_t* vals, uint64_t k) { return witness_set(w, idx, vals, k, true); }

int reef_witness_read(reef_witness* w, uint64_t first, uint64_t k, uint8_t* out) {
  REEF_REQUIRE(w && out, REEF_EINVAL, "reef_witness_read: NULL argument");
  REEF_REQUIRE(first <= w->n && k <= w->n - first, REEF_EASSERT, "reef_witness_read: range out of bounds");
  reef_ctx* c = w->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_witness_read");
  REEF_CUDA(cudaSetDevice(c->device));
  REEF_CUDA(cudaMemcpyAsync(out, (char*)w->d + 32 * first, (size_t)k * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

void* reef_witness_dev(reef_witness* w) { return w ? w->d : nullptr; }
uint64_t reef_witness_len(reef_witness* w) { return w ? w->n : 0; }

void reef_witness_free(reef_witness* w) {
  if (!w) return;
  reef_ctx* c = w->ctx;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(w->d);
  }
  delete w;
  ctx_release(c);
}

// ---- multi-GPU sharded sum-check: see include/reef_b200.h
int reef_nl_shard_begin(reef_ctx* c, int tag, const reef_table* local_table, uint32_t rank, uint32_t world, const uint64_t* q,
                        const uint8_t* v, uint32_t m, const uint8_t* prev_q, const uint8_t* prev_v, const uint8_t* doc_hash,
                        reef_nlookup_out* out, reef_nl_session** session) {
  REEF_REQUIRE(c && local_table && out && session, REEF_EINVAL, "reef_nl_shard_begin: NULL argument");
  REEF_REQUIRE(local_table->ctx == c, REEF_EINVAL, "reef_nl_shard_begin: table belongs to another context");
  REEF_REQUIRE(world >= 1 && (world & (world - 1)) == 0 && rank < world, REEF_EINVAL, "reef_nl_shard_begin: world must be a power of two");
  REEF_REQUIRE(local_table->n_orig == local_table->n_pad, REEF_EINVAL, "reef_nl_shard_begin: local shard must be a power of two long");
  const uint64_t n_global = local_table->n_pad * world;
  NlPrep P;
  int rc = nl_prepare(tag, n_global, n_global, (world == 1 || rank == 0) ? local_table->first : nullptr, q, v, m, prev_q, prev_v,
                      doc_hash, out, P);
  if (rc) return rc;
  NlookupArgs a;
  a.tag = tag;
  a.d_table = local_table->d;
  a.table_is_u32 = local_table->is_u32;
  a.n = local_table->n_pad;
  a.ell = P.ell;
  a.m = m;
  a.h_q = q;
  a.h_query = P.query.data();
  a.n_query = P.first;
  a.h_prev_q = P.prevq.data();
  io_pattern_tag_le32(P.ops.data(), (uint32_t)P.ops.size(), 0, a.tag_le);
  a.out_claim_r = a.out_rounds = a.out_last_claim = a.out_next_v = nullptr;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  table_wait(c, local_table);
  return nl_shard_begin(c, a, rank, world, session);
}

// Every session call serialises on the session's context like the other entry points (the sharded
// path mutates context state: mailbox sequence, staging buffers, profile records).
#define REEF_SHARD_ENTER(s, what)                                 \
  reef_ctx* c = nl_shard_ctx(s);                                  \
  REEF_CTX_LIVE(c, what);                                         \
  std::lock_guard<std::mutex> lk(c->mu);                          \
  REEF_CUDA(cudaSetDevice(c->device))

int reef_nl_shard_round_local(reef_nl_session* s, void* out_triple_dev) {
  REEF_REQUIRE(s && out_triple_dev, REEF_EINVAL, "reef_nl_shard_round_local: NULL argument");
  REEF_SHARD_ENTER(s, "reef_nl_shard_round_local");
  return nl_shard_round_local(s, out_triple_dev);
}

int reef_nl_shard_round_finish(reef_nl_session* s, const void* all_triples_dev) {
  REEF_REQUIRE(s && all_triples_dev, REEF_EINVAL, "reef_nl_shard_round_finish: NULL argument");
  REEF_SHARD_ENTER(s, "reef_nl_shard_round_finish");
  return nl_shard_round_finish(s, all_triples_dev);
}

int reef_nl_shard_export(reef_nl_session* s, void* out_pair_dev) {
  REEF_REQUIRE(s && out_pair_dev, REEF_EINVAL, "reef_nl_shard_export: NULL argument");
  REEF_SHARD_ENTER(s, "reef_nl_shard_export");
  return nl_shard_export(s, out_pair_dev);
}

int reef_nl_shard_finish(reef_nl_session* s, const void* all_pairs_dev, reef_nlookup_out* out) {
  REEF_REQUIRE(s && all_pairs_dev && out, REEF_EINVAL, "reef_nl_shard_finish: NULL argument");
  REEF_REQUIRE(out->claim_r && out->rounds && out->sc_last_claim && out->next_running_claim, REEF_EINVAL,
               "reef_nl_shard_finish: NULL output buffer");
  REEF_SHARD_ENTER(s, "reef_nl_shard_finish");
  return nl_shard_finish(s, all_pairs_dev, out->claim_r, out->rounds, out->sc_last_claim, out->next_running_claim);
}

int reef_nl_shard_round_p2p(reef_nl_session* s) {
  REEF_REQUIRE(s, REEF_EINVAL, "reef_nl_shard_round_p2p: NULL argument");
  REEF_SHARD_ENTER(s, "reef_nl_shard_round_p2p");
  return nl_shard_round_p2p(s);
}

int reef_nl_shard_finish_p2p(reef_nl_session* s, reef_nlookup_out* out) {
  REEF_REQUIRE(s && out, REEF_EINVAL, "reef_nl_shard_finish_p2p: NULL argument");
  REEF_REQUIRE(out->claim_r && out->rounds && out->sc_last_claim && out->next_running_claim, REEF_EINVAL,
               "reef_nl_shard_finish_p2p: NULL output buffer");
  REEF_SHARD_ENTER(s, "reef_nl_shard_finish_p2p");
  return nl_shard_finish_p2p(s, out->claim_r, out->rounds, out->sc_last_claim, out->next_running_claim);
}

void reef_nl_shard_free(reef_nl_session* s) {
  if (!s) return;
  reef_ctx* c = nl_shard_ctx(s);
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    nl_shard_free(s);
  }
  ctx_release(c);
}

int reef_gen_eq_table(reef_ctx* c, const uint8_t* rs, const uint64_t* qs, uint32_t m, const uint8_t* last_q, uint32_t ell,
                      uint8_t* out) {
  REEF_REQUIRE(c && rs && last_q && out && (qs || m == 0), REEF_EINVAL, "reef_gen_eq_table: NULL argument");
  REEF_REQUIRE(ell >= 1 && ell <= 30, REEF_EINVAL, "reef_gen_eq_table: ell out of range for a host output buffer");
  int rc = check_canonical(rs, m + 1, "reef_gen_eq_table: rs");
  if (rc) return rc;
  rc = check_canonical(last_q, ell, "reef_gen_eq_table: last_q");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  const uint64_t n = (uint64_t)1 << ell;
  void* d_out;
  rc = ctx_scratch2(c, (size_t)n * 32, &d_out);
  if (rc) return rc;
  rc = launch_gen_eq_table(c, rs, qs, m, last_q, ell, d_out);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

int reef_linear_mle_product(reef_ctx* c, reef_table* tt, reef_table* te, uint32_t ell, uint32_t i, reef_sponge* sp,
                            uint8_t out[128]) {
  REEF_REQUIRE(c && tt && te && sp && out, REEF_EINVAL, "reef_linear_mle_product: NULL argument");
  REEF_REQUIRE(!tt->is_u32 && !te->is_u32, REEF_EINVAL, "reef_linear_mle_product: tables must hold field elements");
  REEF_REQUIRE(ell >= 1 && ell < 63, REEF_EINVAL, "reef_linear_mle_product: ell out of range");
  const uint64_t n = (uint64_t)1 << ell;
  if (tt->n_pad != n || te->n_pad != n) return fail(REEF_EASSERT, "assertion failed: table.len() == base.pow(ell)");
  uint8_t g[96];  // (xsq, x, con)
  {
    std::lock_guard<std::mutex> lk(c->mu);
    REEF_CUDA(cudaSetDevice(c->device));
    table_wait(c, tt);
    table_wait(c, te);
    int rc = launch_mle_round_coeffs(c, tt->d, te->d, ell, i, g);
    if (rc) return rc;
  }
  uint8_t query[96];  // absorb [con, x, xsq]  (r1cs_helper.rs:479-486)
  memcpy(query, g + 64, 32);
  memcpy(query + 32, g + 32, 32);
  memcpy(query + 64, g, 32);
  int rc = reef_sponge_absorb(sp, query, 3);
  if (rc) return rc;
  uint8_t r[32];
  rc = reef_sponge_squeeze(sp, 1, r);
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    rc = launch_mle_round_fold(c, tt->d, te->d, ell, i, r);
    if (rc) return rc;
    REEF_CUDA(cudaStreamSynchronize(c->stream));
  }
  memcpy(out, r, 32);
  memcpy(out + 32, g, 96);
  return REEF_OK;
}

int reef_verifier_mle_eval(reef_ctx* c, const reef_table* t, const uint8_t* q, uint32_t ell, uint8_t out[32]) {
  REEF_REQUIRE(c && t && q && out, REEF_EINVAL, "reef_verifier_mle_eval: NULL argument");
  // prover_mle_partial_eval asserts 2^(m-1) <= prods.len() <= 2^m  (r1cs_helper.rs:562-563)
  if (ell < 1 || ell > 62 || ((uint64_t)1 << (ell - 1)) > t->n_orig || ((uint64_t)1 << ell) < t->n_orig)
    return fail(REEF_EASSERT, "assertion failed: base.pow(m - 1) <= prods.len() <= base.pow(m)");
  int rc = check_canonical(q, ell, "reef_verifier_mle_eval: q");
  if (rc) return rc;
  REEF_REQUIRE(t->n_pad == ((uint64_t)1 << ell), REEF_EINVAL, "reef_verifier_mle_eval: padded length mismatch");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  table_wait(c, t);
  return launch_mle_eval(c, t->d, t->is_u32, t->n_pad, q, ell, out);
}

int reef_hyrax_lz(reef_ctx* c, const reef_table* t, uint64_t rows, uint64_t cols, const uint8_t* L, uint8_t* out) {
  REEF_REQUIRE(c && t && L && out, REEF_EINVAL, "reef_hyrax_lz: NULL argument");
  REEF_REQUIRE(t->ctx == c, REEF_EINVAL, "reef_hyrax_lz: table belongs to another context");
  REEF_REQUIRE(rows >= 1 && cols >= 1 && rows * cols == t->n_pad, REEF_EASSERT, "reef_hyrax_lz: rows * cols must equal the table length");
  int rc = check_canonical(L, rows, "reef_hyrax_lz: L");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  table_wait(c, t);
  return launch_lz(c, t->d, t->is_u32, rows, cols, L, out);
}

int reef_prover_mle_partial_eval(reef_ctx* c, const reef_table* t, const uint8_t* x, uint32_t ell, int32_t hole,
                                 uint8_t out_coeff[32], uint8_t out_const[32]) {
  REEF_REQUIRE(c && t && x && out_coeff && out_const, REEF_EINVAL, "reef_prover_mle_partial_eval: NULL argument");
  REEF_REQUIRE(hole >= -1 && hole < (int32_t)ell, REEF_EINVAL, "reef_prover_mle_partial_eval: hole index out of range");
  if (hole < 0) {
    // (crap, full result): the reference returns hole_coeff = -minus_coeff here
    int rc = reef_verifier_mle_eval(c, t, x, ell, out_const);
    if (rc) return rc;
    // hole_coeff = 0 - minus_coeff  (mod q)
    uint32_t borrow = 0;
    bool zero = true;
    for (int i = 0; i < 32; i++) zero = zero && out_const[i] == 0;
    if (zero) {
      memset(out_coeff, 0, 32);
    } else {
      for (int i = 0; i < 32; i++) {
        int d = (int)FQ_LE[i] - (int)out_const[i] - (int)borrow;
        borrow = d < 0;
        out_coeff[i] = (uint8_t)(d + (borrow ? 256 : 0));
      }
    }
    return REEF_OK;
  }
  std::vector<uint8_t> xx(x, x + (size_t)ell * 32);
  uint8_t c0[32], c1[32];
  memset(&xx[(size_t)hole * 32], 0, 32);
  int rc = reef_verifier_mle_eval(c, t, xx.data(), ell, c0);
  if (rc) return rc;
  xx[(size_t)hole * 32] = 1;
  rc = reef_verifier_mle_eval(c, t, xx.data(), ell, c1);
  if (rc) return rc;
  // coeff = c1 - c0 (mod q)
  int borrow = 0;
  uint8_t d[32];
  for (int i = 0; i < 32; i++) {
    int v = (int)c1[i] - (int)c0[i] - borrow;
    borrow = v < 0;
    d[i] = (uint8_t)(v + (borrow ? 256 : 0));
  }
  if (borrow) {
    int carry = 0;
    for (int i = 0; i < 32; i++) {
      int v = (int)d[i] + (int)FQ_LE[i] + carry;
      d[i] = (uint8_t)v;
      carry = v >> 8;
    }
  }
  memcpy(out_coeff, d, 32);
  memcpy(out_const, c0, 32);
  return REEF_OK;
}

// ---------------------------------------------------------------------------------------
// test hooks (include/reef_b200_testing.h): host evaluation of the SHARED __host__ __device__
// field / permutation code so its structure is unit-tested without a GPU.  Not a product path.
// ---------------------------------------------------------------------------------------
int reef_hosttest_field_op(int field, int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  if (field == 0) hosttest_field_op<FqCfg>(op, a, b, out);
  else hosttest_field_op<FpCfg>(op, a, b, out);
  return 0;
}

int reef_hosttest_mul_wide(const uint8_t a[32], const uint8_t b[32], uint8_t out[64]) {
  uint32_t x[8], y[8], r[16];
  load_le(x, a);
  load_le(y, b);
  mul_wide(r, x, y);
  store_le(out, r);
  store_le(out + 32, r + 8);
  return 0;
}

int reef_hosttest_poseidon_permute(const uint8_t in[160], uint8_t out[160]) {
  Fq s[5];
  for (int i = 0; i < 5; i++) {
    load_le(s[i].v, in + 32 * i);
    s[i] = to_mont<FqCfg>(s[i]);
  }
  poseidon_permute_host(s);
  for (int i = 0; i < 5; i++) {
    Fq o = from_mont<FqCfg>(s[i]);
    store_le(out + 32 * i, o.v);
  }
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// B3: MSM
// ---------------------------------------------------------------------------------------
#include "ec.cuh"


template <class C>
static void hosttest_ec(int op, const uint8_t* p, const uint8_t* q, uint8_t* out) {
  Affine<C> a, b;
  load_le(a.x.v, p);
  load_le(a.y.v, p + 32);
  load_le(b.x.v, q);
  load_le(b.y.v, q + 32);
  bool binf = affine_is_inf<C>(b);
  a.x = to_mont<C>(a.x);
  a.y = to_mont<C>(a.y);
  uint64_t k = 0;
  for (int i = 0; i < 8; i++) k |= (uint64_t)q[i] << (8 * i);
  if (op != 4) {
    b.x = to_mont<C>(b.x);
    b.y = to_mont<C>(b.y);
  }
  (void)binf;
  XYZZ<C> r = xyzz_from_affine<C>(a);
  switch (op) {
    case 0: {
      XYZZ<C> t = xyzz_from_affine<C>(b);
      // randomise the representation of t: scale by (zz, zzz) = (4, 8)  i.e. Z = 2
      if (!xyzz_is_inf<C>(t)) {
        Fe<C> four = fe_from_u64<C>(4), eight = fe_from_u64<C>(8);
        t.x = mont_mul<C>(t.x, four);
        t.y = mont_mul<C>(t.y, eight);
        t.zz = four;
        t.zzz = eight;
      }
      xyzz_add<C>(r, t);
      break;
    }
    case 1: xyzz_add_affine<C>(r, b, false); break;
    case 2: r = xyzz_dbl<C>(r); break;
    case 3: xyzz_add_affine<C>(r, b, true); break;
    case 4: r = xyzz_mul_small<C>(r, k); break;
  }
  Affine<C> o = xyzz_to_affine<C>(r);
  o.x = from_mont<C>(o.x);
  o.y = from_mont<C>(o.y);
  store_le(out, o.x.v);
  store_le(out + 32, o.y.v);
}

extern "C" {

int reef_hosttest_ec_op(int curve, int op, const uint8_t p[64], const uint8_t q[64], uint8_t out[64]) {
  if (curve == 0) hosttest_ec<FpCfg>(op, p, q, out);
  else hosttest_ec<FqCfg>(op, p, q, out);
  return 0;
}

// (f2) Registered generator levels are a pure function of (device, curve, scalar width, the points): a process-wide,
// reference-counted cache keyed by a 128-bit content hash lets every context / proof that commits with the same
// `CommitmentGens` (the prover registers the same key on several streams, the verifier derives it again,
// framework.rs:297-303, 770, 910-976) share ONE set of window levels in HBM and skip k_precompute.  REEF_BASES_CACHE=0
// turns it off.
namespace {
struct LevelsEntry {
  int device, curve;
  uint64_t n;
  uint32_t scalar_bits;
  uint64_t h1, h2;
  void* d_levels;
  int refs;
};
std::mutex g_levels_mu;
std::vector<LevelsEntry> g_levels;
std::atomic<unsigned long long> g_levels_hits{0};
bool levels_cache_on() {
  static const bool on = !(getenv("REEF_BASES_CACHE") && atoi(getenv("REEF_BASES_CACHE")) == 0);
  return on;
}
void hash128(const uint8_t* p, size_t bytes, uint64_t& h1, uint64_t& h2) {
  uint64_t a = 0xcbf29ce484222325ull, b = 0x9e3779b97f4a7c15ull;
  const size_t words = bytes / 8;
  for (size_t i = 0; i < words; i++) {
    uint64_t w;
    memcpy(&w, p + 8 * i, 8);
    a = (a ^ w) * 0x100000001b3ull;
    b = (b + w) * 0xff51afd7ed558ccdull;
    b ^= b >> 29;
  }
  h1 = a;
  h2 = b;
}
}  // namespace

int reef_bases_register(reef_ctx* c, int curve, const uint8_t* bases, uint64_t n, uint32_t scalar_bits, reef_bases** out) {
  REEF_REQUIRE(c && bases && out, REEF_EINVAL, "reef_bases_register: NULL argument");
  REEF_REQUIRE(curve == REEF_CURVE_PALLAS || curve == REEF_CURVE_VESTA, REEF_EINVAL, "reef_bases_register: unknown curve");
  REEF_REQUIRE(n >= 1, REEF_EINVAL, "reef_bases_register: no bases");
  if (scalar_bits == 0 || scalar_bits > 255) scalar_bits = 255;
  uint64_t h1 = 0, h2 = 0;
  const bool cache = levels_cache_on();
  if (cache) hash128(bases, (size_t)n * 64, h1, h2);
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_bases_register");
  REEF_CUDA(cudaSetDevice(c->device));
  MsmPlanPublic pl = msm_make_plan(n, scalar_bits, (uint64_t)24 << 30);
  void* d_levels = nullptr;
  if (cache) {
    std::lock_guard<std::mutex> gl(g_levels_mu);
    for (auto& e : g_levels)
      if (e.device == c->device && e.curve == curve && e.n == n && e.scalar_bits == scalar_bits && e.h1 == h1 && e.h2 == h2) {
        e.refs++;
        d_levels = e.d_levels;
        g_levels_hits.fetch_add(1, std::memory_order_relaxed);
        break;
      }
  }
  if (!d_levels) {
    int rc = msm_bases_register(c, curve, bases, n, pl, &d_levels);
    if (rc) return rc;
    if (cache) {
      std::lock_guard<std::mutex> gl(g_levels_mu);
      g_levels.push_back(LevelsEntry{c->device, curve, n, scalar_bits, h1, h2, d_levels, 1});
    }
  }
  reef_bases* b = new reef_bases;
  b->ctx = c;
  b->curve = curve;
  b->n = n;
  b->scalar_bits = scalar_bits;
  b->plan = pl;
  b->d_levels = d_levels;
  ctx_retain(c);
  *out = b;
  return REEF_OK;
}

void reef_bases_free(reef_bases* b) {
  if (!b) return;
  reef_ctx* c = b->ctx;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    bool shared_alive = false, found = false;
    {
      std::lock_guard<std::mutex> gl(g_levels_mu);
      for (size_t i = 0; i < g_levels.size(); i++)
        if (g_levels[i].d_levels == b->d_levels) {
          found = true;
          if (--g_levels[i].refs > 0) shared_alive = true;
          else g_levels.erase(g_levels.begin() + i);
          break;
        }
    }
    (void)found;
    if (!shared_alive) cudaFree(b->d_levels);
  }
  delete b;
  ctx_release(c);
}

/* hits since process start / live entries of the generator-level cache */
int reef_bases_cache_stats(uint64_t* hits, uint64_t* entries) {
  std::lock_guard<std::mutex> gl(g_levels_mu);
  if (hits) *hits = g_levels_hits.load();
  if (entries) *entries = g_levels.size();
  return REEF_OK;
}

uint32_t reef_bases_windows(const reef_bases* b) { return b ? b->plan.W : 0; }
uint32_t reef_bases_window_bits(const reef_bases* b) { return b ? b->plan.c : 0; }

static int msm_dispatch(reef_ctx* c, const reef_bases* b, const void* d_scalars, int is_u32, uint64_t n, uint32_t w0,
                        uint32_t w1, uint8_t* out_aff, uint8_t* out_xyzz) {
  MsmRunArgs a;
  a.plan = b->plan;
  a.d_levels = b->d_levels;
  a.n_bases = b->n;
  a.d_scalars = d_scalars;
  a.scalars_u32 = is_u32;
  a.n = n;
  a.w_begin = w0;
  a.w_end = w1;
  a.h_out_affine = out_aff;
  a.h_out_xyzz = out_xyzz;
  a.h_extra_xyzz_mont = nullptr;
  a.n_extra = 0;
  return msm_run(c, b->curve, a);
}

extern "C" int reef_msm_sharded_dev(reef_ctx* c, const reef_bases* b, const void* scalars_dev, uint64_t n, uint8_t out[64]) {
  REEF_REQUIRE(c && b && out && scalars_dev, REEF_EINVAL, "reef_msm_sharded_dev: NULL argument");
  REEF_REQUIRE(b->ctx == c, REEF_EINVAL, "reef_msm_sharded_dev: bases belong to another context");
  REEF_REQUIRE(n >= 1 && n <= b->n, REEF_EASSERT, "reef_msm_sharded_dev: scalar count out of range");
  REEF_REQUIRE(b->scalar_bits == 255, REEF_EINVAL, "reef_msm_sharded_dev: full-width scalars need generators registered with scalar_bits = 255");
  REEF_REQUIRE(c->mb_world >= 1, REEF_EINVAL, "reef_msm_sharded_dev: mailbox not connected (reef_mailbox_connect)");
  REEF_REQUIRE(b->plan.W >= c->mb_world, REEF_EINVAL, "reef_msm_sharded_dev: fewer Pippenger windows than ranks");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  const uint32_t W = b->plan.W, G = c->mb_world, g = c->mb_rank;
  MsmRunArgs a;
  a.plan = b->plan;
  a.d_levels = b->d_levels;
  a.n_bases = b->n;
  a.d_scalars = scalars_dev;
  a.scalars_u32 = 0;
  a.n = n;
  a.w_begin = W * g / G;
  a.w_end = W * (g + 1) / G;
  a.h_out_affine = out;
  a.h_out_xyzz = nullptr;
  a.h_extra_xyzz_mont = nullptr;
  a.n_extra = 0;
  a.p2p_combine = 1;
  int rc = msm_run(c, b->curve, a);
  if (rc) return rc;
  uint32_t e = 0;
  REEF_CUDA(cudaMemcpy(&e, c->mb_err_dev, 4, cudaMemcpyDeviceToHost));
  if (e) {
    cudaMemset(c->mb_err_dev, 0, 4);
    return fail(REEF_ECUDA, "reef_msm_sharded_dev: exchange " + std::to_string(e & 0x7fffffffu) + " failed (a peer never posted, or a peer reported a timeout)");
  }
  return REEF_OK;
}

static int msm_host_scalars(reef_ctx* c, const reef_bases* b, const void* scalars, int is_u32, uint64_t n, uint8_t out[64]) {
  REEF_REQUIRE(c && b && out && (scalars || n == 0), REEF_EINVAL, "reef_msm: NULL argument");
  REEF_REQUIRE(b->ctx == c, REEF_EINVAL, "reef_msm: bases belong to another context");
  REEF_REQUIRE(n <= b->n, REEF_EASSERT, "reef_msm: more scalars than generators (assertion failed: gens.len() >= v.len())");
  if (n == 0) {
    memset(out, 0, 64);
    return REEF_OK;
  }
  if (is_u32) REEF_REQUIRE(b->scalar_bits >= 32, REEF_EINVAL, "reef_msm_u32: bases were registered for narrower scalars");
  else REEF_REQUIRE(b->scalar_bits == 255, REEF_EINVAL, "reef_msm: full-width scalars need generators registered with scalar_bits = 255 (higher bits would be dropped)");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * (is_u32 ? 4 : 32);
  void* d_s;
  int rc = ctx_scratch2(c, bytes, &d_s);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(d_s, scalars, bytes, cudaMemcpyHostToDevice, c->stream));
  return msm_dispatch(c, b, d_s, is_u32, n, 0, b->plan.W, out, nullptr);
}

int reef_msm(reef_ctx* c, const reef_bases* b, const uint8_t* scalars, uint64_t n, uint8_t out[64]) {
  return msm_host_scalars(c, b, scalars, 0, n, out);
}

int reef_msm_u32(reef_ctx* c, const reef_bases* b, const uint32_t* scalars, uint64_t n, uint8_t out[64]) {
  return msm_host_scalars(c, b, scalars, 1, n, out);
}

int reef_msm_dev(reef_ctx* c, const reef_bases* b, const void* scalars_dev, uint64_t n, uint8_t out[64]) {
  REEF_REQUIRE(c && b && out && scalars_dev, REEF_EINVAL, "reef_msm_dev: NULL argument");
  REEF_REQUIRE(b->ctx == c, REEF_EINVAL, "reef_msm_dev: bases belong to another context");
  REEF_REQUIRE(n >= 1 && n <= b->n, REEF_EASSERT, "reef_msm_dev: scalar count out of range");
  REEF_REQUIRE(b->scalar_bits == 255, REEF_EINVAL, "reef_msm_dev: full-width scalars need generators registered with scalar_bits = 255");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return msm_dispatch(c, b, scalars_dev, 0, n, 0, b->plan.W, out, nullptr);
}

int reef_msm_partial_dev(reef_ctx* c, const reef_bases* b, const void* scalars_dev, uint64_t n, uint32_t w_begin,
                         uint32_t w_end, uint8_t out_xyzz[128]) {
  REEF_REQUIRE(c && b && out_xyzz && scalars_dev, REEF_EINVAL, "reef_msm_partial_dev: NULL argument");
  REEF_REQUIRE(b->ctx == c, REEF_EINVAL, "reef_msm_partial_dev: bases belong to another context");
  REEF_REQUIRE(n >= 1 && n <= b->n, REEF_EASSERT, "reef_msm_partial_dev: scalar count out of range");
  REEF_REQUIRE(b->scalar_bits == 255, REEF_EINVAL, "reef_msm_partial_dev: full-width scalars need generators registered with scalar_bits = 255");
  REEF_REQUIRE(w_begin <= w_end && w_end <= b->plan.W, REEF_EINVAL, "reef_msm_partial_dev: window range out of bounds");
  if (w_begin == w_end) {
    memset(out_xyzz, 0, 128);
    return REEF_OK;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return msm_dispatch(c, b, scalars_dev, 0, n, w_begin, w_end, nullptr, out_xyzz);
}

// caller holds c->mu; d_rows_out (optional) receives the device address of the rows x 64 B affine results, valid
// until the next call that uses the context's scratch
static int msm_rows_host_locked(reef_ctx* c, const reef_bases* b, const void* matrix, int is_u32, uint64_t rows, uint64_t cols,
                                uint32_t entry_bits, const uint8_t* blinds, uint8_t* out, void** d_rows_out) {
  REEF_REQUIRE(b->ctx == c, REEF_EINVAL, "reef_msm_rows: bases belong to another context");
  REEF_REQUIRE(rows >= 1 && cols >= 1, REEF_EINVAL, "reef_msm_rows: empty matrix");
  REEF_REQUIRE(cols + (blinds ? 1 : 0) <= b->n, REEF_EASSERT, "reef_msm_rows: not enough generators (assertion failed: gens.len() >= cols)");
  if (entry_bits == 0 || entry_bits > 255) entry_bits = 255;
  REEF_REQUIRE(entry_bits <= b->scalar_bits, REEF_EINVAL, "reef_msm_rows: generators registered for narrower scalars");
  REEF_REQUIRE(!blinds || b->scalar_bits == 255, REEF_EINVAL, "reef_msm_rows: blinds need generators registered with scalar_bits = 255");
  REEF_CUDA(cudaSetDevice(c->device));
  const size_t mbytes = (size_t)rows * cols * (is_u32 ? 4 : 32);
  const size_t mpad = (mbytes + 255) & ~(size_t)255;
  void* d_s;
  int rc = ctx_scratch2(c, mpad + (blinds ? rows * 32 : 0), &d_s);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(d_s, matrix, mbytes, cudaMemcpyHostToDevice, c->stream));
  void* d_b = nullptr;
  if (blinds) {
    d_b = (char*)d_s + mpad;
    REEF_CUDA(cudaMemcpyAsync(d_b, blinds, (size_t)rows * 32, cudaMemcpyHostToDevice, c->stream));
  }
  MsmRowsArgs a;
  a.plan = b->plan;
  a.d_levels = b->d_levels;
  a.n_bases = b->n;
  a.d_scalars = d_s;
  a.scalars_u32 = is_u32;
  a.scalar_bits = entry_bits;
  a.rows = rows;
  a.cols = cols;
  a.d_blinds = d_b;
  a.blind_base = cols;
  a.h_out = out;
  a.d_rows_out = d_rows_out;
  return msm_rows_run(c, b->curve, a);
}

static int msm_rows_host(reef_ctx* c, const reef_bases* b, const void* matrix, int is_u32, uint64_t rows, uint64_t cols,
                         uint32_t entry_bits, const uint8_t* blinds, uint8_t* out) {
  REEF_REQUIRE(c && b && matrix && out, REEF_EINVAL, "reef_msm_rows: NULL argument");
  std::lock_guard<std::mutex> lk(c->mu);
  return msm_rows_host_locked(c, b, matrix, is_u32, rows, cols, entry_bits, blinds, out, nullptr);
}

int reef_msm_rows_u32(reef_ctx* c, const reef_bases* b, const uint32_t* matrix, uint64_t rows, uint64_t cols,
                      uint32_t entry_bits, const uint8_t* blinds, uint8_t* out) {
  if (entry_bits == 0 || entry_bits > 32) entry_bits = 32;
  return msm_rows_host(c, b, matrix, 1, rows, cols, entry_bits, blinds, out);
}

int reef_msm_rows(reef_ctx* c, const reef_bases* b, const uint8_t* matrix, uint64_t rows, uint64_t cols,
                  const uint8_t* blinds, uint8_t* out) {
  return msm_rows_host(c, b, matrix, 0, rows, cols, 255, blinds, out);
}

// scalars resident: rows x cols canonical 32-byte values, row-major (e.g. the witness W and the cross term T of one
// fold as two rows over the same commitment key: ONE sort / accumulate / reduce chain for both commitments)
int reef_msm_rows_dev(reef_ctx* c, const reef_bases* b, const void* matrix_dev, uint64_t rows, uint64_t cols, uint8_t* out) {
  REEF_REQUIRE(c && b && matrix_dev && out, REEF_EINVAL, "reef_msm_rows_dev: NULL argument");
  REEF_REQUIRE(b->ctx == c, REEF_EINVAL, "reef_msm_rows_dev: bases belong to another context");
  REEF_REQUIRE(rows >= 1 && cols >= 1, REEF_EINVAL, "reef_msm_rows_dev: empty matrix");
  REEF_REQUIRE(cols <= b->n, REEF_EASSERT, "reef_msm_rows_dev: not enough generators (assertion failed: gens.len() >= cols)");
  REEF_REQUIRE(b->scalar_bits == 255, REEF_EINVAL, "reef_msm_rows_dev: generators registered for narrower scalars");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_msm_rows_dev");
  REEF_CUDA(cudaSetDevice(c->device));
  MsmRowsArgs a;
  a.plan = b->plan;
  a.d_levels = b->d_levels;
  a.n_bases = b->n;
  a.d_scalars = matrix_dev;
  a.scalars_u32 = 0;
  a.scalar_bits = 255;
  a.rows = rows;
  a.cols = cols;
  a.d_blinds = nullptr;
  a.blind_base = cols;
  a.h_out = out;
  return msm_rows_run(c, b->curve, a);
}

int reef_msm_combine(reef_ctx* c, int curve, const uint8_t* partials, uint32_t k, uint8_t out[64]) {
  REEF_REQUIRE(c && partials && out, REEF_EINVAL, "reef_msm_combine: NULL argument");
  REEF_REQUIRE(curve == REEF_CURVE_PALLAS || curve == REEF_CURVE_VESTA, REEF_EINVAL, "reef_msm_combine: unknown curve");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return msm_combine(c, curve, partials, k, out);
}

// ---------------------------------------------------------------------------------------
// a16: PoseidonRO (nova-snark), doc_commit_hash
// ---------------------------------------------------------------------------------------
static const uint8_t FP_LE[32] = {0x01, 0x00, 0x00, 0x00, 0xed, 0x30, 0x2d, 0x99, 0x1b, 0xf9, 0x4c, 0x09, 0xfc, 0x98, 0x46, 0x22,
                                  0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x00, 0x00, 0x00, 0x40};

static bool lt_le32(const uint8_t* a, const uint8_t* m) {
  for (int i = 31; i >= 0; i--) {
    if (a[i] < m[i]) return true;
    if (a[i] > m[i]) return false;
  }
  return false;
}

// digest (canonical in the base field) -> low num_bits bits -> element of the other field
static void ro_finish(uint8_t h[32], int base_field, uint32_t num_bits) {
  if (num_bits < 256)
    for (uint32_t bit = num_bits; bit < 256; bit++) h[bit >> 3] &= (uint8_t)~(1u << (bit & 7));
  const uint8_t* m = base_field == 0 ? FP_LE : FQ_LE;   // modulus of the OTHER field
  if (!lt_le32(h, m)) {                                 // h < 2^255 < 2 m: one subtraction
    int borrow = 0;
    for (int i = 0; i < 32; i++) {
      int d = (int)h[i] - (int)m[i] - borrow;
      borrow = d < 0;
      h[i] = (uint8_t)(d + (borrow << 8));
    }
  }
}

static int ro_run_host_input(reef_ctx* c, int base_field, const uint8_t* data, size_t bytes, uint64_t n_elems, int triples,
                             uint32_t num_bits, uint8_t out[32], const char* what) {
  REEF_REQUIRE(n_elems >= 1 && (n_elems >> 31) == 0, REEF_EINVAL, std::string(what) + ": element count out of range");
  REEF_REQUIRE(num_bits >= 1 && num_bits <= 256, REEF_EINVAL, std::string(what) + ": num_bits out of range");
  const uint8_t* bm = base_field == 0 ? FQ_LE : FP_LE;
  for (size_t i = 0; i < bytes / 32; i++)
    if (!lt_le32(data + 32 * i, bm)) return fail(REEF_EINVAL, std::string(what) + ": element " + std::to_string(i) + " is not canonical");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_poseidon_ro");
  REEF_CUDA(cudaSetDevice(c->device));
  void* base;
  int rc = ctx_scratch(c, bytes + 32, &base);
  if (rc) return rc;
  char* d_in = (char*)base + 32;
  REEF_CUDA(cudaMemcpyAsync(d_in, data, bytes, cudaMemcpyHostToDevice, c->stream));
  rc = launch_poseidon_ro(c, base_field, d_in, n_elems, triples, base);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(out, base, 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  ro_finish(out, base_field, num_bits);
  return REEF_OK;
}

int reef_poseidon_ro(reef_ctx* c, int base_field, const uint8_t* elems, uint64_t n, uint32_t num_bits, uint8_t out[32]) {
  REEF_REQUIRE(c && elems && out, REEF_EINVAL, "reef_poseidon_ro: NULL argument");
  REEF_REQUIRE(base_field == 0 || base_field == 1, REEF_EINVAL, "reef_poseidon_ro: unknown field");
  return ro_run_host_input(c, base_field, elems, (size_t)n * 32, n, 0, num_bits, out, "reef_poseidon_ro");
}

int reef_poseidon_ro_points(reef_ctx* c, int curve, const uint8_t* points, uint64_t n_points, uint32_t num_bits, uint8_t out[32]) {
  REEF_REQUIRE(c && points && out, REEF_EINVAL, "reef_poseidon_ro_points: NULL argument");
  REEF_REQUIRE(curve == REEF_CURVE_PALLAS || curve == REEF_CURVE_VESTA, REEF_EINVAL, "reef_poseidon_ro_points: unknown curve");
  const int base_field = curve == REEF_CURVE_PALLAS ? 1 : 0;     // coordinates of Pallas points live in Fp
  return ro_run_host_input(c, base_field, points, (size_t)n_points * 64, 3 * n_points, 1, num_bits, out, "reef_poseidon_ro_points");
}

int reef_doc_commit_u32(reef_ctx* c, const reef_bases* gens, const uint32_t* doc, uint64_t rows, uint64_t cols, uint32_t entry_bits,
                        const uint8_t* blinds, uint8_t* out_rows, uint8_t out_hash[32]) {
  REEF_REQUIRE(c && gens && doc && out_rows && out_hash, REEF_EINVAL, "reef_doc_commit_u32: NULL argument");
  REEF_REQUIRE(gens->curve == REEF_CURVE_PALLAS, REEF_EINVAL, "reef_doc_commit_u32: the document commitment lives on Pallas (G1)");
  REEF_REQUIRE((3 * rows) >> 31 == 0, REEF_EINVAL, "reef_doc_commit_u32: too many rows");
  if (entry_bits == 0 || entry_bits > 32) entry_bits = 32;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_doc_commit_u32");
  void* d_rows = nullptr;
  int rc = msm_rows_host_locked(c, gens, doc, 1, rows, cols, entry_bits, blinds, out_rows, &d_rows);
  if (rc) return rc;
  // the row commitments are still in the context's scratch; the document codes in scratch2 are spent: digest goes there
  rc = launch_poseidon_ro(c, 1, d_rows, 3 * rows, 1, c->scratch2);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(out_hash, c->scratch2, 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  ro_finish(out_hash, 1, 256);
  return REEF_OK;
}

int reef_hosttest_poseidon_ro(int field, const uint8_t* elems, uint64_t n, uint8_t out[32]) {
  if (!elems || !out || n == 0 || (field != 0 && field != 1)) return REEF_EINVAL;
  poseidon_ro_host(field, elems, n, out);
  return REEF_OK;
}
int reef_hosttest_poseidon_ro_fast_ok(int field) {
  if (field != 0 && field != 1) return -1;
  return poseidon_ro_fast_ok_host(field);
}
int reef_hosttest_poseidon_ro_constants(int field, uint8_t* rc, uint8_t* mds) {
  if (!rc || !mds || (field != 0 && field != 1)) return REEF_EINVAL;
  poseidon_ro_constants_host(field, rc, mds);
  return REEF_OK;
}

}  // extern "C"
