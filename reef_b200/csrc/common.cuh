// Shared device/host helpers for libreef_b200: 256-bit vector loads, warp shuffles of
// field elements, error plumbing, context layout.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/reef_b200.h"
#include "fp.cuh"
#include "fp29.cuh"
#include "poseidon.cuh"
#include "poseidon_lp.cuh"

namespace reef {

// ---------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 on success; the message of the last
// failure on the calling thread is available through reef_last_error().
// ---------------------------------------------------------------------------------------
// status codes: REEF_OK / REEF_EINVAL / REEF_ECUDA / REEF_EASSERT / REEF_ENOMEM from the public header

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define REEF_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::reef::fail(REEF_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// after every kernel launch: count it (reef_launch_count) and surface launch errors
extern std::atomic<unsigned long long> g_launches;
#define REEF_LAUNCHED()                                   \
  do {                                                    \
    ::reef::g_launches.fetch_add(1, std::memory_order_relaxed); \
    REEF_CUDA(cudaGetLastError());                        \
  } while (0)

#define REEF_REQUIRE(cond, code, msg)                  \
  do {                                                 \
    if (!(cond)) return ::reef::fail((code), (msg));   \
  } while (0)

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#if defined(__CUDACC__)

// One LDG.E.ENL2.256 / STG.E.ENL2.256 per field element (sm_100+): a warp reading 32
// consecutive elements issues one fully-coalesced 1 KiB request.
template <class F>
__device__ __forceinline__ F ld256(const F* p) {
  F r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

// streaming (read-once) variant: no L1 allocation, read-only path
template <class F>
__device__ __forceinline__ F ld256_stream(const F* p) {
  F r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

template <class F>
__device__ __forceinline__ void st256(F* p, const F& r) {
  asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p)
               : "memory");
}

template <class F>
__device__ __forceinline__ F shfl_fe(const F& x, int src_lane) {
  F r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, x.v[i], src_lane);
  return r;
}

template <class F>
__device__ __forceinline__ F shfl_xor_fe(const F& x, int mask) {
  F r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, x.v[i], mask);
  return r;
}

template <class F>
__device__ __forceinline__ F select_fe(bool c, const F& a, const F& b) {
  F r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may become resident while its stream predecessor still runs; pdl_wait() blocks until that predecessor has
// completed and its writes are visible (a no-op for an ordinary launch), pdl_trigger() lets the stream successor
// start being scheduled.  Rule kept throughout: every kernel of such a chain calls pdl_wait() before its first
// access to memory another kernel of the chain writes or reads-then-overwrites.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// warp-wide sum (mod p) of one field element per lane; every lane gets the total
template <class C>
__device__ __forceinline__ Fe<C> warp_sum_fe(Fe<C> v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fe_add<C>(v, shfl_xor_fe(v, m));
  return v;
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct SpongeTags {
  Fq a2s1;  // IOPattern [Absorb(2), Squeeze(1)]   (Montgomery form)
  Fq a4s1;  // IOPattern [Absorb(4), Squeeze(1)]
};

enum ProfClass : int {
  PROF_SWEEP_FIRST = 0,   // k_sweep without fold (round 1)
  PROF_SWEEP_FOLD = 1,    // k_sweep fold + accumulate (rounds >= 2)
  PROF_ROUND = 2,         // k_round (transcript, A fold)
  PROF_TAIL = 3,          // k_tail
  PROF_NL_SETUP = 4,      // k_nl_begin + k_eq_tables
  PROF_MSM_SORT = 5,      // digits, histogram, scans, scatter
  PROF_MSM_ACCUM = 6,     // bucket accumulation passes
  PROF_MSM_REDUCE = 7,    // bucket gather + bit-decomposition sum + final
  PROF_POSEIDON = 8,      // hash batch / merkle levels / sponge
  PROF_NCLASS = 9
};
struct ProfRec {
  int cls;
  uint64_t units;         // class-specific work units (e.g. input length of a sweep)
  cudaEvent_t e0, e1;
};

}  // namespace reef

struct reef_ctx {
  // Lifetime: one reference for the caller's handle (dropped by reef_shutdown) plus one per live
  // child handle (table, bases, sponge, sum-check / sharded session).  The context is destroyed
  // when the last reference goes, so ANY destruction order of handles is safe; after
  // reef_shutdown the children can still be freed (and only freed).
  std::atomic<int> refs{1};
  std::atomic<bool> closed{false};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;    // asynchronous table uploads (reef_table_upload_u32_async), created on first use
  std::mutex mu;                         // one in-flight call per context
  reef::PoseidonTables* d_pos = nullptr; // Montgomery-form tables in global memory
  reef::PoseidonLpTables* d_lp = nullptr; // tables of the lane-parallel transcript permutation
  reef::SpongeTags tags;
  void* d_ro_fq = nullptr;               // PoseidonRO (width 25) tables, uploaded at first use (poseidon_ro.cu)
  void* d_ro_fp = nullptr;
  // reusable device scratch (grown on demand)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* scratch2 = nullptr;
  size_t scratch2_bytes = 0;
  // pinned host staging for small results
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  int sm_count = 148;
  // background contexts (reef_init_prio(.., 0, ..)): SMs of the green-context partition their stream lives in (0 = none)
  uint32_t partition_sms = 0;
  bool polite = false;                   // background context + REEF_MSM_POLITE=1: long MSM grids at two CTAs per SM (msm.cu)
  // device buffers of freed tables, reused by the next upload of the same size: a prover re-uploads
  // a same-sized table per proof, and cudaMalloc/cudaFree synchronise the whole device
  std::vector<std::pair<size_t, void*>> table_cache;
  // one retired buffer of a sharded sum-check session, reused by the next session (same reason)
  void* shard_cache = nullptr;
  size_t shard_cache_bytes = 0;
  // peer mailbox (p2p.cu): this rank's buffer, the peers' buffers (device array of pointers), and the
  // sequence number of the next exchange (every rank issues the same sequence of exchanges)
  void* mb_mine = nullptr;
  void** mb_peers_dev = nullptr;
  uint32_t* mb_err_dev = nullptr;
  std::vector<void*> mb_ipc_opened;
  uint32_t mb_world = 0, mb_rank = 0, mb_seq = 0;
  // optional per-kernel-class event timing (reef_profile_enable); resolved lazily
  bool profile = false;
  std::vector<reef::ProfRec> prof;
};

namespace reef {
void ctx_retain(reef_ctx* c);
void ctx_release(reef_ctx* c);   // never call with c->mu held
#define REEF_CTX_LIVE(c, what)                                                                         \
  do {                                                                                                 \
    if ((c)->closed.load()) return ::reef::fail(REEF_EINVAL, what ": the context was shut down");      \
  } while (0)
int ctx_scratch(reef_ctx* c, size_t bytes, void** out);
int ctx_scratch2(reef_ctx* c, size_t bytes, void** out);
int ctx_stage(reef_ctx* c, size_t bytes, void** out);
// RAII event pair around a group of launches (no-op unless profiling is enabled)
struct ProfScope {
  reef_ctx* c;
  reef::ProfRec r;
  bool on;
  ProfScope(reef_ctx* ctx, int cls, uint64_t units) : c(ctx), on(ctx->profile) {
    if (!on) return;
    r.cls = cls;
    r.units = units;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, c->stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.e1, c->stream);
    c->prof.push_back(r);
  }
};
}  // namespace reef
