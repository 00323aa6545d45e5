// Shared device/host helpers for libreef_b200: 256-bit vector loads, warp shuffles of
// field elements, error plumbing, context layout.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/reef_b200.h"
#include "fp.cuh"
#include "poseidon.cuh"

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

}  // namespace reef

struct reef_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;                         // one in-flight call per context
  reef::PoseidonTables* d_pos = nullptr; // Montgomery-form tables in global memory
  reef::SpongeTags tags;
  // reusable device scratch (grown on demand)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* scratch2 = nullptr;
  size_t scratch2_bytes = 0;
  // pinned host staging for small results
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  int sm_count = 148;
};

namespace reef {
int ctx_scratch(reef_ctx* c, size_t bytes, void** out);
int ctx_scratch2(reef_ctx* c, size_t bytes, void** out);
int ctx_stage(reef_ctx* c, size_t bytes, void** out);
}  // namespace reef
