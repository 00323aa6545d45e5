// Device side of the peer mailbox (see p2p.cu): post this rank's payload into every peer's
// mailbox over NVLink P2P stores, acquire the peers' payloads from this rank's own mailbox.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace reef {

static constexpr uint32_t MB_ENTRY = 256;
static constexpr uint32_t MB_SEQ_OFF = 248;   // payload <= 248 B (a partial MSM point in XYZZ coordinates is 128 B)
static constexpr uint32_t MB_MAX_WORLD = 32;
static constexpr uint32_t MB_SPIN_LIMIT = 1u << 25;   // ~30 s of polling (host-side skew between ranks is legal)

// Passed by value to kernels; peers == nullptr means "no mailbox exchange in this launch".
struct MbRef {
  void* const* peers;          // device array: peers[g] = rank g's mailbox (own entry: local pointer)
  const unsigned char* mine;   // this rank's mailbox
  uint32_t* err;               // device error flag (a peer never posted)
  uint32_t world, rank, seq;
};

#if defined(__CUDACC__)
// Failure protocol: a rank whose wait timed out sets its own error flag; from then on every post of
// that rank carries the POISON bit, a receiver that sees a poisoned post sets its own flag (and so
// poisons its later posts): within one more exchange every rank knows, and the *_finish_p2p /
// reef_p2p_status call of each rank returns an error instead of a silently wrong transcript.
static constexpr uint32_t MB_POISON = 0x80000000u;

// Thread `dest` (< world) stores nwords 32-bit words into rank `dest`'s mailbox and publishes them.
__device__ __forceinline__ void mb_post(const MbRef& mb, uint32_t dest, const uint32_t* payload, uint32_t nwords) {
  unsigned char* e = (unsigned char*)mb.peers[dest] + ((size_t)(mb.seq & 1u) * mb.world + mb.rank) * MB_ENTRY;
  volatile uint32_t* w = (volatile uint32_t*)e;
  for (uint32_t k = 0; k < nwords; k++) w[k] = payload[k];
  const uint32_t failed = *(volatile uint32_t*)mb.err;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(e + MB_SEQ_OFF), "r"(failed ? (mb.seq | MB_POISON) : mb.seq) : "memory");
}

// Thread `src` (< world) waits for rank `src`'s entry of this exchange and copies it to dst[0..nwords).
// On failure (time-out, or a poisoned post) dst is zero-filled, the error flag is set and false returned.
__device__ __forceinline__ bool mb_wait_copy(const MbRef& mb, uint32_t src, uint32_t* dst, uint32_t nwords) {
  const unsigned char* e = mb.mine + ((size_t)(mb.seq & 1u) * mb.world + src) * MB_ENTRY;
  // once this rank has failed, later waits give up quickly instead of burning the full limit each
  const uint32_t limit = *(volatile uint32_t*)mb.err ? (1u << 12) : MB_SPIN_LIMIT;
  uint32_t got = 0, spins = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(e + MB_SEQ_OFF) : "memory");
    if ((got & ~MB_POISON) == mb.seq || ++spins >= limit) break;
    __nanosleep(64);
  }
  if (got != mb.seq) {
    atomicExch(mb.err, mb.seq | MB_POISON);
    for (uint32_t k = 0; k < nwords; k++) dst[k] = 0;
    return false;
  }
  const volatile uint32_t* w = (const volatile uint32_t*)e;
  for (uint32_t k = 0; k < nwords; k++) dst[k] = w[k];
  return true;
}
#endif

}  // namespace reef
