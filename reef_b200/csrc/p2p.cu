// Small-message all-gather over NVLink peer memory for the sharded sum-check.
//
// The reference is a single process (no collective exists in it); on an 8 x B200 box the nldoc
// sum-check is sharded over the GPUs and every round needs the ranks' (const, g(1), xsq) triples
// -- 96 bytes per rank -- on every rank before the transcript can continue.  At that size an
// NCCL all-gather is pure launch/protocol latency (~20 us plus its host call) on a path that runs
// ~50 times per pass, so the exchange is done by the producer itself: each rank owns a MAILBOX in
// its HBM, mapped into every peer with CUDA IPC; one tiny kernel stores this rank's payload into
// all peers' mailboxes with plain P2P stores over NVLink, publishes it with a system-scope release
// store of a sequence number, and then acquires the peers' sequence numbers in its own mailbox.
// No host round trip, no NCCL kernel: the exchange is one stream-ordered launch.
//
// Mailbox layout: [2 slots][world] entries of 256 bytes: payload (<= 248 B) + sequence number at
// byte 248.  Two slots suffice: a rank can be at most one exchange ahead of any peer (it cannot
// finish exchange k+1 before every peer has posted k+1, which a peer does only after it has
// consumed exchange k).
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "p2p.cuh"

namespace reef {

__global__ void __launch_bounds__(32) k_p2p_allgather(MbRef mb, const uint32_t* __restrict__ src, uint32_t nwords,
                                                      uint32_t* __restrict__ dst) {
  const uint32_t t = threadIdx.x;
  if (t >= mb.world) return;
  mb_post(mb, t, src, nwords);
  mb_wait_copy(mb, t, dst + (size_t)t * nwords, nwords);
}

}  // namespace reef

namespace reef {
// stream-ordered all-gather of `nwords` 32-bit words per rank through the mailboxes (context lock held by the caller)
int launch_p2p_allgather(reef_ctx* c, const void* src_dev, uint32_t nwords, void* dst_dev) {
  REEF_REQUIRE(c->mb_world >= 1, REEF_EINVAL, "p2p all-gather: mailbox not connected (reef_mailbox_connect)");
  REEF_REQUIRE(nwords >= 1 && nwords * 4 <= MB_SEQ_OFF, REEF_EINVAL, "p2p all-gather: payload must be 4..248 bytes");
  MbRef mb{c->mb_peers_dev, (const unsigned char*)c->mb_mine, c->mb_err_dev, c->mb_world, c->mb_rank, c->mb_seq + 1};
  k_p2p_allgather<<<1, 32, 0, c->stream>>>(mb, (const uint32_t*)src_dev, nwords, (uint32_t*)dst_dev);
  REEF_LAUNCHED();
  c->mb_seq++;
  return REEF_OK;
}
}  // namespace reef

using namespace reef;

static int mailbox_alloc(reef_ctx* c, uint32_t world) {
  if (c->mb_mine) return REEF_OK;
  const size_t bytes = (size_t)2 * MB_MAX_WORLD * MB_ENTRY;
  REEF_CUDA(cudaMalloc(&c->mb_mine, bytes));
  REEF_CUDA(cudaMemset(c->mb_mine, 0, bytes));
  REEF_CUDA(cudaMalloc((void**)&c->mb_peers_dev, MB_MAX_WORLD * sizeof(void*)));
  REEF_CUDA(cudaMalloc((void**)&c->mb_err_dev, 256));
  REEF_CUDA(cudaMemset(c->mb_err_dev, 0, 256));
  REEF_CUDA(cudaDeviceSynchronize());
  (void)world;
  return REEF_OK;
}

static int mailbox_set_peers(reef_ctx* c, uint32_t rank, uint32_t world, void* const* ptrs) {
  REEF_CUDA(cudaMemcpy(c->mb_peers_dev, ptrs, world * sizeof(void*), cudaMemcpyHostToDevice));
  cudaFuncAttributes fa;
  REEF_CUDA(cudaFuncGetAttributes(&fa, (const void*)k_p2p_allgather));   // load it now, see nl_shard_preload
  int rc = nl_shard_preload();
  if (rc) return rc;
  rc = msm_preload();
  if (rc) return rc;
  REEF_CUDA(cudaMemset(c->mb_err_dev, 0, 256));   // a fresh connection starts without a recorded failure
  c->mb_world = world;
  c->mb_rank = rank;
  // mb_seq stays monotonic over the life of the context: the mailbox still holds the sequence
  // numbers of earlier exchanges, and a peer may already be posting into it
  return REEF_OK;
}

extern "C" {

int reef_mailbox_create(reef_ctx* c, uint32_t world, uint8_t out_handle[64]) {
  REEF_REQUIRE(c && out_handle, REEF_EINVAL, "reef_mailbox_create: NULL argument");
  REEF_REQUIRE(world >= 1 && world <= MB_MAX_WORLD, REEF_EINVAL, "reef_mailbox_create: world must be in 1..32");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  int rc = mailbox_alloc(c, world);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  REEF_CUDA(cudaIpcGetMemHandle(&h, c->mb_mine));
  memcpy(out_handle, &h, 64);
  return REEF_OK;
}

void* reef_mailbox_ptr(reef_ctx* c) { return c ? c->mb_mine : nullptr; }

int reef_mailbox_connect(reef_ctx* c, uint32_t rank, uint32_t world, const uint8_t* handles) {
  REEF_REQUIRE(c && handles, REEF_EINVAL, "reef_mailbox_connect: NULL argument");
  REEF_REQUIRE(world >= 1 && world <= MB_MAX_WORLD && rank < world, REEF_EINVAL, "reef_mailbox_connect: bad rank / world");
  REEF_REQUIRE(c->mb_mine != nullptr, REEF_EINVAL, "reef_mailbox_connect: call reef_mailbox_create first");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  std::vector<void*> ptrs(world, nullptr);
  for (uint32_t g = 0; g < world; g++) {
    if (g == rank) {
      ptrs[g] = c->mb_mine;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)g * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(REEF_ECUDA, std::string("reef_mailbox_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(g) + "): " + cudaGetErrorString(e));
    c->mb_ipc_opened.push_back(p);
    ptrs[g] = p;
  }
  return mailbox_set_peers(c, rank, world, ptrs.data());
}

int reef_mailbox_connect_local(reef_ctx* c, uint32_t rank, uint32_t world, void* const* mailboxes) {
  REEF_REQUIRE(c && mailboxes, REEF_EINVAL, "reef_mailbox_connect_local: NULL argument");
  REEF_REQUIRE(world >= 1 && world <= MB_MAX_WORLD && rank < world, REEF_EINVAL, "reef_mailbox_connect_local: bad rank / world");
  REEF_REQUIRE(c->mb_mine != nullptr && mailboxes[rank] == c->mb_mine, REEF_EINVAL,
               "reef_mailbox_connect_local: mailboxes[rank] must be this context's own mailbox");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return mailbox_set_peers(c, rank, world, mailboxes);
}

int reef_p2p_allgather(reef_ctx* c, const void* mine_dev, uint32_t nbytes, void* out_dev) {
  REEF_REQUIRE(c && mine_dev && out_dev, REEF_EINVAL, "reef_p2p_allgather: NULL argument");
  REEF_REQUIRE(c->mb_world >= 1, REEF_EINVAL, "reef_p2p_allgather: mailbox not connected");
  REEF_REQUIRE(nbytes >= 4 && nbytes <= MB_SEQ_OFF && (nbytes & 3) == 0, REEF_EINVAL, "reef_p2p_allgather: payload must be 4..248 bytes, a multiple of 4");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return launch_p2p_allgather(c, mine_dev, nbytes / 4, out_dev);
}

int reef_p2p_status(reef_ctx* c) {
  REEF_REQUIRE(c, REEF_EINVAL, "reef_p2p_status: NULL argument");
  if (!c->mb_err_dev) return REEF_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  uint32_t e = 0;
  REEF_CUDA(cudaMemcpyAsync(&e, c->mb_err_dev, 4, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  if (e) {
    cudaMemset(c->mb_err_dev, 0, 4);
    return fail(REEF_ECUDA, "reef_p2p_allgather: exchange " + std::to_string(e & 0x7fffffffu) + " failed (a peer never posted, or a peer reported a timeout)");
  }
  return REEF_OK;
}

}  // extern "C"
