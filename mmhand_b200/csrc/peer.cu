// Synchronised-BatchNorm statistics exchange over NVLink peer memory, fused into the BN finalise kernels.
//
// The reference wraps its three networks in apex SyncBatchNorm (models/MMHandModel.py:109-116): every BN layer
// all-reduces 2*C floats in forward and again in backward, ~200 latency-bound collectives per step. Here the
// finalise kernel of a layer (one thread per channel, <= 2 blocks) does the exchange itself:
//
//   * every rank owns a mailbox  box[slot][src rank][word]  of 8-byte words, mapped into all peers (CUDA IPC);
//   * a thread posts its two partial sums to all ranks' mailboxes as (sequence number << 32 | float bits): one
//     aligned 64-bit store carries the datum AND its arrival flag (the LL idea of NCCL), so no fence is needed;
//   * it then polls its own mailbox until the word of every rank carries this exchange's sequence number and adds
//     the values in rank order -- every rank computes bit-identical global sums -- and goes on to the usual
//     mean / rstd / coefficient arithmetic.
//
// Slot reuse: exchange k+2 may overwrite slot (k & 1) because a rank can only finish exchange k+1 once every peer
// has *started* k+1, i.e. has finished reading k (launches of one rank are stream-ordered). Four slots are used.
// A wait that exceeds kPeerTimeoutNs (60 s; a peer died) raises the group's status word instead of hanging the GPU.
#include <stdint.h>
#include <string.h>

#include "../../include/mmhand_sm100.h"
#include "bn_finalize.h"
#include "peer.cuh"
#include "host_common.h"

#ifdef MMH_HOST_EMU
// The host emulation (CPU tests) has no peer memory: the entry points exist and fail loudly.
namespace mmh {
int peer_dev(const MmhPeer* g, uint32_t, int, PeerDev* d) {
  MMH_CHECK(g == nullptr, "peer mailboxes need CUDA devices");
  for (int r = 0; r < kPeerMaxWorld; ++r) d->box[r] = nullptr;
  d->status = nullptr; d->rank = 0; d->world = 1; d->seq = 0;
  return 0;
}
}  // namespace mmh
extern "C" int mmh_peer_create(int32_t, int32_t, MmhPeer**) { MMH_CHECK(false, "peer mailboxes need CUDA devices"); }
extern "C" int mmh_peer_handle(MmhPeer*, void*) { MMH_CHECK(false, "peer mailboxes need CUDA devices"); }
extern "C" int mmh_peer_connect(MmhPeer*, const void*) { MMH_CHECK(false, "peer mailboxes need CUDA devices"); }
extern "C" int mmh_peer_status(MmhPeer*) { return -1; }
extern "C" int mmh_peer_destroy(MmhPeer*) { return 0; }
extern "C" int mmh_peer_sum(MmhPeer*, uint32_t, float*, int32_t, void*) {
  MMH_CHECK(false, "peer mailboxes need CUDA devices");
}
extern "C" int mmh_bn_finalize_sync(MmhPeer*, uint32_t, float*, float, const float*, const float*, float*, float*,
                                    float, float, int32_t, float*, float*, void*) {
  MMH_CHECK(false, "peer mailboxes need CUDA devices");
}
extern "C" int mmh_bn_bwd_finalize_sync(MmhPeer*, uint32_t, const float*, float*, float, float*, float*, float*,
                                        int32_t, void*) {
  MMH_CHECK(false, "peer mailboxes need CUDA devices");
}
#else

#include <cuda_runtime.h>

namespace mmh {

__global__ void __launch_bounds__(256) peer_sum_kernel(const PeerDev p, float* __restrict__ data, const int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  peer_post(p, data[i], i);
  data[i] = peer_collect(p, i);
}

// sums (local partial sums) are replaced by the global sums, then the usual finalisation runs on them
__global__ void __launch_bounds__(256) bn_finalize_sync_kernel(const BnFinalizeF f, const PeerDev p, float* sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.C) return;
  peer_post(p, sums[c], c);
  peer_post(p, sums[f.C + c], f.C + c);
  sums[c] = peer_collect(p, c);
  sums[f.C + c] = peer_collect(p, f.C + c);
  f(c);
}
__global__ void __launch_bounds__(256) bn_bwd_finalize_sync_kernel(const BnBwdFinalizeF f, const PeerDev p,
                                                                   float* sums_global) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.C) return;
  peer_post(p, f.sl[c], c);
  peer_post(p, f.sl[f.C + c], f.C + c);
  sums_global[c] = peer_collect(p, c);
  sums_global[f.C + c] = peer_collect(p, f.C + c);
  f(c);
}

}  // namespace mmh

using namespace mmh;

struct MmhPeer {
  int rank, world, connected;
  unsigned long long* local;
  unsigned long long* peers[kPeerMaxWorld];
  int* status_host;
  int* status_dev;
};

static size_t mailbox_bytes(int world) {
  return static_cast<size_t>(kPeerSlots) * world * kPeerWords * sizeof(unsigned long long);
}

extern "C" int mmh_peer_create(int32_t rank, int32_t world, MmhPeer** out) {
  MMH_CHECK(out != nullptr, "null argument");
  MMH_CHECK(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "rank %d / world %d unsupported", rank,
            world);
  MmhPeer* g = new MmhPeer();
  memset(g, 0, sizeof(*g));
  g->rank = rank; g->world = world;
  // the mailbox is an IPC object: it must be the base of its own allocation, hence owned by the library
  MMH_CUDA(cudaMalloc(&g->local, mailbox_bytes(world)));
  MMH_CUDA(cudaMemset(g->local, 0, mailbox_bytes(world)));
  MMH_CUDA(cudaHostAlloc(&g->status_host, sizeof(int), cudaHostAllocMapped));
  *g->status_host = 0;
  MMH_CUDA(cudaHostGetDevicePointer(&g->status_dev, g->status_host, 0));
  MMH_CUDA(cudaDeviceSynchronize());
  g->peers[rank] = g->local;
  g->connected = world == 1;
  *out = g;
  return 0;
}

extern "C" int mmh_peer_handle(MmhPeer* g, void* handle64) {
  MMH_CHECK(g && handle64, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MMH_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  MMH_CUDA(cudaIpcGetMemHandle(&h, g->local));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}

extern "C" int mmh_peer_connect(MmhPeer* g, const void* handles) {
  MMH_CHECK(g && handles, "null argument");
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + static_cast<size_t>(r) * MMH_PEER_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    MMH_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->peers[r] = static_cast<unsigned long long*>(p);
  }
  g->connected = 1;
  return 0;
}

extern "C" int mmh_peer_status(MmhPeer* g) { return g == nullptr ? -1 : *reinterpret_cast<volatile int*>(g->status_host); }

extern "C" int mmh_peer_destroy(MmhPeer* g) {
  if (g == nullptr) return 0;
  for (int r = 0; r < g->world; ++r)
    if (r != g->rank && g->peers[r] != nullptr) cudaIpcCloseMemHandle(g->peers[r]);
  if (g->local) cudaFree(g->local);
  if (g->status_host) cudaFreeHost(g->status_host);
  delete g;
  return 0;
}

namespace mmh {
int peer_dev(const MmhPeer* g, uint32_t seq, int words, PeerDev* d) {
  if (g == nullptr) {                 // single GPU: no exchange
    for (int r = 0; r < kPeerMaxWorld; ++r) d->box[r] = nullptr;
    d->status = nullptr; d->rank = 0; d->world = 1; d->seq = 0;
    return 0;
  }
  MMH_CHECK(g->connected, "peer group not connected");
  MMH_CHECK(seq != 0, "sequence numbers start at 1");
  MMH_CHECK(words >= 1 && words <= kPeerWords, "%d words exceed the mailbox (%d)", words, kPeerWords);
  for (int r = 0; r < kPeerMaxWorld; ++r) d->box[r] = r < g->world ? g->peers[r] : nullptr;
  d->status = g->status_dev; d->rank = g->rank; d->world = g->world; d->seq = seq;
  return 0;
}
}  // namespace mmh

extern "C" int mmh_peer_sum(MmhPeer* g, uint32_t seq, float* data, int32_t n, void* stream) {
  MMH_CHECK(data != nullptr && g != nullptr, "null argument");
  PeerDev d;
  if (peer_dev(g, seq, n, &d)) return 1;
  peer_sum_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d, data, n);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mmh_bn_finalize_sync(MmhPeer* g, uint32_t seq, float* sums, float count_global, const float* gamma,
                                    const float* beta, float* running_mean, float* running_var, float momentum,
                                    float eps, int32_t C, float* coef, float* save, void* stream) {
  MMH_CHECK(sums && coef && save && g, "null argument");
  PeerDev d;
  if (peer_dev(g, seq, 2 * C, &d)) return 1;
  BnFinalizeF f;
  f.sums = sums; f.gamma = gamma; f.beta = beta; f.rm = running_mean; f.rv = running_var; f.coef = coef; f.save = save;
  f.count = count_global; f.momentum = momentum; f.eps = eps; f.train = 1; f.C = C;
  bn_finalize_sync_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(f, d, sums);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int mmh_bn_bwd_finalize_sync(MmhPeer* g, uint32_t seq, const float* sums_local, float* sums_global,
                                        float count_global, float* k, float* dgamma, float* dbeta, int32_t C,
                                        void* stream) {
  MMH_CHECK(sums_local && sums_global && k && g, "null argument");
  PeerDev d;
  if (peer_dev(g, seq, 2 * C, &d)) return 1;
  BnBwdFinalizeF f;
  f.sg = sums_global; f.sl = sums_local; f.k = k; f.dgamma = dgamma; f.dbeta = dbeta; f.count = count_global; f.C = C;
  bn_bwd_finalize_sync_kernel<<<(C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(f, d, sums_global);
  MMH_CUDA(cudaGetLastError());
  return 0;
}

#endif  // MMH_HOST_EMU
