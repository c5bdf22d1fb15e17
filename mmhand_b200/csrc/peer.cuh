// Device side of the peer-memory exchange (see peer.cu for the protocol) + the finalisers that run in the last
// block of a statistics kernel: (exchange over NVLink) -> BatchNorm finalisation -> accumulators back to zero.
#pragma once
#include <stdint.h>

#include "../../include/mmhand_sm100.h"
#include "bn_finalize.h"

namespace mmh {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerSlots = 4;
constexpr int kPeerWords = 2048;                      // >= 2 * C of the widest BN layer (C <= 1024)

struct PeerDev {
  unsigned long long* box[kPeerMaxWorld];
  int* status;
  int rank, world;                                    // world == 1: no exchange
  uint32_t seq;
};

// word w of source rank `src` in the slot of exchange p.seq (same arithmetic on every rank's mailbox)
MMH_HD size_t peer_off(const PeerDev& p, int src, int w) {
  return (static_cast<size_t>(p.seq & (kPeerSlots - 1)) * p.world + src) * kPeerWords + w;
}

#ifndef MMH_HOST_EMU
constexpr unsigned long long kPeerTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;   // once: later waits fail fast
__device__ __forceinline__ void peer_post(const PeerDev& p, float v, int w) {
  const unsigned long long word = (static_cast<unsigned long long>(p.seq) << 32) | __float_as_uint(v);
  const size_t off = peer_off(p, p.rank, w);
  for (int r = 0; r < p.world; ++r)
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.box[r] + off), "l"(word) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float peer_collect(const PeerDev& p, int w) {
  float s = 0.f;
  unsigned long long t0 = 0;
  // a wait already timed out (a peer is gone): do not spend the time-out again on every later exchange
  if (*reinterpret_cast<volatile int*>(p.status) != 0) return __int_as_float(0x7FC00000);
  for (int r = 0; r < p.world; ++r) {
    const unsigned long long* src = p.box[p.rank] + peer_off(p, r, w);
    unsigned long long word;
    uint32_t spins = 0;
    while (true) {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(src) : "memory");
      if (static_cast<uint32_t>(word >> 32) == p.seq) break;
      if ((++spins & 0x3FFu) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > kPeerTimeoutNs) {
          *p.status = 1;                              // host-mapped: visible to mmh_peer_status without a sync
          return __int_as_float(0x7FC00000);
        }
      }
    }
    s += __uint_as_float(static_cast<uint32_t>(word));
  }
  return s;
}
#endif
// the accumulators were written with atomics (L2) by other blocks: read them past L1
MMH_HD float ld_acc(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}

// channel c: local sums -> (exchange) -> finalisation. `sums` is zeroed by the launcher afterwards.
struct BnFwdFin {
  BnFinalizeF f; const float* sums; PeerDev px;
  MMH_HD void operator()(int c) const {
    float s0 = ld_acc(sums + c), s1 = ld_acc(sums + f.C + c);
#if defined(__CUDA_ARCH__)
    if (px.world > 1) {
      peer_post(px, s0, c);
      peer_post(px, s1, f.C + c);
      s0 = peer_collect(px, c);
      s1 = peer_collect(px, f.C + c);
    }
#endif
    f.apply(c, s0, s1);
  }
};
struct BnBwdFin {
  BnBwdFinalizeF f; const float* sums; PeerDev px;
  MMH_HD void operator()(int c) const {
    const float l0 = ld_acc(sums + c), l1 = ld_acc(sums + f.C + c);
    float g0 = l0, g1 = l1;
#if defined(__CUDA_ARCH__)
    if (px.world > 1) {
      peer_post(px, l0, c);
      peer_post(px, l1, f.C + c);
      g0 = peer_collect(px, c);
      g1 = peer_collect(px, f.C + c);
    }
#endif
    f.apply(c, g0, g1, l0, l1);
  }
};

// host side (peer.cu): device view of a connected group; g == NULL gives the single-GPU view (world = 1)
int peer_dev(const MmhPeer* g, uint32_t seq, int words, PeerDev* d);

}  // namespace mmh
