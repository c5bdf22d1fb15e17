// Host simulation of the peer-mailbox exchange protocol of csrc/peer.cuh (test infrastructure): `world` threads play
// the ranks, each owning a mailbox of 64-bit atomics; every rank runs `rounds` exchanges of `words` values with the
// kernel's own slot / offset arithmetic (peer_off), posting (sequence << 32 | float bits) into every mailbox and
// polling its own until all sources carry the sequence number. Random delays skew the ranks by whole exchanges.
// Checks: every rank obtains the rank-ordered sum of every exchange (slot reuse never exposes a stale or a future
// word) and nobody waits forever. Usage: peer_sim <world> <rounds> <words> <seed>; exit code 0 = ok.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#define MMH_HOST_EMU 1
#include "../mmhand_b200/csrc/peer.cuh"

using namespace mmh;

static float value_of(int rank, uint32_t seq, int w) { return static_cast<float>((rank + 1) * 1000 + (seq % 97) * 7 + w % 13); }

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 8;
  const int rounds = argc > 2 ? atoi(argv[2]) : 2000;
  const int words = argc > 3 ? atoi(argv[3]) : 64;
  const unsigned seed = argc > 4 ? atoi(argv[4]) : 1;
  const size_t box_words = static_cast<size_t>(kPeerSlots) * world * kPeerWords;
  std::vector<std::vector<std::atomic<uint64_t>>> box(world);
  for (auto& b : box) { b = std::vector<std::atomic<uint64_t>>(box_words); for (auto& w : b) w.store(0); }
  std::atomic<int> failures{0};
  auto rank_fn = [&](int rank) {
    std::mt19937 rng(seed * 131 + rank);
    PeerDev p;
    memset(&p, 0, sizeof(p));
    p.rank = rank; p.world = world;
    for (uint32_t seq = 1; seq <= static_cast<uint32_t>(rounds) && failures.load() == 0; ++seq) {
      p.seq = seq;
      if (rng() % 50 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 300));   // skew
      for (int w = 0; w < words; ++w) {        // one "GPU thread" per word: post, then collect
        const float v = value_of(rank, seq, w);
        uint32_t bits; memcpy(&bits, &v, 4);
        const uint64_t word = (static_cast<uint64_t>(seq) << 32) | bits;
        for (int r = 0; r < world; ++r) box[r][peer_off(p, rank, w)].store(word, std::memory_order_relaxed);
        float s = 0.f, want = 0.f;
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < world; ++r) {
          uint64_t got;
          long spins = 0;
          while (true) {
            got = box[rank][peer_off(p, r, w)].load(std::memory_order_relaxed);
            if (static_cast<uint32_t>(got >> 32) == seq) break;
            if ((++spins & 0xFFF) == 0) {
              if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                fprintf(stderr, "rank %d: exchange %u word %d never arrived from rank %d (found seq %u)\n", rank, seq, w, r,
                        static_cast<uint32_t>(got >> 32));
                failures.fetch_add(1);
                return;
              }
              std::this_thread::yield();
            }
          }
          float f; const uint32_t lo = static_cast<uint32_t>(got); memcpy(&f, &lo, 4);
          s += f;
          want += value_of(r, seq, w);
        }
        if (s != want) {
          fprintf(stderr, "rank %d: exchange %u word %d: sum %g, expected %g\n", rank, seq, w, s, want);
          failures.fetch_add(1);
          return;
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(rank_fn, r);
  for (auto& t : th) t.join();
  if (failures.load() == 0) printf("ok: world %d, %d exchanges x %d words\n", world, rounds, words);
  return failures.load() == 0 ? 0 : 1;
}
