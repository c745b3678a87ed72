// Host emulation of the cross-GPU epoch barrier (coll/src/barrier.cuh: the text nvcc compiles), classic flag exchange and — built with
// -DB200COLL_VARIANT_MCBAR — the multicast-counter flavour, whose multimem.red is emulated as "add to the same word of every arena".
// Same idea as p2p_emu.cc: CTAs are groups of host threads, the flag primitives are C++ atomics of the same strength, ThreadSanitizer
// checks that the data phase between the two barriers is ordered by them. One "kernel" = one launch of all CTAs of all ranks (thread
// creation / join stands in for stream order), structured like the library's kernels:
//   s = load_seq; barrier<relaxed>(2s+1); every thread stores a stamp into every peer's arena; barrier<release>(2s+2); ticket bumps s.
// After each launch the host checks that every stamp of that launch is there; launches alternate grid sizes (CTAs that sit out a launch
// must not confuse the counters) and ranks start skewed.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#include "../src/layout.h"

struct Dim { unsigned x; };
static thread_local Dim threadIdx{0}, blockIdx{0}, blockDim{1}, gridDim{1};
struct CtaBarrier {
  std::atomic<unsigned> arrived{0}, phase{0};
  std::atomic<int> any{0};            // __syncthreads_or: sticky (a watchdog hit is terminal)
  unsigned n = 1;
  void wait() {
    const unsigned p = phase.load(std::memory_order_acquire);
    if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { arrived.store(0, std::memory_order_relaxed); phase.store(p + 1, std::memory_order_release); }
    else while (phase.load(std::memory_order_acquire) == p) std::this_thread::yield();
  }
};
static thread_local CtaBarrier* g_bar = nullptr;
#define __device__
#define __forceinline__ inline
static inline void __syncthreads() { g_bar->wait(); }
static inline int __syncthreads_or(int pred) { if (pred) g_bar->any.store(1, std::memory_order_relaxed); g_bar->wait(); return g_bar->any.load(std::memory_order_relaxed); }
static inline void __threadfence() {}      /* ThreadSanitizer does not model fences; the ticket below is an acq_rel RMW, which it does */
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }

struct Machine;
static Machine* g_machine = nullptr;

namespace b200coll {
static inline unsigned long long globaltimer_ns() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (unsigned long long)t.tv_sec * 1000000000ull + t.tv_nsec; }
static inline void record_fault(const CommDev& c, uint32_t code, uint32_t peer, uint32_t expected, uint32_t observed, uint32_t op) {
  uint32_t zero = 0;
  if (__atomic_compare_exchange_n(&c.fault->code, &zero, code, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) { c.fault->rank = c.rank; c.fault->peer = peer; c.fault->block = blockIdx.x; c.fault->expected = expected; c.fault->observed = observed; c.fault->op = op; }
}
static inline uint32_t ld_volatile_u32(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline uint32_t ld_relaxed_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline void st_relaxed_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
static inline void mc_red_add(void* mc_addr, uint32_t v, int order);
static inline void mc_red_add_release_u32(void* a, uint32_t v) { mc_red_add(a, v, __ATOMIC_RELEASE); }
static inline void mc_red_add_relaxed_u32(void* a, uint32_t v) { mc_red_add(a, v, __ATOMIC_RELAXED); }
}  // namespace b200coll

#include "../src/barrier.cuh"

using namespace b200coll;

struct Rank { char* arena = nullptr; std::vector<uint32_t> state = std::vector<uint32_t>(kStateWords, 0); b200collFault fault{}; CommDev dev{}; };
struct Machine {
  int n; std::vector<Rank> ranks; char* mc_base;
  Machine(int n_, bool multicast) : n(n_), ranks(n_), mc_base(multicast ? reinterpret_cast<char*>((uintptr_t)1 << 46) : nullptr) {     // an address nothing else uses: only offsets from it matter
    for (auto& r : ranks) r.arena = static_cast<char*>(calloc(1, 2u << 20));
    for (int i = 0; i < n; i++) {
      CommDev& d = ranks[i].dev;
      d.rank = i; d.nranks = n; d.mc = mc_base; d.state = ranks[i].state.data(); d.fault = &ranks[i].fault; d.timeout_ns = 20000ull * 1000000ull; d.mcbar = multicast ? 1 : 0;
      for (int p = 0; p < kMaxRanks; p++) d.peer[p] = ranks[p < n ? p : i].arena;
    }
  }
  ~Machine() { for (auto& r : ranks) free(r.arena); }
};
namespace b200coll {
static inline void mc_red_add(void* mc_addr, uint32_t v, int order) {       // the switch: the same word of every rank's arena
  const size_t off = (size_t)(static_cast<char*>(mc_addr) - g_machine->mc_base);
  for (auto& r : g_machine->ranks) __atomic_fetch_add(reinterpret_cast<uint32_t*>(r.arena + off), v, order);
}
}  // namespace b200coll

constexpr size_t kOffData = 1u << 20;      // stamps: u32 [source rank][CTA][thread], in the second megabyte of each arena
constexpr int kMaxCtas = 8, kMaxThreads = 8;
static uint32_t* stamp(char* arena, int src, int cta, int t) { return reinterpret_cast<uint32_t*>(arena + kOffData) + ((size_t)src * kMaxCtas + cta) * kMaxThreads + t; }

static std::atomic<int> g_early_read_failures{0};
static void kernel(CommDev c, uint32_t launch_no) {
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, 0);
  for (int p = 0; p < c.nranks; p++) *stamp(c.peer[p], c.rank, blockIdx.x, threadIdx.x) = launch_no * 1000 + c.rank;      // plain stores into peers
  barrier_blocks<true>(c, 2 * s + 2, 0);
  // what this rank's NEXT kernel may do while other ranks are still inside this one: read what the namesake CTAs of every peer wrote
  for (int p = 0; p < c.nranks; p++)
    for (unsigned t = 0; t < blockDim.x; t++)
      if (*stamp(c.peer[c.rank], p, blockIdx.x, t) != launch_no * 1000 + p) g_early_read_failures.fetch_add(1);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

static int run(int n, bool multicast, int launches, unsigned T) {
  Machine m(n, multicast);
  g_machine = &m;
  int failures = 0;
  for (int it = 1; it <= launches; it++) {
    const int grid = (it % 3 == 0) ? 2 : 5;                     // CTAs 2..4 sit some launches out
    std::vector<std::thread> threads;
    std::vector<std::unique_ptr<CtaBarrier>> bars;
    for (int r = 0; r < n; r++)
      for (int b = 0; b < grid; b++) {
        bars.emplace_back(new CtaBarrier()); bars.back()->n = T;
        CtaBarrier* bar = bars.back().get();
        for (unsigned t = 0; t < T; t++)
          threads.emplace_back([&m, bar, r, b, t, T, grid, it] {
            if ((r + it) % 3 == 0) { timespec ts{0, 300000}; nanosleep(&ts, nullptr); }       // this rank enters the kernel late
            g_bar = bar; threadIdx.x = t; blockDim.x = T; blockIdx.x = (unsigned)b; gridDim.x = (unsigned)grid;
            kernel(m.ranks[r].dev, (uint32_t)it);
          });
      }
    for (auto& t : threads) t.join();
    for (int r = 0; r < n; r++) {
      if (m.ranks[r].fault.code) { printf("FAIL fault on rank %d launch %d: code=%u peer=%u expected=%u observed=%u\n", r, it, m.ranks[r].fault.code, m.ranks[r].fault.peer, m.ranks[r].fault.expected, m.ranks[r].fault.observed); return 1; }
      if (m.ranks[r].state[kSeqBarrier] != (uint32_t)it) { printf("FAIL rank %d launch %d: sequence word %u\n", r, it, m.ranks[r].state[kSeqBarrier]); failures++; }
      for (int src = 0; src < n; src++) for (int b = 0; b < grid; b++) for (unsigned t = 0; t < T; t++)
        if (*stamp(m.ranks[r].arena, src, b, t) != (uint32_t)it * 1000 + src) { if (failures < 5) printf("FAIL rank %d launch %d: stamp from rank %d cta %d thread %u is %u\n", r, it, src, b, t, *stamp(m.ranks[r].arena, src, b, t)); failures++; }
    }
  }
  return failures + g_early_read_failures.exchange(0);
}

int main(int argc, char** argv) {
  const int launches = argc > 1 ? atoi(argv[1]) : 40;
  int failures = 0;
  for (int n : {2, 4, 8}) {
    const unsigned T = (unsigned)n;                              // the flag exchange uses one thread per peer
    failures += run(n, false, launches, T);
#ifdef B200COLL_VARIANT_MCBAR
    failures += run(n, true, launches, T);
#endif
  }
#ifdef B200COLL_VARIANT_MCBAR
  printf(failures ? "barrier_emu (flags + multicast counter): %d FAILURES\n" : "barrier_emu (flags + multicast counter): all launches consistent\n", failures);
#else
  printf(failures ? "barrier_emu (flags): %d FAILURES\n" : "barrier_emu (flags): all launches consistent\n", failures);
#endif
  return failures ? 1 : 0;
}
