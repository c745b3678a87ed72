// Host emulation of the point-to-point protocol: compiles coll/src/p2p.cuh — the very text nvcc compiles for sm_100a — with g++,
// every CTA of every rank a group of std::threads (--threads N per CTA, default 1; thread 0 runs the protocol, the others move data and
// meet it at __syncthreads), arenas in host memory,
// the cross-GPU flag primitives mapped to C++ atomics of the same strength (relaxed / release / acquire). The scenarios below are the
// ones tests/test_coll_gpu.py runs on hardware; here they run under a real scheduler and, in the `tsan` build, under ThreadSanitizer,
// which checks exactly what the protocol promises: every plain payload access is ordered by a release/acquire pair on a flag.
// What this cannot show: PTX memory-model corner cases, NVLink write ordering, co-residency of CTAs. What it does show: sequence
// numbers, slot reuse, window reuse, chunking, share splitting, fault paths and graph-replay safety (state lives in "device" memory).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#include "../src/layout.h"

// ---- stand-ins for the CUDA side ------------------------------------------------------------------------------------------------
struct uint4 { uint32_t x, y, z, w; };
struct Dim { unsigned x; };
static thread_local Dim threadIdx{0}, blockIdx{0}, blockDim{1}, gridDim{1};
// A CTA is a group of host threads that share one P2pShared and one barrier (bar.sync). The barrier is a plain sense-reversing counter on
// C++ atomics (acq_rel), so ThreadSanitizer sees exactly the ordering __syncthreads() gives and flags anything that relies on more.
struct CtaBarrier {
  std::atomic<unsigned> arrived{0}, phase{0};
  unsigned n = 1;
  void wait() {
    const unsigned p = phase.load(std::memory_order_acquire);
    if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == n) { arrived.store(0, std::memory_order_relaxed); phase.store(p + 1, std::memory_order_release); }
    else while (phase.load(std::memory_order_acquire) == p) std::this_thread::yield();
  }
};
struct CtaCtx;
static thread_local CtaCtx* g_cta = nullptr;
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(n)
#define __grid_constant__
#define COMM_PARAM CommDev c
#define P2P_SHARED(name) P2pShared& name = emu_shared()
namespace b200coll { struct P2pShared; }
static inline b200coll::P2pShared& emu_shared();
static inline void __syncthreads();

namespace b200coll {
static inline void pdl_prologue() {}
static inline unsigned long long globaltimer_ns() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (unsigned long long)t.tv_sec * 1000000000ull + t.tv_nsec; }
static inline void record_fault(const CommDev& c, uint32_t code, uint32_t peer, uint32_t expected, uint32_t observed, uint32_t op) {
  uint32_t zero = 0;
  if (__atomic_compare_exchange_n(&c.fault->code, &zero, code, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
    c.fault->rank = c.rank; c.fault->peer = peer; c.fault->block = blockIdx.x; c.fault->expected = expected; c.fault->observed = observed; c.fault->op = op;
  }
}
static inline uint32_t ld_volatile_u32(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline uint32_t ld_relaxed_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline unsigned long long ld_acquire_sys_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
static inline void st_release_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline uint4 ld_vec(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
static inline uint4 ld_vec_volatile(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
static inline void st_vec(void* p, const uint4& v) { memcpy(p, &v, 16); }
}  // namespace b200coll

#include "../src/p2p.cuh"

using namespace b200coll;

struct CtaCtx { CtaBarrier bar; P2pShared shared{}; };
static inline P2pShared& emu_shared() { return g_cta->shared; }
static inline void __syncthreads() { g_cta->bar.wait(); }
static unsigned g_threads_per_cta = 1;      // --threads N: N host threads per CTA (exercises the bar.sync placement and the strided loops)

// ---- a "machine": n ranks, each with an arena, state words and a fault record -----------------------------------------------------
struct Rank {
  char* arena = nullptr;
  std::vector<uint32_t> state = std::vector<uint32_t>(kStateWords, 0);
  b200collFault fault{};
  CommDev dev{};
};
struct Machine {
  int n;
  size_t arena_bytes;
  std::vector<Rank> ranks;
  Machine(int n_, size_t heap_bytes, unsigned long long timeout_ms) : n(n_), arena_bytes(kOffHeap + heap_bytes), ranks(n_) {
    for (auto& r : ranks) { r.arena = static_cast<char*>(calloc(1, arena_bytes)); if (!r.arena) { perror("calloc"); exit(2); } }
    for (int i = 0; i < n; i++) {
      CommDev& d = ranks[i].dev;
      d.rank = i; d.nranks = n; d.mc = nullptr; d.state = ranks[i].state.data(); d.fault = &ranks[i].fault; d.timeout_ns = timeout_ms * 1000000ull;
      for (int p = 0; p < kMaxRanks; p++) d.peer[p] = ranks[p < n ? p : i].arena;
    }
  }
  ~Machine() { for (auto& r : ranks) free(r.arena); }
};

struct Op { bool send; int peer; char* buf; size_t bytes; bool in_arena; };

// mirrors p2p_blocks / p2p_launch of collectives.cu (that code is checked through b200collDebugPlanP2p in tests/test_coll_cpu.py)
static int blocks_for(size_t bytes, int cap) { return (int)std::max<size_t>(1, std::min<size_t>((bytes + (128u << 10) - 1) / (128u << 10), (size_t)cap)); }
static P2pArgs plan(const Machine& m, int rank, const std::vector<Op>& ops, int cap, size_t window_cap) {
  P2pArgs a = {};
  int nstaged = 0;
  for (const Op& o : ops) if (!o.send && !o.in_arena) nstaged++;
  const size_t share = nstaged ? (2 * kStageHalfBytes / (size_t)nstaged) / 1024 * 1024 : 0;
  const size_t window = window_cap ? std::min(window_cap, share / 2) : share / 2;
  int blocks = 0, seen = 0;
  for (int pass = 0; pass < 2; pass++)
    for (const Op& o : ops) {
      if (o.send != (pass == 0)) continue;
      const int i = a.nops++;
      a.first_block[i] = blocks; a.lanes[i] = blocks_for(o.bytes, cap); blocks += (!o.send && o.in_arena) ? 1 : a.lanes[i];
      a.peer[i] = o.peer; a.bytes[i] = o.bytes;
      if (o.send) { a.nsend++; a.src[i] = o.buf; }
      else {
        a.dst[i] = o.buf;
        if (o.in_arena) { a.staged[i] = 0; a.win_off[i] = (size_t)(o.buf - m.ranks[rank].arena); a.win_bytes[i] = o.bytes; }
        else { a.staged[i] = 1; a.win_off[i] = kOffStage + (size_t)seen++ * share; a.win_bytes[i] = window; }
      }
    }
  a.first_block[a.nops] = blocks;
  return a;
}

// one "kernel launch" per rank, all at once: every CTA is a thread
static void launch_all(Machine& m, const std::vector<std::vector<Op>>& per_rank, int cap, size_t window_cap, unsigned skew_us = 0) {
  std::vector<std::thread> threads;
  std::vector<P2pArgs> args(m.n);
  for (int r = 0; r < m.n; r++) args[r] = plan(m, r, per_rank[r], cap, window_cap);
  std::vector<std::unique_ptr<CtaCtx>> ctas;
  const unsigned T = g_threads_per_cta;
  for (int r = 0; r < m.n; r++) {
    if (per_rank[r].empty()) continue;
    const int grid = args[r].first_block[args[r].nops];
    for (int b = 0; b < grid; b++) {
      ctas.emplace_back(new CtaCtx());
      CtaCtx* ctx = ctas.back().get();
      ctx->bar.n = T;
      for (unsigned t = 0; t < T; t++)
        threads.emplace_back([&m, &args, ctx, r, b, t, T, grid, skew_us] {
          if (skew_us && (r + b + t) % 3 == 0) { timespec ts{0, (long)skew_us * 1000}; nanosleep(&ts, nullptr); }
          g_cta = ctx;
          threadIdx.x = t; blockDim.x = T; blockIdx.x = (unsigned)b; gridDim.x = (unsigned)grid;
          k_p2p(m.ranks[r].dev, args[r], 6u);
        });
    }
  }
  for (auto& t : threads) t.join();
}

static void fill(char* p, size_t n, unsigned seed) { for (size_t i = 0; i < n; i++) p[i] = (char)((i * (2 * seed + 7) + seed * 31 + (i >> 8)) % 251); }
static int g_failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d %s  ", __FILE__, __LINE__, #cond); printf(__VA_ARGS__); printf("\n"); g_failures++; } } while (0)
static bool no_faults(Machine& m) { for (auto& r : m.ranks) if (r.fault.code) { printf("  fault code=%u rank=%u peer=%u block=%u expected=%u observed=%u\n", r.fault.code, r.fault.rank, r.fault.peer, r.fault.block, r.fault.expected, r.fault.observed); return false; } return true; }

// ring step, repeated with sizes that grow and shrink (CTA counts change between launches), arena or staged receive
static void scenario_ring(int n, bool arena_recv, size_t window_cap, int cap) {
  const size_t sizes[] = {16, 1003, (1u << 20) + 48, 200, (3u << 20) + 5, 4096, (3u << 20) + 5};
  const size_t maxb = 4u << 20;
  Machine m(n, 2 * maxb + 8192, 20000);
  std::vector<char*> src(n), dst(n);
  std::vector<std::vector<char>> outside(n, std::vector<char>(arena_recv ? 16 : maxb + 16));
  for (int r = 0; r < n; r++) {
    src[r] = m.ranks[r].arena + kOffHeap;
    dst[r] = arena_recv ? m.ranks[r].arena + kOffHeap + maxb + 4096 : reinterpret_cast<char*>(((uintptr_t)outside[r].data() + 15) & ~(uintptr_t)15);
  }
  int rep = 0;
  for (size_t bytes : sizes) {
    for (int r = 0; r < n; r++) { fill(src[r], bytes, r + 10 * rep); memset(dst[r], 0xEE, bytes + 1); }
    std::vector<std::vector<Op>> ops(n);
    for (int r = 0; r < n; r++) ops[r] = {Op{true, (r + 1) % n, src[r], bytes, true}, Op{false, (r + n - 1) % n, dst[r], bytes, arena_recv}};
    launch_all(m, ops, cap, window_cap, rep % 2 ? 200 : 0);
    CHECK(no_faults(m), "ring n=%d bytes=%zu", n, bytes);
    for (int r = 0; r < n; r++) {
      CHECK(memcmp(dst[r], src[(r + n - 1) % n], bytes) == 0, "ring n=%d rank %d bytes=%zu arena=%d window=%zu", n, r, bytes, (int)arena_recv, window_cap);
      CHECK((unsigned char)dst[r][bytes] == 0xEE, "ring wrote past the end: rank %d bytes=%zu", r, bytes);
    }
    rep++;
  }
}

// every pair exchanges messages of different sizes in ONE launch per rank; mixed arena / staged receives
static void scenario_all_pairs(int n, int cap) {
  Machine m(n, (size_t)n * n * (1u << 20), 20000);
  auto size_of = [](int s, int d) { return (size_t)4096 * (1 + s) + 16 * d + (s * 7 + d) % 13; };
  std::vector<std::vector<Op>> ops(n);
  std::vector<std::vector<std::vector<char>>> outside(n, std::vector<std::vector<char>>(n));
  std::vector<std::vector<char*>> out(n, std::vector<char*>(n)), in(n, std::vector<char*>(n));
  for (int r = 0; r < n; r++) {
    size_t cursor = kOffHeap;
    for (int p = 0; p < n; p++) {
      if (p == r) continue;
      out[r][p] = m.ranks[r].arena + cursor; cursor += (size_of(r, p) + 4095) / 4096 * 4096;
      fill(out[r][p], size_of(r, p), 17 * r + p);
      const bool arena = (r + p) % 2 == 0;
      if (arena) { in[r][p] = m.ranks[r].arena + cursor; cursor += (size_of(p, r) + 4095) / 4096 * 4096; }
      else { outside[r][p].resize(size_of(p, r) + 32); in[r][p] = reinterpret_cast<char*>(((uintptr_t)outside[r][p].data() + 15) & ~(uintptr_t)15); }
      ops[r].push_back(Op{true, p, out[r][p], size_of(r, p), true});
      ops[r].push_back(Op{false, p, in[r][p], size_of(p, r), arena});
    }
  }
  for (int rep = 0; rep < 3; rep++) {
    launch_all(m, ops, cap, 4096, rep == 1 ? 300 : 0);       // 4 KiB windows: most staged messages take several chunks
    CHECK(no_faults(m), "all pairs n=%d rep=%d", n, rep);
    for (int r = 0; r < n; r++) for (int p = 0; p < n; p++) if (p != r) CHECK(memcmp(in[r][p], out[p][r], size_of(p, r)) == 0, "all pairs %d<-%d rep %d", r, p, rep);
  }
}

// pipeline hand-over: lone send on rank 0, lone recv on rank 1, the receiver arriving late; then the other direction
static void scenario_handover() {
  Machine m(2, 1u << 20, 20000);
  char* a = m.ranks[0].arena + kOffHeap; char* b = m.ranks[1].arena + kOffHeap;
  fill(a, 70000, 5); memset(b, 0, 70000);
  launch_all(m, {{Op{true, 1, a, 70000, true}}, {Op{false, 0, b, 70000, true}}}, 16, 0, 2000);
  CHECK(no_faults(m) && memcmp(a, b, 70000) == 0, "hand-over 0->1");
  std::vector<char> back(200 + 16);
  char* bk = reinterpret_cast<char*>(((uintptr_t)back.data() + 15) & ~(uintptr_t)15);
  launch_all(m, {{Op{false, 1, bk, 200, false}}, {Op{true, 0, b, 200, true}}}, 16, 0, 0);
  CHECK(no_faults(m) && memcmp(bk, a, 200) == 0, "hand-over 1->0 staged");
}

// nobody receives: the sender's watchdog fires (code 3) and the kernel returns; sizes that differ are code 4 on the sender
static void scenario_faults() {
  {
    Machine m(2, 1u << 20, 200);
    launch_all(m, {{Op{true, 1, m.ranks[0].arena + kOffHeap, 4096, true}}, {}}, 16, 0);
    CHECK(m.ranks[0].fault.code == 3 && m.ranks[0].fault.peer == 1, "lone send: code=%u", m.ranks[0].fault.code);
  }
  {
    Machine m(2, 1u << 20, 200);
    launch_all(m, {{}, {Op{false, 0, m.ranks[1].arena + kOffHeap, 4096, true}}}, 16, 0);
    CHECK(m.ranks[1].fault.code == 3 && m.ranks[1].fault.peer == 0, "lone recv: code=%u", m.ranks[1].fault.code);
  }
  {
    Machine m(2, 1u << 20, 300);
    launch_all(m, {{Op{true, 1, m.ranks[0].arena + kOffHeap, 1024, true}}, {Op{false, 0, m.ranks[1].arena + kOffHeap, 4096, true}}}, 16, 0);
    CHECK(m.ranks[0].fault.code == 4, "size mismatch: sender code=%u", m.ranks[0].fault.code);
    CHECK(m.ranks[1].fault.code == 3, "size mismatch: receiver code=%u (times out waiting for the rest)", m.ranks[1].fault.code);
  }
}

int main(int argc, char** argv) {
  bool quick = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--quick")) quick = true;
    else if (!strcmp(argv[i], "--threads") && i + 1 < argc) g_threads_per_cta = (unsigned)atoi(argv[++i]);
  }
  if (g_threads_per_cta < 1 || g_threads_per_cta > 8) { fprintf(stderr, "--threads 1..8\n"); return 2; }
  for (int n : {2, 4}) {
    scenario_ring(n, true, 0, 16);
    scenario_ring(n, false, 1u << 20, 16);       // 3 MiB + 5 B through 1 MiB windows: 4 chunks, both windows reused
    scenario_ring(n, false, 512, 2);             // tiny windows, two CTAs: thousands of chunk hand-shakes per message
    if (quick) break;
  }
  scenario_all_pairs(quick ? 3 : 8, 2);
  scenario_handover();
  scenario_faults();
  printf(g_failures ? "p2p_emu: %d FAILURES\n" : "p2p_emu: all scenarios passed\n", g_failures);
  return g_failures ? 1 : 0;
}
