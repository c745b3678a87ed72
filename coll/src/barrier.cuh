// Epoch barrier between the namesake CTAs of all ranks, the launch-sequence bookkeeping, and the one-instruction multicast flavour
// (CommDev::mcbar). Included by device.cuh after the flag primitives; also compiled for the HOST by coll/tests/barrier_emu.cc (CTAs as
// thread groups, arenas in host memory, ThreadSanitizer) — keep this file to those primitives, layout.h and plain C++.
#pragma once

namespace b200coll {

// All blocks read the launch sequence number; the last block to finish bumps it (graph-replay safe:
// nothing about the epoch lives in kernel arguments).
__device__ __forceinline__ uint32_t load_seq(const CommDev& c, int which) { return ld_volatile_u32(c.state + which); }

__device__ __forceinline__ bool last_block_ticket(const CommDev& c) {
  // call from one thread per block after the block's work is complete
  if (gridDim.x == 1) return true;
  __threadfence();
  uint32_t t = atomicAdd(c.state + kTicket, 1u);
  if (t == gridDim.x - 1) { c.state[kTicket] = 0; return true; }
  return false;
}

// Block b of every rank meets block b of every other rank. RELEASE=true publishes this block's prior
// writes (local or peer) system-wide before signalling; the wait side is always an acquire.

// Multicast barrier (CommDev::mcbar, B200COLL_MCBAR=0 turns it off): with NVLS, a barrier is ONE multimem.red on a multicast
// counter (the switch adds 1 to block b's counter on every rank) and ONE polled local word, instead of N remote stores + N polled flags.
// Counters grow by nranks per barrier and are never reset; how many barriers block b has taken is kept next to them (only block b of
// this rank touches that word, kernels of a communicator are stream-ordered). Both arrays live in the zero-initialised part of the flag
// megabyte that the flag matrix does not use. Without a multicast mapping the classic flag exchange below is used.
constexpr size_t kOffMcCounter = 64 << 10;    // u32[kMaxBlocks], multicast-addressed
constexpr size_t kOffMcCalls = 128 << 10;     // u32[kMaxBlocks], local bookkeeping
template <bool RELEASE>
__device__ __forceinline__ int barrier_blocks_mc(const CommDev& c, uint32_t op) {      // -1: no multicast mapping, use the flag exchange; 1 ok; 0 watchdog
  if (c.mc == nullptr) return -1;
  __syncthreads();
  int bad = 0;
  if (threadIdx.x == 0) {
    uint32_t* calls = reinterpret_cast<uint32_t*>(c.peer[c.rank] + kOffMcCalls) + blockIdx.x;
    const uint32_t k = *calls + 1;
    *calls = k;
    char* counter_mc = c.mc + kOffMcCounter + 4 * (size_t)blockIdx.x;
    if (RELEASE) mc_red_add_release_u32(counter_mc, 1u); else mc_red_add_relaxed_u32(counter_mc, 1u);
    const uint32_t want = k * (uint32_t)c.nranks;
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(c.peer[c.rank] + kOffMcCounter) + blockIdx.x;
    uint32_t v = ld_relaxed_sys(mine);
    if ((int32_t)(v - want) < 0) {
      const unsigned long long t0 = globaltimer_ns();
      uint32_t spins = 0;
      while ((int32_t)((v = ld_relaxed_sys(mine)) - want) < 0) {
        if (((++spins) & 0x3FF) == 0) {
          if (c.fault->code != 0 || globaltimer_ns() - t0 > c.timeout_ns) { record_fault(c, 1, 0xFFu, want, v, op); bad = 1; break; }
        }
      }
    }
    (void)ld_acquire_sys(mine);
  }
  return __syncthreads_or(bad) ? 0 : 1;
}

// Returns false when the watchdog fired (a peer never arrived): the caller must not touch peer data and returns at once — the
// communicator is poisoned, the host reports b200collRemoteError, and nothing unsynchronised is read or written.
template <bool RELEASE>
__device__ __forceinline__ bool barrier_blocks(const CommDev& c, uint32_t epoch, uint32_t op) {
  if (c.mcbar) { const int r = barrier_blocks_mc<RELEASE>(c, op); if (r >= 0) return r != 0; }
  __syncthreads();
  const int t = threadIdx.x;
  int bad = 0;
  if (t < c.nranks) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(c.peer[t] + kOffFlags) + (size_t)blockIdx.x * kMaxRanks + c.rank;
    if (RELEASE) st_release_sys(remote, epoch); else st_relaxed_sys(remote, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(c.peer[c.rank] + kOffFlags) + (size_t)blockIdx.x * kMaxRanks + t;
    // Poll with relaxed loads, then a single load-acquire of the same flag: acquire ordering for the data that
    // follows without a full fence (a fence.sys here would also wait for my own flag store's round trip).
    uint32_t v = ld_relaxed_sys(mine);
    if ((int32_t)(v - epoch) < 0) {
      const unsigned long long t0 = globaltimer_ns();
      uint32_t spins = 0;
      while ((int32_t)((v = ld_relaxed_sys(mine)) - epoch) < 0) {
        if (((++spins) & 0x3FF) == 0) {
          if (c.fault->code != 0 || globaltimer_ns() - t0 > c.timeout_ns) { record_fault(c, 1, t, epoch, v, op); bad = 1; break; }
        }
      }
    }
    (void)ld_acquire_sys(mine);
  }
  return __syncthreads_or(bad) == 0;
}

}  // namespace b200coll
