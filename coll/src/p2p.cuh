// libb200coll point-to-point kernel. Included by kernels.cuh after device.cuh; also compiled for the HOST by coll/tests/p2p_emu.cc, which
// supplies stand-ins for the handful of primitives used here (flag loads / stores, 16-byte moves, globaltimer, record_fault) and runs
// every CTA as a thread — the protocol below is exercised under a real scheduler and ThreadSanitizer before it ever meets a GPU.
// Keep this file free of anything but those primitives, layout.h and plain C++.
#pragma once

namespace b200coll {

// ------------------------------------------------------------------------------------------------
// Point-to-point: one kernel carries every send and every recv of a group (at most one of each per peer), a few CTAs per
// operation, so a ring step or a pipeline hand-over is a single launch and a send can never starve the matching recv.
// Rendezvous per (pair, CTA), no all-rank barrier:
//   receiver CTA j   stores "write n bytes at offset off of my arena" into mailbox [me][j] of the SENDER's arena (two u64 words, each
//                    tagged with the chunk sequence number, two slots so that two windows can be outstanding);
//   sender CTA j     waits for the post with the next sequence number, pushes its 1/nb share of the n bytes straight into the
//                    receiver's memory (plain 16-byte stores to the peer mapping), then st.release.sys the sequence number into
//                    done [me][j] of the RECEIVER's arena;
//   receiver CTA j   waits for that number. A buffer inside the symmetric arena is the window itself (zero copy, one chunk, and a single
//                    receiver CTA whose thread j plays "CTA j" for the flags: there is nothing to move on this side);
//                    any other buffer is received through two staging windows that CTA j empties into it while the next
//                    chunk is already in flight.
// Sequence numbers live in device memory per (peer, CTA) and only ever grow, so a captured graph can be replayed; both sides
// derive the CTA count from the message size alone, so CTA j always meets CTA j. Every rank runs this same kernel whatever its
// role (see docs/protocols.md §5 on why different kernels per role deadlock virtual ranks).
struct P2pArgs {
  int nops, nsend;                               // operations [0, nsend) are sends, [nsend, nops) recvs
  int first_block[2 * kMaxRanks + 1];            // prefix sum of CTAs per operation
  int peer[2 * kMaxRanks];
  int staged[2 * kMaxRanks];                     // recv: 1 = two staging windows + copy-out, 0 = peers write the buffer itself
  int lanes[2 * kMaxRanks];                      // how many sender CTAs move this message (a function of its size). Sends and staged recvs run
                                                 // one CTA per lane; a direct recv has nothing to move, so ONE CTA serves all its lanes
  unsigned long long bytes[2 * kMaxRanks];       // message size; must be the same on both sides
  const char* src[2 * kMaxRanks];                // send: local source (any device memory)
  char* dst[2 * kMaxRanks];                      // recv: destination buffer
  unsigned long long win_off[2 * kMaxRanks];     // recv: arena offset of the (first) window
  unsigned long long win_bytes[2 * kMaxRanks];   // recv: bytes per window (>= bytes when not staged)
};
struct P2pShared { unsigned long long off, n; int bad; };
#ifndef P2P_SHARED                               // the host emulator maps this to per-CTA storage shared by the CTA's threads
#define P2P_SHARED(name) __shared__ P2pShared name
#endif
constexpr int kP2pThreads = 512;
constexpr unsigned long long kP2pValueMask = (1ull << kP2pValueBits) - 1;

// CTA `sub` of `nb` moves its share of a chunk of n bytes. Shares are fixed byte ranges of the WINDOW (`span` = size of the message's
// first chunk: the whole message when it is a single chunk, else the window size), clipped to n — not an even split of n: a staging
// window is reused every other chunk and flow control is per CTA, so CTA j must only ever touch the part of a window that CTA j of the
// other side has finished with, also when the last chunk is shorter than the others. (An even split of each chunk's n was the first
// version; the host emulator coll/tests/p2p_emu.cc caught it corrupting the tail of multi-chunk messages.)
// dst may be peer memory (send) or local (copy-out of a staging window); both pointers are 16-byte aligned.
// FRESH: src is a staging window a peer has just written and that this kernel has read before (two chunks ago): volatile loads, so
// that nothing an earlier read may have left in L1 can be returned.
template <bool FRESH>
__device__ __forceinline__ void p2p_move_share(char* dst, const char* src, unsigned long long n, unsigned long long span, int sub, int nb) {
  constexpr int U = 4;
  const unsigned long long wv = span / 16;
  unsigned long long b0 = wv * (unsigned)sub / (unsigned)nb * 16;
  unsigned long long b1 = sub == nb - 1 ? span : wv * (unsigned)(sub + 1) / (unsigned)nb * 16;
  if (b0 > n) b0 = n;
  if (b1 > n) b1 = n;
  const unsigned long long v0 = b0 / 16, v1 = b1 / 16;          // b0 is a multiple of 16 unless it was clipped to n (then the range is empty)
  for (unsigned long long base = v0 + threadIdx.x; base < v1; base += (unsigned long long)blockDim.x * U) {
    uint4 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const unsigned long long v = base + (unsigned long long)u * blockDim.x; if (v < v1) d[u] = FRESH ? ld_vec_volatile(src + v * 16) : ld_vec(src + v * 16); }
#pragma unroll
    for (int u = 0; u < U; u++) { const unsigned long long v = base + (unsigned long long)u * blockDim.x; if (v < v1) st_vec(dst + v * 16, d[u]); }
  }
  const unsigned long long t0 = v1 * 16 > b0 ? v1 * 16 : b0;    // odd bytes at the end of the chunk go with the CTA whose range they fall into
  for (unsigned long long e = t0 + threadIdx.x; e < b1; e += blockDim.x) dst[e] = FRESH ? *reinterpret_cast<const volatile char*>(src + e) : src[e];
}

__global__ void __launch_bounds__(kP2pThreads) k_p2p(COMM_PARAM, const __grid_constant__ P2pArgs a, uint32_t op) {
  pdl_prologue();
  P2P_SHARED(sh);                                // per-CTA: what thread 0 learnt from a mailbox, for the other threads
  int o = 0;
  while (o + 1 < a.nops && (int)blockIdx.x >= a.first_block[o + 1]) o++;
  const int sub = (int)blockIdx.x - a.first_block[o], nb = a.lanes[o];
  const int peer = a.peer[o];
  const unsigned long long total = a.bytes[o];
  char* my = c.peer[c.rank];
  if (o < a.nsend) {
    // ------------------------------------------------------------------ sender
    uint32_t* seqp = c.state + kP2pSendSeq0 + peer * kP2pMaxBlocks + sub;
    uint32_t seq = ld_volatile_u32(seqp);
    uint32_t* done = reinterpret_cast<uint32_t*>(c.peer[peer] + kOffP2pDone) + c.rank * kP2pMaxBlocks + sub;
    const char* src = a.src[o];
    unsigned long long sent = 0, span = 0;       // span: size of the first chunk = the receiver's window (or the whole message)
    while (sent < total) {
      seq++;
      if (threadIdx.x == 0) {
        const unsigned long long* box = reinterpret_cast<const unsigned long long*>(my + kOffP2pPost) + (((size_t)peer * kP2pMaxBlocks + sub) * 2 + (seq & 1u)) * 2;
        const unsigned long long tag = (unsigned long long)(seq & ((1u << kP2pTagBits) - 1)) << kP2pValueBits;
        unsigned long long wa = ld_relaxed_sys_u64(box), wb = ld_relaxed_sys_u64(box + 1);
        int bad = 0;
        if ((wa & ~kP2pValueMask) != tag || (wb & ~kP2pValueMask) != tag) {
          const unsigned long long t0 = globaltimer_ns();
          uint32_t spins = 0;
          for (;;) {
            wa = ld_relaxed_sys_u64(box); wb = ld_relaxed_sys_u64(box + 1);
            if ((wa & ~kP2pValueMask) == tag && (wb & ~kP2pValueMask) == tag) break;
            if (((++spins) & 0x3FF) == 0 && (c.fault->code != 0 || globaltimer_ns() - t0 > c.timeout_ns)) {
              record_fault(c, 3, (uint32_t)peer, seq, (uint32_t)(wb >> kP2pValueBits), op);      // the peer never posted this receive
              bad = 1;
              break;
            }
          }
        }
        (void)ld_acquire_sys_u64(box + 1);       // the receiver finished with the window's previous contents before it posted
        sh.off = (wa & kP2pValueMask) * 16; sh.n = wb & kP2pValueMask; sh.bad = bad;
      }
      __syncthreads();
      const unsigned long long off = sh.off, n = sh.n;
      const int bad = sh.bad;
      __syncthreads();                           // sh is rewritten in the next round
      if (bad) break;
      if (n == 0 || n > total - sent) {          // the two sides disagree about the message size
        if (threadIdx.x == 0) record_fault(c, 4, (uint32_t)peer, (uint32_t)(total - sent), (uint32_t)n, op);
        break;
      }
      if (sent == 0) span = n;
      if (n > span) {                            // later chunks never exceed the first one
        if (threadIdx.x == 0) record_fault(c, 4, (uint32_t)peer, (uint32_t)span, (uint32_t)n, op);
        break;
      }
      p2p_move_share<false>(c.peer[peer] + off, src + sent, n, span, sub, nb);
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys(done, seq);
      sent += n;
    }
    if (threadIdx.x == 0) *seqp = seq;
  } else {
    // ------------------------------------------------------------------ receiver
    if (!a.staged[o]) {
      // The buffer is in the arena: every lane gets the same post ("the whole message goes at this offset") and this one CTA then waits
      // for each lane's done flag — thread j looks after lane j (a loop only so that the host emulation can run with fewer threads).
      const unsigned long long off16 = a.win_off[o] / 16;
      for (int j = (int)threadIdx.x; j < nb; j += (int)blockDim.x) {
        const uint32_t seq = ld_volatile_u32(c.state + kP2pRecvSeq0 + peer * kP2pMaxBlocks + j) + 1;
        const unsigned long long tag = (unsigned long long)(seq & ((1u << kP2pTagBits) - 1)) << kP2pValueBits;
        unsigned long long* b = reinterpret_cast<unsigned long long*>(c.peer[peer] + kOffP2pPost) + (((size_t)c.rank * kP2pMaxBlocks + j) * 2 + (seq & 1u)) * 2;
        st_relaxed_sys_u64(b, tag | off16);
        st_release_sys_u64(b + 1, tag | total);
      }
      for (int j = (int)threadIdx.x; j < nb; j += (int)blockDim.x) {
        uint32_t* seqp = c.state + kP2pRecvSeq0 + peer * kP2pMaxBlocks + j;
        const uint32_t want = ld_volatile_u32(seqp) + 1;
        const uint32_t* done = reinterpret_cast<const uint32_t*>(my + kOffP2pDone) + peer * kP2pMaxBlocks + j;
        uint32_t v = ld_relaxed_sys(done);
        if ((int32_t)(v - want) < 0) {
          const unsigned long long t0 = globaltimer_ns();
          uint32_t spins = 0;
          while ((int32_t)((v = ld_relaxed_sys(done)) - want) < 0) {
            if (((++spins) & 0x3FF) == 0 && (c.fault->code != 0 || globaltimer_ns() - t0 > c.timeout_ns)) {
              record_fault(c, 3, (uint32_t)peer, want, v, op);        // the peer never sent (or died half-way)
              break;
            }
          }
        }
        (void)ld_acquire_sys(done);
        *seqp = want;
      }
      return;
    }
    uint32_t* seqp = c.state + kP2pRecvSeq0 + peer * kP2pMaxBlocks + sub;
    const uint32_t base = ld_volatile_u32(seqp);
    const unsigned long long win = a.win_bytes[o];
    const unsigned long long chunks = total <= win ? 1 : (total + win - 1) / win;
    const int staged = 1;                        // direct receives returned above
    unsigned long long* box = reinterpret_cast<unsigned long long*>(c.peer[peer] + kOffP2pPost) + ((size_t)c.rank * kP2pMaxBlocks + sub) * 4;
    const uint32_t* done = reinterpret_cast<const uint32_t*>(my + kOffP2pDone) + peer * kP2pMaxBlocks + sub;
    auto post = [&](unsigned long long k) {      // thread 0: chunk k may now be written into window k & 1
      const uint32_t seq = base + 1 + (uint32_t)k;
      const unsigned long long tag = (unsigned long long)(seq & ((1u << kP2pTagBits) - 1)) << kP2pValueBits;
      const unsigned long long off = a.win_off[o] + (staged ? (k & 1) * win : 0);
      const unsigned long long n = chunks == 1 ? total : (total - k * win < win ? total - k * win : win);
      unsigned long long* b = box + (seq & 1u) * 2;
      st_relaxed_sys_u64(b, tag | (off / 16));
      st_release_sys_u64(b + 1, tag | n);        // release: my reads of this window (copy-out of chunk k-2) are complete
    };
    if (threadIdx.x == 0) { post(0); if (chunks > 1) post(1); }
    for (unsigned long long k = 0; k < chunks; k++) {
      if (threadIdx.x == 0) {
        const uint32_t want = base + 1 + (uint32_t)k;
        int bad = 0;
        uint32_t v = ld_relaxed_sys(done);
        if ((int32_t)(v - want) < 0) {
          const unsigned long long t0 = globaltimer_ns();
          uint32_t spins = 0;
          while ((int32_t)((v = ld_relaxed_sys(done)) - want) < 0) {
            if (((++spins) & 0x3FF) == 0 && (c.fault->code != 0 || globaltimer_ns() - t0 > c.timeout_ns)) {
              record_fault(c, 3, (uint32_t)peer, want, v, op);        // the peer never sent (or died half-way)
              bad = 1;
              break;
            }
          }
        }
        (void)ld_acquire_sys(done);
        sh.bad = bad;
      }
      __syncthreads();
      const int bad = sh.bad;
      __syncthreads();
      if (bad) break;
      if (staged) {
        const unsigned long long n = chunks == 1 ? total : (total - k * win < win ? total - k * win : win);
        p2p_move_share<true>(a.dst[o] + k * win, my + a.win_off[o] + (k & 1) * win, n, chunks == 1 ? total : win, sub, nb);
        __syncthreads();
        if (threadIdx.x == 0 && k + 2 < chunks) post(k + 2);
      }
    }
    if (threadIdx.x == 0) *seqp = base + (uint32_t)chunks;
  }
}

}  // namespace b200coll
