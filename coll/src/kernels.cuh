// libb200coll kernels (sm_100a). One kernel per (algorithm); each is templated on <InT, OutT> so the
// cast/scale epilogue runs in registers between the reduction and the (peer / multicast) store —
// no separate elementwise kernel or extra HBM pass on any path (BASELINE.json north-star).
//
//   k_copy_scale      nranks==1                    out = cast(in*scale)
//   k_ll              Lamport flag-in-payload      all four ops, zero barriers, <= 512 KiB per source
//   k_pull_reduce     one-shot AR / P2P or NVLS RS barrier, pull (or multimem.ld_reduce), reduce, local store
//   k_ar_twoshot      P2P all-reduce               pull+reduce own slice, push result to every peer (single pass)
//   k_ar_nvls         NVLS all-reduce              multimem.ld_reduce own slice -> scale/cast -> multimem.st
//   k_ag_push         all-gather                   P2P push or one multimem.st per vector
//   k_a2av_push       all-to-all(v)                per-peer row ranges, flattened for load balance
//   k_bcast           broadcast                    root pushes (P2P or one multimem.st per vector); the others only synchronise
//   k_reduce_root     reduce                       the root pulls (or multimem.ld_reduce) the whole buffer; the others only synchronise
//   k_p2p             send / recv (grouped)        receiver posts where to write, sender pushes over NVLink and signals; no all-rank barrier
//   k_barrier
#pragma once
#include "device.cuh"

namespace b200coll {

constexpr int kThreads = 512;

template <typename InT, typename OutT, int E>
__device__ __forceinline__ void finish_store_local(OutT* dst, const float* acc, float scale) {
  float f[E];
#pragma unroll
  for (int i = 0; i < E; i++) f[i] = acc[i] * scale;
  uint32_t w[Pack<OutT, E>::W];
  Pack<OutT, E>::run(f, w);
  st_words<Pack<OutT, E>::W>(dst, w);
}

// ------------------------------------------------------------------------------------------------
template <typename InT, typename OutT>
__global__ void __launch_bounds__(kThreads) k_copy_scale(const InT* __restrict__ in, OutT* __restrict__ out, size_t count, float scale) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int U = 4;                       // 4 x 16 B loads in flight per thread: HBM-latency hiding for a pure streaming kernel
  const size_t nvec = count / E;
  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < nvec; base += stride) {
    uint4 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const size_t v = base + (size_t)u * blockDim.x; if (v < nvec) d[u] = ld_vec(in + v * E); }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t v = base + (size_t)u * blockDim.x;
      if (v < nvec) { float acc[E] = {}; unpack_add<InT>(acc, d[u]); finish_store_local<InT, OutT, E>(out + v * E, acc, scale); }
    }
  }
  if (blockIdx.x == 0) {
    const size_t e = nvec * E + threadIdx.x;
    if (e < count) out[e] = from_float<OutT>(to_float<InT>(in[e]) * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// Lamport path. Scratch slot (buf, src, i) of MY arena is written by rank `src` (or by the switch on
// its behalf) and read only by me. Three buffers rotate per launch: use k%3, clear (k-1)%3 (last
// touched one launch ago, next written two launches from now), (k+1)%3 was cleared by the previous
// launch. A slot is "empty" while any of its four words equals kLLSentinel.
__device__ __forceinline__ char* ll_slot(char* base, uint32_t buf, int src, size_t i) {
  return base + kOffLL + (((size_t)buf * kMaxRanks + src) * kLLMaxVecs + i) * 16;
}
__device__ __forceinline__ uint4 ll_sanitize(uint4 v) {
  if (v.x == kLLSentinel) v.x = kLLSanitized;
  if (v.y == kLLSentinel) v.y = kLLSanitized;
  if (v.z == kLLSentinel) v.z = kLLSanitized;
  if (v.w == kLLSentinel) v.w = kLLSanitized;
  return v;
}
__device__ __forceinline__ bool ll_ready(const uint4& v) {
  return v.x != kLLSentinel && v.y != kLLSentinel && v.z != kLLSentinel && v.w != kLLSentinel;
}
__device__ __forceinline__ uint4 ll_wait(const CommDev& c, const char* slot, int src, uint32_t op) {
  uint4 v = ld_vec_volatile(slot);
  if (!ll_ready(v)) {
    const unsigned long long t0 = globaltimer_ns();
    uint32_t spins = 0;
    while (true) {
      v = ld_vec_volatile(slot);
      if (ll_ready(v)) break;
      if (((++spins) & 0x3FF) == 0) {
        if (c.fault->code != 0 || globaltimer_ns() - t0 > c.timeout_ns) { record_fault(c, 2, src, 0, v.x, op); break; }
      }
    }
  }
  return v;
}
template <typename InT, int E>
__device__ __forceinline__ uint4 load_in_guarded(const InT* in, size_t i, size_t count) {
  if ((i + 1) * E <= count) return ld_vec(in + i * E);
  union { InT e[E]; uint4 v; } u;
  u.v = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < E; k++) if (i * E + k < count) u.e[k] = in[i * E + k];
  return u.v;
}
template <typename OutT, int E>
__device__ __forceinline__ void store_out_guarded(OutT* out, size_t i, size_t count, const float* acc, float scale) {
  if ((i + 1) * E <= count) {
    float f[E];
#pragma unroll
    for (int k = 0; k < E; k++) f[k] = acc[k] * scale;
    uint32_t w[Pack<OutT, E>::W];
    Pack<OutT, E>::run(f, w);
    st_words<Pack<OutT, E>::W>(out + i * E, w);
  } else {
#pragma unroll
    for (int k = 0; k < E; k++) if (i * E + k < count) out[i * E + k] = from_float<OutT>(acc[k] * scale);
  }
}

// SLICED=false: every peer gets my whole `count`-element input (AR, AG). SLICED=true: peer p gets block p (RS, A2A).
// SUM=true: out[i] = sum over sources (AR, RS). SUM=false: out[src*count + i] = source's data (AG, A2A).
template <typename InT, typename OutT, bool SLICED, bool SUM, bool MC>
__global__ void __launch_bounds__(kThreads) k_ll(COMM_PARAM, const InT* __restrict__ in, OutT* __restrict__ out, size_t count, float scale, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  const uint32_t k = load_seq(c, kSeqLL);
  const uint32_t buf = k % 3, prev = (k + 2) % 3;
  const size_t nv = (count + E - 1) / E;
  const size_t used_prev = c.state[kLLUsedLo0 + prev], used_prev_hi = c.state[kLLUsedHi0 + prev];
  size_t span = nv > used_prev ? nv : used_prev;
  if (used_prev_hi > span) span = used_prev_hi;
  char* const me = c.peer[c.rank];
  const uint4 empty = make_uint4(kLLSentinel, kLLSentinel, kLLSentinel, kLLSentinel);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < span; i += (size_t)gridDim.x * blockDim.x) {
    uint4 own = make_uint4(0, 0, 0, 0);
    if (i < nv) {
      if (!SLICED) {
        own = ll_sanitize(load_in_guarded<InT, E>(in, i, count));
        if (MC) {
          const uint32_t w[4] = {own.x, own.y, own.z, own.w};
          mc_st_words(ll_slot(c.mc, buf, c.rank, i), w, 4);
        } else {
#pragma unroll
          for (int j = 1; j < kMaxRanks; j++) if (j < c.nranks) {
            int p = c.rank + j; if (p >= c.nranks) p -= c.nranks;
            st_vec_volatile(ll_slot(c.peer[p], buf, c.rank, i), own);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
          int p = c.rank + j; if (p >= c.nranks) p -= c.nranks;
          uint4 v = ll_sanitize(ld_vec(in + (size_t)p * count + i * E));
          if (j == 0) own = v; else st_vec_volatile(ll_slot(c.peer[p], buf, c.rank, i), v);
        }
      }
    }
    if (i < used_prev) {
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) st_vec(ll_slot(me, prev, r, i), empty);
    }
    if (i < used_prev_hi) {
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) st_vec(ll_slot(me, prev, r, kLLHalfVecs + i), empty);
    }
    if (i < nv) {
      if (SUM) {
        float acc[E] = {};
#pragma unroll
        for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) {
          uint4 d = (r == c.rank) ? own : ll_wait(c, ll_slot(me, buf, r, i), r, op);
          unpack_add<InT>(acc, d);
        }
        store_out_guarded<OutT, E>(out, i, count, acc, scale);
      } else {
#pragma unroll
        for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) {
          uint4 d = (r == c.rank) ? own : ll_wait(c, ll_slot(me, buf, r, i), r, op);
          float acc[E] = {};
          unpack_add<InT>(acc, d);
          store_out_guarded<OutT, E>(out + (size_t)r * count, i, count, acc, scale);
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && last_block_ticket(c)) {
    c.state[kLLUsedLo0 + buf] = (uint32_t)nv;
    c.state[kLLUsedHi0 + buf] = 0;
    c.state[kSeqLL] = k + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Two-shot Lamport all-reduce: zero barriers like k_ll, but each rank receives 2S instead of N*S bytes.
// Phase 1 (reduce-scatter): rank s pushes vector i of slice p into peer p's lo slot (buf, s, i); rank p sums the
// N copies of its slice in fp32, applies scale/cast. Phase 2 (all-gather): p publishes the finished vector into
// everybody's hi slot (buf, p, i) — one multimem.st when MC — and every rank copies the N slices into out.
// One thread owns vector i of *every* slice end to end, so there is no intra-kernel dependency between threads
// and in-place is safe (a thread reads all its inputs before its first write). Needs sizeof(OutT) == sizeof(InT).
template <typename OutT, int E>
__device__ __forceinline__ void store_raw_guarded(OutT* out, size_t vec, size_t count, const uint4& w) {
  if ((vec + 1) * E <= count) { st_vec(out + vec * E, w); return; }
  union { OutT e[E]; uint4 v; } u;
  u.v = w;
#pragma unroll
  for (int k = 0; k < E; k++) if (vec * E + k < count) out[vec * E + k] = u.e[k];
}

template <typename InT, typename OutT, bool MC>
__global__ void __launch_bounds__(kThreads) k_ll_twoshot(COMM_PARAM, const InT* __restrict__ in, OutT* __restrict__ out, size_t count, float scale, uint32_t op) {
  pdl_prologue();
  static_assert(sizeof(InT) == sizeof(OutT), "two-shot LL keeps one vector geometry for both phases");
  constexpr int E = Epv<InT>::value;
  const uint32_t k = load_seq(c, kSeqLL);
  const uint32_t buf = k % 3, prev = (k + 2) % 3;
  const size_t nvec = (count + E - 1) / E;
  const size_t nslice = (nvec + c.nranks - 1) / c.nranks;
  const size_t used_lo = c.state[kLLUsedLo0 + prev], used_hi = c.state[kLLUsedHi0 + prev];
  size_t span = nslice > used_lo ? nslice : used_lo;
  if (used_hi > span) span = used_hi;
  char* const me = c.peer[c.rank];
  const uint4 empty = make_uint4(kLLSentinel, kLLSentinel, kLLSentinel, kLLSentinel);
  const size_t my0 = (size_t)c.rank * nslice;
  const size_t my_len = my0 >= nvec ? 0 : (nvec - my0 < nslice ? nvec - my0 : nslice);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < span; i += (size_t)gridDim.x * blockDim.x) {
    uint4 own = make_uint4(0, 0, 0, 0);
    if (i < nslice) {
#pragma unroll
      for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
        int p = c.rank + j; if (p >= c.nranks) p -= c.nranks;
        const size_t v = (size_t)p * nslice + i;
        if (v < nvec) {
          uint4 d = ll_sanitize(load_in_guarded<InT, E>(in, v, count));
          if (j == 0) own = d; else st_vec_volatile(ll_slot(c.peer[p], buf, c.rank, i), d);
        }
      }
    }
    if (i < used_lo) {
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) st_vec(ll_slot(me, prev, r, i), empty);
    }
    if (i < used_hi) {
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) st_vec(ll_slot(me, prev, r, kLLHalfVecs + i), empty);
    }
    if (i < my_len) {
      float acc[E] = {};
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks) {
        uint4 d = (r == c.rank) ? own : ll_wait(c, ll_slot(me, buf, r, i), r, op);
        unpack_add<InT>(acc, d);
      }
#pragma unroll
      for (int e = 0; e < E; e++) acc[e] *= scale;
      uint32_t w[4];
      Pack<OutT, E>::run(acc, w);
      const uint4 red = ll_sanitize(make_uint4(w[0], w[1], w[2], w[3]));
      if (MC) {
        const uint32_t rw[4] = {red.x, red.y, red.z, red.w};
        mc_st_words(ll_slot(c.mc, buf, c.rank, kLLHalfVecs + i), rw, 4);
      } else {
#pragma unroll
        for (int j = 1; j < kMaxRanks; j++) if (j < c.nranks) {
          int p = c.rank + j; if (p >= c.nranks) p -= c.nranks;
          st_vec_volatile(ll_slot(c.peer[p], buf, c.rank, kLLHalfVecs + i), red);
        }
      }
      store_raw_guarded<OutT, E>(out, my0 + i, count, red);
    }
    if (i < nslice) {
#pragma unroll
      for (int r = 0; r < kMaxRanks; r++) if (r < c.nranks && r != c.rank) {
        const size_t v = (size_t)r * nslice + i;
        if (v < nvec) store_raw_guarded<OutT, E>(out, v, count, ll_wait(c, ll_slot(me, buf, r, kLLHalfVecs + i), r, op));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && last_block_ticket(c)) {
    c.state[kLLUsedLo0 + buf] = (uint32_t)nslice;
    c.state[kLLUsedHi0 + buf] = (uint32_t)nslice;
    c.state[kSeqLL] = k + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Scalar tail for the barrier-based all-reduce kernels: rank 0 reduces elements [e0, count) from every
// peer and writes the result into every peer's out (so in-place is safe).
template <typename InT, typename OutT>
__device__ __forceinline__ void ar_tail_rank0(const CommDev& c, size_t in_off, size_t out_off, size_t e0, size_t count, float scale) {
  if (c.rank != 0 || blockIdx.x != 0) return;
  const size_t e = e0 + threadIdx.x;
  if (e >= count) return;
  float acc = 0.f;
  for (int r = 0; r < c.nranks; r++) acc += to_float<InT>(reinterpret_cast<const volatile InT*>(c.peer[r] + in_off)[e]);
  const OutT o = from_float<OutT>(acc * scale);
  for (int r = 0; r < c.nranks; r++) reinterpret_cast<OutT*>(c.peer[r] + out_off)[e] = o;
}

// out_local[i] = scale * sum_r peer_r[in_off + i]   (one-shot all-reduce; reduce-scatter with in_off pointing at my slice)
// FIXED_ORDER: sum in rank order so every rank computes bit-identical results (one-shot AR).
template <typename InT, typename OutT, bool FIXED_ORDER, bool MC>
__global__ void __launch_bounds__(kThreads) k_pull_reduce(COMM_PARAM, size_t in_off, OutT* __restrict__ out, size_t count, float scale, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int U = 2;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  const size_t nvec = count / E;
  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < nvec; base += stride) {
    if (MC) {
      uint4 d[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const size_t v = base + (size_t)u * blockDim.x; if (v < nvec) d[u] = mc_ld_reduce<InT>(c.mc + in_off + v * 16); }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = base + (size_t)u * blockDim.x;
        if (v < nvec) { float acc[E] = {}; unpack_add<InT>(acc, d[u]); finish_store_local<InT, OutT, E>(out + v * E, acc, scale); }
      }
    } else {
      uint4 d[U][kMaxRanks];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = base + (size_t)u * blockDim.x;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
            int r = j;
            if (!FIXED_ORDER) { r = c.rank + j; if (r >= c.nranks) r -= c.nranks; }
            d[u][j] = ld_vec(c.peer[r] + in_off + v * 16);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = base + (size_t)u * blockDim.x;
        if (v < nvec) {
          float acc[E] = {};
#pragma unroll
          for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) unpack_add<InT>(acc, d[u][j]);
          finish_store_local<InT, OutT, E>(out + v * E, acc, scale);
        }
      }
    }
  }
  if (blockIdx.x == 0) {   // scalar tail, every rank for itself (out is local and distinct from in for these ops)
    const size_t e = nvec * E + threadIdx.x;
    if (e < count) {
      float acc = 0.f;
      for (int r = 0; r < c.nranks; r++) acc += to_float<InT>(reinterpret_cast<const volatile InT*>(c.peer[r] + in_off)[e]);
      out[e] = from_float<OutT>(acc * scale);
    }
  }
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

// ------------------------------------------------------------------------------------------------
// Two-shot fused into one pass: rank r owns vector slice r; pulls it from every peer, reduces in fp32,
// applies scale/cast and pushes the finished vector into every peer's out. Bytes per GPU and direction:
// S(N-1)/N pulled + S(N-1)/N pushed -> bus bandwidth bound = link bandwidth.
template <typename InT, typename OutT>
__global__ void __launch_bounds__(kThreads) k_ar_twoshot(COMM_PARAM, size_t in_off, size_t out_off, size_t count, float scale, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int W = Pack<OutT, E>::W;
  constexpr int U = 2;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  const size_t nvec = count / E;
  const size_t v0 = nvec * c.rank / c.nranks, v1 = nvec * (c.rank + 1) / c.nranks;
  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = v0 + (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < v1; base += stride) {
    uint4 d[U][kMaxRanks];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t v = base + (size_t)u * blockDim.x;
      if (v < v1) {
#pragma unroll
        for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
          int r = c.rank + j; if (r >= c.nranks) r -= c.nranks;
          d[u][j] = ld_vec(c.peer[r] + in_off + v * 16);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t v = base + (size_t)u * blockDim.x;
      if (v < v1) {
        float acc[E] = {};
#pragma unroll
        for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) unpack_add<InT>(acc, d[u][j]);
#pragma unroll
        for (int i = 0; i < E; i++) acc[i] *= scale;
        uint32_t w[W];
        Pack<OutT, E>::run(acc, w);
#pragma unroll
        for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
          int r = c.rank + j; if (r >= c.nranks) r -= c.nranks;
          st_words<W>(c.peer[r] + out_off + v * (E * sizeof(OutT)), w);
        }
      }
    }
  }
  ar_tail_rank0<InT, OutT>(c, in_off, out_off, nvec * E, count, scale);
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

// NVLS all-reduce: the switch does the reduction (fp32 accumulate) and the broadcast.
// Per GPU and direction: ~S(1+1/N) bytes -> algbw bound = link/(1+1/N).
// U = vectors in flight per thread. Large messages use U=4; mid sizes use U=1 so that a slice takes several passes and
// the multimem.st of pass k (ingress-heavy) overlaps the ld_reduce of pass k+1 (egress-heavy) instead of running back to back.
template <typename InT, typename OutT, int U>
__global__ void __launch_bounds__(kThreads) k_ar_nvls(COMM_PARAM, size_t in_off, size_t out_off, size_t count, float scale, int identity, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int W = Pack<OutT, E>::W;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  const size_t nvec = count / E;
  const size_t v0 = nvec * c.rank / c.nranks, v1 = nvec * (c.rank + 1) / c.nranks;
  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = v0 + (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < v1; base += stride) {
    uint4 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const size_t v = base + (size_t)u * blockDim.x; if (v < v1) d[u] = mc_ld_reduce<InT>(c.mc + in_off + v * 16); }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t v = base + (size_t)u * blockDim.x;
      if (v < v1) {
        uint32_t w[W];
        if (identity) {   // same dtype, scale == 1: forward the switch's result untouched
          const uint32_t raw[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
          for (int i = 0; i < W && i < 4; i++) w[i] = raw[i];
        } else {
          float acc[E] = {};
          unpack_add<InT>(acc, d[u]);
#pragma unroll
          for (int i = 0; i < E; i++) acc[i] *= scale;
          Pack<OutT, E>::run(acc, w);
        }
        mc_st_wordsW<W>(c.mc + out_off + v * (E * sizeof(OutT)), w);
      }
    }
  }
  ar_tail_rank0<InT, OutT>(c, in_off, out_off, nvec * E, count, scale);
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

// ------------------------------------------------------------------------------------------------
// All-gather push: my `count` elements land at element offset rank*count of every peer's out.
// MC: one multimem.st per vector (egress S/N instead of S(N-1)/N).
template <typename InT, typename OutT, bool MC>
__global__ void __launch_bounds__(kThreads) k_ag_push(COMM_PARAM, const InT* __restrict__ in, size_t out_off, size_t count, float scale, int identity, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int W = Pack<OutT, E>::W;
  constexpr int U = 4;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  const size_t nvec = count / E;
  const size_t dst0 = out_off + (size_t)c.rank * count * sizeof(OutT);
  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < nvec; base += stride) {
    uint4 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const size_t v = base + (size_t)u * blockDim.x; if (v < nvec) d[u] = ld_vec(in + v * E); }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t v = base + (size_t)u * blockDim.x;
      if (v < nvec) {
        uint32_t w[W];
        if (identity) {
          const uint32_t raw[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
          for (int i = 0; i < W && i < 4; i++) w[i] = raw[i];
        } else {
          float acc[E] = {};
          unpack_add<InT>(acc, d[u]);
#pragma unroll
          for (int i = 0; i < E; i++) acc[i] *= scale;
          Pack<OutT, E>::run(acc, w);
        }
        const size_t off = dst0 + v * (E * sizeof(OutT));
        if (MC) {
          mc_st_wordsW<W>(c.mc + off, w);
        } else {
#pragma unroll
          for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
            int r = c.rank + j; if (r >= c.nranks) r -= c.nranks;
            st_words<W>(c.peer[r] + off, w);
          }
        }
      }
    }
  }
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

// ------------------------------------------------------------------------------------------------
// All-to-all(v) push. Per destination p: nvec[p] input vectors starting at vector src_vec[p] of my send
// buffer go to output-vector offset dst_vec[p] (units of E*sizeof(OutT) bytes) of peer p's recv buffer.
// The (peer, vector) space is flattened with a prefix table so skewed expert loads still spread over all CTAs.
struct A2AvArgs {
  unsigned long long src_vec[kMaxRanks];
  unsigned long long dst_vec[kMaxRanks];
  unsigned long long prefix[kMaxRanks + 1];   // prefix over the staggered order j = 0..nranks-1 (peer = rank+1+j)
};
template <typename InT, typename OutT>
__global__ void __launch_bounds__(kThreads) k_a2av_push(COMM_PARAM, const InT* __restrict__ in, size_t out_off, A2AvArgs a, float scale, int identity, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int W = Pack<OutT, E>::W;
  constexpr int U = 4;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  const size_t total = a.prefix[c.nranks];
  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < total; base += stride) {
    uint4 d[U];
    int pj[U];
    size_t lv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t g = base + (size_t)u * blockDim.x;
      pj[u] = 0; lv[u] = 0;
      if (g < total) {
        int j = 0;
#pragma unroll
        for (int q = 1; q < kMaxRanks; q++) if (q < c.nranks && g >= a.prefix[q]) j = q;
        int p = c.rank + 1 + j; if (p >= c.nranks) p -= c.nranks;
        pj[u] = p; lv[u] = g - a.prefix[j];
        d[u] = ld_vec(in + (a.src_vec[p] + lv[u]) * E);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t g = base + (size_t)u * blockDim.x;
      if (g < total) {
        uint32_t w[W];
        if (identity) {
          const uint32_t raw[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
          for (int i = 0; i < W && i < 4; i++) w[i] = raw[i];
        } else {
          float acc[E] = {};
          unpack_add<InT>(acc, d[u]);
#pragma unroll
          for (int i = 0; i < E; i++) acc[i] *= scale;
          Pack<OutT, E>::run(acc, w);
        }
        st_words<W>(c.peer[pj[u]] + out_off + (a.dst_vec[pj[u]] + lv[u]) * (E * sizeof(OutT)), w);
      }
    }
  }
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

// ------------------------------------------------------------------------------------------------
// Broadcast: the root's `count` elements land at out_off of every rank (its own copy included, so in-place and
// out-of-place behave the same). MC: one multimem.st per vector -> root egress S instead of S(N-1).
// Every rank launches the same grid; non-root CTAs only take part in the two barriers.
template <typename InT, typename OutT, bool MC>
__global__ void __launch_bounds__(kThreads) k_bcast(COMM_PARAM, const InT* __restrict__ in, size_t out_off, size_t count, float scale, int identity, int root, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int W = Pack<OutT, E>::W;
  constexpr int U = 4;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);          // every rank is done with the previous contents of its out
  if (c.rank == root) {
    const size_t nvec = count / E;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < nvec; base += stride) {
      uint4 d[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const size_t v = base + (size_t)u * blockDim.x; if (v < nvec) d[u] = ld_vec(in + v * E); }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = base + (size_t)u * blockDim.x;
        if (v < nvec) {
          uint32_t w[W];
          if (identity) {
            const uint32_t raw[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
            for (int i = 0; i < W && i < 4; i++) w[i] = raw[i];
          } else {
            float acc[E] = {};
            unpack_add<InT>(acc, d[u]);
#pragma unroll
            for (int i = 0; i < E; i++) acc[i] *= scale;
            Pack<OutT, E>::run(acc, w);
          }
          const size_t off = out_off + v * (E * sizeof(OutT));
          if (MC) {
            mc_st_wordsW<W>(c.mc + off, w);
          } else {
#pragma unroll
            for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
              int r = c.rank + j; if (r >= c.nranks) r -= c.nranks;
              st_words<W>(c.peer[r] + off, w);
            }
          }
        }
      }
    }
    if (blockIdx.x == 0) {                          // scalar tail (count not a multiple of one vector)
      const size_t e = nvec * E + threadIdx.x;
      if (e < count) {
        OutT o = from_float<OutT>(to_float<InT>(in[e]) * scale);
        if constexpr (sizeof(InT) == sizeof(OutT)) {
          if (identity) { const InT x = in[e]; memcpy(&o, &x, sizeof(o)); }   // bit-exact, like the vector path (payload may not be a float at all)
        }
        for (int r = 0; r < c.nranks; r++) reinterpret_cast<OutT*>(c.peer[r] + out_off)[e] = o;
      }
    }
  }
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

// Rooted reduce: out@root[i] = scale * sum_r peer_r[in_off + i]. Every rank launches this same kernel with the same grid
// (a different kernel on the non-root ranks would need its own lazy module load, and with virtual ranks that load cannot
// complete while the root's kernel is already spinning on the GPU -> deadlock); non-root CTAs only take the two barriers:
// the first tells the root my send buffer is ready, the second tells me the root has finished reading it.
// MC: the root issues multimem.ld_reduce (the switch adds, the root receives S bytes instead of S(N-1)).
template <typename InT, typename OutT, bool MC>
__global__ void __launch_bounds__(kThreads) k_reduce_root(COMM_PARAM, size_t in_off, OutT* __restrict__ out, size_t count, float scale, int root, uint32_t op) {
  pdl_prologue();
  constexpr int E = Epv<InT>::value;
  constexpr int U = 2;
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  if (c.rank == root) {
    const size_t nvec = count / E;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t base = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < nvec; base += stride) {
      if (MC) {
        uint4 d[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t v = base + (size_t)u * blockDim.x; if (v < nvec) d[u] = mc_ld_reduce<InT>(c.mc + in_off + v * 16); }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = base + (size_t)u * blockDim.x;
          if (v < nvec) { float acc[E] = {}; unpack_add<InT>(acc, d[u]); finish_store_local<InT, OutT, E>(out + v * E, acc, scale); }
        }
      } else {
        uint4 d[U][kMaxRanks];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = base + (size_t)u * blockDim.x;
          if (v < nvec) {
#pragma unroll
            for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) d[u][j] = ld_vec(c.peer[j] + in_off + v * 16);   // rank order: deterministic sum
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = base + (size_t)u * blockDim.x;
          if (v < nvec) {
            float acc[E] = {};
#pragma unroll
            for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) unpack_add<InT>(acc, d[u][j]);
            finish_store_local<InT, OutT, E>(out + v * E, acc, scale);
          }
        }
      }
    }
    if (blockIdx.x == 0) {   // scalar tail
      const size_t e = nvec * E + threadIdx.x;
      if (e < count) {
        float acc = 0.f;
        for (int r = 0; r < c.nranks; r++) acc += to_float<InT>(reinterpret_cast<const volatile InT*>(c.peer[r] + in_off)[e]);
        out[e] = from_float<OutT>(acc * scale);
      }
    }
  }
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

#ifdef B200COLL_VARIANT_BULK
// ------------------------------------------------------------------------------------------------
// A/B candidate (make VARIANT=bulk -> lib/libb200coll_bulk.so): all-gather push with the copy engine of the SM instead of
// LDG/STG. One elected thread per CTA streams its share of the local buffer through a small shared-memory ring:
//   cp.async.bulk global -> shared (completion on an mbarrier), then one cp.async.bulk shared -> peer global per rank, committed as a
//   bulk group; a ring slot is reused once the group that read it has finished reading (cp.async.bulk.wait_group.read).
// Loads of chunk k+1 overlap the stores of chunk k; no registers carry data, so a CTA is one warp's worth of control code.
// Identity epilogue only (same dtype, scale 1): the bytes are never looked at.
constexpr int kBulkStages = 4;
constexpr uint32_t kBulkChunk = 8192;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k_ag_bulk(COMM_PARAM, const char* __restrict__ in, size_t out_off, size_t bytes, uint32_t op) {
  pdl_prologue();
  __shared__ __align__(128) char ring[kBulkStages][kBulkChunk];
  __shared__ __align__(8) unsigned long long full[kBulkStages];
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<false>(c, 2 * s + 1, op);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBulkStages; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t dst0 = out_off + (size_t)c.rank * bytes;
    const size_t nchunks = (bytes + kBulkChunk - 1) / kBulkChunk;
    uint32_t it = 0;
    for (size_t i = blockIdx.x; i < nchunks; i += gridDim.x, it++) {
      const int stage = (int)(it % kBulkStages);
      const uint32_t parity = (it / kBulkStages) & 1u;
      if (it >= (uint32_t)kBulkStages) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBulkStages - 1) : "memory");   // the stores that read this slot are done reading
      const size_t off = i * (size_t)kBulkChunk;
      const uint32_t n = (uint32_t)(bytes - off < kBulkChunk ? bytes - off : kBulkChunk);
      const uint32_t bar = smem_u32(&full[stage]), dst_s = smem_u32(&ring[stage][0]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s), "l"(in + off), "r"(n), "r"(bar) : "memory");
      uint32_t done = 0;
      while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#pragma unroll
      for (int j = 0; j < kMaxRanks; j++) if (j < c.nranks) {
        int r = c.rank + j; if (r >= c.nranks) r -= c.nranks;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(c.peer[r] + dst0 + off), "r"(dst_s), "r"(n) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // every store has been performed, not just read
    asm volatile("fence.proxy.async.global;" ::: "memory");             // order the copy engine's writes before the flag stores of the barrier
  }
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}
#endif

// ------------------------------------------------------------------------------------------------
// Point-to-point: one kernel carries every send and every recv of a group (at most one of each per peer), a few CTAs per
// operation, so a ring step or a pipeline hand-over is a single launch and a send can never starve the matching recv.
// Rendezvous per (pair, CTA), no all-rank barrier:
//   receiver CTA j   stores "write n bytes at offset off of my arena" into mailbox [me][j] of the SENDER's arena (two u64 words, each
//                    tagged with the chunk sequence number, two slots so that two windows can be outstanding);
//   sender CTA j     waits for the post with the next sequence number, pushes its 1/nb share of the n bytes straight into the
//                    receiver's memory (plain 16-byte stores to the peer mapping), then st.release.sys the sequence number into
//                    done [me][j] of the RECEIVER's arena;
//   receiver CTA j   waits for that number. A buffer inside the symmetric arena is the window itself (zero copy, one chunk);
//                    any other buffer is received through two staging windows that the CTA empties into it while the next
//                    chunk is already in flight.
// Sequence numbers live in device memory per (peer, CTA) and only ever grow, so a captured graph can be replayed; both sides
// derive the CTA count from the message size alone, so CTA j always meets CTA j. Every rank runs this same kernel whatever its
// role (see docs/protocols.md §5 on why different kernels per role deadlock virtual ranks).
struct P2pArgs {
  int nops, nsend;                               // operations [0, nsend) are sends, [nsend, nops) recvs
  int first_block[2 * kMaxRanks + 1];            // prefix sum of CTAs per operation
  int peer[2 * kMaxRanks];
  int staged[2 * kMaxRanks];                     // recv: 1 = two staging windows + copy-out, 0 = peers write the buffer itself
  unsigned long long bytes[2 * kMaxRanks];       // message size; must be the same on both sides
  const char* src[2 * kMaxRanks];                // send: local source (any device memory)
  char* dst[2 * kMaxRanks];                      // recv: destination buffer
  unsigned long long win_off[2 * kMaxRanks];     // recv: arena offset of the (first) window
  unsigned long long win_bytes[2 * kMaxRanks];   // recv: bytes per window (>= bytes when not staged)
};
constexpr int kP2pThreads = 512;
constexpr unsigned long long kP2pValueMask = (1ull << kP2pValueBits) - 1;

// CTA `sub` of `nb` moves its share of n bytes: whole 16-byte vectors split evenly, the odd bytes at the end go with the last CTA.
// dst may be peer memory (send) or local (copy-out of a staging window); both pointers are 16-byte aligned.
// FRESH: src is a staging window a peer has just written and that this kernel has read before (two chunks ago): volatile loads, so
// that nothing an earlier read may have left in L1 can be returned.
template <bool FRESH>
__device__ __forceinline__ void p2p_move_share(char* dst, const char* src, unsigned long long n, int sub, int nb) {
  constexpr int U = 4;
  const unsigned long long nvec = n / 16;
  const unsigned long long v0 = nvec * (unsigned)sub / (unsigned)nb, v1 = nvec * (unsigned)(sub + 1) / (unsigned)nb;
  for (unsigned long long base = v0 + threadIdx.x; base < v1; base += (unsigned long long)blockDim.x * U) {
    uint4 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const unsigned long long v = base + (unsigned long long)u * blockDim.x; if (v < v1) d[u] = FRESH ? ld_vec_volatile(src + v * 16) : ld_vec(src + v * 16); }
#pragma unroll
    for (int u = 0; u < U; u++) { const unsigned long long v = base + (unsigned long long)u * blockDim.x; if (v < v1) st_vec(dst + v * 16, d[u]); }
  }
  if (sub == nb - 1)
    for (unsigned long long e = nvec * 16 + threadIdx.x; e < n; e += blockDim.x) dst[e] = FRESH ? *reinterpret_cast<const volatile char*>(src + e) : src[e];
}

__global__ void __launch_bounds__(kP2pThreads) k_p2p(COMM_PARAM, const __grid_constant__ P2pArgs a, uint32_t op) {
  pdl_prologue();
  __shared__ unsigned long long sh_off, sh_n;
  __shared__ int sh_bad;
  int o = 0;
  while (o + 1 < a.nops && (int)blockIdx.x >= a.first_block[o + 1]) o++;
  const int sub = (int)blockIdx.x - a.first_block[o], nb = a.first_block[o + 1] - a.first_block[o];
  const int peer = a.peer[o];
  const unsigned long long total = a.bytes[o];
  char* my = c.peer[c.rank];
  if (o < a.nsend) {
    // ------------------------------------------------------------------ sender
    uint32_t* seqp = c.state + kP2pSendSeq0 + peer * kP2pMaxBlocks + sub;
    uint32_t seq = ld_volatile_u32(seqp);
    uint32_t* done = reinterpret_cast<uint32_t*>(c.peer[peer] + kOffP2pDone) + c.rank * kP2pMaxBlocks + sub;
    const char* src = a.src[o];
    unsigned long long sent = 0;
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
        sh_off = (wa & kP2pValueMask) * 16; sh_n = wb & kP2pValueMask; sh_bad = bad;
      }
      __syncthreads();
      const unsigned long long off = sh_off, n = sh_n;
      const int bad = sh_bad;
      __syncthreads();                           // sh_* are rewritten in the next round
      if (bad) break;
      if (n == 0 || n > total - sent) {          // the two sides disagree about the message size
        if (threadIdx.x == 0) record_fault(c, 4, (uint32_t)peer, (uint32_t)(total - sent), (uint32_t)n, op);
        break;
      }
      p2p_move_share<false>(c.peer[peer] + off, src + sent, n, sub, nb);
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys(done, seq);
      sent += n;
    }
    if (threadIdx.x == 0) *seqp = seq;
  } else {
    // ------------------------------------------------------------------ receiver
    uint32_t* seqp = c.state + kP2pRecvSeq0 + peer * kP2pMaxBlocks + sub;
    const uint32_t base = ld_volatile_u32(seqp);
    const unsigned long long win = a.win_bytes[o];
    const unsigned long long chunks = total <= win ? 1 : (total + win - 1) / win;
    const int staged = a.staged[o];
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
        sh_bad = bad;
      }
      __syncthreads();
      const int bad = sh_bad;
      __syncthreads();
      if (bad) break;
      if (staged) {
        const unsigned long long n = chunks == 1 ? total : (total - k * win < win ? total - k * win : win);
        p2p_move_share<true>(a.dst[o] + k * win, my + a.win_off[o] + (k & 1) * win, n, sub, nb);
        __syncthreads();
        if (threadIdx.x == 0 && k + 2 < chunks) post(k + 2);
      }
    }
    if (threadIdx.x == 0) *seqp = base + (uint32_t)chunks;
  }
}

__global__ void k_barrier(COMM_PARAM, uint32_t op) {
  pdl_prologue();
  const uint32_t s = load_seq(c, kSeqBarrier);
  barrier_blocks<true>(c, 2 * s + 2, op);
  if (threadIdx.x == 0 && last_block_ticket(c)) c.state[kSeqBarrier] = s + 1;
}

__global__ void k_fill_u32(uint32_t* p, size_t n, uint32_t v) {
  pdl_prologue();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace b200coll
