// libb200coll host path: NUMA placement of the rank, pinned host memory on the GPU's node, and the end-to-end all-reduce
// "pinned host -> GPU -> NVSwitch -> GPU -> pinned host" as ONE library call with the three legs overlapped.
//
// Why this is part of a transport library: the reference's benchmark is what a user sees when running *_perf in the pod
// (reference: gpudirect-rdma/nccl-test-a4x-max-jobset.yaml:141-153), and on an 8-GPU box the host legs are where the time goes:
//   * 8 ranks pulling 55 GB/s each out of host DRAM only scale when every rank's buffer sits on the socket its GPU hangs off
//     (GPUs 0-3 / 4-7 are on different sockets; four unplaced ranks crossing the socket link halve everybody's rate);
//   * copy, collective and copy-back run back to back on one stream unless somebody chunks them.
// b200collCommInitRank therefore binds the calling thread to the GPU-local CPUs (B200COLL_AFFINITY=0 opts out, =2 restores the
// previous mask after init), b200collHostAlloc returns registered memory on that node, and b200collAllReduceHost pipelines
// H2D(i+1) | all-reduce(i) | D2H(i-1) over two pairs of arena staging buffers. Small messages skip the copy engines altogether:
// the Lamport kernel reads the pinned input over PCIe and stores the result straight into pinned host memory (one launch).
#include <dirent.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <string>

#include "comm.h"

namespace b200coll {

static std::string read_small_file(const std::string& path) {
  std::string out;
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return out;
  char buf[4096];
  size_t n = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[n] = 0;
  out = buf;
  while (!out.empty() && (out.back() == '\n' || out.back() == ' ')) out.pop_back();
  return out;
}

// "0-31,64-95" -> cpu_set_t. Returns the number of CPUs set. Pure: unit-tested through b200collDebugParseCpuList.
int parse_cpulist(const char* s, cpu_set_t* set) {
  CPU_ZERO(set);
  int n = 0;
  while (s && *s) {
    char* end = nullptr;
    long a = strtol(s, &end, 10);
    if (end == s) break;
    long b = a;
    if (*end == '-') { s = end + 1; b = strtol(s, &end, 10); if (end == s) break; }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++) if (c >= 0 && !CPU_ISSET(c, set)) { CPU_SET(c, set); n++; }
    s = (*end == ',') ? end + 1 : end;
    if (*end != ',' ) break;
  }
  return n;
}

// sysfs directory of the GPU: /sys/bus/pci/devices/0000:1b:00.0 (B200COLL_SYSFS_PCI overrides the root for tests)
static std::string gpu_sysfs_dir(int device) {
  char bus[32] = {};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { (void)cudaGetLastError(); return ""; }
  for (char* p = bus; *p; p++) *p = (char)tolower(*p);
  const char* root = getenv("B200COLL_SYSFS_PCI");
  return std::string(root && *root ? root : "/sys/bus/pci/devices") + "/" + bus;
}

// The part of bind_to_gpu_numa that needs no GPU: read <dir>/numa_node and <dir>/local_cpulist, and (mode != 0) narrow the calling
// thread's affinity to the listed CPUs that its current mask allows. Unit-tested on CPU through b200collDebugApplyLocality.
static void apply_gpu_locality(b200collComm* c, const std::string& dir, int mode) {
  const std::string node = read_small_file(dir + "/numa_node");
  if (!node.empty()) c->numa_node = atoi(node.c_str());
  const std::string cpus = read_small_file(dir + "/local_cpulist");
  c->local_cpulist = cpus;
  if (mode == 0 || cpus.empty()) return;
  cpu_set_t want, have, both;
  if (parse_cpulist(cpus.c_str(), &want) == 0) return;
  if (sched_getaffinity(0, sizeof(have), &have) != 0) return;
  CPU_AND(&both, &want, &have);
  if (CPU_COUNT(&both) == 0) { dbg(1, "rank %d: GPU-local CPUs %s are outside this process's cpuset; affinity unchanged", c->rank, cpus.c_str()); return; }
  if (CPU_EQUAL(&both, &have)) return;
  c->affinity_saved = have; c->affinity_changed = true;
  if (sched_setaffinity(0, sizeof(both), &both) != 0) { c->affinity_changed = false; dbg(1, "rank %d: sched_setaffinity failed; affinity unchanged", c->rank); return; }
  dbg(1, "rank %d: bound to the CPUs of NUMA node %d (%s)", c->rank, c->numa_node, cpus.c_str());
}

void bind_to_gpu_numa(b200collComm* c, bool allow_bind) {
  c->numa_node = -1;
  const char* e = getenv("B200COLL_AFFINITY");
  const int mode = !allow_bind ? 0 : (e && *e) ? atoi(e) : 1;
  const std::string dir = gpu_sysfs_dir(c->device);
  if (dir.empty()) return;
  apply_gpu_locality(c, dir, mode);
}

void restore_affinity_after_init(b200collComm* c) {
  const char* e = getenv("B200COLL_AFFINITY");
  if (!(e && atoi(e) == 2) || !c->affinity_changed) return;
  sched_setaffinity(0, sizeof(c->affinity_saved), &c->affinity_saved);
  c->affinity_changed = false;
}

void hostpath_destroy(b200collComm* c) {
  HostPath& h = c->host;
  for (int k = 0; k < 2; k++) {
    if (h.ev_in_ready[k]) cudaEventDestroy(h.ev_in_ready[k]);
    if (h.ev_in_free[k]) cudaEventDestroy(h.ev_in_free[k]);
    if (h.ev_out_free[k]) cudaEventDestroy(h.ev_out_free[k]);
    h.ev_in_ready[k] = h.ev_in_free[k] = h.ev_out_free[k] = nullptr;
  }
  if (h.ev_start) { cudaEventDestroy(h.ev_start); h.ev_start = nullptr; }
  if (h.h2d) { cudaStreamDestroy(h.h2d); h.h2d = nullptr; }
  if (h.d2h) { cudaStreamDestroy(h.d2h); h.d2h = nullptr; }
  for (auto& kv : c->host_allocs) { cudaHostUnregister(kv.first); munmap(kv.first, kv.second); }
  c->host_allocs.clear();
}

static b200collResult_t hostpath_init(b200collComm* c, size_t chunk_bytes) {
  HostPath& h = c->host;
  if (!h.h2d) {
    if (cudaStreamCreateWithFlags(&h.h2d, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&h.d2h, cudaStreamNonBlocking) != cudaSuccess) {
      set_last_error("host path: cannot create copy streams"); return b200collUnhandledCudaError;
    }
    for (int k = 0; k < 2; k++) {
      cudaEventCreateWithFlags(&h.ev_in_ready[k], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h.ev_in_free[k], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&h.ev_out_free[k], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&h.ev_start, cudaEventDisableTiming);
  }
  if (h.chunk_bytes < chunk_bytes) {       // (re)allocate the staging rings: a collective contract like MemAlloc (same call sequence on every rank)
    for (int k = 0; k < 2; k++) {
      if (h.in[k]) b200collMemFree(c, h.in[k]);
      if (h.out[k]) b200collMemFree(c, h.out[k]);
      h.in[k] = h.out[k] = nullptr;
    }
    for (int k = 0; k < 2; k++) {
      b200collResult_t rc = b200collMemAlloc(c, &h.in[k], chunk_bytes);
      if (rc == b200collSuccess) rc = b200collMemAlloc(c, &h.out[k], chunk_bytes);
      if (rc != b200collSuccess) { h.chunk_bytes = 0; return rc; }
    }
    h.chunk_bytes = chunk_bytes;
  }
  return b200collSuccess;
}

static long env_l(const char* name, long dflt) { const char* e = getenv(name); return (e && *e) ? atol(e) : dflt; }

}  // namespace b200coll

using namespace b200coll;

extern "C" {

int b200collDebugParseCpuList(const char* s, int* cpus, int max) {
  cpu_set_t set;
  parse_cpulist(s, &set);
  int n = 0;
  for (int i = 0; i < CPU_SETSIZE && n < max; i++) if (CPU_ISSET(i, &set)) cpus[n++] = i;
  return n;
}

// Test hook (no GPU): apply the locality rules to the calling thread from a directory laid out like /sys/bus/pci/devices/<gpu>
// (files numa_node, local_cpulist). Returns the node read (-1 if none); *changed says whether the thread's affinity was narrowed.
int b200collDebugApplyLocality(const char* sysfs_dir, int mode, int* changed) {
  b200collComm fake;
  fake.numa_node = -1;
  apply_gpu_locality(&fake, sysfs_dir ? sysfs_dir : "", mode);
  if (changed) *changed = fake.affinity_changed ? 1 : 0;
  return fake.numa_node;
}

b200collResult_t b200collCommNumaGet(b200collComm_t c, int* numa_node, char* cpulist, size_t len) {
  if (!c) return b200collInvalidArgument;
  if (numa_node) *numa_node = c->numa_node;
  if (cpulist && len) { strncpy(cpulist, c->local_cpulist.c_str(), len - 1); cpulist[len - 1] = 0; }
  return b200collSuccess;
}

// Pinned + device-mapped host memory on the NUMA node of this rank's GPU. mbind(MPOL_PREFERRED) when the container allows it;
// otherwise first touch from a GPU-local CPU (the thread is moved there for the duration of the touch if it is not bound already).
b200collResult_t b200collHostAlloc(b200collComm_t c, void** ptr, size_t bytes) {
  if (!c || !ptr || bytes == 0) return b200collInvalidArgument;
  const size_t len = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) { set_last_error("host alloc: mmap failed"); return b200collSystemError; }
  (void)madvise(p, len, MADV_HUGEPAGE);
  bool placed = false;
  if (c->numa_node >= 0 && c->numa_node < 64) {
    unsigned long mask = 1ul << c->numa_node;
    placed = syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, &mask, 65ul, 0u) == 0;
  }
  cpu_set_t have, want, both;
  bool moved = false;
  if (!placed && !c->local_cpulist.empty() && sched_getaffinity(0, sizeof(have), &have) == 0 && parse_cpulist(c->local_cpulist.c_str(), &want) > 0) {
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) > 0 && !CPU_EQUAL(&both, &have)) moved = sched_setaffinity(0, sizeof(both), &both) == 0;
  }
  memset(p, 0, len);                                  // first touch: pages are allocated now, on the preferred / current node
  if (moved) sched_setaffinity(0, sizeof(have), &have);
  cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e != cudaSuccess) { munmap(p, len); set_last_error(std::string("host alloc: cudaHostRegister failed: ") + cudaGetErrorString(e)); (void)cudaGetLastError(); return b200collUnhandledCudaError; }
  { std::lock_guard<std::mutex> lk(c->mu); c->host_allocs[p] = len; }
  dbg(1, "rank %d: %zu MiB of pinned host memory on node %d (%s)", c->rank, len >> 20, c->numa_node, placed ? "mbind" : "first touch");
  *ptr = p;
  return b200collSuccess;
}

b200collResult_t b200collHostFree(b200collComm_t c, void* ptr) {
  if (!c || !ptr) return b200collInvalidArgument;
  size_t len = 0;
  { std::lock_guard<std::mutex> lk(c->mu); auto it = c->host_allocs.find(ptr); if (it == c->host_allocs.end()) { set_last_error("pointer passed to HostFree was not returned by HostAlloc"); return b200collInvalidArgument; } len = it->second; c->host_allocs.erase(it); }
  cudaHostUnregister(ptr);
  munmap(ptr, len);
  return b200collSuccess;
}

// host_recv = cast(scale * sum over ranks of host_send). Both buffers are pinned host memory (HostAlloc, cudaHostAlloc or
// cudaHostRegister); host_recv == host_send is allowed for equal-size types. Asynchronous on `stream`: the result is in host_recv
// when the stream reaches the point after this call. The same call (count, types) on every rank.
//   <= B200COLL_HOST_ZEROCOPY_KB (default 4096) and within reach of a Lamport kernel: one kernel, reads and writes host memory directly
//   <  B200COLL_HOST_PIPELINE_KB (default 8192): copy in, all-reduce, copy out on `stream`
//   larger: chunks of B200COLL_HOST_CHUNK_KB (default: total/6 clamped to [4 MiB, 16 MiB]) through two staging pairs, three legs overlapped
b200collResult_t b200collAllReduceHost(const void* host_send, void* host_recv, size_t count, const b200collEpilogue* ep, b200collRedOp_t rop,
                                       b200collComm_t c, b200collStream_t stream) {
  if (!c || !host_send || !host_recv || !ep) { set_last_error("null argument"); return b200collInvalidArgument; }
  if (count == 0) return b200collSuccess;
  const size_t is = b200collTypeSize(ep->in_dtype), os = b200collTypeSize(ep->out_dtype);
  if (is == 0 || os == 0) { set_last_error("bad dtype"); return b200collInvalidArgument; }
  if (host_send == host_recv && is != os) { set_last_error("in-place host all-reduce needs in_dtype and out_dtype of equal size"); return b200collInvalidArgument; }
  cudaStream_t main = static_cast<cudaStream_t>(stream);
  const size_t bytes = count * is;
  // Measured on one B200 (profiles/host_path.md): the copy engines move 55.6 GB/s in and 57.3 GB/s out, 49 GB/s each way when both run;
  // a chunk boundary costs ~20 us of bubble (event hand-over between three streams), so the pipeline wants 4-8 big chunks; below 8 MiB
  // there is nothing to win with DMA, but a kernel that reads and writes pinned host memory itself overlaps both directions by nature
  // and needs one launch instead of copy + kernel + copy.
  static const size_t zc_max = (size_t)env_l("B200COLL_HOST_ZEROCOPY_KB", 4096) << 10;
  static const size_t pipe_min = (size_t)env_l("B200COLL_HOST_PIPELINE_KB", 8192) << 10;
  static const size_t chunk_env = (size_t)env_l("B200COLL_HOST_CHUNK_KB", 0) << 10;
  c->stats.host_calls++; c->stats.host_bytes += bytes;

  // One kernel, no copy engine: nranks > 1 takes a Lamport kernel (they only ever read `in` and write `out` locally, so both may be
  // host memory): one-shot up to 512 KiB, two-shot up to 512 KiB x nranks when the types have equal size. One rank: the copy kernel.
  // (the generic reductions — min / max / prod, integer, fp64 — stage non-arena buffers with device-to-device copies: copy engines for them)
  const bool zc_feasible = !needs_generic(ep, rop) && (c->nranks == 1 || bytes <= kLLOneShotMaxBytes || (is == os && bytes <= kLLOneShotMaxBytes * (size_t)c->nranks));
  if (bytes <= zc_max && zc_feasible) {
    void *din = nullptr, *dout = nullptr;
    if (cudaHostGetDevicePointer(&din, const_cast<void*>(host_send), 0) == cudaSuccess && cudaHostGetDevicePointer(&dout, host_recv, 0) == cudaSuccess) {
      c->stats.host_zero_copy++;
      return b200collAllReduce(din, dout, count, ep, rop, c, stream);
    }
    (void)cudaGetLastError();                                              // not mapped: fall through to the copy engines
  }
  size_t chunk = chunk_env ? chunk_env : std::min<size_t>(16u << 20, std::max<size_t>(4u << 20, bytes / 6));
  if (bytes < pipe_min) chunk = bytes;
  chunk = (chunk + 1023) / 1024 * 1024;
  const size_t chunk_elems = chunk / is;
  const size_t out_chunk_bytes = chunk_elems * os;
  b200collResult_t rc = hostpath_init(c, std::max(chunk, out_chunk_bytes));
  if (rc != b200collSuccess) return rc;
  HostPath& h = c->host;
#define HP_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_last_error(std::string("host path: ") + #call + ": " + cudaGetErrorString(e__)); return b200collUnhandledCudaError; } } while (0)
  const char* src = static_cast<const char*>(host_send);
  char* dst = static_cast<char*>(host_recv);
  if (chunk >= bytes) {                                                    // one chunk: nothing to overlap, stay on the caller's stream
    HP_TRY(cudaMemcpyAsync(h.in[0], src, bytes, cudaMemcpyHostToDevice, main));
    rc = b200collAllReduce(h.in[0], h.out[0], count, ep, rop, c, stream);
    if (rc != b200collSuccess) return rc;
    HP_TRY(cudaMemcpyAsync(dst, h.out[0], count * os, cudaMemcpyDeviceToHost, main));
    return b200collSuccess;
  }
  c->stats.host_pipelined++;
  HP_TRY(cudaEventRecord(h.ev_start, main));                               // earlier work on the caller's stream (and the previous call's use of the rings) comes first
  HP_TRY(cudaStreamWaitEvent(h.h2d, h.ev_start, 0));
  HP_TRY(cudaStreamWaitEvent(h.d2h, h.ev_start, 0));
  size_t i = 0;
  for (size_t done = 0; done < count; done += chunk_elems, i++) {
    const int k = (int)(i & 1);
    const size_t n = std::min(chunk_elems, count - done);
    if (i >= 2) HP_TRY(cudaStreamWaitEvent(h.h2d, h.ev_in_free[k], 0));   // the all-reduce that read in[k] two chunks ago is done
    HP_TRY(cudaMemcpyAsync(h.in[k], src + done * is, n * is, cudaMemcpyHostToDevice, h.h2d));
    HP_TRY(cudaEventRecord(h.ev_in_ready[k], h.h2d));
    HP_TRY(cudaStreamWaitEvent(main, h.ev_in_ready[k], 0));
    if (i >= 2) HP_TRY(cudaStreamWaitEvent(main, h.ev_out_free[k], 0));   // the copy-back that read out[k] two chunks ago is done
    rc = b200collAllReduce(h.in[k], h.out[k], n, ep, rop, c, stream);
    if (rc != b200collSuccess) return rc;
    HP_TRY(cudaEventRecord(h.ev_in_free[k], main));
    HP_TRY(cudaStreamWaitEvent(h.d2h, h.ev_in_free[k], 0));
    HP_TRY(cudaMemcpyAsync(dst + done * os, h.out[k], n * os, cudaMemcpyDeviceToHost, h.d2h));
    HP_TRY(cudaEventRecord(h.ev_out_free[k], h.d2h));
  }
  HP_TRY(cudaStreamWaitEvent(main, h.ev_out_free[0], 0));                  // join: the caller's stream continues once the last copy-backs have landed
  if (i >= 2) HP_TRY(cudaStreamWaitEvent(main, h.ev_out_free[1], 0));
#undef HP_TRY
  return b200collSuccess;
}

}  // extern "C"
