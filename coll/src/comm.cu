// libb200coll host runtime: driver-table loading, symmetric arena (CUDA VMM + NVLS multicast object),
// rendezvous, control-page initialisation, symmetric heap allocator, self-check.
#include "comm.h"

#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>

namespace b200coll {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) {
  g_last_error = s;
  if (debug_level() >= 1) fprintf(stderr, "[b200coll] error: %s\n", s.c_str());
}

int debug_level() {
  static int lvl = [] {
    const char* e = getenv("B200COLL_DEBUG");
    if (!e) return 0;
    if (!strcasecmp(e, "INFO")) return 1;
    if (!strcasecmp(e, "TRACE")) return 2;
    if (!strcasecmp(e, "WARN")) return 0;
    return atoi(e);
  }();
  return lvl;
}

void dbg(int level, const char* fmt, ...) {
  if (debug_level() < level) return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[b200coll] ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}

const Drv& drv() {
  static Drv d = [] {
    Drv t;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
      t.why = std::string("no CUDA device: ") + cudaGetErrorString(e);
      (void)cudaGetLastError();
      return t;
    }
    bool all = true;
#define X(n)                                                                                                    \
  {                                                                                                             \
    void* p = nullptr;                                                                                          \
    cudaDriverEntryPointQueryResult q;                                                                          \
    cudaError_t r = cudaGetDriverEntryPoint(#n, &p, cudaEnableDefault, &q);                                     \
    if (r != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { all = false; t.why += std::string(" missing ") + #n; (void)cudaGetLastError(); } \
    t.n = reinterpret_cast<decltype(t.n)>(p);                                                                   \
  }
    B200COLL_DRV_FUNCS(X)
#undef X
    t.ok = all;
    return t;
  }();
  return d;
}

static std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (drv().cuGetErrorString) drv().cuGetErrorString(r, &s);
  return std::string(s ? s : "unknown") + " (" + std::to_string((int)r) + ")";
}

#define CU_TRY(call)                                                                                   \
  do {                                                                                                 \
    CUresult r__ = (call);                                                                             \
    if (r__ != CUDA_SUCCESS) { set_last_error(std::string(#call) + " failed: " + cu_err(r__)); return b200collUnhandledCudaError; } \
  } while (0)
#define RT_TRY(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) { set_last_error(std::string(#call) + " failed: " + cudaGetErrorString(e__)); return b200collUnhandledCudaError; } \
  } while (0)
#define BOOT_TRY(call)                                                                                 \
  do {                                                                                                 \
    std::string e__ = (call);                                                                          \
    if (!e__.empty()) { set_last_error(std::string("bootstrap: ") + e__); return b200collSystemError; } \
  } while (0)

static size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

static long env_long(const char* name, long dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atol(e) : dflt;
}

// In-process groups: rank 0's record owns nothing special, but arenas must outlive every mapping.
struct SharedGroup {
  std::atomic<int> alive{0};
  std::atomic<int> mc_refs{0};
};

static CUmemAllocationProp arena_prop(int device) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

struct Geometry { size_t total, gran, mc_gran; };

static b200collResult_t compute_geometry(int device, int nranks, size_t arena_bytes, bool nvls, Geometry* g) {
  const Drv& d = drv();
  CUmemAllocationProp prop = arena_prop(device);
  size_t gran = 0;
  CU_TRY(d.cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  size_t want = kOffHeap + arena_bytes;
  size_t mc_gran = 0;
  if (nvls) {
    CUmulticastObjectProp mp = {};
    mp.numDevices = nranks; mp.size = round_up(want, gran); mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gmin = 0, grec = 0;
    CU_TRY(d.cuMulticastGetGranularity(&gmin, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
    CU_TRY(d.cuMulticastGetGranularity(&grec, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    mc_gran = (want >= ((size_t)1 << 30)) ? grec : gmin;   // don't burn 512 MiB of HBM on a test-sized arena
    if (mc_gran > gran) gran = mc_gran;
  }
  g->gran = gran; g->mc_gran = mc_gran ? mc_gran : gran;
  g->total = round_up(want, gran);
  return b200collSuccess;
}

static b200collResult_t map_handle(int device, CUmemGenericAllocationHandle h, size_t total, size_t align, CUdeviceptr* va) {
  const Drv& d = drv();
  CU_TRY(d.cuMemAddressReserve(va, total, align, 0, 0));
  CUresult r = d.cuMemMap(*va, total, 0, h, 0);
  if (r != CUDA_SUCCESS) { d.cuMemAddressFree(*va, total); *va = 0; set_last_error("cuMemMap failed: " + cu_err(r)); return b200collUnhandledCudaError; }
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = device; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.cuMemSetAccess(*va, total, &acc, 1);
  if (r != CUDA_SUCCESS) { d.cuMemUnmap(*va, total); d.cuMemAddressFree(*va, total); *va = 0; set_last_error("cuMemSetAccess failed: " + cu_err(r)); return b200collUnhandledCudaError; }
  return b200collSuccess;
}

static void unmap_va(CUdeviceptr va, size_t total) {
  if (!va) return;
  drv().cuMemUnmap(va, total);
  drv().cuMemAddressFree(va, total);
}

static void stats_page_open(b200collComm* c) {
  // Exported counters for the node agent's /metrics endpoint. Best effort.
  char name[96];
  snprintf(name, sizeof(name), "/b200coll.%d.%d", (int)getpid(), c->rank);
  int fd = shm_open(name, O_CREAT | O_RDWR | O_TRUNC, 0644);
  if (fd < 0) return;
  const size_t len = 4096;
  if (ftruncate(fd, len) != 0) { close(fd); shm_unlink(name); return; }
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { shm_unlink(name); return; }
  c->stats_shm = p; c->stats_shm_name = name;
}

// Page layout: 64-byte header, then b200collStats. version 1 had 4 ops (17 counters), version 2 has 6 (21 counters) and a
// last-update time in the header so exporters can drop pages left behind by processes that died without CommDestroy.
// Called at init, on CommStatsGet and from the collective path every 256 calls (collectives.cu account()).
void stats_page_publish(b200collComm* c) {
  if (!c->stats_shm) return;
  struct Header { char magic[8]; uint32_t version, pid, rank, nranks, device, nvls; uint64_t updated_unix_s; } h = {};
  static_assert(sizeof(Header) <= 64, "counters start at byte 64");
  memcpy(h.magic, "B200COLL", 8);
  h.version = 2; h.pid = (uint32_t)getpid(); h.rank = c->rank; h.nranks = c->nranks; h.device = c->device; h.nvls = c->nvls;
  h.updated_unix_s = (uint64_t)time(nullptr);
  memcpy(c->stats_shm, &h, sizeof(h));
  memcpy(static_cast<char*>(c->stats_shm) + 64, &c->stats, sizeof(c->stats));
}

static b200collResult_t init_local_state(b200collComm* c) {
  RT_TRY(cudaMalloc(&c->state_dev, kStateWords * sizeof(uint32_t)));
  RT_TRY(cudaMemset(c->state_dev, 0, kStateWords * sizeof(uint32_t)));
  RT_TRY(cudaHostAlloc(reinterpret_cast<void**>(&c->fault_host), sizeof(b200collFault), cudaHostAllocMapped));
  memset(c->fault_host, 0, sizeof(b200collFault));
  RT_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->fault_dev), c->fault_host, 0));
  // control page: flags = 0, Lamport slots = sentinel (0xFFFFFFFF words)
  char* base = reinterpret_cast<char*>(c->peer_va[c->rank]);
  RT_TRY(cudaMemset(base + kOffFlags, 0, kOffLL));
  RT_TRY(cudaMemset(base + kOffLL, 0xFF, kLLBytes));
  RT_TRY(cudaDeviceSynchronize());
  return b200collSuccess;
}

static void finalize_comm(b200collComm* c) {
  CommDev& d = c->dev;
  d.rank = c->rank; d.nranks = c->nranks;
  for (int r = 0; r < kMaxRanks; r++) d.peer[r] = reinterpret_cast<char*>(c->peer_va[r < c->nranks ? r : c->rank]);
  d.mc = c->nvls ? reinterpret_cast<char*>(c->mc_va) : nullptr;
  d.state = c->state_dev;
  d.fault = c->fault_dev;
  // one multimem.red per barrier instead of N flag stores + N polled flags. Measured (profiles/latency_ab.md): ~1 us SLOWER per kernel at
  // 2 ranks (a detour through the switch against two direct stores) and no faster at 8 (6.12 vs 6.15 us at 1 KiB, 23.5 vs 23.2 at 4 MiB):
  // the barrier is bound by the NVLink round trip, not by the number of flag stores. Off unless B200COLL_MCBAR=1.
  d.mcbar = (c->nvls && env_long("B200COLL_MCBAR", 0) != 0) ? 1 : 0;
  d.timeout_ns = c->cfg.timeout_ms == 0 ? ~0ull : (unsigned long long)c->cfg.timeout_ms * 1000000ull;   // 0 = no watchdog
  c->free_list.clear();
  c->free_list.push_back({kOffHeap, c->arena.total - kOffHeap});
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if (drv().cuDriverGetVersion) drv().cuDriverGetVersion(&c->driver_version);
  if (c->cfg.max_ctas > 0) c->max_ctas = c->cfg.max_ctas;
  if (c->max_ctas <= 0) c->max_ctas = c->sm_count > 0 ? 2 * c->sm_count : 128;
  if (c->max_ctas > kMaxBlocks) c->max_ctas = kMaxBlocks;
  c->shape[0].max_ctas = std::min<int>((int)env_long("B200COLL_NVLS_CTAS", c->shape[0].max_ctas), c->max_ctas);
  c->shape[0].threads = (int)env_long("B200COLL_NVLS_THREADS", c->shape[0].threads);
  c->shape[1].max_ctas = std::min<int>((int)env_long("B200COLL_P2P_CTAS", c->max_ctas), c->max_ctas);
  c->shape[1].threads = (int)env_long("B200COLL_P2P_THREADS", 0);
  c->shape[2].max_ctas = std::min(c->shape[2].max_ctas, c->max_ctas);
  c->shape[3].max_ctas = std::min(c->shape[3].max_ctas, c->max_ctas);
  c->shape[4].max_ctas = std::min<int>((int)env_long("B200COLL_ROOTED_CTAS", c->shape[4].max_ctas), c->max_ctas);
  c->shape[4].threads = (int)env_long("B200COLL_ROOTED_THREADS", c->shape[4].threads);
  const char* fa = getenv("B200COLL_ALGO");
  if (fa && *fa) {
    for (int a = 0; a < b200collNumAlgos; a++) if (!strcasecmp(fa, b200collAlgoName((b200collAlgo_t)a))) c->forced_algo = (b200collAlgo_t)a;
  }
  stats_page_open(c);
  stats_page_publish(c);
  if (c->loopback) { c->loopback_counted = true; g_loopback_comms.fetch_add(1); }
  dbg(1, "rank %d/%d dev %d arena %zu MiB nvls=%d loopback=%d max_ctas=%d", c->rank, c->nranks, c->device, c->arena.total >> 20, (int)c->nvls,
      (int)c->loopback, c->max_ctas);
}

std::atomic<int> g_loopback_comms{0};

static void destroy_resources(b200collComm* c) {
  const Drv& d = drv();
  hostpath_destroy(c);
  if (c->loopback_counted) { c->loopback_counted = false; g_loopback_comms.fetch_sub(1); }
  if (c->mc_va) { unmap_va(c->mc_va, c->arena.total); c->mc_va = 0; }
  if (c->mc_bound && c->mc_handle) {
    CUdevice dev;
    if (d.cuDeviceGet(&dev, c->device) == CUDA_SUCCESS) d.cuMulticastUnbind(c->mc_handle, dev, 0, c->arena.total);
  }
  bool release_mc = c->mc_owned;
  if (c->group) release_mc = c->mc_handle && (--c->group->mc_refs == 0);   // in-process group: last one out releases
  if (c->mc_handle && release_mc) d.cuMemRelease(c->mc_handle);
  c->mc_handle = 0;
  for (int r = 0; r < c->nranks; r++) {
    if (c->peer_va[r]) { unmap_va(c->peer_va[r], c->arena.total); c->peer_va[r] = 0; }
    if (c->peer_handle_owned[r] && c->peer_handle[r]) d.cuMemRelease(c->peer_handle[r]);
    c->peer_handle[r] = 0;
  }
  if (c->arena.handle) { d.cuMemRelease(c->arena.handle); c->arena.handle = 0; }
  if (c->state_dev) { cudaFree(c->state_dev); c->state_dev = nullptr; }
  if (c->fault_host) { cudaFreeHost(c->fault_host); c->fault_host = nullptr; }
  if (c->stats_shm) { munmap(c->stats_shm, 4096); shm_unlink(c->stats_shm_name.c_str()); c->stats_shm = nullptr; }
}

}  // namespace b200coll

using namespace b200coll;

extern "C" {

const char* b200collGetErrorString(b200collResult_t r) {
  switch (r) {
    case b200collSuccess: return "success";
    case b200collUnhandledCudaError: return "unhandled CUDA error";
    case b200collSystemError: return "system error";
    case b200collInternalError: return "internal error";
    case b200collInvalidArgument: return "invalid argument";
    case b200collInvalidUsage: return "invalid usage";
    case b200collRemoteError: return "remote rank timed out or aborted";
    case b200collInProgress: return "in progress";
    case b200collOutOfMemory: return "symmetric arena exhausted";
    case b200collNoDriver: return "no CUDA driver / GPU on this host";
  }
  return "unknown";
}

const char* b200collGetLastError(void) { return g_last_error.c_str(); }
int b200collGetVersion(void) { return B200COLL_VERSION_MAJOR * 1000 + B200COLL_VERSION_MINOR; }

void b200collConfigDefault(b200collConfig* cfg) {
  cfg->arena_bytes = (size_t)env_long("B200COLL_ARENA_MB", 2560) << 20;
  cfg->enable_nvls = (int)env_long("B200COLL_NVLS", -1);
  cfg->max_ctas = (int)env_long("B200COLL_MAX_CTAS", 0);
  cfg->timeout_ms = (int)env_long("B200COLL_TIMEOUT_MS", 600000);   // c10d / NCCL scale: ordinary rank skew (checkpoints, evaluation) must not trip it; 0 = never
  cfg->debug = debug_level();
}

b200collResult_t b200collGetUniqueId(b200collUniqueId* id) {
  if (!id) return b200collInvalidArgument;
  memset(id, 0, sizeof(*id));
  unsigned long long rnd = 0;
  int fd = open("/dev/urandom", O_RDONLY | O_CLOEXEC);
  if (fd >= 0) { if (read(fd, &rnd, sizeof(rnd)) != (ssize_t)sizeof(rnd)) rnd = 0; close(fd); }
  if (!rnd) rnd = ((unsigned long long)getpid() << 32) ^ (unsigned long long)time(nullptr);
  snprintf(id->internal, sizeof(id->internal), "u%016llx-%d", rnd, (int)getpid());
  return b200collSuccess;
}

b200collResult_t b200collUniqueIdFromString(const char* s, b200collUniqueId* id) {
  if (!s || !id) return b200collInvalidArgument;
  memset(id, 0, sizeof(*id));
  // FNV-1a so arbitrary characters (':' '/') are fine in the socket name
  unsigned long long h = 1469598103934665603ull;
  for (const char* p = s; *p; p++) { h ^= (unsigned char)*p; h *= 1099511628211ull; }
  snprintf(id->internal, sizeof(id->internal), "s%016llx", h);
  return b200collSuccess;
}

b200collResult_t b200collCommInitRank(b200collComm_t* out, int nranks, const b200collUniqueId* id, int rank, const b200collConfig* cfg_in) {
  if (!out || !id || nranks < 1 || nranks > B200COLL_MAX_RANKS || rank < 0 || rank >= nranks) { set_last_error("bad arguments to CommInitRank"); return b200collInvalidArgument; }
  *out = nullptr;
  const Drv& d = drv();
  if (!d.ok) { set_last_error("driver unavailable:" + d.why); return b200collNoDriver; }
  std::unique_ptr<b200collComm> c(new b200collComm());
  c->rank = rank; c->nranks = nranks;
  if (cfg_in) c->cfg = *cfg_in; else b200collConfigDefault(&c->cfg);
  if (c->cfg.timeout_ms < 0) c->cfg.timeout_ms = (int)env_long("B200COLL_TIMEOUT_MS", 600000);
  RT_TRY(cudaGetDevice(&c->device));
  RT_TRY(cudaFree(0));
  CUdevice cudev;
  CU_TRY(d.cuDeviceGet(&cudev, c->device));
  bind_to_gpu_numa(c.get(), true);      // before anything is allocated: host-side buffers created from here on land on the GPU's node

  c->boot.reset(new Bootstrap());
  std::string name(id->internal, strnlen(id->internal, sizeof(id->internal)));
  BOOT_TRY(c->boot->init(name, rank, nranks, c->cfg.timeout_ms == 0 ? 0 : std::max(c->cfg.timeout_ms, 60000)));   // 0 = wait for ever
  c->boot_name = name;

  struct Info { unsigned char uuid[16]; int mc; int p2p_all; unsigned long long arena_bytes; } mine = {};
  {
    CUuuid u; CU_TRY(d.cuDeviceGetUuid(&u, cudev)); memcpy(mine.uuid, u.bytes, 16);
    CU_TRY(d.cuDeviceGetAttribute(&mine.mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev));
    mine.arena_bytes = c->cfg.arena_bytes;
  }
  std::vector<char> all;
  BOOT_TRY(c->boot->allgather(&mine, sizeof(mine), &all));
  const Info* infos = reinterpret_cast<const Info*>(all.data());
  bool mc_all = true, dup = false;
  for (int r = 0; r < nranks; r++) {
    mc_all &= infos[r].mc != 0;
    if (infos[r].arena_bytes != mine.arena_bytes) { set_last_error("ranks disagree on arena_bytes"); return b200collInvalidUsage; }
    for (int q = 0; q < r; q++) if (!memcmp(infos[r].uuid, infos[q].uuid, 16)) dup = true;
  }
  c->loopback = dup;
  bool want_nvls = c->cfg.enable_nvls != 0 && mc_all && !dup && nranks > 1;
  if (c->cfg.enable_nvls == 1 && !want_nvls) { set_last_error("NVLS required (B200COLL_NVLS=1) but multicast is not available on every rank"); return b200collInvalidUsage; }

  Geometry g;
  b200collResult_t rc = compute_geometry(c->device, nranks, c->cfg.arena_bytes, want_nvls, &g);
  if (rc != b200collSuccess) return rc;
  c->arena.total = g.total; c->arena.device = c->device;
  CUmemAllocationProp prop = arena_prop(c->device);
  CU_TRY(d.cuMemCreate(&c->arena.handle, g.total, &prop, 0));
  int my_fd = -1;
  CU_TRY(d.cuMemExportToShareableHandle(&my_fd, c->arena.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  std::vector<int> fds;
  {
    std::string e = c->boot->exchange_fds(my_fd, &fds);
    close(my_fd);
    if (!e.empty()) { for (int f : fds) if (f >= 0) close(f); set_last_error("bootstrap: " + e); destroy_resources(c.get()); return b200collSystemError; }
  }
  rc = b200collSuccess;
  for (int r = 0; r < nranks && rc == b200collSuccess; r++) {
    CUmemGenericAllocationHandle h = c->arena.handle;
    if (r != rank) {
      CUresult cr = d.cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (cr != CUDA_SUCCESS) { set_last_error("cuMemImportFromShareableHandle failed: " + cu_err(cr)); rc = b200collUnhandledCudaError; break; }
      c->peer_handle[r] = h; c->peer_handle_owned[r] = true;
    }
    rc = map_handle(c->device, h, g.total, g.gran, &c->peer_va[r]);
  }
  for (int f : fds) if (f >= 0) close(f);
  if (rc != b200collSuccess) { destroy_resources(c.get()); return rc; }

  // ---- NVLS multicast object: rank 0 creates, everyone adds its device, binds its arena, maps the alias.
  if (want_nvls) {
    auto agree = [&](bool ok_here) -> bool {
      char b = ok_here ? 1 : 0; std::vector<char> v;
      if (!c->boot->allgather(&b, 1, &v).empty()) return false;
      for (char x : v) if (!x) return false;
      return true;
    };
    bool ok = true;
    int mc_fd = -1;
    if (rank == 0) {
      CUmulticastObjectProp mp = {};
      mp.numDevices = nranks; mp.size = g.total; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      CUresult cr = d.cuMulticastCreate(&c->mc_handle, &mp);
      if (cr == CUDA_SUCCESS) { c->mc_owned = true; cr = d.cuMemExportToShareableHandle(&mc_fd, c->mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0); }
      if (cr != CUDA_SUCCESS) { ok = false; dbg(1, "multicast create/export failed: %s", cu_err(cr).c_str()); }
    }
    ok = agree(ok);
    if (ok) {
      int got = -1;
      std::string e = c->boot->broadcast_fd(0, mc_fd, &got);
      bool here = e.empty();
      if (here && rank != 0) {
        CUresult cr = d.cuMemImportFromShareableHandle(&c->mc_handle, (void*)(uintptr_t)got, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        if (cr != CUDA_SUCCESS) { here = false; dbg(1, "multicast import failed: %s", cu_err(cr).c_str()); } else c->mc_owned = true;
      }
      if (got >= 0) close(got);
      if (here) { CUresult cr = d.cuMulticastAddDevice(c->mc_handle, cudev); if (cr != CUDA_SUCCESS) { here = false; dbg(1, "cuMulticastAddDevice failed: %s", cu_err(cr).c_str()); } }
      ok = agree(here);
    }
    if (mc_fd >= 0) close(mc_fd);
    if (ok) {
      CUresult cr = d.cuMulticastBindMem(c->mc_handle, 0, c->arena.handle, 0, g.total, 0);
      if (cr != CUDA_SUCCESS) dbg(1, "cuMulticastBindMem failed: %s", cu_err(cr).c_str()); else c->mc_bound = true;
      ok = agree(cr == CUDA_SUCCESS);
    }
    if (ok) {
      b200collResult_t mr = map_handle(c->device, c->mc_handle, g.total, g.mc_gran, &c->mc_va);
      ok = agree(mr == b200collSuccess);
    }
    c->nvls = ok;
    if (!ok) {
      if (c->mc_va) { unmap_va(c->mc_va, g.total); c->mc_va = 0; }
      if (c->cfg.enable_nvls == 1) { set_last_error("NVLS required but multicast setup failed"); destroy_resources(c.get()); return b200collUnhandledCudaError; }
      dbg(1, "NVLS disabled: multicast setup failed on some rank; P2P paths only");
    }
  }

  rc = init_local_state(c.get());
  if (rc != b200collSuccess) { destroy_resources(c.get()); return rc; }
  finalize_comm(c.get());
  if (c->loopback) {
    int same = 0;
    for (int r = 0; r < nranks; r++) if (!memcmp(infos[r].uuid, mine.uuid, 16)) same++;
    c->max_ctas = std::max(1, std::min(c->max_ctas, c->sm_count / std::max(1, same)));
    for (auto& sh : c->shape) sh.max_ctas = std::min(sh.max_ctas, c->max_ctas);
  }
  BOOT_TRY(c->boot->barrier());
  restore_affinity_after_init(c.get());
  *out = c.release();
  return b200collSuccess;
}

b200collResult_t b200collCommInitAll(b200collComm_t* comms, int n, const int* devs, const b200collConfig* cfg_in) {
  if (!comms || n < 1 || n > B200COLL_MAX_RANKS) { set_last_error("bad arguments to CommInitAll"); return b200collInvalidArgument; }
  const Drv& d = drv();
  if (!d.ok) { set_last_error("driver unavailable:" + d.why); return b200collNoDriver; }
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  std::vector<std::unique_ptr<b200collComm>> cs;
  auto fail = [&](b200collResult_t r) { for (auto& c : cs) { cudaSetDevice(c->device); destroy_resources(c.get()); } cudaSetDevice(prev_dev); return r; };
  auto group = std::make_shared<SharedGroup>();
  bool dup = false, mc_all = true;
  for (int i = 0; i < n; i++) {
    std::unique_ptr<b200collComm> c(new b200collComm());
    c->rank = i; c->nranks = n; c->device = devs ? devs[i] : i;
    if (cfg_in) c->cfg = *cfg_in; else b200collConfigDefault(&c->cfg);
    if (c->cfg.timeout_ms < 0) c->cfg.timeout_ms = (int)env_long("B200COLL_TIMEOUT_MS", 600000);
    c->group = group;
    for (int q = 0; q < i; q++) if (cs[q]->device == c->device) dup = true;
    CUdevice cudev; int mc = 0;
    if (d.cuDeviceGet(&cudev, c->device) != CUDA_SUCCESS) { set_last_error("bad device ordinal"); return fail(b200collInvalidArgument); }
    d.cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev);
    mc_all &= mc != 0;
    cs.push_back(std::move(c));
  }
  const bool want_nvls = cs[0]->cfg.enable_nvls != 0 && mc_all && !dup && n > 1;
  if (cs[0]->cfg.enable_nvls == 1 && !want_nvls) { set_last_error("NVLS required but not available for this device list"); return fail(b200collInvalidUsage); }
  Geometry g;
  b200collResult_t rc = compute_geometry(cs[0]->device, n, cs[0]->cfg.arena_bytes, want_nvls, &g);
  if (rc != b200collSuccess) return fail(rc);
  for (auto& c : cs) {
    if (cudaSetDevice(c->device) != cudaSuccess || cudaFree(0) != cudaSuccess) { set_last_error("cudaSetDevice failed"); return fail(b200collUnhandledCudaError); }
    if (!dup) {
      for (auto& o : cs) if (o->device != c->device) { cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0); if (e != cudaSuccess) (void)cudaGetLastError(); }
    }
    CUmemAllocationProp prop = arena_prop(c->device);
    CUresult cr = d.cuMemCreate(&c->arena.handle, g.total, &prop, 0);
    if (cr != CUDA_SUCCESS) { set_last_error("cuMemCreate failed: " + cu_err(cr)); return fail(b200collUnhandledCudaError); }
    c->arena.total = g.total; c->arena.device = c->device; c->loopback = dup;
  }
  for (auto& c : cs)
    for (int r = 0; r < n; r++) {
      rc = map_handle(c->device, cs[r]->arena.handle, g.total, g.gran, &c->peer_va[r]);
      if (rc != b200collSuccess) return fail(rc);
    }
  if (want_nvls) {
    bool ok = true;
    CUmulticastObjectProp mp = {};
    mp.numDevices = n; mp.size = g.total; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle mch = 0;
    CUresult cr = d.cuMulticastCreate(&mch, &mp);
    ok = cr == CUDA_SUCCESS;
    for (int i = 0; ok && i < n; i++) { CUdevice cudev; d.cuDeviceGet(&cudev, cs[i]->device); ok = d.cuMulticastAddDevice(mch, cudev) == CUDA_SUCCESS; }
    for (int i = 0; ok && i < n; i++) { ok = d.cuMulticastBindMem(mch, 0, cs[i]->arena.handle, 0, g.total, 0) == CUDA_SUCCESS; if (ok) { cs[i]->mc_bound = true; cs[i]->mc_handle = mch; } }
    for (int i = 0; ok && i < n; i++) ok = map_handle(cs[i]->device, mch, g.total, g.mc_gran, &cs[i]->mc_va) == b200collSuccess;
    if (ok) { group->mc_refs = n; for (auto& c : cs) { c->mc_handle = mch; c->nvls = true; } }
    else {
      dbg(1, "NVLS disabled for in-process group: %s", cu_err(cr).c_str());
      for (auto& c : cs) { if (c->mc_va) { unmap_va(c->mc_va, g.total); c->mc_va = 0; } c->nvls = false; }
      if (cs[0]->cfg.enable_nvls == 1) { set_last_error("NVLS required but multicast setup failed"); return fail(b200collUnhandledCudaError); }
    }
  }
  int per_dev[64] = {};
  for (auto& c : cs) if (c->device >= 0 && c->device < 64) per_dev[c->device]++;
  for (auto& c : cs) {
    cudaSetDevice(c->device);
    rc = init_local_state(c.get());
    if (rc != b200collSuccess) return fail(rc);
    finalize_comm(c.get());
    if (dup) {
      c->max_ctas = std::max(1, std::min(c->max_ctas, c->sm_count / std::max(1, per_dev[c->device % 64])));
      for (auto& sh : c->shape) sh.max_ctas = std::min(sh.max_ctas, c->max_ctas);
    }
  }
  bool one_gpu = true;
  for (auto& c : cs) one_gpu &= c->device == cs[0]->device;
  for (auto& c : cs) bind_to_gpu_numa(c.get(), one_gpu && c->rank == 0);   // one process driving several GPUs cannot sit on all their nodes
  group->alive = n;
  for (int i = 0; i < n; i++) comms[i] = cs[i].release();
  cudaSetDevice(prev_dev);
  return b200collSuccess;
}

// Who ends up where after a split: ranks with my colour, ordered by (key, old rank). Pure, so it is testable without a GPU.
static void split_plan(int nranks, int rank, const int* colors, const int* keys, int* new_rank, int* new_size) {
  std::vector<std::pair<std::pair<int, int>, int>> members;          // ((key, old rank), old rank)
  for (int r = 0; r < nranks; r++) if (colors[r] == colors[rank]) members.push_back({{keys[r], r}, r});
  std::sort(members.begin(), members.end());
  *new_size = (int)members.size();
  *new_rank = 0;
  for (size_t i = 0; i < members.size(); i++) if (members[i].second == rank) *new_rank = (int)i;
}

b200collResult_t b200collDebugSplitPlan(int nranks, int rank, const int* colors, const int* keys, int* new_rank, int* new_size) {
  if (nranks < 1 || nranks > B200COLL_MAX_RANKS || rank < 0 || rank >= nranks || !colors || !keys || !new_rank || !new_size) return b200collInvalidArgument;
  split_plan(nranks, rank, colors, keys, new_rank, new_size);
  return b200collSuccess;
}

b200collResult_t b200collCommSplit(b200collComm_t parent, int color, int key, b200collComm_t* out, const b200collConfig* cfg) {
  if (!parent || !out) { set_last_error("null argument"); return b200collInvalidArgument; }
  *out = nullptr;
  if (!parent->boot) { set_last_error("only communicators created with CommInitRank can be split (in-process groups: call CommInitAll on the subset)"); return b200collInvalidUsage; }
  struct Mine { int color, key; } mine = {color, key};
  std::vector<char> all;
  std::string e = parent->boot->allgather(&mine, sizeof(mine), &all);
  if (!e.empty()) { set_last_error("bootstrap: " + e); return b200collSystemError; }
  const uint32_t seq = parent->split_seq++;
  if (color < 0) return b200collSuccess;                               // NCCL_SPLIT_NOCOLOR: took part in the exchange, gets no communicator
  int colors[B200COLL_MAX_RANKS], keys[B200COLL_MAX_RANKS];
  for (int r = 0; r < parent->nranks; r++) { Mine m; memcpy(&m, all.data() + r * sizeof(Mine), sizeof(Mine)); colors[r] = m.color; keys[r] = m.key; }
  int new_rank = 0, new_size = 0;
  split_plan(parent->nranks, parent->rank, colors, keys, &new_rank, &new_size);
  b200collUniqueId id;
  const std::string name = parent->boot_name + "/split" + std::to_string(seq) + "/color" + std::to_string(color);
  b200collUniqueIdFromString(name.c_str(), &id);
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != parent->device) cudaSetDevice(parent->device);          // the child lives on the parent's GPU
  const b200collConfig child_cfg = cfg ? *cfg : parent->cfg;
  const b200collResult_t rc = b200collCommInitRank(out, new_size, &id, new_rank, &child_cfg);
  if (prev != parent->device) cudaSetDevice(prev);
  return rc;
}

static std::atomic<b200collComm*> g_alloc_comm{nullptr};      // communicator behind b200collTorchAlloc / b200collTorchFree (below)

b200collResult_t b200collCommDestroy(b200collComm_t c) {
  if (!c) return b200collInvalidArgument;
  { b200collComm* expected = c; g_alloc_comm.compare_exchange_strong(expected, nullptr); }      // frees arriving later become no-ops instead of touching a dead arena
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->boot) (void)c->boot->barrier();   // peers may still be reading my arena until they reach this point
  destroy_resources(c);
  cudaSetDevice(prev);
  delete c;
  return b200collSuccess;
}

b200collResult_t b200collCommInfoGet(b200collComm_t c, b200collCommInfo* info) {
  if (!c || !info) return b200collInvalidArgument;
  info->rank = c->rank; info->nranks = c->nranks; info->device = c->device; info->nvls = c->nvls; info->p2p_ok = 1;
  info->same_device_loopback = c->loopback; info->arena_bytes = c->arena.total - kOffHeap;
  size_t used = 0;
  { std::lock_guard<std::mutex> lk(c->mu); for (auto& kv : c->live) used += kv.second; }
  info->arena_used = used; info->sm_count = c->sm_count; info->driver_version = c->driver_version;
  return b200collSuccess;
}

b200collResult_t b200collCommStatsGet(b200collComm_t c, b200collStats* s) {
  if (!c || !s) return b200collInvalidArgument;
  *s = c->stats;
  stats_page_publish(c);
  return b200collSuccess;
}

b200collResult_t b200collCommGetAsyncError(b200collComm_t c, b200collFault* f) {
  if (!c) return b200collInvalidArgument;
  b200collFault cur;
  __sync_synchronize();
  memcpy(&cur, c->fault_host, sizeof(cur));
  if (f) *f = cur;
  return cur.code ? b200collRemoteError : b200collSuccess;
}

b200collResult_t b200collHostBarrier(b200collComm_t c) {
  if (!c) return b200collInvalidArgument;
  if (!c->boot) return b200collSuccess;
  BOOT_TRY(c->boot->barrier());
  return b200collSuccess;
}

// ---- symmetric heap: deterministic first-fit, so identical call sequences give identical offsets on every rank.
b200collResult_t b200collMemAlloc(b200collComm_t c, void** ptr, size_t bytes) {
  if (!c || !ptr || bytes == 0) return b200collInvalidArgument;
  const size_t need = round_up(bytes, 512);
  std::lock_guard<std::mutex> lk(c->mu);
  for (size_t i = 0; i < c->free_list.size(); i++) {
    FreeBlock& fb = c->free_list[i];
    if (fb.len >= need) {
      const size_t off = fb.off;
      fb.off += need; fb.len -= need;
      if (fb.len == 0) c->free_list.erase(c->free_list.begin() + i);
      c->live[off] = need;
      *ptr = reinterpret_cast<void*>(c->peer_va[c->rank] + off);
      return b200collSuccess;
    }
  }
  set_last_error("symmetric arena exhausted (raise B200COLL_ARENA_MB)");
  return b200collOutOfMemory;
}

b200collResult_t b200collMemFree(b200collComm_t c, void* ptr) {
  if (!c || !ptr) return b200collInvalidArgument;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t off = reinterpret_cast<CUdeviceptr>(ptr) - c->peer_va[c->rank];
  auto it = c->live.find(off);
  if (it == c->live.end()) { set_last_error("pointer passed to MemFree was not returned by MemAlloc"); return b200collInvalidArgument; }
  FreeBlock nb{off, it->second};
  c->live.erase(it);
  auto pos = std::lower_bound(c->free_list.begin(), c->free_list.end(), nb, [](const FreeBlock& a, const FreeBlock& b) { return a.off < b.off; });
  pos = c->free_list.insert(pos, nb);
  if (pos + 1 != c->free_list.end() && pos->off + pos->len == (pos + 1)->off) { pos->len += (pos + 1)->len; c->free_list.erase(pos + 1); }
  if (pos != c->free_list.begin() && (pos - 1)->off + (pos - 1)->len == pos->off) { (pos - 1)->len += pos->len; c->free_list.erase(pos); }
  return b200collSuccess;
}

int b200collIsSymmetric(b200collComm_t c, const void* ptr, size_t bytes) {
  if (!c || !ptr) return 0;
  const CUdeviceptr p = reinterpret_cast<CUdeviceptr>(ptr), base = c->peer_va[c->rank];
  return p >= base + kOffStage && p + bytes <= base + c->arena.total;
}

// ---- PyTorch pluggable allocator (torch.cuda.memory.CUDAPluggableAllocator + torch.cuda.MemPool): tensors created under the pool live in
// the symmetric arena of the communicator named here, so the collectives on them are zero-copy and NVLS-capable without the application
// calling MemAlloc itself (what ncclMemAlloc is to PyTorch's NCCL pools). Symmetry needs the same allocation sequence on every rank, which
// SPMD training code has by construction. One allocator communicator per process.
b200collResult_t b200collSetAllocatorComm(b200collComm_t c) { g_alloc_comm.store(c); return b200collSuccess; }
void* b200collTorchAlloc(size_t size, int /*device*/, void* /*stream*/) {
  b200collComm* c = g_alloc_comm.load();
  void* p = nullptr;
  if (!c || b200collMemAlloc(c, &p, size) != b200collSuccess) return nullptr;       // PyTorch turns nullptr into its out-of-memory error
  return p;
}
void b200collTorchFree(void* ptr, size_t /*size*/, int /*device*/, void* /*stream*/) {
  b200collComm* c = g_alloc_comm.load();
  if (c && ptr) b200collMemFree(c, ptr);
}

b200collAlgo_t b200collCommGetAlgo(b200collComm_t c) { return c ? c->forced_algo : b200collAlgoAuto; }

b200collResult_t b200collCommSetAlgo(b200collComm_t c, b200collAlgo_t a) {
  if (!c || a < 0 || a >= b200collNumAlgos) return b200collInvalidArgument;
  c->forced_algo = a;
  return b200collSuccess;
}

b200collResult_t b200collCommSetMaxCtas(b200collComm_t c, int m) {
  if (!c || m < 1) return b200collInvalidArgument;
  c->max_ctas = std::min(m, kMaxBlocks);
  for (auto& sh : c->shape) sh.max_ctas = c->max_ctas;
  return b200collSuccess;
}

b200collResult_t b200collCommSetP2pWindow(b200collComm_t c, size_t bytes) {
  if (!c || (bytes != 0 && bytes < 512)) return b200collInvalidArgument;
  c->p2p_window = bytes / 512 * 512;
  return b200collSuccess;
}

b200collResult_t b200collCommSetLaunchShape(b200collComm_t c, int kind, int max_ctas, int threads) {
  if (!c || kind < 0 || kind > 4) return b200collInvalidArgument;
  if (threads != 0 && (threads < 32 || threads > 512 || threads % 32)) return b200collInvalidArgument;
  if (max_ctas > 0) c->shape[kind].max_ctas = std::min(max_ctas, kMaxBlocks);
  c->shape[kind].threads = threads;
  return b200collSuccess;
}

b200collResult_t b200collSelfCheck(char* buf, size_t buflen) {
  // The guest-config-checker role: refuse to start on a box that cannot run the fast paths, and say why.
  std::string rep;
  bool ok = true;
  const Drv& d = drv();
  if (!d.ok) { rep = "driver: unavailable (" + d.why + ")\n"; ok = false; }
  else {
    int n = 0, drvver = 0;
    cudaGetDeviceCount(&n);
    d.cuDriverGetVersion(&drvver);
    rep += "driver_version: " + std::to_string(drvver) + "\n";
    if (drvver < 12010) { rep += "FAIL driver older than CUDA 12.1 (multicast objects need it)\n"; ok = false; }
    for (int i = 0; i < n; i++) {
      CUdevice dev; d.cuDeviceGet(&dev, i);
      int vmm = 0, fd = 0, mc = 0, maj = 0, min = 0;
      d.cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
      d.cuDeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
      d.cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
      d.cuDeviceGetAttribute(&maj, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev);
      d.cuDeviceGetAttribute(&min, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev);
      char line[160];
      snprintf(line, sizeof(line), "gpu %d: cc=%d.%d vmm=%d posix_fd=%d multicast=%d\n", i, maj, min, vmm, fd, mc);
      rep += line;
      if (maj != 10) { rep += "FAIL gpu " + std::to_string(i) + " is not sm_100 (kernels are built for sm_100a only)\n"; ok = false; }
      if (!vmm || !fd) { rep += "FAIL gpu " + std::to_string(i) + " lacks VMM/POSIX-fd handles\n"; ok = false; }
      if (!mc) rep += "WARN gpu " + std::to_string(i) + " has no multicast: NVLS paths disabled, P2P only\n";
      for (int j = 0; j < n; j++) if (j != i) {
        int can = 0; cudaDeviceCanAccessPeer(&can, i, j);
        if (!can) { rep += "FAIL no P2P access " + std::to_string(i) + "->" + std::to_string(j) + "\n"; ok = false; }
      }
    }
    if (n == 0) { rep += "FAIL no GPUs visible\n"; ok = false; }
  }
  rep += ok ? "self-check: OK\n" : "self-check: FAILED\n";
  if (buf && buflen) { strncpy(buf, rep.c_str(), buflen - 1); buf[buflen - 1] = 0; }
  return ok ? b200collSuccess : (d.ok ? b200collInvalidUsage : b200collNoDriver);
}

}  // extern "C"

// Test hook (CPU-only): run the rendezvous protocol with arbitrary descriptors (memfd in the unit tests) so the
// SCM_RIGHTS plumbing is covered without a GPU. out_fds must hold nranks ints; bcast_fd receives rank 0's descriptor.
// How many connections rank 0's last BootstrapSelfTest dropped (wrong user / token / malformed hello).
static std::atomic<int> g_selftest_rejected{0};
extern "C" int b200collBootstrapSelfTestRejected(void) { return g_selftest_rejected.load(); }

extern "C" int b200collBootstrapSelfTest(const char* name, int rank, int nranks, int my_fd, int* out_fds, int* bcast_fd, int timeout_ms) {
  b200coll::Bootstrap b;
  std::string e = b.init(name ? name : "selftest", rank, nranks, timeout_ms);
  if (!e.empty()) { b200coll::set_last_error(e); return 1; }
  std::vector<char> all;
  int mine = rank * 7 + 1;
  if (!(e = b.allgather(&mine, sizeof(mine), &all)).empty()) { b200coll::set_last_error(e); return 2; }
  for (int r = 0; r < nranks; r++) { int v; memcpy(&v, all.data() + r * sizeof(int), sizeof(int)); if (v != r * 7 + 1) { b200coll::set_last_error("allgather payload mismatch"); return 3; } }
  std::vector<int> fds;
  if (!(e = b.exchange_fds(my_fd, &fds)).empty()) { b200coll::set_last_error(e); return 4; }
  for (int r = 0; r < nranks; r++) out_fds[r] = fds[r];
  if (!(e = b.broadcast_fd(0, rank == 0 ? my_fd : -1, bcast_fd)).empty()) { b200coll::set_last_error(e); return 5; }
  if (!(e = b.barrier()).empty()) { b200coll::set_last_error(e); return 6; }
  g_selftest_rejected.store(b.rejected());
  return 0;
}
