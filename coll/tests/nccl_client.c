/* A plain NCCL program — it includes the SYSTEM's <nccl.h> (not a header of this repository) and uses nothing but the NCCL C API,
 * the way nccl-tests' *_perf binaries do in one-process-many-GPUs mode: ncclCommInitAll, then per collective one
 * ncclGroupStart / per-device call / ncclGroupEnd, results checked on the host. Linked against libb200coll_nccl.so it runs on
 * libb200coll (tests/test_coll_gpu.py::test_unmodified_nccl_program_runs_on_the_shim); linked against a real libnccl it is an ordinary
 * NCCL test (with distinct devices). This is the "drop-in" claim as an executable: reference role = the pods that run
 * /third_party/nccl-tests/build/all_reduce_perf against whatever NCCL the installer dropped (gpudirect-tcpx/nccl-config.yaml:22,61).
 *
 *   nccl_client [ndev] [--same-device]      --same-device: every rank on GPU 0 (virtual ranks; only the shim accepts that)
 */
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXDEV 8
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(2); } } while (0)
#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { fprintf(stderr, "NCCL %s:%d %s: %s\n", __FILE__, __LINE__, #x, ncclGetErrorString(r_)); exit(3); } } while (0)

static int ndev = 2, devs[MAXDEV];
static ncclComm_t comms[MAXDEV];
static cudaStream_t streams[MAXDEV];
static void *d_in[MAXDEV], *d_out[MAXDEV];
static size_t cap_bytes = 8u << 20;
static int failures = 0;

static void sync_all(void) { for (int i = 0; i < ndev; i++) { CK(cudaSetDevice(devs[i])); CK(cudaStreamSynchronize(streams[i])); } }
static void upload(int i, const void* h, size_t bytes) { CK(cudaSetDevice(devs[i])); CK(cudaMemcpy(d_in[i], h, bytes, cudaMemcpyHostToDevice)); }
static void download(int i, void* h, size_t bytes) { CK(cudaSetDevice(devs[i])); CK(cudaMemcpy(h, d_out[i], bytes, cudaMemcpyDeviceToHost)); }
static void report(const char* what, int ok) { printf("%-44s %s\n", what, ok ? "ok" : "WRONG"); if (!ok) failures++; }

/* value rank r contributes at element e (small integers: every type holds them, products stay tiny) */
static long val(int r, size_t e) { return (long)((e * 7 + (size_t)r * 13) % 5) + 1; }

#define DEFINE_ALLREDUCE(NAME, T, NCCLT)                                                                          \
  static void NAME(const char* what, ncclRedOp_t op, size_t count) {                                              \
    T* h = (T*)malloc(count * sizeof(T));                                                                         \
    for (int i = 0; i < ndev; i++) { for (size_t e = 0; e < count; e++) h[e] = (T)val(i, e); upload(i, h, count * sizeof(T)); } \
    NC(ncclGroupStart());                                                                                         \
    for (int i = 0; i < ndev; i++) { CK(cudaSetDevice(devs[i])); NC(ncclAllReduce(d_in[i], d_out[i], count, NCCLT, op, comms[i], streams[i])); } \
    NC(ncclGroupEnd());                                                                                           \
    sync_all();                                                                                                   \
    int ok = 1;                                                                                                   \
    for (int i = 0; i < ndev && ok; i++) {                                                                        \
      download(i, h, count * sizeof(T));                                                                          \
      for (size_t e = 0; e < count; e++) {                                                                        \
        double want = (double)val(0, e);                                                                          \
        for (int r = 1; r < ndev; r++) { double v = (double)val(r, e); want = op == ncclSum ? want + v : op == ncclProd ? want * v : op == ncclMax ? (v > want ? v : want) : (v < want ? v : want); } \
        if ((double)h[e] != want) { ok = 0; fprintf(stderr, "%s: rank %d element %zu: got %g want %g\n", what, i, e, (double)h[e], want); break; } \
      }                                                                                                           \
    }                                                                                                             \
    free(h);                                                                                                      \
    report(what, ok);                                                                                             \
  }
DEFINE_ALLREDUCE(ar_i32, int32_t, ncclInt32)
DEFINE_ALLREDUCE(ar_i64, int64_t, ncclInt64)
DEFINE_ALLREDUCE(ar_u8, uint8_t, ncclUint8)
DEFINE_ALLREDUCE(ar_f32, float, ncclFloat32)
DEFINE_ALLREDUCE(ar_f64, double, ncclFloat64)

static void bcast_and_gather(size_t count) {
  int32_t* h = (int32_t*)malloc(count * ndev * sizeof(int32_t));
  for (int i = 0; i < ndev; i++) { for (size_t e = 0; e < count; e++) h[e] = (int32_t)(val(i, e) * 1000 + i); upload(i, h, count * 4); }
  const int root = ndev - 1;
  NC(ncclGroupStart());
  for (int i = 0; i < ndev; i++) { CK(cudaSetDevice(devs[i])); NC(ncclBroadcast(d_in[i], d_out[i], count, ncclInt32, root, comms[i], streams[i])); }
  NC(ncclGroupEnd());
  sync_all();
  int ok = 1;
  for (int i = 0; i < ndev; i++) { download(i, h, count * 4); for (size_t e = 0; e < count; e++) if (h[e] != (int32_t)(val(root, e) * 1000 + root)) ok = 0; }
  report("ncclBroadcast int32", ok);
  NC(ncclGroupStart());
  for (int i = 0; i < ndev; i++) { CK(cudaSetDevice(devs[i])); NC(ncclAllGather(d_in[i], d_out[i], count, ncclInt32, comms[i], streams[i])); }
  NC(ncclGroupEnd());
  sync_all();
  ok = 1;
  for (int i = 0; i < ndev; i++) { download(i, h, count * ndev * 4); for (int r = 0; r < ndev; r++) for (size_t e = 0; e < count; e++) if (h[(size_t)r * count + e] != (int32_t)(val(r, e) * 1000 + r)) ok = 0; }
  report("ncclAllGather int32", ok);
  free(h);
}

static void ring_send_recv(size_t count) {
  float* h = (float*)malloc(count * sizeof(float));
  for (int i = 0; i < ndev; i++) { for (size_t e = 0; e < count; e++) h[e] = (float)(val(i, e) + 10 * i); upload(i, h, count * 4); }
  NC(ncclGroupStart());
  for (int i = 0; i < ndev; i++) {
    CK(cudaSetDevice(devs[i]));
    NC(ncclSend(d_in[i], count, ncclFloat32, (i + 1) % ndev, comms[i], streams[i]));
    NC(ncclRecv(d_out[i], count, ncclFloat32, (i + ndev - 1) % ndev, comms[i], streams[i]));
  }
  NC(ncclGroupEnd());
  sync_all();
  int ok = 1;
  for (int i = 0; i < ndev; i++) { const int from = (i + ndev - 1) % ndev; download(i, h, count * 4); for (size_t e = 0; e < count; e++) if (h[e] != (float)(val(from, e) + 10 * from)) ok = 0; }
  report("ncclSend / ncclRecv ring (one group)", ok);
  free(h);
}

int main(int argc, char** argv) {
  int same = 0;
  for (int a = 1; a < argc; a++) { if (!strcmp(argv[a], "--same-device")) same = 1; else ndev = atoi(argv[a]); }
  if (ndev < 1 || ndev > MAXDEV) { fprintf(stderr, "ndev must be 1..%d\n", MAXDEV); return 1; }
  int visible = 0;
  CK(cudaGetDeviceCount(&visible));
  for (int i = 0; i < ndev; i++) devs[i] = same ? 0 : i % visible;
  int version = 0;
  NC(ncclGetVersion(&version));
  printf("NCCL API level of the header %d, of the library %d; %d ranks on devices", NCCL_VERSION_CODE, version, ndev);
  for (int i = 0; i < ndev; i++) printf(" %d", devs[i]);
  printf("\n");
  NC(ncclCommInitAll(comms, ndev, devs));
  for (int i = 0; i < ndev; i++) {
    CK(cudaSetDevice(devs[i]));
    CK(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
    CK(cudaMalloc(&d_in[i], cap_bytes)); CK(cudaMalloc(&d_out[i], cap_bytes));
    int n = 0, r = -1;
    NC(ncclCommCount(comms[i], &n)); NC(ncclCommUserRank(comms[i], &r));
    if (n != ndev || r != i) { fprintf(stderr, "communicator %d reports rank %d of %d\n", i, r, n); return 4; }
  }
  ar_i32("ncclAllReduce int32 max (1003 elements)", ncclMax, 1003);
  ar_i32("ncclAllReduce int32 sum (256 Ki elements)", ncclSum, 1u << 18);
  ar_i64("ncclAllReduce int64 sum", ncclSum, 5000);
  ar_u8("ncclAllReduce uint8 min", ncclMin, 4096);
  ar_f32("ncclAllReduce float sum (1 Mi elements)", ncclSum, 1u << 20);
  ar_f32("ncclAllReduce float prod", ncclProd, 777);
  ar_f64("ncclAllReduce double max", ncclMax, 2048);
  bcast_and_gather(4096);
  ring_send_recv(3001);
  for (int i = 0; i < ndev; i++) { CK(cudaSetDevice(devs[i])); NC(ncclCommDestroy(comms[i])); CK(cudaFree(d_in[i])); CK(cudaFree(d_out[i])); CK(cudaStreamDestroy(streams[i])); }
  printf(failures ? "FAILED: %d check(s)\n" : "all checks passed (%d failures)\n", failures);
  return failures ? 5 : 0;
}
