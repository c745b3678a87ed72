// nvls_probe: single-process capability probe for the symmetric-memory runtime.
// Reports VMM / POSIX-fd / fabric / multicast support, the P2P matrix, and runs
// a minimal multimem.st / multimem.ld_reduce / multimem.red smoke across all visible GPUs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/nvls_probe.cu -o build/nvls_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <unistd.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); \
  printf("FAIL %s:%d %s -> %d %s\n", __FILE__, __LINE__, #x, (int)r_, s_ ? s_ : "?"); fflush(stdout); ok = false; goto done; } } while (0)
#define CKS(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); \
  printf("soft-fail %s -> %d %s\n", #x, (int)r_, s_ ? s_ : "?"); fflush(stdout); } } while (0)
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); fflush(stdout); ok = false; goto done; } } while (0)

__global__ void fill_bf16(__nv_bfloat16* p, size_t n, float v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = __float2bfloat16(v + (float)(i % 7));
}

// every rank: multimem.st its rank id pattern into slot [rank] of the multicast window
__global__ void mc_store(uint32_t* mc, int rank, int words_per_rank) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < words_per_rank) {
    uint32_t v = 0x1000u * (rank + 1) + i;
    asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(mc + rank * words_per_rank + i), "r"(v) : "memory");
  }
}

__global__ void mc_red_flag(uint32_t* mc_flag) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_flag), "r"(1u) : "memory");
}

// out[i] = sum over devices of in[i] (bf16x2 x4, fp32 accumulate in the switch)
__global__ void mc_ldreduce(const uint4* mc_in, uint4* out, size_t nvec) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_in + i) : "memory");
    out[i] = v;
  }
}

// NVLS all-reduce slice: ld_reduce own slice, multimem.st it to everybody
__global__ void mc_allreduce_slice(const uint4* mc_in, uint4* mc_out, size_t v0, size_t v1) {
  size_t i = v0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < v1; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_in + i) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_out + i),
                 "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
  }
}

__global__ void p2p_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < nvec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char** argv) {
  bool ok = true;
  int ndev = 0;
  size_t bytes = (argc > 1 ? atol(argv[1]) : 256) << 20;
  CUdeviceptr va[16][16] = {};      // va[viewer][owner]
  CUdeviceptr mcva[16] = {};
  CUmemGenericAllocationHandle h[16] = {};
  CUmemGenericAllocationHandle mch = 0;
  size_t gran = 0, mcgran_min = 0, mcgran_rec = 0, sz = 0;
  std::vector<CUmemAccessDesc> acc;
  int mc_ok_all = 1;
  cudaStream_t st[16];
  CK(cuInit(0));
  CK(cuDeviceGetCount(&ndev));
  { int drv = 0; cuDriverGetVersion(&drv); printf("driver_version %d ndev %d\n", drv, ndev); }
  if (ndev > 16) ndev = 16;
  for (int d = 0; d < ndev; d++) {
    CUdevice dev; CK(cuDeviceGet(&dev, d));
    char name[128]; cuDeviceGetName(name, 128, dev);
    int vmm = 0, fd = 0, fab = 0, mc = 0, sms = 0, gdr = 0, cc_maj = 0, cc_min = 0;
    cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
    cuDeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
    cuDeviceGetAttribute(&fab, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev);
    cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
    cuDeviceGetAttribute(&sms, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
    cuDeviceGetAttribute(&gdr, CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED, dev);
    cuDeviceGetAttribute(&cc_maj, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev);
    cuDeviceGetAttribute(&cc_min, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev);
    size_t tot = 0; cuDeviceTotalMem(&tot, dev);
    printf("dev %d name=\"%s\" cc=%d.%d sms=%d mem=%zuMiB vmm=%d posix_fd=%d fabric=%d multicast=%d\n", d, name, cc_maj, cc_min, sms, tot >> 20, vmm, fd, fab, mc);
    mc_ok_all &= mc;
  }
  printf("p2p matrix (can_access / native_atomics):\n");
  for (int a = 0; a < ndev; a++) {
    for (int b = 0; b < ndev; b++) {
      int ca = (a == b), at = (a == b);
      if (a != b) { cudaDeviceCanAccessPeer(&ca, a, b); cudaDeviceGetP2PAttribute(&at, cudaDevP2PAttrNativeAtomicSupported, a, b); }
      printf(" %d/%d", ca, at);
    }
    printf("\n");
  }
  fflush(stdout);
  // --- VMM allocations, mapped on every device
  {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    prop.location.id = 0;
    CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    printf("vmm granularity recommended=%zu\n", gran);
    if (mc_ok_all && ndev >= 1) {
      CUmulticastObjectProp mp = {};
      mp.numDevices = ndev; mp.size = bytes; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      CKS(cuMulticastGetGranularity(&mcgran_min, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
      CKS(cuMulticastGetGranularity(&mcgran_rec, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
      printf("multicast granularity min=%zu recommended=%zu\n", mcgran_min, mcgran_rec);
    }
    size_t g = gran; if (mcgran_rec > g) g = mcgran_rec;
    sz = (bytes + g - 1) / g * g;
    for (int d = 0; d < ndev; d++) {
      RT(cudaSetDevice(d)); RT(cudaFree(0)); RT(cudaStreamCreateWithFlags(&st[d], cudaStreamNonBlocking));
      prop.location.id = d;
      CK(cuMemCreate(&h[d], sz, &prop, 0));
      int fdh = -1;
      CUresult r = cuMemExportToShareableHandle(&fdh, h[d], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      printf("dev %d cuMemCreate %zu MiB ok; export fd -> %d (fd=%d)\n", d, sz >> 20, (int)r, fdh);
      if (fdh >= 0) close(fdh);
      CUmemAccessDesc a = {}; a.location.type = CU_MEM_LOCATION_TYPE_DEVICE; a.location.id = d; a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
      acc.push_back(a);
    }
    for (int v = 0; v < ndev; v++)
      for (int o = 0; o < ndev; o++) {
        CK(cuMemAddressReserve(&va[v][o], sz, g, 0, 0));
        CK(cuMemMap(va[v][o], sz, 0, h[o], 0));
        CK(cuMemSetAccess(va[v][o], sz, &acc[v], 1));
      }
    printf("peer mapping ok\n"); fflush(stdout);
  }
  // --- P2P bandwidth (kernel copy dev0 <- dev1, and bidirectional)
  if (ndev >= 2) {
    size_t nvec = bytes / 2 / 16;
    cudaEvent_t e0, e1; RT(cudaSetDevice(0)); RT(cudaEventCreate(&e0)); RT(cudaEventCreate(&e1));
    for (int mode = 0; mode < 2; mode++) {   // 0 = pull (load remote), 1 = push (store remote)
      for (int it = 0; it < 3; it++) {
        RT(cudaEventRecord(e0, st[0]));
        if (mode == 0) p2p_copy<<<148 * 4, 512, 0, st[0]>>>((const uint4*)va[0][1], (uint4*)(va[0][0] + bytes / 2), nvec);
        else p2p_copy<<<148 * 4, 512, 0, st[0]>>>((const uint4*)va[0][0], (uint4*)(va[0][1] + bytes / 2), nvec);
        RT(cudaEventRecord(e1, st[0])); RT(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it == 2) printf("p2p %s dev0<->dev1 %zu MiB: %.3f ms  %.1f GB/s\n", mode ? "push" : "pull", (bytes / 2) >> 20, ms, bytes / 2 / ms / 1e6);
      }
    }
    fflush(stdout);
  }
  // --- multicast
  if (mc_ok_all && ndev >= 2) {
    CUmulticastObjectProp mp = {};
    mp.numDevices = ndev; mp.size = sz; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CK(cuMulticastCreate(&mch, &mp));
    { int fdh = -1; CUresult r = cuMemExportToShareableHandle(&fdh, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0); printf("multicast export fd -> %d (fd=%d)\n", (int)r, fdh); if (fdh >= 0) close(fdh); }
    for (int d = 0; d < ndev; d++) { CUdevice dev; CK(cuDeviceGet(&dev, d)); CK(cuMulticastAddDevice(mch, dev)); }
    for (int d = 0; d < ndev; d++) CK(cuMulticastBindMem(mch, 0, h[d], 0, sz, 0));
    for (int d = 0; d < ndev; d++) {
      CK(cuMemAddressReserve(&mcva[d], sz, mcgran_rec ? mcgran_rec : gran, 0, 0));
      CK(cuMemMap(mcva[d], sz, 0, mch, 0));
      CK(cuMemSetAccess(mcva[d], sz, &acc[d], 1));
    }
    printf("multicast create/bind/map ok\n"); fflush(stdout);
    // 1) multimem.st broadcast
    const int W = 256;
    for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); RT(cudaMemsetAsync((void*)va[d][d], 0, 1 << 20, st[d])); RT(cudaStreamSynchronize(st[d])); }
    for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); mc_store<<<1, W, 0, st[d]>>>((uint32_t*)mcva[d], d, W); mc_red_flag<<<1, 32, 0, st[d]>>>((uint32_t*)(mcva[d] + 65536)); }
    for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); RT(cudaStreamSynchronize(st[d])); }
    {
      int bad = 0; std::vector<uint32_t> host(W * ndev);
      for (int d = 0; d < ndev; d++) {
        RT(cudaSetDevice(d)); RT(cudaMemcpy(host.data(), (void*)va[d][d], W * ndev * 4, cudaMemcpyDeviceToHost));
        for (int r = 0; r < ndev; r++) for (int i = 0; i < W; i++) if (host[r * W + i] != 0x1000u * (r + 1) + i) bad++;
        uint32_t flag = 0; RT(cudaMemcpy(&flag, (void*)(va[d][d] + 65536), 4, cudaMemcpyDeviceToHost));
        printf("dev %d multimem.red flag = %u (expect %d)\n", d, flag, ndev);
      }
      printf("multimem.st broadcast mismatches=%d\n", bad);
    }
    // 2) ld_reduce correctness + NVLS all-reduce bandwidth
    size_t half = bytes / 2, n = half / 2, nvec = half / 16;
    for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); fill_bf16<<<1024, 256, 0, st[d]>>>((__nv_bfloat16*)va[d][d], n, (float)(d + 1)); RT(cudaStreamSynchronize(st[d])); }
    RT(cudaSetDevice(0));
    mc_ldreduce<<<148 * 2, 512, 0, st[0]>>>((const uint4*)mcva[0], (uint4*)(va[0][0] + half), 4096);
    RT(cudaStreamSynchronize(st[0]));
    {
      std::vector<__nv_bfloat16> host(64); RT(cudaMemcpy(host.data(), (void*)(va[0][0] + half), 128, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int i = 0; i < 64; i++) { float exp = 0; for (int d = 0; d < ndev; d++) exp += (float)(d + 1) + (float)(i % 7); if (fabsf(__bfloat162float(host[i]) - exp) > 0.02f * exp) bad++; }
      printf("multimem.ld_reduce mismatches=%d (first=%f)\n", bad, __bfloat162float(host[0]));
    }
    for (int blocks = 16; blocks <= 148 * 2; blocks *= 2) {
      cudaEvent_t e0[16], e1[16];
      float worst = 0;
      for (int it = 0; it < 4; it++) {
        for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); if (it == 0) { RT(cudaEventCreate(&e0[d])); RT(cudaEventCreate(&e1[d])); } RT(cudaStreamSynchronize(st[d])); }
        for (int d = 0; d < ndev; d++) {
          RT(cudaSetDevice(d));
          size_t v0 = nvec * d / ndev, v1 = nvec * (d + 1) / ndev;
          RT(cudaEventRecord(e0[d], st[d]));
          mc_allreduce_slice<<<blocks, 512, 0, st[d]>>>((const uint4*)mcva[d], (uint4*)(mcva[d] + half), v0, v1);
          RT(cudaEventRecord(e1[d], st[d]));
        }
        worst = 0;
        for (int d = 0; d < ndev; d++) { RT(cudaSetDevice(d)); RT(cudaEventSynchronize(e1[d])); float ms; cudaEventElapsedTime(&ms, e0[d], e1[d]); if (ms > worst) worst = ms; }
      }
      double algbw = half / worst / 1e6;
      printf("nvls allreduce (no barrier) %zu MiB blocks=%d: %.3f ms algbw %.1f GB/s busbw %.1f GB/s\n", half >> 20, blocks, worst, algbw, algbw * 2 * (ndev - 1) / ndev);
      fflush(stdout);
    }
  } else {
    printf("multicast path skipped (supported_all=%d ndev=%d)\n", mc_ok_all, ndev);
  }
done:
  printf("probe %s\n", ok ? "OK" : "FAILED");
  return ok ? 0 : 1;
}
