// mps_probe — prints what an MPS client actually gets: free/total memory and the SM count visible to this process,
// per device. Used to verify the limits the device plugin sets in Allocate (CUDA_MPS_ACTIVE_THREAD_PERCENTAGE,
// CUDA_MPS_PINNED_DEVICE_MEM_LIMIT; reference manager.go:333-346; probe role: example/cuda-mps/cuda_mem_and_sm_count.c:19-59).
// Adds machine-readable output (--json) and the limits' env values so an e2e can assert on them.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char** argv) {
  const bool json = argc > 1 && !strcmp(argv[1], "--json");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { fprintf(stderr, "cudaGetDeviceCount returned: %s\n", cudaGetErrorString(e)); return 1; }
  if (n == 0) { printf("No GPU devices found\n"); return 1; }
  const char* pct = getenv("CUDA_MPS_ACTIVE_THREAD_PERCENTAGE");
  const char* lim = getenv("CUDA_MPS_PINNED_DEVICE_MEM_LIMIT");
  if (json) printf("{\"active_thread_percentage\":\"%s\",\"pinned_device_mem_limit\":\"%s\",\"devices\":[", pct ? pct : "", lim ? lim : "");
  for (int i = 0; i < n; i++) {
    if ((e = cudaSetDevice(i)) != cudaSuccess) { fprintf(stderr, "cudaSetDevice(%d): %s\n", i, cudaGetErrorString(e)); return (int)e; }
    size_t free_b = 0, total_b = 0;
    if ((e = cudaMemGetInfo(&free_b, &total_b)) != cudaSuccess) { fprintf(stderr, "cudaMemGetInfo returned: %s\n", cudaGetErrorString(e)); return (int)e; }
    int sms = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, i)) != cudaSuccess) { fprintf(stderr, "cudaDeviceGetAttribute returned: %s\n", cudaGetErrorString(e)); return (int)e; }
    if (json) printf("%s{\"device\":%d,\"free_mib\":%zu,\"total_mib\":%zu,\"sm_count\":%d}", i ? "," : "", i, free_b >> 20, total_b >> 20, sms);
    else {
      printf("For device %d:  Free memory: %zu M, Total memory: %zu M\n", i, free_b >> 20, total_b >> 20);
      printf("For device %d:  multiProcessorCount: %d\n", i, sms);
    }
  }
  if (json) printf("]}\n");
  return 0;
}
