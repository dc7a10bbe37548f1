// Scratch: peer row gathers through a CUDA VMM allocation shared across processes by POSIX fd
// (cuMemCreate / cuMemExportToShareableHandle / pidfd_getfd / cuMemImportFromShareableHandle / cuMemMap).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/syscall.h>
#include <sys/wait.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("%s: %s\n", #x, s); exit(1);} } while (0)

__global__ void gather(const float* tab, int nrows, int P, int per_warp, float* sink) {
   int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   float acc = 0.f;
   for (int i = 0; i < per_warp; i++) {
      uint32_t r = (uint32_t)((warp * 2654435761u + i * 40503u) * 2246822519u) % (uint32_t)nrows;
      const float* p = tab + (size_t)r * P;
      for (int q = 0; q < 2; q++) {
         int off = (q * 32 + lane) * 4;
         if (off < P) { float4 v = __ldcg(reinterpret_cast<const float4*>(p + off)); acc += v.x + v.w; }
      }
   }
   if (acc == 12345.678f) sink[0] = acc;
}

int main() {
   const int nrows = 2000000, P = 200;
   int to_child[2], to_parent[2];
   if (pipe(to_child) || pipe(to_parent)) return 1;
   pid_t pid = fork();
   CUmemAllocationProp prop = {};
   prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
   prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
   prop.location.id = 1;
   prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
   size_t gran = 0;
   if (pid == 0) {
      CK(cudaSetDevice(1)); CK(cudaFree(0));
      CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
      size_t bytes = ((size_t)nrows * P * 4 + gran - 1) / gran * gran;
      CUmemGenericAllocationHandle h; CU(cuMemCreate(&h, bytes, &prop, 0));
      int fd = -1; CU(cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
      long msg[3] = {(long)fd, (long)bytes, (long)gran};
      if (write(to_parent[1], msg, sizeof(msg)) != sizeof(msg)) return 1;
      char c; if (read(to_child[0], &c, 1) != 1) return 1;
      return 0;
   }
   CK(cudaSetDevice(0)); CK(cudaFree(0));
   long msg[3];
   if (read(to_parent[0], msg, sizeof(msg)) != sizeof(msg)) return 1;
   int pidfd = (int)syscall(SYS_pidfd_open, pid, 0);
   int fd = (int)syscall(438 /* pidfd_getfd */, pidfd, (int)msg[0], 0);
   printf("granularity %ld, bytes %ld, pidfd %d, fd %d\n", msg[2], msg[1], pidfd, fd);
   if (fd < 0) { perror("pidfd_getfd"); return 1; }
   CUmemGenericAllocationHandle h;
   CU(cuMemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
   CUdeviceptr ptr; CU(cuMemAddressReserve(&ptr, (size_t)msg[1], 0, 0, 0));
   CU(cuMemMap(ptr, (size_t)msg[1], 0, h, 0));
   CUmemAccessDesc acc = {};
   acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
   CU(cuMemSetAccess(ptr, (size_t)msg[1], &acc, 1));
   float* sink; CK(cudaMalloc(&sink, 64));
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   for (int rep = 0; rep < 2; rep++) {
      int per_warp = 512, threads = 768, warps = 148 * threads / 32;
      cudaEventRecord(e0);
      gather<<<148, threads>>>((const float*)ptr, nrows, P, per_warp, sink);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("VMM (fd-shared) peer gather: %.2f ms, %.1f GB/s\n", ms, (double)warps * per_warp * P * 4 / ms / 1e6);
   }
   char c = 1; if (write(to_child[1], &c, 1) != 1) return 1;
   waitpid(pid, nullptr, 0);
   return 0;
}
