// Scratch: peer row gathers through a CUDA-IPC mapping (two processes) vs cudaDeviceEnablePeerAccess (one process).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

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
   if (pid == 0) {   // child: owns the table on GPU 1
      CK(cudaSetDevice(1));
      float* t; CK(cudaMalloc(&t, (size_t)nrows * P * 4)); CK(cudaMemset(t, 0, (size_t)nrows * P * 4));
      cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, t));
      if (write(to_parent[1], &h, sizeof(h)) != sizeof(h)) return 1;
      char c; if (read(to_child[0], &c, 1) != 1) return 1;   // wait until the parent is done
      return 0;
   }
   CK(cudaSetDevice(0));
   cudaIpcMemHandle_t h;
   if (read(to_parent[0], &h, sizeof(h)) != sizeof(h)) return 1;
   void* p = nullptr;
   CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
   float* sink; CK(cudaMalloc(&sink, 64));
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   for (int rep = 0; rep < 2; rep++) {
      int per_warp = 512, threads = 768, warps = 148 * threads / 32;
      cudaEventRecord(e0);
      gather<<<148, threads>>>((const float*)p, nrows, P, per_warp, sink);
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("IPC-mapped peer gather: %.2f ms, %.1f GB/s\n", ms, (double)warps * per_warp * P * 4 / ms / 1e6);
   }
   char c = 1; if (write(to_child[1], &c, 1) != 1) return 1;
   waitpid(pid, nullptr, 0);
   return 0;
}
