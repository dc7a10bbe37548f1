// Scratch: what NVLink gives to the access patterns of the partitioned trainer (2 GPUs, one process).
// (a) random 800-byte row gathers from peer memory with ld.global.cg.v4 (b) random row scatters with st (c) with red.add.v4.f32
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void rows(float* tab, int nrows, int P, int per_warp, int mode, int inflight, float* sink) {
   int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   float acc = 0.f;
   for (int i = 0; i < per_warp; i += inflight) {
      float4 v[4][2];
      for (int k = 0; k < inflight; k++) {
         uint32_t r = (uint32_t)((warp * 2654435761u + (i + k) * 40503u) * 2246822519u) % (uint32_t)nrows;
         float* p = tab + (size_t)r * P;
         for (int q = 0; q < 2; q++) {
            int off = (q * 32 + lane) * 4;
            if (off < P) {
               if (mode == 0) v[k][q] = __ldcg(reinterpret_cast<const float4*>(p + off));
               else if (mode == 1) __stcg(reinterpret_cast<float4*>(p + off), make_float4(1.f, 2.f, 3.f, 4.f));
               else asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p + off), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
            } else if (mode == 0) v[k][q] = make_float4(0, 0, 0, 0);
         }
      }
      if (mode == 0) for (int k = 0; k < inflight; k++) for (int q = 0; q < 2; q++) acc += v[k][q].x + v[k][q].w;
   }
   if (acc == 12345.678f) sink[0] = acc;
}

int main() {
   int n; CK(cudaGetDeviceCount(&n));
   if (n < 2) { printf("need 2 GPUs\n"); return 0; }
   const int nrows = 2000000, P = 200;
   float *t0, *t1, *sink;
   CK(cudaSetDevice(1)); CK(cudaMalloc(&t1, (size_t)nrows * P * 4)); CK(cudaMemset(t1, 0, (size_t)nrows * P * 4));
   CK(cudaSetDevice(0)); CK(cudaMalloc(&t0, (size_t)nrows * P * 4)); CK(cudaMemset(t0, 0, (size_t)nrows * P * 4)); CK(cudaMalloc(&sink, 64));
   CK(cudaDeviceEnablePeerAccess(1, 0));
   cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
   const char* names[3] = {"ld.cg.v4 gather", "st.cg.v4 scatter", "red.add.v4 scatter"};
   for (int remote = 0; remote < 2; remote++)
      for (int mode = 0; mode < 3; mode++)
         for (int inflight : {1, 4})
            for (int threads : {768}) {
               int per_warp = 64;
               int warps = 148 * threads / 32;
               float* tab = remote ? t1 : t0;
               rows<<<148, threads>>>(tab, nrows, P, per_warp, mode, inflight, sink);
               CK(cudaDeviceSynchronize());
               cudaEventRecord(e0);
               rows<<<148, threads>>>(tab, nrows, P, per_warp, mode, inflight, sink);
               cudaEventRecord(e1); CK(cudaDeviceSynchronize());
               float ms; cudaEventElapsedTime(&ms, e0, e1);
               double bytes = (double)warps * per_warp * P * 4;
               printf("%-6s %-20s rows in flight/warp %d: %.3f ms, %.1f GB/s\n", remote ? "PEER" : "local", names[mode], inflight, ms, bytes / ms / 1e6);
            }
   // sustained + bidirectional: both GPUs gather / scatter from each other at the same time, ~3 GB each
   CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0));
   float* sink1; CK(cudaMalloc(&sink1, 64));
   cudaEvent_t f0, f1; cudaEventCreate(&f0); cudaEventCreate(&f1);
   for (int mode = 0; mode < 3; mode++)
      for (int both = 0; both < 2; both++) {
         int per_warp = 1024, threads = 768, warps = 148 * threads / 32;
         CK(cudaSetDevice(0)); cudaEventRecord(e0);
         rows<<<148, threads>>>(t1, nrows, P, per_warp, mode, 1, sink);
         cudaEventRecord(e1);
         if (both) { CK(cudaSetDevice(1)); cudaEventRecord(f0); rows<<<148, threads>>>(t0, nrows, P, per_warp, mode, 1, sink1); cudaEventRecord(f1); }
         CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize());
         CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
         float ms, ms1 = 0; cudaEventElapsedTime(&ms, e0, e1); if (both) cudaEventElapsedTime(&ms1, f0, f1);
         double bytes = (double)warps * per_warp * P * 4;
         printf("sustained %-20s %s: GPU0 %.2f ms %.1f GB/s", names[mode], both ? "both directions" : "one direction ", ms, bytes / ms / 1e6);
         if (both) printf(" | GPU1 %.2f ms %.1f GB/s", ms1, bytes / ms1 / 1e6);
         printf("\n");
      }
   return 0;
}
