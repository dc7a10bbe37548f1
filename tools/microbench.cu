// Scratch microbenchmarks that size the design constants of the persistent training kernel on the
// actual B200: L2 load latency (dependent chain), grid-barrier cost, RED.128 throughput.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// 1. pointer chase through L2 (.cg) over a buffer of `n` ints
__global__ void chase(const int* p, int start, int iters, long long* out) {
   int idx = start;
   long long t0 = clock64();
   for (int i = 0; i < iters; i++) idx = __ldcg(p + idx);
   long long t1 = clock64();
   out[0] = t1 - t0; out[1] = idx;
}

__device__ __forceinline__ uint32_t ld_acq(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_rlx(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// 2. grid barrier variants, `iters` barriers back to back; mode 0: fence+red.release+acquire spin+fence
//    mode 1: red.release + relaxed spin + one fence; mode 2: atomicAdd relaxed + relaxed spin (no ordering)
__global__ void barrier_test(uint32_t* counter, int iters, int mode, long long* out) {
   uint32_t target = 0;
   long long t0 = clock64();
   for (int i = 0; i < iters; i++) {
      __syncthreads();
      if (threadIdx.x == 0) {
         target += gridDim.x;
         if (mode == 0) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(counter), "r"(1u) : "memory");
            while ((int32_t)(ld_acq(counter) - target) < 0) {}
            __threadfence();
         } else if (mode == 1) {
            asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(counter), "r"(1u) : "memory");
            while ((int32_t)(ld_rlx(counter) - target) < 0) {}
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
         } else {
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l"(counter), "r"(1u) : "memory");
            while ((int32_t)(ld_rlx(counter) - target) < 0) {}
         }
      }
      __syncthreads();
   }
   long long t1 = clock64();
   if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// 2b. hierarchical variant (round-1 review suggestion): barrier.cluster over a 4-CTA cluster, ONE arrival per cluster on the
//     global counter (37 instead of 148), the cluster leader polls, a second barrier.cluster releases its three siblings
__global__ void barrier_cluster_test(uint32_t* counter, int iters, long long* out) {
   uint32_t target = 0;
   uint32_t rank;
   asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
   const uint32_t n_clusters = gridDim.x / 4;
   long long t0 = clock64();
   for (int i = 0; i < iters; i++) {
      // every thread of the cluster takes part in barrier.cluster (it counts threads, not CTAs): it stands in for bar.sync
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      target += n_clusters;
      if (rank == 0 && threadIdx.x == 0) {
         asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(counter), "r"(1u) : "memory");
         while ((int32_t)(ld_rlx(counter) - target) < 0) {}
         asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
   }
   long long t1 = clock64();
   if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

// 3. RED.128 throughput: every thread adds float4 to pseudo-random rows of a table [rows][pitch]
__global__ void red_test(float* tab, int rows, int pitch4, int per_thread, long long* out) {
   int tid = blockIdx.x * blockDim.x + threadIdx.x;
   int lane = threadIdx.x & 31;
   int warp = tid >> 5;
   long long t0 = clock64();
   for (int i = 0; i < per_thread; i++) {
      uint32_t r = (uint32_t)(warp * 2654435761u + i * 40503u) % (uint32_t)rows;
      if (lane < pitch4) {
         float* p = tab + ((size_t)r * pitch4 + lane) * 4;
         asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
      }
   }
   __threadfence();
   long long t1 = clock64();
   if (tid == 0) out[0] = t1 - t0;
}

// 3b. same with plain vector stores, and with scalar REDs, for comparison
__global__ void st_test(float* tab, int rows, int pitch4, int per_thread, int scalar_red, long long* out) {
   int tid = blockIdx.x * blockDim.x + threadIdx.x;
   int lane = threadIdx.x & 31;
   int warp = tid >> 5;
   long long t0 = clock64();
   for (int i = 0; i < per_thread; i++) {
      uint32_t r = (uint32_t)(warp * 2654435761u + i * 40503u) % (uint32_t)rows;
      if (lane < pitch4) {
         float* p = tab + ((size_t)r * pitch4 + lane) * 4;
         if (scalar_red) { atomicAdd(p, 1.f); atomicAdd(p + 1, 1.f); atomicAdd(p + 2, 1.f); atomicAdd(p + 3, 1.f); }
         else __stcg(reinterpret_cast<float4*>(p), make_float4(1.f, 1.f, 1.f, 1.f));
      }
   }
   __threadfence();
   long long t1 = clock64();
   if (tid == 0) out[0] = t1 - t0;
}

int main() {
   cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
   int sms = prop.multiProcessorCount;
   double ghz = prop.clockRate * 1e-6;
   printf("%s, %d SMs, %.2f GHz nominal\n", prop.name, sms, ghz);
   long long* out; CK(cudaMallocManaged(&out, 64));
   // chase
   for (int mb : {1, 16, 64, 256}) {
      int n = mb * 1024 * 1024 / 4;
      int* h = (int*)malloc((size_t)n * 4);
      // random cycle with stride so that each hop is a different line
      uint32_t x = 12345; for (int i = 0; i < n; i++) h[i] = 0;
      int cur = 0; int hops = 20000;
      for (int i = 0; i < hops; i++) { x = x * 1664525u + 1013904223u; int nxt = (int)((x >> 4) % (uint32_t)n); h[cur] = nxt; cur = nxt; }
      int* d; CK(cudaMalloc(&d, (size_t)n * 4)); CK(cudaMemcpy(d, h, (size_t)n * 4, cudaMemcpyHostToDevice));
      chase<<<1, 1>>>(d, 0, 4000, out); CK(cudaDeviceSynchronize());
      chase<<<1, 1>>>(d, 0, 4000, out); CK(cudaDeviceSynchronize());
      printf("ld.cg dependent chain over %3d MB: %.0f cycles/hop\n", mb, (double)out[0] / 4000);
      cudaFree(d); free(h);
   }
   // barrier
   uint32_t* counter; CK(cudaMalloc(&counter, 4));
   for (int threads : {1024, 256}) for (int mode = 0; mode < 3; mode++) {
      CK(cudaMemset(counter, 0, 4));
      void* args[] = {&counter, nullptr, &mode, &out};
      int iters = 2000; args[1] = &iters;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      CK(cudaLaunchCooperativeKernel((void*)barrier_test, dim3(sms), dim3(threads), args, 0, 0));
      cudaEventRecord(e1); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("grid barrier (%d CTAs x %d thr, mode %d): %.0f ns each (%.0f cycles)\n", sms, threads, mode, ms * 1e6 / iters, (double)out[0] / iters);
   }
   for (int threads : {1024, 256}) {
      // 148 = 37 clusters of 4; every CTA is resident (one per SM), so the plain launch is as good as a cooperative one here
      CK(cudaMemset(counter, 0, 4));
      int iters = 2000;
      // a spinning barrier needs every cluster resident at once: ask how many fit (GPCs with an SM count that is not a
      // multiple of 4 leave SMs unused) and launch no more than that
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((sms / 4) * 4); cfg.blockDim = dim3(threads);
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int max_clusters = 0;
      CK(cudaOccupancyMaxActiveClusters(&max_clusters, barrier_cluster_test, &cfg));
      const int clusters = max_clusters < sms / 4 ? max_clusters : sms / 4;
      if (clusters < 1) { printf("cluster barrier: no resident cluster\n"); break; }
      cfg.gridDim = dim3(clusters * 4);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      CK(cudaLaunchKernelEx(&cfg, barrier_cluster_test, counter, iters, out));
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("cluster barrier launch: %s\n", cudaGetErrorString(err)); break; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("grid barrier (%d CTAs x %d thr in %d resident clusters of 4 (of %d wanted): barrier.cluster + one arrival per cluster): %.0f ns each (%.0f cycles)\n",
             clusters * 4, threads, clusters, sms / 4, ms * 1e6 / iters, (double)out[0] / iters);
   }
   // RED throughput: rows of 100 floats (25 float4), table of 16296 rows, total ~ 480k RED.128 (one batch at alpha=1)
   int rows = 16296, pitch4 = 25;
   float* tab; CK(cudaMalloc(&tab, (size_t)rows * pitch4 * 16)); CK(cudaMemset(tab, 0, (size_t)rows * pitch4 * 16));
   for (int per_thread : {4, 16}) {
      for (int variant = 0; variant < 3; variant++) {
         cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
         for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (variant == 0) red_test<<<sms, 1024>>>(tab, rows, pitch4, per_thread, out);
            else st_test<<<sms, 1024>>>(tab, rows, pitch4, per_thread, variant == 2, out);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
         }
         float ms; cudaEventElapsedTime(&ms, e0, e1);
         double n = (double)sms * 32 * per_thread * pitch4;  // 16-byte ops
         printf("%s x %.0fk 16-byte ops: %.1f us -> %.1f G ops/s, %.0f GB/s\n", variant == 0 ? "RED.128    " : variant == 1 ? "ST.128     " : "4xRED.32   ",
                n / 1e3, ms * 1e3, n / ms / 1e6, n * 16 / ms / 1e6);
      }
   }
   return 0;
}
