// Device-side building blocks shared by the training and ranking kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kb2e {

// ---- counter-based RNG: Philox4x32-10 (Salmon et al., SC'11) ------------------------------------
// Replaces the reference's global std::rand stream (common/utils.cpp:18-38,113-120); keyed by
// -seed, indexed by (sample, global batch, attempt, stream) so results do not depend on the launch
// geometry or the number of GPUs.  oracle/kb2e_oracle.c:orc_philox is the CPU twin.
__host__ __device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
   for (int round = 0; round < 10; round++) {
      uint64_t p0 = (uint64_t)0xD2511F53u * c0;
      uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
      uint32_t n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
      uint32_t n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
   }
   out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- triple hash set (open addressing, linear probing, 64-bit keys) -----------------------------
// Replaces std::map<pair<h,r>,map<t,int>> triples_ (common/trainer.h:49) for the sampler's rejection test.
constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kEntityBits = 24;  // entity ids < 16,777,216
constexpr int kRelationBits = 16;

__host__ __device__ __forceinline__ uint64_t pack_triple(int h, int r, int t) {
   return ((uint64_t)(uint32_t)r << (2 * kEntityBits)) | ((uint64_t)(uint32_t)h << kEntityBits) | (uint64_t)(uint32_t)t;
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // murmur3 fmix64
   x ^= x >> 33;
   x *= 0xff51afd7ed558ccdull;
   x ^= x >> 33;
   x *= 0xc4ceb9fe1a85ec53ull;
   x ^= x >> 33;
   return x;
}

__device__ __forceinline__ bool hash_contains(const uint64_t* __restrict__ table, uint64_t mask, uint64_t key) {
   uint64_t slot = mix64(key) & mask;
   while (true) {
      uint64_t v = __ldg(table + slot);
      if (v == key) return true;
      if (v == kEmptyKey) return false;
      slot = (slot + 1) & mask;
   }
}

// ---- memory access helpers ----------------------------------------------------------------------
// Embedding tables are rewritten by other SMs between the phases of one persistent launch, so
// they are always read/written through L2 (.cg), never through the non-coherent L1 path.
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_cg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
// The same load as ld_cg4, pinned in program order: the compiler sinks an ordinary load into the conditional block that
// uses it (measured in the list publish of train.cu: the row loads ended up BEHIND the claim they were meant to travel
// with, one extra L2 round trip per batch); a volatile asm stays where it is written.
__device__ __forceinline__ float4 ld_cg4_pinned(const float* p) {
   float4 v;
   asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
   return v;
}

__device__ __forceinline__ uint32_t ld_cg_u32_pinned(const uint32_t* p) {
   uint32_t v;
   asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
   return v;
}

// Vector reduction into global memory: one 16-byte RED instead of four scalar atomics (sm_90+).
__device__ __forceinline__ void red_add4(float* p, float4 v) {
   asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// read-only loads pinned in program order (the staged sampler of train_device.cuh issues them one phase ahead of their use)
__device__ __forceinline__ int4 ld_nc_int4_pinned(const int4* p) {
   int4 v;
   asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ uint64_t ld_nc_u64_pinned(const uint64_t* p) {
   uint64_t v;
   asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
   uint32_t v;
   asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
   return v;
}

__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
   asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
   for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) { return __umul64hi(a, b); }

}  // namespace kb2e
