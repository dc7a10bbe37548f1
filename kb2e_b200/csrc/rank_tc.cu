// Squared-L2 all-entity scoring on the 5th-generation tensor cores (tcgen05 + TMEM) with a fused
// compare-and-count epilogue: the scores never reach HBM.
//
// Replaces the score/sort/scan of common/evaluation.cpp:129-164 for the squared-L2 scorers
// (transe/transe.cpp:22-24) where the path is a dense contraction:
//     E(c) = sum_i ((V_i - C[c]_i) - d'_i)^2 = |u|^2 + n_c - 2 u.C[c],   u = V - d',  n_c = |C[c]|^2
// so "candidate c ranks before the truth" <=> s(c) = n_c - 2 u.C[c] < T = E_true - |u|^2.
// G = U x C^T is a (queries x entities x D) GEMM.  Precision: both operands are split into two
// bf16 terms (x = hi + lo, |x - hi - lo| <= 2^-17 |x|) and three products hi*hi + hi*lo + lo*hi are
// accumulated in fp32 in TMEM (relative error <= 2^-15 of sum |u_i||c_i|).  Every candidate whose
// approximate score lies within a proven band around T is re-scored in exact fp64 with the
// reference's operation order (recheck_kernel in rank.cu), so the integer ranks stay bit-exact.
//
// Kernel shape (one CTA = 128 queries x a range of 128-candidate tiles, 4 warps, 2 CTAs per SM):
//   A (u_hi, u_lo) is loaded once into shared memory, K-major, no swizzle: core matrix = 8 rows x 16 B,
//   laid out [16-byte K chunk][row] so LBO = 2048 B (next K chunk), SBO = 128 B (next 8 rows);
//   per candidate tile: B (c_hi, c_lo) + n_c -> shared memory; one elected thread issues
//   21 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) into a 128-column TMEM accumulator
//   and commits to an mbarrier; each warp then reads its 32 TMEM lanes (= 32 queries) with
//   tcgen05.ld 32x32b.x32 and counts / collects band candidates.  With two CTAs resident per SM one
//   CTA's epilogue and loads overlap the other's MMAs.

#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "rank_tc.h"

namespace kb2e {
namespace tc {

constexpr int BM = 128;          // queries per CTA (TMEM lanes)
constexpr int BN = 128;          // candidates per tile (TMEM columns)
constexpr int KC = kRowChunks;   // 16-byte chunks per operand row (8 bf16 each)
constexpr int KSTEPS = KC / 2;   // one MMA consumes K = 16 bf16 = 2 chunks
constexpr int TILE_BYTES = BM * KC * 16;
constexpr int SMEM_BYTES = 4 * TILE_BYTES + BN * 4 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE, version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
   uint64_t d = 0;
   d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
   d |= (uint64_t)((BM * 16) >> 4) << 16;      // leading byte offset: next 16-byte K chunk
   d |= (uint64_t)(128 >> 4) << 32;            // stride byte offset: next group of 8 rows
   d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
   return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
   asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
   uint32_t ok;
   do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
   } while (!ok);
}

// global [rows][KC*8] bf16 row-major -> shared [chunk][row] (128 rows)
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ g, long long row0, unsigned char* s) {
   const uint4* src = reinterpret_cast<const uint4*>(g) + row0 * KC;
   for (int item = threadIdx.x; item < BM * KC; item += blockDim.x) {
      int row = item / KC, kc = item - row * KC;
      uint4 v = __ldg(src + item);
      *reinterpret_cast<uint4*>(s + kc * (BM * 16) + row * 16) = v;
   }
}

__global__ void __launch_bounds__(128, 2) rank_l2_tc_kernel(const TcArgs a) {
   extern __shared__ __align__(128) unsigned char smem[];
   unsigned char* sA_hi = smem;
   unsigned char* sA_lo = smem + TILE_BYTES;
   unsigned char* sB_hi = smem + 2 * TILE_BYTES;
   unsigned char* sB_lo = smem + 3 * TILE_BYTES;
   float* s_nc = reinterpret_cast<float*>(smem + 4 * TILE_BYTES);
   uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + 4 * TILE_BYTES + BN * 4);
   uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + 4 * TILE_BYTES + BN * 4 + 16);

   const int warp = threadIdx.x >> 5;
   const long long q0 = (long long)blockIdx.x * BM;
   const int tiles_total = a.n_pad / BN;
   const int t_begin = (int)((long long)tiles_total * blockIdx.y / gridDim.y);
   const int t_end = (int)((long long)tiles_total * (blockIdx.y + 1) / gridDim.y);

   if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"((uint32_t)BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
   }
   if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(s_bar)), "r"(1u) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   load_tile(a.u_hi, q0, sA_hi);
   load_tile(a.u_lo, q0, sA_lo);
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   __syncthreads();
   asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
   const uint32_t tmem_base = *s_tmem;

   // this thread's query
   const long long q = q0 + threadIdx.x;
   const float thr_lo = a.thr_lo[q];   // s <  thr_lo            -> certainly ranked before the truth
   const float thr_hi = a.thr_hi[q];   // thr_lo <= s <= thr_hi  -> exact fp64 recheck
   int less = 0;
   uint32_t parity = 0;

   for (int t = t_begin; t < t_end; t++) {
      const long long c0 = (long long)t * BN;
      load_tile(a.c_hi, c0, sB_hi);
      load_tile(a.c_lo, c0, sB_lo);
      if (threadIdx.x < BN) s_nc[threadIdx.x] = __ldg(a.n_c + c0 + threadIdx.x);
      // generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
         asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
         const uint32_t ah = smem_u32(sA_hi), al = smem_u32(sA_lo), bh = smem_u32(sB_hi), bl = smem_u32(sB_lo);
#pragma unroll
         for (int ks = 0; ks < KSTEPS; ks++) {
            const uint32_t off = ks * 2 * (BM * 16);
            mma_bf16(tmem_base, make_desc(ah + off), make_desc(bh + off), ks > 0 ? 1u : 0u);
            mma_bf16(tmem_base, make_desc(ah + off), make_desc(bl + off), 1u);
            mma_bf16(tmem_base, make_desc(al + off), make_desc(bh + off), 1u);
         }
         // arrives on the mbarrier when every MMA above has completed (implies fence::before_thread_sync)
         asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(s_bar)) : "memory");
      }
      mbar_wait(smem_u32(s_bar), parity);
      parity ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // epilogue: warp w owns TMEM lanes 32w .. 32w+31 (= queries), 32 columns (= candidates) per load
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
         uint32_t v[32];
         const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
         asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr) : "memory");
         asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
         for (int j = 0; j < 32; j++) {
            const float s = fmaf(-2.f, __uint_as_float(v[j]), s_nc[cb + j]);
            if (s < thr_lo) {
               less++;
            } else if (s <= thr_hi) {
               unsigned int slot = atomicAdd(a.band_count, 1u);
               if (slot < a.band_cap) a.band[slot] = make_int2((int)q, (int)(c0 + cb + j));
            }
         }
      }
      // TMEM and the B buffers are overwritten by the next tile
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
   }
   if (q < a.nq) {
      if (gridDim.y == 1) a.q_less[q] = less;
      else if (less) atomicAdd(a.q_less + q, less);
   }
   __syncthreads();
   if (warp == 0) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)BN) : "memory");
   }
}

// ---- operand preparation ---------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(double x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
   float xf = (float)x;
   hi = __float2bfloat16_rn(xf);
   lo = __float2bfloat16_rn(xf - __bfloat162float(hi));
}

// candidates: C (fp64 [n][D] row-major) -> c_hi, c_lo (bf16 [n_pad][KC*8]), n_c = |c|^2, and max |c|
__global__ void prep_candidates_kernel(const double* __restrict__ c64, int n, int n_pad, int D,
                                       __nv_bfloat16* c_hi, __nv_bfloat16* c_lo, float* n_c, unsigned int* cmax_bits) {
   int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (row >= n_pad) return;
   double s = 0.0;
   for (int i = lane; i < KC * 8; i += 32) {
      double x = (row < n && i < D) ? c64[(size_t)row * D + i] : 0.0;
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      c_hi[(size_t)row * (KC * 8) + i] = h;
      c_lo[(size_t)row * (KC * 8) + i] = l;
      s += x * x;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   if (lane == 0) {
      n_c[row] = row < n ? (float)s : __int_as_float(0x7f800000);  // padding never counts
      if (row < n) atomicMax(cmax_bits, __float_as_uint((float)sqrt(s) * 1.0000002f));
   }
}

// queries: u = V - d' (exact fp64) -> u_hi, u_lo; thresholds from E_true and |u|^2
__global__ void prep_queries_kernel(const double* __restrict__ c64, const double* __restrict__ rel64,
                                    const int32_t* q_fixed, const int32_t* q_rel, const int32_t* q_side, const double* q_etrue,
                                    long long nq, long long q_pad, int D, const unsigned int* cmax_bits,
                                    __nv_bfloat16* u_hi, __nv_bfloat16* u_lo, float* thr_lo, float* thr_hi) {
   long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (q >= q_pad) return;
   double s = 0.0;
   const bool real = q < nq;
   const double* v = real ? c64 + (size_t)q_fixed[q] * D : nullptr;
   const double* d = real ? rel64 + (size_t)q_rel[q] * D : nullptr;
   const double sign = real && q_side[q] ? -1.0 : 1.0;
   for (int i = lane; i < KC * 8; i += 32) {
      double x = (real && i < D) ? v[i] - sign * d[i] : 0.0;
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      u_hi[(size_t)q * (KC * 8) + i] = h;
      u_lo[(size_t)q * (KC * 8) + i] = l;
      s += x * x;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   if (lane == 0) {
      if (!real) {
         thr_lo[q] = __int_as_float(0xff800000);  // -inf: padding rows never count, never recheck
         thr_hi[q] = __int_as_float(0xff800000);
      } else {
         const double cmax = (double)__uint_as_float(*cmax_bits);
         const double T = q_etrue[q] - s;
         // |s~ - s| <= 2 * 1.9e-5 * |u| |c| (three-product bf16 split) + ~4e-5 |u| |c| (fp32 accumulation of
         // 336 products) + fp32 rounding of n_c and of the final fma; the band 2^-12 |u| |c|max is > 2x that.
         const double unorm = sqrt(s);
         const double delta = 8.0 * (1.0 / 32768.0) * unorm * cmax + 6e-6 * (1.0 + unorm * cmax + cmax * cmax);
         thr_lo[q] = __double2float_rd(T - delta);
         thr_hi[q] = __double2float_ru(T + delta);
      }
   }
}

}  // namespace tc

// ---- host -------------------------------------------------------------------------------------------
static inline unsigned nblk2(long long n, int t) { return (unsigned)((n + t - 1) / t); }

bool tc_supported(const kb2e_ctx* c) {
   return c->cfg.model == KB2E_MODEL_TRANSE && c->cfg.distance == KB2E_DISTANCE_L2 && c->D <= tc::kRowChunks * 8 &&
          !(c->cfg.flags & KB2E_FLAG_RANK_EXACT_ONLY);
}

int tc_prepare_candidates(kb2e_ctx* c, TcState* s) {
   const int n_pad = ((c->nE + tc::BN - 1) / tc::BN) * tc::BN;
   if (n_pad != s->n_pad) {
      cudaFree(s->c_hi); cudaFree(s->c_lo); cudaFree(s->n_c);
      s->c_hi = s->c_lo = nullptr; s->n_c = nullptr;
      KB2E_CUDA(c, cudaMalloc(&s->c_hi, (size_t)n_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, cudaMalloc(&s->c_lo, (size_t)n_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, cudaMalloc(&s->n_c, (size_t)n_pad * sizeof(float)));
      s->n_pad = n_pad;
   }
   if (!s->scalars) KB2E_CUDA(c, cudaMalloc(&s->scalars, 4 * sizeof(unsigned int)));
   KB2E_CUDA(c, cudaMemsetAsync(s->scalars, 0, 4 * sizeof(unsigned int), c->stream));
   tc::prep_candidates_kernel<<<nblk2((long long)n_pad * 32, 256), 256, 0, c->stream>>>(
      c->ent64, c->nE, n_pad, c->D, (__nv_bfloat16*)s->c_hi, (__nv_bfloat16*)s->c_lo, s->n_c, s->scalars);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int tc_run(kb2e_ctx* c, TcState* s, const int32_t* q_fixed, const int32_t* q_rel, const int32_t* q_side, const double* q_etrue,
           long long nq, int32_t* q_less, bool* overflow) {
   const long long q_pad = ((nq + tc::BM - 1) / tc::BM) * tc::BM;
   if (q_pad > s->q_cap) {
      cudaFree(s->u_hi); cudaFree(s->u_lo); cudaFree(s->thr_lo); cudaFree(s->thr_hi); cudaFree(s->band);
      s->u_hi = s->u_lo = nullptr; s->thr_lo = s->thr_hi = nullptr; s->band = nullptr;
      s->q_cap = 0;
      KB2E_CUDA(c, cudaMalloc(&s->u_hi, (size_t)q_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, cudaMalloc(&s->u_lo, (size_t)q_pad * tc::kRowChunks * 16));
      KB2E_CUDA(c, cudaMalloc(&s->thr_lo, (size_t)q_pad * sizeof(float)));
      KB2E_CUDA(c, cudaMalloc(&s->thr_hi, (size_t)q_pad * sizeof(float)));
      s->band_cap = (unsigned int)std::min<long long>(std::max<long long>(1 << 20, 64 * q_pad), 1ll << 28);
      KB2E_CUDA(c, cudaMalloc(&s->band, (size_t)s->band_cap * sizeof(int2)));
      s->q_cap = q_pad;
   }
   unsigned int* band_count = s->scalars + 1;
   KB2E_CUDA(c, cudaMemsetAsync(band_count, 0, sizeof(unsigned int), c->stream));
   tc::prep_queries_kernel<<<nblk2(q_pad * 32, 256), 256, 0, c->stream>>>(
      c->ent64, c->rel64, q_fixed, q_rel, q_side, q_etrue, nq, q_pad, c->D, s->scalars,
      (__nv_bfloat16*)s->u_hi, (__nv_bfloat16*)s->u_lo, s->thr_lo, s->thr_hi);
   TcArgs a;
   a.u_hi = (const __nv_bfloat16*)s->u_hi; a.u_lo = (const __nv_bfloat16*)s->u_lo;
   a.c_hi = (const __nv_bfloat16*)s->c_hi; a.c_lo = (const __nv_bfloat16*)s->c_lo;
   a.n_c = s->n_c; a.thr_lo = s->thr_lo; a.thr_hi = s->thr_hi;
   a.q_less = q_less; a.band = s->band; a.band_count = band_count; a.band_cap = s->band_cap;
   a.nq = nq; a.n_pad = s->n_pad;
   const unsigned mtiles = (unsigned)(q_pad / tc::BM);
   const int ntiles = s->n_pad / tc::BN;
   long long splits = std::max<long long>(1, (2ll * 2 * c->num_sms + mtiles - 1) / mtiles);
   splits = std::min<long long>(splits, ntiles);
   KB2E_CUDA(c, cudaFuncSetAttribute(tc::rank_l2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
   if (splits > 1) KB2E_CUDA(c, cudaMemsetAsync(q_less, 0, (size_t)nq * sizeof(int32_t), c->stream));
   KB2E_CUDA(c, cudaEventRecord(s->e0, c->stream));
   tc::rank_l2_tc_kernel<<<dim3(mtiles, (unsigned)splits), 128, tc::SMEM_BYTES, c->stream>>>(a);
   KB2E_CUDA(c, cudaEventRecord(s->e1, c->stream));
   KB2E_CUDA(c, cudaGetLastError());
   unsigned int count = 0;
   KB2E_CUDA(c, cudaMemcpyAsync(&count, band_count, sizeof(count), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, s->e0, s->e1));
   s->last_ms = ms;
   s->last_band = count;
   *overflow = count > s->band_cap;
   return KB2E_OK;
}

int tc_init(kb2e_ctx* c, TcState* s) {
   if (!s->e0) {
      KB2E_CUDA(c, cudaEventCreate(&s->e0));
      KB2E_CUDA(c, cudaEventCreate(&s->e1));
   }
   return KB2E_OK;
}

void tc_free(TcState* s) {
   cudaFree(s->c_hi); cudaFree(s->c_lo); cudaFree(s->n_c); cudaFree(s->u_hi); cudaFree(s->u_lo);
   cudaFree(s->thr_lo); cudaFree(s->thr_hi); cudaFree(s->band); cudaFree(s->scalars);
   if (s->e0) { cudaEventDestroy(s->e0); cudaEventDestroy(s->e1); }
   *s = TcState();
}

}  // namespace kb2e
