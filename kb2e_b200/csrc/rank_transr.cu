// TransR ranking: the batched projection P_r = E M_r (every entity under every relation of the test set) on the
// 5th-generation tensor cores, and the exact fp64 pieces the integer ranks still need, computed on demand.
//
// Replaces, for TransR (citations into eriq-augustine/KB2E):
//   transr::tripleEnergy's projection loop   transr/transr.cpp:20-25   (headVec[i] += M[r][j][i] * e[j], per candidate)
//   called N_E times per query by            common/evaluation.cpp:129-136 through transr/evaluation.cpp:26-32
//
// The reference projects every candidate for every query (2 D^2 flops each).  All candidates of one relation share M_r,
// so the projection is a dense contraction  P_r[c][i] = sum_j E[c][j] M_r[j][i]:  an (N_E x D x D) GEMM per relation,
// batched over the relations of the pass.  The L1 (or squared-L2) scoring of the projected candidates is NOT a contraction
// and stays on the fp32 CUDA-core pre-filter (rank_f32.cu), which only needs the projected matrix in fp32 together with a
// rigorous bound on its error:
//   * operands split into bf16 hi + lo (|x - hi - lo| <= 2^-18 |x|), three products hi*hi + hi*lo + lo*hi accumulated in
//     fp32 in TMEM:  |P~ - P| <= eps_p * sum_j |e_j| |M_ji|  with eps_p = trp_eps(D) (split error + a worst-case model of
//     the tensor core's truncating accumulation; checked against exact values by tests/test_gpu_rank.py through
//     kb2e_debug_transr_projection).  Per projected row:  |eta|_1 <= eps_p |e|_2 sqrt(sum_j |M_j.|_1^2),
//     |eta|_2 <= eps_p |e|_2 |M|_F  (Cauchy-Schwarz) -- these widen the pre-filter's undecided band;
//   * everything that decides an integer rank is exact fp64 with the reference's operation order (j ascending from a zero
//     accumulator, separate multiply and add), projected ON DEMAND: the query's fixed and true entities (query_kernel),
//     the candidates left in the undecided band (recheck_kernel) and the known-true neighbours of the filter pass
//     (filter_kernel).  About 2 + 3 + 15 exact projections per query instead of 14,951.
//
// Kernel shape (project_tc_kernel): one CTA = one 128-entity operand tile (A: loaded once) x a range of relations
// (B = M_r^T tiles streamed through a 2-stage ring of 1-D bulk copies); one elected thread issues 3 * K/16
// tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = D rounded up to 16, K = 16) per relation into one of two TMEM
// accumulators; four epilogue warps read their 32 TMEM lanes (= 32 entities) with tcgen05.ld and store the D projected
// values as coalesced 128-byte rows of the transposed fp32 candidate matrix [slot][D][ld] -- the layout rank_f32_kernel
// reads.  Bound: HBM writes (N_E * D * 4 bytes per relation; the MMAs are ~10 % of the time).

#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "rank_transr.h"

namespace kb2e {
namespace trp {

constexpr int BM = 128;          // entities per operand tile (TMEM lanes)
constexpr int STAGES = 2;        // B-tile ring and TMEM accumulator ring
constexpr int EPI_WARPS = 4;
constexpr int THREADS = 32 * (EPI_WARPS + 2);   // epilogue warps, copy producer, MMA issuer
constexpr int EX_WARPS = 8;      // warps per CTA of the exact kernels
constexpr int MAXD = 128;

// ---- PTX wrappers (same forms as rank_tc.cu, with run-time tile shapes) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrix = 8 rows x 16 bytes; leading byte offset = distance
// between core matrices along K (rows * 16 bytes in the [chunk][row][16 B] tile layout), stride byte offset = 128 bytes
// (the next 8 rows).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
   uint64_t d = 0;
   d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
   d |= (uint64_t)(lbo_bytes >> 4) << 16;
   d |= (uint64_t)(128 >> 4) << 32;
   d |= (uint64_t)1 << 46;
   return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
   asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
   uint32_t ok;
   do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
   } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
   asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

struct ProjArgs {
   const unsigned char* e_hi;   // [tile][kc][128][16 B]
   const unsigned char* e_lo;
   const unsigned char* m_hi;   // [slot][kc][ncols][16 B]
   const unsigned char* m_lo;
   float* out;                  // [slot][D][ld]
   int slots, D, ld, kc, ncols, acc_stride;
   uint32_t idesc, tmem_cols;
};

// ---- the projection kernel ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) project_tc_kernel(const ProjArgs a) {
   extern __shared__ __align__(128) unsigned char smem[];
   const uint32_t a_bytes = (uint32_t)BM * a.kc * 16u;
   const uint32_t b_bytes = (uint32_t)a.ncols * a.kc * 16u;
   unsigned char* sA = smem;                      // [hi | lo]
   unsigned char* sB = smem + 2 * a_bytes;        // STAGES x [hi | lo]
   uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * a_bytes + 2 * STAGES * b_bytes);
   uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 12);
   const uint32_t bar_a = smem_u32(bars + 0);
   const uint32_t bar_full = smem_u32(bars + 1);      // + stage
   const uint32_t bar_empty = smem_u32(bars + 3);     // + stage
   const uint32_t bar_tfull = smem_u32(bars + 5);     // + accumulator
   const uint32_t bar_tempty = smem_u32(bars + 7);    // + accumulator

   const int warp = threadIdx.x >> 5;
   const int lane = threadIdx.x & 31;
   const int tile = blockIdx.x;
   const int s_begin = (int)((long long)a.slots * blockIdx.y / gridDim.y);
   const int s_end = (int)((long long)a.slots * (blockIdx.y + 1) / gridDim.y);
   const int n_iter = s_end - s_begin;

   if (threadIdx.x == 0) {
      mbar_init(bar_a, 1);
      for (int s = 0; s < STAGES; s++) {
         mbar_init(bar_full + 8 * s, 1);
         mbar_init(bar_empty + 8 * s, 1);
         mbar_init(bar_tfull + 8 * s, 1);
         mbar_init(bar_tempty + 8 * s, EPI_WARPS);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (warp == EPI_WARPS + 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(a.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
   }
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   __syncthreads();
   asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
   const uint32_t tmem_base = *s_tmem;

   if (warp == EPI_WARPS) {
      // ===== copy producer: operand tiles lie in global memory already in their shared-memory layout =====
      if (lane == 0) {
         mbar_expect_tx(bar_a, 2 * a_bytes);
         bulk_copy(smem_u32(sA), a.e_hi + (size_t)tile * a_bytes, a_bytes, bar_a);
         bulk_copy(smem_u32(sA + a_bytes), a.e_lo + (size_t)tile * a_bytes, a_bytes, bar_a);
         for (int i = 0; i < n_iter; i++) {
            const int s = i & 1;
            mbar_wait(bar_empty + 8 * s, ((i >> 1) & 1) ^ 1);
            const size_t off = (size_t)(s_begin + i) * b_bytes;
            mbar_expect_tx(bar_full + 8 * s, 2 * b_bytes);
            bulk_copy(smem_u32(sB + (2 * s) * b_bytes), a.m_hi + off, b_bytes, bar_full + 8 * s);
            bulk_copy(smem_u32(sB + (2 * s + 1) * b_bytes), a.m_lo + off, b_bytes, bar_full + 8 * s);
         }
      }
   } else if (warp == EPI_WARPS + 1) {
      // ===== MMA issuer =====
      if (lane == 0) {
         mbar_wait(bar_a, 0);
         const uint32_t ah = smem_u32(sA), al = smem_u32(sA + a_bytes);
         const uint32_t a_lbo = BM * 16u, b_lbo = (uint32_t)a.ncols * 16u;
         const int ksteps = a.kc / 2;   // one MMA consumes K = 16 bf16 = two 16-byte chunks
         for (int i = 0; i < n_iter; i++) {
            const int s = i & 1;
            const uint32_t ph = (i >> 1) & 1;
            mbar_wait(bar_tempty + 8 * s, ph ^ 1);
            mbar_wait(bar_full + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bh = smem_u32(sB + (2 * s) * b_bytes), bl = smem_u32(sB + (2 * s + 1) * b_bytes);
            const uint32_t acc = tmem_base + (uint32_t)(s * a.acc_stride);
            for (int ks = 0; ks < ksteps; ks++) {
               const uint32_t oa = ks * 2 * a_lbo, ob = ks * 2 * b_lbo;
               mma_bf16(acc, make_desc(ah + oa, a_lbo), make_desc(bh + ob, b_lbo), a.idesc, ks > 0 ? 1u : 0u);   // hi * hi
               mma_bf16(acc, make_desc(ah + oa, a_lbo), make_desc(bl + ob, b_lbo), a.idesc, 1u);                 // hi * lo
               mma_bf16(acc, make_desc(al + oa, a_lbo), make_desc(bh + ob, b_lbo), a.idesc, 1u);                 // lo * hi
            }
            umma_commit(bar_empty + 8 * s);
            umma_commit(bar_tfull + 8 * s);
         }
      }
   } else {
      // ===== epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = 32 entities; one output dimension per register =====
      const int c = tile * BM + warp * 32 + lane;
      const bool in = c < a.ld;
      for (int i = 0; i < n_iter; i++) {
         const int s = i & 1;
         mbar_wait(bar_tfull + 8 * s, (i >> 1) & 1);
         asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
         float* out = a.out + (size_t)(s_begin + i) * a.D * a.ld + c;
         for (int cb = 0; cb < a.D; cb += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * a.acc_stride + cb);
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
            if (in) {
#pragma unroll
               for (int j = 0; j < 32; j++)
                  if (cb + j < a.D) __stcs(out + (size_t)(cb + j) * a.ld, __uint_as_float(v[j]));   // streaming: written once, read by the next kernel
            }
         }
         asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
         __syncwarp();
         if (lane == 0) mbar_arrive(bar_tempty + 8 * s);
      }
   }
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   __syncthreads();
   if (warp == EPI_WARPS + 1) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(a.tmem_cols) : "memory");
   }
}

// ---- operand preparation -------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(double x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
   const float xf = (float)x;
   hi = __float2bfloat16_rn(xf);
   lo = __float2bfloat16_rn(xf - __bfloat162float(hi));
}

// entities: fp64 [nE][D] -> tiles [tile][chunk][row][8 bf16] (hi and lo); max |e|_2 as ordered bits of a non-negative double
__global__ void prep_entities_kernel(const double* __restrict__ ent64, int nE, int n_pad, int D, int kc, __nv_bfloat16* e_hi,
                                     __nv_bfloat16* e_lo, unsigned long long* emax_bits) {
   const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   const int lane = threadIdx.x & 31;
   if (row >= n_pad) return;
   const int tile = row / BM, r = row - tile * BM;
   const size_t base = (size_t)tile * (BM * kc * 8);
   double s = 0.0;
   for (int i = lane; i < kc * 8; i += 32) {
      const double x = (row < nE && i < D) ? ent64[(size_t)row * D + i] : 0.0;
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      const size_t idx = base + (size_t)(i >> 3) * (BM * 8) + (size_t)r * 8 + (i & 7);
      e_hi[idx] = h;
      e_lo[idx] = l;
      s += x * x;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   if (lane == 0 && row < nE) atomicMax(emax_bits, (unsigned long long)__double_as_longlong(sqrt(s) * (1.0 + 1e-12)));
}

// relations of the pass: M_r (fp64, M[j][i]) -> B tile = M_r^T rows (output dimension i) x K (input dimension j), and the
// error / magnitude bounds of the slot's projected rows
__global__ void prep_relations_kernel(const double* __restrict__ w64, const int32_t* __restrict__ slot_rel, int D, int kc, int ncols,
                                      __nv_bfloat16* m_hi, __nv_bfloat16* m_lo, const unsigned long long* emax_bits, double eps,
                                      double* __restrict__ bounds) {
   __shared__ double s_r1[MAXD], s_r2[MAXD];
   const int slot = blockIdx.x;
   const double* M = w64 + (size_t)slot_rel[slot] * D * D;
   const size_t base = (size_t)slot * ((size_t)ncols * kc * 8);
   for (int idx = threadIdx.x; idx < ncols * kc * 8; idx += blockDim.x) {
      const int j = idx / ncols, i = idx - j * ncols;   // consecutive threads: consecutive output dims of one input row j
      const double x = (i < D && j < D) ? M[(size_t)j * D + i] : 0.0;
      __nv_bfloat16 h, l;
      split_bf16(x, h, l);
      const size_t o = base + (size_t)(j >> 3) * (ncols * 8) + (size_t)i * 8 + (j & 7);
      m_hi[o] = h;
      m_lo[o] = l;
   }
   for (int j = threadIdx.x; j < D; j += blockDim.x) {
      double r1 = 0.0, r2 = 0.0;
      for (int i = 0; i < D; i++) {
         const double x = M[(size_t)j * D + i];
         r1 += fabs(x);
         r2 += x * x;
      }
      s_r1[j] = r1 * r1;
      s_r2[j] = r2;
   }
   __syncthreads();
   if (threadIdx.x == 0) {
      double b1 = 0.0, b2 = 0.0;
      for (int j = 0; j < D; j++) { b1 += s_r1[j]; b2 += s_r2[j]; }
      const double emax = __longlong_as_double((long long)*emax_bits);
      const double B1 = sqrt(b1) * (1.0 + 1e-9), B2 = sqrt(b2) * (1.0 + 1e-9);
      const double eta1 = eps * emax * B1, eta2 = eps * emax * B2;
      const double c2 = B2 * emax + eta2;          // |P~ row|_2 <= |M|_F |e|_2 + |eta|_2
      bounds[4 * slot + 0] = sqrt((double)D) * c2;  // |.|_1 <= sqrt(D) |.|_2
      bounds[4 * slot + 1] = c2 * c2;
      bounds[4 * slot + 2] = eta1;
      bounds[4 * slot + 3] = eta2;
   }
}

// ---- exact fp64 pieces (the reference's operation order) -----------------------------------------------------------
// p_i = sum_j M[j][i] * e_j, j ascending from a zero accumulator (transr/transr.cpp:20-25); lane owns i = lane + 32 k.
__device__ __forceinline__ void project_exact(const double* __restrict__ M, const double* __restrict__ e, int D, int lane, double (&acc)[4]) {
#pragma unroll
   for (int k = 0; k < 4; k++) acc[k] = 0.0;
   for (int j = 0; j < D; j++) {
      const double x = __ldg(e + j);
      const double* m = M + (size_t)j * D + lane;
#pragma unroll
      for (int k = 0; k < 4; k++)
         if (lane + 32 * k < D) acc[k] = __dadd_rn(acc[k], __dmul_rn(__ldg(m + 32 * k), x));
   }
}

// Energy of candidate row p (this warp's registers) against the query's projected fixed entity V:
// sum_i f((V_i - p_i) - d'_i), i ascending (exact_energy of rank.cu).  The result is valid on lane 0.
template <int L2>
__device__ __forceinline__ double energy_exact(const double (&p)[4], const double* __restrict__ V, const double* __restrict__ d, double dsign,
                                               int D, int lane, double* s_terms) {
#pragma unroll
   for (int k = 0; k < 4; k++) {
      const int i = lane + 32 * k;
      if (i < D) {
         const double v = __dsub_rn(__dsub_rn(V[i], p[k]), dsign * __ldg(d + i));
         s_terms[i] = L2 ? __dmul_rn(v, v) : fabs(v);
      }
   }
   __syncwarp();
   double e = 0.0;
   if (lane == 0)
      for (int i = 0; i < D; i++) e = __dadd_rn(e, s_terms[i]);
   __syncwarp();
   return e;
}

struct QueryRefs {
   const double* ent64;
   const double* rel64;
   const double* w64;
   const int32_t* q_fixed;
   const int32_t* q_truth;
   const int32_t* q_rel;
   const int32_t* q_side;
   int D;
};

// V[q] = M_r^T e_fixed and E_true[q] (one warp per query)
template <int L2>
__global__ void __launch_bounds__(32 * EX_WARPS) query_kernel(const QueryRefs r, long long q_begin, long long q_end, double* __restrict__ V,
                                                              double* __restrict__ q_etrue) {
   __shared__ double s_terms[EX_WARPS][MAXD];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const long long q = q_begin + (long long)blockIdx.x * EX_WARPS + warp;
   if (q >= q_end) return;
   const int D = r.D, rel = r.q_rel[q];
   const double* M = r.w64 + (size_t)rel * D * D;
   double* Vq = V + (size_t)q * D;
   double p[4];
   project_exact(M, r.ent64 + (size_t)r.q_fixed[q] * D, D, lane, p);
#pragma unroll
   for (int k = 0; k < 4; k++)
      if (lane + 32 * k < D) Vq[lane + 32 * k] = p[k];
   __syncwarp();
   project_exact(M, r.ent64 + (size_t)r.q_truth[q] * D, D, lane, p);
   const double e = energy_exact<L2>(p, Vq, r.rel64 + (size_t)rel * D, r.q_side[q] ? -1.0 : 1.0, D, lane, s_terms[warp]);
   if (lane == 0) q_etrue[q] = e;
}

// w = fl32(V - d') and the thresholds E_true -+ delta of the fp32 pre-filter: the rounding-error bound of rank_f32.cu with
// the projection-error terms eta_1 / eta_2 of the slot added to the per-row error budget
template <int L2>
__global__ void thresholds_kernel(const QueryRefs r, const int32_t* __restrict__ q_slot, const double* __restrict__ V,
                                  const double* __restrict__ q_etrue, const double* __restrict__ bounds, long long q_begin, long long q_end,
                                  float* __restrict__ wq, float* __restrict__ thr_lo, float* __restrict__ thr_hi) {
   const long long q = q_begin + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   const int lane = threadIdx.x & 31;
   if (q >= q_end) return;
   const int D = r.D;
   const double* d = r.rel64 + (size_t)r.q_rel[q] * D;
   const double dsign = r.q_side[q] ? -1.0 : 1.0;
   double a1 = 0.0, a2 = 0.0;
   for (int i = lane; i < D; i += 32) {
      const double w = V[(size_t)q * D + i] - dsign * d[i];
      wq[(size_t)q * D + i] = (float)w;
      a1 += fabs(w);
      a2 += w * w;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
   }
   if (lane == 0) {
      const double u = 5.9604644775390625e-08;   // 2^-24
      const double E = q_etrue[q];
      const double* b = bounds + 4 * (size_t)q_slot[q];
      double first;
      if (L2) {
         const double eN = 2.01 * u * sqrt(2.0 * (a2 + b[1]) * 1.0000001) + b[3];
         first = 2.1 * eN * sqrt(E) + 2.0 * eN * eN;
      } else {
         first = 2.01 * u * (a1 + b[0]) * 1.0000001 + b[2];
      }
      double delta = first + 1.01 * (D + 1) * u * (E + first);   // fp32 accumulation error of the (perturbed) sum
      delta += 1e-12 * (a1 + b[0] + E) + 2.2250738585072014e-308;
      thr_lo[q] = __double2float_rd(E - delta);
      thr_hi[q] = __double2float_ru(E + delta);
   }
}

// undecided band of the pre-filter: exact energies, candidates projected on demand (one warp per entry)
template <int L2>
__global__ void __launch_bounds__(32 * EX_WARPS) recheck_kernel(const QueryRefs r, const int2* __restrict__ band, unsigned int* band_count,
                                                                unsigned int band_cap, const double* __restrict__ V,
                                                                const double* __restrict__ q_etrue, int32_t* q_cnt, long long nq) {
   __shared__ double s_terms[EX_WARPS][MAXD];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const unsigned int n = min(*band_count, band_cap);
   if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(band_count + 2, n);
   const unsigned int warps = gridDim.x * EX_WARPS;
   const int D = r.D;
   // contiguous shares: consecutive entries mostly belong to the same query tile = the same relation (M_r stays in L1)
   const unsigned int per = (n + warps - 1) / warps;
   const unsigned int w = blockIdx.x * EX_WARPS + warp;
   const unsigned int k_end = min(n, (w + 1) * per);
   for (unsigned int k = w * per; k < k_end; k++) {
      const int q = band[k].x, c = band[k].y;
      if (c == r.q_truth[q]) continue;
      const int rel = r.q_rel[q];
      double p[4];
      project_exact(r.w64 + (size_t)rel * D * D, r.ent64 + (size_t)c * D, D, lane, p);
      const double e = energy_exact<L2>(p, V + (size_t)q * D, r.rel64 + (size_t)rel * D, r.q_side[q] ? -1.0 : 1.0, D, lane, s_terms[warp]);
      if (lane == 0) {
         const double et = q_etrue[q];
         if (e < et) atomicAdd(q_cnt + q, 1);
         else if (e == et) atomicAdd(q_cnt + nq + q, 1);
      }
   }
}

// filter pass: (query, known-true neighbour) pairs (rank.cu: filter_plan_kernel), neighbours projected on demand, one
// warp per pair over contiguous shares of the pair list (neighbouring pairs share the query, hence M_r stays in L1)
template <int L2>
__global__ void __launch_bounds__(32 * EX_WARPS) filter_kernel(const QueryRefs r, const int2* __restrict__ pairs,
                                                               const unsigned int* __restrict__ pair_count, unsigned int pair_cap,
                                                               const double* __restrict__ V, const double* __restrict__ q_etrue, int32_t* q_cnt,
                                                               long long nq) {
   __shared__ double s_terms[EX_WARPS][MAXD];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const unsigned int n = min(*pair_count, pair_cap);
   const unsigned int warps = gridDim.x * EX_WARPS;
   const unsigned int per = (n + warps - 1) / warps;
   const unsigned int w = blockIdx.x * EX_WARPS + warp;
   const unsigned int k_end = min(n, (w + 1) * per);
   const int D = r.D;
   for (unsigned int k = w * per; k < k_end; k++) {
      const int2 pr = __ldg(pairs + k);
      if (pr.y < 0) continue;
      const long long q = pr.x;
      const int rel = r.q_rel[q];
      double p[4];
      project_exact(r.w64 + (size_t)rel * D * D, r.ent64 + (size_t)pr.y * D, D, lane, p);
      const double e = energy_exact<L2>(p, V + (size_t)q * D, r.rel64 + (size_t)rel * D, r.q_side[q] ? -1.0 : 1.0, D, lane, s_terms[warp]);
      if (lane == 0) {
         const double et = q_etrue[q];
         if (e < et) atomicAdd(q_cnt + 2 * nq + q, 1);
         else if (e == et) atomicAdd(q_cnt + 3 * nq + q, 1);
      }
   }
}

}  // namespace trp

// ---- host -----------------------------------------------------------------------------------------------------------
static inline unsigned nblk4(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// |P~[c][i] - P[c][i]| <= trp_eps(D) * sum_j |e_cj| |M_ji|:
//   operand split:  x = hi + lo + dx, |dx| <= (2^-18 + 2^-24) |x|, |lo| <= 2^-9 |x|; the dropped lo*lo and the two
//                   dx terms: (2^-18 + 2 * 2^-17.98) < 2^-16.4 of |x y| per product;
//   accumulation:   3 K/16 MMA instructions per output, each adding 16 exact products to the fp32 accumulator with
//                   alignment to the largest exponent and truncation: <= 17 * 2^-23 of the largest addend (<= the sum of
//                   absolute products) per instruction;
//   x 1.25 safety.  tests/test_gpu_rank.py::test_transr_tensor_core_projection_error measures the real error against it.
double trp_eps(int D) {
   const int k_pad = (D + 15) / 16 * 16;
   return 1.25 * (std::ldexp(1.0, -16) * 0.76 + 3.0 * (k_pad / 16) * 17.0 * std::ldexp(1.0, -23));
}

bool trp_supported(const kb2e_ctx* c) {
   return c->cfg.model == KB2E_MODEL_TRANSR && c->D <= trp::MAXD && !(c->cfg.flags & KB2E_FLAG_RANK_EXACT_ONLY);
}

template <typename T>
static int grow_bytes(kb2e_ctx* c, T** p, size_t* cap, size_t bytes) {
   if (bytes <= *cap && *p) return KB2E_OK;
   pool_free(c, *p);
   *p = nullptr;
   *cap = 0;
   KB2E_CUDA(c, pool_alloc(c, p, std::max<size_t>(16, bytes)));
   *cap = std::max<size_t>(16, bytes);
   return KB2E_OK;
}

static trp::QueryRefs make_refs(kb2e_ctx* c, const int32_t* q_int, long long nq) {
   trp::QueryRefs r;
   r.ent64 = c->ent64; r.rel64 = c->rel64; r.w64 = c->w64;
   r.q_fixed = q_int; r.q_truth = q_int + nq; r.q_rel = q_int + 2 * nq; r.q_side = q_int + 3 * nq;
   r.D = c->D;
   return r;
}

int trp_prepare_entities(kb2e_ctx* c, TrpState* s) {
   s->kc = 2 * ((c->D + 15) / 16);
   s->ncols = (c->D + 15) / 16 * 16;
   s->n_pad = (c->nE + trp::BM - 1) / trp::BM * trp::BM;
   const size_t bytes = (size_t)s->n_pad * s->kc * 16;
   if (bytes > s->e_cap || !s->e_hi) {
      pool_free(c, s->e_hi); pool_free(c, s->e_lo);
      s->e_hi = s->e_lo = nullptr;
      s->e_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->e_hi, bytes));
      KB2E_CUDA(c, pool_alloc(c, &s->e_lo, bytes));
      s->e_cap = bytes;
   }
   if (!s->scalars) KB2E_CUDA(c, pool_alloc(c, &s->scalars, 2 * sizeof(double)));
   if (!s->e0) {
      KB2E_CUDA(c, cudaEventCreate(&s->e0));
      KB2E_CUDA(c, cudaEventCreate(&s->e1));
   }
   KB2E_CUDA(c, cudaMemsetAsync(s->scalars, 0, 2 * sizeof(double), c->stream));
   trp::prep_entities_kernel<<<nblk4((long long)s->n_pad * 32, 256), 256, 0, c->stream>>>(
      c->ent64, c->nE, s->n_pad, c->D, s->kc, (__nv_bfloat16*)s->e_hi, (__nv_bfloat16*)s->e_lo, reinterpret_cast<unsigned long long*>(s->scalars));
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int trp_project(kb2e_ctx* c, TrpState* s, const std::vector<int32_t>& rels, int ld, float* ct32) {
   const size_t slots = rels.size();
   if (slots == 0) return KB2E_OK;
   const size_t b_bytes = (size_t)s->ncols * s->kc * 16;
   if (slots > s->slot_cap) {
      pool_free(c, s->m_hi); pool_free(c, s->m_lo); pool_free(c, s->bounds); pool_free(c, s->slot_rel);
      s->m_hi = s->m_lo = nullptr; s->bounds = nullptr; s->slot_rel = nullptr;
      s->slot_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->m_hi, slots * b_bytes));
      KB2E_CUDA(c, pool_alloc(c, &s->m_lo, slots * b_bytes));
      KB2E_CUDA(c, pool_alloc(c, &s->bounds, slots * 4 * sizeof(double)));
      KB2E_CUDA(c, pool_alloc(c, &s->slot_rel, slots * sizeof(int32_t)));
      s->slot_cap = slots;
   }
   // pageable source: staged by the runtime before the call returns
   KB2E_CUDA(c, cudaMemcpyAsync(s->slot_rel, rels.data(), slots * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   trp::prep_relations_kernel<<<(unsigned)slots, 128, 0, c->stream>>>(c->w64, s->slot_rel, c->D, s->kc, s->ncols, (__nv_bfloat16*)s->m_hi,
                                                                      (__nv_bfloat16*)s->m_lo, reinterpret_cast<unsigned long long*>(s->scalars),
                                                                      trp_eps(c->D), s->bounds);
   KB2E_CUDA(c, cudaGetLastError());
   trp::ProjArgs a;
   a.e_hi = (const unsigned char*)s->e_hi; a.e_lo = (const unsigned char*)s->e_lo;
   a.m_hi = (const unsigned char*)s->m_hi; a.m_lo = (const unsigned char*)s->m_lo;
   a.out = ct32;
   a.slots = (int)slots; a.D = c->D; a.ld = ld; a.kc = s->kc; a.ncols = s->ncols;
   a.acc_stride = s->ncols <= 32 ? 32 : (s->ncols <= 64 ? 64 : 128);
   a.tmem_cols = (uint32_t)(trp::STAGES * a.acc_stride);
   // cute::UMMA::InstrDescriptor: D = F32, A = B = BF16, both K-major, N = ncols, M = 128
   a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(s->ncols >> 3) << 17) | ((uint32_t)(trp::BM >> 4) << 24);
   const size_t a_bytes = (size_t)trp::BM * s->kc * 16;
   const size_t smem = 2 * a_bytes + 2 * trp::STAGES * b_bytes + 128;
   KB2E_CUDA(c, cudaFuncSetAttribute(trp::project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   int per_sm = 0;
   KB2E_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trp::project_tc_kernel, trp::THREADS, smem));
   per_sm = std::max(1, std::min(per_sm, (int)(512u / a.tmem_cols)));
   // cut the relations of the pass into ranges so that the work items fill whole waves of resident CTAs
   const long long tiles = s->n_pad / trp::BM, cap = (long long)c->num_sms * per_sm;
   long long splits = 1, best = 1;
   double best_eff = 0.0;
   for (splits = 1; splits <= (long long)slots; splits++) {
      const long long items = tiles * splits, waves = (items + cap - 1) / cap;
      const double eff = (double)items / (double)(waves * cap);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = splits; }
      if (items >= 2 * cap && eff > 0.93) { best = splits; break; }
      if ((long long)slots / splits < 4 && best_eff > 0.5) break;   // ranges of a few relations: the per-CTA set-up would dominate
   }
   splits = std::min<long long>(best, (long long)slots);
   KB2E_CUDA(c, cudaEventRecord(s->e0, c->stream));
   trp::project_tc_kernel<<<dim3((unsigned)tiles, (unsigned)splits), trp::THREADS, smem, c->stream>>>(a);
   KB2E_CUDA(c, cudaEventRecord(s->e1, c->stream));
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int trp_queries(kb2e_ctx* c, TrpState* s, bool l2, const int32_t* q_int, long long nq_total, long long q_begin, long long q_end, double* q_etrue) {
   if (nq_total > s->q_cap) {
      pool_free(c, s->V);
      s->V = nullptr;
      s->q_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->V, (size_t)nq_total * c->D * sizeof(double)));
      s->q_cap = nq_total;
   }
   const trp::QueryRefs r = make_refs(c, q_int, nq_total);
   const unsigned blocks = nblk4(q_end - q_begin, trp::EX_WARPS);
   if (l2) trp::query_kernel<1><<<blocks, 32 * trp::EX_WARPS, 0, c->stream>>>(r, q_begin, q_end, s->V, q_etrue);
   else trp::query_kernel<0><<<blocks, 32 * trp::EX_WARPS, 0, c->stream>>>(r, q_begin, q_end, s->V, q_etrue);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int trp_thresholds(kb2e_ctx* c, TrpState* s, F32State* f, bool l2, const int32_t* q_int, long long nq_total, long long q_begin, long long q_end,
                   const double* q_etrue) {
   const trp::QueryRefs r = make_refs(c, q_int, nq_total);
   const int32_t* q_slot = q_int + 4 * nq_total;
   const unsigned blocks = nblk4((q_end - q_begin) * 32, 256);
   if (l2) trp::thresholds_kernel<1><<<blocks, 256, 0, c->stream>>>(r, q_slot, s->V, q_etrue, s->bounds, q_begin, q_end, f->wq, f->thr_lo, f->thr_hi);
   else trp::thresholds_kernel<0><<<blocks, 256, 0, c->stream>>>(r, q_slot, s->V, q_etrue, s->bounds, q_begin, q_end, f->wq, f->thr_lo, f->thr_hi);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int trp_recheck(kb2e_ctx* c, TrpState* s, F32State* f, bool l2, const int32_t* q_int, long long nq_total, const double* q_etrue, int32_t* q_cnt) {
   const trp::QueryRefs r = make_refs(c, q_int, nq_total);
   const unsigned blocks = 8 * c->num_sms;
   if (l2) trp::recheck_kernel<1><<<blocks, 32 * trp::EX_WARPS, 0, c->stream>>>(r, f->band, f->band_count, f->band_cap, s->V, q_etrue, q_cnt, nq_total);
   else trp::recheck_kernel<0><<<blocks, 32 * trp::EX_WARPS, 0, c->stream>>>(r, f->band, f->band_count, f->band_cap, s->V, q_etrue, q_cnt, nq_total);
   KB2E_CUDA(c, cudaGetLastError());
   // the largest band of the call decides whether the list overflowed (checked by the caller after its one synchronisation)
   KB2E_CUDA(c, cudaMemcpyAsync(f->host_count, f->band_count, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
   return KB2E_OK;
}

int trp_filter(kb2e_ctx* c, TrpState* s, bool l2, const int32_t* q_int, long long nq_total, const double* q_etrue, const int2* pairs,
               const unsigned int* pair_count, unsigned int pair_cap, int32_t* q_cnt, cudaStream_t stream) {
   const trp::QueryRefs r = make_refs(c, q_int, nq_total);
   const unsigned blocks = 8 * c->num_sms;
   if (l2) trp::filter_kernel<1><<<blocks, 32 * trp::EX_WARPS, 0, stream>>>(r, pairs, pair_count, pair_cap, s->V, q_etrue, q_cnt, nq_total);
   else trp::filter_kernel<0><<<blocks, 32 * trp::EX_WARPS, 0, stream>>>(r, pairs, pair_count, pair_cap, s->V, q_etrue, q_cnt, nq_total);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int trp_debug_project(kb2e_ctx* c, TrpState* s, int relation, float* out, double* eps_rel) {
   const int ld = (c->nE + 31) / 32 * 32;
   int rc = trp_prepare_entities(c, s);
   if (rc) return rc;
   float* dev = nullptr;
   KB2E_CUDA(c, pool_alloc(c, &dev, (size_t)c->D * ld * sizeof(float)));
   rc = trp_project(c, s, std::vector<int32_t>(1, relation), ld, dev);
   std::vector<float> host((size_t)c->D * ld);
   if (rc == KB2E_OK) {
      cudaError_t e = cudaMemcpyAsync(host.data(), dev, host.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = cuda_fail(c, e, "kb2e_debug_transr_projection copy");
   }
   pool_free(c, dev);
   if (rc) return rc;
   for (int i = 0; i < c->D; i++)
      for (int e = 0; e < c->nE; e++) out[(size_t)e * c->D + i] = host[(size_t)i * ld + e];
   if (eps_rel) *eps_rel = trp_eps(c->D);
   return KB2E_OK;
}

void trp_free(kb2e_ctx* c, TrpState* s) {
   pool_free(c, s->e_hi); pool_free(c, s->e_lo); pool_free(c, s->m_hi); pool_free(c, s->m_lo); pool_free(c, s->bounds);
   pool_free(c, s->V); pool_free(c, s->scalars); pool_free(c, s->slot_rel);
   if (s->e0) { cudaEventDestroy(s->e0); cudaEventDestroy(s->e1); }
   *s = TrpState();
}

}  // namespace kb2e
