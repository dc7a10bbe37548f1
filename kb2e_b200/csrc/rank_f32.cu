// fp32 pre-filter for the all-candidates ranking pass of every model / distance that has no tensor-core path
// (TransE L1, TransH, TransR): common/evaluation.cpp:124-179 scores all N_E candidates per query in fp64, which is
// FP64-issue bound on the GPU (rank_exact_kernel).  Here the candidates are scored in fp32 on the CUDA cores -- L1 is
// not a contraction, so no tensor cores -- with a RIGOROUS error bound, and only the candidates the bound cannot decide
// are re-scored exactly:
//
//   E(c) = sum_i f((V_i - C_i) - d'_i) = sum_i f(W_i - C_i),   W = V - d' (fp64, then rounded to fp32), f = |.| or (.)^2
//   S32(c) = the same sum in fp32 (round to nearest, sequential accumulation, fma allowed)
//
// With u = 2^-24:  |fl32(W_i) - W_i| <= u|W_i|,  |fl32(C_i) - C_i| <= u|C_i|,  the subtraction adds u|w_i - c_i|, so the
// residual error is e_i <= 2u(1+u)(|W_i| + |C_i|);  n sequential additions add at most 1.01 n u times the sum.  Hence
//   L1:  |S32 - E| <= 2.01 u (|W|_1 + |C|_1)                               + 1.01 (D+1) u E
//   L2:  |S32 - E| <= 2.1 eN sqrt(E) + 2 eN^2,  eN = 2.01 u sqrt(2(|W|_2^2 + |C|_2^2)),  + 1.01 (D+1) u E
// evaluated at E = E_true with |C| replaced by its maximum over the slot's candidates (the bound grows with E slower than
// E itself, so thresholds taken at E_true are valid on both sides).  A candidate with S32 < E_true - delta is counted as
// ranked before the truth, one with S32 > E_true + delta is ignored, and the band in between goes to a list that
// recheck_ct_kernel re-scores with exact_energy (the reference's fp64 operation order), so the final counts -- and the
// ranks -- are bit-identical to rank_exact_kernel's.

#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "internal.h"
#include "rank_f32.h"

namespace kb2e {
namespace f32 {

constexpr int QT = 32;         // queries per tile
constexpr int THREADS = 256;
constexpr int CPT = 2;         // candidates per thread and step

// ---- operand preparation ---------------------------------------------------------------------------------
// candidates: fp64 transposed [slot][D][ld] -> fp32, plus the slot's max |C|_1 and max |C|_2^2 (fp64, as ordered bits)
__global__ void prep_candidates_kernel(const double* __restrict__ ct, float* __restrict__ ct32, int nE, int D, int ld,
                                       unsigned long long* __restrict__ slot_max /* [slots][2] */) {
   const int c = blockIdx.x * blockDim.x + threadIdx.x;
   const int slot = blockIdx.y;
   const double* src = ct + (size_t)slot * D * ld;
   float* dst = ct32 + (size_t)slot * D * ld;
   double n1 = 0.0, n2 = 0.0;
   if (c < ld) {
      for (int i = 0; i < D; i++) {
         const double v = c < nE ? src[(size_t)i * ld + c] : 0.0;
         dst[(size_t)i * ld + c] = (float)v;
         n1 += fabs(v);
         n2 += v * v;
      }
   }
   // non-negative doubles order like their bit patterns
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      n1 = fmax(n1, __shfl_xor_sync(0xffffffffu, n1, o));
      n2 = fmax(n2, __shfl_xor_sync(0xffffffffu, n2, o));
   }
   if ((threadIdx.x & 31) == 0) {
      atomicMax(slot_max + 2 * slot, (unsigned long long)__double_as_longlong(n1));
      atomicMax(slot_max + 2 * slot + 1, (unsigned long long)__double_as_longlong(n2));
   }
}

// queries: w = fl32(V - d') as [q][D], thresholds E_true -+ delta (thr_lo rounded down, thr_hi rounded up)
template <int L2>
__global__ void prep_queries_kernel(const double* __restrict__ ct, const double* __restrict__ rel64, const int32_t* __restrict__ q_fixed,
                                    const int32_t* __restrict__ q_rel, const int32_t* __restrict__ q_side, const int32_t* __restrict__ q_slot,
                                    const double* __restrict__ q_etrue, const unsigned long long* __restrict__ slot_max, long long q_begin,
                                    long long q_end, int D, int ld, float* __restrict__ wq, float* __restrict__ thr_lo, float* __restrict__ thr_hi) {
   const long long q = q_begin + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
   const int lane = threadIdx.x & 31;
   if (q >= q_end) return;
   const int slot = q_slot[q];
   const double* c = ct + (size_t)slot * D * ld;
   const double* d = rel64 + (size_t)q_rel[q] * D;
   const int fixed = q_fixed[q];
   const double dsign = q_side[q] ? -1.0 : 1.0;
   double a1 = 0.0, a2 = 0.0;
   for (int i = lane; i < D; i += 32) {
      const double w = c[(size_t)i * ld + fixed] - dsign * d[i];
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
      const double b1 = __longlong_as_double((long long)slot_max[2 * slot]);
      const double b2 = __longlong_as_double((long long)slot_max[2 * slot + 1]);
      double delta;
      if (L2) {
         const double eN = 2.01 * u * sqrt(2.0 * (a2 + b2) * 1.0000001);   // a2 itself carries a few fp64 ulps
         delta = 2.1 * eN * sqrt(E) + 2.0 * eN * eN + 1.01 * (D + 1) * u * E;
      } else {
         delta = 2.01 * u * (a1 + b1) * 1.0000001 + 1.01 * (D + 1) * u * E;
      }
      delta += 1e-12 * (a1 + b1 + E) + DBL_MIN;   // fp64 rounding of the reference sums themselves
      thr_lo[q] = __double2float_rd(E - delta);
      thr_hi[q] = __double2float_ru(E + delta);
   }
}

// ---- the all-candidates kernel -----------------------------------------------------------------------------
// grid.x = query tiles (<= QT queries of one slot), grid.y = candidate splits.  The tile's w vectors sit in shared memory
// as [D][QT] so that one 16-byte broadcast load feeds four running sums per candidate; each thread owns CPT candidates per step.
template <int L2>
__global__ void __launch_bounds__(THREADS) rank_f32_kernel(const F32Args a) {
   extern __shared__ float s_w[];   // [D][QT]
   __shared__ float s_lo[QT], s_hi[QT];
   __shared__ int s_cnt[QT];
   const int4 tile = a.tiles[blockIdx.x];
   const int q0 = tile.x, nq = tile.y;
   const float* ct = a.ct32 + (size_t)tile.z * a.D * a.ld;
   const int D = a.D, ld = a.ld;
   for (int k = threadIdx.x; k < D * QT; k += THREADS) {
      const int i = k / QT, q = k % QT;
      s_w[k] = q < nq ? a.wq[(size_t)(q0 + q) * D + i] : 0.f;
   }
   if (threadIdx.x < QT) {
      const int q = threadIdx.x;
      // padding queries of the tile can never count or reach the band
      s_lo[q] = q < nq ? a.thr_lo[q0 + q] : -FLT_MAX;
      s_hi[q] = q < nq ? a.thr_hi[q0 + q] : -FLT_MAX;
      s_cnt[q] = 0;
   }
   __syncthreads();
   // CPT candidates per thread and step: every 16-byte broadcast load of four query values then feeds 4 * CPT running
   // sums.  With one candidate per thread the kernel was bound by the shared-memory pipe (8 LDS.128 per 64 FP32
   // operations), not by FP32 issue.
   const int steps = (a.nE + THREADS * CPT - 1) / (THREADS * CPT);
   const int s_begin = (int)((long long)steps * blockIdx.y / a.splits);
   const int s_end = (int)((long long)steps * (blockIdx.y + 1) / a.splits);
   int cnt[QT];
#pragma unroll
   for (int q = 0; q < QT; q++) cnt[q] = 0;
   for (int step = s_begin; step < s_end; step++) {
      int c[CPT];
      bool valid[CPT];
      const float* col[CPT];
#pragma unroll
      for (int k = 0; k < CPT; k++) {
         c[k] = (step * CPT + k) * THREADS + threadIdx.x;
         valid[k] = c[k] < a.nE;
         col[k] = ct + (valid[k] ? c[k] : 0);
      }
      float acc[CPT][QT];
#pragma unroll
      for (int k = 0; k < CPT; k++)
#pragma unroll
         for (int q = 0; q < QT; q++) acc[k][q] = 0.f;
#pragma unroll 2
      for (int i = 0; i < D; i++) {
         float ci[CPT];
#pragma unroll
         for (int k = 0; k < CPT; k++) ci[k] = __ldg(col[k] + (size_t)i * ld);
         const float4* w4 = reinterpret_cast<const float4*>(s_w + i * QT);
#pragma unroll
         for (int g = 0; g < QT / 4; g++) {
            const float4 w = w4[g];
#pragma unroll
            for (int k = 0; k < CPT; k++) {
               if (L2) {
                  const float r0 = w.x - ci[k], r1 = w.y - ci[k], r2 = w.z - ci[k], r3 = w.w - ci[k];
                  acc[k][4 * g + 0] = fmaf(r0, r0, acc[k][4 * g + 0]);
                  acc[k][4 * g + 1] = fmaf(r1, r1, acc[k][4 * g + 1]);
                  acc[k][4 * g + 2] = fmaf(r2, r2, acc[k][4 * g + 2]);
                  acc[k][4 * g + 3] = fmaf(r3, r3, acc[k][4 * g + 3]);
               } else {
                  acc[k][4 * g + 0] += fabsf(w.x - ci[k]);
                  acc[k][4 * g + 1] += fabsf(w.y - ci[k]);
                  acc[k][4 * g + 2] += fabsf(w.z - ci[k]);
                  acc[k][4 * g + 3] += fabsf(w.w - ci[k]);
               }
            }
         }
      }
#pragma unroll
      for (int k = 0; k < CPT; k++) {
         if (!valid[k]) continue;
#pragma unroll
         for (int q = 0; q < QT; q++) {
            const float s = acc[k][q];
            if (s < s_lo[q]) {
               cnt[q]++;
            } else if (s <= s_hi[q]) {
               const unsigned int slot = atomicAdd(a.band_count, 1u);
               if (slot < a.band_cap) a.band[slot] = make_int2(q0 + q, c[k]);
               else a.band_count[1] = 1u;   // overflow: the caller redoes the call with the exact kernel
            }
         }
      }
   }
#pragma unroll
   for (int q = 0; q < QT; q++) {
      int v = cnt[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[q], v);
   }
   __syncthreads();
   if (threadIdx.x < nq && s_cnt[threadIdx.x]) atomicAdd(a.q_less + q0 + threadIdx.x, s_cnt[threadIdx.x]);
}

// exact re-score of the band: same arithmetic and order as exact_energy (rank.cu), candidates read from the transposed
// fp64 matrix of the query's slot
template <int L2>
__global__ void recheck_ct_kernel(const int2* __restrict__ band, const unsigned int* __restrict__ band_count, unsigned int band_cap,
                                  const double* __restrict__ ct, const double* __restrict__ rel64, const int32_t* q_fixed,
                                  const int32_t* q_truth, const int32_t* q_rel, const int32_t* q_side, const int32_t* q_slot,
                                  const double* q_etrue, int32_t* q_cnt, long long nq, int D, int ld) {
   const unsigned int n = min(*band_count, band_cap);
   if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(const_cast<unsigned int*>(band_count) + 2, n);
   for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
      const int q = band[k].x, c = band[k].y;
      if (c == q_truth[q]) continue;
      const double* m = ct + (size_t)q_slot[q] * D * ld;
      const double* d = rel64 + (size_t)q_rel[q] * D;
      const double dsign = q_side[q] ? -1.0 : 1.0;
      const int fixed = q_fixed[q];
      double acc = 0.0;
      for (int i = 0; i < D; i++) {
         const double t = __dsub_rn(__dsub_rn(m[(size_t)i * ld + fixed], m[(size_t)i * ld + c]), dsign * d[i]);
         acc = L2 ? __dadd_rn(acc, __dmul_rn(t, t)) : __dadd_rn(acc, fabs(t));
      }
      const double et = q_etrue[q];
      if (acc < et) atomicAdd(q_cnt + q, 1);
      else if (acc == et) atomicAdd(q_cnt + nq + q, 1);
   }
}

}  // namespace f32

// ---- host ------------------------------------------------------------------------------------------------
static inline unsigned nblk3(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// Buffers only (grow-only): the fp32 candidate matrices of `slots` slots, the per-slot maxima and the per-query arrays.
int f32_ensure(kb2e_ctx* c, F32State* s, size_t slots, int ld, long long nq, bool first_pass) {
   const size_t need_ct = slots * (size_t)c->D * ld;
   if (need_ct > s->ct32_cap) {
      pool_free(c, s->ct32);
      s->ct32 = nullptr; s->ct32_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->ct32, need_ct * sizeof(float)));
      s->ct32_cap = need_ct;
   }
   if (slots > s->slot_cap) {
      pool_free(c, s->slot_max);
      s->slot_max = nullptr; s->slot_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->slot_max, slots * 2 * sizeof(unsigned long long)));
      s->slot_cap = slots;
   }
   if (nq > s->q_cap) {
      pool_free(c, s->wq); pool_free(c, s->thr_lo); pool_free(c, s->thr_hi); pool_free(c, s->band);
      s->wq = nullptr; s->thr_lo = s->thr_hi = nullptr; s->band = nullptr; s->q_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &s->wq, (size_t)nq * c->D * sizeof(float)));
      KB2E_CUDA(c, pool_alloc(c, &s->thr_lo, (size_t)nq * sizeof(float)));
      KB2E_CUDA(c, pool_alloc(c, &s->thr_hi, (size_t)nq * sizeof(float)));
      s->band_cap = (unsigned int)std::min<long long>(std::max<long long>(1 << 20, 64 * nq), 1ll << 28);
      KB2E_CUDA(c, pool_alloc(c, &s->band, (size_t)s->band_cap * sizeof(int2)));
      s->q_cap = nq;
   }
   if (!s->band_count) {
      KB2E_CUDA(c, pool_alloc(c, &s->band_count, 4 * sizeof(unsigned int)));
      KB2E_CUDA(c, cudaMallocHost(&s->host_count, 4 * sizeof(unsigned int)));
   }
   if (first_pass) KB2E_CUDA(c, cudaMemsetAsync(s->band_count, 0, 4 * sizeof(unsigned int), c->stream));   // overflow flag, call total
   return KB2E_OK;
}

int f32_prepare(kb2e_ctx* c, F32State* s, const double* ct, size_t slots, int ld, long long nq, bool first_pass) {
   int rc = f32_ensure(c, s, slots, ld, nq, first_pass);
   if (rc) return rc;
   KB2E_CUDA(c, cudaMemsetAsync(s->slot_max, 0, slots * 2 * sizeof(unsigned long long), c->stream));
   f32::prep_candidates_kernel<<<dim3(nblk3(ld, 256), (unsigned)slots), 256, 0, c->stream>>>(ct, s->ct32, c->nE, c->D, ld, s->slot_max);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

// The all-candidates kernel alone, on s->ct32 / s->wq / s->thr_* as they stand (band list reset first).
int f32_main(kb2e_ctx* c, F32State* s, bool l2, int ld, const int4* tiles, unsigned ntiles, int32_t* q_cnt, cudaEvent_t e0, cudaEvent_t e1) {
   KB2E_CUDA(c, cudaMemsetAsync(s->band_count, 0, sizeof(unsigned int), c->stream));
   F32Args a;
   a.ct32 = s->ct32; a.wq = s->wq; a.thr_lo = s->thr_lo; a.thr_hi = s->thr_hi; a.tiles = tiles;
   a.q_less = q_cnt; a.band = s->band; a.band_count = s->band_count; a.band_cap = s->band_cap;
   a.nE = c->nE; a.D = c->D; a.ld = ld;
   const int steps = (c->nE + f32::THREADS * f32::CPT - 1) / (f32::THREADS * f32::CPT);
   long long splits = std::max<long long>(1, (2ll * c->num_sms + ntiles - 1) / ntiles);
   splits = std::min<long long>(splits, steps);
   a.splits = (int)splits;
   const size_t smem = (size_t)c->D * f32::QT * sizeof(float);
   if (smem > 200 * 1024) return fail(c, KB2E_ERR_LIMIT, "embedding size too large for the fp32 ranking kernel");
   KB2E_CUDA(c, cudaFuncSetAttribute(f32::rank_f32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   KB2E_CUDA(c, cudaFuncSetAttribute(f32::rank_f32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   KB2E_CUDA(c, cudaEventRecord(e0, c->stream));
   if (l2) f32::rank_f32_kernel<1><<<dim3(ntiles, (unsigned)splits), f32::THREADS, smem, c->stream>>>(a);
   else f32::rank_f32_kernel<0><<<dim3(ntiles, (unsigned)splits), f32::THREADS, smem, c->stream>>>(a);
   KB2E_CUDA(c, cudaEventRecord(e1, c->stream));
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int f32_run(kb2e_ctx* c, F32State* s, bool l2, const double* ct, int ld, const int32_t* q_int, long long nq_total, const double* q_etrue,
            long long q_begin, long long q_end, const int4* tiles, unsigned ntiles, int32_t* q_cnt, cudaEvent_t e0, cudaEvent_t e1) {
   const int32_t* q_fixed = q_int;
   const int32_t* q_truth = q_int + nq_total;
   const int32_t* q_rel = q_int + 2 * nq_total;
   const int32_t* q_side = q_int + 3 * nq_total;
   const int32_t* q_slot = q_int + 4 * nq_total;
   const long long pq = q_end - q_begin;
   if (l2)
      f32::prep_queries_kernel<1><<<nblk3(pq * 32, 256), 256, 0, c->stream>>>(ct, c->rel64, q_fixed, q_rel, q_side, q_slot, q_etrue, s->slot_max,
                                                                             q_begin, q_end, c->D, ld, s->wq, s->thr_lo, s->thr_hi);
   else
      f32::prep_queries_kernel<0><<<nblk3(pq * 32, 256), 256, 0, c->stream>>>(ct, c->rel64, q_fixed, q_rel, q_side, q_slot, q_etrue, s->slot_max,
                                                                             q_begin, q_end, c->D, ld, s->wq, s->thr_lo, s->thr_hi);
   int rc = f32_main(c, s, l2, ld, tiles, ntiles, q_cnt, e0, e1);
   if (rc) return rc;
   if (l2)
      f32::recheck_ct_kernel<1><<<4 * c->num_sms, 128, 0, c->stream>>>(s->band, s->band_count, s->band_cap, ct, c->rel64, q_fixed, q_truth, q_rel,
                                                                       q_side, q_slot, q_etrue, q_cnt, nq_total, c->D, ld);
   else
      f32::recheck_ct_kernel<0><<<4 * c->num_sms, 128, 0, c->stream>>>(s->band, s->band_count, s->band_cap, ct, c->rel64, q_fixed, q_truth, q_rel,
                                                                       q_side, q_slot, q_etrue, q_cnt, nq_total, c->D, ld);
   KB2E_CUDA(c, cudaGetLastError());
   // the largest band of the call decides whether the list overflowed (checked by the caller after its one synchronisation)
   KB2E_CUDA(c, cudaMemcpyAsync(s->host_count, s->band_count, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
   return KB2E_OK;
}

void f32_free(kb2e_ctx* c, F32State* s) {
   pool_free(c, s->ct32); pool_free(c, s->slot_max); pool_free(c, s->wq); pool_free(c, s->thr_lo); pool_free(c, s->thr_hi);
   pool_free(c, s->band); pool_free(c, s->band_count);
   if (s->host_count) cudaFreeHost(s->host_count);
   *s = F32State();
}

}  // namespace kb2e
