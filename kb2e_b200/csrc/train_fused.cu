// TransE epochs with ONE grid barrier per batch for the small-batch shapes (FB15k / WN18: a batch is a few
// microseconds of work, so the two barriers and the stamp scan of train.cu are more than half of it).
//
// Same semantics as train_kernel (train.cu) -- directions from the pre-batch snapshot, all deltas accumulated,
// every touched row normalised ONCE -- but the publish of a row is done by whichever sample finishes with it
// last, inside phase 1, instead of by a second grid-wide phase:
//   * the sampler does not depend on the embeddings, so the rows a batch will reference are known in advance:
//     between "arrive" and "wait" of the barrier that ends batch b every group draws its sample of batch b + 2
//     and adds 1 to pend[(b + 2) % 3][row] for its four rows (h, t, corrupting entity, relation)
//   * in batch b a group gathers its rows, scores, adds its update into the delta rows with vector REDs and stamps
//     the rows it changed; then one lane per row issues fence + atomic decrement of pend[b % 3][row]
//   * the group that takes a row's counter to zero knows that every reader of the row in this batch has read it and
//     every update has landed (release / acquire through the counter), so it folds the row on the spot:
//     row += delta, delta = 0, normalise, store -- if the row was stamped at all
//   * one grid barrier ends the batch (the next batch gathers the folded rows); the counters are back at zero by
//     themselves
// Measured on B200 (profiles/README.md): a win when the batch is small against the machine (WN18 shape, 1,414 samples:
// 13.4 -> 7.4 us per batch); at FB15k shape (4,831 samples = two thirds of the resident groups) the per-row atomics
// and the fold on the sample's critical path cost as much as the barrier they save (11.5 vs 10.4 us), so the host
// picks this kernel only when the batch needs at most half of the resident groups.  Sharding the counters of the hot
// relation rows (32 sub-counters + a top counter) was tried and made it slower, not faster.

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "train_device.cuh"

namespace kb2e {

struct FusedArgs {
   TrainArgs base;
   uint32_t* pend;   // [3][nE + nR] outstanding references of the row in batch b, b + 1, b + 2 (by b % 3)
};

__device__ __forceinline__ void red_inc_u32(uint32_t* p) {
   asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(p) : "memory");
}
// release (this group's REDs / stamps are performed) + acquire (the other groups' are visible if we are last)
__device__ __forceinline__ uint32_t atom_dec_acq_rel(uint32_t* p) {
   uint32_t old;
   asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 0xffffffff;" : "=r"(old) : "l"(p) : "memory");
   return old;
}

__device__ __forceinline__ long long pair_row(const TrainArgs& a, const Pair& s, int w) {
   return w == 0 ? (long long)s.h : (w == 1 ? (long long)s.t : (w == 2 ? (long long)s.c : (long long)a.nE + s.r));
}

template <int LPS, int NV, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) train_fused_kernel(const __grid_constant__ FusedArgs fa) {
   __shared__ double s_loss[THREADS / 32];
   const TrainArgs& a = fa.base;
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const int gshift = (lane / LPS) * LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << gshift);
   const int groups_per_block = blockDim.x / LPS;
   const long long g0 = (long long)(threadIdx.x / LPS) * gridDim.x + blockIdx.x;   // round-robin over CTAs
   const long long R = (long long)a.nE + a.nR;
   const int P = a.P, D = a.D;
   const bool has = g0 < a.batchsize;   // one sample per group and batch (the host checks batchsize <= #groups)
   const uint32_t gb_first = (uint32_t)a.first_epoch * (uint32_t)a.batches;
   const uint32_t total = (uint32_t)a.n_epochs * (uint32_t)a.batches;
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   int trace_slot = 0;
#define KB2E_FTRACE()                                                                                 \
   if (a.trace != nullptr && threadIdx.x == 0 && trace_slot < kTraceSlots) {                          \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      a.trace[(size_t)blockIdx.x * kTraceSlots + trace_slot++] = t_;                                  \
   }
   auto reg_rows = [&](const Pair& s, uint32_t gb) {
      if (gl < 4) red_inc_u32(fa.pend + (size_t)(gb % 3u) * R + pair_row(a, s, gl));
   };
   // prologue: the samples of the first two batches, registered before anybody starts
   Pair cur, nxt;
   cur.h = cur.t = cur.r = cur.c = 0; cur.corruptTail = false;
   nxt = cur;
   if (has) {
      cur = draw_pair(a, (uint32_t)g0, gb_first);
      reg_rows(cur, gb_first);
      if (total > 1u) {
         nxt = draw_pair(a, (uint32_t)g0, gb_first + 1u);
         reg_rows(nxt, gb_first + 1u);
      }
   }
   grid_barrier(a.barrier, bar_target);

   uint32_t rel_batch = 0;
   for (int ep = 0; ep < a.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < a.batches; batch++, rel_batch++) {
         const uint32_t gb = gb_first + rel_batch;
         const uint32_t stamp = a.stamp_base + rel_batch + 1u;   // counts the batches this context has run (train.cu)
         uint32_t* pend = fa.pend + (size_t)(gb % 3u) * R;
         KB2E_FTRACE();
         if (has) {
            const Pair s = cur;
            const long long xr = (long long)a.nE + s.r;
            float4 vh[NV], vt[NV], vc[NV], vr[NV];
            load_row<LPS, NV>(a.tab + (size_t)s.h * P, P, gl, vh);
            load_row<LPS, NV>(a.tab + (size_t)s.t * P, P, gl, vt);
            load_row<LPS, NV>(a.tab + (size_t)s.c * P, P, gl, vc);
            load_row<LPS, NV>(a.tab + (size_t)xr * P, P, gl, vr);
            const bool l1 = a.distance == KB2E_DISTANCE_L1;
            float4 rp[NV], rn[NV];
            float ep_ = 0.f, en_ = 0.f;
#pragma unroll
            for (int q = 0; q < NV; q++) {
               rp[q] = (vt[q] - vh[q]) - vr[q];
               rn[q] = s.corruptTail ? (vc[q] - vh[q]) - vr[q] : (vt[q] - vc[q]) - vr[q];
               if (l1) { ep_ += abs4(rp[q]); en_ += abs4(rn[q]); }
               else { ep_ += dot4(rp[q], rp[q]); en_ += dot4(rn[q], rn[q]); }
            }
            ep_ = gsum<LPS>(ep_, gmask);
            en_ = gsum<LPS>(en_, gmask);
            if (ep_ + a.margin > en_) {   // common/trainer.cpp:138, strict '>'
               if (gl == 0) {
                  loss_acc += (double)(a.margin + ep_ - en_);
                  active_acc++;
               }
               const float lr = a.lr;
               float4 gp[NV], gn[NV], u[NV];
#pragma unroll
               for (int q = 0; q < NV; q++) {
                  const int idx = (q * LPS + gl) * 4;
                  if (l1) { gp[q] = lr * sign4(rp[q], idx, D); gn[q] = lr * sign4(rn[q], idx, D); }
                  else { gp[q] = (2.f * lr) * rp[q]; gn[q] = (2.f * lr) * rn[q]; }
               }
               float* dh = a.dtab + (size_t)s.h * P;
               float* dt = a.dtab + (size_t)s.t * P;
               float* dc = a.dtab + (size_t)s.c * P;
               float* dr = a.dtab + (size_t)xr * P;
               // transe/trainer.cpp:37-41: relation -= m*lr*x, head -= m*lr*x, tail += m*lr*x, m = -1 positive / +1 negative
#pragma unroll
               for (int q = 0; q < NV; q++) u[q] = gp[q] - gn[q];
               red_row<LPS, NV>(dr, P, gl, u);
               if (s.corruptTail) {
                  red_row<LPS, NV>(dh, P, gl, u);
#pragma unroll
                  for (int q = 0; q < NV; q++) u[q] = -1.f * gp[q];
                  red_row<LPS, NV>(dt, P, gl, u);
                  red_row<LPS, NV>(dc, P, gl, gn);
               } else {
                  red_row<LPS, NV>(dh, P, gl, gp);
#pragma unroll
                  for (int q = 0; q < NV; q++) u[q] = gn[q] - gp[q];
                  red_row<LPS, NV>(dt, P, gl, u);
#pragma unroll
                  for (int q = 0; q < NV; q++) u[q] = -1.f * gn[q];
                  red_row<LPS, NV>(dc, P, gl, u);
               }
               if (gl < 4) a.flag[pair_row(a, s, gl)] = stamp;
            }
            // this group is done with its four rows: the lane that owns a row releases it; whoever takes the
            // counter to zero folds the row
            __syncwarp(gmask);
            bool last = false;
            if (gl < 4) {
               const long long row = pair_row(a, s, gl);
               if (atom_dec_acq_rel(pend + row) == 1u) last = __ldcg(a.flag + row) == stamp;
            }
            uint32_t m = (__ballot_sync(gmask, last) >> gshift) & 0xfu;
            while (m) {
               const int w0 = __ffs(m) - 1;
               m &= m - 1;
               int w1 = -1;
               if (m) { w1 = __ffs(m) - 1; m &= m - 1; }
               const long long r0 = pair_row(a, s, w0);
               const long long r1 = w1 >= 0 ? pair_row(a, s, w1) : -1;
               float4 x0[NV], d0[NV], x1[NV], d1[NV];
               load_row<LPS, NV>(a.tab + (size_t)r0 * P, P, gl, x0);
               load_row<LPS, NV>(a.dtab + (size_t)r0 * P, P, gl, d0);
               if (r1 >= 0) {
                  load_row<LPS, NV>(a.tab + (size_t)r1 * P, P, gl, x1);
                  load_row<LPS, NV>(a.dtab + (size_t)r1 * P, P, gl, d1);
               }
               auto fold = [&](long long r, float4 (&x)[NV], float4 (&d)[NV]) {
#pragma unroll
                  for (int q = 0; q < NV; q++) { x[q] = x[q] + d[q]; d[q] = f4(0.f); }
                  store_row<LPS, NV>(a.dtab + (size_t)r * P, P, gl, d);
                  norm_row<LPS, NV>(x, true, gmask);   // transe/trainer.cpp:43-45, once per batch
                  store_row<LPS, NV>(a.tab + (size_t)r * P, P, gl, x);
                  if (gl == 0) { if (r >= a.nE) trel_acc++; else tent_acc++; }
               };
               fold(r0, x0, d0);
               if (r1 >= 0) fold(r1, x1, d1);
            }
         }
         KB2E_FTRACE();
         grid_arrive(a.barrier, bar_target);
         // two batches ahead: draw and register while the other CTAs arrive
         cur = nxt;
         if (has && rel_batch + 2u < total) {
            nxt = draw_pair(a, (uint32_t)g0, gb + 2u);
            reg_rows(nxt, gb + 2u);
         }
         grid_wait(a.barrier, bar_target);
      }
      double v = (gl == 0) ? loss_acc : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_loss[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
         double t = 0.0;
         for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += s_loss[i];
         if (t != 0.0) atomicAdd(a.loss + ep, t);
      }
      __syncthreads();
   }
   uint32_t c0 = (gl == 0) ? active_acc : 0u, c1 = (gl == 0) ? tent_acc : 0u, c2 = (gl == 0) ? trel_acc : 0u;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
   }
   if (lane == 0) {
      if (c0) atomicAdd(a.counters + 0, (unsigned long long)c0);
      if (c1) atomicAdd(a.counters + 1, (unsigned long long)c1);
      if (c2) atomicAdd(a.counters + 2, (unsigned long long)c2);
   }
   (void)groups_per_block;
}

// ---- host ------------------------------------------------------------------------------------------------
bool train_fused_wanted(const kb2e_ctx* c, long long batchsize, int lps, int threads) {
   // Opt-in since round 2 (KB2E_TRAIN_FUSED=1: whenever a batch fits in one pass; =2: only for batches that need at most
   // half of the resident groups, the round-1 default): with touched-row lists and the pipelined sampler the two-barrier
   // kernel of train.cu is faster at every measured shape (WN18 shape, TransE size 100: 6.7 vs 7.4 us per batch).
   const char* env = getenv("KB2E_TRAIN_FUSED");
   if (!env || atoi(env) == 0) return false;
   if (c->cfg.model != KB2E_MODEL_TRANSE) return false;
   if (threads != 1024 && threads != 768 && threads != 512) return false;   // (640: the list kernel of train.cu)   // KB2E_TRAIN_THREADS override without an instantiation
   const long long groups = (long long)c->num_sms * (threads / lps);
   if (env && atoi(env) == 1) return batchsize <= groups;
   return 2 * batchsize <= groups;
}

int train_fused_launch(kb2e_ctx* c, const TrainArgs& base, int lps, int nv, int threads) {
   const size_t R = (size_t)c->nE + c->nR;
   if (!c->pend) KB2E_CUDA(c, pool_alloc(c, &c->pend, 3 * R * sizeof(uint32_t)));
   KB2E_CUDA(c, cudaMemsetAsync(c->pend, 0, 3 * R * sizeof(uint32_t), c->stream));
   FusedArgs a;
   a.base = base;
   a.pend = c->pend;
   void (*k)(const FusedArgs) = nullptr;
#define KB2E_FUSED(L_, N_, T_) if (lps == L_ && nv == N_ && threads == T_) k = train_fused_kernel<L_, N_, T_>;
   KB2E_FUSED(8, 1, 1024) KB2E_FUSED(16, 1, 1024) KB2E_FUSED(32, 1, 1024)
   KB2E_FUSED(8, 2, 768) KB2E_FUSED(16, 2, 768) KB2E_FUSED(32, 2, 768)
   KB2E_FUSED(8, 4, 512) KB2E_FUSED(16, 4, 512) KB2E_FUSED(32, 4, 512)
#undef KB2E_FUSED
   if (!k) return KB2E_ERR_LIMIT;   // no instantiation for this (lanes, vectors, threads): the caller falls back to train_kernel
   void* params[] = {&a};
   KB2E_CUDA(c, cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(threads), params, 0, c->stream));
   return KB2E_OK;
}

}  // namespace kb2e
