// TransE epochs with ONE grid barrier per batch ("lazy publish") for the small-batch shapes
// (FB15k / WN18: a batch is a few microseconds of work, so the two barriers of train.cu are >50 % of it).
//
// Same semantics as train_kernel (train.cu) -- directions from the pre-batch snapshot, all deltas
// accumulated, every touched row normalised once -- but the publish of batch b runs CONCURRENTLY with
// the samples of batch b+1, on different warps of the same persistent launch:
//   * two value buffers V[0], V[1]; three delta buffers D[b % 3]; per row a touch stamp T[b % 3][row]
//     (= batch + 1 of the last batch that added into D[b % 3][row]) and a location word
//     L[row] = (folded_stamp << 1) | loc, written only by the materialiser
//   * a sample of batch b reads row x as  V[loc][x]                          if x was not touched in b-1,
//                                          norm(V[loc][x] + D[(b-1)%3][x])   if it was and is not folded yet,
//                                          V[loc'][x] (already folded)        if the materialiser got there first
//     -- all three give the same number, and nothing a reader looks at is modified during the batch:
//     the materialiser writes the OTHER value buffer, fences, then flips L[x] with one store
//   * the materialiser warps fold the rows touched in b-1 and zero the delta rows of b-2
//   * one grid barrier ends the batch
// At the end of the launch the rows are folded completely and copied back into V[0] (= ctx->tab), so the
// rest of the library never sees the double buffering.

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "train_device.cuh"

namespace kb2e {

constexpr int kLazyThreads = 768;

struct LazyArgs {
   TrainArgs base;
   float* V1;          // second value buffer [R][P]
   float* Dbuf[3];     // delta buffers [R][P]; Dbuf[0] = ctx->dtab
   uint32_t* T;        // [3][R] touch stamps
   uint32_t* L;        // [R] (folded stamp << 1) | loc
   int sample_groups;  // groups per CTA that process samples; the rest materialise
};

// value of row x as the previous batch (stamp Sp; 0 = there is none in this launch) left it
template <int LPS, int NV>
__device__ __forceinline__ void lazy_load(const LazyArgs& a, long long x, uint32_t Sp, const uint32_t* Tprev, const float* Dprev,
                                          int gl, uint32_t gmask, float4 (&v)[NV]) {
   const int P = a.base.P;
   const uint32_t tp = __ldcg(Tprev + x);
   const uint32_t l = __ldcg(a.L + x);
   const float* base = ((l & 1u) ? a.V1 : a.base.tab) + (size_t)x * P;
   load_row<LPS, NV>(base, P, gl, v);
   if (Sp != 0u && tp == Sp && (l >> 1) != Sp) {
      float4 d[NV];
      load_row<LPS, NV>(Dprev + (size_t)x * P, P, gl, d);
#pragma unroll
      for (int q = 0; q < NV; q++) v[q] = v[q] + d[q];
      norm_row<LPS, NV>(v, true, gmask);   // transe/trainer.cpp:43-45, once per batch
   }
}

template <int LPS, int NV>
__device__ __forceinline__ void lazy_process_pair(const LazyArgs& a, const Pair s, uint32_t S, uint32_t Sp, const uint32_t* Tprev, const float* Dprev,
                                                  uint32_t* Tcur, float* Dcur, int gl, uint32_t gmask, double& loss_acc, uint32_t& active_acc) {
   const TrainArgs& b = a.base;
   const int P = b.P, D = b.D;
   const long long xr = (long long)b.nE + s.r;
   float4 vh[NV], vt[NV], vc[NV], vr[NV];
   lazy_load<LPS, NV>(a, s.h, Sp, Tprev, Dprev, gl, gmask, vh);
   lazy_load<LPS, NV>(a, s.t, Sp, Tprev, Dprev, gl, gmask, vt);
   lazy_load<LPS, NV>(a, s.c, Sp, Tprev, Dprev, gl, gmask, vc);
   lazy_load<LPS, NV>(a, xr, Sp, Tprev, Dprev, gl, gmask, vr);
   const bool l1 = b.distance == KB2E_DISTANCE_L1;
   float4 rp[NV], rn[NV];
   float ep = 0.f, en = 0.f;
#pragma unroll
   for (int q = 0; q < NV; q++) {
      rp[q] = (vt[q] - vh[q]) - vr[q];
      rn[q] = s.corruptTail ? (vc[q] - vh[q]) - vr[q] : (vt[q] - vc[q]) - vr[q];
      if (l1) { ep += abs4(rp[q]); en += abs4(rn[q]); }
      else { ep += dot4(rp[q], rp[q]); en += dot4(rn[q], rn[q]); }
   }
   ep = gsum<LPS>(ep, gmask);
   en = gsum<LPS>(en, gmask);
   if (!(ep + b.margin > en)) return;   // common/trainer.cpp:138
   if (gl == 0) {
      loss_acc += (double)(b.margin + ep - en);
      active_acc++;
   }
   const float lr = b.lr;
   float4 gp[NV], gn[NV], u[NV];
#pragma unroll
   for (int q = 0; q < NV; q++) {
      int idx = (q * LPS + gl) * 4;
      if (l1) { gp[q] = lr * sign4(rp[q], idx, D); gn[q] = lr * sign4(rn[q], idx, D); }
      else { gp[q] = (2.f * lr) * rp[q]; gn[q] = (2.f * lr) * rn[q]; }
   }
   float* dh = Dcur + (size_t)s.h * P;
   float* dt = Dcur + (size_t)s.t * P;
   float* dc = Dcur + (size_t)s.c * P;
   float* dr = Dcur + (size_t)xr * P;
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
   if (gl < 4) Tcur[gl == 0 ? (long long)s.h : (gl == 1 ? (long long)s.t : (gl == 2 ? (long long)s.c : xr))] = S;
}

// fold the rows touched in the previous batch (stamp Sp) and zero the delta rows of the batch before it (Spp)
template <int LPS, int NV>
__device__ __forceinline__ void lazy_materialise(const LazyArgs& a, long long first, long long end, uint32_t Sp, const uint32_t* Tp,
                                                 float* Dp, uint32_t Spp, const uint32_t* Tpp, float* Dpp, int gl, uint32_t gmask,
                                                 int lane, uint32_t& tent, uint32_t& trel) {
   const int P = a.base.P;
   auto stamped = [&](long long r) {
      return (Sp != 0u && __ldcg(Tp + r) == Sp) || (Spp != 0u && __ldcg(Tpp + r) == Spp);
   };
   for_stamped_rows<LPS>(first, end, gl, gmask, lane, stamped, [&](long long r0, long long r1) {
#pragma unroll 1
      for (int k = 0; k < 2; k++) {
         const long long r = k == 0 ? r0 : r1;
         if (r < 0) break;
         if (Spp != 0u && __ldcg(Tpp + r) == Spp) {
            float4 z[NV];
#pragma unroll
            for (int q = 0; q < NV; q++) z[q] = f4(0.f);
            store_row<LPS, NV>(Dpp + (size_t)r * P, P, gl, z);
         }
         if (Sp != 0u && __ldcg(Tp + r) == Sp) {
            const uint32_t l = __ldcg(a.L + r);
            if ((l >> 1) == Sp) continue;   // already folded (end-of-launch pass after a partial batch)
            const uint32_t loc = l & 1u;
            float4 x[NV], d[NV];
            load_row<LPS, NV>((loc ? a.V1 : a.base.tab) + (size_t)r * P, P, gl, x);
            load_row<LPS, NV>(Dp + (size_t)r * P, P, gl, d);
#pragma unroll
            for (int q = 0; q < NV; q++) x[q] = x[q] + d[q];
            norm_row<LPS, NV>(x, true, gmask);
            store_row<LPS, NV>((loc ? a.base.tab : a.V1) + (size_t)r * P, P, gl, x);
            __threadfence();                 // the new value is visible before the location flips
            __syncwarp(gmask);
            if (gl == 0) {
               __stcg(a.L + r, (Sp << 1) | (loc ^ 1u));
               if (r >= a.base.nE) trel++; else tent++;
            }
         }
      }
   });
}

template <int LPS, int NV>
__global__ void __launch_bounds__(kLazyThreads, 1) train_lazy_kernel(const __grid_constant__ LazyArgs a) {
   __shared__ double s_loss[kLazyThreads / 32];
   const TrainArgs& b = a.base;
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((lane / LPS) * LPS));
   const int groups_per_block = blockDim.x / LPS;
   const int grp = threadIdx.x / LPS;
   const bool sampler = grp < a.sample_groups;
   // sample groups and materialiser groups are both dealt round-robin over the CTAs
   const long long Gs = (long long)gridDim.x * a.sample_groups;
   const long long gs0 = (long long)grp * gridDim.x + blockIdx.x;
   const int mat_groups = groups_per_block - a.sample_groups;
   const long long Gm = (long long)gridDim.x * mat_groups;
   const long long gm0 = (long long)(grp - a.sample_groups) * gridDim.x + blockIdx.x;
   const long long Gall = (long long)gridDim.x * groups_per_block;
   const long long gall0 = (long long)grp * gridDim.x + blockIdx.x;
   const long long R = (long long)b.nE + b.nR;
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   const uint32_t gb_first = (uint32_t)b.first_epoch * (uint32_t)b.batches;
   long long mfirst = 0, mend = 0;
   if (!sampler) group_range(0, R, gm0, Gm, mfirst, mend);
   Pair pre;
   const bool has_first = sampler && gs0 < b.batchsize;
   if (has_first) pre = draw_pair(b, (uint32_t)gs0, gb_first);
   uint32_t S = 0;

   for (int ep = 0; ep < b.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < b.batches; batch++) {
         const uint32_t gb = gb_first + (uint32_t)ep * (uint32_t)b.batches + (uint32_t)batch;
         const uint32_t rel = (uint32_t)(ep * b.batches + batch);   // batches since the launch started
         S = gb + 1u;
         const uint32_t Sp = rel >= 1u ? S - 1u : 0u;    // stamps only count within this launch (T, L start zeroed)
         const uint32_t Spp = rel >= 2u ? S - 2u : 0u;
         const int ic = (int)(gb % 3u), ip = (int)((gb + 2u) % 3u), ipp = (int)((gb + 1u) % 3u);
         uint32_t* Tcur = a.T + (size_t)ic * R;
         const uint32_t* Tp = a.T + (size_t)ip * R;
         const uint32_t* Tpp = a.T + (size_t)ipp * R;
         if (sampler) {
            if (has_first) lazy_process_pair<LPS, NV>(a, pre, S, Sp, Tp, a.Dbuf[ip], Tcur, a.Dbuf[ic], gl, gmask, loss_acc, active_acc);
            for (long long k = gs0 + Gs; k < b.batchsize; k += Gs) {
               Pair s = draw_pair(b, (uint32_t)k, gb);
               lazy_process_pair<LPS, NV>(a, s, S, Sp, Tp, a.Dbuf[ip], Tcur, a.Dbuf[ic], gl, gmask, loss_acc, active_acc);
            }
         } else {
            lazy_materialise<LPS, NV>(a, mfirst, mend, Sp, Tp, a.Dbuf[ip], Spp, Tpp, a.Dbuf[ipp], gl, gmask, lane, tent_acc, trel_acc);
         }
         grid_arrive(b.barrier, bar_target);
         if (has_first && !(ep == b.n_epochs - 1 && batch == b.batches - 1)) pre = draw_pair(b, (uint32_t)gs0, gb + 1u);
         grid_wait(b.barrier, bar_target);
      }
      double v = (sampler && gl == 0) ? loss_acc : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_loss[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
         double t = 0.0;
         for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += s_loss[i];
         if (t != 0.0) atomicAdd(b.loss + ep, t);
      }
      __syncthreads();
   }
   // ---- end of launch: fold what is still pending, zero the deltas, bring every row home to V[0] ----
   {
      const uint32_t total = (uint32_t)(b.n_epochs * b.batches);
      const uint32_t gb_last = gb_first + total - 1u;
      const int ic = (int)(gb_last % 3u), ip = (int)((gb_last + 2u) % 3u);
      long long first, end;
      group_range(0, R, gall0, Gall, first, end);
      // pass A: rows touched in the last batch are folded; the deltas of the batch before it are zeroed
      lazy_materialise<LPS, NV>(a, first, end, S, a.T + (size_t)ic * R, a.Dbuf[ic], total >= 2u ? S - 1u : 0u, a.T + (size_t)ip * R,
                                a.Dbuf[ip], gl, gmask, lane, tent_acc, trel_acc);
      grid_barrier(b.barrier, bar_target);
      // pass B: zero the last batch's deltas; rows living in V[1] go back to V[0]
      const int P = b.P;
      const uint32_t* Tl = a.T + (size_t)ic * R;
      for_stamped_rows<LPS>(first, end, gl, gmask, lane,
                            [&](long long r) { return __ldcg(Tl + r) == S || (__ldcg(a.L + r) & 1u); },
                            [&](long long r0, long long r1) {
#pragma unroll 1
         for (int k = 0; k < 2; k++) {
            const long long r = k == 0 ? r0 : r1;
            if (r < 0) break;
            if (__ldcg(Tl + r) == S) {
               float4 z[NV];
#pragma unroll
               for (int q = 0; q < NV; q++) z[q] = f4(0.f);
               store_row<LPS, NV>(a.Dbuf[ic] + (size_t)r * P, P, gl, z);
            }
            if (__ldcg(a.L + r) & 1u) {
               float4 x[NV];
               load_row<LPS, NV>(a.V1 + (size_t)r * P, P, gl, x);
               store_row<LPS, NV>(b.tab + (size_t)r * P, P, gl, x);
            }
         }
      });
   }
   uint32_t c0 = (sampler && gl == 0) ? active_acc : 0u, c1 = tent_acc, c2 = trel_acc;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o);
   }
   if (lane == 0) {
      if (c0) atomicAdd(b.counters + 0, (unsigned long long)c0);
      if (c1) atomicAdd(b.counters + 1, (unsigned long long)c1);
      if (c2) atomicAdd(b.counters + 2, (unsigned long long)c2);
   }
}

// ---- host ----------------------------------------------------------------------------------------------
struct LazyBuffers {
   float* V1 = nullptr;
   float* D1 = nullptr;
   float* D2 = nullptr;
   uint32_t* T = nullptr;
   uint32_t* L = nullptr;
};

bool train_lazy_wanted(const kb2e_ctx* c, long long batchsize, int lps, int nv) {
   // Experimental, opt-in (KB2E_TRAIN_LAZY=1; =2 forces it for any batch size): measured on B200 at FB15k shape
   // it is SLOWER than the two-barrier kernel (18.7 vs 10.4 us per batch, profiles/README.md): the extra stamp /
   // location loads in front of every row gather and the per-row fence of the materialiser cost more than the
   // barrier they remove.
   const char* env = getenv("KB2E_TRAIN_LAZY");
   if (!env || atoi(env) == 0) return false;
   if (c->cfg.model != KB2E_MODEL_TRANSE || nv > 2) return false;
   // worth it only while a batch fits in about one pass over the resident groups (barrier-bound regime)
   const long long groups = (long long)c->num_sms * (kLazyThreads / lps);
   return batchsize <= groups * 4 / 5 || atoi(env) == 2;
}

int train_lazy_launch(kb2e_ctx* c, const TrainArgs& base, int lps, int nv, int* threads_out) {
   if (!c->lazy) c->lazy = new LazyBuffers();
   LazyBuffers* lb = c->lazy;
   const size_t R = (size_t)c->nE + c->nR;
   const size_t tab_bytes = R * c->P * sizeof(float);
   if (!lb->V1) {
      KB2E_CUDA(c, pool_alloc(c, &lb->V1, tab_bytes));
      KB2E_CUDA(c, pool_alloc(c, &lb->D1, tab_bytes));
      KB2E_CUDA(c, pool_alloc(c, &lb->D2, tab_bytes));
      KB2E_CUDA(c, pool_alloc(c, &lb->T, 3 * R * sizeof(uint32_t)));
      KB2E_CUDA(c, pool_alloc(c, &lb->L, R * sizeof(uint32_t)));
      KB2E_CUDA(c, cudaMemsetAsync(lb->D1, 0, tab_bytes, c->stream));
      KB2E_CUDA(c, cudaMemsetAsync(lb->D2, 0, tab_bytes, c->stream));
   }
   KB2E_CUDA(c, cudaMemsetAsync(lb->T, 0, 3 * R * sizeof(uint32_t), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(lb->L, 0, R * sizeof(uint32_t), c->stream));
   LazyArgs a;
   a.base = base;
   a.V1 = lb->V1;
   a.Dbuf[0] = c->dtab; a.Dbuf[1] = lb->D1; a.Dbuf[2] = lb->D2;
   a.T = lb->T; a.L = lb->L;
   const int groups_per_block = kLazyThreads / lps;
   int need = (int)((base.batchsize + c->num_sms - 1) / c->num_sms);
   // at least a quarter of the groups materialise; the samplers take what a one-pass batch needs
   const int gpw = 32 / lps;   // groups per warp: roles are assigned per warp
   a.sample_groups = std::max(1, std::min(need, groups_per_block - std::max(2, groups_per_block / 4)));
   a.sample_groups = std::min(groups_per_block - gpw, (a.sample_groups + gpw - 1) / gpw * gpw);
   void (*k)(const LazyArgs) = nullptr;
#define KB2E_LAZY(L_, N_) if (lps == L_ && nv == N_) k = train_lazy_kernel<L_, N_>;
   KB2E_LAZY(8, 1) KB2E_LAZY(8, 2) KB2E_LAZY(16, 1) KB2E_LAZY(16, 2) KB2E_LAZY(32, 1) KB2E_LAZY(32, 2)
#undef KB2E_LAZY
   if (!k) return fail(c, KB2E_ERR_LIMIT, "no lazy training kernel for this shape");
   void* params[] = {&a};
   *threads_out = kLazyThreads;
   KB2E_CUDA(c, cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(kLazyThreads), params, 0, c->stream));
   return KB2E_OK;
}

void train_lazy_free(kb2e_ctx* c) {
   if (!c->lazy) return;
   pool_free(c, c->lazy->V1); pool_free(c, c->lazy->D1); pool_free(c, c->lazy->D2); pool_free(c, c->lazy->T); pool_free(c, c->lazy->L);
   delete c->lazy;
   c->lazy = nullptr;
}

}  // namespace kb2e
