// Margin-ranking SGD epochs for TransE / TransH / TransR as ONE persistent cooperative launch.
//
// Replaces (citations into eriq-augustine/KB2E):
//   common::Trainer::bfgs            common/trainer.cpp:69-107   epoch x batch x sample loop
//   the negative sampler             common/trainer.cpp:78-98    (counter RNG instead of std::rand)
//   common::Trainer::train_kb        common/trainer.cpp:130-149  hinge, strict '>'
//   transe::tripleEnergy / gradientUpdate   transe/transe.cpp:10-28, transe/trainer.cpp:25-46
//   transh::tripleEnergy / gradientUpdate   transh/transh.cpp:10-29, transh/trainer.cpp:11-59
//   prebatch / postbatch             transe/trainer.cpp:48-56 (whole-table deep copies -> touched rows only)
//   common::norm (both overloads)    common/utils.cpp:70-111
//
// Batch semantics (oracle/kb2e_oracle.c:orc_train_batch_dfr is the CPU twin):
//   phase 1  every sample reads the frozen tables `tab`/`w` (the reference's entityVec_/relationVec_),
//            and adds its update into the zero-based delta tables with vector REDs (the reference's
//            *_next_ minus the snapshot); rows it touches are flagged.
//   phase 2a flagged relation-side rows:  row += delta, normalise ONCE, publish, delta = 0.
//   phase 2b flagged entity rows, likewise (TransH/TransR: plus the soft constraint against the
//            lowest/highest relation id that touched the row).
//   Phases are separated by a grid-wide barrier inside the launch; all table traffic goes through
//   L2 (.cg) because other SMs rewrite the rows between phases.
//
// Work distribution: a "group" of LPS lanes owns one sample (LPS*NV float4 >= row pitch), so a
// D=50 row uses 16 lanes and a D=100 row 16 lanes x 2 vectors or 32 x 1, whichever lets one batch
// fit in one pass over the resident groups (148 CTAs x 1024 threads).

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace kb2e {

constexpr int kTrainThreads = 1024;

struct TrainArgs {
   float* tab;
   float* dtab;
   float* w;
   float* dw;
   uint8_t* flagE;   // [nE]
   uint8_t* flagR;   // [2][nR], indexed by global-batch parity
   int* rmin;
   int* rmax;
   const int4* triples;
   const uint64_t* hash;
   uint64_t hash_mask;
   const double* pr;
   const int32_t* pairs;  // test hook: explicit (pos,neg) pairs instead of the sampler
   uint32_t* barrier;
   double* loss;                  // [n_epochs]
   unsigned long long* counters;  // [0] active, [1] touched_ent, [2] touched_rel
   long long n_train;
   long long batchsize;
   size_t w_row;
   int nE, nR, D, P;
   int batches, first_epoch, n_epochs, distance;
   float lr, margin;
   uint32_t seed_lo, seed_hi, flags;
};

// ---- small float4 helpers ----------------------------------------------------------------------
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float abs4(float4 a) { return fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(a.w); }
// L1 direction: x > 0 ? 1 : -1 (zero maps to -1, transe/trainer.cpp:31-35); padding lanes get 0.
__device__ __forceinline__ float4 sign4(float4 a, int idx, int D) {
   return make_float4(idx + 0 < D ? (a.x > 0.f ? 1.f : -1.f) : 0.f, idx + 1 < D ? (a.y > 0.f ? 1.f : -1.f) : 0.f,
                      idx + 2 < D ? (a.z > 0.f ? 1.f : -1.f) : 0.f, idx + 3 < D ? (a.w > 0.f ? 1.f : -1.f) : 0.f);
}

template <int LPS>
__device__ __forceinline__ float gsum(float v, uint32_t gmask) {
#pragma unroll
   for (int o = LPS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
   return v;
}

template <int LPS, int NV>
__device__ __forceinline__ void load_row(const float* base, int P, int gl, float4 (&v)[NV]) {
#pragma unroll
   for (int q = 0; q < NV; q++) {
      int off = (q * LPS + gl) * 4;
      v[q] = off < P ? ld_cg4(base + off) : f4(0.f);
   }
}

template <int LPS, int NV>
__device__ __forceinline__ void red_row(float* base, int P, int gl, const float4 (&v)[NV]) {
#pragma unroll
   for (int q = 0; q < NV; q++) {
      int off = (q * LPS + gl) * 4;
      if (off < P) red_add4(base + off, v[q]);
   }
}

template <int LPS, int NV>
__device__ __forceinline__ void store_row(float* base, int P, int gl, const float4 (&v)[NV]) {
#pragma unroll
   for (int q = 0; q < NV; q++) {
      int off = (q * LPS + gl) * 4;
      if (off < P) st_cg4(base + off, v[q]);
   }
}

template <int LPS, int NV>
__device__ __forceinline__ float len2_row(const float4 (&v)[NV], uint32_t gmask) {
   float s = 0.f;
#pragma unroll
   for (int q = 0; q < NV; q++) s += dot4(v[q], v[q]);
   return gsum<LPS>(s, gmask);
}

// common::norm(a, ignoreShort), common/utils.cpp:70-77
template <int LPS, int NV>
__device__ __forceinline__ void norm_row(float4 (&v)[NV], bool ignoreShort, uint32_t gmask) {
   float len = sqrtf(len2_row<LPS, NV>(v, gmask));
   if (!ignoreShort || len > 1.f) {
#pragma unroll
      for (int q = 0; q < NV; q++) v[q] = make_float4(v[q].x / len, v[q].y / len, v[q].z / len, v[q].w / len);
   }
}

// The loop of common::norm(a, b, rate), common/utils.cpp:83-108 (`sum` is deliberately not reset
// between iterations, as in the reference).  b must already be unit length.  Returns #corrective steps.
template <int LPS, int NV>
__device__ __forceinline__ int soft_orth_loop(float4 (&a)[NV], float4 (&b)[NV], float rate, uint32_t gmask) {
   float sum = 0.f;
   int iters = 0;
   while (true) {
      sum += len2_row<LPS, NV>(b, gmask);
      sum = sqrtf(sum);
      float x = 0.f;
#pragma unroll
      for (int q = 0; q < NV; q++) {
         b[q] = make_float4(b[q].x / sum, b[q].y / sum, b[q].z / sum, b[q].w / sum);
         x += dot4(a[q], b[q]);
      }
      x = gsum<LPS>(x, gmask);
      if (x > 0.1f && iters < 1000) {
#pragma unroll
         for (int q = 0; q < NV; q++) {
            a[q] = a[q] - rate * b[q];
            b[q] = b[q] - rate * a[q];
         }
         iters++;
      } else {
         break;
      }
   }
   return iters;
}

// ---- grid-wide barrier (all CTAs are co-resident: cooperative launch) ---------------------------
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& target) {
   __syncthreads();
   if (threadIdx.x == 0) {
      target += gridDim.x;
      __threadfence();
      red_release_add_u32(counter, 1u);
      while ((int32_t)(ld_acquire_u32(counter) - target) < 0) {
      }
      __threadfence();
   }
   __syncthreads();
}

// ---- the sampler: common/trainer.cpp:78-98 with a counter RNG -----------------------------------
struct Pair {
   int h, t, r, c;    // positive triple and the corrupting entity
   bool corruptTail;  // true: (h, r, c) is the negative; false: (c, r, t)
};

__device__ __forceinline__ Pair draw_pair(const TrainArgs& a, uint32_t k, uint32_t gb) {
   Pair s;
   if (a.pairs != nullptr) {
      const int32_t* p = a.pairs + 6ll * k;
      s.h = __ldg(p + 0); s.t = __ldg(p + 1); s.r = __ldg(p + 2);
      int nh = __ldg(p + 3), nt = __ldg(p + 4);
      s.corruptTail = (nh == s.h);
      s.c = s.corruptTail ? nt : nh;
      return s;
   }
   uint32_t x[4];
   philox4x32(k, gb, 0u, 0u, a.seed_lo, a.seed_hi, x);
   uint64_t i = mulhi64(((uint64_t)x[0] << 32) | x[1], (uint64_t)a.n_train);
   int4 tr = __ldg(a.triples + i);
   s.h = tr.x; s.t = tr.y; s.r = tr.z;
   int coin = (int)(x[2] % 1000u);
   int j = (int)mulhi32(x[3], (uint32_t)a.nE);
   s.corruptTail = (double)coin < __ldg(a.pr + tr.z);
   for (uint32_t att = 1; att < 64; att++) {
      uint64_t key = s.corruptTail ? pack_triple(s.h, s.r, j) : pack_triple(j, s.r, s.t);
      if (!hash_contains(a.hash, a.hash_mask, key)) break;
      philox4x32(k, gb, att, 0u, a.seed_lo, a.seed_hi, x);
      j = (int)mulhi32(x[0], (uint32_t)a.nE);
   }
   s.c = j;
   return s;
}

// ---- phase 1: one (positive, negative) pair, TransE and TransH -----------------------------------
template <int MODEL, int LPS, int NV>
__device__ __forceinline__ void process_pair(const TrainArgs& a, const Pair s, int gl, uint32_t gmask, uint32_t par,
                                             double& loss_acc, uint32_t& active_acc) {
   const int P = a.P, D = a.D;
   const float* eh = a.tab + (size_t)s.h * P;
   const float* et = a.tab + (size_t)s.t * P;
   const float* ec = a.tab + (size_t)s.c * P;
   const float* er = a.tab + ((size_t)a.nE + s.r) * P;
   float4 vh[NV], vt[NV], vc[NV], vr[NV], vw[NV];
   load_row<LPS, NV>(eh, P, gl, vh);
   load_row<LPS, NV>(et, P, gl, vt);
   load_row<LPS, NV>(ec, P, gl, vc);
   load_row<LPS, NV>(er, P, gl, vr);
   float hs = 0.f, ts = 0.f, cs = 0.f;
   if (MODEL == KB2E_MODEL_TRANSH) {
      load_row<LPS, NV>(a.w + (size_t)s.r * P, P, gl, vw);
#pragma unroll
      for (int q = 0; q < NV; q++) {
         hs += dot4(vw[q], vh[q]);
         ts += dot4(vw[q], vt[q]);
         cs += dot4(vw[q], vc[q]);
      }
      hs = gsum<LPS>(hs, gmask);
      ts = gsum<LPS>(ts, gmask);
      cs = gsum<LPS>(cs, gmask);
      // project onto the hyperplane: e - (w.e) w   (transh/transh.cpp:25)
#pragma unroll
      for (int q = 0; q < NV; q++) {
         vh[q] = vh[q] - hs * vw[q];
         vt[q] = vt[q] - ts * vw[q];
         vc[q] = vc[q] - cs * vw[q];
      }
   }
   const bool l1 = (MODEL == KB2E_MODEL_TRANSH) || a.distance == KB2E_DISTANCE_L1;
   float4 rp[NV], rn[NV];
   float ep = 0.f, en = 0.f;
#pragma unroll
   for (int q = 0; q < NV; q++) {
      rp[q] = (vt[q] - vh[q]) - vr[q];
      rn[q] = s.corruptTail ? (vc[q] - vh[q]) - vr[q] : (vt[q] - vc[q]) - vr[q];
      if (l1) {
         ep += abs4(rp[q]);
         en += abs4(rn[q]);
      } else {
         ep += dot4(rp[q], rp[q]);
         en += dot4(rn[q], rn[q]);
      }
   }
   ep = gsum<LPS>(ep, gmask);
   en = gsum<LPS>(en, gmask);
   // common/trainer.cpp:138: strict '>'
   if (!(ep + a.margin > en)) return;

   if (gl == 0) {
      loss_acc += (double)(a.margin + ep - en);
      active_acc++;
   }
   const float lr = a.lr;
   float4 gp[NV], gn[NV];  // lr * x for the positive / negative triple
#pragma unroll
   for (int q = 0; q < NV; q++) {
      int idx = (q * LPS + gl) * 4;
      if (l1) {
         gp[q] = lr * sign4(rp[q], idx, D);
         gn[q] = lr * sign4(rn[q], idx, D);
      } else {
         gp[q] = (2.f * lr) * rp[q];
         gn[q] = (2.f * lr) * rn[q];
      }
   }
   float* dh = a.dtab + (size_t)s.h * P;
   float* dt = a.dtab + (size_t)s.t * P;
   float* dc = a.dtab + (size_t)s.c * P;
   float* dr = a.dtab + ((size_t)a.nE + s.r) * P;
   float4 u[NV];
   // relation row: -= m*lr*x  with m = -1 (positive), +1 (negative)
#pragma unroll
   for (int q = 0; q < NV; q++) u[q] = gp[q] - gn[q];
   red_row<LPS, NV>(dr, P, gl, u);
   if (s.corruptTail) {
      // negative = (h, r, c): head -= gn, c += gn; positive: head += gp, tail -= gp
      red_row<LPS, NV>(dh, P, gl, u);
#pragma unroll
      for (int q = 0; q < NV; q++) u[q] = -1.f * gp[q];
      red_row<LPS, NV>(dt, P, gl, u);
      red_row<LPS, NV>(dc, P, gl, gn);
   } else {
      // negative = (c, r, t): c -= gn, tail += gn
      red_row<LPS, NV>(dh, P, gl, gp);
#pragma unroll
      for (int q = 0; q < NV; q++) u[q] = gn[q] - gp[q];
      red_row<LPS, NV>(dt, P, gl, u);
#pragma unroll
      for (int q = 0; q < NV; q++) u[q] = -1.f * gn[q];
      red_row<LPS, NV>(dc, P, gl, u);
   }
   if (MODEL == KB2E_MODEL_TRANSH) {
      // transh/trainer.cpp:33,39-40,44-45: w += beta*lr*(x*(hs-ts) + sum_x*(h - t)), with the RAW h, t.
      float sxp = 0.f, sxn = 0.f;
#pragma unroll
      for (int q = 0; q < NV; q++) {
         sxp += dot4(gp[q], vw[q]);
         sxn += dot4(gn[q], vw[q]);
      }
      sxp = gsum<LPS>(sxp, gmask);  // = lr * sum_x (positive)
      sxn = gsum<LPS>(sxn, gmask);
      const float nhs = s.corruptTail ? hs : cs;  // negative triple's head / tail projections
      const float nts = s.corruptTail ? cs : ts;
#pragma unroll
      for (int q = 0; q < NV; q++) {
         // raw rows back from the projected ones: e = p + (w.e) w
         float4 rh = vh[q] + hs * vw[q], rt = vt[q] + ts * vw[q], rc = vc[q] + cs * vw[q];
         float4 nh = s.corruptTail ? rh : rc, nt = s.corruptTail ? rc : rt;
         float4 pos = (hs - ts) * gp[q] + sxp * (rh - rt);
         float4 neg = (nhs - nts) * gn[q] + sxn * (nh - nt);
         u[q] = neg - pos;
      }
      red_row<LPS, NV>(a.dw + (size_t)s.r * P, P, gl, u);
   }
   // flag the touched rows (+ relation range per entity for the TransH/TransR constraints)
   if (gl < 3) {
      int e = gl == 0 ? s.h : (gl == 1 ? s.t : s.c);
      a.flagE[e] = 1;
      if (MODEL != KB2E_MODEL_TRANSE) {
         atomicMin(a.rmin + e, s.r);
         atomicMax(a.rmax + e, s.r);
      }
   } else if (gl == 3) {
      a.flagR[(size_t)par * a.nR + s.r] = 1;
   }
}

// ---- phase 2 -------------------------------------------------------------------------------------
// Relation-side row r: d_r (and w_r).  transe/trainer.cpp:43, transh/trainer.cpp:48,52,56.
template <int MODEL, int LPS, int NV>
__device__ __forceinline__ void publish_relation(const TrainArgs& a, int r, int gl, uint32_t gmask) {
   const int P = a.P;
   float* cur = a.tab + ((size_t)a.nE + r) * P;
   float* del = a.dtab + ((size_t)a.nE + r) * P;
   float4 x[NV], d[NV];
   load_row<LPS, NV>(cur, P, gl, x);
   load_row<LPS, NV>(del, P, gl, d);
#pragma unroll
   for (int q = 0; q < NV; q++) { x[q] = x[q] + d[q]; d[q] = f4(0.f); }
   store_row<LPS, NV>(del, P, gl, d);
   norm_row<LPS, NV>(x, true, gmask);
   if (MODEL == KB2E_MODEL_TRANSH) {
      float* wc = a.w + (size_t)r * P;
      float* wd = a.dw + (size_t)r * P;
      float4 b[NV], db[NV];
      load_row<LPS, NV>(wc, P, gl, b);
      load_row<LPS, NV>(wd, P, gl, db);
#pragma unroll
      for (int q = 0; q < NV; q++) { b[q] = b[q] + db[q]; db[q] = f4(0.f); }
      store_row<LPS, NV>(wd, P, gl, db);
      norm_row<LPS, NV>(b, false, gmask);          // transh/trainer.cpp:52
      norm_row<LPS, NV>(b, false, gmask);          // common/utils.cpp:82
      soft_orth_loop<LPS, NV>(x, b, a.lr, gmask);  // common/utils.cpp:83-108
      norm_row<LPS, NV>(b, false, gmask);          // common/utils.cpp:110
      store_row<LPS, NV>(wc, P, gl, b);
   }
   store_row<LPS, NV>(cur, P, gl, x);
}

// Entity row e.  transe/trainer.cpp:44-45, transh/trainer.cpp:49-50,57-58.
template <int MODEL, int LPS, int NV>
__device__ __forceinline__ void publish_entity(const TrainArgs& a, int e, int gl, uint32_t gmask, uint32_t next_par) {
   const int P = a.P;
   float* cur = a.tab + (size_t)e * P;
   float* del = a.dtab + (size_t)e * P;
   float4 x[NV], d[NV];
   load_row<LPS, NV>(cur, P, gl, x);
   load_row<LPS, NV>(del, P, gl, d);
#pragma unroll
   for (int q = 0; q < NV; q++) { x[q] = x[q] + d[q]; d[q] = f4(0.f); }
   store_row<LPS, NV>(del, P, gl, d);
   norm_row<LPS, NV>(x, true, gmask);
   if (MODEL == KB2E_MODEL_TRANSH) {
      int r0 = __ldcg(a.rmin + e), r1 = __ldcg(a.rmax + e);
      if (gl == 0) { a.rmin[e] = 0x7fffffff; a.rmax[e] = -1; }
      for (int pass = 0; pass < 2; pass++) {
         int r = pass == 0 ? r0 : r1;
         if (pass == 1 && r1 == r0) break;
         float4 b[NV], b0[NV];
         load_row<LPS, NV>(a.w + (size_t)r * P, P, gl, b0);
#pragma unroll
         for (int q = 0; q < NV; q++) b[q] = b0[q];
         int iters = soft_orth_loop<LPS, NV>(x, b, a.lr, gmask);
         if (iters > 0) {
            // The reference also perturbs w_r here (common/utils.cpp:103,110); the perturbation is
            // folded into the NEXT batch's delta so that w_r stays read-only in this phase.
            norm_row<LPS, NV>(b, false, gmask);
#pragma unroll
            for (int q = 0; q < NV; q++) b[q] = b[q] - b0[q];
            red_row<LPS, NV>(a.dw + (size_t)r * P, P, gl, b);
            if (gl == 0) a.flagR[(size_t)next_par * a.nR + r] = 1;
         }
      }
   }
   store_row<LPS, NV>(cur, P, gl, x);
}

template <int MODEL, int LPS, int NV>
__global__ void __launch_bounds__(kTrainThreads, 1) train_kernel(const __grid_constant__ TrainArgs a) {
   __shared__ double s_loss[kTrainThreads / 32];
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((lane / LPS) * LPS));
   const int groups_per_block = blockDim.x / LPS;
   const long long G = (long long)gridDim.x * groups_per_block;
   // Samples (and rows) are dealt round-robin over CTAs so every SM gets an equal share.
   const long long g0 = (long long)(threadIdx.x / LPS) * gridDim.x + blockIdx.x;
   const long long tid_global = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   const long long nthreads = (long long)gridDim.x * blockDim.x;
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;

   for (int ep = 0; ep < a.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < a.batches; batch++) {
         const uint32_t gb = (uint32_t)(a.first_epoch + ep) * (uint32_t)a.batches + (uint32_t)batch;
         const uint32_t par = gb & 1u;
         // ---- phase 1 ----
         // The relation flags of the previous batch (other parity) are dead by now: clear them.
         for (long long i = tid_global; i < a.nR; i += nthreads) a.flagR[(size_t)(par ^ 1u) * a.nR + i] = 0;
         for (long long k = g0; k < a.batchsize; k += G) {
            Pair s = draw_pair(a, (uint32_t)k, gb);
            process_pair<MODEL, LPS, NV>(a, s, gl, gmask, par, loss_acc, active_acc);
         }
         grid_barrier(a.barrier, bar_target);
         // ---- phase 2a: relation-side rows ----
         for (long long r = g0; r < a.nR; r += G) {
            if (__ldcg(a.flagR + (size_t)par * a.nR + r)) {
               publish_relation<MODEL, LPS, NV>(a, (int)r, gl, gmask);
               if (gl == 0) trel_acc++;
            }
         }
         if (MODEL != KB2E_MODEL_TRANSE) grid_barrier(a.barrier, bar_target);
         // ---- phase 2b: entity rows ----
         for (long long e = g0; e < a.nE; e += G) {
            if (__ldcg(a.flagE + e)) {
               publish_entity<MODEL, LPS, NV>(a, (int)e, gl, gmask, par ^ 1u);
               if (gl == 0) { a.flagE[e] = 0; tent_acc++; }
            }
         }
         grid_barrier(a.barrier, bar_target);
      }
      // epoch loss: group leaders -> warp -> block -> one atomic per CTA
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
   // counters
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
}

// ---- test hooks ---------------------------------------------------------------------------------
__global__ void sample_kernel(const TrainArgs a, uint32_t gb, long long count, int32_t* out) {
   long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= count) return;
   Pair s = draw_pair(a, (uint32_t)k, gb);
   int32_t* p = out + 6 * k;
   p[0] = s.h; p[1] = s.t; p[2] = s.r; p[5] = s.r;
   if (s.corruptTail) { p[3] = s.h; p[4] = s.c; } else { p[3] = s.c; p[4] = s.t; }
}

// fp32 energies with exactly the arithmetic of process_pair (one warp per triple).
template <int MODEL>
__global__ void score32_kernel(const float* tab, const float* w, int nE, int D, int P, int distance,
                               const int32_t* h, const int32_t* t, const int32_t* r, long long n, double* out) {
   long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (k >= n) return;
   const float* eh = tab + (size_t)h[k] * P;
   const float* et = tab + (size_t)t[k] * P;
   const float* er = tab + ((size_t)nE + r[k]) * P;
   const float* wr = (MODEL == KB2E_MODEL_TRANSH) ? w + (size_t)r[k] * P : nullptr;
   float hs = 0.f, ts = 0.f;
   if (MODEL == KB2E_MODEL_TRANSH) {
      for (int off = lane * 4; off < P; off += 128) {
         float4 vw = ld_cg4(wr + off);
         hs += dot4(vw, ld_cg4(eh + off));
         ts += dot4(vw, ld_cg4(et + off));
      }
      hs = gsum<32>(hs, 0xffffffffu);
      ts = gsum<32>(ts, 0xffffffffu);
   }
   const bool l1 = (MODEL == KB2E_MODEL_TRANSH) || distance == KB2E_DISTANCE_L1;
   float e = 0.f;
   for (int off = lane * 4; off < P; off += 128) {
      float4 vh = ld_cg4(eh + off), vt = ld_cg4(et + off), vr = ld_cg4(er + off);
      if (MODEL == KB2E_MODEL_TRANSH) {
         float4 vw = ld_cg4(wr + off);
         vh = vh - hs * vw;
         vt = vt - ts * vw;
      }
      float4 res = (vt - vh) - vr;
      e += l1 ? abs4(res) : dot4(res, res);
   }
   e = gsum<32>(e, 0xffffffffu);
   if (lane == 0) out[k] = (double)e;
}

// ---- table maintenance ----------------------------------------------------------------------------
__global__ void widen_kernel(const float* src, double* dst, long long rows, int D, int P) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= rows * D) return;
   dst[i] = (double)src[(i / D) * P + (i % D)];
}

__global__ void narrow_kernel(const double* src, float* dst, long long rows, int D, int P) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= rows * P) return;
   int c = (int)(i % P);
   dst[i] = c < D ? (float)src[(i / P) * D + c] : 0.f;
}

__global__ void hash_insert_kernel(const int4* triples, long long n, uint64_t* table, uint64_t mask) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   int4 tr = triples[i];
   uint64_t key = pack_triple(tr.x, tr.z, tr.y);
   uint64_t slot = mix64(key) & mask;
   while (true) {
      unsigned long long prev = atomicCAS((unsigned long long*)(table + slot), (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (prev == kEmptyKey || prev == key) return;
      slot = (slot + 1) & mask;
   }
}

__global__ void pack_triples_kernel(const int32_t* h, const int32_t* t, const int32_t* r, long long n, int4* out) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) out[i] = make_int4(h[i], t[i], r[i], 0);
}

// Initial values: N(0, (1/D)^2) per element (the reference's rejection-sampled truncated normal never
// truncates), then the row normalisation prepTrain applies (common/trainer.cpp:45-57;
// transh/trainer.cpp:80-87 unit-normalises w_r; transr/trainer.cpp:73-86 sets M_r = I).
__global__ void init_rows_kernel(float* tab, long long rows, int D, int P, uint32_t k0, uint32_t k1, uint32_t stream, int mode) {
   // one warp per row; mode 0: clip to the unit ball, 1: unit length
   long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (row >= rows) return;
   float* p = tab + row * P;
   float s2 = 0.f;
   const float sigma = 1.0f / (float)D;
   for (int base = lane * 4; base < P; base += 128) {
      uint32_t x[4];
      philox4x32((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)base, stream, k0, k1, x);
      float v[4];
#pragma unroll
      for (int q = 0; q < 2; q++) {
         float u1 = ((float)x[2 * q] + 1.0f) * 2.3283064e-10f;  // (0, 1]
         float u2 = (float)x[2 * q + 1] * 2.3283064e-10f;
         float rad = sqrtf(-2.0f * logf(u1));
         v[2 * q] = rad * cospif(2.0f * u2) * sigma;
         v[2 * q + 1] = rad * sinpif(2.0f * u2) * sigma;
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
         if (base + q >= D) v[q] = 0.f;
         s2 += v[q] * v[q];
      }
      *reinterpret_cast<float4*>(p + base) = make_float4(v[0], v[1], v[2], v[3]);
   }
   s2 = gsum<32>(s2, 0xffffffffu);
   float len = sqrtf(s2);
   __syncwarp();
   if (mode == 1 || len > 1.f) {
      for (int base = lane * 4; base < P; base += 128) {
         float4 v = *reinterpret_cast<float4*>(p + base);
         *reinterpret_cast<float4*>(p + base) = make_float4(v.x / len, v.y / len, v.z / len, v.w / len);
      }
   }
}

__global__ void identity_kernel(float* w, long long nR, int D, int P) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= nR * D * P) return;
   int col = (int)(i % P);
   int row = (int)((i / P) % D);
   w[i] = (col == row) ? 1.f : 0.f;
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) p[i] = v;
}

// ---- host side ------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int train_alloc(kb2e_ctx* c) {
   if (c->tab) return KB2E_OK;
   size_t rows = (size_t)c->nE + c->nR;
   KB2E_CUDA(c, cudaMalloc(&c->tab, rows * c->P * sizeof(float)));
   KB2E_CUDA(c, cudaMalloc(&c->dtab, rows * c->P * sizeof(float)));
   KB2E_CUDA(c, cudaMemsetAsync(c->tab, 0, rows * c->P * sizeof(float), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->dtab, 0, rows * c->P * sizeof(float), c->stream));
   KB2E_CUDA(c, cudaMalloc(&c->flag, (size_t)c->nE + 2 * (size_t)c->nR));
   KB2E_CUDA(c, cudaMemsetAsync(c->flag, 0, (size_t)c->nE + 2 * (size_t)c->nR, c->stream));
   if (c->cfg.model != KB2E_MODEL_TRANSE) {
      c->w_row = c->cfg.model == KB2E_MODEL_TRANSH ? (size_t)c->P : (size_t)c->D * c->P;
      KB2E_CUDA(c, cudaMalloc(&c->w, (size_t)c->nR * c->w_row * sizeof(float)));
      KB2E_CUDA(c, cudaMalloc(&c->dw, (size_t)c->nR * c->w_row * sizeof(float)));
      KB2E_CUDA(c, cudaMemsetAsync(c->w, 0, (size_t)c->nR * c->w_row * sizeof(float), c->stream));
      KB2E_CUDA(c, cudaMemsetAsync(c->dw, 0, (size_t)c->nR * c->w_row * sizeof(float), c->stream));
      KB2E_CUDA(c, cudaMalloc(&c->rmin, (size_t)c->nE * sizeof(int)));
      KB2E_CUDA(c, cudaMalloc(&c->rmax, (size_t)c->nE * sizeof(int)));
      fill_int_kernel<<<blocks_for(c->nE, 256), 256, 0, c->stream>>>(c->rmin, c->nE, 0x7fffffff);
      fill_int_kernel<<<blocks_for(c->nE, 256), 256, 0, c->stream>>>(c->rmax, c->nE, -1);
   }
   KB2E_CUDA(c, cudaMalloc(&c->barrier, 64));
   KB2E_CUDA(c, cudaMalloc(&c->counters, 8 * sizeof(unsigned long long)));
   KB2E_CUDA(c, cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
   KB2E_CUDA(c, cudaMalloc(&c->pr, (size_t)c->nR * sizeof(double)));
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

void train_free(kb2e_ctx* c) {
   cudaFree(c->tab); cudaFree(c->dtab); cudaFree(c->w); cudaFree(c->dw); cudaFree(c->flag);
   cudaFree(c->rmin); cudaFree(c->rmax); cudaFree(c->triples); cudaFree(c->hash); cudaFree(c->pr);
   cudaFree(c->barrier); cudaFree(c->loss_dev); cudaFree(c->counters); cudaFree(c->pairs_dev);
   cudaFree(c->ent64); cudaFree(c->rel64); cudaFree(c->w64);
}

int train_set_triples(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n) {
   int rc = train_alloc(c);
   if (rc) return rc;
   for (int64_t i = 0; i < n; i++) {
      if (h[i] < 0 || h[i] >= c->nE || t[i] < 0 || t[i] >= c->nE || r[i] < 0 || r[i] >= c->nR)
         return fail(c, KB2E_ERR_ARG, "train triple " + std::to_string(i) + " has an id out of range");
   }
   cudaFree(c->triples); c->triples = nullptr;
   cudaFree(c->hash); c->hash = nullptr;
   c->n_train = n;
   if (n == 0) return KB2E_OK;
   int32_t* tmp = nullptr;
   KB2E_CUDA(c, cudaMalloc(&tmp, 3 * (size_t)n * sizeof(int32_t)));
   KB2E_CUDA(c, cudaMemcpyAsync(tmp, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(tmp + n, t, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(tmp + 2 * n, r, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMalloc(&c->triples, (size_t)n * sizeof(int4)));
   pack_triples_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(tmp, tmp + n, tmp + 2 * n, n, c->triples);
   uint64_t slots = 1024;
   while (slots < 2 * (uint64_t)n) slots <<= 1;
   c->hash_mask = slots - 1;
   KB2E_CUDA(c, cudaMalloc(&c->hash, slots * sizeof(uint64_t)));
   KB2E_CUDA(c, cudaMemsetAsync(c->hash, 0xff, slots * sizeof(uint64_t), c->stream));
   hash_insert_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->triples, n, c->hash, c->hash_mask);
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   cudaFree(tmp);
   return KB2E_OK;
}

int train_init_embeddings(kb2e_ctx* c) {
   int rc = train_alloc(c);
   if (rc) return rc;
   uint32_t k0 = (uint32_t)c->cfg.seed, k1 = (uint32_t)(c->cfg.seed >> 32);
   long long rows = (long long)c->nE + c->nR;
   init_rows_kernel<<<blocks_for(rows * 32, 256), 256, 0, c->stream>>>(c->tab, rows, c->D, c->P, k0, k1, 1u, 0);
   if (c->cfg.model == KB2E_MODEL_TRANSH) {
      init_rows_kernel<<<blocks_for((long long)c->nR * 32, 256), 256, 0, c->stream>>>(c->w, c->nR, c->D, c->P, k0, k1, 2u, 1);
   } else if (c->cfg.model == KB2E_MODEL_TRANSR) {
      identity_kernel<<<blocks_for((long long)c->nR * c->D * c->P, 256), 256, 0, c->stream>>>(c->w, c->nR, c->D, c->P);
   }
   KB2E_CUDA(c, cudaGetLastError());
   c->have32 = true;
   c->have64 = false;
   return KB2E_OK;
}

int tables_32_to_64(kb2e_ctx* c) {
   if (c->have64) return KB2E_OK;
   if (!c->have32) return fail(c, KB2E_ERR_ARG, "no embeddings: call kb2e_init_embeddings or kb2e_upload first");
   if (!c->ent64) {
      KB2E_CUDA(c, cudaMalloc(&c->ent64, (size_t)c->nE * c->D * sizeof(double)));
      KB2E_CUDA(c, cudaMalloc(&c->rel64, (size_t)c->nR * c->D * sizeof(double)));
      if (c->cfg.model != KB2E_MODEL_TRANSE) {
         size_t per = c->cfg.model == KB2E_MODEL_TRANSH ? (size_t)c->D : (size_t)c->D * c->D;
         KB2E_CUDA(c, cudaMalloc(&c->w64, (size_t)c->nR * per * sizeof(double)));
      }
   }
   widen_kernel<<<blocks_for((long long)c->nE * c->D, 256), 256, 0, c->stream>>>(c->tab, c->ent64, c->nE, c->D, c->P);
   widen_kernel<<<blocks_for((long long)c->nR * c->D, 256), 256, 0, c->stream>>>(c->tab + (size_t)c->nE * c->P, c->rel64, c->nR, c->D, c->P);
   if (c->cfg.model == KB2E_MODEL_TRANSH) {
      widen_kernel<<<blocks_for((long long)c->nR * c->D, 256), 256, 0, c->stream>>>(c->w, c->w64, c->nR, c->D, c->P);
   } else if (c->cfg.model == KB2E_MODEL_TRANSR) {
      widen_kernel<<<blocks_for((long long)c->nR * c->D * c->D, 256), 256, 0, c->stream>>>(c->w, c->w64, (long long)c->nR * c->D, c->D, c->P);
   }
   KB2E_CUDA(c, cudaGetLastError());
   c->have64 = true;
   return KB2E_OK;
}

int tables_64_to_32(kb2e_ctx* c) {
   int rc = train_alloc(c);
   if (rc) return rc;
   narrow_kernel<<<blocks_for((long long)c->nE * c->P, 256), 256, 0, c->stream>>>(c->ent64, c->tab, c->nE, c->D, c->P);
   narrow_kernel<<<blocks_for((long long)c->nR * c->P, 256), 256, 0, c->stream>>>(c->rel64, c->tab + (size_t)c->nE * c->P, c->nR, c->D, c->P);
   if (c->cfg.model == KB2E_MODEL_TRANSH) {
      narrow_kernel<<<blocks_for((long long)c->nR * c->P, 256), 256, 0, c->stream>>>(c->w64, c->w, c->nR, c->D, c->P);
   } else if (c->cfg.model == KB2E_MODEL_TRANSR) {
      narrow_kernel<<<blocks_for((long long)c->nR * c->D * c->P, 256), 256, 0, c->stream>>>(c->w64, c->w, (long long)c->nR * c->D, c->D, c->P);
   }
   KB2E_CUDA(c, cudaGetLastError());
   c->have32 = true;
   return KB2E_OK;
}

static void fill_args(kb2e_ctx* c, TrainArgs& a) {
   memset(&a, 0, sizeof(a));
   a.tab = c->tab; a.dtab = c->dtab; a.w = c->w; a.dw = c->dw;
   a.flagE = c->flag; a.flagR = c->flag + c->nE;
   a.rmin = c->rmin; a.rmax = c->rmax;
   a.triples = c->triples; a.hash = c->hash; a.hash_mask = c->hash_mask; a.pr = c->pr;
   a.barrier = c->barrier; a.loss = c->loss_dev; a.counters = c->counters;
   a.n_train = c->n_train;
   a.w_row = c->w_row;
   a.nE = c->nE; a.nR = c->nR; a.D = c->D; a.P = c->P;
   a.batches = c->cfg.batches;
   a.distance = c->cfg.distance;
   a.lr = (float)c->cfg.rate;
   a.margin = (float)c->cfg.margin;
   a.seed_lo = (uint32_t)c->cfg.seed;
   a.seed_hi = (uint32_t)(c->cfg.seed >> 32);
   a.flags = c->cfg.flags;
}

typedef void (*TrainKernel)(const TrainArgs);

template <int MODEL>
static TrainKernel pick_kernel(int lps, int nv) {
#define KB2E_PICK(L, N) if (lps == L && nv == N) return train_kernel<MODEL, L, N>;
   KB2E_PICK(8, 1) KB2E_PICK(8, 2) KB2E_PICK(8, 4)
   KB2E_PICK(16, 1) KB2E_PICK(16, 2) KB2E_PICK(16, 4)
   KB2E_PICK(32, 1) KB2E_PICK(32, 2) KB2E_PICK(32, 4)
#undef KB2E_PICK
   return nullptr;
}

// Choose lanes-per-sample (LPS) / float4-vectors-per-lane (NV): among the shapes that waste the
// fewest lanes on padding, take the widest group whose group count still covers the whole batch in
// one pass over the resident groups; KB2E_TRAIN_LPS overrides (tuning aid).
static void choose_shape(const kb2e_ctx* c, long long batchsize, int& lps, int& nv) {
   const int vecs = (c->P + 3) / 4;
   const long long threads = (long long)c->num_sms * kTrainThreads;
   const int maxnv = (c->cfg.model == KB2E_MODEL_TRANSH) ? 2 : 4;
   const char* env = getenv("KB2E_TRAIN_LPS");
   const int forced = env ? atoi(env) : 0;
   int Ls[3] = {32, 16, 8}, Ns[3];
   double eff[3], best = 0.0;
   for (int i = 0; i < 3; i++) {
      int N = (vecs + Ls[i] - 1) / Ls[i];
      Ns[i] = N <= 1 ? 1 : (N <= 2 ? 2 : (N <= 4 ? 4 : 99));
      eff[i] = (Ns[i] <= maxnv || Ls[i] == 32) ? (double)vecs / (Ls[i] * Ns[i]) : 0.0;
      if (Ns[i] > 4) eff[i] = 0.0;
      best = std::max(best, eff[i]);
   }
   lps = 32; nv = Ns[0];
   if (best == 0.0) { nv = 99; return; }
   int widest = -1, fit = -1;
   for (int i = 0; i < 3; i++) {
      if (forced == Ls[i] && eff[i] > 0.0) { lps = Ls[i]; nv = Ns[i]; return; }
      if (eff[i] < 0.75 * best) continue;
      if (widest < 0) widest = i;
      if (fit < 0 && threads / Ls[i] >= batchsize) fit = i;
   }
   int pick = fit >= 0 ? fit : widest;
   lps = Ls[pick]; nv = Ns[pick];
}

int train_run(kb2e_ctx* c, int first_epoch, int n_epochs, const int32_t* pairs_dev, int64_t n_pairs, double* loss_out) {
   if (c->cfg.model == KB2E_MODEL_TRANSR) return fail(c, KB2E_ERR_LIMIT, "TransR training kernel not built yet");
   if (!c->have32) return fail(c, KB2E_ERR_ARG, "no embeddings: call kb2e_init_embeddings or kb2e_upload first");
   TrainArgs a;
   if (n_epochs <= 0) return KB2E_OK;
   if (n_epochs > c->loss_cap) {
      cudaFree(c->loss_dev);
      KB2E_CUDA(c, cudaMalloc(&c->loss_dev, (size_t)n_epochs * sizeof(double)));
      c->loss_cap = n_epochs;
   }
   fill_args(c, a);
   a.first_epoch = first_epoch;
   a.n_epochs = n_epochs;
   if (pairs_dev) {
      a.pairs = pairs_dev;
      a.batches = 1;
      a.batchsize = n_pairs;
   } else {
      if (!c->triples || c->n_train == 0) return fail(c, KB2E_ERR_ARG, "no training triples: call kb2e_set_train_triples first");
      if (!c->have_pr) return fail(c, KB2E_ERR_ARG, "no corruption probabilities: call kb2e_set_bern first");
      if (c->cfg.batches <= 0) return fail(c, KB2E_ERR_ARG, "batches must be positive");
      a.batchsize = c->n_train / c->cfg.batches;  // common/trainer.cpp:70
   }
   if ((int64_t)a.batchsize >= (1ll << 32)) return fail(c, KB2E_ERR_LIMIT, "batch larger than 2^32 samples");
   int lps, nv;
   choose_shape(c, a.batchsize, lps, nv);
   if (nv > 4) return fail(c, KB2E_ERR_LIMIT, "embedding size above 512 is not supported by the training kernel");
   TrainKernel k = c->cfg.model == KB2E_MODEL_TRANSE ? pick_kernel<KB2E_MODEL_TRANSE>(lps, nv) : pick_kernel<KB2E_MODEL_TRANSH>(lps, nv);
   if (!k) return fail(c, KB2E_ERR_LIMIT, "no training kernel for this embedding size");
   int per_sm = 0;
   KB2E_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kTrainThreads, 0));
   if (per_sm < 1) return fail(c, KB2E_ERR_CUDA, "training kernel does not fit on an SM");
   KB2E_CUDA(c, cudaMemsetAsync(c->barrier, 0, 64, c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->loss_dev, 0, (size_t)n_epochs * sizeof(double), c->stream));
   void* params[] = {&a};
   KB2E_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   KB2E_CUDA(c, cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(kTrainThreads), params, 0, c->stream));
   KB2E_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   std::vector<double> loss(n_epochs);
   unsigned long long cnt[3];
   KB2E_CUDA(c, cudaMemcpyAsync(loss.data(), c->loss_dev, (size_t)n_epochs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(cnt, c->counters, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   c->tstats.kernel_ms += ms;
   c->tstats.launches += 1;
   c->tstats.samples += (uint64_t)a.batchsize * (uint64_t)a.batches * (uint64_t)n_epochs;
   c->tstats.active = cnt[0];
   c->tstats.touched_ent = cnt[1];
   c->tstats.touched_rel = cnt[2];
   if (loss_out) memcpy(loss_out, loss.data(), (size_t)n_epochs * sizeof(double));
   c->have64 = false;
   return KB2E_OK;
}

int train_sample(kb2e_ctx* c, int epoch, int batch, int64_t count, int32_t* out_dev) {
   if (!c->triples || !c->have_pr) return fail(c, KB2E_ERR_ARG, "sampler needs kb2e_set_train_triples and kb2e_set_bern");
   TrainArgs a;
   fill_args(c, a);
   uint32_t gb = (uint32_t)epoch * (uint32_t)c->cfg.batches + (uint32_t)batch;
   sample_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(a, gb, count, out_dev);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int train_score32(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n, double* out) {
   if (!c->have32) {
      if (!c->have64) return fail(c, KB2E_ERR_ARG, "no embeddings");
      int rc = tables_64_to_32(c);
      if (rc) return rc;
   }
   if (c->cfg.model == KB2E_MODEL_TRANSR) return fail(c, KB2E_ERR_LIMIT, "fp32 TransR scoring not built yet");
   unsigned blocks = blocks_for(n * 32, 256);
   if (c->cfg.model == KB2E_MODEL_TRANSE)
      score32_kernel<KB2E_MODEL_TRANSE><<<blocks, 256, 0, c->stream>>>(c->tab, c->w, c->nE, c->D, c->P, c->cfg.distance, h, t, r, n, out);
   else
      score32_kernel<KB2E_MODEL_TRANSH><<<blocks, 256, 0, c->stream>>>(c->tab, c->w, c->nE, c->D, c->P, c->cfg.distance, h, t, r, n, out);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

}  // namespace kb2e
