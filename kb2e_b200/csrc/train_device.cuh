// Device-side building blocks of the persistent training kernels (train.cu, train_dist.cu):
// row-vector helpers, normalisations (common/utils.cpp:70-111), the grid-wide barrier and the
// counter-RNG sampler (common/trainer.cpp:78-98).
#pragma once

#include "common.cuh"
#include "internal.h"

namespace kb2e {

constexpr int kMaxTrainThreads = 1024;
constexpr int kTraceSlots = 64;

struct TrainArgs {
   float* tab;
   float* dtab;
   float* w;
   float* dw;
   uint32_t* flag;   // [nE + nR] stamp (global batch + 1) of the last batch that touched the row; 0 = never
   uint32_t* cflag;  // [nR] LIST kernels, TransH: stamp of the batch the relation row was last marked for by an entity-side constraint step
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
   uint32_t stamp_base;           // batches this context has run before this launch: batch k of the launch stamps rows with stamp_base + k + 1
   int replicas;                  // batched training (train_sweep_kernel): K models stacked in tab / dtab / flag, K x (nE + nR) rows
   const struct RepParams* rep;   // [replicas] per-model learning rate, margin, seed
   int tasks_per_group;           // sweep: (model, sample) tasks one group handles per batch (<= lanes per group)
   float* relbuf1;                // train_transh_sr_kernel: second buffer of the relation-side deltas, [2][nR][P]
   int phase1_only;               // test hook (kb2e_train_batch_deltas): stop after the accumulation phase
   int cap_ent, cap_rel;          // LIST kernels: capacities of the per-CTA touched-row lists (train.cu)
   unsigned long long* trace;     // tuning aid (KB2E_TRAIN_TRACE): per-CTA clock stamps of the first batches
};

struct RepParams {
   float lr, margin;
   uint32_t seed_lo, seed_hi;
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
__device__ __forceinline__ void load_row_pinned(const float* base, int P, int gl, float4 (&v)[NV]) {
#pragma unroll
   for (int q = 0; q < NV; q++) {
      int off = (q * LPS + gl) * 4;
      v[q] = off < P ? ld_cg4_pinned(base + off) : f4(0.f);
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
      // one IEEE reciprocal, then multiplications: eight IEEE divisions by the same number compile to eight serialised
      // MUFU.RCP + FCHK + 5-FFMA sequences (each its own reconvergence region), measured at ~1 us per row
      // (tools/trace_transh_sr_fine.py) on the critical path of every publish; <= 1 ulp from the divided value
      const float inv = 1.f / len;
      for (int q = 0; q < NV; q++) v[q] = make_float4(v[q].x * inv, v[q].y * inv, v[q].z * inv, v[q].w * inv);
   }
}

// The loop of common::norm(a, b, rate), common/utils.cpp:83-108 (`sum` is deliberately not reset
// between iterations, as in the reference).  b must already be unit length.  Returns #corrective steps.
//
// Every pass of the reference's loop rescales b and mixes a and b linearly (b /= sum; a -= rate b; b -= rate a),
// so a and b stay in the plane of the initial vectors: a = a1 a0 + a2 b0, b = b1 a0 + b2 b0.  The loop therefore
// runs on the three scalars A = |a|^2, B = |b|^2, X = a.b (one fused group reduction up front) and the four
// coefficients -- no per-iteration shuffle reductions on the dependent path (tens of iterations are common when
// d_r . w_r starts far above the 0.1 threshold) -- and the vectors are rebuilt once at the end.
template <int LPS, int NV>
__device__ __forceinline__ int soft_orth_loop(float4 (&a)[NV], float4 (&b)[NV], float rate, uint32_t gmask) {
   float A = 0.f, B = 0.f, X = 0.f;
#pragma unroll
   for (int q = 0; q < NV; q++) {
      A += dot4(a[q], a[q]);
      B += dot4(b[q], b[q]);
      X += dot4(a[q], b[q]);
   }
#pragma unroll
   for (int o = LPS / 2; o > 0; o >>= 1) {
      A += __shfl_xor_sync(gmask, A, o);
      B += __shfl_xor_sync(gmask, B, o);
      X += __shfl_xor_sync(gmask, X, o);
   }
   float a1 = 1.f, a2 = 0.f, b1 = 0.f, b2 = 1.f, sum = 0.f;
   int iters = 0;
   while (true) {
      // sum = sqrt(sum + B), b /= sum: one approximate reciprocal square root (2 ulp) instead of an IEEE square root and an
      // IEEE division on the loop-carried path -- the chain per iteration is ~45 instead of ~110 cycles, and tens of
      // iterations are common.  inv scales b, B and X consistently, so its rounding is a slightly different normalisation of
      // b (renormalised after the loop), not an inconsistency.
      const float s2 = sum + B;
      const float inv = rsqrtf(s2);
      sum = s2 * inv;
      b1 *= inv; b2 *= inv;            // b /= sum
      B *= inv * inv;
      X *= inv;                        // x = a . b
      if (X > 0.1f && iters < 1000) {
         a1 -= rate * b1; a2 -= rate * b2;                           // a -= rate * b
         const float A1 = A - 2.f * rate * X + rate * rate * B;
         const float X1 = X - rate * B;                              // a_new . b
         b1 -= rate * a1; b2 -= rate * a2;                           // b -= rate * a_new
         const float B1 = B - 2.f * rate * X1 + rate * rate * A1;
         X = X1 - rate * A1;
         A = A1; B = B1;
         iters++;
      } else {
         break;
      }
   }
#pragma unroll
   for (int q = 0; q < NV; q++) {
      const float4 a0 = a[q], b0 = b[q];
      if (iters > 0) a[q] = a1 * a0 + a2 * b0;
      b[q] = b1 * a0 + b2 * b0;
   }
   return iters;
}

// ---- grid-wide barrier (all CTAs are co-resident: cooperative launch) ---------------------------
// bar.sync orders the CTA's writes before thread 0's gpu-scope release; pollers use relaxed loads and
// one acquire fence at the end.  Measured on B200 (tools/microbench.cu): ~1.3 us per barrier.
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
   uint32_t v;
   asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
   return v;
}

// Split in two so that work which does not depend on other CTAs can run between arrive and wait.
__device__ __forceinline__ void grid_arrive(uint32_t* counter, uint32_t& target) {
   __syncthreads();
   if (threadIdx.x == 0) {
      target += gridDim.x;
      red_release_add_u32(counter, 1u);
   }
}

__device__ __forceinline__ void grid_wait(const uint32_t* counter, uint32_t target) {
   if (threadIdx.x == 0) {
      while ((int32_t)(ld_relaxed_u32(counter) - target) < 0) {
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
   }
   __syncthreads();
}

__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& target) {
   grid_arrive(counter, target);
   grid_wait(counter, target);
}

// ---- the sampler: common/trainer.cpp:78-98 with a counter RNG -----------------------------------
struct Pair {
   int h, t, r, c;    // positive triple and the corrupting entity
   bool corruptTail;  // true: (h, r, c) is the negative; false: (c, r, t)
};

// KB2E_FLAG_SAMPLER_RANDMAX: common/utils.cpp:113-120 with two uniform 31-bit draws in place of the two std::rand()
// calls -- (rand() * rand()) % x in wrapping 32-bit int arithmetic, then "while (res < 0) res += x" -- i.e. the
// reference's (non-uniform) index distribution, driven by the counter RNG (the CPU twin in the test checker bears the same name).
__device__ __forceinline__ int randmax_from(uint32_t x0, uint32_t x1, int x) {
   int res = (int)((x0 >> 1) * (x1 >> 1));
   res = res % x;
   if (res < 0) res += x;
   return res;
}

__device__ __forceinline__ Pair draw_pair(const TrainArgs& a, uint32_t k, uint32_t gb, uint32_t seed_lo, uint32_t seed_hi) {
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
   philox4x32(k, gb, 0u, 0u, seed_lo, seed_hi, x);
   const bool randmax = (a.flags & KB2E_FLAG_SAMPLER_RANDMAX) != 0u;
   uint64_t i = mulhi64(((uint64_t)x[0] << 32) | x[1], (uint64_t)a.n_train);
   int j = (int)mulhi32(x[3], (uint32_t)a.nE);
   if (randmax) {
      uint32_t y[4];
      philox4x32(k, gb, 0u, 1u, seed_lo, seed_hi, y);
      i = (uint64_t)randmax_from(x[0], x[1], (int)a.n_train);
      j = randmax_from(y[0], y[1], a.nE);
   }
   int4 tr = __ldg(a.triples + i);
   s.h = tr.x; s.t = tr.y; s.r = tr.z;
   int coin = (int)(x[2] % 1000u);
   s.corruptTail = (double)coin < __ldg(a.pr + tr.z);
   for (uint32_t att = 1; att < 64; att++) {
      uint64_t key = s.corruptTail ? pack_triple(s.h, s.r, j) : pack_triple(j, s.r, s.t);
      if (!hash_contains(a.hash, a.hash_mask, key)) break;
      philox4x32(k, gb, att, 0u, seed_lo, seed_hi, x);
      j = randmax ? randmax_from(x[0], x[1], a.nE) : (int)mulhi32(x[0], (uint32_t)a.nE);
   }
   s.c = j;
   return s;
}


__device__ __forceinline__ Pair draw_pair(const TrainArgs& a, uint32_t k, uint32_t gb) { return draw_pair(a, k, gb, a.seed_lo, a.seed_hi); }

// ---- the same sampler in three stages, for software pipelining across the phases of a batch -------------------------
// draw_pair is a chain of dependent loads (triple -> corruption side -> hash probe): 2-3 L2 round trips.  Run in one piece
// between "arrive" and "wait" of the end-of-batch barrier it ends up on the critical path of whichever CTA arrives last.
// Split in three, every load is issued one phase before its value is needed:
//   draw_begin    counter RNG -> triple index, candidate entity; ISSUES the triple fetch.  triples[i].w holds the
//                 relation's corruption threshold ceil(pr[r]) (coin < pr  <=>  coin < ceil(pr) for an integer coin), so
//                 the corruption side needs no second dependent load
//   draw_probe    corruption side -> key of the candidate negative; ISSUES the first two slots of its probe sequence
//   draw_finish   decides from the two slots (empty / match); only an unresolved sequence or a hit (candidate is a known
//                 triple: resample, common/trainer.cpp:89-96) falls back to the loop of draw_pair.  Same samples, bit for bit.
struct DrawStage {
   int4 tr;
   uint64_t v0, v1;
   int j;      // candidate entity
   int coin;   // 0..999, compared with the relation's corruption threshold
};

__device__ __forceinline__ void draw_begin(const TrainArgs& a, uint32_t k, uint32_t gb, DrawStage& d, uint32_t seed_lo, uint32_t seed_hi) {
   if (a.pairs != nullptr) return;
   uint32_t x[4];
   philox4x32(k, gb, 0u, 0u, seed_lo, seed_hi, x);
   uint64_t i = mulhi64(((uint64_t)x[0] << 32) | x[1], (uint64_t)a.n_train);
   int j = (int)mulhi32(x[3], (uint32_t)a.nE);
   if (a.flags & KB2E_FLAG_SAMPLER_RANDMAX) {
      uint32_t y[4];
      philox4x32(k, gb, 0u, 1u, seed_lo, seed_hi, y);
      i = (uint64_t)randmax_from(x[0], x[1], (int)a.n_train);
      j = randmax_from(y[0], y[1], a.nE);
   }
   d.j = j;
   d.coin = (int)(x[2] % 1000u);
   d.tr = ld_nc_int4_pinned(a.triples + i);
}

__device__ __forceinline__ void draw_begin(const TrainArgs& a, uint32_t k, uint32_t gb, DrawStage& d) {
   draw_begin(a, k, gb, d, a.seed_lo, a.seed_hi);
}

__device__ __forceinline__ uint64_t draw_key(const DrawStage& d, bool& corruptTail) {
   corruptTail = d.coin < d.tr.w;
   return corruptTail ? pack_triple(d.tr.x, d.tr.z, d.j) : pack_triple(d.j, d.tr.z, d.tr.y);
}

__device__ __forceinline__ void draw_probe(const TrainArgs& a, DrawStage& d) {
   if (a.pairs != nullptr) return;
   bool ct;
   const uint64_t slot = mix64(draw_key(d, ct)) & a.hash_mask;
   d.v0 = ld_nc_u64_pinned(a.hash + slot);
   d.v1 = ld_nc_u64_pinned(a.hash + ((slot + 1) & a.hash_mask));
}

__device__ __forceinline__ Pair draw_finish(const TrainArgs& a, uint32_t k, uint32_t gb, const DrawStage& d, uint32_t seed_lo, uint32_t seed_hi) {
   if (a.pairs != nullptr) return draw_pair(a, k, gb);
   Pair s;
   s.h = d.tr.x; s.t = d.tr.y; s.r = d.tr.z;
   const uint64_t key = draw_key(d, s.corruptTail);
   int j = d.j;
   bool hit;
   if (d.v0 == key || (d.v0 != kEmptyKey && d.v1 == key)) hit = true;
   else if (d.v0 == kEmptyKey || d.v1 == kEmptyKey) hit = false;
   else {   // two occupied slots of other keys: walk on
      uint64_t slot = (mix64(key) + 2) & a.hash_mask;
      while (true) {
         const uint64_t v = __ldg(a.hash + slot);
         if (v == key) { hit = true; break; }
         if (v == kEmptyKey) { hit = false; break; }
         slot = (slot + 1) & a.hash_mask;
      }
   }
   if (hit) {
      for (uint32_t att = 1; att < 64; att++) {
         uint32_t x[4];
         philox4x32(k, gb, att, 0u, seed_lo, seed_hi, x);
         j = (a.flags & KB2E_FLAG_SAMPLER_RANDMAX) ? randmax_from(x[0], x[1], a.nE) : (int)mulhi32(x[0], (uint32_t)a.nE);
         const uint64_t key2 = s.corruptTail ? pack_triple(s.h, s.r, j) : pack_triple(j, s.r, s.t);
         if (!hash_contains(a.hash, a.hash_mask, key2)) break;
      }
   }
   s.c = j;
   return s;
}

__device__ __forceinline__ Pair draw_finish(const TrainArgs& a, uint32_t k, uint32_t gb, const DrawStage& d) {
   return draw_finish(a, k, gb, d, a.seed_lo, a.seed_hi);
}

// ---- phase-2 row walk ------------------------------------------------------------------------------
// Every group owns a contiguous range of rows.  The stamps of LPS rows are fetched with one coalesced
// load per lane and turned into a bit mask by a ballot, then the stamped rows are handed to `body` two
// at a time (r1 = -1 when only one is left) so that the loads of both rows are in flight together and
// no row waits on its own stamp load.  `stamped(row)` decides whether a row was touched in this batch.
template <int LPS, typename Pred, typename Body>
__device__ __forceinline__ void for_stamped_rows(long long first, long long end, int gl, uint32_t gmask, int lane,
                                                 Pred&& stamped, Body&& body) {
   const int gshift = (lane / LPS) * LPS;
   for (long long base = first; base < end; base += LPS) {
      const long long r = base + gl;
      const bool f = r < end && stamped(r);
      uint32_t m = (__ballot_sync(gmask, f) >> gshift) & (LPS == 32 ? 0xffffffffu : ((1u << LPS) - 1u));
      while (m) {
         const int b0 = __ffs(m) - 1;
         m &= m - 1;
         int b1 = -1;
         if (m) { b1 = __ffs(m) - 1; m &= m - 1; }
         body(base + b0, b1 >= 0 ? base + b1 : -1ll);
      }
   }
}

// this group's contiguous share of [row_begin, row_end) when the rows are dealt to G groups
__device__ __forceinline__ void group_range(long long row_begin, long long row_end, long long g0, long long G,
                                            long long& first, long long& end) {
   const long long per = (row_end - row_begin + G - 1) / G;
   first = row_begin + g0 * per;
   end = first + per < row_end ? first + per : row_end;
   if (first > row_end) first = row_end;
}

}  // namespace kb2e
