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
// Touched rows, two ways (template parameter LIST):
//   scan  every group walks its share of the stamp array in phase 2 (batches that touch most of the table);
//   list  every active sample appends its four rows to a list in its CTA's shared memory (no global traffic); in
//         phase 2 a CTA walks its own list, one row per group: the group claims the row with an atomic exchange of the
//         row's stamp -- issued together with the loads of the row and its delta, so the claim costs no extra round
//         trip -- and publishes it unless another CTA's copy of the same row claimed it first.  No stamp scan, no
//         dependence on where in the row space the touched rows fall, and the lists are balanced by construction
//         (every CTA holds the same number of samples).  At FB15k shape: ~10 active samples = ~40 list entries per CTA
//         against 40 groups (640 threads), so the publish is a single pass.
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
#include "train_device.cuh"

namespace kb2e {

// ---- per-CTA lists of touched rows (LIST kernels) --------------------------------------------------
// Shared-memory layout (ints): [0] #entity-side rows, [1], [2] #relation-side rows of even / odd stamps, [3] unused,
// then ent[cap_ent], rel[0][cap_rel], rel[1][cap_rel].  TransE keeps entity and relation rows in the one `ent` list
// (they are published in the same pass); TransH publishes relation-side rows first, and the soft constraint of an
// entity row may touch a relation row for the NEXT batch, hence the two relation lists indexed by stamp parity.
// Entries are row indices in the unified row space (relation r = nE + r).
struct RowLists {
   int* count;
   int* ent;
   int* rel[2];
   int cap_ent, cap_rel;
   // train_transh_sr_kernel only (nullptr elsewhere, and the branches on it fold away): the relation-side rows live in
   // shared memory for the whole launch, [nR][2][P] = (d_r, w_r), and the relation-side deltas ping-pong between two buffers
   float* srel;
   float* dr_cur;    // [nR][P] delta of d_r, this batch's RED target
   float* dw_cur;    // [nR][P] delta of w_r, this batch's RED target
   float* dw_next;   // [nR][P] delta of w_r the NEXT batch consumes (entity-side constraint carries)
};

template <int LPS, int NV>
__device__ __forceinline__ void load_row_shared(const float* base, int P, int gl, float4 (&v)[NV]) {
#pragma unroll
   for (int q = 0; q < NV; q++) {
      const int off = (q * LPS + gl) * 4;
      v[q] = off < P ? *reinterpret_cast<const float4*>(base + off) : f4(0.f);
   }
}

__device__ __forceinline__ void list_push(int* count, int* list, int cap, int row, unsigned long long* counters) {
   const int slot = atomicAdd(count, 1);
   if (slot < cap) list[slot] = row;
   else counters[6] = 1ull;   // cannot happen with the host's capacities; reported as an error if it ever does
}

// ---- accumulation into the delta tables ---------------------------------------------------------------
// Default: one 16-byte floating-point vector RED per float4 (order of the additions = order of arrival: results agree
// from run to run only to rounding).  KB2E_FLAG_DETERMINISTIC: the same updates as 32-bit FIXED-POINT integers (2^-24
// units, |sum| < 128 per row and batch) added with integer REDs -- integer addition is associative, so the sums, and with
// them every table and every printed loss, are bit-identical from run to run, as the reference's are for a given -seed.
constexpr float kDetScale = 16777216.0f;   // 2^24
constexpr double kDetLossScale = 1048576.0;   // 2^20: per-epoch loss as a 64-bit fixed-point sum

// DET is a template parameter, not a test of the flag at run time: the run-time branches cost the scan kernel of the
// scaled shape 100 bytes of extra spills in its hot loop (332 -> 504 ms per epoch, measured).
template <int LPS, int NV, bool DET>
__device__ __forceinline__ void acc_row(float* base, int P, int gl, const float4 (&v)[NV]) {
   if (!DET) {
      red_row<LPS, NV>(base, P, gl, v);
      return;
   }
#pragma unroll
   for (int q = 0; q < NV; q++) {
      const int off = (q * LPS + gl) * 4;
      if (off < P) {
         int* p = reinterpret_cast<int*>(base + off);
         atomicAdd(p + 0, __float2int_rn(v[q].x * kDetScale));
         atomicAdd(p + 1, __float2int_rn(v[q].y * kDetScale));
         atomicAdd(p + 2, __float2int_rn(v[q].z * kDetScale));
         atomicAdd(p + 3, __float2int_rn(v[q].w * kDetScale));
      }
   }
}

__device__ __forceinline__ float4 det_to_float(float4 bits) {
   const float inv = 1.0f / kDetScale;
   return make_float4((float)__float_as_int(bits.x) * inv, (float)__float_as_int(bits.y) * inv, (float)__float_as_int(bits.z) * inv,
                      (float)__float_as_int(bits.w) * inv);
}

// ---- phase 1: one (positive, negative) pair, TransE and TransH -----------------------------------
// row_base: first row of the model the pair belongs to in the stacked tables (0 unless several models are trained in one
// launch, train_sweep_kernel); lr / margin: that model's learning rate and margin.
template <int MODEL, int LPS, int NV, bool LIST, bool DET>
__device__ __forceinline__ void process_pair_at(const TrainArgs& a, const RowLists& L, const Pair s, size_t row_base, float lr, float margin,
                                                int gl, uint32_t gmask, uint32_t stamp, double& loss_acc, uint32_t& active_acc) {
   const int P = a.P, D = a.D;
   const float* eh = a.tab + (row_base + s.h) * P;
   const float* et = a.tab + (row_base + s.t) * P;
   const float* ec = a.tab + (row_base + s.c) * P;
   const float* er = a.tab + (row_base + a.nE + s.r) * P;
   float4 vh[NV], vt[NV], vc[NV], vr[NV], vw[NV];
   load_row<LPS, NV>(eh, P, gl, vh);
   load_row<LPS, NV>(et, P, gl, vt);
   load_row<LPS, NV>(ec, P, gl, vc);
   const bool sr = MODEL == KB2E_MODEL_TRANSH && L.srel != nullptr;
   if (sr) load_row_shared<LPS, NV>(L.srel + (size_t)(2 * s.r) * P, P, gl, vr);
   else load_row<LPS, NV>(er, P, gl, vr);
   float hs = 0.f, ts = 0.f, cs = 0.f;
   if (MODEL == KB2E_MODEL_TRANSH) {
      if (sr) load_row_shared<LPS, NV>(L.srel + (size_t)(2 * s.r + 1) * P, P, gl, vw);
      else load_row<LPS, NV>(a.w + (size_t)s.r * P, P, gl, vw);
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
   if (!(ep + margin > en)) return;

   if (gl == 0) {
      loss_acc += (double)(margin + ep - en);
      active_acc++;
   }
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
   float* dh = a.dtab + (row_base + s.h) * P;
   float* dt = a.dtab + (row_base + s.t) * P;
   float* dc = a.dtab + (row_base + s.c) * P;
   float* dr = sr ? L.dr_cur + (size_t)s.r * P : a.dtab + (row_base + a.nE + s.r) * P;
   float4 u[NV];
   // relation row: -= m*lr*x  with m = -1 (positive), +1 (negative)
#pragma unroll
   for (int q = 0; q < NV; q++) u[q] = gp[q] - gn[q];
   acc_row<LPS, NV, DET>(dr, P, gl, u);
   if (s.corruptTail) {
      // negative = (h, r, c): head -= gn, c += gn; positive: head += gp, tail -= gp
      acc_row<LPS, NV, DET>(dh, P, gl, u);
#pragma unroll
      for (int q = 0; q < NV; q++) u[q] = -1.f * gp[q];
      acc_row<LPS, NV, DET>(dt, P, gl, u);
      acc_row<LPS, NV, DET>(dc, P, gl, gn);
   } else {
      // negative = (c, r, t): c -= gn, tail += gn
      acc_row<LPS, NV, DET>(dh, P, gl, gp);
#pragma unroll
      for (int q = 0; q < NV; q++) u[q] = gn[q] - gp[q];
      acc_row<LPS, NV, DET>(dt, P, gl, u);
#pragma unroll
      for (int q = 0; q < NV; q++) u[q] = -1.f * gn[q];
      acc_row<LPS, NV, DET>(dc, P, gl, u);
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
      acc_row<LPS, NV, DET>((sr ? L.dw_cur : a.dw) + (size_t)s.r * P, P, gl, u);
   }
   // flag the touched rows (+ relation range per entity for the TransH/TransR constraints)
   if (gl < 3) {
      int e = gl == 0 ? s.h : (gl == 1 ? s.t : s.c);
      if (!LIST) a.flag[e] = stamp;
      if (MODEL != KB2E_MODEL_TRANSE) {
         atomicMin(a.rmin + e, s.r);
         atomicMax(a.rmax + e, s.r);
      }
   } else if (gl == 3) {
      if (!LIST || sr) a.flag[(size_t)a.nE + s.r] = stamp;
   }
   if (LIST) {
      // the sample's rows go on this CTA's lists (duplicates are resolved when the rows are claimed in phase 2)
      const int n_ent = MODEL == KB2E_MODEL_TRANSE ? 4 : 3;
      int slot = 0;
      if (gl == 0) slot = atomicAdd(L.count + 0, n_ent);
      slot = __shfl_sync(gmask, slot, (threadIdx.x & 31) - gl);
      if (gl < n_ent) {
         const int row = (int)row_base + (gl == 0 ? s.h : (gl == 1 ? s.t : (gl == 2 ? s.c : a.nE + s.r)));
         if (slot + gl < L.cap_ent) L.ent[slot + gl] = row;
         else a.counters[6] = 1ull;   // cannot happen with the host's capacities; reported as an error if it ever does
      } else if (gl == 3 && !sr) {
         list_push(L.count + 1 + (stamp & 1u), L.rel[stamp & 1u], L.cap_rel, a.nE + s.r, a.counters);
      }
   }
}

template <int MODEL, int LPS, int NV, bool LIST, bool DET>
__device__ __forceinline__ void process_pair(const TrainArgs& a, const RowLists& L, const Pair s, int gl, uint32_t gmask, uint32_t stamp,
                                             double& loss_acc, uint32_t& active_acc) {
   process_pair_at<MODEL, LPS, NV, LIST, DET>(a, L, s, 0, a.lr, a.margin, gl, gmask, stamp, loss_acc, active_acc);
}

// ---- phase 2 -------------------------------------------------------------------------------------
// Common head of every publish: x = cur + delta, delta = 0 (rows arrive preloaded so that the loads of
// two rows are in flight together).
template <int LPS, int NV, bool DET>
__device__ __forceinline__ void apply_delta(float* del, int P, int gl, float4 (&x)[NV], float4 (&d)[NV]) {
#pragma unroll
   for (int q = 0; q < NV; q++) { x[q] = x[q] + (DET ? det_to_float(d[q]) : d[q]); d[q] = f4(0.f); }
   store_row<LPS, NV>(del, P, gl, d);
}

// Relation-side row r: d_r (and w_r).  transe/trainer.cpp:43, transh/trainer.cpp:48,52,56.
// w_r after its delta, against the finished d_r (x): transh/trainer.cpp:52-54
template <int LPS, int NV>
__device__ __forceinline__ int finish_hyperplane(float4 (&x)[NV], float4 (&b)[NV], float lr, uint32_t gmask) {
   norm_row<LPS, NV>(b, false, gmask);          // transh/trainer.cpp:52
   norm_row<LPS, NV>(b, false, gmask);          // common/utils.cpp:82
   const int steps = soft_orth_loop<LPS, NV>(x, b, lr, gmask);    // common/utils.cpp:83-108
   norm_row<LPS, NV>(b, false, gmask);          // common/utils.cpp:110
   return steps;
}

template <int MODEL, int LPS, int NV, bool DET>
__device__ __forceinline__ void finish_relation(const TrainArgs& a, int r, int gl, uint32_t gmask, float4 (&x)[NV], float4 (&d)[NV]) {
   const int P = a.P;
   float* cur = a.tab + ((size_t)a.nE + r) * P;
   apply_delta<LPS, NV, DET>(a.dtab + ((size_t)a.nE + r) * P, P, gl, x, d);
   norm_row<LPS, NV>(x, true, gmask);
   if (MODEL == KB2E_MODEL_TRANSH) {
      float* wc = a.w + (size_t)r * P;
      float* wd = a.dw + (size_t)r * P;
      float4 b[NV], db[NV];
      load_row<LPS, NV>(wc, P, gl, b);
      load_row<LPS, NV>(wd, P, gl, db);
      apply_delta<LPS, NV, DET>(wd, P, gl, b, db);
      finish_hyperplane<LPS, NV>(x, b, a.lr, gmask);
      store_row<LPS, NV>(wc, P, gl, b);
   }
   store_row<LPS, NV>(cur, P, gl, x);
}

// Entity row e.  transe/trainer.cpp:44-45, transh/trainer.cpp:49-50,57-58.
template <int MODEL, int LPS, int NV, bool LIST, bool DET>
__device__ __forceinline__ void finish_entity(const TrainArgs& a, const RowLists& L, int e, int gl, uint32_t gmask, uint32_t next_stamp,
                                              float4 (&x)[NV], float4 (&d)[NV], bool have_range = false, int pre_r0 = 0, int pre_r1 = -1) {
   const int P = a.P;
   float* cur = a.tab + (size_t)e * P;
   apply_delta<LPS, NV, DET>(a.dtab + (size_t)e * P, P, gl, x, d);
   norm_row<LPS, NV>(x, true, gmask);
   if (MODEL == KB2E_MODEL_TRANSH) {
      // (list publish: the range was requested together with the row, not after the claim came back)
      int r0 = have_range ? pre_r0 : __ldcg(a.rmin + e), r1 = have_range ? pre_r1 : __ldcg(a.rmax + e);
      if (gl == 0) { a.rmin[e] = 0x7fffffff; a.rmax[e] = -1; }
      const bool sr = L.srel != nullptr;
      // both hyperplanes are requested up front (global copies: one L2 round trip instead of one per pass)
      float4 w_hi[NV];
      if (!sr && r1 >= 0 && r1 != r0) load_row<LPS, NV>(a.w + (size_t)r1 * P, P, gl, w_hi);
      for (int pass = 0; pass < 2 && r1 >= 0; pass++) {
         int r = pass == 0 ? r0 : r1;
         if (pass == 1 && r1 == r0) break;
         float4 b[NV], b0[NV];
         if (sr) {
            load_row_shared<LPS, NV>(L.srel + (size_t)(2 * r + 1) * P, P, gl, b0);
         } else if (pass == 0) {
            load_row<LPS, NV>(a.w + (size_t)r * P, P, gl, b0);
         } else {
#pragma unroll
            for (int q = 0; q < NV; q++) b0[q] = w_hi[q];
         }
#pragma unroll
         for (int q = 0; q < NV; q++) b[q] = b0[q];
         int iters = soft_orth_loop<LPS, NV>(x, b, a.lr, gmask);
         if (iters > 0) {
            // The reference also perturbs w_r here (common/utils.cpp:103,110); the perturbation is
            // folded into the NEXT batch's delta so that w_r stays read-only in this phase.
            norm_row<LPS, NV>(b, false, gmask);
#pragma unroll
            for (int q = 0; q < NV; q++) b[q] = b[q] - b0[q];
            acc_row<LPS, NV, DET>((sr ? L.dw_next : a.dw) + (size_t)r * P, P, gl, b);
            if (gl == 0) {
               if (!LIST) {
                  a.flag[(size_t)a.nE + r] = next_stamp;
               } else if (sr) {
                  a.cflag[r] = next_stamp;
               } else {
                  // next batch's relation list of this CTA; cflag keeps the mark across the end of a launch
                  a.cflag[r] = next_stamp;
                  list_push(L.count + 1 + (next_stamp & 1u), L.rel[next_stamp & 1u], L.cap_rel, a.nE + r, a.counters);
               }
            }
         }
      }
   }
   store_row<LPS, NV>(cur, P, gl, x);
}


// ================================ TransR (fp32 scoring hook; training lives in train_transr.cu) ====
// One warp per triple, rows in the float4 layout (lane owns elements 4*lane .. 4*lane+3; D <= 128).
// M_r is [D][P] = M[j = input dim][i = output dim] (transr/trainer.h:31): row j is contiguous in i, so
// the projection M_r^T e streams the matrix row by row with coalesced 16-byte loads while e_j is
// broadcast by shuffle (transr/transr.cpp:20-25, work vectors zeroed -- SURVEY.md 8c).
__device__ __forceinline__ float comp4(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }
__device__ __forceinline__ float4 fma4(float s, float4 m, float4 acc) {
   return make_float4(fmaf(s, m.x, acc.x), fmaf(s, m.y, acc.y), fmaf(s, m.z, acc.z), fmaf(s, m.w, acc.w));
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// y = M^T a for up to three vectors at once (lane owns output dims 4*lane..)
template <int NVEC>
__device__ __forceinline__ void project3(const float* M, int D, int P, int lane, const float4 (&v)[NVEC], float4 (&y)[NVEC]) {
   const bool on = lane * 4 < P;
#pragma unroll
   for (int k = 0; k < NVEC; k++) y[k] = f4(0.f);
#pragma unroll 4
   for (int j = 0; j < D; j++) {
      const float4 m = on ? ld_cg4(M + (size_t)j * P + lane * 4) : f4(0.f);
#pragma unroll
      for (int k = 0; k < NVEC; k++) {
         const float e = __shfl_sync(0xffffffffu, comp4(v[k], j & 3), j >> 2);
         y[k] = fma4(e, m, y[k]);
      }
   }
}

// Rows [row_begin, row_end) of the unified row space (entities, then relations) whose stamp says
// "touched in this batch".  Two rows per group are examined per step, flags first, then both rows'
// loads, then the arithmetic, so that the L2 round trips overlap.
template <int MODEL, int LPS, int NV, bool DET>
__device__ __forceinline__ void publish_rows(const TrainArgs& a, long long row_begin, long long row_end, long long g0, long long G,
                                             uint32_t stamp, uint32_t next_stamp, int gl, uint32_t gmask, uint32_t& tent, uint32_t& trel) {
   const int P = a.P;
   const int lane = threadIdx.x & 31;
   long long first, end;
   group_range(row_begin, row_end, g0, G, first, end);
   const RowLists none{};
   auto stamped = [&](long long r) { return __ldcg(a.flag + r) == stamp; };
   auto finish = [&](long long r, float4 (&x)[NV], float4 (&d)[NV]) {
      if (r >= a.nE) { finish_relation<MODEL, LPS, NV, DET>(a, (int)(r - a.nE), gl, gmask, x, d); trel += (gl == 0); }
      else { finish_entity<MODEL, LPS, NV, false, DET>(a, none, (int)r, gl, gmask, next_stamp, x, d); tent += (gl == 0); }
   };
   for_stamped_rows<LPS>(first, end, gl, gmask, lane, stamped, [&](long long r0, long long r1) {
      float4 x0[NV], d0[NV], x1[NV], d1[NV];
      load_row<LPS, NV>(a.tab + (size_t)r0 * P, P, gl, x0);
      load_row<LPS, NV>(a.dtab + (size_t)r0 * P, P, gl, d0);
      if (r1 >= 0) {
         load_row<LPS, NV>(a.tab + (size_t)r1 * P, P, gl, x1);
         load_row<LPS, NV>(a.dtab + (size_t)r1 * P, P, gl, d1);
      }
      finish(r0, x0, d0);
      if (r1 >= 0) finish(r1, x1, d1);
   });
}

// LIST kernels: the CTA's own list, n entries, one per group and pass (two of a group in flight when the list is longer
// than the CTA has groups).  The claim (atomic exchange of the row's stamp) travels with the loads of the row; a row that
// another copy of the entry -- in this or in another CTA -- claimed first is dropped.
template <int MODEL, int LPS, int NV, bool DET>
__device__ __forceinline__ void publish_list(const TrainArgs& a, const RowLists& L, const int* list, int n, int group, int groups,
                                             uint32_t stamp, uint32_t next_stamp, int gl, uint32_t gmask, uint32_t& tent, uint32_t& trel,
                                             unsigned long long* fine = nullptr, int rows_per_model = 0) {
   const int P = a.P;
   const int leader = (threadIdx.x & 31) - gl;
   // tuning aid (KB2E_TRAIN_TRACE_FINE): thread 0 stamps its own group's claim / first row / second row
   auto mark = [&](int k, unsigned long long extra) {
      if (fine != nullptr && threadIdx.x == 0) {
         unsigned long long t_;
         asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
         fine[k] = t_;
         fine[k + 3] = extra;
      }
   };
   auto finish = [&](int r, float4 (&x)[NV], float4 (&d)[NV], int rlo, int rhi) {
      // (several stacked models, TransE only: row r of the stack; relation and entity rows are finished alike)
      const bool is_rel = rows_per_model ? (r % rows_per_model) >= a.nE : r >= a.nE;
      if (is_rel) { finish_relation<MODEL, LPS, NV, DET>(a, r - a.nE, gl, gmask, x, d); trel += (gl == 0); }
      else { finish_entity<MODEL, LPS, NV, true, DET>(a, L, r, gl, gmask, next_stamp, x, d, MODEL != KB2E_MODEL_TRANSE, rlo, rhi); tent += (gl == 0); }
   };
   for (int i = group; i < n; i += 2 * groups) {
      const int r0 = list[i];
      const int r1 = i + groups < n ? list[i + groups] : -1;
      int mine0 = 0, mine1 = 0;
      if (gl == 0) {
         mine0 = atomicExch(a.flag + r0, stamp) != stamp;
         if (r1 >= 0) mine1 = atomicExch(a.flag + r1, stamp) != stamp;
      }
      float4 x0[NV], d0[NV], x1[NV], d1[NV];
      load_row_pinned<LPS, NV>(a.tab + (size_t)r0 * P, P, gl, x0);   // pinned: issued with the claim, not after it
      load_row_pinned<LPS, NV>(a.dtab + (size_t)r0 * P, P, gl, d0);
      // TransH entity rows: the range of relations that touched the row (a dependent L2 round trip if left to finish_entity)
      int lo0 = 0, hi0 = -1, lo1 = 0, hi1 = -1;
      if (MODEL != KB2E_MODEL_TRANSE && r0 < a.nE) {
         lo0 = (int)ld_cg_u32_pinned(reinterpret_cast<const uint32_t*>(a.rmin + r0));
         hi0 = (int)ld_cg_u32_pinned(reinterpret_cast<const uint32_t*>(a.rmax + r0));
      }
      if (r1 >= 0) {
         load_row_pinned<LPS, NV>(a.tab + (size_t)r1 * P, P, gl, x1);
         load_row_pinned<LPS, NV>(a.dtab + (size_t)r1 * P, P, gl, d1);
         if (MODEL != KB2E_MODEL_TRANSE && r1 < a.nE) {
            lo1 = (int)ld_cg_u32_pinned(reinterpret_cast<const uint32_t*>(a.rmin + r1));
            hi1 = (int)ld_cg_u32_pinned(reinterpret_cast<const uint32_t*>(a.rmax + r1));
         }
      }
      mine0 = __shfl_sync(gmask, mine0, leader);
      mine1 = __shfl_sync(gmask, mine1, leader);
      mark(0, (unsigned long long)n);
      if (mine0) finish(r0, x0, d0, lo0, hi0);
      mark(1, (unsigned long long)(mine0 + 2 * mine1));
      if (mine1) finish(r1, x1, d1, lo1, hi1);
      mark(2, (unsigned long long)(r1 >= 0));
   }
}

template <int MODEL, int LPS, int NV, int THREADS, bool LIST, bool DET>
__global__ void __launch_bounds__(THREADS, 1) train_kernel(const __grid_constant__ TrainArgs a) {
   __shared__ double s_loss[THREADS / 32];
   extern __shared__ int s_lists[];
   RowLists L{};
   if (LIST) {
      L.count = s_lists;
      L.cap_ent = a.cap_ent;
      L.cap_rel = a.cap_rel;
      L.ent = s_lists + 4;
      L.rel[0] = L.ent + a.cap_ent;
      L.rel[1] = L.rel[0] + a.cap_rel;
   }
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((lane / LPS) * LPS));
   const int groups_per_block = blockDim.x / LPS;
   const long long G = (long long)gridDim.x * groups_per_block;
   // Samples (and rows) are dealt round-robin over CTAs so every SM gets an equal share.
   const long long g0 = (long long)(threadIdx.x / LPS) * gridDim.x + blockIdx.x;
   const long long R = (long long)a.nE + a.nR;
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   int trace_slot = 0;
#define KB2E_TRACE()                                                                                  \
   if (a.trace != nullptr && threadIdx.x == 0 && trace_slot < kTraceSlots) {                          \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      a.trace[(size_t)blockIdx.x * kTraceSlots + trace_slot++] = t_;                                  \
   }
   const uint32_t gb_first = (uint32_t)a.first_epoch * (uint32_t)a.batches;
   // The sampler does not depend on the embeddings, so each group draws its first sample of the
   // NEXT batch (triple fetch + rejection probes) before waiting at the end-of-batch barrier.
   Pair pre;
   DrawStage ds;
   const bool has_first = g0 < a.batchsize;
   const uint32_t n_batches = (uint32_t)a.n_epochs * (uint32_t)a.batches;
   // (The pipelined stages are a LIST-kernel feature: small batches, one sample per group.  In the scan kernels -- hundreds
   // of samples per group and batch, drawn inline -- their live registers only add spills to the hot loop: 240 instead of
   // 96 bytes at the scaled shape, 504 instead of 332 ms per epoch, measured.)
   if (has_first) {
      pre = draw_pair(a, (uint32_t)g0, gb_first);
      if (LIST && n_batches > 1u) draw_begin(a, (uint32_t)g0, gb_first + 1u, ds);   // pipelined from here on (train_device.cuh)
   }
   const int group = threadIdx.x / LPS;
   if (LIST) {
      if (threadIdx.x < 4) L.count[threadIdx.x] = 0;
      __syncthreads();
      if (MODEL != KB2E_MODEL_TRANSE) {
         // relation rows the LAST launch's final entity phase marked for this launch's first batch (the perturbation of
         // w_r carried into the next batch's delta): the list they were put on died with that launch
         const uint32_t s0 = a.stamp_base + 1u;
         for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < a.nR; r += gridDim.x * blockDim.x)
            if (__ldcg(a.cflag + r) == s0) list_push(L.count + 1 + (s0 & 1u), L.rel[s0 & 1u], L.cap_rel, a.nE + r, a.counters);
         __syncthreads();
      }
   }

   uint32_t rel_batch = 0;
   for (int ep = 0; ep < a.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < a.batches; batch++, rel_batch++) {
         const uint32_t gb = gb_first + rel_batch;
         // stamps count the batches this CONTEXT has run (not the caller's epoch numbers), so a stamp never recurs
         const uint32_t stamp = a.stamp_base + rel_batch + 1u;
         const uint32_t next_stamp = stamp + 1u;
         KB2E_TRACE();
         // ---- phase 1 ----
         if (has_first) process_pair<MODEL, LPS, NV, LIST, DET>(a, L, pre, gl, gmask, stamp, loss_acc, active_acc);
         for (long long k = g0 + G; k < a.batchsize; k += G) {
            Pair s = draw_pair(a, (uint32_t)k, gb);
            process_pair<MODEL, LPS, NV, LIST, DET>(a, L, s, gl, gmask, stamp, loss_acc, active_acc);
         }
         KB2E_TRACE();
         grid_barrier(a.barrier, bar_target);
         KB2E_TRACE();
         if (a.phase1_only) continue;   // kb2e_train_batch_deltas: the caller reads the raw delta tables (one batch per launch)
         const bool more = rel_batch + 1u < n_batches;
         // next batch's sample, stage 2: its probe loads travel with the row loads of the publish below
         if (LIST && has_first && more) draw_probe(a, ds);
         // ---- phase 2 ----
         if (LIST) {
            if (MODEL == KB2E_MODEL_TRANSE) {
               unsigned long long* fine = nullptr;
               if ((a.flags & 0x80000000u) && a.trace != nullptr && trace_slot + 8 < kTraceSlots) {
                  fine = a.trace + (size_t)blockIdx.x * kTraceSlots + trace_slot;
                  trace_slot += 6;
               }
               publish_list<MODEL, LPS, NV, DET>(a, L, L.ent, min(L.count[0], L.cap_ent), group, groups_per_block, stamp, next_stamp, gl, gmask,
                                            tent_acc, trel_acc, fine);
               KB2E_TRACE();
            } else {
               int* nrel = L.count + 1 + (stamp & 1u);
               publish_list<MODEL, LPS, NV, DET>(a, L, L.rel[stamp & 1u], min(*nrel, L.cap_rel), group, groups_per_block, stamp, next_stamp, gl, gmask,
                                            tent_acc, trel_acc);
               grid_barrier(a.barrier, bar_target);   // (its leading bar.sync also ends every read of *nrel)
               if (threadIdx.x == 0) *nrel = 0;
               KB2E_TRACE();
               publish_list<MODEL, LPS, NV, DET>(a, L, L.ent, min(L.count[0], L.cap_ent), group, groups_per_block, stamp, next_stamp, gl, gmask,
                                            tent_acc, trel_acc);
            }
            __syncthreads();
            if (threadIdx.x == 0) L.count[0] = 0;
         } else if (MODEL == KB2E_MODEL_TRANSE) {
            // no coupling between relation and entity rows: one pass over the whole row space
            publish_rows<MODEL, LPS, NV, DET>(a, 0, R, g0, G, stamp, next_stamp, gl, gmask, tent_acc, trel_acc);
            KB2E_TRACE();
         } else {
            publish_rows<MODEL, LPS, NV, DET>(a, a.nE, R, g0, G, stamp, next_stamp, gl, gmask, tent_acc, trel_acc);
            grid_barrier(a.barrier, bar_target);
            KB2E_TRACE();
            publish_rows<MODEL, LPS, NV, DET>(a, 0, a.nE, g0, G, stamp, next_stamp, gl, gmask, tent_acc, trel_acc);
         }
         KB2E_TRACE();
         grid_arrive(a.barrier, bar_target);
         // next batch's sample, stage 3 (registers only unless the candidate has to be redrawn), and stage 1 of the batch
         // after it: the triple fetch is in flight until the probe stage needs it, one phase from now
         if (has_first && more) {
            if (LIST) {
               pre = draw_finish(a, (uint32_t)g0, gb + 1u, ds);
               if (rel_batch + 2u < n_batches) draw_begin(a, (uint32_t)g0, gb + 2u, ds);
            } else {
               pre = draw_pair(a, (uint32_t)g0, gb + 1u);   // scan kernels: in one piece, while the other CTAs arrive
            }
         }
         grid_wait(a.barrier, bar_target);
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
         if (t != 0.0) {
            if (DET)   // CTA partial sums are reproducible; add them as integers
               atomicAdd(reinterpret_cast<unsigned long long*>(a.loss + ep), (unsigned long long)__double2ll_rn(t * kDetLossScale));
            else
               atomicAdd(a.loss + ep, t);
         }
      }
      __syncthreads();
   }
   // counters
   uint32_t c0 = (gl == 0) ? active_acc : 0u, c1 = tent_acc, c2 = trel_acc;
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


// ================================ TransH with a small relation set: relation rows resident in shared memory ==========
// WN18 has 18 relations (BASELINE config 2).  In train_kernel the relation-side publish is then a phase of its own in which
// 18 groups of the whole grid work and everybody else waits at a barrier (6 us of a 14-us batch, per-phase trace), and every
// entity row of the next phase starts with a dependent L2 load of w_r.  Here EVERY CTA keeps all relation-side rows
// (d_r, w_r) in its shared memory for the whole launch and finishes the touched ones itself after the first barrier --
// the inputs (published row + the batch's delta) are the same everywhere and the arithmetic is deterministic, so the 148
// copies stay bit-identical; CTA 0 writes them back when the launch ends.  What this needs:
//   * the relation-side deltas cannot be zeroed by "the" publisher while other CTAs still read them, so they ping-pong:
//     in-launch batch i accumulates into buffer i & 1 (0 = the context's dtab / dw, 1 = relbuf1), the entity-side
//     constraint carries of batch i go to buffer (i + 1) & 1, and that buffer is zeroed during phase 1 of batch i (its
//     last readers finished before the previous end-of-batch barrier);
//   * at the end of the launch the pending carries are moved to buffer 0 and buffer 1 is left zero, so every other kernel
//     of the context finds the usual state.
// Same samples, same arithmetic, same order of operations per row as train_kernel<TransH, ..., LIST>: in the deterministic
// mode the tables are bit-identical (tests/test_gpu_train.py).  Two grid barriers per batch instead of three.
// Phase 2a of train_transh_sr_kernel: groups of RL lanes, one touched relation each; the stamps and both delta rows are
// requested together (pinned loads: one L2 round trip, not two).  Returns 1 per finished relation in lane 0 of its group, CTA 0.
template <int RL, int RNV, bool DET>
__device__ __forceinline__ uint32_t sr_finish_relations(const TrainArgs& a, const RowLists& L, uint32_t stamp, unsigned long long* fine = nullptr) {
   // tuning aid (KB2E_TRAIN_TRACE_FINE): thread 0 stamps "deltas arrived", "d_r normalised", "w_r finished" + the loop's step count
   auto mark = [&](int k, unsigned long long extra) {
      if (fine != nullptr && threadIdx.x == 0) {
         unsigned long long t_;
         asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
         fine[k] = extra ? extra : t_;
      }
   };
   const int P = a.P;
   const int lane = threadIdx.x & 31;
   const int gl = lane % RL;
   const uint32_t gmask = RL == 32 ? 0xffffffffu : (((1u << RL) - 1u) << ((lane / RL) * RL));
   const int group = threadIdx.x / RL, groups = blockDim.x / RL;
   uint32_t done = 0;
   for (int r = group; r < a.nR; r += groups) {
      float4 x[RNV], d[RNV], b[RNV], db[RNV];
      const uint32_t f0 = ld_cg_u32_pinned(a.flag + a.nE + r), f1 = ld_cg_u32_pinned(a.cflag + r);
      load_row_pinned<RL, RNV>(L.dr_cur + (size_t)r * P, P, gl, d);
      load_row_pinned<RL, RNV>(L.dw_cur + (size_t)r * P, P, gl, db);
      if (f0 != stamp && f1 != stamp) continue;
      if (fine != nullptr) { asm volatile("" :: "f"(d[0].x), "f"(db[0].x)); mark(0, 0); }
      load_row_shared<RL, RNV>(L.srel + (size_t)(2 * r) * P, P, gl, x);
      load_row_shared<RL, RNV>(L.srel + (size_t)(2 * r + 1) * P, P, gl, b);
#pragma unroll
      for (int q = 0; q < RNV; q++) {
         x[q] = x[q] + (DET ? det_to_float(d[q]) : d[q]);
         b[q] = b[q] + (DET ? det_to_float(db[q]) : db[q]);
      }
      norm_row<RL, RNV>(x, true, gmask);
      if (fine != nullptr) { asm volatile("" :: "f"(x[0].x)); mark(1, 0); }
      const int steps = finish_hyperplane<RL, RNV>(x, b, a.lr, gmask);
      if (fine != nullptr) { asm volatile("" :: "f"(b[0].x)); mark(2, 0); mark(3, 1000000ull + (unsigned long long)steps); }
#pragma unroll
      for (int q = 0; q < RNV; q++) {
         const int off = (q * RL + gl) * 4;
         if (off < P) {
            *reinterpret_cast<float4*>(L.srel + (size_t)(2 * r) * P + off) = x[q];
            *reinterpret_cast<float4*>(L.srel + (size_t)(2 * r + 1) * P + off) = b[q];
         }
      }
      done += (gl == 0 && blockIdx.x == 0);
   }
   return done;
}

template <int LPS, int NV, int THREADS, bool DET>
__global__ void __launch_bounds__(THREADS, 1) train_transh_sr_kernel(const __grid_constant__ TrainArgs a) {
   constexpr int MODEL = KB2E_MODEL_TRANSH;
   __shared__ double s_loss[THREADS / 32];
   extern __shared__ int s_lists[];
   const int P = a.P, P4 = P >> 2;
   RowLists L{};
   L.count = s_lists;
   L.cap_ent = a.cap_ent;
   L.cap_rel = 0;
   L.ent = s_lists + 4;
   L.srel = reinterpret_cast<float*>(s_lists + 4 + ((a.cap_ent + 3) & ~3));
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((lane / LPS) * LPS));
   const int groups_per_block = blockDim.x / LPS;
   const long long G = (long long)gridDim.x * groups_per_block;
   const long long g0 = (long long)(threadIdx.x / LPS) * gridDim.x + blockIdx.x;
   const int group = threadIdx.x / LPS;
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   int trace_slot = 0;
   const uint32_t gb_first = (uint32_t)a.first_epoch * (uint32_t)a.batches;
   const uint32_t n_batches = (uint32_t)a.n_epochs * (uint32_t)a.batches;
   float* const dr_buf0 = a.dtab + (size_t)a.nE * P;
   float* const dr_buf1 = a.relbuf1;
   float* const dw_buf0 = a.dw;
   float* const dw_buf1 = a.relbuf1 + (size_t)a.nR * P;
   Pair pre;
   DrawStage ds;
   const bool has_first = g0 < a.batchsize;
   if (has_first) {
      pre = draw_pair(a, (uint32_t)g0, gb_first);
      if (n_batches > 1u) draw_begin(a, (uint32_t)g0, gb_first + 1u, ds);
   }
   for (int i = threadIdx.x; i < a.nR * P4; i += blockDim.x) {
      const int r = i / P4, k = i - r * P4;
      reinterpret_cast<float4*>(L.srel)[(size_t)(2 * r) * P4 + k] = ld_cg4(a.tab + ((size_t)a.nE + r) * P + 4 * k);
      reinterpret_cast<float4*>(L.srel)[(size_t)(2 * r + 1) * P4 + k] = ld_cg4(a.w + (size_t)r * P + 4 * k);
   }
   if (threadIdx.x < 4) L.count[threadIdx.x] = 0;
   __syncthreads();

   uint32_t rel_batch = 0;
   for (int ep = 0; ep < a.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < a.batches; batch++, rel_batch++) {
         const uint32_t gb = gb_first + rel_batch;
         const uint32_t stamp = a.stamp_base + rel_batch + 1u;
         const uint32_t next_stamp = stamp + 1u;
         const int cur = (int)(rel_batch & 1u);
         L.dr_cur = cur ? dr_buf1 : dr_buf0;
         L.dw_cur = cur ? dw_buf1 : dw_buf0;
         L.dw_next = cur ? dw_buf0 : dw_buf1;
         float* const dr_next = cur ? dr_buf0 : dr_buf1;
         KB2E_TRACE();
         // the buffer the carries of this batch go to: consumed one batch ago, zero it (relation r by CTA r % gridDim.x)
         for (int r = blockIdx.x; r < a.nR; r += gridDim.x)
            for (int k = threadIdx.x; k < 2 * P4; k += blockDim.x)
               st_cg4((k < P4 ? dr_next : L.dw_next) + (size_t)r * P + 4 * (k < P4 ? k : k - P4), f4(0.f));
         // ---- phase 1 ----
         if (has_first) process_pair<MODEL, LPS, NV, true, DET>(a, L, pre, gl, gmask, stamp, loss_acc, active_acc);
         for (long long k = g0 + G; k < a.batchsize; k += G) {
            Pair s = draw_pair(a, (uint32_t)k, gb);
            process_pair<MODEL, LPS, NV, true, DET>(a, L, s, gl, gmask, stamp, loss_acc, active_acc);
         }
         KB2E_TRACE();
         grid_barrier(a.barrier, bar_target);   // (ends every read of the shared relation rows in this CTA, too)
         KB2E_TRACE();
         if (a.phase1_only) continue;
         const bool more = rel_batch + 1u < n_batches;
         if (has_first && more) draw_probe(a, ds);
         // ---- phase 2a, in every CTA: the touched relation-side rows, shared memory to shared memory ----
         // 32 x 1 rows are handled by half-warps (16 lanes x 2 vectors): twice the groups, so WN18's 18 relations take one
         // round instead of two.  The sums come out bit-identical: lane i adds elements i and i + 16 first, which is exactly
         // the first stage of the 32-lane butterfly.
         unsigned long long* fine = nullptr;
         if ((a.flags & 0x80000000u) && a.trace != nullptr && trace_slot + 6 < kTraceSlots) {
            fine = a.trace + (size_t)blockIdx.x * kTraceSlots + trace_slot;
            trace_slot += 4;
         }
         if (LPS == 32 && NV == 1) trel_acc += sr_finish_relations<16, 2, DET>(a, L, stamp, fine);
         else trel_acc += sr_finish_relations<LPS, NV, DET>(a, L, stamp, fine);
         __syncthreads();
         KB2E_TRACE();
         // ---- phase 2b: entity rows (w_r from shared memory) ----
         publish_list<MODEL, LPS, NV, DET>(a, L, L.ent, min(L.count[0], L.cap_ent), group, groups_per_block, stamp, next_stamp, gl, gmask,
                                           tent_acc, trel_acc);
         __syncthreads();
         if (threadIdx.x == 0) L.count[0] = 0;
         KB2E_TRACE();
         grid_arrive(a.barrier, bar_target);
         if (has_first && more) {
            pre = draw_finish(a, (uint32_t)g0, gb + 1u, ds);
            if (rel_batch + 2u < n_batches) draw_begin(a, (uint32_t)g0, gb + 2u, ds);
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
         if (t != 0.0) {
            if (DET)
               atomicAdd(reinterpret_cast<unsigned long long*>(a.loss + ep), (unsigned long long)__double2ll_rn(t * kDetLossScale));
            else
               atomicAdd(a.loss + ep, t);
         }
      }
      __syncthreads();
   }
   // ---- end of the launch (after the last end-of-batch barrier: nobody reads a delta buffer any more) ----
   if (blockIdx.x == 0 && !a.phase1_only) {
      for (int i = threadIdx.x; i < a.nR * P4; i += blockDim.x) {
         const int r = i / P4, k = i - r * P4;
         st_cg4(a.tab + ((size_t)a.nE + r) * P + 4 * k, reinterpret_cast<const float4*>(L.srel)[(size_t)(2 * r) * P4 + k]);
         st_cg4(a.w + (size_t)r * P + 4 * k, reinterpret_cast<const float4*>(L.srel)[(size_t)(2 * r + 1) * P4 + k]);
         // pending carries -> buffer 0; buffer 1 zero.  The last batch (index n - 1) accumulated into buffer (n - 1) & 1, which
         // is spent, and carried into buffer n & 1.
         const size_t off = (size_t)r * P + 4 * k;
         if (n_batches & 1u) {
            st_cg4(dr_buf0 + off, ld_cg4(dr_buf1 + off));
            st_cg4(dw_buf0 + off, ld_cg4(dw_buf1 + off));
         }
         st_cg4(dr_buf1 + off, f4(0.f));
         st_cg4(dw_buf1 + off, f4(0.f));
      }
   }
   uint32_t c0 = (gl == 0) ? active_acc : 0u, c1 = tent_acc, c2 = trel_acc;
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


// ================================ batched training: K models in one persistent launch ==============================
// SURVEY.md 8f row 4 (seed / hyper-parameter sweeps: "tables are tiny; many configs fit at once").  At FB15k shape one
// model's batch is ~10 us of DEPENDENT L2 round trips and grid barriers with four fifths of the issue slots idle
// (profiles/README.md): the machine is latency-bound, not throughput-bound.  K independent TransE models of one size on
// one KG -- different seeds, learning rates, margins -- therefore share the launch, the phases and the barriers: the
// K x (nE + nR) rows are stacked in tab / dtab / flag, a batch consists of K * batchsize (model, sample) TASKS dealt
// round-robin over all groups (task tau -> model tau % K, sample tau / K; a group handles tasks_per_group of them, one
// after the other in phase 1), and the publish walks one list of stacked rows.  Every model computes exactly what it would
// compute alone (same counter-RNG stream, same batch semantics; bit-identical tables in the deterministic mode --
// tests/test_gpu_train.py); the batch takes about as long as one model's, so throughput grows almost K-fold.
// Sampling: lane t of a group owns the group's t-th task and runs the three pipelined stages of train_device.cuh for it (in
// the single-model kernel all lanes draw the same sample redundantly); finished samples wait in shared memory.
// The per-model epoch loss is accumulated in shared memory as a 64-bit fixed-point sum (2^-24 units).
constexpr double kSweepLossScale = 16777216.0;

template <int LPS, int NV, int THREADS, bool DET>
__global__ void __launch_bounds__(THREADS, 1) train_sweep_kernel(const __grid_constant__ TrainArgs a) {
   extern __shared__ int s_dyn[];
   const int K = a.replicas, T = a.tasks_per_group;
   const int groups_per_block = THREADS / LPS;
   RowLists L{};
   L.count = s_dyn;
   L.cap_ent = a.cap_ent;
   L.cap_rel = 0;
   L.ent = s_dyn + 4;
   L.rel[0] = L.rel[1] = nullptr;
   int4* s_pairs = reinterpret_cast<int4*>(s_dyn + 4 + ((a.cap_ent + 3) & ~3));           // [groups][T]: h, t, r, c | corruptTail << 31; h = -1: no task
   unsigned long long* s_loss = reinterpret_cast<unsigned long long*>(s_pairs + (size_t)groups_per_block * T);   // [K]
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((lane / LPS) * LPS));
   const int group = threadIdx.x / LPS;
   const long long G = (long long)gridDim.x * groups_per_block;
   const long long g0 = (long long)group * gridDim.x + blockIdx.x;
   const long long n_tasks = (long long)K * a.batchsize;
   const int R = a.nE + a.nR;
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   const uint32_t gb_first = (uint32_t)a.first_epoch * (uint32_t)a.batches;
   const uint32_t n_batches = (uint32_t)a.n_epochs * (uint32_t)a.batches;
   // this lane's sampling task: the gl-th task of the group
   const long long tau = g0 + (long long)gl * G;
   const bool samples = gl < T && tau < n_tasks;
   const int my_model = samples ? (int)(tau % K) : 0;
   const uint32_t my_k = (uint32_t)(tau / K);
   RepParams mine = a.rep[my_model];
   DrawStage ds;
   int trace_slot = 0;
#define KB2E_STRACE()                                                                                 \
   if (a.trace != nullptr && threadIdx.x == 0 && trace_slot < kTraceSlots) {                          \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      a.trace[(size_t)blockIdx.x * kTraceSlots + trace_slot++] = t_;                                  \
   }
   int4* my_slot = s_pairs + (size_t)group * T + gl;
   auto park = [&](const Pair& p) { *my_slot = make_int4(p.h, p.t, p.r, p.c | (p.corruptTail ? (int)0x80000000 : 0)); };
   if (threadIdx.x < 4) L.count[threadIdx.x] = 0;
   for (int m = threadIdx.x; m < K; m += THREADS) s_loss[m] = 0ull;
   if (gl < T) {
      if (samples) {
         park(draw_pair(a, my_k, gb_first, mine.seed_lo, mine.seed_hi));
         if (n_batches > 1u) draw_begin(a, my_k, gb_first + 1u, ds, mine.seed_lo, mine.seed_hi);
      } else {
         *my_slot = make_int4(-1, 0, 0, 0);
      }
   }
   __syncthreads();

   uint32_t rel_batch = 0;
   for (int ep = 0; ep < a.n_epochs; ep++) {
      for (int batch = 0; batch < a.batches; batch++, rel_batch++) {
         const uint32_t gb = gb_first + rel_batch;
         const uint32_t stamp = a.stamp_base + rel_batch + 1u;
         KB2E_STRACE();
         // ---- phase 1: the group's tasks, one after the other ----
         for (int t = 0; t < T; t++) {
            const int4 pk = s_pairs[(size_t)group * T + t];
            if (pk.x < 0) break;
            Pair s;
            s.h = pk.x; s.t = pk.y; s.r = pk.z; s.c = pk.w & 0x7fffffff; s.corruptTail = pk.w < 0;
            const int m = (int)((g0 + (long long)t * G) % K);
            const RepParams rp = a.rep[m];
            double loss = 0.0;
            process_pair_at<KB2E_MODEL_TRANSE, LPS, NV, true, DET>(a, L, s, (size_t)m * R, rp.lr, rp.margin, gl, gmask, stamp, loss, active_acc);
            if (gl == 0 && loss != 0.0) atomicAdd(s_loss + m, (unsigned long long)__double2ll_rn(loss * kSweepLossScale));
         }
         KB2E_STRACE();
         grid_barrier(a.barrier, bar_target);
         KB2E_STRACE();
         const bool more = rel_batch + 1u < n_batches;
         if (samples && more) draw_probe(a, ds);
         // ---- phase 2: the CTA's list of stacked rows ----
         publish_list<KB2E_MODEL_TRANSE, LPS, NV, DET>(a, L, L.ent, min(L.count[0], L.cap_ent), group, groups_per_block, stamp, stamp + 1u, gl, gmask,
                                                  tent_acc, trel_acc, nullptr, R);
         KB2E_STRACE();
         __syncthreads();
         if (threadIdx.x == 0) L.count[0] = 0;
         KB2E_STRACE();
         grid_arrive(a.barrier, bar_target);
         if (samples && more) {
            park(draw_finish(a, my_k, gb + 1u, ds, mine.seed_lo, mine.seed_hi));
            if (rel_batch + 2u < n_batches) draw_begin(a, my_k, gb + 2u, ds, mine.seed_lo, mine.seed_hi);
         }
         grid_wait(a.barrier, bar_target);   // (its closing bar.sync publishes the parked samples to the whole group)
      }
      // epoch losses: one integer atomic per model and CTA (a.loss holds K x n_epochs 64-bit fixed-point sums)
      __syncthreads();
      for (int m = threadIdx.x; m < K; m += THREADS) {
         if (s_loss[m]) atomicAdd(reinterpret_cast<unsigned long long*>(a.loss) + (size_t)m * a.n_epochs + ep, s_loss[m]);
         s_loss[m] = 0ull;
      }
      __syncthreads();
   }
   uint32_t c0 = (gl == 0) ? active_acc : 0u, c1 = tent_acc, c2 = trel_acc;
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
// The sample through BOTH forms of the sampler -- draw_pair (one piece: the TransR, one-barrier and partitioned kernels)
// and the three pipelined stages of train_kernel; a disagreement poisons the row so that the parity test fails loudly.
__global__ void sample_kernel(const TrainArgs a, uint32_t gb, long long count, int32_t* out) {
   long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= count) return;
   Pair s = draw_pair(a, (uint32_t)k, gb);
   DrawStage ds;
   draw_begin(a, (uint32_t)k, gb, ds);
   draw_probe(a, ds);
   const Pair s2 = draw_finish(a, (uint32_t)k, gb, ds);
   if (s2.h != s.h || s2.t != s.t || s2.r != s.r || s2.c != s.c || s2.corruptTail != s.corruptTail) s.h = s.t = s.r = s.c = -1;
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

// TransR fp32 energies with the arithmetic of process_pair_transr (one warp per triple).
__global__ void score32_transr_kernel(const float* tab, const float* w, size_t w_row, int nE, int D, int P, int distance,
                                      const int32_t* h, const int32_t* t, const int32_t* r, long long n, double* out) {
   long long k = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (k >= n) return;
   const bool on = lane * 4 < P;
   float4 v[2], y[2];
   v[0] = on ? ld_cg4(tab + (size_t)h[k] * P + lane * 4) : f4(0.f);
   v[1] = on ? ld_cg4(tab + (size_t)t[k] * P + lane * 4) : f4(0.f);
   const float4 vr = on ? ld_cg4(tab + ((size_t)nE + r[k]) * P + lane * 4) : f4(0.f);
   project3<2>(w + (size_t)r[k] * w_row, D, P, lane, v, y);
   const float4 res = (y[1] - y[0]) - vr;
   float e = distance == KB2E_DISTANCE_L1 ? abs4(res) : dot4(res, res);
   e = wsum(e);
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

// (h, t, r) columns -> packed int4 records; ids are validated here (first offending index -> *bad).
__global__ void pack_triples_kernel(const int32_t* h, const int32_t* t, const int32_t* r, long long n, int nE, int nR,
                                    int4* out, unsigned long long* bad) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   int hh = h[i], tt = t[i], rr = r[i];
   if (hh < 0 || hh >= nE || tt < 0 || tt >= nE || rr < 0 || rr >= nR) atomicMin(bad, (unsigned long long)i);
   out[i] = make_int4(hh, tt, rr, 0);
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

// triples[i].w = the corruption threshold of the triple's relation: coin < pr  <=>  coin < ceil(pr) for an integer coin
// (pr = NaN for a relation without triples compares false: threshold 0), so the staged sampler of train_device.cuh
// decides the corruption side from the triple record alone.
__global__ void fill_threshold_kernel(int4* triples, long long n, const double* __restrict__ pr) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const double p = pr[triples[i].z];
   triples[i].w = (p == p) ? (int)fmin(fmax(ceil(p), 0.0), 1001.0) : 0;
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
   const size_t all = (size_t)c->K * rows;   // K stacked models (kb2e_set_replicas), K = 1 otherwise
   if (all >= (1ull << 31)) return fail(c, KB2E_ERR_LIMIT, "more than 2^31 rows in the stacked tables");
   KB2E_CUDA(c, pool_alloc(c, &c->tab_all, all * c->P * sizeof(float)));
   KB2E_CUDA(c, pool_alloc(c, &c->dtab_all, all * c->P * sizeof(float)));
   KB2E_CUDA(c, cudaMemsetAsync(c->tab_all, 0, all * c->P * sizeof(float), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->dtab_all, 0, all * c->P * sizeof(float), c->stream));
   KB2E_CUDA(c, pool_alloc(c, &c->flag_all, all * sizeof(uint32_t)));
   KB2E_CUDA(c, cudaMemsetAsync(c->flag_all, 0, all * sizeof(uint32_t), c->stream));
   c->tab = c->tab_all + (size_t)c->sel * rows * c->P;
   c->dtab = c->dtab_all + (size_t)c->sel * rows * c->P;
   c->flag = c->flag_all + (size_t)c->sel * rows;
   if (c->cfg.model != KB2E_MODEL_TRANSE) {
      c->w_row = c->cfg.model == KB2E_MODEL_TRANSH ? (size_t)c->P : (size_t)c->D * c->P;
      KB2E_CUDA(c, pool_alloc(c, &c->w, (size_t)c->nR * c->w_row * sizeof(float)));
      KB2E_CUDA(c, pool_alloc(c, &c->dw, (size_t)c->nR * c->w_row * sizeof(float)));
      KB2E_CUDA(c, cudaMemsetAsync(c->w, 0, (size_t)c->nR * c->w_row * sizeof(float), c->stream));
      KB2E_CUDA(c, cudaMemsetAsync(c->dw, 0, (size_t)c->nR * c->w_row * sizeof(float), c->stream));
      KB2E_CUDA(c, pool_alloc(c, &c->rmin, (size_t)c->nE * sizeof(int)));
      KB2E_CUDA(c, pool_alloc(c, &c->rmax, (size_t)c->nE * sizeof(int)));
      fill_int_kernel<<<blocks_for(c->nE, 256), 256, 0, c->stream>>>(c->rmin, c->nE, 0x7fffffff);
      fill_int_kernel<<<blocks_for(c->nE, 256), 256, 0, c->stream>>>(c->rmax, c->nE, -1);
   }
   KB2E_CUDA(c, pool_alloc(c, &c->cflag, (size_t)c->nR * sizeof(uint32_t)));
   KB2E_CUDA(c, cudaMemsetAsync(c->cflag, 0, (size_t)c->nR * sizeof(uint32_t), c->stream));
   KB2E_CUDA(c, pool_alloc(c, &c->barrier, 64));
   KB2E_CUDA(c, pool_alloc(c, &c->counters, 8 * sizeof(unsigned long long)));
   KB2E_CUDA(c, cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
   KB2E_CUDA(c, pool_alloc(c, &c->pr, (size_t)c->nR * sizeof(double)));
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

void train_free(kb2e_ctx* c) {
   pool_free(c, c->tab_all); pool_free(c, c->dtab_all); pool_free(c, c->w); pool_free(c, c->dw); pool_free(c, c->flag_all);
   pool_free(c, c->rep_dev);
   pool_free(c, c->rmin); pool_free(c, c->rmax); pool_free(c, c->triples); pool_free(c, c->stage); pool_free(c, c->hash); pool_free(c, c->pr);
   pool_free(c, c->barrier); pool_free(c, c->loss_dev); pool_free(c, c->counters); pool_free(c, c->pairs_dev);
   pool_free(c, c->ent64); pool_free(c, c->rel64); pool_free(c, c->w64);
   pool_free(c, c->pend);
   pool_free(c, c->cflag);
   pool_free(c, c->transr_aux);
   pool_free(c, c->relbuf1);
   pool_free(c, c->filt_dev);
}

int train_set_triples(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n) {
   int rc = train_alloc(c);
   if (rc) return rc;
   c->n_train = 0;
   if (n == 0) return KB2E_OK;
   // device buffers are kept across calls (capacity only grows): no allocation on the steady path
   if (n > c->triples_cap) {
      pool_free(c, c->triples); pool_free(c, c->stage);
      c->triples = nullptr; c->stage = nullptr; c->triples_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &c->stage, 3 * (size_t)n * sizeof(int32_t)));
      KB2E_CUDA(c, pool_alloc(c, &c->triples, (size_t)n * sizeof(int4)));
      c->triples_cap = n;
   }
   uint64_t slots = 1024;
   while (slots < 2 * (uint64_t)n) slots <<= 1;
   if (slots > c->hash_cap) {
      pool_free(c, c->hash);
      c->hash = nullptr; c->hash_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &c->hash, slots * sizeof(uint64_t)));
      c->hash_cap = slots;
   }
   c->hash_mask = slots - 1;
   int32_t* st = c->stage;
   KB2E_CUDA(c, cudaMemcpyAsync(st, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(st + n, t, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(st + 2 * n, r, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->counters + 7, 0xff, sizeof(unsigned long long), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->hash, 0xff, slots * sizeof(uint64_t), c->stream));
   pack_triples_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(st, st + n, st + 2 * n, n, c->nE, c->nR, c->triples, c->counters + 7);
   hash_insert_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->triples, n, c->hash, c->hash_mask);
   KB2E_CUDA(c, cudaGetLastError());
   unsigned long long bad = 0;
   KB2E_CUDA(c, cudaMemcpyAsync(&bad, c->counters + 7, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   if (bad != ~0ull) return fail(c, KB2E_ERR_ARG, "train triple " + std::to_string(bad) + " has an id out of range");
   c->n_train = n;
   c->thr_valid = false;
   return KB2E_OK;
}

int train_init_embeddings(kb2e_ctx* c) {
   int rc = train_alloc(c);
   if (rc) return rc;
   uint32_t k0 = (uint32_t)c->cfg.seed, k1 = (uint32_t)(c->cfg.seed >> 32);
   long long rows = (long long)c->nE + c->nR;
   if (c->K > 1) {
      // every stacked model from its own seed, exactly as a single-model context with that seed would start
      for (int m = 0; m < c->K; m++)
         init_rows_kernel<<<blocks_for(rows * 32, 256), 256, 0, c->stream>>>(c->tab_all + (size_t)m * rows * c->P, rows, c->D, c->P,
                                                                             (uint32_t)c->rep_seed[m], (uint32_t)(c->rep_seed[m] >> 32), 1u, 0);
   } else {
      init_rows_kernel<<<blocks_for(rows * 32, 256), 256, 0, c->stream>>>(c->tab, rows, c->D, c->P, k0, k1, 1u, 0);
   }
   if (c->cfg.model == KB2E_MODEL_TRANSH) {
      init_rows_kernel<<<blocks_for((long long)c->nR * 32, 256), 256, 0, c->stream>>>(c->w, c->nR, c->D, c->P, k0, k1, 2u, 1);
   } else if (c->cfg.model == KB2E_MODEL_TRANSR) {
      identity_kernel<<<blocks_for((long long)c->nR * c->D * c->P, 256), 256, 0, c->stream>>>(c->w, c->nR, c->D, c->P);
   }
   KB2E_CUDA(c, cudaGetLastError());
   for (int t = 0; t < 3; t++) { c->v32[t] = true; c->v64[t] = false; }
   c->tables_epoch++;
   return KB2E_OK;
}

int num_tables(const kb2e_ctx* c) { return c->cfg.model == KB2E_MODEL_TRANSE ? 2 : 3; }

static long long table_rows(const kb2e_ctx* c, int t) {
   if (t == KB2E_TABLE_ENTITY) return c->nE;
   if (t == KB2E_TABLE_RELATION) return c->nR;
   return c->cfg.model == KB2E_MODEL_TRANSH ? (long long)c->nR : (long long)c->nR * c->D;
}

static float* table32(kb2e_ctx* c, int t) {
   if (t == KB2E_TABLE_ENTITY) return c->tab;
   if (t == KB2E_TABLE_RELATION) return c->tab + (size_t)c->nE * c->P;
   return c->w;
}

static double** table64_slot(kb2e_ctx* c, int t) {
   return t == KB2E_TABLE_ENTITY ? &c->ent64 : (t == KB2E_TABLE_RELATION ? &c->rel64 : &c->w64);
}

double* table64(kb2e_ctx* c, int t) {
   double** slot = table64_slot(c, t);
   if (!*slot) {
      if (pool_alloc(c, slot, (size_t)table_rows(c, t) * c->D * sizeof(double)) != cudaSuccess) return nullptr;
   }
   return *slot;
}

int widen_table(kb2e_ctx* c, int t) {
   if (!c->v32[t]) return fail(c, KB2E_ERR_ARG, "table " + std::to_string(t) + " has no values: call kb2e_init_embeddings or kb2e_upload first");
   double* dst = table64(c, t);
   if (!dst) return fail(c, KB2E_ERR_CUDA, "out of device memory for the fp64 tables");
   long long rows = table_rows(c, t);
   widen_kernel<<<blocks_for(rows * c->D, 256), 256, 0, c->stream>>>(table32(c, t), dst, rows, c->D, c->P);
   KB2E_CUDA(c, cudaGetLastError());
   c->v64[t] = true;
   return KB2E_OK;
}

int narrow_table(kb2e_ctx* c, int t) {
   int rc = train_alloc(c);
   if (rc) return rc;
   if (!c->v64[t]) return fail(c, KB2E_ERR_ARG, "table " + std::to_string(t) + " has no values");
   long long rows = table_rows(c, t);
   narrow_kernel<<<blocks_for(rows * c->P, 256), 256, 0, c->stream>>>(*table64_slot(c, t), table32(c, t), rows, c->D, c->P);
   KB2E_CUDA(c, cudaGetLastError());
   c->v32[t] = true;
   return KB2E_OK;
}

int ensure32(kb2e_ctx* c) {
   for (int t = 0; t < num_tables(c); t++) {
      if (c->v32[t]) continue;
      int rc = narrow_table(c, t);
      if (rc) return rc;
   }
   return KB2E_OK;
}

int ensure64(kb2e_ctx* c) {
   for (int t = 0; t < num_tables(c); t++) {
      if (c->v64[t]) continue;
      int rc = widen_table(c, t);
      if (rc) return rc;
   }
   return KB2E_OK;
}

static void fill_args(kb2e_ctx* c, TrainArgs& a) {
   memset(&a, 0, sizeof(a));
   a.tab = c->tab; a.dtab = c->dtab; a.w = c->w; a.dw = c->dw;
   a.flag = c->flag;
   a.cflag = c->cflag;
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

// One persistent CTA per SM; fewer threads per CTA buy registers (65536 / THREADS per thread) for the
// shapes that keep more vectors per lane in flight: 1024 threads with one float4 per lane, 768 with two, 512 with
// four; TransH always 512 (five rows per sample live in registers: 768 threads = 80 registers spill).
static int threads_for(int model, int nv) {
   if (nv >= 4) return 512;
   if (model == KB2E_MODEL_TRANSH) return 512;
   if (nv == 2) return 768;
   return 1024;
}

template <int MODEL, bool LIST, bool DET>
static TrainKernel pick_kernel(int lps, int nv, int threads) {
#define KB2E_PICK(L, N, T) if (lps == L && nv == N) return train_kernel<MODEL, L, N, T, LIST, DET>;
   if (MODEL == KB2E_MODEL_TRANSH) {
      KB2E_PICK(8, 1, 512) KB2E_PICK(16, 1, 512) KB2E_PICK(32, 1, 512)
      KB2E_PICK(8, 2, 512) KB2E_PICK(16, 2, 512) KB2E_PICK(32, 2, 512)
   } else {
      KB2E_PICK(8, 1, 1024) KB2E_PICK(16, 1, 1024) KB2E_PICK(32, 1, 1024)
      if (threads == 640) { KB2E_PICK(8, 2, 640) KB2E_PICK(16, 2, 640) KB2E_PICK(32, 2, 640) }
      KB2E_PICK(8, 2, 768) KB2E_PICK(16, 2, 768) KB2E_PICK(32, 2, 768)
      KB2E_PICK(8, 4, 512) KB2E_PICK(16, 4, 512) KB2E_PICK(32, 4, 512)
   }
#undef KB2E_PICK
   return nullptr;
}

constexpr int kSmallRelations = 32;   // train_transh_sr_kernel: 2 * 32 rows of <= 512 floats next to the entity list

template <bool DET>
static TrainKernel pick_transh_sr(int lps, int nv) {
#define KB2E_PICK(L, N) if (lps == L && nv == N) return train_transh_sr_kernel<L, N, 512, DET>;
   KB2E_PICK(8, 1) KB2E_PICK(16, 1) KB2E_PICK(32, 1)
   KB2E_PICK(8, 2) KB2E_PICK(16, 2) KB2E_PICK(32, 2)
#undef KB2E_PICK
   return nullptr;
}

// Capacities of the per-CTA touched-row lists (LIST kernels) for a batch of `batchsize` pairs dealt round-robin to
// `groups` groups per CTA: S pairs per CTA and batch, each can be the first toucher of 3 entity rows and 1 relation row;
// a published entity row can mark up to two relation rows for the next batch (TransH).  Returns the dynamic shared
// memory in bytes, or 0 when the scan kernel should be used (lists too large, or a batch that touches most of the table).
static size_t list_shape(const kb2e_ctx* c, long long batchsize, int groups, int& cap_ent, int& cap_rel) {
   const char* env = getenv("KB2E_TRAIN_LIST");   // 0: always the scan kernel (tuning aid / A-B measurements)
   if (env && atoi(env) == 0) return 0;
   const long long G = (long long)c->num_sms * groups;
   const long long S = (long long)groups * ((batchsize + G - 1) / G);
   const bool transe = c->cfg.model == KB2E_MODEL_TRANSE;
   const long long ce = (transe ? 4 : 3) * S;
   const long long cr = transe ? 0 : 7 * S;   // S phase-1 entries + up to two marks per published entity row (3 S), duplicates included
   const size_t bytes = (size_t)(4 + ce + 2 * cr) * sizeof(int);
   if (bytes > 40 * 1024) return 0;   // (large batches -- the scaled shape: 6,757 pairs per CTA -- keep the stamp scan)
   cap_ent = (int)ce;
   cap_rel = (int)cr;
   return bytes;
}

// Choose lanes-per-sample (LPS) / float4-vectors-per-lane (NV): among the shapes that waste the
// fewest lanes on padding, take the widest group whose group count still covers the whole batch in
// one pass over the resident groups; KB2E_TRAIN_LPS overrides (tuning aid).
static void choose_shape(const kb2e_ctx* c, long long batchsize, int& lps, int& nv, int& threads_out) {
   const int vecs = (c->P + 3) / 4;
   const int maxnv = (c->cfg.model == KB2E_MODEL_TRANSH) ? 2 : 4;
   const char* env = getenv("KB2E_TRAIN_LPS");
   const int forced = env ? atoi(env) : 0;
   int Ls[3] = {32, 16, 8}, Ns[3], Ts[3];
   double eff[3], best = 0.0;
   for (int i = 0; i < 3; i++) {
      int N = (vecs + Ls[i] - 1) / Ls[i];
      Ns[i] = N <= 1 ? 1 : (N <= 2 ? 2 : (N <= 4 ? 4 : 99));
      Ts[i] = threads_for(c->cfg.model, Ns[i]);
      eff[i] = (Ns[i] <= maxnv || Ls[i] == 32) ? (double)vecs / (Ls[i] * Ns[i]) : 0.0;
      if (Ns[i] > 4) eff[i] = 0.0;
      best = std::max(best, eff[i]);
   }
   lps = 32; nv = Ns[0]; threads_out = Ts[0];
   if (best == 0.0) { nv = 99; return; }
   int widest = -1, fit = -1;
   for (int i = 0; i < 3; i++) {
      if (forced == Ls[i] && eff[i] > 0.0) { lps = Ls[i]; nv = Ns[i]; threads_out = Ts[i]; return; }
      if (eff[i] < 0.75 * best) continue;
      if (widest < 0) widest = i;
      if (fit < 0 && (long long)c->num_sms * Ts[i] / Ls[i] >= batchsize) fit = i;
   }
   int pick = fit >= 0 ? fit : widest;
   lps = Ls[pick]; nv = Ns[pick]; threads_out = Ts[pick];
}

// ---- batched training of K stacked models (train_sweep_kernel) -----------------------------------------------
int train_set_replicas(kb2e_ctx* c, int K, const double* rates, const double* margins, const uint64_t* seeds) {
   if (c->tab_all) return fail(c, KB2E_ERR_ARG, "kb2e_set_replicas: call it before any table exists (right after kb2e_create)");
   if (c->cfg.model != KB2E_MODEL_TRANSE) return fail(c, KB2E_ERR_LIMIT, "kb2e_set_replicas: batched training is built for TransE");
   if (K < 1 || K > 64) return fail(c, KB2E_ERR_ARG, "kb2e_set_replicas: 1 <= K <= 64");
   c->K = K;
   c->sel = 0;
   c->rep_rate.assign(K, c->cfg.rate);
   c->rep_margin.assign(K, c->cfg.margin);
   c->rep_seed.resize(K);
   for (int m = 0; m < K; m++) {
      if (rates) c->rep_rate[m] = rates[m];
      if (margins) c->rep_margin[m] = margins[m];
      c->rep_seed[m] = seeds ? seeds[m] : c->cfg.seed + (uint64_t)m;
   }
   return KB2E_OK;
}

int train_select_replica(kb2e_ctx* c, int m) {
   if (m < 0 || m >= c->K) return fail(c, KB2E_ERR_ARG, "kb2e_select_replica: no such model");
   int rc = train_alloc(c);
   if (rc) return rc;
   if (m == c->sel) return KB2E_OK;
   const size_t rows = (size_t)c->nE + c->nR;
   c->sel = m;
   c->tab = c->tab_all + (size_t)m * rows * c->P;
   c->dtab = c->dtab_all + (size_t)m * rows * c->P;
   c->flag = c->flag_all + (size_t)m * rows;
   // the fp64 copies (and everything the ranking derives from them) describe the previously selected model
   for (int t = 0; t < 3; t++) c->v64[t] = false;
   c->tables_epoch++;
   return KB2E_OK;
}

typedef void (*SweepKernel)(const TrainArgs);
template <bool DET>
static SweepKernel pick_sweep(int lps, int nv, int threads) {
#define KB2E_SWEEP(L, N, T) if (lps == L && nv == N && threads == T) return train_sweep_kernel<L, N, T, DET>;
   KB2E_SWEEP(8, 1, 1024) KB2E_SWEEP(16, 1, 1024) KB2E_SWEEP(32, 1, 1024)
   KB2E_SWEEP(8, 2, 640) KB2E_SWEEP(16, 2, 640) KB2E_SWEEP(32, 2, 640)
   KB2E_SWEEP(8, 4, 512) KB2E_SWEEP(16, 4, 512) KB2E_SWEEP(32, 4, 512)
#undef KB2E_SWEEP
   return nullptr;
}

static int sweep_run(kb2e_ctx* c, int first_epoch, int n_epochs, double* loss_out) {
   const int K = c->K;
   if (!c->triples || c->n_train == 0) return fail(c, KB2E_ERR_ARG, "no training triples: call kb2e_set_train_triples first");
   if (!c->have_pr) return fail(c, KB2E_ERR_ARG, "no corruption probabilities: call kb2e_set_bern first");
   if (c->cfg.batches <= 0) return fail(c, KB2E_ERR_ARG, "batches must be positive");
   if (!c->v32[0] || !c->v32[1]) return fail(c, KB2E_ERR_ARG, "the stacked models have no values: call kb2e_init_embeddings (or upload every model) first");
   if ((size_t)K * n_epochs > (size_t)c->loss_cap) {
      pool_free(c, c->loss_dev);
      c->loss_dev = nullptr;
      c->loss_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &c->loss_dev, (size_t)K * n_epochs * sizeof(double)));
      c->loss_cap = K * n_epochs;
   }
   TrainArgs a;
   fill_args(c, a);
   a.tab = c->tab_all; a.dtab = c->dtab_all; a.flag = c->flag_all;
   a.first_epoch = first_epoch;
   a.n_epochs = n_epochs;
   a.batchsize = c->n_train / c->cfg.batches;
   a.replicas = K;
   if (!c->thr_valid) {
      fill_threshold_kernel<<<blocks_for(c->n_train, 256), 256, 0, c->stream>>>(c->triples, c->n_train, c->pr);
      KB2E_CUDA(c, cudaGetLastError());
      c->thr_valid = true;
   }
   std::vector<RepParams> rp(K);
   for (int m = 0; m < K; m++) rp[m] = RepParams{(float)c->rep_rate[m], (float)c->rep_margin[m], (uint32_t)c->rep_seed[m], (uint32_t)(c->rep_seed[m] >> 32)};
   if (!c->rep_dev) KB2E_CUDA(c, pool_alloc(c, &c->rep_dev, (size_t)K * sizeof(RepParams)));
   KB2E_CUDA(c, cudaMemcpyAsync(c->rep_dev, rp.data(), (size_t)K * sizeof(RepParams), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));   // rp is a local
   a.rep = static_cast<const RepParams*>(c->rep_dev);
   int lps, nv, threads;
   choose_shape(c, a.batchsize, lps, nv, threads);
   if (nv == 2) threads = 640;
   SweepKernel k = (c->cfg.flags & KB2E_FLAG_DETERMINISTIC) ? pick_sweep<true>(lps, nv, threads) : pick_sweep<false>(lps, nv, threads);
   if (!k) return fail(c, KB2E_ERR_LIMIT, "no batched training kernel for this embedding size");
   const int groups = threads / lps;
   const long long G = (long long)c->num_sms * groups;
   const long long T = ((long long)K * a.batchsize + G - 1) / G;
   if (T > lps) return fail(c, KB2E_ERR_LIMIT, "too many models for one launch at this batch size: K * batch size must not exceed " +
                                                std::to_string(G * lps) + " (model, sample) tasks");
   a.tasks_per_group = (int)T;
   a.cap_ent = (int)(4 * groups * T);
   a.cap_rel = 0;
   const size_t smem = (size_t)(4 + ((a.cap_ent + 3) & ~3)) * sizeof(int) + (size_t)groups * T * sizeof(int4) + (size_t)K * sizeof(unsigned long long);
   if (smem > 96 * 1024) return fail(c, KB2E_ERR_LIMIT, "batched training: task lists do not fit in shared memory");
   KB2E_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   const uint64_t n_batches = (uint64_t)a.batches * (uint64_t)n_epochs;
   if (n_batches >= (1ull << 31)) return fail(c, KB2E_ERR_LIMIT, "more than 2^31 batches in one launch");
   if ((uint64_t)c->stamp_base + n_batches + 2 >= (1ull << 32)) {
      KB2E_CUDA(c, cudaMemsetAsync(c->flag_all, 0, (size_t)K * ((size_t)c->nE + c->nR) * sizeof(uint32_t), c->stream));
      c->stamp_base = 0;
   }
   a.stamp_base = c->stamp_base;
   KB2E_CUDA(c, cudaMemsetAsync(c->barrier, 0, 64, c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->loss_dev, 0, (size_t)K * n_epochs * sizeof(double), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->counters + 6, 0, sizeof(unsigned long long), c->stream));
   const char* trace_path = getenv("KB2E_TRAIN_TRACE");
   unsigned long long* trace_dev = nullptr;
   if (trace_path) {
      KB2E_CUDA(c, pool_alloc(c, &trace_dev, (size_t)c->num_sms * kTraceSlots * sizeof(unsigned long long)));
      KB2E_CUDA(c, cudaMemsetAsync(trace_dev, 0, (size_t)c->num_sms * kTraceSlots * sizeof(unsigned long long), c->stream));
      a.trace = trace_dev;
   }
   void* params[] = {&a};
   KB2E_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   KB2E_CUDA(c, cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(threads), params, smem, c->stream));
   KB2E_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   if (trace_dev) {
      std::vector<unsigned long long> tr((size_t)c->num_sms * kTraceSlots);
      cudaMemcpyAsync(tr.data(), trace_dev, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
      cudaStreamSynchronize(c->stream);
      pool_free(c, trace_dev);
      if (FILE* f = fopen(trace_path, "w")) {
         for (int b = 0; b < c->num_sms; b++)
            for (int q = 0; q < kTraceSlots; q++) fprintf(f, "%llu%c", tr[(size_t)b * kTraceSlots + q], q + 1 == kTraceSlots ? '\n' : ' ');
         fclose(f);
      }
   }
   std::vector<long long> fixed((size_t)K * n_epochs);
   unsigned long long cnt[7];
   KB2E_CUDA(c, cudaMemcpyAsync(fixed.data(), c->loss_dev, fixed.size() * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(cnt, c->counters, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   c->stamp_base += (uint32_t)n_batches;
   if (cnt[6]) return fail(c, KB2E_ERR_LIMIT, "internal: a touched-row list overflowed (tables may be inconsistent)");
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   c->tstats.kernel_ms += ms;
   c->tstats.launches += 1;
   c->tstats.samples += (uint64_t)K * (uint64_t)a.batchsize * (uint64_t)a.batches * (uint64_t)n_epochs;
   c->tstats.active = cnt[0];
   c->tstats.touched_ent = cnt[1];
   c->tstats.touched_rel = cnt[2];
   if (loss_out)
      for (size_t i = 0; i < fixed.size(); i++) loss_out[i] = (double)fixed[i] / kSweepLossScale;
   for (int t = 0; t < 3; t++) c->v64[t] = false;
   c->tables_epoch++;
   return KB2E_OK;
}

int train_run(kb2e_ctx* c, int first_epoch, int n_epochs, const int32_t* pairs_dev, int64_t n_pairs, double* loss_out, bool phase1_only) {
   if (c->K > 1) {
      if (pairs_dev || phase1_only) return fail(c, KB2E_ERR_ARG, "the batch test hooks address one model: not available after kb2e_set_replicas(K > 1)");
      if (n_epochs <= 0) return KB2E_OK;
      return sweep_run(c, first_epoch, n_epochs, loss_out);
   }
   { int rc = ensure32(c); if (rc) return rc; }
   TrainArgs a;
   if (n_epochs <= 0) return KB2E_OK;
   if (n_epochs > c->loss_cap) {
      pool_free(c, c->loss_dev);
      c->loss_dev = nullptr;
      c->loss_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &c->loss_dev, (size_t)n_epochs * sizeof(double)));
      c->loss_cap = n_epochs;
   }
   fill_args(c, a);
   a.first_epoch = first_epoch;
   a.n_epochs = n_epochs;
   a.phase1_only = phase1_only ? 1 : 0;
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
   if (!pairs_dev && !c->thr_valid) {
      fill_threshold_kernel<<<blocks_for(c->n_train, 256), 256, 0, c->stream>>>(c->triples, c->n_train, c->pr);
      KB2E_CUDA(c, cudaGetLastError());
      c->thr_valid = true;
   }
   int lps, nv, threads;
   choose_shape(c, a.batchsize, lps, nv, threads);
   // two vectors per lane: 640 threads (102 registers, nothing spills) when that many still take the batch in one pass
   // and the one-barrier kernel (which has no 640-thread form) is not the better choice
   if (c->cfg.model == KB2E_MODEL_TRANSE && nv == 2 && (long long)c->num_sms * 640 / lps >= a.batchsize &&
       !(!phase1_only && train_fused_wanted(c, a.batchsize, lps, threads)) && !getenv("KB2E_TRAIN_NO640"))
      threads = 640;
   if (nv > 4) return fail(c, KB2E_ERR_LIMIT, "embedding size above 512 is not supported by the training kernel");
   TrainKernel k = nullptr;
   const bool transr = c->cfg.model == KB2E_MODEL_TRANSR;   // own kernel: train_transr.cu
   size_t list_bytes = 0;
   if (transr) {
      if (c->P > 128) return fail(c, KB2E_ERR_LIMIT, "TransR training supports embedding sizes up to 128");
      if (c->cfg.flags & KB2E_FLAG_DETERMINISTIC) return fail(c, KB2E_ERR_LIMIT, "KB2E_FLAG_DETERMINISTIC covers TransE and TransH");
   } else {
      list_bytes = list_shape(c, a.batchsize, threads / lps, a.cap_ent, a.cap_rel);
      const bool det = (c->cfg.flags & KB2E_FLAG_DETERMINISTIC) != 0u;
      if (det && !list_bytes)
         return fail(c, KB2E_ERR_LIMIT, "KB2E_FLAG_DETERMINISTIC is built for the list kernels: the batch is too large for this embedding size");
      if (c->cfg.model == KB2E_MODEL_TRANSE)
         k = !list_bytes ? pick_kernel<KB2E_MODEL_TRANSE, false, false>(lps, nv, threads)
                         : (det ? pick_kernel<KB2E_MODEL_TRANSE, true, true>(lps, nv, threads) : pick_kernel<KB2E_MODEL_TRANSE, true, false>(lps, nv, threads));
      else
         k = !list_bytes ? pick_kernel<KB2E_MODEL_TRANSH, false, false>(lps, nv, threads)
                         : (det ? pick_kernel<KB2E_MODEL_TRANSH, true, true>(lps, nv, threads) : pick_kernel<KB2E_MODEL_TRANSH, true, false>(lps, nv, threads));
      // few relations (WN18: 18): relation-side rows resident in every CTA's shared memory, two barriers per batch
      const char* sr_env = getenv("KB2E_TRANSH_SR");   // 0: off (A-B measurements, the bit-identity test)
      if (c->cfg.model == KB2E_MODEL_TRANSH && list_bytes && c->nR <= kSmallRelations && !(sr_env && atoi(sr_env) == 0)) {
         TrainKernel ksr = det ? pick_transh_sr<true>(lps, nv) : pick_transh_sr<false>(lps, nv);
         const size_t sr_bytes = (size_t)(4 + ((a.cap_ent + 3) & ~3)) * sizeof(int) + 2 * (size_t)c->nR * c->P * sizeof(float);
         if (ksr && sr_bytes <= 48 * 1024) {
            if (!c->relbuf1) {
               KB2E_CUDA(c, pool_alloc(c, &c->relbuf1, 2 * (size_t)c->nR * c->P * sizeof(float)));
               KB2E_CUDA(c, cudaMemsetAsync(c->relbuf1, 0, 2 * (size_t)c->nR * c->P * sizeof(float), c->stream));
            }
            a.relbuf1 = c->relbuf1;
            a.cap_rel = 0;
            list_bytes = sr_bytes;
            k = ksr;
         }
      }
   }
   if (!k && !transr) return fail(c, KB2E_ERR_LIMIT, "no training kernel for this embedding size");
   // row stamps of this launch: stamp_base + 1 ... stamp_base + #batches, never reused by a later launch
   const uint64_t n_batches = (uint64_t)a.batches * (uint64_t)n_epochs;
   if (n_batches >= (1ull << 31)) return fail(c, KB2E_ERR_LIMIT, "more than 2^31 batches in one launch");
   if ((uint64_t)c->stamp_base + n_batches + 2 >= (1ull << 32)) {
      // the stamp counter starts over: every array that holds stamps is cleared with it (a pending TransH carry mark is
      // dropped -- its delta stays in dw and is applied the next time a sample touches the relation)
      KB2E_CUDA(c, cudaMemsetAsync(c->flag, 0, ((size_t)c->nE + c->nR) * sizeof(uint32_t), c->stream));
      KB2E_CUDA(c, cudaMemsetAsync(c->cflag, 0, (size_t)c->nR * sizeof(uint32_t), c->stream));
      if (c->transr_aux) KB2E_CUDA(c, cudaMemsetAsync(c->transr_aux, 0, (2 * (size_t)c->nE + (size_t)c->nR + 2) * sizeof(uint32_t), c->stream));
      c->stamp_base = 0;
   }
   a.stamp_base = c->stamp_base;
   // batches that need at most half of the resident groups: one barrier per batch, every row folded where its last
   // sample finishes (train_fused.cu)
   const bool fused = !transr && !phase1_only && !(c->cfg.flags & KB2E_FLAG_DETERMINISTIC) && train_fused_wanted(c, a.batchsize, lps, threads);
   if (!transr) {
      int per_sm = 0;
      KB2E_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, list_bytes));
      if (per_sm < 1) return fail(c, KB2E_ERR_CUDA, "training kernel does not fit on an SM");
   }
   KB2E_CUDA(c, cudaMemsetAsync(c->barrier, 0, 64, c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->loss_dev, 0, (size_t)n_epochs * sizeof(double), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->counters + 6, 0, sizeof(unsigned long long), c->stream));
   const char* trace_path = getenv("KB2E_TRAIN_TRACE");
   unsigned long long* trace_dev = nullptr;
   if (trace_path) {
      if (getenv("KB2E_TRAIN_TRACE_FINE")) a.flags |= 0x80000000u;   // + per-group stamps inside the list publish (11 slots per batch)
      KB2E_CUDA(c, pool_alloc(c, &trace_dev, (size_t)c->num_sms * kTraceSlots * sizeof(unsigned long long)));
      KB2E_CUDA(c, cudaMemsetAsync(trace_dev, 0, (size_t)c->num_sms * kTraceSlots * sizeof(unsigned long long), c->stream));
      a.trace = trace_dev;
   }
   const bool transr_stats = transr && getenv("KB2E_TRANSR_STATS") != nullptr;   // tuning aid: constraint-loop pass counts on stderr
   if (transr_stats) {
      a.flags |= 0x80000000u;
      KB2E_CUDA(c, cudaMemsetAsync(c->counters + 3, 0, 3 * sizeof(unsigned long long), c->stream));
   }
   void* params[] = {&a};
   KB2E_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   if (transr) {
      int rc = train_transr_launch(c, a, &threads);
      if (rc) return rc;
   } else {
      int rc = fused ? train_fused_launch(c, a, lps, nv, threads) : KB2E_ERR_LIMIT;
      if (rc == KB2E_ERR_LIMIT) {   // not a small batch, or no fused instantiation for this shape: the two-barrier kernel
         KB2E_CUDA(c, cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(threads), params, list_bytes, c->stream));
      } else if (rc) {
         return rc;
      }
   }
   KB2E_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   std::vector<double> loss(n_epochs);
   unsigned long long cnt[7];
   KB2E_CUDA(c, cudaMemcpyAsync(loss.data(), c->loss_dev, (size_t)n_epochs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(cnt, c->counters, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   c->stamp_base += (uint32_t)n_batches;
   if (cnt[6]) return fail(c, KB2E_ERR_LIMIT, "internal: a touched-row list overflowed (tables may be inconsistent)");
   if (transr_stats)
      fprintf(stderr, "transRNorm over %d batches: %llu violating calls, %llu passes, longest call %llu passes; %llu entity rows\n", (int)n_batches,
              cnt[4], cnt[3], cnt[5], cnt[1]);
   if (trace_dev) {
      std::vector<unsigned long long> tr((size_t)c->num_sms * kTraceSlots);
      cudaMemcpy(tr.data(), trace_dev, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      pool_free(c, trace_dev);
      if (FILE* f = fopen(trace_path, "w")) {
         for (int b = 0; b < c->num_sms; b++) {
            for (int k = 0; k < kTraceSlots; k++) fprintf(f, "%llu%c", tr[(size_t)b * kTraceSlots + k], k + 1 == kTraceSlots ? '\n' : ' ');
         }
         fclose(f);
      }
   }
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   c->tstats.kernel_ms += ms;
   c->tstats.launches += 1;
   c->tstats.samples += (uint64_t)a.batchsize * (uint64_t)a.batches * (uint64_t)n_epochs;
   c->tstats.active = cnt[0];
   c->tstats.touched_ent = cnt[1];
   c->tstats.touched_rel = cnt[2];
   if (c->cfg.flags & KB2E_FLAG_DETERMINISTIC) {
      for (double& l : loss) {
         long long fixed;
         memcpy(&fixed, &l, sizeof(fixed));
         l = (double)fixed / kDetLossScale;
      }
   }
   if (loss_out) memcpy(loss_out, loss.data(), (size_t)n_epochs * sizeof(double));
   if (!phase1_only) {
      for (int t = 0; t < 3; t++) c->v64[t] = false;
      c->tables_epoch++;
   }
   return KB2E_OK;
}

// Test hook behind kb2e_train_batch_deltas: the accumulated (pre-normalisation) updates of the last phase1_only
// launch, widened to fp64, then the delta tables, the stamps and the relation ranges go back to their idle state.
__global__ void widen_delta_kernel(float* src, double* dst, long long rows, int D, int P, int det) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= rows * P) return;
   const int col = (int)(i % P);
   if (col < D) dst[(i / P) * D + col] = det ? (double)__float_as_int(src[i]) / (double)kDetScale : (double)src[i];
   src[i] = 0.f;
}

int train_take_deltas(kb2e_ctx* c, double* d_ent, double* d_rel, double* d_w) {
   const long long we = c->cfg.model == KB2E_MODEL_TRANSE ? 0 : (c->cfg.model == KB2E_MODEL_TRANSH ? (long long)c->nR : (long long)c->nR * c->D);
   const long long rows[3] = {c->nE, c->nR, we};
   float* src[3] = {c->dtab, c->dtab + (size_t)c->nE * c->P, c->dw};
   double* dst[3] = {d_ent, d_rel, d_w};
   long long most = std::max(rows[0], std::max(rows[1], rows[2]));
   double* dev = nullptr;
   KB2E_CUDA(c, pool_alloc(c, &dev, (size_t)most * c->D * sizeof(double)));
   int rc = KB2E_OK;
   for (int t = 0; t < 3 && rc == KB2E_OK; t++) {
      if (rows[t] == 0) continue;
      widen_delta_kernel<<<blocks_for(rows[t] * c->P, 256), 256, 0, c->stream>>>(src[t], dev, rows[t], c->D, c->P,
                                                                                  (c->cfg.flags & KB2E_FLAG_DETERMINISTIC) ? 1 : 0);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess && dst[t]) e = cudaMemcpyAsync(dst[t], dev, (size_t)rows[t] * c->D * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = cuda_fail(c, e, "kb2e_train_batch_deltas copy");
   }
   pool_free(c, dev);
   if (rc) return rc;
   KB2E_CUDA(c, cudaMemsetAsync(c->flag, 0, ((size_t)c->nE + c->nR) * sizeof(uint32_t), c->stream));
   if (c->rmin) {
      fill_int_kernel<<<blocks_for(c->nE, 256), 256, 0, c->stream>>>(c->rmin, c->nE, 0x7fffffff);
      fill_int_kernel<<<blocks_for(c->nE, 256), 256, 0, c->stream>>>(c->rmax, c->nE, -1);
   }
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   return KB2E_OK;
}

int train_sample(kb2e_ctx* c, int epoch, int batch, int64_t count, int32_t* out_dev) {
   if (!c->triples || !c->have_pr) return fail(c, KB2E_ERR_ARG, "sampler needs kb2e_set_train_triples and kb2e_set_bern");
   if (!c->thr_valid) {
      fill_threshold_kernel<<<blocks_for(c->n_train, 256), 256, 0, c->stream>>>(c->triples, c->n_train, c->pr);
      KB2E_CUDA(c, cudaGetLastError());
      c->thr_valid = true;
   }
   TrainArgs a;
   fill_args(c, a);
   uint32_t gb = (uint32_t)epoch * (uint32_t)c->cfg.batches + (uint32_t)batch;
   sample_kernel<<<blocks_for(count, 256), 256, 0, c->stream>>>(a, gb, count, out_dev);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int train_score32(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n, double* out) {
   { int rc = ensure32(c); if (rc) return rc; }
   unsigned blocks = blocks_for(n * 32, 256);
   if (c->cfg.model == KB2E_MODEL_TRANSR) {
      if (c->P > 128) return fail(c, KB2E_ERR_LIMIT, "TransR fp32 scoring supports embedding sizes up to 128");
      score32_transr_kernel<<<blocks, 256, 0, c->stream>>>(c->tab, c->w, c->w_row, c->nE, c->D, c->P, c->cfg.distance, h, t, r, n, out);
   } else if (c->cfg.model == KB2E_MODEL_TRANSE)
      score32_kernel<KB2E_MODEL_TRANSE><<<blocks, 256, 0, c->stream>>>(c->tab, c->w, c->nE, c->D, c->P, c->cfg.distance, h, t, r, n, out);
   else
      score32_kernel<KB2E_MODEL_TRANSH><<<blocks, 256, 0, c->stream>>>(c->tab, c->w, c->nE, c->D, c->P, c->cfg.distance, h, t, r, n, out);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

}  // namespace kb2e
