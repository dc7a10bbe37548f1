// TransR epochs (transr/transr.cpp:13-37, transr/trainer.cpp:35-64,144-188) as one persistent cooperative
// launch: the same batch semantics and phases as train_kernel (train.cu), but a work decomposition built
// around the D x D projection matrix.
//
// One warp owns one sample (phase 1) or one touched row (phase 2) and stages the M_r it needs ONCE into its
// own slice of shared memory with coalesced 16-byte loads (all of them in flight together), then runs both
// contractions from there:
//   forward   y_i = sum_j M[j][i] e_j     lane owns output dims i = lane, lane + 32, ...; row j of the staged
//                                          matrix is read with consecutive lanes on consecutive banks and e_j is
//                                          a shared-memory broadcast
//   backward  S_j = sum_i g_i M[j][i]     lane owns input dims j and reads its row with 16-byte shared loads; the
//                                          row pitch is a multiple of 4 floats with pitch/4 odd, which puts the
//                                          8 rows of a quarter-warp on disjoint banks -- no shuffles
//   transRNorm sweep (sequential in i)    four columns per step: one 16-byte shared load per lane-owned row, four
//                                          interleaved warp sums, and a scalar recurrence on the chunk's 4x4 Gram
//                                          block for the dependence between the four steps (transr_constraint)
// Phase 2 work (touched relations / entity rows) is compacted into shared-memory lists and dealt dynamically.
// Staging uses cp.async (16-byte, L2-only) so that all of M_r is in flight at once without holding registers.
// The dM update is issued as flat, fully coalesced vector REDs over the D*P/4 float4 of M_r.
// Everything else (sampler, stamps, deferred renormalisation, the transr/trainer.cpp:187 quirk, carrying the
// constraint's perturbation of M_r into the next batch) is exactly as in train.cu / oracle orc_train_batch_dfr.

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "train_device.cuh"

namespace kb2e {

constexpr int kRMaxThreads = 576;   // 18 warps: 65536 / 576 = 113 registers per thread
enum { S_H = 0, S_T, S_C, S_R, S_GP, S_GN, S_DP, S_DN, S_SP, S_SN, kRSlots };

struct RArgs {
   TrainArgs base;
   int pitch;          // row pitch of the staged matrix in floats (multiple of 4, pitch / 4 odd)
   int vec_off;        // offset of the vector slots inside a warp's slice (floats, multiple of 4)
   int warp_floats;    // floats per warp slice (multiple of 4)
   uint32_t p4_magic;  // ceil(2^32 / (P / 4)): f / (P/4) == umulhi(f, magic) for the flat indices used here
   // phase 2b work list: every entity row a batch must visit, once, in one grid-wide array (built in phase 1)
   uint32_t* claim;    // [nE] stamp of the batch the row was last put on the list for
   uint32_t* qflag;    // [nR] stamp of the batch relation r was last touched by a SAMPLE (flag[nE + r] also moves on constraint carries)
   uint32_t* gcount;   // [2] list length, indexed by stamp parity (the other one is zeroed during the batch)
   int* glist;         // [nE]
};

__device__ __forceinline__ float rcomp(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
   asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stage M (global, [D][P]) into sM ([D][pitch]) with asynchronous 16-byte copies; the caller waits
// (cp_async_wait_all + __syncwarp) before reading.
__device__ __forceinline__ void stage_matrix(const RArgs& a, const float* M, float* sM, int lane) {
   const int P4 = a.base.P >> 2;
   const int nF = a.base.D * P4;
   if (a.pitch == a.base.P) {
      for (int f = lane; f < nF; f += 32) cp_async16(sM + 4 * f, M + 4 * (size_t)f);
   } else {
      for (int f = lane; f < nF; f += 32) {
         const int j = (int)__umulhi((uint32_t)f, a.p4_magic);
         cp_async16(sM + j * a.pitch + 4 * (f - j * P4), M + 4 * (size_t)f);
      }
   }
}

// y[v][k] = sum_j sM[j][lane + 32 k] * vec_v[j]   (transr/transr.cpp:20-25 with zeroed work vectors)
template <int NE, int NVEC>
__device__ __forceinline__ void project_smem(const float* sM, int D, int pitch, int lane, const float* const (&vec)[NVEC],
                                             float (&y)[NVEC][NE]) {
#pragma unroll
   for (int v = 0; v < NVEC; v++)
#pragma unroll
      for (int k = 0; k < NE; k++) y[v][k] = 0.f;
   for (int j0 = 0; j0 < D; j0 += 4) {
      float4 e[NVEC];
#pragma unroll
      for (int v = 0; v < NVEC; v++) e[v] = *reinterpret_cast<const float4*>(vec[v] + j0);
#pragma unroll
      for (int c = 0; c < 4; c++) {
         const int j = j0 + c;
         if (j < D) {
#pragma unroll
            for (int k = 0; k < NE; k++) {
               const int i = lane + 32 * k;
               const float m = i < D ? sM[j * pitch + i] : 0.f;
#pragma unroll
               for (int v = 0; v < NVEC; v++) y[v][k] = fmaf(rcomp(e[v], c), m, y[v][k]);
            }
         }
      }
   }
}

// lane-owned elements (i = lane + 32 k) -> slot; the padding [D, P) is written as zero
template <int NE>
__device__ __forceinline__ void put_slot(float* slot, int D, int P, int lane, const float (&v)[NE]) {
#pragma unroll
   for (int k = 0; k < NE; k++) {
      const int i = lane + 32 * k;
      if (i < P) slot[i] = i < D ? v[k] : 0.f;
   }
}
template <int NE>
__device__ __forceinline__ void get_slot(const float* slot, int D, int lane, float (&v)[NE]) {
#pragma unroll
   for (int k = 0; k < NE; k++) {
      const int i = lane + 32 * k;
      v[k] = i < D ? slot[i] : 0.f;
   }
}

// ---- phase 1 ---------------------------------------------------------------------------------------------
template <int NE>
__device__ __forceinline__ void transr_pair(const RArgs& ra, const Pair s, float* sM, float* sV, int lane, uint32_t stamp,
                                            double& loss_acc, uint32_t& active_acc) {
   const TrainArgs& a = ra.base;
   const int P = a.P, D = a.D, P4 = P >> 2;
   const bool on = lane < P4;
   const float* M = a.w + (size_t)s.r * a.w_row;
   float* dM = a.dw + (size_t)s.r * a.w_row;
   // the four rows and the matrix: everything is in flight together, nothing passes through registers
   if (on) {
      cp_async16(sV + S_H * P + lane * 4, a.tab + (size_t)s.h * P + lane * 4);
      cp_async16(sV + S_T * P + lane * 4, a.tab + (size_t)s.t * P + lane * 4);
      cp_async16(sV + S_C * P + lane * 4, a.tab + (size_t)s.c * P + lane * 4);
      cp_async16(sV + S_R * P + lane * 4, a.tab + ((size_t)a.nE + s.r) * P + lane * 4);
   }
   stage_matrix(ra, M, sM, lane);
   cp_async_wait_all();
   __syncwarp();
   float y[3][NE];
   {
      const float* const vec[3] = {sV + S_H * P, sV + S_T * P, sV + S_C * P};
      project_smem<NE, 3>(sM, D, ra.pitch, lane, vec, y);
   }
   const bool l1 = a.distance == KB2E_DISTANCE_L1;
   float rp[NE], rn[NE], r[NE];
   get_slot<NE>(sV + S_R * P, D, lane, r);
   float ep = 0.f, en = 0.f;
#pragma unroll
   for (int k = 0; k < NE; k++) {
      rp[k] = (y[1][k] - y[0][k]) - r[k];
      rn[k] = s.corruptTail ? (y[2][k] - y[0][k]) - r[k] : (y[1][k] - y[2][k]) - r[k];
      ep += l1 ? fabsf(rp[k]) : rp[k] * rp[k];
      en += l1 ? fabsf(rn[k]) : rn[k] * rn[k];
   }
   ep = warp_sum(ep);
   en = warp_sum(en);
   if (!(ep + a.margin > en)) { __syncwarp(); return; }  // common/trainer.cpp:138
   if (lane == 0) {
      loss_acc += (double)(a.margin + ep - en);
      active_acc++;
   }
   const float lr = a.lr;
   {
      float gp[NE], gn[NE], dp[NE], dn[NE], h[NE], t[NE], c[NE];
      get_slot<NE>(sV + S_H * P, D, lane, h);
      get_slot<NE>(sV + S_T * P, D, lane, t);
      get_slot<NE>(sV + S_C * P, D, lane, c);
#pragma unroll
      for (int k = 0; k < NE; k++) {
         // transr/trainer.cpp:160-165: x = 2 * residual, L1 -> sign (0 -> -1)
         gp[k] = l1 ? (rp[k] > 0.f ? lr : -lr) : (2.f * lr) * rp[k];
         gn[k] = l1 ? (rn[k] > 0.f ? lr : -lr) : (2.f * lr) * rn[k];
         dp[k] = h[k] - t[k];                                             // h - t   (positive)
         dn[k] = (s.corruptTail ? h[k] : c[k]) - (s.corruptTail ? c[k] : t[k]);   // h' - t' (negative)
      }
      put_slot<NE>(sV + S_GP * P, D, P, lane, gp);
      put_slot<NE>(sV + S_GN * P, D, P, lane, gn);
      put_slot<NE>(sV + S_DP * P, D, P, lane, dp);
      put_slot<NE>(sV + S_DN * P, D, P, lane, dn);
   }
   __syncwarp();
   // transr/trainer.cpp:171: r -= beta*lr*x
   if (on) {
      const float4 gp4 = reinterpret_cast<const float4*>(sV + S_GP * P)[lane];
      const float4 gn4 = reinterpret_cast<const float4*>(sV + S_GN * P)[lane];
      red_add4(a.dtab + ((size_t)a.nE + s.r) * P + lane * 4, gp4 - gn4);
   }
   // transr/trainer.cpp:167: M[j][i] -= beta*lr*x_i*(h_j - t_j) -- flat over the D*P/4 vectors of M_r
   {
      const int nF = D * P4;
      for (int f = lane; f < nF; f += 32) {
         const int j = (int)__umulhi((uint32_t)f, ra.p4_magic);
         const int i4 = f - j * P4;
         const float pj = sV[S_DP * P + j], nj = sV[S_DN * P + j];
         const float4 gp4 = reinterpret_cast<const float4*>(sV + S_GP * P)[i4];
         const float4 gn4 = reinterpret_cast<const float4*>(sV + S_GN * P)[i4];
         red_add4(dM + 4 * (size_t)f, pj * gp4 - nj * gn4);
      }
   }
   // transr/trainer.cpp:168-169: e[j] -/+= beta*lr*x_i*M[j][i] summed over i; lane owns j = lane + 32 k
   {
      float sp[NE], sn[NE];
#pragma unroll
      for (int k = 0; k < NE; k++) { sp[k] = 0.f; sn[k] = 0.f; }
      for (int i0 = 0; i0 < D; i0 += 4) {
         const float4 gp4 = *reinterpret_cast<const float4*>(sV + S_GP * P + i0);
         const float4 gn4 = *reinterpret_cast<const float4*>(sV + S_GN * P + i0);
#pragma unroll
         for (int k = 0; k < NE; k++) {
            const int j = lane + 32 * k;
            if (j < D) {
               const float4 m = *reinterpret_cast<const float4*>(sM + j * ra.pitch + i0);   // i0 + 3 <= P - 1; padding columns are zero
               sp[k] += dot4(gp4, m);
               sn[k] += dot4(gn4, m);
            }
         }
      }
      put_slot<NE>(sV + S_SP * P, D, P, lane, sp);
      put_slot<NE>(sV + S_SN * P, D, P, lane, sn);
   }
   __syncwarp();
   if (on) {
      const float4 sp4 = reinterpret_cast<const float4*>(sV + S_SP * P)[lane];
      const float4 sn4 = reinterpret_cast<const float4*>(sV + S_SN * P)[lane];
      float* dh = a.dtab + (size_t)s.h * P + lane * 4;
      float* dt = a.dtab + (size_t)s.t * P + lane * 4;
      float* dc = a.dtab + (size_t)s.c * P + lane * 4;
      if (s.corruptTail) {
         red_add4(dh, sp4 - sn4);        // head is shared by both triples
         red_add4(dt, -1.f * sp4);
         red_add4(dc, sn4);
      } else {
         red_add4(dh, sp4);
         red_add4(dt, sn4 - sp4);        // tail is shared
         red_add4(dc, -1.f * sn4);
      }
   }
   int row = -1;
   if (lane < 3) {
      const int e = lane == 0 ? s.h : (lane == 1 ? s.t : s.c);
      a.flag[e] = stamp;
      atomicMin(a.rmin + e, s.r);
      atomicMax(a.rmax + e, s.r);
      row = e;
   } else if (lane == 3) {
      a.flag[(size_t)a.nE + s.r] = stamp;
      ra.qflag[s.r] = stamp;
      // transr/trainer.cpp:187: entity row r is also visited when RELATION r was touched
      if (!(a.flags & KB2E_FLAG_TRANSR_NO_QUIRK) && s.r < a.nE) row = s.r;
   }
   // first toucher of the batch appends the row to the grid-wide phase 2b list
   const bool first = row >= 0 && atomicExch(ra.claim + row, stamp) != stamp;
   const uint32_t m = __ballot_sync(0xffffffffu, first);
   if (m) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(ra.gcount + (stamp & 1u), (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (first) ra.glist[base + __popc(m & ((1u << lane) - 1u))] = row;
   }
   __syncwarp();
}

// ---- phase 2a: rows [j_begin, j_end) of relation r -- every row M_r[j][.] unit length after the delta
// (transr/trainer.cpp:178-180); the segment that starts at row 0 also publishes r itself (unit length, :174).
// A relation is cut into segments so that the few touched relations of a batch spread over all warps.
template <int NE>
__device__ __forceinline__ void transr_finish_relation(const RArgs& ra, int r, int j_begin, int j_end, float* sM, float* sV, int lane) {
   const TrainArgs& a = ra.base;
   const int P = a.P, P4 = P >> 2;
   const bool on = lane < P4;
   float* M = a.w + (size_t)r * a.w_row + (size_t)j_begin * P;
   float* dM = a.dw + (size_t)r * a.w_row + (size_t)j_begin * P;
   const int rows = j_end - j_begin;
   const int nF = rows * P4;
   float4 x = f4(0.f), d = f4(0.f);
   if (j_begin == 0 && on) {
      x = ld_cg4(a.tab + ((size_t)a.nE + r) * P + lane * 4);
      d = ld_cg4(a.dtab + ((size_t)a.nE + r) * P + lane * 4);
   }
   // M + dM -> shared, dM = 0
   constexpr int U = 4;
   for (int f0 = lane; f0 < nF; f0 += 32 * U) {
      float4 v[U], w[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
         const int f = f0 + 32 * u;
         v[u] = f < nF ? ld_cg4(M + 4 * (size_t)f) : f4(0.f);
         w[u] = f < nF ? ld_cg4(dM + 4 * (size_t)f) : f4(0.f);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
         const int f = f0 + 32 * u;
         if (f < nF) {
            const int j = (int)__umulhi((uint32_t)f, ra.p4_magic);
            *reinterpret_cast<float4*>(sM + j * ra.pitch + 4 * (f - j * P4)) = v[u] + w[u];
            st_cg4(dM + 4 * (size_t)f, f4(0.f));
         }
      }
   }
   if (j_begin == 0) {
      x = x + d;
      const float len = sqrtf(warp_sum(dot4(x, x)));
      if (on) {
         st_cg4(a.dtab + ((size_t)a.nE + r) * P + lane * 4, f4(0.f));
         const float inv = 1.f / len;   // one reciprocal, four products (norm_row, train_device.cuh)
         st_cg4(a.tab + ((size_t)a.nE + r) * P + lane * 4, make_float4(x.x * inv, x.y * inv, x.z * inv, x.w * inv));
      }
   }
   __syncwarp();
   float l[NE];
#pragma unroll
   for (int k = 0; k < NE; k++) {
      const int j = lane + 32 * k;
      float s2 = 0.f;
      if (j < rows) {
         const float4* m = reinterpret_cast<const float4*>(sM + j * ra.pitch);
         for (int i4 = 0; i4 < P4; i4++) { const float4 v = m[i4]; s2 += dot4(v, v); }   // padding columns are zero
      }
      l[k] = sqrtf(s2);
   }
   put_slot<NE>(sV + S_SP * P, rows, P, lane, l);
   __syncwarp();
   for (int f = lane; f < nF; f += 32) {
      const int j = (int)__umulhi((uint32_t)f, ra.p4_magic);
      const float4 m = *reinterpret_cast<const float4*>(sM + j * ra.pitch + 4 * (f - j * P4));
      const float lj = sV[S_SP * P + j];
      const float ij = 1.f / lj;
      st_cg4(M + 4 * (size_t)f, make_float4(m.x * ij, m.y * ij, m.z * ij, m.w * ij));
   }
   __syncwarp();
}

// transRNorm (transr/trainer.cpp:35-64) of the lane-owned entity row x (a copy lives in slot S_H) against the
// published, read-only M_r; the perturbation the reference applies to M_r goes to the NEXT batch's delta.
//
// One pass of the reference's loop is, for i = 0 .. D-1 in order (m_i = column i of M_r):
//    tmp = 2 (m_i . x);   dM[.][i] += delta, delta = -lr tmp x;   x <- x - lr tmp (m_i + delta)
// i.e. x <- x - c (m_i - c x) with c = lr tmp.  The D dependent warp-wide dot products are the critical path, so
// the pass runs in chunks of four columns: the four dots q_a = m_(i0+a) . x are reduced TOGETHER (one 16-byte
// shared load per lane-owned row, four interleaved butterflies), and the effect of step a on the later dots of
// the chunk follows from dotting the update with m_b:  q_b <- q_b - c_a (g_ba - c_a q_b),  g_ba = m_b . m_a  (six
// numbers per chunk, computed once per call when the constraint is violated at all).  Same recurrence, a
// quarter of the dependent reductions.
template <int NE>
__device__ __forceinline__ void transr_constraint(const RArgs& ra, int r, float* sM, float* sV, int lane, uint32_t next_stamp,
                                                  float (&x)[NE]) {
   const TrainArgs& a = ra.base;
   const int P = a.P, D = a.D, P4 = P >> 2, pitch = ra.pitch;
   const float* M = a.w + (size_t)r * a.w_row;
   float* dM = a.dw + (size_t)r * a.w_row;
   float* sG = sV + S_GP * P;   // 6 * P4 floats (slots S_GP, S_GN): g10 g20 g21 g30 g31 g32 per chunk
   stage_matrix(ra, M, sM, lane);
   cp_async_wait_all();
   __syncwarp();
   bool have_gram = false;
   int passes = 0;
   for (int iter = 0; iter < 64; iter++) {
      float y[1][NE];
      const float* const vec[1] = {sV + S_H * P};
      project_smem<NE, 1>(sM, D, pitch, lane, vec, y);
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < NE; k++) n2 += (lane + 32 * k < D) ? y[0][k] * y[0][k] : 0.f;
      if (warp_sum(n2) <= 1.f) break;
      if (!have_gram) {
         for (int t = lane; t < 6 * P4; t += 32) {
            const int chunk = t / 6, pr = t - 6 * chunk;
            const int ca = pr == 0 ? 0 : (pr < 3 ? pr - 1 : pr - 3);   // pairs (1,0) (2,0) (2,1) (3,0) (3,1) (3,2)
            const int cb = pr == 0 ? 1 : (pr < 3 ? 2 : 3);
            const float* col = sM + 4 * chunk;
            float g = 0.f;
            for (int j = 0; j < D; j++) g = fmaf(col[j * pitch + ca], col[j * pitch + cb], g);
            sG[t] = g;
         }
         have_gram = true;
         __syncwarp();
      }
      for (int c4 = 0; c4 < P4; c4++) {
         float4 mk[NE];
         float4 q = f4(0.f);
#pragma unroll
         for (int k = 0; k < NE; k++) {
            const int j = lane + 32 * k;
            mk[k] = j < D ? *reinterpret_cast<const float4*>(sM + j * pitch + 4 * c4) : f4(0.f);
            q.x = fmaf(mk[k].x, x[k], q.x); q.y = fmaf(mk[k].y, x[k], q.y);
            q.z = fmaf(mk[k].z, x[k], q.z); q.w = fmaf(mk[k].w, x[k], q.w);
         }
#pragma unroll
         for (int o = 16; o > 0; o >>= 1) {
            q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
            q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
         }
         const float* g = sG + 6 * c4;
         const float c0 = a.lr * (2.f * q.x);
         float q1 = q.y - c0 * (g[0] - c0 * q.y);
         float q2 = q.z - c0 * (g[1] - c0 * q.z);
         float q3 = q.w - c0 * (g[3] - c0 * q.w);
         const float c1 = a.lr * (2.f * q1);
         q2 = q2 - c1 * (g[2] - c1 * q2);
         q3 = q3 - c1 * (g[4] - c1 * q3);
         const float c2 = a.lr * (2.f * q2);
         q3 = q3 - c2 * (g[5] - c2 * q3);
         const float c3 = a.lr * (2.f * q3);
#pragma unroll
         for (int k = 0; k < NE; k++) {
            const int j = lane + 32 * k;
            if (j < D) {
               float xv = x[k];
               float4 dl;
               dl.x = -(c0 * xv); xv = xv - c0 * (mk[k].x + dl.x);
               dl.y = -(c1 * xv); xv = xv - c1 * (mk[k].y + dl.y);
               dl.z = -(c2 * xv); xv = xv - c2 * (mk[k].z + dl.z);
               dl.w = -(c3 * xv); xv = xv - c3 * (mk[k].w + dl.w);
               x[k] = xv;
               red_add4(dM + (size_t)j * P + 4 * c4, dl);   // padding columns: m = 0 -> q = c = delta = 0
            }
         }
      }
      __syncwarp();
      put_slot<NE>(sV + S_H * P, D, P, lane, x);
      __syncwarp();
      passes++;
   }
   if (have_gram && lane == 0) {
      a.flag[(size_t)a.nE + r] = next_stamp;
      if (a.flags & 0x80000000u) {   // KB2E_TRANSR_STATS: passes of the constraint loop (sum, violating calls, longest call)
         atomicAdd(a.counters + 3, (unsigned long long)passes);
         atomicAdd(a.counters + 4, 1ull);
         atomicMax(a.counters + 5, (unsigned long long)passes);
      }
   }
   __syncwarp();
}

// ---- phase 2b: entity row -- unit length (transr/trainer.cpp:175-176), then transRNorm against the lowest /
// highest relation that touched it, and -- the reference's quirk at :187 -- against M_e when relation e was touched.
template <int NE>
__device__ __forceinline__ void transr_finish_entity(const RArgs& ra, int e, float* sM, float* sV, int lane, uint32_t stamp,
                                                     uint32_t next_stamp, bool own) {
   const TrainArgs& a = ra.base;
   const int P = a.P, D = a.D;
   const bool on = lane < (P >> 2);
   float4 x4 = on ? ld_cg4(a.tab + (size_t)e * P + lane * 4) : f4(0.f);
   int r0 = 0x7fffffff, r1 = -1;
   if (own) {
      const float4 d4 = on ? ld_cg4(a.dtab + (size_t)e * P + lane * 4) : f4(0.f);
      r0 = __ldcg(a.rmin + e);
      r1 = __ldcg(a.rmax + e);
      x4 = x4 + d4;
      const float len = sqrtf(warp_sum(dot4(x4, x4)));
      const float inv = 1.f / len;
      x4 = make_float4(x4.x * inv, x4.y * inv, x4.z * inv, x4.w * inv);
      if (on) st_cg4(a.dtab + (size_t)e * P + lane * 4, f4(0.f));
      if (lane == 0) { a.rmin[e] = 0x7fffffff; a.rmax[e] = -1; }
   }
   if (on) reinterpret_cast<float4*>(sV + S_H * P)[lane] = x4;
   __syncwarp();
   float x[NE];
   get_slot<NE>(sV + S_H * P, D, lane, x);
   if (own && r1 >= 0) {
      transr_constraint<NE>(ra, r0, sM, sV, lane, next_stamp, x);
      if (r1 != r0) transr_constraint<NE>(ra, r1, sM, sV, lane, next_stamp, x);
   }
   const bool quirk = !(a.flags & KB2E_FLAG_TRANSR_NO_QUIRK) && e < a.nR && __ldcg(ra.qflag + e) == stamp;
   if (quirk && !(own && (r0 == e || r1 == e))) transr_constraint<NE>(ra, e, sM, sV, lane, next_stamp, x);
   // slot S_H holds the final row (transr_constraint keeps it current)
   if (on) st_cg4(a.tab + (size_t)e * P + lane * 4, reinterpret_cast<const float4*>(sV + S_H * P)[lane]);
   __syncwarp();
}

// ---- touched-row lists ---------------------------------------------------------------------------------
// Phase 2 work is a small, irregular subset of the rows, and one row costs microseconds (it stages 10 KB
// matrices), so static row ranges leave most warps idle behind the unlucky ones.  Instead the CTA compacts
// the stamps of a window of rows into an ordered list in shared memory (ballot + warp-count prefix) and its
// warps take list entries one at a time.
constexpr int kListCap = 1536;   // FB15k: 1,345 relations in one window; 18 warps of D = 50 still fit next to it

struct RowList {
   int n;
   int next;
   int wcnt[kRMaxThreads / 32];
   int item[kListCap];
};

// rows [begin, end) (end - begin <= kListCap) for which pred holds -> list.item[0 .. list.n), ascending
template <typename Pred>
__device__ __forceinline__ void compact_rows(RowList& list, long long begin, long long end, Pred&& pred) {
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
   if (threadIdx.x == 0) { list.n = 0; list.next = 0; }
   __syncthreads();
   constexpr int U = 4;   // stamps of four strides are loaded together: one L2 round trip instead of four
   for (long long base0 = begin; base0 < end; base0 += (long long)U * blockDim.x) {
      bool fs[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
         const long long r = base0 + (long long)u * blockDim.x + threadIdx.x;
         fs[u] = r < end && pred(r);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
         const long long base = base0 + (long long)u * blockDim.x;
         if (base >= end) break;
         const long long r = base + threadIdx.x;
         const bool f = fs[u];
         const uint32_t m = __ballot_sync(0xffffffffu, f);
         if (lane == 0) list.wcnt[warp] = __popc(m);
         __syncthreads();
         int off = list.n, total = 0;
         for (int w = 0; w < warps; w++) {
            const int c = list.wcnt[w];
            if (w < warp) off += c;
            total += c;
         }
         if (f) list.item[off + __popc(m & ((1u << lane) - 1u))] = (int)(r - begin);
         __syncthreads();
         if (threadIdx.x == 0) list.n += total;
         __syncthreads();
      }
   }
}

template <int NE>
__global__ void __launch_bounds__(kRMaxThreads, 1) train_transr_kernel(const __grid_constant__ RArgs ra) {
   extern __shared__ float4 smem4[];
   __shared__ double s_loss[kRMaxThreads / 32];
   __shared__ RowList s_list;
   const TrainArgs& a = ra.base;
   const int lane = threadIdx.x & 31;
   const int warp = threadIdx.x >> 5;
   float* sM = reinterpret_cast<float*>(smem4) + (size_t)warp * ra.warp_floats;
   float* sV = sM + ra.vec_off;
   const int warps = blockDim.x >> 5;
   const long long G = (long long)gridDim.x * warps;
   const long long g0 = (long long)warp * gridDim.x + blockIdx.x;   // round-robin over CTAs
   uint32_t bar_target = 0;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   int trace_slot = 0;
#define KB2E_RTRACE()                                                                                 \
   if (a.trace != nullptr && threadIdx.x == 0 && trace_slot < kTraceSlots) {                          \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      a.trace[(size_t)blockIdx.x * kTraceSlots + trace_slot++] = t_;                                  \
   }
   const uint32_t gb_first = (uint32_t)a.first_epoch * (uint32_t)a.batches;
   Pair pre;
   const bool has_first = g0 < a.batchsize;
   if (has_first) pre = draw_pair(a, (uint32_t)g0, gb_first);

   for (int ep = 0; ep < a.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < a.batches; batch++) {
         const uint32_t gb = gb_first + (uint32_t)ep * (uint32_t)a.batches + (uint32_t)batch;
         // stamps count the batches this context has run (train.cu), so a stamp never recurs
         const uint32_t stamp = a.stamp_base + (uint32_t)ep * (uint32_t)a.batches + (uint32_t)batch + 1u;
         const uint32_t next_stamp = stamp + 1u;
         KB2E_RTRACE();
         if (has_first) transr_pair<NE>(ra, pre, sM, sV, lane, stamp, loss_acc, active_acc);
         for (long long k = g0 + G; k < a.batchsize; k += G) {
            Pair s = draw_pair(a, (uint32_t)k, gb);
            transr_pair<NE>(ra, s, sM, sV, lane, stamp, loss_acc, active_acc);
         }
         KB2E_RTRACE();
         grid_barrier(a.barrier, bar_target);
         KB2E_RTRACE();
         if (a.phase1_only) continue;   // kb2e_train_batch_deltas: the caller reads the raw delta tables (one batch per launch)
         // ---- phase 2a: every CTA builds the same list of touched relations; entry t is cut into S row segments
         // and (entry, segment) pairs are dealt round-robin over all warps of the grid
         for (long long w0 = 0; w0 < a.nR; w0 += kListCap) {
            const long long w1 = min((long long)a.nR, w0 + kListCap);
            compact_rows(s_list, w0, w1, [&](long long r) { return __ldcg(a.flag + a.nE + r) == stamp; });
            const int n = s_list.n;
            if (n > 0) {
               int S = (int)min((long long)8, max((long long)1, G / n));
               S = min(S, max(1, a.D / 4));
               const int rows_per = (a.D + S - 1) / S;
               for (long long it = g0; it < (long long)n * S; it += G) {
                  const int e = (int)(it / S), seg = (int)(it - (long long)e * S);
                  const int jb = seg * rows_per, je = min(a.D, jb + rows_per);
                  if (jb < je) transr_finish_relation<NE>(ra, (int)w0 + s_list.item[e], jb, je, sM, sV, lane);
                  trel_acc += (lane == 0 && seg == 0);
               }
            }
            __syncthreads();
         }
         KB2E_RTRACE();
         grid_barrier(a.barrier, bar_target);
         KB2E_RTRACE();
         // ---- phase 2b: the batch's list of entity rows (touched, or visited through the quirk), dealt round-robin over
         // all warps of the grid: a row costs microseconds (it stages 10 KB matrices), so the count per warp is what matters
         {
            if (blockIdx.x == 0 && threadIdx.x == 0) ra.gcount[next_stamp & 1u] = 0u;   // last read in the previous batch's phase 2b
            const long long n = __ldcg(ra.gcount + (stamp & 1u));
            for (long long it = g0; it < n; it += G) {
               const int e = __ldcg(ra.glist + it);
               const bool own = __ldcg(a.flag + e) == stamp;
               transr_finish_entity<NE>(ra, e, sM, sV, lane, stamp, next_stamp, own);
               tent_acc += (lane == 0 && own);
            }
         }
         KB2E_RTRACE();
         grid_arrive(a.barrier, bar_target);
         if (has_first && !(ep == a.n_epochs - 1 && batch == a.batches - 1)) pre = draw_pair(a, (uint32_t)g0, gb + 1u);
         grid_wait(a.barrier, bar_target);
      }
      double v = (lane == 0) ? loss_acc : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_loss[warp] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
         double t = 0.0;
         for (int i = 0; i < warps; i++) t += s_loss[i];
         if (t != 0.0) atomicAdd(a.loss + ep, t);
      }
      __syncthreads();
   }
   uint32_t c0 = (lane == 0) ? active_acc : 0u, c1 = tent_acc, c2 = trel_acc;
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

// ---- host ------------------------------------------------------------------------------------------------
int train_transr_launch(kb2e_ctx* c, const TrainArgs& base, int* threads_out) {
   if (c->P > 128) return fail(c, KB2E_ERR_LIMIT, "TransR training supports embedding sizes up to 128");
   RArgs a;
   a.base = base;
   a.pitch = ((c->P / 4) & 1) ? c->P : c->P + 4;   // multiple of 4 floats with pitch/4 odd (bank-conflict-free 16-byte row reads)
   a.vec_off = c->D * a.pitch;
   a.warp_floats = a.vec_off + kRSlots * c->P;
   const int P4 = c->P / 4;
   a.p4_magic = (uint32_t)(((1ull << 32) + P4 - 1) / P4);
   if (!c->transr_aux) {   // claim [nE] | qflag [nR] | glist [nE] | gcount [2]; stamps never recur, so zero once is enough
      const size_t words = 2 * (size_t)c->nE + (size_t)c->nR + 2;
      KB2E_CUDA(c, pool_alloc(c, &c->transr_aux, words * sizeof(uint32_t)));
      KB2E_CUDA(c, cudaMemsetAsync(c->transr_aux, 0, words * sizeof(uint32_t), c->stream));
   }
   a.claim = c->transr_aux;
   a.qflag = a.claim + c->nE;
   a.glist = reinterpret_cast<int*>(a.qflag + c->nR);
   a.gcount = reinterpret_cast<uint32_t*>(a.glist + c->nE);
   KB2E_CUDA(c, cudaMemsetAsync(a.gcount, 0, 2 * sizeof(uint32_t), c->stream));
   int dev_smem = 0;
   KB2E_CUDA(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
   const size_t per_warp = (size_t)a.warp_floats * sizeof(float);
   int warps = (int)std::min<size_t>(kRMaxThreads / 32, ((size_t)dev_smem - sizeof(RowList) - 512) / per_warp);
   if (const char* env = getenv("KB2E_TRANSR_WARPS")) warps = std::max(1, std::min(warps, atoi(env)));
   if (warps < 2) return fail(c, KB2E_ERR_LIMIT, "TransR projection matrix does not fit in shared memory");
   const int ne = (c->D + 31) / 32;
   void (*k)(const RArgs) = ne == 1 ? train_transr_kernel<1> : (ne == 2 ? train_transr_kernel<2> : (ne == 3 ? train_transr_kernel<3> : train_transr_kernel<4>));
   const size_t smem = per_warp * warps;
   KB2E_CUDA(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   void* params[] = {&a};
   *threads_out = warps * 32;
   KB2E_CUDA(c, cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(warps * 32), params, smem, c->stream));
   return KB2E_OK;
}

}  // namespace kb2e
