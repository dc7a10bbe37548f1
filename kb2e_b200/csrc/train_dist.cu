// Entity-table-partitioned TransE training over the GPUs of one NVLink/NVSwitch box (BASELINE
// configs[4]: 4 M entities x D=200, 100 M triples; SURVEY.md 8e "Training, scaled shape").
//
// One process (context) per GPU.  Entity row e lives on GPU e % G at local index e / G; the relation
// table is small and replicated.  Every rank maps every peer's arena with CUDA IPC, so the persistent
// kernel of each GPU dereferences peer pointers directly -- no NCCL call and no host step inside an
// epoch.  Measured on this pool (tools/p2p_bench.cu, tools/ipc_bench.cu, 2 x B200 over NV18): random
// 800-byte row WRITES / REDs to a peer run at ~640 GB/s, but random row READS through an IPC mapping
// reach only ~190 GB/s (315 GB/s through a cuMem fd-shared mapping, 570 GB/s inside one process), and
// remote loads also queue behind posted writes.  So nothing is ever loaded across NVLink: every
// transfer is a posted write issued by the side that holds the data.
//   phase 1a  every rank walks ALL samples of the batch, one per thread (Philox + one triple fetch: the
//             counter RNG makes the sample set independent of G and cheap to re-derive) and keeps those
//             whose positive head it owns -- the head row is then always local, which cuts the rows that
//             cross NVLink from 3(G-1)/G to 2(G-1)/G per sample.  Kept samples are appended to a local
//             list (one atomic per warp); for the tail / corrupting rows it does not own it writes an
//             8-byte request (local row index, launch-unique stamp) into the owner's request table at
//             the slot (list position, t|c) -- the slot IS the return address
//   barrier   (local grid barrier + cross-GPU barrier on peer-mapped counters, system-scope atomics)
//   phase 1s  every owner scans the request tables of its peers (coalesced stamp test + ballot) and
//             writes each requested row into the requester's row cache
//   barrier
//   phase 1b  scores every sample (own rows straight from the table, remote rows from the cache) and
//             accumulates the update with local vector REDs: rows it owns into its delta table, rows
//             owned by a peer into a local staging table; stamps them
//   barrier
//   phase 2a  staged rows are appended to the owner's inbox (one compact region per sender, slots reserved with
//             one local atomic per 32 scanned rows): sequential posted stores instead of random remote REDs,
//             which ran at only ~190 GB/s per GPU in the 8-GPU all-to-all; the entry counts travel with the barrier
//   barrier
//   phase 2a' every owner adds its inbox entries into its delta table with LOCAL vector REDs and stamps the rows
//   (local grid barrier)
//   phase 2b  every owner publishes its stamped rows (row += delta, normalise once); the owner of a
//             relation row also writes the new row into every replica
//   barrier
// Semantics are identical to the single-GPU kernel (same samples, same deferred renormalisation), so
// results agree up to the order of float additions.  Loss and counters are per rank (the host adds them).

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "train_device.cuh"

namespace kb2e {

constexpr int kMaxPeers = 8;
constexpr int kDistThreads = 768;

struct DistArgs {
   TrainArgs base;                     // sampler fields (triples, hash, pr, seeds, nE, n_train) + lr, margin, distance, D, P
   unsigned char* arena[kMaxPeers];    // arena[g] = rank g's arena as mapped into this process
   // arena layout: tab [rows_local][P] | dtab [rows_local][P] | flag [rows_local] |
   //               rel [nR][P] (replica) | drel [nR][P] | rflag [nR] | cross-GPU counter |
   //               req [world][req_cap] int2 (written by the peers) | cache [req_cap][P] (written by the peers)
   //               inbox [world][inbox_cap][4 + P] (written by the peers) | inbox_cnt [world] (written by the peers)
   size_t off_tab, off_dtab, off_flag, off_rel, off_drel, off_rflag, off_xbar, off_req, off_cache, off_inbox, off_inbox_cnt;
   long long req_cap;                  // 3 x the largest per-rank share of a batch
   long long inbox_cap;                // staged-row entries one sender may append per batch
   uint32_t* push_cnt;                 // local: [world] entries appended to each owner's inbox in this batch
   float* stage;                       // local: updates for rows owned by peers, [nE + nR][P] (global row ids)
   uint8_t* sflag;                     // local: stamps of the staged rows, [nE + nR]
   int4* pairs;                        // local: (h, t, r, c | corruptTail << 31) per kept sample
   uint32_t* share;                    // local: [0], [1] kept samples of the even / odd batch, [2] overflow flag
   long long max_share;                // capacity of pairs (and of the request slots / 3)
   uint32_t stamp_base;                // launch-unique high bits of the request stamps
   uint32_t* local_bar;
   int rank, world, wshift;
   int rows_local;
   uint32_t xbase;                     // value of every rank's cross-GPU counter when this launch starts
   unsigned long long timeout_ns;      // longest wait for the peers at one cross-GPU barrier before the launch is abandoned
};

// words of the 64-byte cross-GPU block at off_xbar: [0] arrival counter, [4] abort word (a peer gave up: stop polling),
// [8] error word (a peer dropped samples / updates in this launch: every rank must report it)
constexpr int kXbarAbort = 4, kXbarError = 8;

// value of entity row e as this rank sees it in phase 1b: its own table, or the cache slot the owner filled
__device__ __forceinline__ const float* ent_row(const DistArgs& a, int e, long long slot) {
   if ((e & (a.world - 1)) == a.rank)
      return reinterpret_cast<const float*>(a.arena[a.rank] + a.off_tab) + (size_t)(e >> a.wshift) * a.base.P;
   return reinterpret_cast<const float*>(a.arena[a.rank] + a.off_cache) + (size_t)slot * a.base.P;
}
// where this rank accumulates its update of entity row e, and the stamp that goes with it
__device__ __forceinline__ float* ent_delta(const DistArgs& a, int e, uint8_t*& flag) {
   if ((e & (a.world - 1)) == a.rank) {
      const size_t l = (size_t)(e >> a.wshift);
      flag = a.arena[a.rank] + a.off_flag + l;
      return reinterpret_cast<float*>(a.arena[a.rank] + a.off_dtab) + l * a.base.P;
   }
   flag = a.sflag + e;
   return a.stage + (size_t)e * a.base.P;
}
__device__ __forceinline__ float* rel_delta(const DistArgs& a, int r, uint8_t*& flag) {
   if ((r & (a.world - 1)) == a.rank) {
      flag = a.arena[a.rank] + a.off_rflag + r;
      return reinterpret_cast<float*>(a.arena[a.rank] + a.off_drel) + (size_t)r * a.base.P;
   }
   flag = a.sflag + a.base.nE + r;
   return a.stage + ((size_t)a.base.nE + r) * a.base.P;
}

// all CTAs of this GPU, then all GPUs, then all CTAs again (release / acquire at system scope).  publish: after this
// GPU's CTAs have all arrived, its per-owner inbox counts are written to the owners before the cross-GPU signal.
// The wait for the peers is BOUNDED: a peer that never launched (its host call failed) or died would otherwise hang every
// GPU of the box until an operator resets it.  The poller gives up when a peer has raised the abort word of this arena
// or when timeout_ns have passed; it then raises the abort word on every peer, records the failure in share[3] and the
// whole grid returns (false): the host reports KB2E_ERR_PEER and the context refuses further partitioned calls.
__device__ __forceinline__ bool cross_barrier(const DistArgs& a, uint32_t& ltarget, uint32_t& xtarget, bool publish = false) {
   __threadfence_system();   // this thread's peer stores / REDs are performed before it arrives
   grid_barrier(a.local_bar, ltarget);
   if (blockIdx.x == 0) {
      if (publish) {
         if ((int)threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(a.arena[threadIdx.x] + a.off_inbox_cnt) + a.rank;
            *dst = __ldcg(a.push_cnt + threadIdx.x);
            __threadfence_system();
         }
         __syncthreads();
      }
      if (threadIdx.x == 0) {
         xtarget += (uint32_t)a.world;
         asm volatile("fence.acq_rel.sys;" ::: "memory");
         for (int g = 0; g < a.world; g++) {
            uint32_t* p = reinterpret_cast<uint32_t*>(a.arena[g] + a.off_xbar);
            asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(p), "r"(1u) : "memory");
         }
         const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.arena[a.rank] + a.off_xbar);
         uint32_t v, polls = 0;
         unsigned long long t_start = 0;
         bool dead = false;
         while (true) {
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - xtarget) >= 0) break;
            if ((++polls & 255u) == 0u) {
               uint32_t ab;
               asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(ab) : "l"(mine + kXbarAbort) : "memory");
               unsigned long long now;
               asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
               if (t_start == 0) t_start = now;
               if (ab != 0u || now - t_start > a.timeout_ns) { dead = true; break; }
            }
         }
         if (dead) {
            for (int g = 0; g < a.world; g++) {
               uint32_t* p = reinterpret_cast<uint32_t*>(a.arena[g] + a.off_xbar) + kXbarAbort;
               asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(1u) : "memory");
            }
            __stcg(a.share + 3, 1u);
         }
         asm volatile("fence.acq_rel.sys;" ::: "memory");
      }
   }
   grid_barrier(a.local_bar, ltarget);
   return __ldcg(a.share + 3) == 0u;
}

template <int LPS, int NV>
__device__ __forceinline__ void dist_process_pair(const DistArgs& a, const Pair s, long long j, int gl, uint32_t gmask,
                                                  uint8_t stamp, double& loss_acc, uint32_t& active_acc) {
   const TrainArgs& b = a.base;
   const int P = b.P, D = b.D;
   float4 vh[NV], vt[NV], vc[NV], vr[NV];
   load_row<LPS, NV>(ent_row(a, s.h, 3 * j), P, gl, vh);
   load_row<LPS, NV>(ent_row(a, s.t, 3 * j + 1), P, gl, vt);
   load_row<LPS, NV>(ent_row(a, s.c, 3 * j + 2), P, gl, vc);
   load_row<LPS, NV>(reinterpret_cast<float*>(a.arena[a.rank] + a.off_rel) + (size_t)s.r * P, P, gl, vr);   // local replica
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
   uint8_t *fh, *ft, *fc, *fr;
   float* dh = ent_delta(a, s.h, fh);
   float* dt = ent_delta(a, s.t, ft);
   float* dc = ent_delta(a, s.c, fc);
   float* dr = rel_delta(a, s.r, fr);
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
   if (gl < 4) *(gl == 0 ? fh : (gl == 1 ? ft : (gl == 2 ? fc : fr))) = stamp;
}

template <int LPS, int NV>
__global__ void __launch_bounds__(kDistThreads, 1) train_dist_kernel(const __grid_constant__ DistArgs a) {
   __shared__ double s_loss[kDistThreads / 32];
   const TrainArgs& b = a.base;
   const int lane = threadIdx.x & 31;
   const int gl = lane % LPS;
   const uint32_t gmask = LPS == 32 ? 0xffffffffu : (((1u << LPS) - 1u) << ((lane / LPS) * LPS));
   const int groups_per_block = blockDim.x / LPS;
   const long long G = (long long)gridDim.x * groups_per_block;
   const long long g0 = (long long)(threadIdx.x / LPS) * gridDim.x + blockIdx.x;
   const int P = b.P;
   unsigned char* me = a.arena[a.rank];
   float* tab = reinterpret_cast<float*>(me + a.off_tab);
   float* dtab = reinterpret_cast<float*>(me + a.off_dtab);
   const uint8_t* flag = me + a.off_flag;
   float* drel = reinterpret_cast<float*>(me + a.off_drel);
   const uint8_t* rflag = me + a.off_rflag;
   const long long RL = a.rows_local;
   uint32_t ltarget = 0, xtarget = a.xbase;
   uint32_t active_acc = 0, tent_acc = 0, trel_acc = 0;
   int trace_slot = 0;
#define KB2E_DTRACE()                                                                                 \
   if (b.trace != nullptr && threadIdx.x == 0 && trace_slot < kTraceSlots) {                          \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      b.trace[(size_t)blockIdx.x * kTraceSlots + trace_slot++] = t_;                                  \
   }
   const uint32_t gb_first = (uint32_t)b.first_epoch * (uint32_t)b.batches;
   const long long T = (long long)gridDim.x * blockDim.x;
   const long long t0 = (long long)threadIdx.x * gridDim.x + blockIdx.x;
   // peers may still be zeroing / publishing from the previous launch
   if (!cross_barrier(a, ltarget, xtarget)) return;

   for (int ep = 0; ep < b.n_epochs; ep++) {
      double loss_acc = 0.0;
      for (int batch = 0; batch < b.batches; batch++) {
         const uint32_t rel_batch = (uint32_t)(ep * b.batches + batch);
         const uint32_t gb = gb_first + rel_batch;
         const uint8_t stamp = (uint8_t)(gb % 255u + 1u);
         const int rstamp = (int)(a.stamp_base + rel_batch + 1u);   // unique per batch and per launch, never 0
         KB2E_DTRACE();
         // ---- phase 1a: walk the batch, keep the samples whose head lives here, request their remote rows ----
         uint32_t* my_cnt = a.share + (rel_batch & 1u);
         for (long long k0 = 0; k0 < b.batchsize; k0 += T) {
            const long long k = k0 + t0;
            bool mine = false;
            if (k < b.batchsize) {
               int h;
               if (b.pairs != nullptr) {
                  h = __ldg(b.pairs + 6ll * k);
               } else {
                  uint32_t x[4];
                  philox4x32((uint32_t)k, gb, 0u, 0u, b.seed_lo, b.seed_hi, x);
                  h = __ldg(b.triples + mulhi64(((uint64_t)x[0] << 32) | x[1], (uint64_t)b.n_train)).x;
               }
               mine = (h & (a.world - 1)) == a.rank;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, mine);
            uint32_t base = 0;
            if (lane == 0 && m) base = atomicAdd(my_cnt, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (mine) {
               const long long j = (long long)base + __popc(m & ((1u << lane) - 1u));
               if (j < a.max_share) {
                  const Pair s = draw_pair(b, (uint32_t)k, gb);
                  a.pairs[j] = make_int4(s.h, s.t, s.r, s.c | (s.corruptTail ? 0x80000000 : 0));
#pragma unroll
                  for (int w = 1; w < 3; w++) {
                     const int e = w == 1 ? s.t : s.c;
                     const int g = e & (a.world - 1);
                     if (g != a.rank) {
                        int2* req = reinterpret_cast<int2*>(a.arena[g] + a.off_req) + (size_t)a.rank * a.req_cap + 3 * j + w;
                        *req = make_int2(e >> a.wshift, rstamp);
                     }
                  }
               } else {
                  a.share[2] = 1u;   // more heads on this rank than the buffers were sized for: reported by the host
               }
            }
         }
         KB2E_DTRACE();
         if (!cross_barrier(a, ltarget, xtarget)) return;
         KB2E_DTRACE();
         // ---- phase 1s: serve the peers' requests: own row -> the requester's cache slot ----
         if (a.world > 1) {
            const int2* req = reinterpret_cast<const int2*>(me + a.off_req);
            long long first, end;
            group_range(0, (long long)a.world * a.req_cap, g0, G, first, end);
            for_stamped_rows<LPS>(first, end, gl, gmask, lane, [&](long long r) { return __ldcg(&req[r].y) == rstamp; },
                                  [&](long long r0, long long r1) {
               float4 v0[NV], v1[NV];
               const int l0 = __ldcg(&req[r0].x);
               load_row<LPS, NV>(tab + (size_t)l0 * P, P, gl, v0);
               if (r1 >= 0) {
                  const int l1 = __ldcg(&req[r1].x);
                  load_row<LPS, NV>(tab + (size_t)l1 * P, P, gl, v1);
               }
               auto send = [&](long long r, float4 (&v)[NV]) {
                  const int pr = (int)(r / a.req_cap);
                  const long long slot = r - (long long)pr * a.req_cap;
                  store_row<LPS, NV>(reinterpret_cast<float*>(a.arena[pr] + a.off_cache) + (size_t)slot * P, P, gl, v);
               };
               send(r0, v0);
               if (r1 >= 0) send(r1, v1);
            });
         }
         KB2E_DTRACE();
         if (!cross_barrier(a, ltarget, xtarget)) return;
         KB2E_DTRACE();
         // ---- phase 1b: score + accumulate, everything local ----
         const long long my_count = min((long long)__ldcg(my_cnt), a.max_share);
         if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.share[(rel_batch & 1u) ^ 1u] = 0u;   // the next batch's counter
            atomicAdd(b.counters + 4, (unsigned long long)my_count);
         }
         if (blockIdx.x == 0 && (int)threadIdx.x < a.world) a.push_cnt[threadIdx.x] = 0u;   // this batch's inbox appends
         for (long long j = g0; j < my_count; j += G) {
            const int4 pr = __ldcg(a.pairs + j);
            Pair s;
            s.h = pr.x; s.t = pr.y; s.r = pr.z; s.c = pr.w & 0x7fffffff; s.corruptTail = pr.w < 0;
            dist_process_pair<LPS, NV>(a, s, j, gl, gmask, stamp, loss_acc, active_acc);
         }
         KB2E_DTRACE();
         if (!cross_barrier(a, ltarget, xtarget)) return;
         KB2E_DTRACE();
         // ---- phase 2a: append the staged rows to their owners' inboxes (sequential posted stores) ----
         {
            const int S = P + 4;   // entry: 16-byte header (target row) + the row
            const long long rows_max = (b.nE + a.world - 1) >> a.wshift;
            const int gshift = (lane / LPS) * LPS;
            float4 z[NV];
#pragma unroll
            for (int q = 0; q < NV; q++) z[q] = f4(0.f);
            auto append = [&](int owner, long long idx, long long r, int hdr) {
               float4 d[NV];
               load_row<LPS, NV>(a.stage + (size_t)r * P, P, gl, d);
               if (idx < a.inbox_cap) {
                  float* ent = reinterpret_cast<float*>(a.arena[owner] + a.off_inbox) + ((size_t)a.rank * a.inbox_cap + idx) * S;
                  if (gl == 0) *reinterpret_cast<int4*>(ent) = make_int4(hdr, 0, 0, 0);
                  store_row<LPS, NV>(ent + 4, P, gl, d);
               } else if (gl == 0) {
                  a.share[2] = 1u;   // more staged rows for one owner than the inbox holds: reported by the host
               }
               store_row<LPS, NV>(a.stage + (size_t)r * P, P, gl, z);
               if (gl == 0) a.sflag[r] = 0;   // 8-bit stamps recur every 255 batches: a pushed row must not look staged again
            };
            for (int o = 0; o < a.world; o++) {
               if (o == a.rank) continue;   // own rows never pass through the staging table
               for (long long l0 = g0 * LPS; l0 < rows_max; l0 += G * LPS) {
                  const long long l = l0 + gl;
                  const long long r = (l << a.wshift) + o;
                  const bool f = l < rows_max && r < b.nE && __ldcg(a.sflag + r) == stamp;
                  uint32_t m = (__ballot_sync(gmask, f) >> gshift) & (LPS == 32 ? 0xffffffffu : ((1u << LPS) - 1u));
                  if (!m) continue;
                  uint32_t base = 0;
                  if (gl == 0) base = atomicAdd(a.push_cnt + o, (uint32_t)__popc(m));
                  base = __shfl_sync(gmask, base, gshift);
                  for (int k = 0; m; k++) {
                     const int bit = __ffs(m) - 1;
                     m &= m - 1;
                     const long long lr = l0 + bit;
                     append(o, (long long)base + k, (lr << a.wshift) + o, (int)lr);
                  }
               }
            }
            // relation rows (few): one slot at a time
            for (long long rr = g0; rr < b.nR; rr += G) {
               const int o = (int)(rr & (a.world - 1));
               if (o == a.rank || __ldcg(a.sflag + b.nE + rr) != stamp) continue;
               uint32_t idx = 0;
               if (gl == 0) idx = atomicAdd(a.push_cnt + o, 1u);
               idx = __shfl_sync(gmask, idx, gshift);
               append(o, (long long)idx, (long long)b.nE + rr, -1 - (int)rr);
            }
         }
         KB2E_DTRACE();
         if (!cross_barrier(a, ltarget, xtarget, true)) return;
         KB2E_DTRACE();
         // ---- phase 2a': add the inbox entries into the own delta tables (local REDs), stamp the rows ----
         if (a.world > 1) {
            const int S = P + 4;
            const uint32_t* cnts = reinterpret_cast<const uint32_t*>(me + a.off_inbox_cnt);
            for (int src = 0; src < a.world; src++) {
               if (src == a.rank) continue;
               const long long n = min((long long)__ldcg(cnts + src), a.inbox_cap);
               const float* box = reinterpret_cast<const float*>(me + a.off_inbox) + (size_t)src * a.inbox_cap * S;
               for (long long e = g0; e < n; e += 2 * G) {
                  const long long e1 = e + G;
                  const float* ent0 = box + (size_t)e * S;
                  const float* ent1 = box + (size_t)(e1 < n ? e1 : e) * S;
                  const int h0 = __ldcg(reinterpret_cast<const int*>(ent0));
                  const int h1 = __ldcg(reinterpret_cast<const int*>(ent1));
                  float4 d0[NV], d1[NV];
                  load_row<LPS, NV>(ent0 + 4, P, gl, d0);
                  load_row<LPS, NV>(ent1 + 4, P, gl, d1);
                  auto absorb = [&](int hdr, float4 (&d)[NV]) {
                     if (hdr >= 0) {
                        red_row<LPS, NV>(dtab + (size_t)hdr * P, P, gl, d);
                        if (gl == 0) me[a.off_flag + hdr] = stamp;
                     } else {
                        const int rr = -1 - hdr;
                        red_row<LPS, NV>(drel + (size_t)rr * P, P, gl, d);
                        if (gl == 0) me[a.off_rflag + rr] = stamp;
                     }
                  };
                  absorb(h0, d0);
                  if (e1 < n) absorb(h1, d1);
               }
            }
            grid_barrier(a.local_bar, ltarget);   // every absorbed update has landed before the publish reads the deltas
         }
         KB2E_DTRACE();
         // ---- phase 2b: own entity rows: row += delta, normalise once, publish ----
         {
            long long first, end;
            group_range(0, RL, g0, G, first, end);
            for_stamped_rows<LPS>(first, end, gl, gmask, lane, [&](long long r) { return __ldcg(flag + r) == stamp; },
                                  [&](long long r0, long long r1) {
               float4 x0[NV], d0[NV], x1[NV], d1[NV];
               load_row<LPS, NV>(tab + (size_t)r0 * P, P, gl, x0);
               load_row<LPS, NV>(dtab + (size_t)r0 * P, P, gl, d0);
               if (r1 >= 0) {
                  load_row<LPS, NV>(tab + (size_t)r1 * P, P, gl, x1);
                  load_row<LPS, NV>(dtab + (size_t)r1 * P, P, gl, d1);
               }
               auto publish = [&](long long r, float4 (&x)[NV], float4 (&d)[NV]) {
#pragma unroll
                  for (int q = 0; q < NV; q++) { x[q] = x[q] + d[q]; d[q] = f4(0.f); }
                  store_row<LPS, NV>(dtab + (size_t)r * P, P, gl, d);
                  norm_row<LPS, NV>(x, true, gmask);   // transe/trainer.cpp:44-45
                  store_row<LPS, NV>(tab + (size_t)r * P, P, gl, x);
                  if (gl == 0) me[a.off_flag + r] = 0;   // stamp cleared: only rows with a real delta are published (and counted)
                  tent_acc += (gl == 0);
               };
               publish(r0, x0, d0);
               if (r1 >= 0) publish(r1, x1, d1);
            });
         }
         KB2E_DTRACE();
         // ---- phase 2b: relation rows owned by this rank -> every replica ----
         for (long long r = (long long)a.rank + g0 * a.world; r < b.nR; r += G * a.world) {
            if (__ldcg(rflag + r) != stamp) continue;
            float4 x[NV], d[NV];
            load_row<LPS, NV>(reinterpret_cast<float*>(me + a.off_rel) + (size_t)r * P, P, gl, x);
            load_row<LPS, NV>(drel + (size_t)r * P, P, gl, d);
#pragma unroll
            for (int q = 0; q < NV; q++) { x[q] = x[q] + d[q]; d[q] = f4(0.f); }
            store_row<LPS, NV>(drel + (size_t)r * P, P, gl, d);
            if (gl == 0) me[a.off_rflag + r] = 0;
            norm_row<LPS, NV>(x, true, gmask);   // transe/trainer.cpp:43
            for (int g = 0; g < a.world; g++)
               store_row<LPS, NV>(reinterpret_cast<float*>(a.arena[g] + a.off_rel) + (size_t)r * P, P, gl, x);
            trel_acc += (gl == 0);
         }
         KB2E_DTRACE();
         if (!cross_barrier(a, ltarget, xtarget)) return;
      }
      double v = (gl == 0) ? loss_acc : 0.0;
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
   // one status for the whole job: a rank that dropped samples or updates (overflow flag share[2], set only where the
   // overflow happened) raises the error word of every peer before the last cross-GPU barrier; afterwards every rank
   // folds its own error word into share[2], so that all ranks return the same error (the owner whose updates were
   // dropped by a sender would otherwise report success with silently wrong tables)
   if (blockIdx.x == 0 && (int)threadIdx.x < a.world && __ldcg(a.share + 2) != 0u) {
      uint32_t* p = reinterpret_cast<uint32_t*>(a.arena[threadIdx.x] + a.off_xbar) + kXbarError;
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(1u) : "memory");
   }
   if (!cross_barrier(a, ltarget, xtarget)) return;
   if (blockIdx.x == 0 && threadIdx.x == 0) {
      uint32_t e;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(e) : "l"(reinterpret_cast<const uint32_t*>(me + a.off_xbar) + kXbarError) : "memory");
      if (e != 0u) a.share[2] = 1u;
   }
   uint32_t c0 = (gl == 0) ? active_acc : 0u, c1 = tent_acc, c2 = trel_acc;
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

// init: N(0, (1/D)^2) keyed by the GLOBAL row id, so every world size starts from the same tables
__global__ void dist_init_kernel(float* tab, long long rows_local, int rank, int world, long long first_row, int D, int P, uint32_t k0,
                                 uint32_t k1) {
   long long lrow = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
   int lane = threadIdx.x & 31;
   if (lrow >= rows_local) return;
   long long row = first_row + lrow * world + rank;   // row id in the unified row space of the single-GPU tables
   float* p = tab + lrow * P;
   float s2 = 0.f;
   const float sigma = 1.0f / (float)D;
   for (int base = lane * 4; base < P; base += 128) {
      uint32_t x[4];
      philox4x32((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)base, 1u, k0, k1, x);
      float v[4];
#pragma unroll
      for (int q = 0; q < 2; q++) {
         float u1 = ((float)x[2 * q] + 1.0f) * 2.3283064e-10f;
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
   if (len > 1.f) {
      for (int base = lane * 4; base < P; base += 128) {
         float4 v = *reinterpret_cast<float4*>(p + base);
         *reinterpret_cast<float4*>(p + base) = make_float4(v.x / len, v.y / len, v.z / len, v.w / len);
      }
   }
}

__global__ void dist_widen_kernel(const float* src, double* dst, long long rows, int D, int P) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < rows * D) dst[i] = (double)src[(i / D) * P + (i % D)];
}
__global__ void dist_narrow_kernel(const double* src, float* dst, long long rows, int D, int P) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= rows * P) return;
   int c = (int)(i % P);
   dst[i] = c < D ? (float)src[(i / P) * D + c] : 0.f;
}

}  // namespace kb2e

using namespace kb2e;

struct DistState {
   int rank = 0, world = 1, wshift = 0;
   long long rows_local = 0;
   unsigned char* arena = nullptr;
   size_t arena_bytes = 0;
   size_t off_tab = 0, off_dtab = 0, off_flag = 0, off_rel = 0, off_drel = 0, off_rflag = 0, off_xbar = 0, off_req = 0, off_cache = 0,
          off_inbox = 0, off_inbox_cnt = 0;
   long long inbox_cap = 0;
   uint32_t* push_cnt = nullptr;
   long long req_cap = 0;        // request / cache slots per requester: 3 x the largest per-rank share of a batch
   long long max_share = 0;
   float* stage = nullptr;
   uint8_t* sflag = nullptr;
   int4* pairs = nullptr;
   uint32_t* share = nullptr;
   unsigned long long samples_seen = 0;   // value of counters[4] after the previous launch
   uint32_t launches = 0;
   unsigned char* peers[kMaxPeers] = {};
   bool connected = false;
   uint32_t xcount = 0;   // cross-GPU barrier arrivals so far (identical on every rank: the counters are never reset)
   bool broken = false;   // a launch was abandoned (a peer never arrived): the barrier counters of the ranks no longer agree
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline unsigned dblocks(long long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" {

int kb2e_dist_setup(kb2e_ctx* c, int32_t rank, int32_t world, void* handle_out) {
   if (!c || !handle_out) return KB2E_ERR_ARG;
   KB2E_CUDA(c, cudaSetDevice(c->device));
   if (c->cfg.model != KB2E_MODEL_TRANSE) return fail(c, KB2E_ERR_LIMIT, "partitioned training is built for TransE");
   if (world < 1 || world > kMaxPeers || (world & (world - 1)) || rank < 0 || rank >= world)
      return fail(c, KB2E_ERR_ARG, "kb2e_dist_setup: world must be 1, 2, 4 or 8 and 0 <= rank < world");
   if (!c->triples || c->n_train == 0 || c->cfg.batches <= 0)
      return fail(c, KB2E_ERR_ARG, "kb2e_dist_setup: call kb2e_set_train_triples first (the batch size sizes the exchange buffers)");
   int rc = train_alloc(c);   // barrier counter, counters, pr
   if (rc) return rc;
   kb2e_dist_teardown(c);     // a second setup replaces the first one: release its arena, staging buffers and IPC mappings
   DistState* d = new DistState();
   c->dist = d;               // owned by the context from here on: an error below is cleaned up by kb2e_dist_teardown / kb2e_destroy
   d->rank = rank; d->world = world;
   while ((1 << d->wshift) < world) d->wshift++;
   d->rows_local = ((long long)c->nE - rank + world - 1) / world;
   // every rank uses the SAME offsets (computed from rank 0's row count, the largest), so a peer's arena can be
   // addressed without knowing its private layout
   const long long rows_max = ((long long)c->nE + world - 1) / world;
   const size_t tab_bytes = (size_t)rows_max * c->P * sizeof(float);
   const size_t rel_bytes = (size_t)c->nR * c->P * sizeof(float);
   size_t off = 0;
   d->off_tab = off; off = align_up(off + tab_bytes, 256);
   d->off_dtab = off; off = align_up(off + tab_bytes, 256);
   d->off_flag = off; off = align_up(off + (size_t)rows_max, 256);
   d->off_rel = off; off = align_up(off + rel_bytes, 256);
   d->off_drel = off; off = align_up(off + rel_bytes, 256);
   d->off_rflag = off; off = align_up(off + (size_t)c->nR, 256);
   d->off_xbar = off; off = align_up(off + 64, 256);
   // exchange buffers: this rank's request table (one region per requester) and its row cache
   // samples are dealt by the owner of their head, so a rank's share of a batch is binomial(batch, 1/world):
   // mean + 8 sigma + slack (an overflow is detected by the kernel and reported, never silent)
   {
      const long long batch = c->n_train / c->cfg.batches;
      const double mean = (double)batch / world;
      d->max_share = world == 1 ? batch : std::min<long long>(batch, (long long)(mean + 8.0 * sqrt(mean) + 1024.0));
   }
   d->req_cap = 3 * d->max_share;
   d->off_req = off; off = align_up(off + (size_t)world * d->req_cap * sizeof(int2), 256);
   d->off_cache = off; off = align_up(off + (world > 1 ? (size_t)d->req_cap * c->P * sizeof(float) : 0), 256);
   // inbox: per sender, the staged rows it may append in one batch.  A sender stages at most 2 remote rows per kept
   // sample (tail, corrupting entity), spread evenly over the world - 1 other owners: mean + 50 % + slack, never more
   // than the worst case (an overflow is detected by the kernel and reported)
   d->inbox_cap = world == 1 ? 0 : std::min<long long>(2 * d->max_share + c->nR, (long long)(3.0 * d->max_share / world) + c->nR + 4096);
   d->off_inbox = off; off = align_up(off + (size_t)world * d->inbox_cap * (c->P + 4) * sizeof(float), 256);
   d->off_inbox_cnt = off; off = align_up(off + (size_t)world * sizeof(uint32_t), 256);
   d->arena_bytes = off;
   KB2E_CUDA(c, cudaMalloc(&d->arena, d->arena_bytes));
   KB2E_CUDA(c, cudaMemset(d->arena, 0, d->arena_bytes));
   const size_t stage_rows = (size_t)c->nE + c->nR;
   KB2E_CUDA(c, cudaMalloc(&d->stage, stage_rows * c->P * sizeof(float)));
   KB2E_CUDA(c, cudaMemset(d->stage, 0, stage_rows * c->P * sizeof(float)));
   KB2E_CUDA(c, cudaMalloc(&d->sflag, stage_rows));
   KB2E_CUDA(c, cudaMemset(d->sflag, 0, stage_rows));
   KB2E_CUDA(c, cudaMalloc(&d->pairs, (size_t)std::max<long long>(1, d->max_share) * sizeof(int4)));
   KB2E_CUDA(c, cudaMalloc(&d->share, 4 * sizeof(uint32_t)));
   KB2E_CUDA(c, cudaMalloc(&d->push_cnt, kMaxPeers * sizeof(uint32_t)));
   KB2E_CUDA(c, cudaMemset(d->push_cnt, 0, kMaxPeers * sizeof(uint32_t)));
   cudaIpcMemHandle_t h;
   KB2E_CUDA(c, cudaIpcGetMemHandle(&h, d->arena));
   static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
   memcpy(handle_out, &h, sizeof(h));
   return KB2E_OK;
}

int kb2e_dist_connect(kb2e_ctx* c, const void* handles) {
   if (!c || !c->dist || !handles) return KB2E_ERR_ARG;
   KB2E_CUDA(c, cudaSetDevice(c->device));
   DistState* d = c->dist;
   for (int g = 0; g < d->world; g++) {
      if (g == d->rank) { d->peers[g] = d->arena; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, (const unsigned char*)handles + 64 * g, sizeof(h));
      void* p = nullptr;
      KB2E_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      d->peers[g] = (unsigned char*)p;
   }
   d->connected = true;
   return KB2E_OK;
}

int kb2e_dist_init_embeddings(kb2e_ctx* c) {
   if (!c || !c->dist) return KB2E_ERR_ARG;
   KB2E_CUDA(c, cudaSetDevice(c->device));
   DistState* d = c->dist;
   uint32_t k0 = (uint32_t)c->cfg.seed, k1 = (uint32_t)(c->cfg.seed >> 32);
   dist_init_kernel<<<dblocks(d->rows_local * 32, 256), 256, 0, c->stream>>>(
      reinterpret_cast<float*>(d->arena + d->off_tab), d->rows_local, d->rank, d->world, 0, c->D, c->P, k0, k1);
   // relation replica: rows nE .. nE+nR-1 of the unified row space with the SAME key, identical on every rank and
   // identical to kb2e_init_embeddings (train.cu:init_rows_kernel), so partitioned and single-GPU runs of one seed
   // start from the same tables
   dist_init_kernel<<<dblocks((long long)c->nR * 32, 256), 256, 0, c->stream>>>(
      reinterpret_cast<float*>(d->arena + d->off_rel), c->nR, 0, 1, (long long)c->nE, c->D, c->P, k0, k1);
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   return KB2E_OK;
}

// table: KB2E_TABLE_ENTITY -> this rank's rows (global ids rank, rank+world, ...), [rows_local][dim];
//        KB2E_TABLE_RELATION -> the replica, [num_relations][dim]
static int dist_copy(kb2e_ctx* c, int table, double* host, const double* host_in, int64_t rows, int64_t cols) {
   if (!c || !c->dist) return KB2E_ERR_ARG;
   KB2E_CUDA(c, cudaSetDevice(c->device));
   DistState* d = c->dist;
   const long long want = table == KB2E_TABLE_ENTITY ? d->rows_local : c->nR;
   if (table != KB2E_TABLE_ENTITY && table != KB2E_TABLE_RELATION) return fail(c, KB2E_ERR_ARG, "kb2e_dist: unknown table");
   if (rows != want || cols != c->D) return fail(c, KB2E_ERR_ARG, "kb2e_dist: shape mismatch");
   float* dev32 = reinterpret_cast<float*>(d->arena + (table == KB2E_TABLE_ENTITY ? d->off_tab : d->off_rel));
   double* tmp = nullptr;
   KB2E_CUDA(c, cudaMalloc(&tmp, (size_t)rows * cols * sizeof(double)));
   if (host_in) {
      KB2E_CUDA(c, cudaMemcpyAsync(tmp, host_in, (size_t)rows * cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      dist_narrow_kernel<<<dblocks(rows * c->P, 256), 256, 0, c->stream>>>(tmp, dev32, rows, c->D, c->P);
   } else {
      dist_widen_kernel<<<dblocks(rows * c->D, 256), 256, 0, c->stream>>>(dev32, tmp, rows, c->D, c->P);
      KB2E_CUDA(c, cudaMemcpyAsync(host, tmp, (size_t)rows * cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   }
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   cudaFree(tmp);
   return KB2E_OK;
}

int kb2e_dist_upload(kb2e_ctx* c, int table, const double* host, int64_t rows, int64_t cols) {
   return host ? dist_copy(c, table, nullptr, host, rows, cols) : KB2E_ERR_ARG;
}
int kb2e_dist_download(kb2e_ctx* c, int table, double* host, int64_t rows, int64_t cols) {
   return host ? dist_copy(c, table, host, nullptr, rows, cols) : KB2E_ERR_ARG;
}

int kb2e_dist_train_epochs(kb2e_ctx* c, int32_t first_epoch, int32_t n_epochs, double* loss_per_epoch) {
   if (!c || !c->dist) return KB2E_ERR_ARG;
   KB2E_CUDA(c, cudaSetDevice(c->device));
   DistState* d = c->dist;
   if (!d->connected) return fail(c, KB2E_ERR_ARG, "kb2e_dist_train_epochs: call kb2e_dist_connect first");
   if (d->broken)
      return fail(c, KB2E_ERR_PEER, "kb2e_dist_train_epochs: an earlier launch was abandoned because a peer never arrived; the ranks are out of "
                                    "step -- call kb2e_dist_teardown on every rank and set the job up again");
   if (!c->triples || c->n_train == 0 || !c->have_pr) return fail(c, KB2E_ERR_ARG, "kb2e_dist_train_epochs: set the train triples and bern statistics first");
   if (n_epochs <= 0) return KB2E_OK;
   if (n_epochs > c->loss_cap) {
      cudaFree(c->loss_dev);
      c->loss_dev = nullptr;
      c->loss_cap = 0;
      KB2E_CUDA(c, cudaMalloc(&c->loss_dev, (size_t)n_epochs * sizeof(double)));
      c->loss_cap = n_epochs;
   }
   DistArgs a;
   memset(&a, 0, sizeof(a));
   TrainArgs& b = a.base;
   b.triples = c->triples; b.hash = c->hash; b.hash_mask = c->hash_mask; b.pr = c->pr;
   b.loss = c->loss_dev; b.counters = c->counters;
   b.n_train = c->n_train;
   b.batchsize = c->n_train / c->cfg.batches;
   b.nE = c->nE; b.nR = c->nR; b.D = c->D; b.P = c->P;
   b.batches = c->cfg.batches; b.first_epoch = first_epoch; b.n_epochs = n_epochs; b.distance = c->cfg.distance;
   b.lr = (float)c->cfg.rate; b.margin = (float)c->cfg.margin;
   b.seed_lo = (uint32_t)c->cfg.seed; b.seed_hi = (uint32_t)(c->cfg.seed >> 32);
   for (int g = 0; g < d->world; g++) a.arena[g] = d->peers[g];
   a.off_tab = d->off_tab; a.off_dtab = d->off_dtab; a.off_flag = d->off_flag; a.off_rel = d->off_rel;
   a.off_drel = d->off_drel; a.off_rflag = d->off_rflag; a.off_xbar = d->off_xbar;
   a.off_req = d->off_req; a.off_cache = d->off_cache; a.req_cap = d->req_cap;
   a.off_inbox = d->off_inbox; a.off_inbox_cnt = d->off_inbox_cnt; a.inbox_cap = d->inbox_cap; a.push_cnt = d->push_cnt;
   a.stage = d->stage; a.sflag = d->sflag; a.pairs = d->pairs;
   {
      const double mean = (double)b.batchsize / d->world;
      const long long need = d->world == 1 ? b.batchsize : std::min<long long>(b.batchsize, (long long)(mean + 8.0 * sqrt(mean) + 1024.0));
      if (need > d->max_share)
         return fail(c, KB2E_ERR_ARG, "kb2e_dist_train_epochs: the batch grew after kb2e_dist_setup; set the train triples before kb2e_dist_setup");
   }
   a.share = d->share; a.max_share = d->max_share;
   KB2E_CUDA(c, cudaMemsetAsync(d->share, 0, 4 * sizeof(uint32_t), c->stream));
   if ((uint64_t)c->cfg.batches * (uint64_t)n_epochs >= (1u << 20))
      return fail(c, KB2E_ERR_LIMIT, "kb2e_dist_train_epochs: at most 2^20 batches per call");
   // request stamps: (launch number << 20) | batch-in-launch + 1: an entry left by an earlier launch can never match
   d->launches = (d->launches + 1u) & 0x7ffu;
   a.stamp_base = d->launches << 20;
   a.local_bar = c->barrier;
   a.rank = d->rank; a.world = d->world; a.wshift = d->wshift; a.rows_local = (int)d->rows_local;
   // shape: same rule as the single-GPU kernel, on this rank's share of the batch
   const int vecs = (c->P + 3) / 4;
   int lps = vecs <= 16 ? 16 : 32;
   int nv = (vecs + lps - 1) / lps;
   nv = nv <= 1 ? 1 : (nv <= 2 ? 2 : 99);
   void (*k)(const DistArgs) = nullptr;
   if (lps == 16 && nv == 1) k = train_dist_kernel<16, 1>;
   else if (lps == 32 && nv == 1) k = train_dist_kernel<32, 1>;
   else if (lps == 32 && nv == 2) k = train_dist_kernel<32, 2>;
   if (!k) return fail(c, KB2E_ERR_LIMIT, "partitioned training supports embedding sizes up to 256");
   // the peer-mapped cross-GPU counters are monotonic across launches (a reset could wipe a fast peer's
   // arrival); every rank runs the same barrier sequence, so the start value is known on the host.  Everything that can
   // fail on the host happens BEFORE the counter is advanced; from the launch on, a rank that does not make it is
   // noticed by its peers through the bounded wait of cross_barrier.
   a.xbase = d->xcount;
   {
      const char* env = getenv("KB2E_DIST_TIMEOUT_MS");   // per cross-GPU barrier; default 30 s
      const double ms = env ? atof(env) : 30000.0;
      a.timeout_ns = (unsigned long long)(std::max(1.0, ms) * 1e6);
   }
   KB2E_CUDA(c, cudaMemsetAsync(c->barrier, 0, 64, c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->loss_dev, 0, (size_t)n_epochs * sizeof(double), c->stream));
   // this rank's error word: the peers write it only before the last barrier of a launch they share with this rank
   KB2E_CUDA(c, cudaMemsetAsync(d->arena + d->off_xbar + kXbarError * sizeof(uint32_t), 0, sizeof(uint32_t), c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   const char* trace_path = getenv("KB2E_TRAIN_TRACE");
   unsigned long long* trace_dev = nullptr;
   if (trace_path) {
      KB2E_CUDA(c, cudaMalloc(&trace_dev, (size_t)c->num_sms * kTraceSlots * sizeof(unsigned long long)));
      KB2E_CUDA(c, cudaMemset(trace_dev, 0, (size_t)c->num_sms * kTraceSlots * sizeof(unsigned long long)));
      a.base.trace = trace_dev;
   }
   void* params[] = {&a};
   KB2E_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   if (cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(c->num_sms), dim3(kDistThreads), params, 0, c->stream); e != cudaSuccess) {
      d->broken = true;   // the peers will time out waiting for this rank
      if (trace_dev) cudaFree(trace_dev);
      return cuda_fail(c, e, "kb2e_dist_train_epochs: launch");
   }
   // barrier sequence of one launch: 1 opening + 5 per batch + 1 closing (status exchange)
   d->xcount += (uint32_t)d->world * (2u + 5u * (uint32_t)c->cfg.batches * (uint32_t)n_epochs);
   KB2E_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   std::vector<double> loss(n_epochs);
   unsigned long long cnt[5];
   uint32_t share_host[4] = {0, 0, 0, 0};
   KB2E_CUDA(c, cudaMemcpyAsync(loss.data(), c->loss_dev, (size_t)n_epochs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(cnt, c->counters, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(share_host, d->share, sizeof(share_host), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   if (trace_dev) {
      std::vector<unsigned long long> tr((size_t)c->num_sms * kTraceSlots);
      cudaMemcpy(tr.data(), trace_dev, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      cudaFree(trace_dev);
      std::string path = std::string(trace_path) + "." + std::to_string(d->rank);
      if (FILE* f = fopen(path.c_str(), "w")) {
         for (int bb = 0; bb < c->num_sms; bb++)
            for (int kk = 0; kk < kTraceSlots; kk++) fprintf(f, "%llu%c", tr[(size_t)bb * kTraceSlots + kk], kk + 1 == kTraceSlots ? '\n' : ' ');
         fclose(f);
      }
   }
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   c->tstats.kernel_ms += ms;
   c->tstats.launches += 1;
   c->tstats.samples += cnt[4] - d->samples_seen;   // samples whose head this rank owns (counted by the kernel)
   d->samples_seen = cnt[4];
   if (share_host[3]) {
      d->broken = true;
      return fail(c, KB2E_ERR_PEER, "kb2e_dist_train_epochs: a peer GPU did not arrive at a cross-GPU barrier within the time limit (its call failed or it "
                                    "never launched); the launch was abandoned on every rank -- results of this call are invalid");
   }
   if (share_host[2])
      return fail(c, KB2E_ERR_LIMIT, "kb2e_dist_train_epochs: the heads (or the remote rows) of a batch are too unevenly spread over the ranks for the "
                                     "exchange buffers; samples or updates were dropped on at least one rank -- results of this call are invalid on every rank");
   c->tstats.active = cnt[0];
   c->tstats.touched_ent = cnt[1];
   c->tstats.touched_rel = cnt[2];
   if (loss_per_epoch) memcpy(loss_per_epoch, loss.data(), (size_t)n_epochs * sizeof(double));
   return KB2E_OK;
}

void kb2e_dist_teardown(kb2e_ctx* c) {
   if (!c || !c->dist) return;
   cudaSetDevice(c->device);
   DistState* d = c->dist;
   for (int g = 0; g < d->world; g++)
      if (g != d->rank && d->peers[g]) cudaIpcCloseMemHandle(d->peers[g]);
   cudaFree(d->arena);
   cudaFree(d->stage);
   cudaFree(d->sflag);
   cudaFree(d->pairs);
   cudaFree(d->share);
   cudaFree(d->push_cnt);
   delete d;
   c->dist = nullptr;
}

}  // extern "C"
