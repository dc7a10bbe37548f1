// Filtered link-prediction ranking against all entities (exact fp64 path).
//
// Replaces (citations into eriq-augustine/KB2E):
//   EmbeddingEvaluation::run            common/evaluation.cpp:181-251  relation-major loop, sums, hits@10
//   EmbeddingEvaluation::evalCorruption common/evaluation.cpp:124-179  score all entities, sort, scan
//   cachedTripleEnergy + the N_E^2 cache common/evaluation.cpp:107-120,194-218 (unnecessary here)
//   transe/transh/transr::tripleEnergy   transe/transe.cpp:10-28, transh/transh.cpp:10-29, transr/transr.cpp:13-37
//
// No sort: rank = 1 + #{candidates with strictly smaller energy}; exact ties are counted separately
// (the reference's std::sort places the truth anywhere among its equals).  Energies are computed
// in fp64 with the reference's exact operation order -- per candidate a left-to-right sum over the
// dimensions of |(t - h) - r| (or its square), no FMA contraction (explicit __dadd_rn/__dmul_rn and
// -fmad=false) -- so `<` and `==` decide exactly as the reference's doubles do.
//
// All three models reduce to one kernel over a per-relation candidate matrix C (transposed,
// C^T[i][c], so that consecutive threads = consecutive candidates read consecutive addresses):
//   TransE  C = entity table                       (one slot for every relation)
//   TransH  C_r[c] = e_c - (w_r . e_c) w_r         (transh/transh.cpp:18-25)
//   TransR  C_r[c] = M_r^T e_c                     (transr/transr.cpp:20-25, work vectors zeroed)
// and energy(head-corruption by c) = sum_i f((C[t]_i - C[c]_i) - d_i), energy(tail-corruption by c)
// = sum_i f((C[c]_i - C[h]_i) - d_i) = sum_i f((V_i - C[c]_i) - d'_i) with V = C[fixed], d' = -d
// (negation is exact), f = |.| or (.)^2.
//
// Filtered rank (common/evaluation.cpp:161-163): the known-true neighbours of each query live in a
// device hash table keyed by (side, relation, fixed entity) -> CSR segment; their energies are
// re-scored and those ranked before the truth are subtracted from the raw rank.

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"
#include "internal.h"
#include "rank_f32.h"
#include "rank_tc.h"
#include "rank_transr.h"

struct RankState {
   // filter set as a hashed CSR, built on the device: entries sorted by (segment key, neighbour);
   // hash slot -> offset of the segment's first entry
   uint64_t* seg_key = nullptr;   // hash slots: key or kEmptyKey
   uint32_t* seg_val = nullptr;   // offset of the segment in ent_key / nbr
   uint64_t* ent_key = nullptr;   // sorted segment keys, one per entry
   int32_t* nbr = nullptr;        // neighbour entity ids, sorted within a segment (duplicates adjacent)
   uint32_t* seg_end = nullptr;   // hash slot -> one past the segment's last entry
   uint64_t seg_mask = 0;
   uint32_t n_ent = 0;
   int2* chunks = nullptr;        // (query, known-true neighbour) work items of the filter pass; neighbour -1 = nothing to score
   unsigned int chunk_cap = 0;    // entries needed by the whole test set (both sides): bound for any window
   unsigned int* chunk_count = nullptr;
   // candidate matrices
   double* ct0 = nullptr;         // entity table transposed [D][ld]
   double* pt = nullptr;          // projected slots [slots][D][ld]
   size_t pt_slots = 0;
   int ld = 0;
   // per-call query arrays
   int32_t* q_int = nullptr;      // fixed, truth, rel, side, slot, orig  (6 x cap)
   double* q_etrue = nullptr;
   int32_t* q_cnt = nullptr;      // less, eq, known_less, known_eq (4 x cap)
   int4* tiles = nullptr;         // (first query, nq, slot, unused)
   int64_t q_cap = 0, tile_cap = 0;
   int32_t* slot_rel = nullptr;
   int64_t slot_cap = 0;
   unsigned long long* sums = nullptr;
   int32_t* out = nullptr;        // 4 x cap results in original order
   std::vector<cudaEvent_t> pass_ev;   // two per pass, around the all-candidates kernel
   int32_t* ids = nullptr;        // h | t | r columns of the known triples (test first), kept for the device query builder
   size_t ids_n = 0;
   uint64_t* key_tmp = nullptr;   // sort scratch (grow-only, like every buffer of the filter set)
   int32_t* val_tmp = nullptr;
   char* cub_tmp = nullptr;
   size_t cap_seg_key = 0, cap_seg_val = 0, cap_seg_end = 0, cap_ent_key = 0, cap_nbr = 0, cap_ids = 0, cap_key_tmp = 0,
          cap_val_tmp = 0, cap_cub_tmp = 0, cap_chunks = 0;
   kb2e::TcState tc;
   kb2e::F32State f32;
   kb2e::TrpState trp;
   // the filter pass depends only on the queries' exact energies, not on the all-candidates kernel: in a single-pass call
   // it runs beside that kernel on a second stream (fork after E_true, join before finalize)
   cudaStream_t side = nullptr;
   cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
   uint64_t ct_epoch = 0;   // tables_epoch the transposed entity table was made for
};

namespace kb2e {

constexpr int kQT = 16;          // queries per tile
constexpr int kRankThreads = 256;

__host__ __device__ __forceinline__ uint64_t seg_key_of(int side, int rel, int fixed) {
   return ((uint64_t)side << 40) | ((uint64_t)(uint32_t)rel << 24) | (uint64_t)(uint32_t)fixed;
}

// ---- candidate matrices ----------------------------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int D, int ld) {
   __shared__ double tile[32][33];
   int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
   for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      int r = r0 + j, c = c0 + threadIdx.x;
      tile[j][threadIdx.x] = (r < rows && c < D) ? src[(size_t)r * D + c] : 0.0;
   }
   __syncthreads();
   for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      int c = c0 + j, r = r0 + threadIdx.x;
      if (c < D && r < ld) dst[(size_t)c * ld + r] = (r < rows) ? tile[threadIdx.x][j] : 0.0;
   }
}

// TransH: p_i = e_i - (sum_k w_k e_k) * w_i, the sum taken left to right (transh/transh.cpp:18-25).
__global__ void project_transh_kernel(const double* __restrict__ ct0, const double* __restrict__ w64,
                                      const int32_t* __restrict__ slot_rel, double* __restrict__ pt, int nE, int D, int ld) {
   int c = blockIdx.x * blockDim.x + threadIdx.x;
   int slot = blockIdx.y;
   if (c >= nE) return;
   const double* w = w64 + (size_t)slot_rel[slot] * D;
   double s = 0.0;
   for (int i = 0; i < D; i++) s = __dadd_rn(s, __dmul_rn(__ldg(w + i), ct0[(size_t)i * ld + c]));
   double* out = pt + (size_t)slot * D * ld;
   for (int i = 0; i < D; i++) out[(size_t)i * ld + c] = __dsub_rn(ct0[(size_t)i * ld + c], __dmul_rn(s, __ldg(w + i)));
}

// TransR: p_i = sum_j M[j][i] * e_j, j ascending from a zero accumulator (transr/transr.cpp:20-25).
__global__ void project_transr_kernel(const double* __restrict__ ct0, const double* __restrict__ w64,
                                      const int32_t* __restrict__ slot_rel, double* __restrict__ pt, int nE, int D, int ld) {
   extern __shared__ double sM[];  // [D][D]
   int slot = blockIdx.y;
   const double* M = w64 + (size_t)slot_rel[slot] * D * D;
   for (int k = threadIdx.x; k < D * D; k += blockDim.x) sM[k] = M[k];
   __syncthreads();
   int c = blockIdx.x * blockDim.x + threadIdx.x;
   if (c >= nE) return;
   double* out = pt + (size_t)slot * D * ld;
   // eight output dims per pass over j: one (L1-cached) load of e_j feeds eight multiply-adds; every output is still the
   // plain left-to-right sum over j, so the values are bit-identical to the reference's
   constexpr int W = 8;
   for (int i0 = 0; i0 < D; i0 += W) {
      double acc[W];
#pragma unroll
      for (int k = 0; k < W; k++) acc[k] = 0.0;
      for (int j = 0; j < D; j++) {
         const double e = ct0[(size_t)j * ld + c];
         const double* m = sM + j * D + i0;
#pragma unroll
         for (int k = 0; k < W; k++)
            if (i0 + k < D) acc[k] = __dadd_rn(acc[k], __dmul_rn(m[k], e));
      }
#pragma unroll
      for (int k = 0; k < W; k++)
         if (i0 + k < D) out[(size_t)(i0 + k) * ld + c] = acc[k];
   }
}

// ---- exact energy of one candidate (any thread; strided reads) -------------------------------------
template <int L2>
__device__ __forceinline__ double exact_energy(const double* __restrict__ ct, int ld, int D, int fixed, int cand,
                                               const double* __restrict__ d, double dsign) {
   double acc = 0.0;
   for (int i = 0; i < D; i++) {
      double u = __dsub_rn(ct[(size_t)i * ld + fixed], ct[(size_t)i * ld + cand]);
      double v = __dsub_rn(u, dsign * d[i]);
      acc = L2 ? __dadd_rn(acc, __dmul_rn(v, v)) : __dadd_rn(acc, fabs(v));
   }
   return acc;
}

struct RankArgs {
   const double* ct;      // slot 0 base
   const double* rel64;
   const int32_t* q_fixed;
   const int32_t* q_truth;
   const int32_t* q_rel;
   const int32_t* q_side;
   const int32_t* q_slot;
   double* q_etrue;
   int32_t* q_cnt;        // [4][nq]: less, eq, known_less, known_eq
   const int4* tiles;
   const uint64_t* seg_key;
   const uint32_t* seg_val;
   const uint64_t* ent_key;
   const int32_t* nbr;
   uint64_t seg_mask;
   uint32_t n_ent;
   long long nq;        // total queries of this call (row stride of q_cnt)
   long long q_begin;   // window of queries handled by etrue_kernel / filter_kernel
   long long q_end;
   int nE, D, ld, splits;
};

template <int L2>
__global__ void etrue_kernel(const RankArgs a) {
   long long q = a.q_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (q >= a.q_end) return;
   const double* ct = a.ct + (size_t)a.q_slot[q] * a.D * a.ld;
   const double* d = a.rel64 + (size_t)a.q_rel[q] * a.D;
   a.q_etrue[q] = exact_energy<L2>(ct, a.ld, a.D, a.q_fixed[q], a.q_truth[q], d, a.q_side[q] ? -1.0 : 1.0);
}

// Exact re-score of the (query, candidate) pairs the tensor-core pre-filter could not decide
// (rank_tc.cu): same arithmetic and order as exact_energy, rows read from the row-major fp64 table.
// The band size lives on the device (*band_count, capped by band_cap): the launch is a fixed grid striding over it,
// so the host never waits for the tensor-core kernel before enqueueing the rest of the call.
template <int L2>
__global__ void recheck_kernel(const int2* __restrict__ band, const unsigned int* __restrict__ band_count, unsigned int band_cap,
                               const double* __restrict__ ent64, const double* __restrict__ rel64, const int32_t* q_fixed,
                               const int32_t* q_truth, const int32_t* q_rel, const int32_t* q_side, const double* q_etrue,
                               int32_t* q_cnt, long long nq, int D) {
   const unsigned int n = min(*band_count, band_cap);
   for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
      const int q = band[k].x, c = band[k].y;
      if (c == q_truth[q]) continue;
      const double* v = ent64 + (size_t)q_fixed[q] * D;
      const double* e = ent64 + (size_t)c * D;
      const double* d = rel64 + (size_t)q_rel[q] * D;
      const double dsign = q_side[q] ? -1.0 : 1.0;
      double acc = 0.0;
      for (int i = 0; i < D; i++) {
         double u = __dsub_rn(v[i], e[i]);
         double w = __dsub_rn(u, dsign * d[i]);
         acc = L2 ? __dadd_rn(acc, __dmul_rn(w, w)) : __dadd_rn(acc, fabs(w));
      }
      const double et = q_etrue[q];
      if (acc < et) atomicAdd(q_cnt + q, 1);
      else if (acc == et) atomicAdd(q_cnt + nq + q, 1);
   }
}

// TransE: the query arrays in original test order, built on the device from the resident test triples
// (fixed entity, true answer, relation, side, slot = 0, original index); query 2 i is the head corruption of
// test triple first + i, query 2 i + 1 its tail corruption (common/evaluation.cpp:230-238).
__global__ void build_queries_kernel(const int32_t* __restrict__ th, const int32_t* __restrict__ tt, const int32_t* __restrict__ tr,
                                     long long first, long long count, int32_t* q_int) {
   const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   const long long nq = 2 * count;
   if (q >= nq) return;
   const long long i = first + (q >> 1);
   const int side = (int)(q & 1);
   const int h = th[i], t = tt[i];
   q_int[q] = side == 0 ? t : h;
   q_int[nq + q] = side == 0 ? h : t;
   q_int[2 * nq + q] = tr[i];
   q_int[3 * nq + q] = side;
   q_int[4 * nq + q] = 0;
   q_int[5 * nq + q] = (int32_t)q;
}

// ---- the all-candidates kernel ---------------------------------------------------------------------
// grid.x = query tiles, grid.y = candidate splits; each thread owns one candidate per 256-wide step
// and keeps kQT running energies, reading the tile's (V_i, d'_i) pairs from shared memory.
template <int L2>
__global__ void __launch_bounds__(kRankThreads) rank_exact_kernel(const RankArgs a) {
   extern __shared__ double2 s_vd[];  // [D][kQT]
   __shared__ double s_et[kQT];
   __shared__ int s_truth[kQT];
   __shared__ int s_cnt[kQT][2];
   const int4 tile = a.tiles[blockIdx.x];
   const int q0 = tile.x, nq = tile.y;
   const double* ct = a.ct + (size_t)tile.z * a.D * a.ld;
   const int D = a.D, ld = a.ld;
   for (int k = threadIdx.x; k < D * kQT; k += blockDim.x) {
      int i = k / kQT, q = k % kQT;
      double2 vd = make_double2(0.0, 0.0);
      if (q < nq) {
         int gq = q0 + q;
         vd.x = ct[(size_t)i * ld + a.q_fixed[gq]];
         double dv = a.rel64[(size_t)a.q_rel[gq] * D + i];
         vd.y = a.q_side[gq] ? -dv : dv;
      }
      s_vd[k] = vd;
   }
   if (threadIdx.x < kQT) {
      int q = threadIdx.x;
      s_et[q] = q < nq ? a.q_etrue[q0 + q] : 0.0;
      s_truth[q] = q < nq ? a.q_truth[q0 + q] : -1;
      s_cnt[q][0] = 0;
      s_cnt[q][1] = 0;
   }
   __syncthreads();
   // candidate range of this split, in steps of blockDim.x
   const int steps = (a.nE + kRankThreads - 1) / kRankThreads;
   const int s_begin = (int)((long long)steps * blockIdx.y / a.splits);
   const int s_end = (int)((long long)steps * (blockIdx.y + 1) / a.splits);
   uint32_t cnt[kQT];
#pragma unroll
   for (int q = 0; q < kQT; q++) cnt[q] = 0;
   for (int step = s_begin; step < s_end; step++) {
      const int c = step * kRankThreads + threadIdx.x;
      const bool valid = c < a.nE;
      const int cc = valid ? c : 0;
      double acc[kQT];
#pragma unroll
      for (int q = 0; q < kQT; q++) acc[q] = 0.0;
#pragma unroll 2
      for (int i = 0; i < D; i++) {
         const double ci = ct[(size_t)i * ld + cc];
         const double2* vd = s_vd + i * kQT;
#pragma unroll
         for (int q = 0; q < kQT; q++) {
            double2 p = vd[q];
            double v = __dsub_rn(__dsub_rn(p.x, ci), p.y);
            acc[q] = L2 ? __dadd_rn(acc[q], __dmul_rn(v, v)) : __dadd_rn(acc[q], fabs(v));
         }
      }
      if (valid) {
#pragma unroll
         for (int q = 0; q < kQT; q++) {
            double et = s_et[q];
            cnt[q] += (acc[q] < et ? 1u : 0u) + ((acc[q] == et && c != s_truth[q]) ? 0x10000u : 0u);
         }
      }
   }
   // block reduction of the packed (less | eq << 16) counters; per-thread counts are < 2^16
#pragma unroll
   for (int q = 0; q < kQT; q++) {
      uint32_t lo = cnt[q] & 0xffffu, hi = cnt[q] >> 16;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
         lo += __shfl_xor_sync(0xffffffffu, lo, o);
         hi += __shfl_xor_sync(0xffffffffu, hi, o);
      }
      if ((threadIdx.x & 31) == 0) {
         if (lo) atomicAdd(&s_cnt[q][0], (int)lo);
         if (hi) atomicAdd(&s_cnt[q][1], (int)hi);
      }
   }
   __syncthreads();
   if (threadIdx.x < nq) {
      int q = threadIdx.x;
      if (a.splits == 1) {
         a.q_cnt[0 * a.nq + q0 + q] = s_cnt[q][0];
         a.q_cnt[1 * a.nq + q0 + q] = s_cnt[q][1];
      } else {
         if (s_cnt[q][0]) atomicAdd(a.q_cnt + 0 * a.nq + q0 + q, s_cnt[q][0]);
         if (s_cnt[q][1]) atomicAdd(a.q_cnt + 1 * a.nq + q0 + q, s_cnt[q][1]);
      }
   }
}

// ---- filter set construction (device) --------------------------------------------------------------
// Every known triple (h, t, r) contributes two entries: head-corruption segment (0, r, t) -> h and
// tail-corruption segment (1, r, h) -> t.
__global__ void filter_entries_kernel(const int32_t* h, const int32_t* t, const int32_t* r, long long n,
                                      uint64_t* key, int32_t* val) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   key[2 * i] = seg_key_of(0, r[i], t[i]);
   val[2 * i] = h[i];
   key[2 * i + 1] = seg_key_of(1, r[i], h[i]);
   val[2 * i + 1] = t[i];
}

__global__ void segment_hash_kernel(const uint64_t* ent_key, uint32_t n, uint64_t* slots, uint32_t* vals, uint64_t mask) {
   uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   uint64_t k = ent_key[i];
   if (i > 0 && ent_key[i - 1] == k) return;  // not the first entry of its segment
   uint64_t p = mix64(k) & mask;
   while (true) {
      unsigned long long prev = atomicCAS((unsigned long long*)(slots + p), (unsigned long long)kEmptyKey, (unsigned long long)k);
      if (prev == kEmptyKey) { vals[p] = i; return; }
      p = (p + 1) & mask;
   }
}

// last entry of every segment -> seg_end[slot of its key] = index one past it (runs after segment_hash_kernel)
__global__ void segment_end_kernel(const uint64_t* ent_key, uint32_t n, const uint64_t* slots, uint32_t* ends, uint64_t mask) {
   uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   const uint64_t k = ent_key[i];
   if (i + 1 < n && ent_key[i + 1] == k) return;  // not the last entry of its segment
   uint64_t p = mix64(k) & mask;
   while (slots[p] != k) p = (p + 1) & mask;
   ends[p] = i + 1;
}

// ---- filter adjustment ---------------------------------------------------------------------------------
// The known-true neighbours of a query are one segment of the sorted entry list.  Segment lengths are heavy-tailed
// (FB15k-shape planted KG: 833,075 neighbours for 118,142 queries -- median 1, 99th percentile 126, maximum 775), and one
// exact energy is a chain of D dependent fp64 additions.  Work is therefore flattened to ONE (query, neighbour) PAIR per
// thread: filter_plan_kernel looks the segments up (hash), reserves room with a warp scan and one atomic per warp and
// writes the pairs -- the warp walks its 32 queries together, so a 775-neighbour segment costs 25 coalesced passes, not
// 775 iterations of one lane -- dropping the truth itself and triples listed twice; filter_pairs_kernel (persistent
// grid) scores one pair per thread with every lane busy.  (The previous scheme, one <= 32-neighbour chunk per warp, left
// a warp with one or two active lanes for the median segment: 0.31 ms at FB15k shape.)  No host step: the pair count stays
// on the device.
struct SegRef { uint32_t off, len; };

__device__ __forceinline__ SegRef find_segment(const uint64_t* __restrict__ seg_key, const uint32_t* __restrict__ seg_val,
                                               const uint32_t* __restrict__ seg_end, uint64_t mask, uint64_t key) {
   uint64_t slot = mix64(key) & mask;
   while (true) {
      const uint64_t k = __ldg(seg_key + slot);
      if (k == key) { const uint32_t o = __ldg(seg_val + slot); return SegRef{o, __ldg(seg_end + slot) - o}; }
      if (k == kEmptyKey) return SegRef{0u, 0u};
      slot = (slot + 1) & mask;
   }
}

// upper bound of the pairs any kb2e_rank window can need: all test triples, both sides (runs once per filter build)
__global__ void count_chunks_kernel(const int32_t* __restrict__ th, const int32_t* __restrict__ tt, const int32_t* __restrict__ tr,
                                    long long n_test, const uint64_t* seg_key, const uint32_t* seg_val, const uint32_t* seg_end,
                                    uint64_t mask, unsigned long long* total) {
   const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   unsigned int n = 0;
   if (q < 2 * n_test) {
      const long long i = q >> 1;
      const int side = (int)(q & 1);
      n = find_segment(seg_key, seg_val, seg_end, mask, seg_key_of(side, tr[i], side == 0 ? tt[i] : th[i])).len;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
   if ((threadIdx.x & 31) == 0 && n) atomicAdd(total, (unsigned long long)n);
}

__global__ void filter_plan_kernel(const RankArgs a, const uint32_t* __restrict__ seg_end, int2* pairs, unsigned int* pair_count,
                                   unsigned int pair_cap) {
   const long long q = a.q_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
   const int lane = threadIdx.x & 31;
   SegRef sr{0u, 0u};
   int truth = -1;
   if (q < a.q_end) {
      sr = find_segment(a.seg_key, a.seg_val, seg_end, a.seg_mask, seg_key_of(a.q_side[q], a.q_rel[q], a.q_fixed[q]));
      truth = a.q_truth[q];
   }
   unsigned int x = sr.len;   // inclusive warp scan
#pragma unroll
   for (int o = 1; o < 32; o <<= 1) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
   }
   const unsigned int total = __shfl_sync(0xffffffffu, x, 31);
   if (total == 0) return;
   unsigned int base = 0;
   if (lane == 0) base = atomicAdd(pair_count, total);
   base = __shfl_sync(0xffffffffu, base, 0) + (x - sr.len);
   for (int src = 0; src < 32; src++) {
      const unsigned int len = __shfl_sync(0xffffffffu, sr.len, src);
      if (len == 0) continue;
      const unsigned int off = __shfl_sync(0xffffffffu, sr.off, src), out = __shfl_sync(0xffffffffu, base, src);
      const int tq = __shfl_sync(0xffffffffu, truth, src);
      const int qq = (int)(q - lane + src);
      for (unsigned int k = lane; k < len; k += 32) {
         int c = __ldg(a.nbr + off + k);
         // the truth itself is not a competitor; the same triple listed twice (e.g. in train and valid) counts once
         if (c == tq || (k > 0 && __ldg(a.nbr + off + k - 1) == c)) c = -1;
         if (out + k < pair_cap) pairs[out + k] = make_int2(qq, c);
      }
   }
}

// ROWS = true (TransE): the candidate matrix IS the entity table, so the energies are taken from the row-major
// fp64 table (contiguous 8*D-byte rows) instead of the transposed copy; arithmetic and order are exact_energy's.
template <int L2, bool ROWS>
__global__ void filter_pairs_kernel(const RankArgs a, const double* __restrict__ ent64, const int2* __restrict__ pairs,
                                    const unsigned int* __restrict__ pair_count, unsigned int pair_cap) {
   const unsigned int n = min(*pair_count, pair_cap);
   for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
      const int2 pr = __ldg(pairs + k);
      const int c = pr.y;
      if (c < 0) continue;
      const long long q = pr.x;
      const int fixed = a.q_fixed[q];
      const double dsign = a.q_side[q] ? -1.0 : 1.0;
      const double* d = a.rel64 + (size_t)a.q_rel[q] * a.D;
      double e;
      if (ROWS) {
         const double* v = ent64 + (size_t)fixed * a.D;
         const double* x = ent64 + (size_t)c * a.D;
         e = 0.0;
#pragma unroll 8
         for (int i = 0; i < a.D; i++) {
            const double t = __dsub_rn(__dsub_rn(__ldg(v + i), __ldg(x + i)), dsign * __ldg(d + i));
            e = L2 ? __dadd_rn(e, __dmul_rn(t, t)) : __dadd_rn(e, fabs(t));
         }
      } else {
         const double* ct = a.ct + (size_t)a.q_slot[q] * a.D * a.ld;
         e = exact_energy<L2>(ct, a.ld, a.D, fixed, c, d, dsign);
      }
      const double et = a.q_etrue[q];
      if (e < et) atomicAdd(a.q_cnt + 2 * a.nq + q, 1);
      else if (e == et) atomicAdd(a.q_cnt + 3 * a.nq + q, 1);
   }
}

// ranks in original order + the four sums (common/evaluation.cpp:169-178)
__global__ void finalize_kernel(const int32_t* q_cnt, const int32_t* q_orig, long long nq, int32_t* out, unsigned long long* sums) {
   long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   unsigned long long v[4] = {0, 0, 0, 0};
   if (q < nq) {
      int raw = 1 + q_cnt[q], ties = q_cnt[nq + q];
      int filt = raw - q_cnt[2 * nq + q], fties = ties - q_cnt[3 * nq + q];
      int o = q_orig[q];
      out[0 * nq + o] = raw;
      out[1 * nq + o] = filt;
      out[2 * nq + o] = ties;
      out[3 * nq + o] = fties;
      v[0] = raw; v[1] = filt; v[2] = raw <= 10; v[3] = filt <= 10;
   }
#pragma unroll
   for (int k = 0; k < 4; k++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(sums + k, v[k]);
   }
}

// one thread per triple, exact energies (precision = 1 of kb2e_score)
template <int L2>
__global__ void score64_kernel(const double* ct, const double* rel64, const int32_t* slot_of_rel, int D, int ld,
                               const int32_t* h, const int32_t* t, const int32_t* r, long long n, double* out) {
   long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (k >= n) return;
   const double* c = ct + (size_t)(slot_of_rel ? slot_of_rel[r[k]] : 0) * D * ld;
   // head-corruption form with the true head as candidate: (C[t] - C[h]) - d
   out[k] = exact_energy<L2>(c, ld, D, t[k], h[k], rel64 + (size_t)r[k] * D, 1.0);
}

// ---- host ------------------------------------------------------------------------------------------
static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

static int ensure_state(kb2e_ctx* c) {
   if (!c->rank) c->rank = new RankState();
   RankState* s = c->rank;
   s->ld = ((c->nE + 31) / 32) * 32;
   if (!s->ct0) KB2E_CUDA(c, pool_alloc(c, &s->ct0, (size_t)c->D * s->ld * sizeof(double)));
   if (!s->sums) KB2E_CUDA(c, pool_alloc(c, &s->sums, 4 * sizeof(unsigned long long)));
   return KB2E_OK;
}

// All buffers of the filter set are grow-only (capacities in RankState): rebuilding the CSR for a new test / filter
// set of the same size -- the per-step pattern of an evaluation loop -- allocates nothing.
template <typename T>
static int grow(kb2e_ctx* c, T** p, size_t* cap, size_t need) {
   if (need <= *cap && *p) return KB2E_OK;
   pool_free(c, *p);
   *p = nullptr;
   *cap = 0;
   KB2E_CUDA(c, pool_alloc(c, p, std::max<size_t>(1, need) * sizeof(T)));
   *cap = std::max<size_t>(1, need);
   return KB2E_OK;
}

static int build_filter(kb2e_ctx* c) {
   RankState* s = c->rank;
   if (!c->filter_dirty && s->seg_key) return KB2E_OK;
   // known triples = test + filter (common/evaluation.cpp:59-61)
   const size_t nt = c->test_h.size(), nf = c->filt_n, n = nt + nf;
   if (2 * n >= (1ull << 32)) return fail(c, KB2E_ERR_LIMIT, "filter set larger than 2^31 triples");
   s->n_ent = (uint32_t)(2 * n);
   uint64_t slots = 1024;
   while (slots < 4 * n) slots <<= 1;  // <= 2n segments, load <= 0.5
   s->seg_mask = slots - 1;
   int rc;
   if ((rc = grow(c, &s->seg_key, &s->cap_seg_key, slots)) || (rc = grow(c, &s->seg_val, &s->cap_seg_val, slots)) ||
       (rc = grow(c, &s->seg_end, &s->cap_seg_end, slots)) || (rc = grow(c, &s->ent_key, &s->cap_ent_key, 2 * n)) ||
       (rc = grow(c, &s->nbr, &s->cap_nbr, 2 * n)) || (rc = grow(c, &s->ids, &s->cap_ids, 3 * n)) ||
       (rc = grow(c, &s->key_tmp, &s->cap_key_tmp, 2 * n)) || (rc = grow(c, &s->val_tmp, &s->cap_val_tmp, 2 * n)))
      return rc;
   if (!s->chunk_count) KB2E_CUDA(c, pool_alloc(c, &s->chunk_count, 2 * sizeof(unsigned long long)));
   s->ids_n = n;
   KB2E_CUDA(c, cudaMemsetAsync(s->seg_key, 0xff, slots * sizeof(uint64_t), c->stream));
   if (n == 0) { c->filter_dirty = false; return KB2E_OK; }
   int32_t* ids = s->ids;      // h | t | r columns of test then filter triples
   uint64_t* key_tmp = s->key_tmp;
   int32_t* val_tmp = s->val_tmp;
   const std::vector<int32_t>* test_cols[3] = {&c->test_h, &c->test_t, &c->test_r};
   for (int k = 0; k < 3; k++) {
      if (nt) KB2E_CUDA(c, cudaMemcpyAsync(ids + k * n, test_cols[k]->data(), nt * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      if (nf) KB2E_CUDA(c, cudaMemcpyAsync(ids + k * n + nt, c->filt_dev + k * c->filt_cap, nf * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
   }
   filter_entries_kernel<<<nblk((long long)n, 256), 256, 0, c->stream>>>(ids, ids + n, ids + 2 * n, (long long)n, key_tmp, val_tmp);
   // stable LSD radix sort: by neighbour (24 bits), then by segment key (41 bits) -> (key, neighbour) order
   size_t bytes1 = 0, bytes2 = 0;
   cub::DeviceRadixSort::SortPairs(nullptr, bytes1, val_tmp, s->nbr, key_tmp, s->ent_key, (int)(2 * n), 0, kEntityBits, c->stream);
   cub::DeviceRadixSort::SortPairs(nullptr, bytes2, s->ent_key, key_tmp, s->nbr, val_tmp, (int)(2 * n), 0, 41, c->stream);
   if ((rc = grow(c, &s->cub_tmp, &s->cap_cub_tmp, std::max(bytes1, bytes2)))) return rc;
   KB2E_CUDA(c, cub::DeviceRadixSort::SortPairs(s->cub_tmp, bytes1, val_tmp, s->nbr, key_tmp, s->ent_key, (int)(2 * n), 0, kEntityBits, c->stream));
   KB2E_CUDA(c, cub::DeviceRadixSort::SortPairs(s->cub_tmp, bytes2, s->ent_key, key_tmp, s->nbr, val_tmp, (int)(2 * n), 0, 41, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(s->ent_key, key_tmp, 2 * n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(s->nbr, val_tmp, 2 * n * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
   segment_hash_kernel<<<nblk((long long)(2 * n), 256), 256, 0, c->stream>>>(s->ent_key, s->n_ent, s->seg_key, s->seg_val, s->seg_mask);
   segment_end_kernel<<<nblk((long long)(2 * n), 256), 256, 0, c->stream>>>(s->ent_key, s->n_ent, s->seg_key, s->seg_end, s->seg_mask);
   // (query, neighbour) pairs the filter pass of the whole test set needs (kb2e_rank windows need at most as many)
   unsigned long long* total_dev = reinterpret_cast<unsigned long long*>(s->chunk_count) + 1;
   unsigned long long total = 0;
   KB2E_CUDA(c, cudaMemsetAsync(total_dev, 0, sizeof(unsigned long long), c->stream));
   if (nt) count_chunks_kernel<<<nblk(2 * (long long)nt, 256), 256, 0, c->stream>>>(ids, ids + n, ids + 2 * n, (long long)nt, s->seg_key, s->seg_val,
                                                                                  s->seg_end, s->seg_mask, total_dev);
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaMemcpyAsync(&total, total_dev, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   if (total >= (1ull << 31)) return fail(c, KB2E_ERR_LIMIT, "filter pass needs more than 2^31 (query, neighbour) pairs");
   s->chunk_cap = (unsigned int)std::max<unsigned long long>(1, total);
   if ((rc = grow(c, &s->chunks, &s->cap_chunks, (size_t)s->chunk_cap))) return rc;
   c->filter_dirty = false;
   return KB2E_OK;
}

static int ensure_query_buffers(kb2e_ctx* c, int64_t nq) {
   RankState* s = c->rank;
   if (nq <= s->q_cap) return KB2E_OK;
   pool_free(c, s->q_int); pool_free(c, s->q_etrue); pool_free(c, s->q_cnt); pool_free(c, s->out);
   s->q_int = nullptr; s->q_etrue = nullptr; s->q_cnt = nullptr; s->out = nullptr;
   s->q_cap = 0;
   KB2E_CUDA(c, pool_alloc(c, &s->q_int, (size_t)nq * 6 * sizeof(int32_t)));
   KB2E_CUDA(c, pool_alloc(c, &s->q_etrue, (size_t)nq * sizeof(double)));
   KB2E_CUDA(c, pool_alloc(c, &s->q_cnt, (size_t)nq * 4 * sizeof(int32_t)));
   KB2E_CUDA(c, pool_alloc(c, &s->out, (size_t)nq * 4 * sizeof(int32_t)));
   s->q_cap = nq;
   return KB2E_OK;
}

// Project the relations in `rels` into slots (TransH / TransR); TransE uses ct0 directly.
static int project(kb2e_ctx* c, const std::vector<int32_t>& rels) {
   RankState* s = c->rank;
   size_t slots = rels.size();
   if (slots > s->pt_slots) {
      pool_free(c, s->pt);
      s->pt = nullptr;
      KB2E_CUDA(c, pool_alloc(c, &s->pt, slots * (size_t)c->D * s->ld * sizeof(double)));
      s->pt_slots = slots;
   }
   if ((int64_t)slots > s->slot_cap) {
      pool_free(c, s->slot_rel);
      s->slot_rel = nullptr;
      KB2E_CUDA(c, pool_alloc(c, &s->slot_rel, slots * sizeof(int32_t)));
      s->slot_cap = (int64_t)slots;
   }
   KB2E_CUDA(c, cudaMemcpyAsync(s->slot_rel, rels.data(), slots * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   dim3 grid(nblk(c->nE, 128), (unsigned)slots);
   if (c->cfg.model == KB2E_MODEL_TRANSH) {
      project_transh_kernel<<<grid, 128, 0, c->stream>>>(s->ct0, c->w64, s->slot_rel, s->pt, c->nE, c->D, s->ld);
   } else {
      size_t smem = (size_t)c->D * c->D * sizeof(double);
      if (smem > 200 * 1024) return fail(c, KB2E_ERR_LIMIT, "TransR ranking supports embedding sizes up to 160");
      KB2E_CUDA(c, cudaFuncSetAttribute(project_transr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      project_transr_kernel<<<grid, 128, smem, c->stream>>>(s->ct0, c->w64, s->slot_rel, s->pt, c->nE, c->D, s->ld);
   }
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

static int prepare_tables(kb2e_ctx* c) {
   int rc = ensure_state(c);
   if (rc) return rc;
   rc = ensure64(c);
   if (rc) return rc;
   RankState* s = c->rank;
   if (s->ct_epoch == c->tables_epoch) return KB2E_OK;   // the transposed copy is current: tables unchanged since the last call
   dim3 grid(nblk(s->ld, 32), nblk(c->D, 32));
   transpose_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(c->ent64, s->ct0, c->nE, c->D, s->ld);
   KB2E_CUDA(c, cudaGetLastError());
   s->ct_epoch = c->tables_epoch;
   return KB2E_OK;
}

int rank_run(kb2e_ctx* c, int64_t first, int64_t count, int32_t* raw_rank, int32_t* filt_rank,
             int32_t* raw_ties, int32_t* filt_ties, int64_t sums[4]) {
   if (first < 0 || count < 0 || first + count > (int64_t)c->test_h.size())
      return fail(c, KB2E_ERR_ARG, "rank window outside the test triples");
   if (sums) sums[0] = sums[1] = sums[2] = sums[3] = 0;
   if (count == 0) return KB2E_OK;
   // KB2E_RANK_TIMING=1: host-side phase times of this call on stderr (tuning aid; adds stream synchronisations)
   const bool timing = getenv("KB2E_RANK_TIMING") != nullptr;
   auto t_last = std::chrono::steady_clock::now();
   auto lap = [&](const char* what) {
      if (!timing) return;
      cudaStreamSynchronize(c->stream);
      auto now = std::chrono::steady_clock::now();
      fprintf(stderr, "[kb2e_rank] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
   };
   int rc = prepare_tables(c);
   if (rc) return rc;
   lap("fp64 tables + transpose");
   rc = build_filter(c);
   if (rc) return rc;
   lap("filter CSR build");
   RankState* s = c->rank;
   const bool per_rel = c->cfg.model != KB2E_MODEL_TRANSE;
   const bool l2 = c->cfg.model != KB2E_MODEL_TRANSH && c->cfg.distance == KB2E_DISTANCE_L2;
   const int64_t nq = 2 * count;
   const bool use_tc = tc_supported(c);
   // every other model / distance: fp32 CUDA-core pre-filter + exact recheck (rank_f32.cu) unless the exact kernel is forced
   const bool use_f32 = !use_tc && !(c->cfg.flags & KB2E_FLAG_RANK_EXACT_ONLY);
   // TransR: candidates projected by the tensor cores (fp32, bounded error), exact projections only on demand (rank_transr.cu)
   const bool use_trp = use_f32 && trp_supported(c);
   const int tile_q = use_f32 ? kF32QueriesPerTile : kQT;

   rc = ensure_query_buffers(c, nq);
   if (rc) return rc;
   KB2E_CUDA(c, cudaMemsetAsync(s->q_cnt, 0, (size_t)nq * 4 * sizeof(int32_t), c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(s->sums, 0, 4 * sizeof(unsigned long long), c->stream));

   struct Pass { int64_t q_begin, q_end; size_t tile_begin, tile_end; std::vector<int32_t> rels; };
   std::vector<Pass> passes(1);
   std::vector<int4> tiles;  // (first query, #queries, slot, -)
   if (!per_rel) {
      // TransE: one pass over the entity table, queries in test order, built on the device
      build_queries_kernel<<<nblk(nq, 256), 256, 0, c->stream>>>(s->ids, s->ids + s->ids_n, s->ids + 2 * s->ids_n, first, count, s->q_int);
      KB2E_CUDA(c, cudaGetLastError());
      passes[0].q_begin = 0;
      passes[0].q_end = nq;
      passes[0].tile_begin = 0;
      if (!use_tc)
         for (int64_t q = 0; q < nq; q += tile_q) tiles.push_back(make_int4((int)q, (int)std::min<int64_t>(tile_q, nq - q), 0, 0));
      passes[0].tile_end = tiles.size();
   } else {
      // query order: grouped by relation (slot) for TransH/TransR
      std::vector<int64_t> order(count);
      std::iota(order.begin(), order.end(), first);
      std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return c->test_r[x] < c->test_r[y]; });
      // slots per pass bounded by a memory budget for the projected matrices
      const size_t slot_bytes = (size_t)c->D * s->ld * (use_trp ? sizeof(float) : sizeof(double));
      const size_t budget = (size_t)16 << 30;
      const size_t max_slots = std::max<size_t>(1, budget / slot_bytes);
      // query arrays (SoA, nq each): fixed entity, true answer, relation, side, slot, original index
      std::vector<int32_t> qi(6 * (size_t)nq);
      int32_t* qf = qi.data();
      int32_t* qt = qf + nq;
      int32_t* qr = qt + nq;
      int32_t* qs = qr + nq;
      int32_t* qslot = qs + nq;
      int32_t* qo = qslot + nq;
      int64_t q = 0, tile_start = 0;
      int cur_slot = 0;
      passes[0].q_begin = 0;
      passes[0].tile_begin = 0;
      auto close_tile = [&](int64_t end) {
         if (end > tile_start) tiles.push_back(make_int4((int)tile_start, (int)(end - tile_start), cur_slot, 0));
         tile_start = end;
      };
      for (int64_t i = 0; i < count; i++) {
         const int64_t ti = order[i];
         const int32_t h = c->test_h[ti], t = c->test_t[ti], r = c->test_r[ti];
         Pass& cur = passes.back();
         if (cur.rels.empty() || cur.rels.back() != r) {
            close_tile(q);
            if (cur.rels.size() == max_slots) {
               cur.q_end = q;
               cur.tile_end = tiles.size();
               Pass next;
               next.q_begin = q;
               next.tile_begin = tiles.size();
               passes.push_back(next);
            }
            passes.back().rels.push_back(r);
            cur_slot = (int)passes.back().rels.size() - 1;
         }
         for (int side = 0; side < 2; side++) {
            if (q - tile_start == tile_q) close_tile(q);
            qf[q] = side == 0 ? t : h;  // the entity that stays
            qt[q] = side == 0 ? h : t;  // the true answer among the candidates
            qr[q] = r;
            qs[q] = side;
            qslot[q] = cur_slot;
            qo[q] = (int32_t)(2 * (ti - first) + side);
            q++;
         }
      }
      close_tile(q);
      passes.back().q_end = q;
      passes.back().tile_end = tiles.size();
      // pageable source: the copy is staged before the call returns, so the local vector may go out of scope
      KB2E_CUDA(c, cudaMemcpyAsync(s->q_int, qi.data(), qi.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   }
   if ((int64_t)tiles.size() > s->tile_cap) {
      pool_free(c, s->tiles);
      s->tiles = nullptr;
      KB2E_CUDA(c, pool_alloc(c, &s->tiles, tiles.size() * sizeof(int4)));
      s->tile_cap = (int64_t)tiles.size();
   }
   if (!tiles.empty())
      KB2E_CUDA(c, cudaMemcpyAsync(s->tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
   lap("query arrays + H2D");

   RankArgs a;
   memset(&a, 0, sizeof(a));
   a.rel64 = c->rel64;
   a.q_fixed = s->q_int; a.q_truth = s->q_int + nq; a.q_rel = s->q_int + 2 * nq; a.q_side = s->q_int + 3 * nq;
   a.q_slot = s->q_int + 4 * nq;
   a.q_etrue = s->q_etrue; a.q_cnt = s->q_cnt;
   a.seg_key = s->seg_key; a.seg_val = s->seg_val; a.ent_key = s->ent_key; a.nbr = s->nbr; a.seg_mask = s->seg_mask;
   a.n_ent = s->n_ent;
   a.nq = nq; a.nE = c->nE; a.D = c->D; a.ld = s->ld;

   const size_t smem = (size_t)c->D * kQT * sizeof(double2);
   if (smem > 200 * 1024) return fail(c, KB2E_ERR_LIMIT, "embedding size too large for the ranking kernel");
   KB2E_CUDA(c, cudaFuncSetAttribute(rank_exact_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   KB2E_CUDA(c, cudaFuncSetAttribute(rank_exact_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

   if (use_tc) {
      rc = tc_init(c, &s->tc);
      if (rc) return rc;
      if (s->tc.epoch != c->tables_epoch) {   // operand tiles + max |c| are a function of the entity table alone
         rc = tc_prepare_candidates(c, &s->tc);
         if (rc) return rc;
         s->tc.epoch = c->tables_epoch;
      }
   }
   while (s->pass_ev.size() < 2 * passes.size()) {
      cudaEvent_t e;
      KB2E_CUDA(c, cudaEventCreate(&e));
      s->pass_ev.push_back(e);
   }
   if (use_trp) {
      if (s->trp.epoch != c->tables_epoch) {
         rc = trp_prepare_entities(c, &s->trp);
         if (rc) return rc;
         s->trp.epoch = c->tables_epoch;
      }
      while (s->pass_ev.size() < 4 * passes.size()) {   // + two per pass around the tensor-core projection
         cudaEvent_t e;
         KB2E_CUDA(c, cudaEventCreate(&e));
         s->pass_ev.push_back(e);
      }
   }
   lap("tensor-core operand prep");
   bool f32_started = false;
   // Opt-in (KB2E_RANK_OVERLAP=1): measured on B200 at FB15k shape the filter pass beside the tensor-core kernel costs that
   // kernel more issue slots than the overlap saves (1.83 ms with, 1.43 ms without: profiles/r02_rank_probes.txt)
   const char* ov_env = getenv("KB2E_RANK_OVERLAP");
   const bool overlap = passes.size() == 1 && s->n_ent != 0 && ov_env != nullptr && atoi(ov_env) != 0;
   // KB2E_RANK_OVERLAP=2 (tensor-core path): the filter pass is enqueued AFTER the MMA kernel, on a low-priority stream and
   // with a small footprint (two 256-thread CTAs per SM), so that it runs in the threads and registers the one-CTA-per-SM
   // MMA kernel leaves free instead of delaying its CTAs.  Measured: whole call 1.303 ms against 1.348 ms without overlap
   // (the MMA kernel itself stretches from 0.89 to 1.08 ms beside the filter warps) -- a 3 % gain that blurs the per-kernel
   // accounting of the bench, so it stays opt-in too.
   const bool overlap_late = overlap && atoi(ov_env) == 2;
   if (overlap && !s->side) {
      int lo = 0, hi = 0;
      KB2E_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
      KB2E_CUDA(c, cudaStreamCreateWithPriority(&s->side, cudaStreamNonBlocking, lo));
      KB2E_CUDA(c, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
      KB2E_CUDA(c, cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
   }
   // filter pass of the current pass window (a.q_begin / a.q_end / a.ct as set below), enqueued on `st`
   auto run_filter = [&](cudaStream_t st) -> int {
      KB2E_CUDA(c, cudaMemsetAsync(s->chunk_count, 0, sizeof(unsigned int), st));
      filter_plan_kernel<<<nblk(a.q_end - a.q_begin, 256), 256, 0, st>>>(a, s->seg_end, s->chunks, s->chunk_count, s->chunk_cap);
      const unsigned fb = (overlap_late && st != c->stream ? 2 : 8) * c->num_sms;   // persistent: 8 (2) x 256 threads per SM striding over the pair list
      if (use_trp) {
         int rc2 = trp_filter(c, &s->trp, l2, s->q_int, nq, s->q_etrue, s->chunks, s->chunk_count, s->chunk_cap, s->q_cnt, st);
         if (rc2) return rc2;
      } else if (per_rel) {
         if (l2) filter_pairs_kernel<1, false><<<fb, 256, 0, st>>>(a, c->ent64, s->chunks, s->chunk_count, s->chunk_cap);
         else filter_pairs_kernel<0, false><<<fb, 256, 0, st>>>(a, c->ent64, s->chunks, s->chunk_count, s->chunk_cap);
      } else {
         // tensor-core path, filter enqueued behind tc_run of this window: most pairs are decided from its operand tiles
         if (use_tc && st == c->stream && !getenv("KB2E_RANK_NO_FILTER_PREFILTER")) {
            int rc2 = tc_filter_prefilter(c, &s->tc, a.q_begin, s->chunks, s->chunk_count, s->chunk_cap, s->q_cnt + 2 * nq, st);
            if (rc2) return rc2;
         }
         if (l2) filter_pairs_kernel<1, true><<<fb, 256, 0, st>>>(a, c->ent64, s->chunks, s->chunk_count, s->chunk_cap);
         else filter_pairs_kernel<0, true><<<fb, 256, 0, st>>>(a, c->ent64, s->chunks, s->chunk_count, s->chunk_cap);
      }
      KB2E_CUDA(c, cudaGetLastError());
      return KB2E_OK;
   };
   // fork: everything enqueued on the main stream so far (E_true of the pass) happens before the filter pass starts
   auto fork_mark = [&]() -> int {
      KB2E_CUDA(c, cudaEventRecord(s->ev_fork, c->stream));
      return KB2E_OK;
   };
   auto fork_launch = [&]() -> int {
      KB2E_CUDA(c, cudaStreamWaitEvent(s->side, s->ev_fork, 0));
      int rc2 = run_filter(s->side);
      if (rc2) return rc2;
      KB2E_CUDA(c, cudaEventRecord(s->ev_join, s->side));
      return KB2E_OK;
   };
   auto fork_filter = [&]() -> int {
      int rc2 = fork_mark();
      return rc2 ? rc2 : fork_launch();
   };
   // Everything below is enqueued without a host wait; the one synchronisation is at the end of the call.
   KB2E_CUDA(c, cudaEventRecord(c->ev0, c->stream));
   for (size_t p = 0; p < passes.size(); p++) {
      const Pass& ps = passes[p];
      const long long pq = ps.q_end - ps.q_begin;
      if (pq == 0) continue;
      if (per_rel && !use_trp) {
         rc = project(c, ps.rels);
         if (rc) return rc;
         a.ct = s->pt;
      } else {
         a.ct = s->ct0;
      }
      a.tiles = s->tiles + ps.tile_begin;
      a.q_begin = ps.q_begin;
      a.q_end = ps.q_end;
      if (use_trp) {
         const unsigned ntiles = (unsigned)(ps.tile_end - ps.tile_begin);
         rc = f32_ensure(c, &s->f32, ps.rels.size(), s->ld, nq, !f32_started);
         f32_started = true;
         if (rc) return rc;
         KB2E_CUDA(c, cudaEventRecord(s->pass_ev[2 * passes.size() + 2 * p], c->stream));
         rc = trp_project(c, &s->trp, ps.rels, s->ld, s->f32.ct32);                               // tcgen05: P_r = E M_r, all slots
         if (rc) return rc;
         KB2E_CUDA(c, cudaEventRecord(s->pass_ev[2 * passes.size() + 2 * p + 1], c->stream));
         rc = trp_queries(c, &s->trp, l2, s->q_int, nq, ps.q_begin, ps.q_end, s->q_etrue);         // exact V[q], E_true[q]
         if (rc) return rc;
         if (overlap && (rc = fork_filter())) return rc;
         rc = trp_thresholds(c, &s->trp, &s->f32, l2, s->q_int, nq, ps.q_begin, ps.q_end, s->q_etrue);
         if (rc) return rc;
         rc = f32_main(c, &s->f32, l2, s->ld, a.tiles, ntiles, s->q_cnt, s->pass_ev[2 * p], s->pass_ev[2 * p + 1]);
         if (rc) return rc;
         rc = trp_recheck(c, &s->trp, &s->f32, l2, s->q_int, nq, s->q_etrue, s->q_cnt);
         if (rc) return rc;
         c->rstats.launches += 6;
      } else {
         if (l2) etrue_kernel<1><<<nblk(pq, 128), 128, 0, c->stream>>>(a);
         else etrue_kernel<0><<<nblk(pq, 128), 128, 0, c->stream>>>(a);
         if (overlap && use_tc && overlap_late) { if ((rc = fork_mark())) return rc; }
         else if (overlap && (rc = fork_filter())) return rc;
      }
      if (use_trp) {
         // scored above
      } else if (use_tc) {
         // tensor-core pre-filter + exact recheck of the undecided band (rank_tc.cu)
         rc = tc_run(c, &s->tc, a.q_fixed + ps.q_begin, a.q_rel + ps.q_begin, a.q_side + ps.q_begin, a.q_etrue + ps.q_begin,
                     pq, s->q_cnt + ps.q_begin);
         if (rc) return rc;
         if (overlap && overlap_late && (rc = fork_launch())) return rc;   // behind the MMA kernel in issue order
         // band entries hold query indices relative to the pass window
         recheck_kernel<1><<<4 * c->num_sms, 128, 0, c->stream>>>(
            s->tc.band, s->tc.scalars + 1, s->tc.band_cap, c->ent64, c->rel64, a.q_fixed + ps.q_begin, a.q_truth + ps.q_begin,
            a.q_rel + ps.q_begin, a.q_side + ps.q_begin, a.q_etrue + ps.q_begin, s->q_cnt + ps.q_begin, nq, c->D);
         c->rstats.launches += 4;
      } else if (use_f32) {
         const unsigned ntiles = (unsigned)(ps.tile_end - ps.tile_begin);
         rc = f32_prepare(c, &s->f32, a.ct, per_rel ? ps.rels.size() : 1, s->ld, nq, !f32_started);
         f32_started = true;
         if (rc) return rc;
         rc = f32_run(c, &s->f32, l2, a.ct, s->ld, s->q_int, nq, s->q_etrue, ps.q_begin, ps.q_end, a.tiles, ntiles, s->q_cnt,
                      s->pass_ev[2 * p], s->pass_ev[2 * p + 1]);
         if (rc) return rc;
         c->rstats.launches += 4;
      } else {
         const unsigned ntiles = (unsigned)(ps.tile_end - ps.tile_begin);
         // enough CTAs to fill the machine twice even when a pass has few query tiles; per-thread
         // counters are 16-bit, so a split never covers more than 32768 candidate steps
         const int steps = (c->nE + kRankThreads - 1) / kRankThreads;
         long long splits = std::max<long long>(1, (2ll * c->num_sms + ntiles - 1) / ntiles);
         splits = std::max<long long>(splits, (steps + 32767) / 32768);
         splits = std::min<long long>(splits, steps);
         a.splits = (int)splits;
         KB2E_CUDA(c, cudaEventRecord(s->pass_ev[2 * p], c->stream));
         if (l2) rank_exact_kernel<1><<<dim3(ntiles, (unsigned)splits), kRankThreads, smem, c->stream>>>(a);
         else rank_exact_kernel<0><<<dim3(ntiles, (unsigned)splits), kRankThreads, smem, c->stream>>>(a);
         KB2E_CUDA(c, cudaEventRecord(s->pass_ev[2 * p + 1], c->stream));
      }
      if (s->n_ent && !overlap) {
         rc = run_filter(c->stream);
         if (rc) return rc;
      }
      KB2E_CUDA(c, cudaGetLastError());
      c->rstats.launches += per_rel ? 5 : 4;
   }
   if (overlap) KB2E_CUDA(c, cudaStreamWaitEvent(c->stream, s->ev_join, 0));   // join: the filter counts are complete
   finalize_kernel<<<nblk(nq, 256), 256, 0, c->stream>>>(s->q_cnt, s->q_int + 5 * nq, nq, s->out, s->sums);
   KB2E_CUDA(c, cudaGetLastError());
   KB2E_CUDA(c, cudaEventRecord(c->ev1, c->stream));
   // results go straight into the caller's buffers (a true DMA when they are pinned)
   unsigned long long hs[4];
   int32_t* dst[4] = {raw_rank, filt_rank, raw_ties, filt_ties};
   for (int k = 0; k < 4; k++)
      if (dst[k]) KB2E_CUDA(c, cudaMemcpyAsync(dst[k], s->out + (size_t)k * nq, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(hs, s->sums, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   lap("kernels + D2H");
   if (use_tc) {
      bool overflow = false;
      rc = tc_collect(c, &s->tc, &overflow);
      if (rc) return rc;
      if (overflow) {
         // the band list was too small (degenerate tables): redo the call with the exact kernel only
         const uint32_t saved = c->cfg.flags;
         c->cfg.flags |= KB2E_FLAG_RANK_EXACT_ONLY;
         rc = rank_run(c, first, count, raw_rank, filt_rank, raw_ties, filt_ties, sums);
         c->cfg.flags = saved;
         return rc;
      }
      c->rstats.main_kernel_ms += s->tc.last_ms;
      c->rstats.rechecked += s->tc.last_band;
   } else {
      if (use_f32) {
         if (s->f32.host_count[1]) {
            // a band list overflowed (degenerate tables): redo the call with the exact kernel only
            const uint32_t saved = c->cfg.flags;
            c->cfg.flags |= KB2E_FLAG_RANK_EXACT_ONLY;
            rc = rank_run(c, first, count, raw_rank, filt_rank, raw_ties, filt_ties, sums);
            c->cfg.flags = saved;
            return rc;
         }
         c->rstats.rechecked += s->f32.host_count[2];
      }
      for (size_t p = 0; p < passes.size(); p++) {
         if (passes[p].q_end == passes[p].q_begin) continue;
         float pms = 0.f;
         KB2E_CUDA(c, cudaEventElapsedTime(&pms, s->pass_ev[2 * p], s->pass_ev[2 * p + 1]));
         c->rstats.main_kernel_ms += pms;
         if (use_trp) {
            // operand preparation + tensor-core projection of every pass; the projection kernel alone for the last pass
            KB2E_CUDA(c, cudaEventElapsedTime(&pms, s->pass_ev[2 * passes.size() + 2 * p], s->pass_ev[2 * passes.size() + 2 * p + 1]));
            c->rstats.project_ms += pms;
            if (p + 1 == passes.size()) {
               KB2E_CUDA(c, cudaEventElapsedTime(&pms, s->trp.e0, s->trp.e1));
               c->rstats.project_kernel_ms += pms;
            }
         }
      }
   }
   float ms = 0.f;
   KB2E_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
   c->rstats.kernel_ms += ms;
   c->rstats.queries += (uint64_t)nq;
   c->rstats.launches += 2;
   if (sums) for (int k = 0; k < 4; k++) sums[k] = (int64_t)hs[k];
   return KB2E_OK;
}

int rank_score64(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n, double* out) {
   int rc = prepare_tables(c);
   if (rc) return rc;
   RankState* s = c->rank;
   const bool l2 = c->cfg.model != KB2E_MODEL_TRANSH && c->cfg.distance == KB2E_DISTANCE_L2;
   const double* ct = s->ct0;
   int32_t* slot_of_rel = nullptr;
   if (c->cfg.model != KB2E_MODEL_TRANSE) {
      // project every relation (test hook: small tables only)
      std::vector<int32_t> rels(c->nR);
      std::iota(rels.begin(), rels.end(), 0);
      rc = project(c, rels);
      if (rc) return rc;
      ct = s->pt;
      slot_of_rel = s->slot_rel;  // identity map
   }
   if (l2) score64_kernel<1><<<nblk(n, 128), 128, 0, c->stream>>>(ct, c->rel64, slot_of_rel, c->D, s->ld, h, t, r, n, out);
   else score64_kernel<0><<<nblk(n, 128), 128, 0, c->stream>>>(ct, c->rel64, slot_of_rel, c->D, s->ld, h, t, r, n, out);
   KB2E_CUDA(c, cudaGetLastError());
   return KB2E_OK;
}

int rank_debug_transr_projection(kb2e_ctx* c, int relation, float* out, double* eps_rel) {
   if (c->cfg.model != KB2E_MODEL_TRANSR || c->D > 128) return fail(c, KB2E_ERR_ARG, "TransR with an embedding size up to 128 only");
   if (relation < 0 || relation >= c->nR) return fail(c, KB2E_ERR_ARG, "relation out of range");
   int rc = ensure_state(c);
   if (rc) return rc;
   rc = ensure64(c);
   if (rc) return rc;
   return trp_debug_project(c, &c->rank->trp, relation, out, eps_rel);
}

void rank_free(kb2e_ctx* c) {
   RankState* s = c->rank;
   if (!s) return;
   pool_free(c, s->seg_key); pool_free(c, s->seg_val); pool_free(c, s->seg_end); pool_free(c, s->nbr); pool_free(c, s->ent_key);
   pool_free(c, s->chunks); pool_free(c, s->chunk_count); pool_free(c, s->key_tmp); pool_free(c, s->val_tmp); pool_free(c, s->cub_tmp);
   pool_free(c, s->ct0); pool_free(c, s->pt);
   pool_free(c, s->q_int); pool_free(c, s->q_etrue); pool_free(c, s->q_cnt); pool_free(c, s->tiles); pool_free(c, s->slot_rel);
   pool_free(c, s->sums); pool_free(c, s->out);
   for (cudaEvent_t e : s->pass_ev) cudaEventDestroy(e);
   if (s->side) { cudaStreamDestroy(s->side); cudaEventDestroy(s->ev_fork); cudaEventDestroy(s->ev_join); }
   pool_free(c, s->ids);
   tc_free(c, &s->tc);
   f32_free(c, &s->f32);
   trp_free(c, &s->trp);
   delete s;
   c->rank = nullptr;
}

}  // namespace kb2e
