// Interface between rank.cu and rank_f32.cu (fp32 CUDA-core pre-filter with a rigorous error bound + exact recheck).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct kb2e_ctx;

namespace kb2e {

struct F32Args {
   const float* ct32;     // [slot][D][ld] candidates, fp32
   const float* wq;       // [query][D]  fl32(V - d')
   const float* thr_lo;   // per query: below -> certainly ranked before the truth
   const float* thr_hi;   // per query: above -> certainly not
   const int4* tiles;     // (first query, #queries <= 32, slot, -)
   int32_t* q_less;
   int2* band;            // undecided (query, candidate) pairs
   unsigned int* band_count;   // [0] entries of this pass, [1] sticky overflow flag, [2] total over the call
   unsigned int band_cap;
   int nE, D, ld, splits;
};

struct F32State {
   float* ct32 = nullptr;
   size_t ct32_cap = 0;
   unsigned long long* slot_max = nullptr;   // [slot][2]: max |C|_1, max |C|_2^2 over the slot's candidates (fp64 bits)
   size_t slot_cap = 0;
   float* wq = nullptr;
   float* thr_lo = nullptr;
   float* thr_hi = nullptr;
   int2* band = nullptr;
   unsigned int band_cap = 0;
   unsigned int* band_count = nullptr;
   unsigned int* host_count = nullptr;       // pinned copy of band_count[0..2]
   long long q_cap = 0;
};

constexpr int kF32QueriesPerTile = 32;

// Grow-only buffers for `slots` fp32 candidate matrices and nq queries (first_pass also clears the call-wide counters).
int f32_ensure(kb2e_ctx* c, F32State* s, size_t slots, int ld, long long nq, bool first_pass);
// The all-candidates kernel alone on the prepared s->ct32 / s->wq / s->thr_lo / s->thr_hi (rank_transr.cu fills them itself).
int f32_main(kb2e_ctx* c, F32State* s, bool l2, int ld, const int4* tiles, unsigned ntiles, int32_t* q_cnt, cudaEvent_t e0, cudaEvent_t e1);
// Convert the pass's candidate matrices (fp64, transposed, `slots` of them) and size the per-call buffers.
int f32_prepare(kb2e_ctx* c, F32State* s, const double* ct, size_t slots, int ld, long long nq, bool first_pass);
// Enqueue thresholds + pre-filter + exact recheck for the queries [q_begin, q_end) of one pass (no host synchronisation).
int f32_run(kb2e_ctx* c, F32State* s, bool l2, const double* ct, int ld, const int32_t* q_int, long long nq_total, const double* q_etrue,
            long long q_begin, long long q_end, const int4* tiles, unsigned ntiles, int32_t* q_cnt, cudaEvent_t e0, cudaEvent_t e1);
void f32_free(kb2e_ctx* c, F32State* s);

}  // namespace kb2e
