// Interface between rank.cu and rank_transr.cu: TransR ranking with the batched M_r^T E projection on the tensor cores.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "rank_f32.h"

struct kb2e_ctx;

namespace kb2e {

struct TrpState {
   void* e_hi = nullptr;          // entity operand tiles, bf16 hi / lo parts: [n_pad / 128][kc][128][8]
   void* e_lo = nullptr;
   void* m_hi = nullptr;          // per slot: M_r^T operand tile [kc][ncols][8] (row = output dimension, K = input dimension)
   void* m_lo = nullptr;
   double* bounds = nullptr;      // [slot][4]: bound on max |C|_1, on max |C|_2^2, eta_1 (L1 error of a projected row), eta_2 (L2 error)
   double* V = nullptr;           // [query][D] exact fp64 projection of the query's fixed entity
   double* scalars = nullptr;     // [0] max |e|_2 over the entity table
   int32_t* slot_rel = nullptr;
   unsigned int* band_sorted = nullptr;
   int n_pad = 0, kc = 0, ncols = 0;
   size_t slot_cap = 0;
   long long q_cap = 0;
   size_t e_cap = 0;
   cudaEvent_t e0 = nullptr, e1 = nullptr;   // around the tensor-core projection of the last pass
   uint64_t epoch = 0;                       // kb2e_ctx::tables_epoch the entity operand tiles were made for
};

// TransR, embedding sizes the tensor-core tiles cover, pre-filter not disabled
bool trp_supported(const kb2e_ctx* c);
// Once per kb2e_rank call: entity operand tiles + max |e|_2 from the fp64 entity table.
int trp_prepare_entities(kb2e_ctx* c, TrpState* s);
// Per pass: operand tiles and error bounds of the pass's relations, then the tensor-core projection of every entity under
// every relation of the pass into ct32 [slot][D][ld] (fp32).
int trp_project(kb2e_ctx* c, TrpState* s, const std::vector<int32_t>& rels, int ld, float* ct32);
// Exact fp64 pieces: V[q] = M_r^T e_fixed and E_true[q] for the queries [q_begin, q_end).
int trp_queries(kb2e_ctx* c, TrpState* s, bool l2, const int32_t* q_int, long long nq_total, long long q_begin, long long q_end, double* q_etrue);
// fp32 query vectors + thresholds (fp32 rounding bound + projection error bound) into the F32State, for f32_main.
int trp_thresholds(kb2e_ctx* c, TrpState* s, F32State* f, bool l2, const int32_t* q_int, long long nq_total, long long q_begin, long long q_end,
                   const double* q_etrue);
// Exact re-score of the undecided band of f32_main (candidates projected on demand in fp64, reference operation order).
int trp_recheck(kb2e_ctx* c, TrpState* s, F32State* f, bool l2, const int32_t* q_int, long long nq_total, const double* q_etrue, int32_t* q_cnt);
// Filter pass: (query, known-true neighbour) pairs, neighbours projected on demand in fp64.
int trp_filter(kb2e_ctx* c, TrpState* s, bool l2, const int32_t* q_int, long long nq_total, const double* q_etrue, const int2* pairs,
               const unsigned int* pair_count, unsigned int pair_cap, int32_t* q_cnt, cudaStream_t stream);
// Test hook (kb2e_debug_transr_projection): the tensor-core projection of every entity under one relation, [num_entities][dim],
// and the relative error bound eps_p it is guaranteed to meet: |P~[c][i] - P[c][i]| <= eps_p * sum_j |e_cj| |M_ji|.
int trp_debug_project(kb2e_ctx* c, TrpState* s, int relation, float* out, double* eps_rel);
double trp_eps(int D);
void trp_free(kb2e_ctx* c, TrpState* s);

}  // namespace kb2e
