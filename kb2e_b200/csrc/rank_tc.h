// Interface between rank.cu (exact fp64 ranking, filter, finalize) and rank_tc.cu (tcgen05 pre-filter).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct kb2e_ctx;

namespace kb2e {

namespace tc {
constexpr int kRowChunks = 14;  // operand rows hold 14 x 16 B = 112 bf16: embedding sizes up to 106 + 6 bias columns (norm, threshold); larger: fp32 pre-filter
}

struct TcArgs {
   const __nv_bfloat16* u_hi;
   const __nv_bfloat16* u_lo;
   const __nv_bfloat16* c_hi;
   const __nv_bfloat16* c_lo;
   const float* thr_lo;
   const float* thr_hi;
   int32_t* q_less;
   int2* band;
   unsigned int* band_count;
   unsigned int band_cap;
   long long nq;
   int n_pad;
   int experiment;   // KB2E_TC_EXPERIMENT (timing only, WRONG ranks): 1 = TMEM loads without the compare-and-count, 2 = one TMEM load in four
};

struct TcState {
   void* c_hi = nullptr;
   void* c_lo = nullptr;
   int n_pad = 0;
   void* u_hi = nullptr;
   void* u_lo = nullptr;
   float* thr_lo = nullptr;
   float* thr_hi = nullptr;
   long long q_cap = 0;
   int2* band = nullptr;
   unsigned int band_cap = 0;
   unsigned int* scalars = nullptr;  // [0] max |c| (float bits), [1] band count
   cudaEvent_t e0 = nullptr, e1 = nullptr;
   unsigned int* host_count = nullptr;   // pinned: band count of the last tc_run, valid after tc_collect
   float last_ms = 0.f;
   unsigned int last_band = 0;
   uint64_t epoch = 0;   // kb2e_ctx::tables_epoch the candidate operand tiles were made for
};

bool tc_supported(const kb2e_ctx* c);
int tc_init(kb2e_ctx* c, TcState* s);
int tc_prepare_candidates(kb2e_ctx* c, TcState* s);
// tc_run only enqueues work (no host synchronisation); after the stream has been synchronised, tc_collect reads the
// kernel time and the size of the undecided band (overflow: the band list was too small -> rerun the pass exactly).
int tc_run(kb2e_ctx* c, TcState* s, const int32_t* q_fixed, const int32_t* q_rel, const int32_t* q_side, const double* q_etrue,
           long long nq, int32_t* q_less);
int tc_collect(kb2e_ctx* c, TcState* s, bool* overflow);
// After tc_run of the same query window: decide the filter pass's (query, neighbour) pairs from the operand tiles and
// thresholds of that run where the band allows it (counted into q_filt_less, struck from the list); the rest stays for the
// exact kernel.  q_base = first query of the window (pairs carry global query indices).
int tc_filter_prefilter(kb2e_ctx* c, TcState* s, long long q_base, int2* pairs, const unsigned int* pair_count, unsigned int pair_cap,
                        int32_t* q_filt_less, cudaStream_t stream);
void tc_free(kb2e_ctx* c, TcState* s);

}  // namespace kb2e
