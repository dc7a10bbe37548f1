// Internal context shared by the translation units of libkb2e_b200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/kb2e_b200.h"

struct DistState;
namespace kb2e { struct TrainArgs; }

struct kb2e_ctx {
   kb2e_config cfg;
   int D = 0;       // -size
   int P = 0;       // row pitch in elements: D rounded up to a multiple of 4 (16-byte vector access)
   int nE = 0, nR = 0;
   int device = 0;
   int num_sms = 0;
   cudaStream_t stream = nullptr;
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   std::string err;

   // ---- fp32 training tables -------------------------------------------------------------------
   // tab  : [nE + nR][P]  entity rows, then relation rows (one row-index space for the touched lists)
   // dtab : same shape, the batch's accumulated update ("next - cur" of the reference's *_next_ copies)
   // w    : TransH [nR][P] hyperplane normals; TransR [nR][D][P] = M[r][j=in][i=out]; dw same shape
   float* tab = nullptr;    // (with several stacked models, kb2e_set_replicas: the SELECTED model's rows; *_all = the allocation)
   float* dtab = nullptr;
   float* tab_all = nullptr;
   float* dtab_all = nullptr;
   uint32_t* flag_all = nullptr;
   int K = 1, sel = 0;                       // stacked models and the one upload / download / score / rank address
   std::vector<double> rep_rate, rep_margin;
   std::vector<uint64_t> rep_seed;
   void* rep_dev = nullptr;                  // kb2e::RepParams[K]
   float* w = nullptr;
   float* dw = nullptr;
   size_t w_row = 0;           // elements per relation in w
   uint32_t* flag = nullptr;   // [nE + nR] stamp (global batch + 1) of the last batch that touched the row
   float* relbuf1 = nullptr;         // TransH, few relations: second buffer of the relation-side deltas, [2][nR][P] (train_transh_sr_kernel)
   uint32_t* transr_aux = nullptr;   // TransR training: claim stamps, sample-touched relation stamps, the phase 2b row list (train_transr.cu)
   uint32_t* cflag = nullptr;  // [nR] TransH: batch a relation row was marked for by an entity-side constraint step (list kernels)
   bool thr_valid = false;     // triples[i].w holds the corruption threshold of the current pr table
   int* rmin = nullptr;        // [nE] lowest / highest relation id that touched the entity row (TransH/R)
   int* rmax = nullptr;
   bool v32[3] = {false, false, false};  // per table (entity, relation, weights): fp32 copy is current

   // ---- training set ---------------------------------------------------------------------------
   int4* triples = nullptr;    // (h, t, r, 0)
   int32_t* stage = nullptr;   // H2D staging of the three id columns
   int64_t triples_cap = 0;
   uint64_t hash_cap = 0;
   int64_t n_train = 0;
   uint64_t* hash = nullptr;   // open-addressing set of packed (h, r, t)
   uint64_t hash_mask = 0;
   double* pr = nullptr;       // [nR] the reference's `pr` in thousandths (common/trainer.cpp:82-86)
   bool have_pr = false;

   // persistent-launch scratch
   uint32_t* barrier = nullptr;
   double* loss_dev = nullptr;
   int loss_cap = 0;
   unsigned long long* counters = nullptr;  // [0] active [1] touched_ent [2] touched_rel
   int32_t* pairs_dev = nullptr;
   int64_t pairs_cap = 0;
   int hook_batches = 0;      // batches run through kb2e_train_batch_pairs (each gets its own sampler / global-batch index)
   uint32_t stamp_base = 0;   // batches launched so far: row stamps of a launch are stamp_base + 1 ... (never reused)
   kb2e_train_stats tstats{};

   // ---- fp64 tables for ranking (exact copies of what was uploaded, or widened fp32 state) -----
   double* ent64 = nullptr;  // [nE][D]
   double* rel64 = nullptr;  // [nR][D]
   double* w64 = nullptr;    // TransH [nR][D]; TransR [nR][D][D]
   bool v64[3] = {false, false, false};  // per table: fp64 copy is current

   // ---- evaluation set -------------------------------------------------------------------------
   std::vector<int32_t> test_h, test_t, test_r;   // host copy: the per-relation models order their queries on the host
   int32_t* filt_dev = nullptr;   // known-true triples besides the test set, device-resident: h | t | r columns of filt_cap each
   size_t filt_n = 0, filt_cap = 0;
   bool filter_dirty = true;
   uint64_t tables_epoch = 1;     // bumped whenever a table changes (upload, init, training): keys the ranking's derived operands
   struct RankState* rank = nullptr;
   DistState* dist = nullptr;  // entity-partitioned multi-GPU training (train_dist.cu)
   uint32_t* pend = nullptr;           // [3][nE + nR] per-row reference counters of the one-barrier kernel (train_fused.cu)
   kb2e_rank_stats rstats{};
};

namespace kb2e {

int fail(kb2e_ctx* ctx, int code, const std::string& msg);
int cuda_fail(kb2e_ctx* ctx, cudaError_t e, const char* what);

#define KB2E_CUDA(ctx, call)                                         \
   do {                                                              \
      cudaError_t e__ = (call);                                      \
      if (e__ != cudaSuccess) return kb2e::cuda_fail(ctx, e__, #call); \
   } while (0)

// Device allocation helpers.  A stream-ordered pool (cudaMallocAsync / cudaFreeAsync with the release threshold
// lifted) was tried here and measured SLOWER on this pool's B200 boxes for the create -> rank -> destroy cycle
// (context create 25-90 ms, destroy up to 400 ms, vs 3 ms / 5-55 ms with plain cudaMalloc / cudaFree), so the
// helpers stay thin wrappers; long-lived contexts (the intended use) allocate once and regrow rarely.
template <typename T>
inline cudaError_t pool_alloc(kb2e_ctx*, T** p, size_t bytes) {
   return cudaMalloc(reinterpret_cast<void**>(p), bytes ? bytes : 1);
}
inline void pool_free(kb2e_ctx*, void* p) {
   if (p) cudaFree(p);
}

// train.cu
int train_alloc(kb2e_ctx* ctx);
void train_free(kb2e_ctx* ctx);
int train_set_triples(kb2e_ctx* ctx, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n);
int train_init_embeddings(kb2e_ctx* ctx);
int train_run(kb2e_ctx* ctx, int first_epoch, int n_epochs, const int32_t* pairs_dev, int64_t n_pairs, double* loss_out,
              bool phase1_only = false);
int train_take_deltas(kb2e_ctx* ctx, double* d_ent, double* d_rel, double* d_w);
int train_set_replicas(kb2e_ctx* ctx, int K, const double* rates, const double* margins, const uint64_t* seeds);
int train_select_replica(kb2e_ctx* ctx, int m);
int train_sample(kb2e_ctx* ctx, int epoch, int batch, int64_t count, int32_t* pairs_dev);
int train_score32(kb2e_ctx* ctx, const int32_t* h_dev, const int32_t* t_dev, const int32_t* r_dev, int64_t n, double* out_dev);
int num_tables(const kb2e_ctx* ctx);
int widen_table(kb2e_ctx* ctx, int table);    // fp32 -> fp64 copy of one table
int narrow_table(kb2e_ctx* ctx, int table);   // fp64 -> fp32
int ensure32(kb2e_ctx* ctx);                  // every table of the model current in fp32 (training)
int ensure64(kb2e_ctx* ctx);                  // ... in fp64 (ranking)
double* table64(kb2e_ctx* ctx, int table);

// train_fused.cu
bool train_fused_wanted(const kb2e_ctx* ctx, long long batchsize, int lps, int threads);
int train_fused_launch(kb2e_ctx* ctx, const TrainArgs& base, int lps, int nv, int threads);

// train_transr.cu
int train_transr_launch(kb2e_ctx* ctx, const TrainArgs& base, int* threads_out);

// rank.cu
int rank_run(kb2e_ctx* ctx, int64_t first, int64_t count, int32_t* raw_rank, int32_t* filt_rank,
             int32_t* raw_ties, int32_t* filt_ties, int64_t sums[4]);
int rank_score64(kb2e_ctx* ctx, const int32_t* h_dev, const int32_t* t_dev, const int32_t* r_dev, int64_t n, double* out_dev);
void rank_free(kb2e_ctx* ctx);
int rank_debug_transr_projection(kb2e_ctx* ctx, int relation, float* out, double* eps_rel);

}  // namespace kb2e
