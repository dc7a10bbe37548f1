// The C ABI of libkb2e_b200.so (include/kb2e_b200.h): context life cycle, host<->device staging,
// argument checking.  The kernels live in train.cu / rank.cu.  There is no CPU fallback anywhere:
// without a usable CUDA device kb2e_create fails with KB2E_ERR_NO_GPU.

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace kb2e {

static std::mutex g_err_mutex;
static std::string g_create_error;

int fail(kb2e_ctx* ctx, int code, const std::string& msg) {
   if (ctx) {
      ctx->err = msg;
   } else {
      std::lock_guard<std::mutex> lock(g_err_mutex);
      g_create_error = msg;
   }
   return code;
}

int cuda_fail(kb2e_ctx* ctx, cudaError_t e, const char* what) {
   return fail(ctx, KB2E_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace kb2e

using namespace kb2e;

extern "C" {

int kb2e_create(const kb2e_config* cfg, kb2e_ctx** out) {
   if (!cfg || !out) return fail(nullptr, KB2E_ERR_ARG, "kb2e_create: null argument");
   *out = nullptr;
   if (cfg->model < 0 || cfg->model > 2) return fail(nullptr, KB2E_ERR_ARG, "kb2e_create: unknown model");
   if (cfg->dim <= 0) return fail(nullptr, KB2E_ERR_ARG, "kb2e_create: embedding size must be positive");
   if (cfg->num_entities <= 0 || cfg->num_relations <= 0)
      return fail(nullptr, KB2E_ERR_ARG, "kb2e_create: need at least one entity and one relation");
   if (cfg->num_entities >= (1ll << kEntityBits) || cfg->num_relations >= (1ll << kRelationBits))
      return fail(nullptr, KB2E_ERR_LIMIT, "kb2e_create: at most 2^24 entities and 2^16 relations");
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
      return fail(nullptr, KB2E_ERR_NO_GPU, std::string("kb2e_create: no CUDA device (") + cudaGetErrorString(e) +
                                                "); this library has no CPU fallback");
   if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, KB2E_ERR_ARG, "kb2e_create: bad device ordinal");
   cudaDeviceProp prop;
   if ((e = cudaSetDevice(cfg->device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess)
      return fail(nullptr, KB2E_ERR_CUDA, std::string("kb2e_create: ") + cudaGetErrorString(e));
   if (prop.major != 10)
      return fail(nullptr, KB2E_ERR_NO_GPU, std::string("kb2e_create: device '") + prop.name +
                                                "' is not sm_100 (this library carries sm_100a code only)");
   kb2e_ctx* c = new kb2e_ctx();
   c->cfg = *cfg;
   c->D = cfg->dim;
   c->P = (cfg->dim + 3) / 4 * 4;
   c->nE = (int)cfg->num_entities;
   c->nR = (int)cfg->num_relations;
   c->device = cfg->device;
   c->num_sms = prop.multiProcessorCount;
   if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
       (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess) {
      if (c->ev1) cudaEventDestroy(c->ev1);
      if (c->ev0) cudaEventDestroy(c->ev0);
      if (c->stream) cudaStreamDestroy(c->stream);
      delete c;
      return fail(nullptr, KB2E_ERR_CUDA, std::string("kb2e_create: ") + cudaGetErrorString(e));
   }
   *out = c;
   return KB2E_OK;
}

void kb2e_destroy(kb2e_ctx* c) {
   if (!c) return;
   cudaSetDevice(c->device);
   cudaStreamSynchronize(c->stream);
   kb2e_dist_teardown(c);
   rank_free(c);
   train_free(c);
   cudaEventDestroy(c->ev0);
   cudaEventDestroy(c->ev1);
   cudaStreamDestroy(c->stream);
   delete c;
}

const char* kb2e_last_error(const kb2e_ctx* c) {
   if (c) return c->err.c_str();
   std::lock_guard<std::mutex> lock(g_err_mutex);
   return g_create_error.c_str();
}

void* kb2e_stream(kb2e_ctx* c) { return c ? (void*)c->stream : nullptr; }

#define KB2E_ENTER(c)                                     \
   if (!(c)) return KB2E_ERR_ARG;                         \
   KB2E_CUDA(c, cudaSetDevice((c)->device))

int kb2e_set_train_triples(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n) {
   KB2E_ENTER(c);
   if (n < 0 || (n > 0 && (!h || !t || !r))) return fail(c, KB2E_ERR_ARG, "kb2e_set_train_triples: bad arguments");
   return train_set_triples(c, h, t, r, n);
}

int kb2e_set_bern(kb2e_ctx* c, const double* head_mean, const double* tail_mean) {
   KB2E_ENTER(c);
   int rc = train_alloc(c);
   if (rc) return rc;
   std::vector<double> pr(c->nR);
   for (int i = 0; i < c->nR; i++) {
      // common/trainer.cpp:82-86: pr = 1000 * T / (T + H), or 500 for unif
      if (c->cfg.method == KB2E_METHOD_UNIF) {
         pr[i] = 500;
      } else {
         if (!head_mean || !tail_mean) return fail(c, KB2E_ERR_ARG, "kb2e_set_bern: bern needs both statistics");
         pr[i] = 1000 * tail_mean[i] / (tail_mean[i] + head_mean[i]);
      }
   }
   KB2E_CUDA(c, cudaMemcpyAsync(c->pr, pr.data(), pr.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   c->have_pr = true;
   c->thr_valid = false;
   return KB2E_OK;
}

int kb2e_init_embeddings(kb2e_ctx* c) {
   KB2E_ENTER(c);
   return train_init_embeddings(c);
}

static int table_shape(kb2e_ctx* c, int table, int64_t& rows, int64_t& cols) {
   cols = c->D;
   if (table == KB2E_TABLE_ENTITY) { rows = c->nE; return KB2E_OK; }
   if (table == KB2E_TABLE_RELATION) { rows = c->nR; return KB2E_OK; }
   if (table == KB2E_TABLE_WEIGHTS && c->cfg.model == KB2E_MODEL_TRANSH) { rows = c->nR; return KB2E_OK; }
   if (table == KB2E_TABLE_WEIGHTS && c->cfg.model == KB2E_MODEL_TRANSR) { rows = (int64_t)c->nR * c->D; return KB2E_OK; }
   return fail(c, KB2E_ERR_ARG, "unknown table for this model");
}

int kb2e_upload(kb2e_ctx* c, int table, const double* host, int64_t rows, int64_t cols) {
   KB2E_ENTER(c);
   if (!host) return fail(c, KB2E_ERR_ARG, "kb2e_upload: null buffer");
   int64_t er, ec;
   int rc = table_shape(c, table, er, ec);
   if (rc) return rc;
   if (rows != er || cols != ec) return fail(c, KB2E_ERR_ARG, "kb2e_upload: shape mismatch");
   rc = train_alloc(c);
   if (rc) return rc;
   double* dev = table64(c, table);
   if (!dev) return fail(c, KB2E_ERR_CUDA, "kb2e_upload: out of device memory");
   // the exact fp64 values stay on the device for ranking; training gets the fp32 rounding of them
   KB2E_CUDA(c, cudaMemcpyAsync(dev, host, (size_t)rows * cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
   c->v64[table] = true;
   c->tables_epoch++;
   rc = narrow_table(c, table);
   if (rc) return rc;
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   return KB2E_OK;
}

int kb2e_download(kb2e_ctx* c, int table, double* host, int64_t rows, int64_t cols) {
   KB2E_ENTER(c);
   if (!host) return fail(c, KB2E_ERR_ARG, "kb2e_download: null buffer");
   int64_t er, ec;
   int rc = table_shape(c, table, er, ec);
   if (rc) return rc;
   if (rows != er || cols != ec) return fail(c, KB2E_ERR_ARG, "kb2e_download: shape mismatch");
   if (!c->v64[table]) {
      rc = widen_table(c, table);
      if (rc) return rc;
   }
   KB2E_CUDA(c, cudaMemcpyAsync(host, table64(c, table), (size_t)rows * cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   return KB2E_OK;
}

int kb2e_train_epochs(kb2e_ctx* c, int32_t first_epoch, int32_t n_epochs, double* loss_per_epoch) {
   KB2E_ENTER(c);
   if (first_epoch < 0 || n_epochs < 0) return fail(c, KB2E_ERR_ARG, "kb2e_train_epochs: negative epoch");
   return train_run(c, first_epoch, n_epochs, nullptr, 0, loss_per_epoch);
}

int kb2e_set_replicas(kb2e_ctx* c, int32_t n_models, const double* rates, const double* margins, const uint64_t* seeds) {
   KB2E_ENTER(c);
   return train_set_replicas(c, n_models, rates, margins, seeds);
}

int kb2e_select_replica(kb2e_ctx* c, int32_t model_index) {
   KB2E_ENTER(c);
   return train_select_replica(c, model_index);
}

int kb2e_get_train_stats(kb2e_ctx* c, kb2e_train_stats* out) {
   if (!c || !out) return KB2E_ERR_ARG;
   *out = c->tstats;
   return KB2E_OK;
}

int kb2e_get_rank_stats(kb2e_ctx* c, kb2e_rank_stats* out) {
   if (!c || !out) return KB2E_ERR_ARG;
   *out = c->rstats;
   return KB2E_OK;
}

static int stage_triples(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n, int32_t** dev) {
   for (int64_t i = 0; i < n; i++) {
      if (h[i] < 0 || h[i] >= c->nE || t[i] < 0 || t[i] >= c->nE || r[i] < 0 || r[i] >= c->nR)
         return fail(c, KB2E_ERR_ARG, "triple " + std::to_string(i) + " has an id out of range");
   }
   KB2E_CUDA(c, pool_alloc(c, dev, 3 * (size_t)n * sizeof(int32_t)));
   KB2E_CUDA(c, cudaMemcpyAsync(*dev, h, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(*dev + n, t, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemcpyAsync(*dev + 2 * n, r, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   return KB2E_OK;
}

int kb2e_score(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n, int32_t precision, double* out) {
   KB2E_ENTER(c);
   if (n < 0 || (n > 0 && (!h || !t || !r || !out))) return fail(c, KB2E_ERR_ARG, "kb2e_score: bad arguments");
   if (n == 0) return KB2E_OK;
   int32_t* dev = nullptr;
   int rc = stage_triples(c, h, t, r, n, &dev);
   if (rc) { pool_free(c, dev); return rc; }
   double* dout = nullptr;
   if (cudaError_t e = pool_alloc(c, &dout, (size_t)n * sizeof(double)); e != cudaSuccess) {
      pool_free(c, dev);
      return cuda_fail(c, e, "kb2e_score: device buffer");
   }
   rc = precision == 0 ? train_score32(c, dev, dev + n, dev + 2 * n, n, dout) : rank_score64(c, dev, dev + n, dev + 2 * n, n, dout);
   if (rc == KB2E_OK) {
      cudaError_t e = cudaMemcpyAsync(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = cuda_fail(c, e, "kb2e_score copy");
   }
   pool_free(c, dev);
   pool_free(c, dout);
   return rc;
}

int kb2e_set_test_triples(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n) {
   KB2E_ENTER(c);
   if (n < 0 || (n > 0 && (!h || !t || !r))) return fail(c, KB2E_ERR_ARG, "kb2e_set_test_triples: bad arguments");
   for (int64_t i = 0; i < n; i++) {
      if (h[i] < 0 || h[i] >= c->nE || t[i] < 0 || t[i] >= c->nE || r[i] < 0 || r[i] >= c->nR)
         return fail(c, KB2E_ERR_ARG, "test triple " + std::to_string(i) + " has an id out of range");
   }
   c->test_h.assign(h, h + n);
   c->test_t.assign(t, t + n);
   c->test_r.assign(r, r + n);
   c->filter_dirty = true;
   return KB2E_OK;
}

// first index in [0, n) whose ids are out of range -> *bad (atomicMin), else *bad stays ~0
__global__ void validate_triples_kernel(const int32_t* h, const int32_t* t, const int32_t* r, long long n, int nE, int nR, unsigned long long* bad) {
   long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n) return;
   if (h[i] < 0 || h[i] >= nE || t[i] < 0 || t[i] >= nE || r[i] < 0 || r[i] >= nR) atomicMin(bad, (unsigned long long)i);
}

int kb2e_add_filter_triples(kb2e_ctx* c, const int32_t* h, const int32_t* t, const int32_t* r, int64_t n) {
   KB2E_ENTER(c);
   if (n < 0) return fail(c, KB2E_ERR_ARG, "kb2e_add_filter_triples: negative count");
   if (n == 0) {
      c->filt_n = 0;
      c->filter_dirty = true;
      return KB2E_OK;
   }
   if (!h || !t || !r) return fail(c, KB2E_ERR_ARG, "kb2e_add_filter_triples: null buffer");
   int rc = train_alloc(c);   // counters
   if (rc) return rc;
   // The columns go straight from the caller's buffers to the device (a DMA when they are pinned) and are checked there:
   // no host loop over the set, no host copy of it.  Capacity only grows (doubling), old entries are kept.
   const size_t need = c->filt_n + (size_t)n;
   if (need > c->filt_cap) {
      const size_t cap = std::max(need, 2 * c->filt_cap);
      int32_t* fresh = nullptr;
      KB2E_CUDA(c, pool_alloc(c, &fresh, 3 * cap * sizeof(int32_t)));
      for (int k = 0; k < 3 && c->filt_n; k++) {
         cudaError_t e = cudaMemcpyAsync(fresh + k * cap, c->filt_dev + k * c->filt_cap, c->filt_n * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream);
         if (e != cudaSuccess) { pool_free(c, fresh); return cuda_fail(c, e, "kb2e_add_filter_triples: regrow"); }
      }
      KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
      pool_free(c, c->filt_dev);
      c->filt_dev = fresh;
      c->filt_cap = cap;
   }
   const int32_t* src[3] = {h, t, r};
   for (int k = 0; k < 3; k++)
      KB2E_CUDA(c, cudaMemcpyAsync(c->filt_dev + k * c->filt_cap + c->filt_n, src[k], (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   KB2E_CUDA(c, cudaMemsetAsync(c->counters + 7, 0xff, sizeof(unsigned long long), c->stream));
   validate_triples_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->filt_dev + c->filt_n, c->filt_dev + c->filt_cap + c->filt_n,
                                                                              c->filt_dev + 2 * c->filt_cap + c->filt_n, n, c->nE, c->nR, c->counters + 7);
   KB2E_CUDA(c, cudaGetLastError());
   unsigned long long bad = 0;
   KB2E_CUDA(c, cudaMemcpyAsync(&bad, c->counters + 7, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
   KB2E_CUDA(c, cudaStreamSynchronize(c->stream));
   if (bad != ~0ull) return fail(c, KB2E_ERR_ARG, "filter triple " + std::to_string(bad) + " has an id out of range");
   c->filt_n += (size_t)n;
   c->filter_dirty = true;
   return KB2E_OK;
}

int kb2e_rank(kb2e_ctx* c, int64_t first, int64_t count, int32_t* raw_rank, int32_t* filt_rank,
              int32_t* raw_ties, int32_t* filt_ties, int64_t sums[4]) {
   KB2E_ENTER(c);
   return rank_run(c, first, count, raw_rank, filt_rank, raw_ties, filt_ties, sums);
}

int kb2e_debug_transr_projection(kb2e_ctx* c, int32_t relation, float* proj, double* eps_rel) {
   KB2E_ENTER(c);
   if (!proj) return fail(c, KB2E_ERR_ARG, "kb2e_debug_transr_projection: null buffer");
   return rank_debug_transr_projection(c, relation, proj, eps_rel);
}

int kb2e_sample_batch(kb2e_ctx* c, int32_t epoch, int32_t batch, int64_t count, int32_t* pairs_out) {
   KB2E_ENTER(c);
   if (count <= 0 || !pairs_out) return fail(c, KB2E_ERR_ARG, "kb2e_sample_batch: bad arguments");
   int32_t* dev = nullptr;
   KB2E_CUDA(c, pool_alloc(c, &dev, 6 * (size_t)count * sizeof(int32_t)));
   int rc = train_sample(c, epoch, batch, count, dev);
   if (rc == KB2E_OK) {
      cudaError_t e = cudaMemcpyAsync(pairs_out, dev, 6 * (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = cuda_fail(c, e, "kb2e_sample_batch copy");
   }
   pool_free(c, dev);
   return rc;
}

static int stage_pairs(kb2e_ctx* c, const char* who, const int32_t* pairs, int64_t n) {
   if (n <= 0 || !pairs) return fail(c, KB2E_ERR_ARG, std::string(who) + ": bad arguments");
   for (int64_t i = 0; i < n; i++) {
      const int32_t* p = pairs + 6 * i;
      if (p[0] < 0 || p[0] >= c->nE || p[1] < 0 || p[1] >= c->nE || p[3] < 0 || p[3] >= c->nE || p[4] < 0 || p[4] >= c->nE ||
          p[2] < 0 || p[2] >= c->nR || p[5] != p[2] || (p[3] != p[0] && p[4] != p[1]))
         return fail(c, KB2E_ERR_ARG, "pair " + std::to_string(i) + " is not (h,t,r) with one side corrupted");
   }
   int rc = train_alloc(c);
   if (rc) return rc;
   if (n > c->pairs_cap) {
      pool_free(c, c->pairs_dev);
      c->pairs_dev = nullptr;
      c->pairs_cap = 0;
      KB2E_CUDA(c, pool_alloc(c, &c->pairs_dev, 6 * (size_t)n * sizeof(int32_t)));
      c->pairs_cap = n;
   }
   KB2E_CUDA(c, cudaMemcpyAsync(c->pairs_dev, pairs, 6 * (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
   return KB2E_OK;
}

int kb2e_train_batch_pairs(kb2e_ctx* c, const int32_t* pairs, int64_t n, double* loss, int64_t* n_active) {
   KB2E_ENTER(c);
   int rc = stage_pairs(c, "kb2e_train_batch_pairs", pairs, n);
   if (rc) return rc;
   const uint64_t before = c->tstats.active;
   double l = 0.0;
   rc = train_run(c, c->hook_batches, 1, c->pairs_dev, n, &l);
   if (rc) return rc;
   c->hook_batches++;
   if (loss) *loss = l;
   if (n_active) *n_active = (int64_t)(c->tstats.active - before);
   return KB2E_OK;
}

int kb2e_train_batch_deltas(kb2e_ctx* c, const int32_t* pairs, int64_t n, double* d_entity, double* d_relation, double* d_weights,
                            double* loss, int64_t* n_active) {
   KB2E_ENTER(c);
   int rc = stage_pairs(c, "kb2e_train_batch_deltas", pairs, n);
   if (rc) return rc;
   const uint64_t before = c->tstats.active;
   double l = 0.0;
   rc = train_run(c, c->hook_batches, 1, c->pairs_dev, n, &l, /*phase1_only=*/true);
   if (rc) return rc;
   c->hook_batches++;
   rc = train_take_deltas(c, d_entity, d_relation, d_weights);
   if (rc) return rc;
   if (loss) *loss = l;
   if (n_active) *n_active = (int64_t)(c->tstats.active - before);
   return KB2E_OK;
}

}  // extern "C"
