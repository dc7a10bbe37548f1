/* kb2e_b200 -- C ABI of the B200-native KB2E hot path (TransE / TransH / TransR margin-ranking SGD
 * epochs and filtered link-prediction ranking).  Plain pointers and sizes only; every function
 * returns 0 on success or a KB2E_ERR_* code and never exits the process -- the host mains turn
 * errors into the reference's messages / exit codes.
 *
 * Each entry point names the reference interface it replaces (citations into eriq-augustine/KB2E):
 *   kb2e_create / kb2e_destroy      common::Trainer::Trainer (common/trainer.cpp:15-24) and
 *                                   common::EmbeddingEvaluation ctor (common/evaluation.cpp:23-33)
 *   kb2e_set_train_triples          Trainer::add / heads_,tails_,relations_,triples_ (common/trainer.cpp:26-32)
 *   kb2e_set_bern                   relation{Head,Tail}MeanCooccurrence_ (common/trainer.cpp:171-194)
 *   kb2e_init_embeddings            Trainer::prepTrain + initialEmbeddingValue (common/trainer.cpp:34-58,
 *                                   transe/trainer.cpp:21, transh/trainer.cpp:61,77-88, transr/trainer.cpp:66-86)
 *   kb2e_upload / kb2e_download     the nested-vector tables entityVec_/relationVec_/weights_
 *                                   (common/trainer.h:36-37, transh/trainer.h:17, transr/trainer.h:31),
 *                                   TransR seeding (transr/trainer.cpp:88-113), loadEmbeddings
 *                                   (common/evaluation.cpp:74-105), write (common/trainer.cpp:109-127)
 *   kb2e_set_replicas /
 *   kb2e_select_replica             (no counterpart: one reference process = one model)
 *   kb2e_train_epochs               Trainer::bfgs (common/trainer.cpp:69-107) incl. the sampler (:78-98),
 *                                   train_kb (:130-149), <model>::gradientUpdate / prebatch / postbatch
 *   kb2e_score                      <model>::tripleEnergy (transe/transe.cpp:10, transh/transh.cpp:10, transr/transr.cpp:13)
 *   kb2e_set_test_triples /
 *   kb2e_add_filter_triples         EmbeddingEvaluation::loadTriples / add (common/evaluation.cpp:41-72)
 *   kb2e_rank                       EmbeddingEvaluation::run + evalCorruption (common/evaluation.cpp:124-251)
 *   kb2e_sample_batch,
 *   kb2e_train_batch_pairs,
 *   kb2e_train_batch_deltas         test hooks onto the sampler, onto one batch of train_kb calls, and onto the
 *                                   accumulation half of gradientUpdate alone
 *
 * Threading: one context per host thread; a context owns one CUDA device, its own stream and all
 * device memory.  Multi-GPU = one process (context) per GPU; ranks shard queries through
 * kb2e_rank's [first, first+count) window and add the four int64 sums with their own collective.
 */
#ifndef KB2E_B200_H_
#define KB2E_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB2E_MODEL_TRANSE 0
#define KB2E_MODEL_TRANSH 1
#define KB2E_MODEL_TRANSR 2

#define KB2E_METHOD_UNIF 0 /* common/constants.h:8-9 */
#define KB2E_METHOD_BERN 1
#define KB2E_DISTANCE_L1 0 /* common/constants.h:16-17; "L2" is the SQUARED L2 distance */
#define KB2E_DISTANCE_L2 1

#define KB2E_TABLE_ENTITY 0   /* [num_entities][dim] */
#define KB2E_TABLE_RELATION 1 /* [num_relations][dim] */
#define KB2E_TABLE_WEIGHTS 2  /* TransH: [num_relations][dim]; TransR: [num_relations*dim][dim] = M[r][j=in][i=out] */

#define KB2E_OK 0
#define KB2E_ERR_ARG 1    /* bad argument / wrong call order */
#define KB2E_ERR_CUDA 2   /* CUDA runtime failure (message in kb2e_last_error) */
#define KB2E_ERR_NO_GPU 3 /* no usable CUDA device: there is NO CPU fallback */
#define KB2E_ERR_LIMIT 4  /* size outside what the kernels support */
#define KB2E_ERR_PEER 5   /* partitioned training: a peer GPU never arrived; every rank abandons the launch and reports this */

/* flags */
#define KB2E_FLAG_RANK_EXACT_ONLY 1u /* rank with the exact fp64 kernel only (no fp32 / tensor-core pre-filter) */
#define KB2E_FLAG_TRANSR_NO_QUIRK 2u /* do not replicate transr/trainer.cpp:187 (constraint on entity[relation]) */
/* Sampler indices with the DISTRIBUTION of the reference's randMax (common/utils.cpp:113-120: the wrapped 32-bit
 * product of two rand() values modulo x -- e.g. 75 % even indices) instead of uniform ones; still the counter RNG.
 * For trained-model parity with the shipped reference; needs n_train < 2^31. */
#define KB2E_FLAG_SAMPLER_RANDMAX 4u
/* Reproducible training (TransE, TransH): updates are accumulated as 32-bit fixed-point integers (2^-24 units) instead of
 * floating-point REDs, and the epoch loss as a 64-bit fixed-point sum, so tables, losses and output files are
 * bit-identical from run to run for a given seed -- what the single-threaded reference gives its users.  Costs four
 * scalar REDs per float4; needs |accumulated update| < 128 per row and batch. */
#define KB2E_FLAG_DETERMINISTIC 8u

typedef struct kb2e_ctx kb2e_ctx;

typedef struct kb2e_config {
   int32_t model;    /* KB2E_MODEL_* */
   int32_t dim;      /* -size     (common/constants.h:45) */
   int32_t method;   /* -method   0 unif, 1 bern */
   int32_t distance; /* -distance 0 L1, 1 squared L2 (ignored by TransH, transh/transh.cpp:24-26) */
   int32_t batches;  /* -batches  */
   int32_t device;   /* CUDA device ordinal */
   int64_t num_entities;
   int64_t num_relations;
   double rate;   /* -rate   */
   double margin; /* -margin */
   uint64_t seed; /* -seed: keys the counter-based RNG (sampler and initialisation) */
   uint32_t flags;
   uint32_t reserved;
} kb2e_config;

typedef struct kb2e_train_stats {
   uint64_t samples;      /* (pos, neg) pairs processed since the context was created */
   uint64_t active;       /* pairs whose hinge was active (normal + margin > corrupted) */
   uint64_t touched_ent;  /* entity rows renormalised + published, summed over batches */
   uint64_t touched_rel;  /* relation-side rows renormalised + published, summed over batches */
   uint64_t launches;     /* kernels launched by kb2e_train_epochs / kb2e_train_batch_pairs */
   double kernel_ms;      /* CUDA-event time of those launches */
} kb2e_train_stats;

typedef struct kb2e_rank_stats {
   uint64_t queries;      /* queries ranked since the context was created */
   uint64_t rechecked;    /* (query, candidate) pairs re-scored in exact fp64 after a pre-filter */
   uint64_t launches;
   double kernel_ms;      /* CUDA-event time of all ranking kernels */
   double main_kernel_ms; /* of which: the all-candidates scoring kernel(s) */
   double project_ms;        /* TransR: operand preparation + tensor-core projection, all passes */
   double project_kernel_ms; /* TransR: the tensor-core projection kernel alone (last pass of each call) */
} kb2e_rank_stats;

int kb2e_create(const kb2e_config* cfg, kb2e_ctx** out);
void kb2e_destroy(kb2e_ctx* ctx);
/* Message of the last failure on ctx (or of the last failed kb2e_create when ctx is NULL). */
const char* kb2e_last_error(const kb2e_ctx* ctx);
/* The context's CUDA stream (a cudaStream_t), so callers can time with events on the launching stream. */
void* kb2e_stream(kb2e_ctx* ctx);

/* ---- training ------------------------------------------------------------------------------- */
int kb2e_set_train_triples(kb2e_ctx* ctx, const int32_t* heads, const int32_t* tails, const int32_t* relations, int64_t n);
/* Per relation: mean #triples per distinct head / per distinct tail (0 for unused relations). */
int kb2e_set_bern(kb2e_ctx* ctx, const double* head_mean, const double* tail_mean);
int kb2e_init_embeddings(kb2e_ctx* ctx);
int kb2e_upload(kb2e_ctx* ctx, int table, const double* host, int64_t rows, int64_t cols);
int kb2e_download(kb2e_ctx* ctx, int table, double* host, int64_t rows, int64_t cols);
/* Epochs [first_epoch, first_epoch + n_epochs): every batch of every epoch runs on the device in
 * one persistent launch.  loss_per_epoch[n_epochs] = the value the reference prints per epoch. */
int kb2e_train_epochs(kb2e_ctx* ctx, int32_t first_epoch, int32_t n_epochs, double* loss_per_epoch);
int kb2e_get_train_stats(kb2e_ctx* ctx, kb2e_train_stats* out);

/* ---- batched training: several models in one launch (seed / rate / margin sweeps; TransE) ------------------------
 * The reference trains one model per process and has no sweep driver; its users run it once per setting
 * (common/args.cpp flags -rate, -margin, -seed).  At FB15k / WN18 shape one model leaves most of a B200 idle (a batch is
 * a chain of L2 round trips and grid barriers), so n_models independent models of one size on one KG can share every
 * launch, phase and barrier: each computes exactly what it would compute alone with its seed / rate / margin.
 *   kb2e_set_replicas     right after kb2e_create, before any table exists.  rates / margins / seeds: n_models values
 *                         each, or NULL for the context's own (seeds default to seed + index).
 *   kb2e_init_embeddings  initialises every model from its own seed.
 *   kb2e_train_epochs     trains all of them; loss_per_epoch receives n_models x n_epochs values, model-major.
 *   kb2e_select_replica   chooses the model that kb2e_upload / kb2e_download / kb2e_score / kb2e_rank address (0 at first). */
int kb2e_set_replicas(kb2e_ctx* ctx, int32_t n_models, const double* rates, const double* margins, const uint64_t* seeds);
int kb2e_select_replica(kb2e_ctx* ctx, int32_t model_index);

/* ---- scoring / ranking ---------------------------------------------------------------------- */
/* Energies of n triples from the CURRENT tables: precision 0 = the fp32 scoring code the training
 * kernels use, 1 = the exact fp64 code the ranking uses (reference operation order). */
int kb2e_score(kb2e_ctx* ctx, const int32_t* heads, const int32_t* tails, const int32_t* relations, int64_t n,
               int32_t precision, double* out);
int kb2e_set_test_triples(kb2e_ctx* ctx, const int32_t* heads, const int32_t* tails, const int32_t* relations, int64_t n);
/* Known-true triples besides the test set (train, valid); may be called repeatedly.  n = 0 with
 * NULL pointers clears the set. */
int kb2e_add_filter_triples(kb2e_ctx* ctx, const int32_t* heads, const int32_t* tails, const int32_t* relations, int64_t n);
/* Rank test triples [first, first+count): entry 2*i is the head corruption of triple first+i,
 * 2*i+1 its tail corruption.  rank = 1 + #{candidates with strictly lower energy}; *_ties = number
 * of OTHER candidates with exactly equal energy (the reference's std::sort places the truth
 * anywhere inside [rank, rank+ties]).  sums = {raw rank sum, filtered rank sum, raw hits@10,
 * filtered hits@10}.  Any output pointer may be NULL. */
int kb2e_rank(kb2e_ctx* ctx, int64_t first, int64_t count,
              int32_t* raw_rank, int32_t* filt_rank, int32_t* raw_ties, int32_t* filt_ties, int64_t sums[4]);
int kb2e_get_rank_stats(kb2e_ctx* ctx, kb2e_rank_stats* out);

/* ---- entity-partitioned training across the GPUs of one NVLink box (TransE; BASELINE configs[4]) ----
 * One context per process per GPU.  Entity row e lives on rank e % world at local index e / world;
 * relation rows are replicated.  Train triples and bern statistics are set per rank with
 * kb2e_set_train_triples / kb2e_set_bern BEFORE kb2e_dist_setup (the batch size sizes the exchange
 * buffers).  kb2e_dist_setup allocates this rank's arena and returns a 64-byte CUDA IPC handle; the
 * caller exchanges the handles of all ranks (any transport) and passes the world x 64 bytes, in rank
 * order, to kb2e_dist_connect.  All ranks must then call kb2e_dist_train_epochs with the same
 * arguments: the kernels exchange row requests, rows and updates with posted peer stores / vector
 * REDs over NVLink and synchronise on peer-mapped counters; no host step in between.
 * loss_per_epoch receives this rank's share of the loss (add across ranks).
 * Failure is a property of the JOB, not of a rank: a rank whose buffers overflowed raises an error word on every
 * peer, so all ranks return KB2E_ERR_LIMIT together; a rank that never launches (its call failed before the
 * launch) is noticed by the others at their next cross-GPU barrier, which waits at most KB2E_DIST_TIMEOUT_MS
 * (environment, default 30000) -- they abandon the launch and return KB2E_ERR_PEER instead of hanging the GPUs.
 * Tables: KB2E_TABLE_ENTITY = this rank's rows [ceil((N_E - rank) / world)][dim] (global ids rank,
 * rank + world, ...); KB2E_TABLE_RELATION = the replica [N_R][dim]. */
#define KB2E_DIST_HANDLE_BYTES 64
int kb2e_dist_setup(kb2e_ctx* ctx, int32_t rank, int32_t world, void* handle_out);
int kb2e_dist_connect(kb2e_ctx* ctx, const void* handles);
int kb2e_dist_init_embeddings(kb2e_ctx* ctx);
int kb2e_dist_upload(kb2e_ctx* ctx, int table, const double* host, int64_t rows, int64_t cols);
int kb2e_dist_download(kb2e_ctx* ctx, int table, double* host, int64_t rows, int64_t cols);
int kb2e_dist_train_epochs(kb2e_ctx* ctx, int32_t first_epoch, int32_t n_epochs, double* loss_per_epoch);
void kb2e_dist_teardown(kb2e_ctx* ctx);

/* ---- test hooks ----------------------------------------------------------------------------- */
/* The device sampler's output for samples [0, count) of (epoch, batch): count x {h,t,r,h',t',r'}. */
int kb2e_sample_batch(kb2e_ctx* ctx, int32_t epoch, int32_t batch, int64_t count, int32_t* pairs_out);
/* One batch made of the given (pos, neg) pairs instead of sampled ones, through the same kernel. */
int kb2e_train_batch_pairs(kb2e_ctx* ctx, const int32_t* pairs, int64_t n, double* loss, int64_t* n_active);
/* The same pairs through the same kernel, stopped after the accumulation phase: returns the summed PRE-normalisation
 * updates -- what <model>::Trainer::gradientUpdate adds to *_next_ before its norm() calls (transe/trainer.cpp:37-41,
 * transh/trainer.cpp:25-45, transr/trainer.cpp:158-172) -- as [num_entities][dim], [num_relations][dim] and the
 * KB2E_TABLE_WEIGHTS shape (any may be NULL).  The tables themselves are left untouched. */
int kb2e_train_batch_deltas(kb2e_ctx* ctx, const int32_t* pairs, int64_t n, double* d_entity, double* d_relation,
                            double* d_weights, double* loss, int64_t* n_active);

/* TransR: the tensor-core (bf16 hi/lo split, fp32 accumulate) projection M_r^T e of every entity under `relation` as
 * the ranking pre-filter computes it, proj[num_entities][dim], and the relative error bound the ranking assumes for it:
 * |proj[c][i] - exact| <= *eps_rel * sum_j |e_cj| |M_r[j][i]| (transr/transr.cpp:20-25 is the exact form). */
int kb2e_debug_transr_projection(kb2e_ctx* ctx, int32_t relation, float* proj, double* eps_rel);

#ifdef __cplusplus
}
#endif
#endif /* KB2E_B200_H_ */
