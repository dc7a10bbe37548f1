/* TEST INFRASTRUCTURE ONLY -- the product path never includes, links or calls this.
 *
 * kb2e_oracle: a plain-C, fp64, single-threaded CPU restatement of the KB2E hot path
 * (eriq-augustine/KB2E): per-triple energies, gradient updates, normalisations, the hinge,
 * the all-entity ranking, and -- in the "dfr" section -- the batch-deferred / counter-RNG
 * semantics the CUDA kernels implement, so the kernels can be checked sample by sample.
 *
 * PINNING: the reference ships no tests, golden vectors or fixtures, so this
 * oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF compiled here from /root/reference
 * (oracle/_ref/libkb2e_ref.so, built by oracle/Makefile): tests/golden/ holds fixtures generated
 * by tests/golden/make_golden.py through that library, and tests/test_oracle_vs_reference.py
 * demands BITWISE equality for every "ref" function below on those fixtures (and live against
 * libkb2e_ref.so whenever it is present).
 *
 * Layout: tables are row-major double[rows][D]; the TransR matrices are double[nR][D][D] indexed
 * [relation][j = input dim][i = output dim] exactly like weights_[r][j][i] (transr/trainer.h:31).
 * model: 0 TransE, 1 TransH, 2 TransR.  distance: 0 L1, 1 squared L2 (common/constants.h:16-17).
 */
#ifndef KB2E_ORACLE_H_
#define KB2E_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference semantics ("ref") ------------------------------------------------------------ */

/* transe/transe.cpp:10-28, transh/transh.cpp:10-29, transr/transr.cpp:13-37 (work vectors zeroed). */
double orc_energy(int model, int distance, int D, const double* ent, const double* rel, const double* w,
                  int head, int tail, int relation);
/* transr/transr.cpp:13-37 AS SHIPPED: accumulates into caller-owned work vectors (never zeroed). */
double orc_energy_transr_shipped(int distance, int D, const double* ent, const double* rel, const double* w,
                                 int head, int tail, int relation, double* head_work, double* tail_work);
void orc_energy_many(int model, int distance, int D, const double* ent, const double* rel, const double* w,
                     long n, const int* h, const int* t, const int* r, double* out);

double orc_vec_len(const double* a, int n);                       /* common/utils.cpp:44-51 */
void orc_norm(double* a, int n, int ignore_short);                /* common/utils.cpp:70-77 */
void orc_norm2(double* a, double* b, int n, double rate);         /* common/utils.cpp:79-111 */
void orc_transr_norm(double* a, double* M, int D, double lr);     /* transr/trainer.cpp:35-64 */

/* gradientUpdate on the *_next_ tables, directions from the cur tables
 * (transe/trainer.cpp:25-46, transh/trainer.cpp:11-59, transr/trainer.cpp:144-188 incl. the
 * entityVec_next_[relation] quirk at :187). */
void orc_grad(int model, int distance, int D, double lr,
              const double* ent, const double* rel, const double* w,
              double* ent_next, double* rel_next, double* w_next,
              int head, int tail, int relation, int corrupted);
/* The accumulation part of orc_grad only (no normalisation tail), and the tail alone: raw + tail == orc_grad,
 * so next - cur after orc_grad_raw is the pre-normalisation update the CUDA kernels RED into their delta tables. */
void orc_grad_raw(int model, int distance, int D, double lr,
                  const double* ent, const double* rel, const double* w,
                  double* ent_next, double* rel_next, double* w_next,
                  int head, int tail, int relation, int corrupted);
void orc_grad_tail(int model, int D, double lr, double* ent_next, double* rel_next, double* w_next,
                   int head, int tail, int relation);

/* One reference batch: next = cur; for each pair in order: train_kb (common/trainer.cpp:130-149).
 * pairs = n x {h,t,r,h',t',r'}.  Returns the summed loss; losses[k] per pair if not NULL. */
double orc_train_batch_ref(int model, int distance, int D, int nE, int nR, double lr, double margin,
                           const double* ent, const double* rel, const double* w,
                           long n, const int* pairs,
                           double* ent_next, double* rel_next, double* w_next, double* losses);

/* bern statistics, common/trainer.cpp:163-194: mean triples per distinct head / tail, per relation. */
void orc_bern(long n, const int* h, const int* t, const int* r, int nR, double* head_mean, double* tail_mean);

/* All-entity ranking, common/evaluation.cpp:124-179, one entry per query q = 2*i + side
 * (side 0 = head corruption, 1 = tail corruption; :230-238).  The reference's std::sort leaves the
 * order of exact ties arbitrary, so the oracle reports the interval a conforming rank must lie in:
 *   raw_lo  = 1 + #{c : E_c <  E_true}            raw_hi  = raw_lo  + #{c != true : E_c == E_true}
 *   filt_lo = 1 + #{c unknown : E_c < E_true}     filt_hi = filt_lo + #{c != true, unknown : E_c == E_true}
 * "unknown" = corrupted triple not in test + filter triples (:161). */
void orc_rank(int model, int distance, int D, int nE, int nR,
              const double* ent, const double* rel, const double* w,
              long nTest, const int* th, const int* tt, const int* tr,
              long nFilter, const int* fh, const int* ft, const int* fr,
              int* raw_lo, int* raw_hi, int* filt_lo, int* filt_hi);

/* ---- the semantics the CUDA path implements ("dfr": deferred renorm + counter RNG) ------------ */

/* Philox4x32-10; ctr/key as in Salmon et al. (SC'11).  out[4]. */
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out);

typedef struct orc_sampler orc_sampler;
/* Train-set membership + bern table for the sampler.  tail_pr[r] = the reference's `pr`
 * (common/trainer.cpp:82-86): 1000*T/(T+H) for bern, 500 for unif. */
orc_sampler* orc_sampler_create(long n, const int* h, const int* t, const int* r, int nE, int nR, int method);
void orc_sampler_destroy(orc_sampler*);
const double* orc_sampler_pr(const orc_sampler*);
/* mode 0 (default): uniform indices; mode 1: the index distribution of the reference's randMax
 * (common/utils.cpp:113-120: the wrapped 32-bit product of two rand() values, modulo x), driven by the counter RNG. */
void orc_sampler_set_mode(orc_sampler*, int mode);
void orc_randmax_draws(uint64_t seed, int x, long n, int* out);   /* the mode-1 index generator alone */
/* Restates common/trainer.cpp:78-98 with the counter RNG: sample k of global batch gb draws
 * block = philox(k, gb, attempt, 0; seed): i = mulhi64(x0:x1, n), coin = x2 % 1000, j = mulhi32(x3, nE);
 * resample j = mulhi32(philox(k, gb, a, 0).x0, nE) for a = 1.. while the corrupted triple is in train
 * (at most 64 attempts).  pairs_out = count x {h,t,r,h',t',r'}. */
void orc_sample_batch(const orc_sampler*, uint64_t seed, uint32_t global_batch, long count, int* pairs_out);

/* One batch with the deferred semantics: directions and energies from cur (fp64 here), all deltas
 * accumulated, then every touched row normalised ONCE (instead of after every update):
 *   2a relation-side rows: d_r/r clip (E,H) or unit (R); w_r unit + norm(d_r,w_r,lr) (H); M_r += dM, rows unit (R)
 *   2b entity rows: clip (E,H) or unit (R); then the soft constraint (H: norm(e,w_r,lr); R: transRNorm(e,M_r))
 *      against the lowest and highest relation id that touched the row, with w_r/M_r read-only; the
 *      perturbation the reference applies to w_r/M_r inside those loops goes to `carry` and enters
 *      the NEXT batch's delta.  carry: in/out, same shape as w (NULL for TransE / to drop it).
 * In-place on ent/rel/w.  Returns summed loss; *n_active = hinge-active pairs. */
double orc_train_batch_dfr(int model, int distance, int D, int nE, int nR, double lr, double margin,
                           double* ent, double* rel, double* w, double* carry,
                           long n, const int* pairs, long* n_active);

/* epochs x batches of orc_sample_batch + orc_train_batch_dfr (the CPU port of the product path;
 * bench.py's cpu_baseline "port" leg).  loss_out[epochs]. */
void orc_set_dfr_no_carry(int on);   /* study switch: see kb2e_oracle.c */
void orc_train_epochs_dfr(const orc_sampler*, int model, int distance, int D, int nE, int nR,
                          double lr, double margin, int batches, int first_epoch, int epochs, uint64_t seed,
                          double* ent, double* rel, double* w, double* loss_out);

/* Same loop with the REFERENCE's sequential batch semantics (orc_train_batch_ref) under the uniform
 * counter sampler: isolates the effect of the reference's non-uniform randMax from everything else. */
void orc_train_epochs_ref(const orc_sampler*, int model, int distance, int D, int nE, int nR,
                          double lr, double margin, int batches, int first_epoch, int epochs, uint64_t seed,
                          double* ent, double* rel, double* w, double* loss_out);

#ifdef __cplusplus
}
#endif
#endif  /* KB2E_ORACLE_H_ */
