// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI harness around the UNMODIFIED reference (eriq-augustine/KB2E), compiled from the
// sources where they lie under /root/reference by oracle/Makefile; the output
// (oracle/_ref/libkb2e_ref.so) is git-ignored and travels to the GPU box.  Nothing from the
// reference is copied here: this file only #includes its headers and reaches protected members
// through derived classes (the reference's own plug-in seam, common/trainer.h:59-77,
// common/evaluation.h:50-61).
//
// What it exposes:
//   ref_energy      -> transe|transh|transr::tripleEnergy        (transe/transe.cpp:10, transh/transh.cpp:10, transr/transr.cpp:13)
//   ref_grad        -> <model>::Trainer::prebatch + gradientUpdate (transe/trainer.cpp:25, transh/trainer.cpp:11, transr/trainer.cpp:144)
//   ref_train_batch -> prebatch + N x common::Trainer::train_kb   (common/trainer.cpp:130) -- the reference's sequential batch semantics
//   ref_norm/ref_norm2 -> common::norm overloads                  (common/utils.cpp:70, :79)
//   ref_randmax     -> common::randMax                              (common/utils.cpp:113)
//   ref_rank        -> common::EmbeddingEvaluation::evalCorruption (common/evaluation.cpp:124) per query
//   ref_train_files -> loadFiles + prepTrain + bfgs (+write)       (common/trainer.cpp:151,34,69,109) timed around bfgs
//   ref_bern        -> the per-relation statistics loadFiles computes (common/trainer.cpp:171-194)
//
// TransR: transr::tripleEnergy accumulates into caller-owned, never-zeroed work vectors
// (transr/transr.cpp:20-25).  `zero_work=1` zeroes them before every call (the "zero-patched"
// behaviour SURVEY.md section 8c names as the TransR parity target); `zero_work=0` is as shipped.

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "common/args.h"
#include "common/constants.h"
#include "common/evaluation.h"
#include "common/trainer.h"
#include "common/utils.h"
#include "transe/evaluation.h"
#include "transe/trainer.h"
#include "transe/transe.h"
#include "transh/trainer.h"
#include "transh/transh.h"
#include "transr/trainer.h"
#include "transr/transr.h"

namespace {

typedef std::vector<std::vector<double>> Table;
typedef std::vector<std::vector<std::vector<double>>> Table3;

void fill(Table& t, const double* src, int rows, int cols) {
   t.assign(rows, std::vector<double>(cols));
   for (int i = 0; i < rows; i++) {
      for (int j = 0; j < cols; j++) {
         t[i][j] = src[(size_t)i * cols + j];
      }
   }
}

void fill3(Table3& t, const double* src, int n, int d) {
   t.assign(n, Table(d, std::vector<double>(d)));
   for (int i = 0; i < n; i++) {
      for (int j = 0; j < d; j++) {
         for (int k = 0; k < d; k++) {
            t[i][j][k] = src[((size_t)i * d + j) * d + k];
         }
      }
   }
}

void dump(const Table& t, double* dst) {
   if (dst == NULL) return;
   size_t cols = t.empty() ? 0 : t[0].size();
   for (size_t i = 0; i < t.size(); i++) {
      for (size_t j = 0; j < cols; j++) {
         dst[i * cols + j] = t[i][j];
      }
   }
}

void dump3(const Table3& t, double* dst) {
   if (dst == NULL) return;
   size_t d = t.empty() ? 0 : t[0].size();
   for (size_t i = 0; i < t.size(); i++) {
      for (size_t j = 0; j < d; j++) {
         for (size_t k = 0; k < d; k++) {
            dst[(i * d + j) * d + k] = t[i][j][k];
         }
      }
   }
}

common::EmbeddingArguments makeArgs(int D, double lr, double margin, int method, int distance, int batches, int epochs) {
   common::EmbeddingArguments a;
   a.embeddingSize = D;
   a.learningRate = lr;
   a.margin = margin;
   a.method = method;
   a.distanceType = distance;
   a.numBatches = batches;
   a.maxEpochs = epochs;
   return a;
}

// ---- trainer probes ---------------------------------------------------------------------------

struct ProbeE : transe::Trainer {
   explicit ProbeE(common::EmbeddingArguments a) : transe::Trainer(a) {}
   void set(int nE, int nR, const double* ent, const double* rel) {
      numEntities_ = nE;
      numRelations_ = nR;
      fill(entityVec_, ent, nE, embeddingSize_);
      fill(relationVec_, rel, nR, embeddingSize_);
   }
   void pre() { prebatch(); }
   void post() { postbatch(); }
   void grad(int h, int t, int r, bool c) { gradientUpdate(h, t, r, c); }
   double kb(const int* p) { return train_kb(p[0], p[1], p[2], p[3], p[4], p[5]); }
   void outNext(double* e, double* r, double*) { dump(entityVec_next_, e); dump(relationVec_next_, r); }
   void outCur(double* e, double* r, double*) { dump(entityVec_, e); dump(relationVec_, r); }
   void prep() { prepTrain(); }
   void loop() { bfgs(); }
};

struct ProbeH : transh::Trainer {
   explicit ProbeH(common::EmbeddingArguments a) : transh::Trainer(a) {}
   void set(int nE, int nR, const double* ent, const double* rel, const double* w) {
      numEntities_ = nE;
      numRelations_ = nR;
      fill(entityVec_, ent, nE, embeddingSize_);
      fill(relationVec_, rel, nR, embeddingSize_);
      fill(weights_, w, nR, embeddingSize_);
   }
   void pre() { prebatch(); }
   void post() { postbatch(); }
   void grad(int h, int t, int r, bool c) { gradientUpdate(h, t, r, c); }
   double kb(const int* p) { return train_kb(p[0], p[1], p[2], p[3], p[4], p[5]); }
   void outNext(double* e, double* r, double* w) { dump(entityVec_next_, e); dump(relationVec_next_, r); dump(weights_next_, w); }
   void outCur(double* e, double* r, double* w) { dump(entityVec_, e); dump(relationVec_, r); dump(weights_, w); }
   void prep() { prepTrain(); }
   void loop() { bfgs(); }
};

struct ProbeR : transr::Trainer {
   bool zeroWork;
   ProbeR(common::EmbeddingArguments a, bool zw) : transr::Trainer(a), zeroWork(zw) {}
   void set(int nE, int nR, const double* ent, const double* rel, const double* w) {
      numEntities_ = nE;
      numRelations_ = nR;
      fill(entityVec_, ent, nE, embeddingSize_);
      fill(relationVec_, rel, nR, embeddingSize_);
      fill3(weights_, w, nR, embeddingSize_);
   }
   double tripleEnergy(int head, int tail, int relation) override {
      if (zeroWork) {
         headWorkVec_.assign(embeddingSize_, 0.0);
         tailWorkVec_.assign(embeddingSize_, 0.0);
      }
      return transr::Trainer::tripleEnergy(head, tail, relation);
   }
   void pre() { prebatch(); }
   void post() { postbatch(); }
   void grad(int h, int t, int r, bool c) { gradientUpdate(h, t, r, c); }
   double kb(const int* p) { return train_kb(p[0], p[1], p[2], p[3], p[4], p[5]); }
   void outNext(double* e, double* r, double* w) { dump(entityVec_next_, e); dump(relationVec_next_, r); dump3(weights_next_, w); }
   void outCur(double* e, double* r, double* w) { dump(entityVec_, e); dump(relationVec_, r); dump3(weights_, w); }
   void prep() { prepTrain(); }
   void loop() { bfgs(); }
};

// ---- evaluation probes ------------------------------------------------------------------------

// Shared: drive the reference's evalCorruption one query at a time (cache disabled; it never
// changes results, common/evaluation.cpp:107-120).
template <class Base>
struct EvalDriver : Base {
   explicit EvalDriver(common::EmbeddingArguments a) : Base(a) {}
   void setTables(int nE, int nR, const double* ent, const double* rel) {
      this->numEntities_ = nE;
      this->numRelations_ = nR;
      fill(this->entityVec_, ent, nE, this->embeddingSize_);
      fill(this->relationVec_, rel, nR, this->embeddingSize_);
   }
   void addTriple(int h, int t, int r, bool working) { this->add(h, t, r, working); }
   void rankAll(int* raw, int* filt) {
      std::vector<std::pair<int, double>> work(this->numEntities_);
      for (size_t i = 0; i < this->heads_.size(); i++) {
         for (int side = 0; side < 2; side++) {
            int rawSum = 0, filtSum = 0, rawHit = 0, filtHit = 0;
            this->evalCorruption(this->heads_[i], this->tails_[i], this->relations_[i], side == 0,
                                 &rawSum, &filtSum, &rawHit, &filtHit, work);
            raw[2 * i + side] = rawSum;
            filt[2 * i + side] = filtSum;
         }
      }
   }
};

struct EvalH : common::EmbeddingEvaluation {
   Table weights;
   explicit EvalH(common::EmbeddingArguments a) : common::EmbeddingEvaluation(a) {}
   double tripleEnergy(int head, int tail, int relation) override {
      return transh::tripleEnergy(head, tail, relation, embeddingSize_, entityVec_, relationVec_, weights);
   }
};

struct EvalR : common::EmbeddingEvaluation {
   Table3 weights;
   int distanceType;
   bool zeroWork;
   std::vector<double> headWork, tailWork;
   explicit EvalR(common::EmbeddingArguments a)
         : common::EmbeddingEvaluation(a), distanceType(a.distanceType), zeroWork(true),
           headWork(a.embeddingSize), tailWork(a.embeddingSize) {}
   double tripleEnergy(int head, int tail, int relation) override {
      if (zeroWork) {
         headWork.assign(embeddingSize_, 0.0);
         tailWork.assign(embeddingSize_, 0.0);
      }
      return transr::tripleEnergy(head, tail, relation, embeddingSize_, entityVec_, relationVec_, weights,
                                  distanceType, headWork, tailWork);
   }
};

template <class P>
void batchRun(P& p, long nPairs, const int* pairs, double* losses) {
   p.pre();
   for (long k = 0; k < nPairs; k++) {
      double l = p.kb(pairs + 6 * k);
      if (losses) losses[k] = l;
   }
}

template <class P>
double timedLoop(P& p, bool doWrite) {
   p.prep();
   auto t0 = std::chrono::steady_clock::now();
   p.loop();
   auto t1 = std::chrono::steady_clock::now();
   if (doWrite) p.write();
   return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

int ref_energy(int model, int distance, int D, int nE, int nR,
               const double* ent, const double* rel, const double* w,
               long n, const int* h, const int* t, const int* r, int zero_work, double* out) {
   Table E, R;
   fill(E, ent, nE, D);
   fill(R, rel, nR, D);
   if (model == 0) {
      for (long k = 0; k < n; k++) {
         out[k] = transe::tripleEnergy(h[k], t[k], r[k], D, E, R, distance == L1_DISTANCE);
      }
   } else if (model == 1) {
      Table W;
      fill(W, w, nR, D);
      for (long k = 0; k < n; k++) {
         out[k] = transh::tripleEnergy(h[k], t[k], r[k], D, E, R, W);
      }
   } else if (model == 2) {
      Table3 W;
      fill3(W, w, nR, D);
      std::vector<double> hv(D, 0.0), tv(D, 0.0);
      for (long k = 0; k < n; k++) {
         if (zero_work) {
            hv.assign(D, 0.0);
            tv.assign(D, 0.0);
         }
         out[k] = transr::tripleEnergy(h[k], t[k], r[k], D, E, R, W, distance, hv, tv);
      }
   } else {
      return 1;
   }
   return 0;
}

// prebatch(); gradientUpdate(head, tail, relation, corrupted); return the *_next_ tables.
int ref_grad(int model, int distance, int D, int nE, int nR, double lr,
             const double* ent, const double* rel, const double* w,
             int head, int tail, int relation, int corrupted,
             double* ent_next, double* rel_next, double* w_next) {
   common::EmbeddingArguments a = makeArgs(D, lr, 1.0, 0, distance, 1, 1);
   if (model == 0) {
      ProbeE p(a);
      p.set(nE, nR, ent, rel);
      p.pre();
      p.grad(head, tail, relation, corrupted != 0);
      p.outNext(ent_next, rel_next, NULL);
   } else if (model == 1) {
      ProbeH p(a);
      p.set(nE, nR, ent, rel, w);
      p.pre();
      p.grad(head, tail, relation, corrupted != 0);
      p.outNext(ent_next, rel_next, w_next);
   } else if (model == 2) {
      ProbeR p(a, true);
      p.set(nE, nR, ent, rel, w);
      p.pre();
      p.grad(head, tail, relation, corrupted != 0);
      p.outNext(ent_next, rel_next, w_next);
   } else {
      return 1;
   }
   return 0;
}

// One reference batch: prebatch(), then train_kb on each (pos, neg) pair IN ORDER (the reference's
// sequential semantics: directions from the snapshot, accumulate+normalise on *_next_).
// pairs = nPairs x {h, t, r, h', t', r'}.  Returns the *_next_ tables and per-pair losses.
int ref_train_batch(int model, int distance, int D, int nE, int nR, double lr, double margin,
                    const double* ent, const double* rel, const double* w,
                    long nPairs, const int* pairs, int zero_work,
                    double* ent_next, double* rel_next, double* w_next, double* losses) {
   common::EmbeddingArguments a = makeArgs(D, lr, margin, 0, distance, 1, 1);
   if (model == 0) {
      ProbeE p(a);
      p.set(nE, nR, ent, rel);
      batchRun(p, nPairs, pairs, losses);
      p.outNext(ent_next, rel_next, NULL);
   } else if (model == 1) {
      ProbeH p(a);
      p.set(nE, nR, ent, rel, w);
      batchRun(p, nPairs, pairs, losses);
      p.outNext(ent_next, rel_next, w_next);
   } else if (model == 2) {
      ProbeR p(a, zero_work != 0);
      p.set(nE, nR, ent, rel, w);
      batchRun(p, nPairs, pairs, losses);
      p.outNext(ent_next, rel_next, w_next);
   } else {
      return 1;
   }
   return 0;
}

void ref_norm(double* a, int n, int ignore_short) {
   std::vector<double> v(a, a + n);
   common::norm(v, ignore_short != 0);
   std::memcpy(a, v.data(), sizeof(double) * n);
}

void ref_norm2(double* a, double* b, int n, double rate) {
   std::vector<double> va(a, a + n), vb(b, b + n);
   common::norm(va, vb, rate);
   std::memcpy(a, va.data(), sizeof(double) * n);
   std::memcpy(b, vb.data(), sizeof(double) * n);
}

// n draws of the reference's own randMax(x) (common/utils.cpp:113-120) after srand(seed): pins the DISTRIBUTION that
// the counter-RNG emulation (orc_sampler_set_mode(1), KB2E_FLAG_SAMPLER_RANDMAX) reproduces.
void ref_randmax(unsigned seed, int x, long n, int* out) {
   srand(seed);
   for (long i = 0; i < n; i++) out[i] = common::randMax(x);
}

double ref_vec_len(const double* a, int n) {
   std::vector<double> v(a, a + n);
   return common::vec_len(v);
}

// Per-query ranks from the reference's evalCorruption: out[2*i] = head corruption of test triple i,
// out[2*i+1] = tail corruption (the order run() uses, common/evaluation.cpp:230-238).
// Filter set = test + filter triples (common/evaluation.cpp:59-61).
int ref_rank(int model, int distance, int D, int nE, int nR,
             const double* ent, const double* rel, const double* w,
             long nTest, const int* th, const int* tt, const int* tr,
             long nFilter, const int* fh, const int* ft, const int* fr,
             int zero_work, int* raw, int* filt) {
   common::EmbeddingArguments a = makeArgs(D, 0.0, 1.0, 0, distance, 1, 1);
   if (model == 0) {
      EvalDriver<transe::Evaluation> ev(a);
      ev.setTables(nE, nR, ent, rel);
      for (long i = 0; i < nTest; i++) ev.addTriple(th[i], tt[i], tr[i], true);
      for (long i = 0; i < nFilter; i++) ev.addTriple(fh[i], ft[i], fr[i], false);
      ev.rankAll(raw, filt);
   } else if (model == 1) {
      EvalDriver<EvalH> ev(a);
      ev.setTables(nE, nR, ent, rel);
      fill(ev.weights, w, nR, D);
      for (long i = 0; i < nTest; i++) ev.addTriple(th[i], tt[i], tr[i], true);
      for (long i = 0; i < nFilter; i++) ev.addTriple(fh[i], ft[i], fr[i], false);
      ev.rankAll(raw, filt);
   } else if (model == 2) {
      EvalDriver<EvalR> ev(a);
      ev.zeroWork = zero_work != 0;
      ev.setTables(nE, nR, ent, rel);
      fill3(ev.weights, w, nR, D);
      for (long i = 0; i < nTest; i++) ev.addTriple(th[i], tt[i], tr[i], true);
      for (long i = 0; i < nFilter; i++) ev.addTriple(fh[i], ft[i], fr[i], false);
      ev.rankAll(raw, filt);
   } else {
      return 1;
   }
   return 0;
}

// Whole reference training run on files: loadFiles(); prepTrain(); bfgs(); [write()].
// Returns the wall seconds spent inside bfgs() only (load, rejection-sampled init and write
// excluded -- SURVEY.md 8d), or a negative value on error.  srand(seed) as the mains do
// (transe/bin/trainTransE.cpp:13).  The reference prints its per-epoch loss line to stdout.
double ref_train_files(int model, const char* datadir, const char* outdir, int D, double lr, double margin,
                       int method, int distance, int batches, int epochs, unsigned seed,
                       const char* seeddir, int seedmethod, int zero_work, int do_write) {
   common::EmbeddingArguments a = makeArgs(D, lr, margin, method, distance, batches, epochs);
   a.dataDir = datadir;
   a.outputDir = outdir;
   a.seedDataDir = seeddir ? seeddir : ".";
   a.seedMethod = seedmethod;
   a.seed = seed;
   srand(seed);
   double secs = -1.0;
   if (model == 0) {
      ProbeE p(a);
      p.loadFiles();
      secs = timedLoop(p, do_write != 0);
   } else if (model == 1) {
      ProbeH p(a);
      p.loadFiles();
      secs = timedLoop(p, do_write != 0);
   } else if (model == 2) {
      ProbeR p(a, zero_work != 0);
      p.loadFiles();
      secs = timedLoop(p, do_write != 0);
   }
   fflush(stdout);
   return secs;
}

// The bern statistics exactly as loadFiles() computes them (common/trainer.cpp:163-194), from
// in-memory triples: tail_mean[r] = #triples(r) / #distinct tails(r); head_mean likewise.
struct BernProbe : transe::Trainer {
   explicit BernProbe(common::EmbeddingArguments a) : transe::Trainer(a) {}
   void stats(const char* datadir, int nR, double* headMean, double* tailMean) {
      dataDir_ = datadir;
      loadFiles();
      for (int i = 0; i < nR; i++) {
         headMean[i] = relationHeadMeanCooccurrence_[i];
         tailMean[i] = relationTailMeanCooccurrence_[i];
      }
   }
};

int ref_bern(const char* datadir, int nR, double* head_mean, double* tail_mean) {
   BernProbe p(makeArgs(4, 0.0, 1.0, 1, 0, 1, 1));
   p.stats(datadir, nR, head_mean, tail_mean);
   fflush(stdout);
   return 0;
}

// ---- persistent reference trainer (bench.py --impl reference: load + init once, then time epochs) ----
struct RefTrainer {
   int model;
   ProbeE* e;
   ProbeH* h;
   ProbeR* r;
};

struct EpochsE : ProbeE { using ProbeE::ProbeE; void epochs(int n) { maxEpochs_ = n; bfgs(); } };
struct EpochsH : ProbeH { using ProbeH::ProbeH; void epochs(int n) { maxEpochs_ = n; bfgs(); } };
struct EpochsR : ProbeR { using ProbeR::ProbeR; void epochs(int n) { maxEpochs_ = n; bfgs(); } };

void* ref_trainer_create(int model, const char* datadir, const char* outdir, int D, double lr, double margin,
                         int method, int distance, int batches, unsigned seed, const char* seeddir, int seedmethod,
                         int zero_work) {
   common::EmbeddingArguments a = makeArgs(D, lr, margin, method, distance, batches, 1);
   a.dataDir = datadir;
   a.outputDir = outdir;
   a.seedDataDir = seeddir ? seeddir : ".";
   a.seedMethod = seedmethod;
   a.seed = seed;
   srand(seed);
   RefTrainer* t = new RefTrainer();
   t->model = model;
   t->e = NULL; t->h = NULL; t->r = NULL;
   if (model == 0) { t->e = new EpochsE(a); t->e->loadFiles(); t->e->prep(); }
   else if (model == 1) { t->h = new EpochsH(a); t->h->loadFiles(); t->h->prep(); }
   else { t->r = new EpochsR(a, zero_work != 0); t->r->loadFiles(); t->r->prep(); }
   fflush(stdout);
   return t;
}

// Runs n more epochs of the reference's bfgs() (the printed epoch index restarts at 0 every call);
// returns the wall seconds spent inside bfgs().
double ref_trainer_epochs(void* handle, int n) {
   RefTrainer* t = (RefTrainer*)handle;
   auto t0 = std::chrono::steady_clock::now();
   if (t->model == 0) static_cast<EpochsE*>(t->e)->epochs(n);
   else if (t->model == 1) static_cast<EpochsH*>(t->h)->epochs(n);
   else static_cast<EpochsR*>(t->r)->epochs(n);
   auto t1 = std::chrono::steady_clock::now();
   fflush(stdout);
   return std::chrono::duration<double>(t1 - t0).count();
}

void ref_trainer_write(void* handle) {
   RefTrainer* t = (RefTrainer*)handle;
   if (t->model == 0) t->e->write(); else if (t->model == 1) t->h->write(); else t->r->write();
}

void ref_trainer_destroy(void* handle) {
   RefTrainer* t = (RefTrainer*)handle;
   delete t->e; delete t->h; delete t->r;
   delete t;
}

}  // extern "C"
