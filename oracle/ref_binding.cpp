// TEST INFRASTRUCTURE ONLY -- the reference-side binding of INTEGRATION.md section 1, compiled for real.
//
// This is the file a maintainer of eriq-augustine/KB2E would add to the reference tree: one subclass per model that
// overrides the virtual common::Trainer::bfgs() (common/trainer.h:59) and forwards the epoch loop to the C ABI of
// libkb2e_b200.so; everything else -- argument parsing (common/args.cpp), loadFiles() with its bern statistics
// (common/trainer.cpp:151-201), train(), write() and the mains' call sequence (transe/bin/trainTransE.cpp:9-20) -- is
// the reference's own code, linked from the objects oracle/Makefile builds out of /root/reference (no source copied).
// oracle/Makefile builds it into oracle/_ref/bin/gpuTrainTrans{E,H,R} (-DBIND_MODEL=0|1|2);
// tests/test_gpu_programs.py runs them next to kb2e_b200/bin/trainTrans* on the same files and compares the outputs.
//
// Initial tables: TransE / TransH let the library initialise on the device (prepTrain() only sizes the reference's
// tables), so that the run is comparable bit for bit with kb2e_b200/bin/trainTrans{E,H}; TransR keeps the reference's own
// prepTrain() (identity matrices + tables seeded from a TransE run's files, transr/trainer.cpp:70-114) and uploads them.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common/args.h"
#include "common/trainer.h"
#include "kb2e_b200.h"
#include "transe/trainer.h"
#include "transh/trainer.h"
#include "transr/trainer.h"

namespace gpu {

typedef std::vector<std::vector<double> > Table;

static void check(int rc, kb2e_ctx* ctx, const char* what) {
   if (rc != KB2E_OK) {
      printf("%s failed: %s\n", what, kb2e_last_error(ctx));
      exit(3);
   }
}

static void upload(kb2e_ctx* c, int table, const Table& t) {
   std::vector<double> flat;
   for (size_t i = 0; i < t.size(); i++) flat.insert(flat.end(), t[i].begin(), t[i].end());
   check(kb2e_upload(c, table, flat.data(), (int64_t)t.size(), (int64_t)t[0].size()), c, "kb2e_upload");
}

static void download(kb2e_ctx* c, int table, Table& t) {
   std::vector<double> flat(t.size() * t[0].size());
   check(kb2e_download(c, table, flat.data(), (int64_t)t.size(), (int64_t)t[0].size()), c, "kb2e_download");
   for (size_t i = 0; i < t.size(); i++) t[i].assign(flat.begin() + i * t[0].size(), flat.begin() + (i + 1) * t[0].size());
}

// The part every model shares: context, training set, bern statistics, epochs, the reference's per-epoch line.
template <class Base>
class Binding : public Base {
 public:
   Binding(common::EmbeddingArguments args, int model) : Base(args), args_(args), model_(model), ctx_(NULL) {}

 protected:
   common::EmbeddingArguments args_;
   int model_;
   kb2e_ctx* ctx_;

   void open() {
      // bit-reproducible accumulation on request (the reference's parser has no such flag: environment); TransE / TransH
      const unsigned flags = (getenv("KB2E_DETERMINISTIC") && model_ != KB2E_MODEL_TRANSR) ? KB2E_FLAG_DETERMINISTIC : 0u;
      kb2e_config cfg = {model_, this->embeddingSize_, this->method_, args_.distanceType, this->numBatches_, /*device*/ 0,
                         this->numEntities_, this->numRelations_, this->learningRate_, this->margin_, args_.seed, flags, 0};
      if (kb2e_create(&cfg, &ctx_) != KB2E_OK) {
         printf("kb2e_create failed: %s\n", kb2e_last_error(NULL));
         exit(3);
      }
      check(kb2e_set_train_triples(ctx_, this->heads_.data(), this->tails_.data(), this->relations_.data(), (int64_t)this->heads_.size()), ctx_,
            "kb2e_set_train_triples");
      std::vector<double> hm(this->numRelations_), tm(this->numRelations_);   // common/trainer.cpp:171-194
      for (int r = 0; r < this->numRelations_; r++) {
         hm[r] = this->relationHeadMeanCooccurrence_[r];
         tm[r] = this->relationTailMeanCooccurrence_[r];
      }
      check(kb2e_set_bern(ctx_, hm.data(), tm.data()), ctx_, "kb2e_set_bern");
   }

   void epochs() {   // replaces the loop of common/trainer.cpp:69-107
      std::vector<double> loss(this->maxEpochs_ > 0 ? this->maxEpochs_ : 1);
      check(kb2e_train_epochs(ctx_, 0, this->maxEpochs_, loss.data()), ctx_, "kb2e_train_epochs");
      for (int e = 0; e < this->maxEpochs_; e++) printf("Epoch: %d, Loss: %f\n", e, loss[e]);
   }

   void sizeTables() {   // what prepTrain() allocates (common/trainer.cpp:34-43), without the rejection-sampled values
      this->relationVec_.assign(this->numRelations_, std::vector<double>(this->embeddingSize_, 0.0));
      this->entityVec_.assign(this->numEntities_, std::vector<double>(this->embeddingSize_, 0.0));
   }
};

class TransETrainer : public Binding<transe::Trainer> {
 public:
   explicit TransETrainer(common::EmbeddingArguments args) : Binding<transe::Trainer>(args, KB2E_MODEL_TRANSE) {}

 protected:
   void prepTrain() override { sizeTables(); }
   void bfgs() override {
      open();
      check(kb2e_init_embeddings(ctx_), ctx_, "kb2e_init_embeddings");
      epochs();
      download(ctx_, KB2E_TABLE_ENTITY, entityVec_);   // write() then emits the usual files (common/trainer.cpp:109-127)
      download(ctx_, KB2E_TABLE_RELATION, relationVec_);
      kb2e_destroy(ctx_);
   }
};

class TransHTrainer : public Binding<transh::Trainer> {
 public:
   explicit TransHTrainer(common::EmbeddingArguments args) : Binding<transh::Trainer>(args, KB2E_MODEL_TRANSH) {}

 protected:
   void prepTrain() override {
      sizeTables();
      weights_.assign(numRelations_, std::vector<double>(embeddingSize_, 0.0));   // transh/trainer.h:17
   }
   void bfgs() override {
      open();
      check(kb2e_init_embeddings(ctx_), ctx_, "kb2e_init_embeddings");
      epochs();
      download(ctx_, KB2E_TABLE_ENTITY, entityVec_);
      download(ctx_, KB2E_TABLE_RELATION, relationVec_);
      download(ctx_, KB2E_TABLE_WEIGHTS, weights_);    // transh::Trainer::write() emits weights.<method> (transh/trainer.cpp:94-105)
      kb2e_destroy(ctx_);
   }
};

class TransRTrainer : public Binding<transr::Trainer> {
 public:
   explicit TransRTrainer(common::EmbeddingArguments args) : Binding<transr::Trainer>(args, KB2E_MODEL_TRANSR) {}

 protected:
   // prepTrain() is the reference's: M_r = I, entity / relation rows from the seed run's files (transr/trainer.cpp:70-114)
   void bfgs() override {
      open();
      const size_t D = (size_t)embeddingSize_;
      upload(ctx_, KB2E_TABLE_ENTITY, entityVec_);
      upload(ctx_, KB2E_TABLE_RELATION, relationVec_);
      std::vector<double> flat((size_t)numRelations_ * D * D);   // M[r][j][i], the order transr/trainer.cpp:128-142 writes
      for (int r = 0; r < numRelations_; r++)
         for (size_t j = 0; j < D; j++)
            for (size_t i = 0; i < D; i++) flat[((size_t)r * D + j) * D + i] = weights_[r][j][i];
      check(kb2e_upload(ctx_, KB2E_TABLE_WEIGHTS, flat.data(), (int64_t)numRelations_ * (int64_t)D, (int64_t)D), ctx_, "kb2e_upload");
      epochs();
      download(ctx_, KB2E_TABLE_ENTITY, entityVec_);
      download(ctx_, KB2E_TABLE_RELATION, relationVec_);
      check(kb2e_download(ctx_, KB2E_TABLE_WEIGHTS, flat.data(), (int64_t)numRelations_ * (int64_t)D, (int64_t)D), ctx_, "kb2e_download");
      for (int r = 0; r < numRelations_; r++)
         for (size_t j = 0; j < D; j++)
            for (size_t i = 0; i < D; i++) weights_[r][j][i] = flat[((size_t)r * D + j) * D + i];
      kb2e_destroy(ctx_);
   }
};

}  // namespace gpu

// The reference's main, with the one changed line (transe/bin/trainTransE.cpp:15).
int main(int argc, char** argv) {
   common::EmbeddingArguments args = common::parseArgs(argc, argv);
   printf("%s\n", args.to_string().c_str());
   srand(args.seed);
#if BIND_MODEL == 0
   common::Trainer* trainer = new gpu::TransETrainer(args);
#elif BIND_MODEL == 1
   common::Trainer* trainer = new gpu::TransHTrainer(args);
#else
   common::Trainer* trainer = new gpu::TransRTrainer(args);
#endif
   trainer->loadFiles();
   trainer->train();
   trainer->write();
   delete (trainer);
   return 0;
}
