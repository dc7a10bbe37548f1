#include "trainer.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <set>
#include <utility>

#include "kb2e_b200.h"
#include "loader.h"

namespace kb2e_host {

Trainer::Trainer(int model, const EmbeddingArguments& args) : model_(model), args_(args) {}

Trainer::~Trainer() {
   if (ctx_) kb2e_destroy(ctx_);
}

void Trainer::die(const char* what) {
   printf("%s failed: %s\n", what, kb2e_last_error(ctx_));
   exit(3);
}

void Trainer::add(int head, int tail, int relation) {
   heads_.push_back(head);
   tails_.push_back(tail);
   relations_.push_back(relation);
}

void Trainer::setCounts(int numEntities, int numRelations) {
   numEntities_ = numEntities;
   numRelations_ = numRelations;
}

// Per relation: (#triples with r) / (#distinct heads of r) and / (#distinct tails of r), 0 for an
// unused relation -- what the reference accumulates in two map-of-maps (common/trainer.cpp:157-194).
void Trainer::computeBernStatistics() {
   headMean_.assign(numRelations_, 0.0);
   tailMean_.assign(numRelations_, 0.0);
   std::vector<double> count(numRelations_, 0.0);
   std::vector<std::pair<int, int>> byHead(heads_.size()), byTail(heads_.size());
   for (size_t i = 0; i < heads_.size(); i++) {
      count[relations_[i]] += 1.0;
      byHead[i] = std::make_pair(relations_[i], heads_[i]);
      byTail[i] = std::make_pair(relations_[i], tails_[i]);
   }
   for (int side = 0; side < 2; side++) {
      std::vector<std::pair<int, int>>& v = side == 0 ? byHead : byTail;
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      std::vector<double> distinct(numRelations_, 0.0);
      for (size_t i = 0; i < v.size(); i++) distinct[v[i].first] += 1.0;
      std::vector<double>& out = side == 0 ? headMean_ : tailMean_;
      for (int r = 0; r < numRelations_; r++) out[r] = distinct[r] > 0 ? count[r] / distinct[r] : 0.0;
   }
}

void Trainer::loadFiles() {
   IdMap entity2id, relation2id;
   if (!loadIdFile(args_.dataDir + "/entity2id.txt", entity2id) || !loadIdFile(args_.dataDir + "/relation2id.txt", relation2id)) {
      printf("Could not read the id files in: %s\n", args_.dataDir.c_str());
      exit(2);
   }
   // the id files fix the table sizes (common/trainer.cpp:196-197); ids must be 0..N-1
   numEntities_ = (int)entity2id.size();
   numRelations_ = (int)relation2id.size();
   const int nE = numEntities_, nR = numRelations_;
   bool ok = loadTripleFile(args_.dataDir + "/train.txt", entity2id, relation2id, [this, nE, nR](int h, int t, int r) {
      if (h < 0 || h >= nE || t < 0 || t >= nE || r < 0 || r >= nR) {
         printf("Triple (%d, %d, %d) has an id outside 0..N-1; ids in the id files must be dense.\n", h, t, r);
         exit(1);
      }
      this->add(h, t, r);
   });
   if (!ok) {
      printf("Could not read: %s/train.txt\n", args_.dataDir.c_str());
      exit(2);
   }
   computeBernStatistics();
   std::cout << "Number of Relations: " << numRelations_ << std::endl;
   std::cout << "Number of Entities: " << numEntities_ << std::endl;
}

void Trainer::prepTrain() {
   kb2e_config cfg;
   cfg.model = model_;
   cfg.dim = args_.embeddingSize;
   cfg.method = args_.method;
   cfg.distance = args_.distanceType;
   cfg.batches = args_.numBatches;
   cfg.device = args_.device;
   cfg.num_entities = numEntities_;
   cfg.num_relations = numRelations_;
   cfg.rate = args_.learningRate;
   cfg.margin = args_.margin;
   cfg.seed = args_.seed;
   cfg.flags = (args_.samplerRandMax ? KB2E_FLAG_SAMPLER_RANDMAX : 0u) | (args_.deterministic ? KB2E_FLAG_DETERMINISTIC : 0u);
   cfg.reserved = 0;
   if (kb2e_create(&cfg, &ctx_) != KB2E_OK) {
      printf("kb2e_create failed: %s\n", kb2e_last_error(NULL));
      exit(3);
   }
   if (kb2e_set_train_triples(ctx_, heads_.data(), tails_.data(), relations_.data(), (int64_t)heads_.size())) die("kb2e_set_train_triples");
   if (kb2e_set_bern(ctx_, headMean_.data(), tailMean_.data())) die("kb2e_set_bern");
   if (kb2e_init_embeddings(ctx_)) die("kb2e_init_embeddings");
   if (args_.resume) {
      loadResumeTables();
      return;
   }
   if (model_ == KB2E_MODEL_TRANSR && args_.seedEpochs > 0) {
      seedTransRInProcess();
      return;
   }
   if (model_ == KB2E_MODEL_TRANSR) {
      // transr/trainer.cpp:88-113: entity and relation tables come from a previous (TransE) run;
      // entity rows are scaled to unit length, relation rows are taken as they are.
      const size_t D = (size_t)args_.embeddingSize;
      std::vector<double> table;
      std::string path = args_.seedDataDir + "/entity2vec." + methodName(args_.seedMethod);
      if (!loadTable(path, (size_t)numEntities_, D, table)) {
         printf("Failed to read embedding values from seed file: '%s'\n", path.c_str());
         exit(1);
      }
      for (int i = 0; i < numEntities_; i++) {
         double len = 0;
         for (size_t j = 0; j < D; j++) len += table[i * D + j] * table[i * D + j];
         len = std::sqrt(len);
         for (size_t j = 0; j < D; j++) table[i * D + j] /= len;
      }
      if (kb2e_upload(ctx_, KB2E_TABLE_ENTITY, table.data(), numEntities_, (int64_t)D)) die("kb2e_upload");
      path = args_.seedDataDir + "/relation2vec." + methodName(args_.seedMethod);
      if (!loadTable(path, (size_t)numRelations_, D, table)) {
         printf("Failed to read embedding values from seed file: '%s'\n", path.c_str());
         exit(1);
      }
      if (kb2e_upload(ctx_, KB2E_TABLE_RELATION, table.data(), numRelations_, (int64_t)D)) die("kb2e_upload");
   }
}

// --resume: the previous run's output files are the initial tables (6-decimal text is the reference's only format).
void Trainer::loadResumeTables() {
   const std::string suffix = std::string(".") + methodName(args_.method);
   const size_t D = (size_t)args_.embeddingSize;
   std::vector<double> table;
   auto restore = [&](const std::string& base, int which, size_t rows) {
      const std::string path = args_.outputDir + "/" + base + suffix;
      if (!loadTable(path, rows, D, table)) {
         printf("Failed to read embedding values from file: '%s'\n", path.c_str());
         exit(1);
      }
      if (kb2e_upload(ctx_, which, table.data(), (int64_t)rows, (int64_t)D)) die("kb2e_upload");
   };
   restore("entity2vec", KB2E_TABLE_ENTITY, (size_t)numEntities_);
   restore("relation2vec", KB2E_TABLE_RELATION, (size_t)numRelations_);
   if (model_ != KB2E_MODEL_TRANSE)
      restore("weights", KB2E_TABLE_WEIGHTS, model_ == KB2E_MODEL_TRANSH ? (size_t)numRelations_ : (size_t)numRelations_ * D);
}

// --seed-epochs N (trainTransR): the seed model -- TransE with --seedmethod on the same training set -- is trained here and
// handed over as doubles (transr/trainer.cpp:88-113 reads it back from 6-decimal text): entity rows scaled to unit
// length, relation rows as they are.
void Trainer::seedTransRInProcess() {
   kb2e_config cfg;
   cfg.model = KB2E_MODEL_TRANSE;
   cfg.dim = args_.embeddingSize;
   cfg.method = args_.seedMethod;
   cfg.distance = args_.distanceType;
   cfg.batches = args_.numBatches;
   cfg.device = args_.device;
   cfg.num_entities = numEntities_;
   cfg.num_relations = numRelations_;
   cfg.rate = args_.learningRate;
   cfg.margin = args_.margin;
   cfg.seed = args_.seed;
   cfg.flags = args_.samplerRandMax ? KB2E_FLAG_SAMPLER_RANDMAX : 0u;
   cfg.reserved = 0;
   kb2e_ctx* seed = nullptr;
   if (kb2e_create(&cfg, &seed) != KB2E_OK) {
      printf("kb2e_create failed: %s\n", kb2e_last_error(NULL));
      exit(3);
   }
   kb2e_ctx* mine = ctx_;
   ctx_ = seed;   // die() reports the seed context's error
   if (kb2e_set_train_triples(seed, heads_.data(), tails_.data(), relations_.data(), (int64_t)heads_.size())) die("kb2e_set_train_triples");
   if (kb2e_set_bern(seed, headMean_.data(), tailMean_.data())) die("kb2e_set_bern");
   if (kb2e_init_embeddings(seed)) die("kb2e_init_embeddings");
   std::vector<double> loss((size_t)args_.seedEpochs);
   if (kb2e_train_epochs(seed, 0, args_.seedEpochs, loss.data())) die("kb2e_train_epochs");
   printf("Seed model (TransE %s): %d epochs, loss %f -> %f\n", methodName(args_.seedMethod), args_.seedEpochs, loss.front(), loss.back());
   const size_t D = (size_t)args_.embeddingSize;
   std::vector<double> ent((size_t)numEntities_ * D), rel((size_t)numRelations_ * D);
   if (kb2e_download(seed, KB2E_TABLE_ENTITY, ent.data(), numEntities_, (int64_t)D)) die("kb2e_download");
   if (kb2e_download(seed, KB2E_TABLE_RELATION, rel.data(), numRelations_, (int64_t)D)) die("kb2e_download");
   kb2e_destroy(seed);
   ctx_ = mine;
   for (int i = 0; i < numEntities_; i++) {
      double len = 0;
      for (size_t j = 0; j < D; j++) len += ent[i * D + j] * ent[i * D + j];
      len = std::sqrt(len);
      for (size_t j = 0; j < D; j++) ent[i * D + j] /= len;
   }
   if (kb2e_upload(ctx_, KB2E_TABLE_ENTITY, ent.data(), numEntities_, (int64_t)D)) die("kb2e_upload");
   if (kb2e_upload(ctx_, KB2E_TABLE_RELATION, rel.data(), numRelations_, (int64_t)D)) die("kb2e_upload");
}

// valid.txt / test.txt for --eval-every / --eval-after; unknown names are reported and skipped as everywhere else
void Trainer::loadEvalSets() {
   if (evalSetsLoaded_) return;
   IdMap entity2id, relation2id;
   if (!loadIdFile(args_.dataDir + "/entity2id.txt", entity2id) || !loadIdFile(args_.dataDir + "/relation2id.txt", relation2id)) {
      printf("Could not read the id files in: %s\n", args_.dataDir.c_str());
      exit(2);
   }
   loadTripleFile(args_.dataDir + "/valid.txt", entity2id, relation2id, [this](int h, int t, int r) {
      validH_.push_back(h); validT_.push_back(t); validR_.push_back(r);
   });
   loadTripleFile(args_.dataDir + "/test.txt", entity2id, relation2id, [this](int h, int t, int r) {
      testH_.push_back(h); testT_.push_back(t); testR_.push_back(r);
   });
   evalSetsLoaded_ = true;
}

// Filtered ranking of valid.txt (which = 0) or test.txt (which = 1) on the context that is training: the tables never
// leave the device.  Filter set = train + valid + test, as in common/evaluation.cpp:59-61.
void Trainer::rankOnDevice(int which, int epoch) {
   loadEvalSets();
   const std::vector<int>& wh = which == 0 ? validH_ : testH_;
   const std::vector<int>& wt = which == 0 ? validT_ : testT_;
   const std::vector<int>& wr = which == 0 ? validR_ : testR_;
   if (wh.empty()) return;
   if (rankingSet_ != which) {
      const std::vector<int>& oh = which == 0 ? testH_ : validH_;
      const std::vector<int>& ot = which == 0 ? testT_ : validT_;
      const std::vector<int>& orr = which == 0 ? testR_ : validR_;
      if (kb2e_set_test_triples(ctx_, wh.data(), wt.data(), wr.data(), (int64_t)wh.size())) die("kb2e_set_test_triples");
      if (kb2e_add_filter_triples(ctx_, NULL, NULL, NULL, 0)) die("kb2e_add_filter_triples");
      if (kb2e_add_filter_triples(ctx_, heads_.data(), tails_.data(), relations_.data(), (int64_t)heads_.size())) die("kb2e_add_filter_triples");
      if (!oh.empty() && kb2e_add_filter_triples(ctx_, oh.data(), ot.data(), orr.data(), (int64_t)oh.size())) die("kb2e_add_filter_triples");
      rankingSet_ = which;
   }
   int64_t sums[4] = {0, 0, 0, 0};
   if (kb2e_rank(ctx_, 0, (int64_t)wh.size(), NULL, NULL, NULL, NULL, sums)) die("kb2e_rank");
   const double n2 = wh.size() * 2.0;
   if (which == 0) {
      printf("Valid @ epoch %d -- Raw Rank: %f, Hits@10: %f; Filtered Rank: %f, Hits@10: %f\n", epoch, sums[0] / n2, sums[2] / n2,
             sums[1] / n2, sums[3] / n2);
   } else {
      printf("Raw      -- Rank: %f, Hits@10: %f\n", sums[0] / n2, sums[2] / n2);
      printf("Filtered -- Rank: %f, Hits@10: %f\n", sums[1] / n2, sums[3] / n2);
   }
   fflush(stdout);
}

void Trainer::bfgs() {
   // Every batch of every epoch runs inside persistent device launches; the host only prints the
   // reference's per-epoch line (common/trainer.cpp:105).  Epochs go down in chunks so that output
   // appears while a long run is in flight; a chunk also ends where a validation pass or a checkpoint is due.
   epochLoss_.assign((size_t)std::max(0, args_.maxEpochs), 0.0);
   const int chunk = 50;
   const int base = std::max(0, args_.firstEpoch);   // a resumed run continues the RNG stream and the numbering
   auto due = [](int every, int done) { return every > 0 && done % every == 0; };
   for (int first = 0; first < args_.maxEpochs;) {
      int n = std::min(chunk, args_.maxEpochs - first);
      for (int every : {args_.evalEvery, args_.checkpointEvery})
         if (every > 0) n = std::min(n, every - first % every);
      if (kb2e_train_epochs(ctx_, base + first, n, epochLoss_.data() + first)) die("kb2e_train_epochs");
      for (int e = first; e < first + n; e++) printf("Epoch: %d, Loss: %f\n", base + e, epochLoss_[e]);
      fflush(stdout);
      first += n;
      if (due(args_.evalEvery, first)) rankOnDevice(0, base + first);
      if (due(args_.checkpointEvery, first) && first < args_.maxEpochs) write();
   }
   if (args_.evalAfter) rankOnDevice(1, base + args_.maxEpochs);
}

void Trainer::train() {
   prepTrain();
   bfgs();
}

void Trainer::write() {
   const std::string suffix = std::string(".") + methodName(args_.method);
   const int64_t D = args_.embeddingSize;
   // every table is checked: a failed or short write (disk full, permissions) must not exit 0 with a partial checkpoint
   auto save = [](const std::string& path, size_t rows, size_t cols, const double* data) {
      if (!writeTable(path, rows, cols, data)) {
         printf("Could not write to: %s\n", path.c_str());
         exit(2);
      }
   };
   std::vector<double> table((size_t)numRelations_ * D);
   if (kb2e_download(ctx_, KB2E_TABLE_RELATION, table.data(), numRelations_, D)) die("kb2e_download");
   save(args_.outputDir + "/relation2vec" + suffix, (size_t)numRelations_, (size_t)D, table.data());
   table.resize((size_t)numEntities_ * D);
   if (kb2e_download(ctx_, KB2E_TABLE_ENTITY, table.data(), numEntities_, D)) die("kb2e_download");
   save(args_.outputDir + "/entity2vec" + suffix, (size_t)numEntities_, (size_t)D, table.data());
   if (model_ != KB2E_MODEL_TRANSE) {
      // transh/trainer.cpp:94-105: one row per relation; transr/trainer.cpp:128-142: D rows per relation
      const int64_t rows = model_ == KB2E_MODEL_TRANSH ? numRelations_ : (int64_t)numRelations_ * D;
      table.resize((size_t)rows * D);
      if (kb2e_download(ctx_, KB2E_TABLE_WEIGHTS, table.data(), rows, D)) die("kb2e_download");
      save(args_.outputDir + "/weights" + suffix, (size_t)rows, (size_t)D, table.data());
   }
}

}  // namespace kb2e_host
