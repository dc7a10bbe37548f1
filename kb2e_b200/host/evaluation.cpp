#include "evaluation.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <thread>

#include "kb2e_b200.h"
#include "loader.h"

namespace kb2e_host {

EmbeddingEvaluation::EmbeddingEvaluation(int model, const EmbeddingArguments& args) : model_(model), args_(args) {
   const std::string suffix = std::string(".") + methodName(args.method);
   relationEmbeddingPath_ = args.outputDir + "/relation2vec" + suffix;
   entityEmbeddingPath_ = args.outputDir + "/entity2vec" + suffix;
   weightEmbeddingPath_ = args.outputDir + "/weights" + suffix;
}

EmbeddingEvaluation::~EmbeddingEvaluation() {
   for (kb2e_ctx* c : contexts_) kb2e_destroy(c);
}

void EmbeddingEvaluation::uploadTable(int table, const double* data, long long rows, long long cols) {
   for (kb2e_ctx* c : contexts_) {
      if (kb2e_upload(c, table, data, rows, cols)) {
         ctx_ = c;
         die("kb2e_upload");
      }
   }
}

void EmbeddingEvaluation::die(const char* what) {
   printf("%s failed: %s\n", what, kb2e_last_error(ctx_));
   exit(3);
}

void EmbeddingEvaluation::loadTriples() {
   IdMap entity2id, relation2id;
   if (!loadIdFile(args_.dataDir + "/entity2id.txt", entity2id) || !loadIdFile(args_.dataDir + "/relation2id.txt", relation2id)) {
      printf("Could not read the id files in: %s\n", args_.dataDir.c_str());
      exit(2);
   }
   numEntities_ = (int)entity2id.size();
   numRelations_ = (int)relation2id.size();
   const int nE = numEntities_, nR = numRelations_;
   auto check = [nE, nR](int h, int t, int r) {
      if (h < 0 || h >= nE || t < 0 || t >= nE || r < 0 || r >= nR) {
         printf("Triple (%d, %d, %d) has an id outside 0..N-1; ids in the id files must be dense.\n", h, t, r);
         exit(1);
      }
   };
   // test = working set + filter; train and valid = filter only (common/evaluation.cpp:59-61)
   loadTripleFile(args_.dataDir + "/test.txt", entity2id, relation2id, [&](int h, int t, int r) {
      check(h, t, r);
      heads_.push_back(h); tails_.push_back(t); relations_.push_back(r);
   });
   auto known = [&](int h, int t, int r) {
      check(h, t, r);
      filterHeads_.push_back(h); filterTails_.push_back(t); filterRelations_.push_back(r);
   };
   loadTripleFile(args_.dataDir + "/train.txt", entity2id, relation2id, known);
   loadTripleFile(args_.dataDir + "/valid.txt", entity2id, relation2id, known);
}

void EmbeddingEvaluation::loadEmbeddings() {
   const size_t D = (size_t)args_.embeddingSize;
   if (model_ != KB2E_MODEL_TRANSE && !fileExists(weightEmbeddingPath_)) {
      printf("Could not find weight embedding file: %s. Make sure to specify the path and/or train.\n", weightEmbeddingPath_.c_str());
      exit(2);
   }
   std::vector<double> table;
   if (!loadTable(relationEmbeddingPath_, (size_t)numRelations_, D, table)) {
      printf("Failed to read embedding values from file: '%s'\n", relationEmbeddingPath_.c_str());
      exit(1);
   }
   uploadTable(KB2E_TABLE_RELATION, table.data(), numRelations_, (long long)D);
   if (!loadTable(entityEmbeddingPath_, (size_t)numEntities_, D, table)) {
      printf("Failed to read embedding values from file: '%s'\n", entityEmbeddingPath_.c_str());
      exit(1);
   }
   for (int i = 0; i < numEntities_; i++) {
      // the reference's informational length check (common/evaluation.cpp:100-102)
      double len = 0;
      for (size_t j = 0; j < D; j++) len += table[i * D + j] * table[i * D + j];
      len = std::sqrt(len);
      if (len - 1 > 1e-3) std::cout << "wrong_entity" << i << ' ' << len << std::endl;
   }
   uploadTable(KB2E_TABLE_ENTITY, table.data(), numEntities_, (long long)D);
   if (model_ != KB2E_MODEL_TRANSE) {
      const size_t rows = model_ == KB2E_MODEL_TRANSH ? (size_t)numRelations_ : (size_t)numRelations_ * D;
      if (!loadTable(weightEmbeddingPath_, rows, D, table)) {
         printf("Failed to read embedding weight values from seed file: '%s'\n", weightEmbeddingPath_.c_str());
         exit(1);
      }
      uploadTable(KB2E_TABLE_WEIGHTS, table.data(), (long long)rows, (long long)D);
   }
}

void EmbeddingEvaluation::prepare() {
   if (!fileExists(relationEmbeddingPath_)) {
      printf("Could not find relation embedding file: %s. Make sure to specify the path and/or train.\n", relationEmbeddingPath_.c_str());
      exit(2);
   }
   if (!fileExists(entityEmbeddingPath_)) {
      printf("Could not find entity embedding file: %s. Make sure to specify the path and/or train.\n", entityEmbeddingPath_.c_str());
      exit(2);
   }
   loadTriples();
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
   cfg.flags = 0;
   cfg.reserved = 0;
   const int gpus = std::max(1, args_.gpus);
   for (int g = 0; g < gpus; g++) {
      cfg.device = args_.device + g;
      kb2e_ctx* c = nullptr;
      if (kb2e_create(&cfg, &c) != KB2E_OK) {
         printf("kb2e_create failed: %s\n", kb2e_last_error(NULL));
         exit(3);
      }
      contexts_.push_back(c);
   }
   ctx_ = contexts_[0];
   loadEmbeddings();
   for (kb2e_ctx* c : contexts_) {
      ctx_ = c;
      if (kb2e_set_test_triples(c, heads_.data(), tails_.data(), relations_.data(), (int64_t)heads_.size())) die("kb2e_set_test_triples");
      if (!filterHeads_.empty() &&
          kb2e_add_filter_triples(c, filterHeads_.data(), filterTails_.data(), filterRelations_.data(), (int64_t)filterHeads_.size()))
         die("kb2e_add_filter_triples");
   }
   ctx_ = contexts_[0];
}

void EmbeddingEvaluation::run() {
   // every GPU ranks a contiguous window of the test triples on its own host thread; the four sums add up
   // (common/evaluation.cpp:169-178 accumulates them in exactly this way)
   const size_t G = contexts_.size();
   const int64_t n = (int64_t)heads_.size();
   std::vector<int64_t> part(4 * G, 0);
   std::vector<int> status(G, 0);
   std::vector<std::thread> workers;
   for (size_t g = 0; g < G; g++) {
      const int64_t lo = n * (int64_t)g / (int64_t)G, hi = n * (int64_t)(g + 1) / (int64_t)G;
      workers.emplace_back([this, g, lo, hi, &part, &status]() {
         status[g] = kb2e_rank(contexts_[g], lo, hi - lo, NULL, NULL, NULL, NULL, part.data() + 4 * g);
      });
   }
   for (std::thread& w : workers) w.join();
   int64_t sums[4] = {0, 0, 0, 0};
   for (size_t g = 0; g < G; g++) {
      if (status[g]) {
         ctx_ = contexts_[g];
         die("kb2e_rank");
      }
      for (int k = 0; k < 4; k++) sums[k] += part[4 * g + k];
   }
   // the reference prints a progress line per relation (common/evaluation.cpp:243); all relations
   // are ranked in one pass here, so only the final state of that line is shown
   printf("\rProcessed %05.2f%% ...", 100.0);
   printf("\n");
   double numberCorruptions = heads_.size() * 2.0;
   result_.queries = (long long)(heads_.size() * 2);
   result_.rawMeanRank = sums[0] / numberCorruptions;
   result_.filteredMeanRank = sums[1] / numberCorruptions;
   result_.rawHitsAt10 = sums[2] / numberCorruptions;
   result_.filteredHitsAt10 = sums[3] / numberCorruptions;
   printf("Raw      -- Rank: %f, Hits@10: %f\n", result_.rawMeanRank, result_.rawHitsAt10);
   printf("Filtered -- Rank: %f, Hits@10: %f\n", result_.filteredMeanRank, result_.filteredHitsAt10);
}

}  // namespace kb2e_host
