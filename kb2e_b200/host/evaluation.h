// Host driver of an evaluation run: the public surface of common::EmbeddingEvaluation
// (common/evaluation.h:15-21: prepare, run); the model plug-ins' tripleEnergy / loadEmbeddings
// overrides (evaluation.h:60-61) become a `model` field because scoring runs on the GPU.
#ifndef KB2E_HOST_EVALUATION_H_
#define KB2E_HOST_EVALUATION_H_

#include <string>
#include <vector>

#include "args.h"

struct kb2e_ctx;

namespace kb2e_host {

struct EvaluationResult {
   double rawMeanRank = 0, rawHitsAt10 = 0, filteredMeanRank = 0, filteredHitsAt10 = 0;
   long long queries = 0;
};

class EmbeddingEvaluation {
   public:
      EmbeddingEvaluation(int model, const EmbeddingArguments& args);
      ~EmbeddingEvaluation();

      void prepare();  // common/evaluation.cpp:253-266
      void run();      // common/evaluation.cpp:181-251

      const EvaluationResult& result() const { return result_; }

   private:
      int model_;
      EmbeddingArguments args_;
      int numEntities_ = 0, numRelations_ = 0;
      std::string relationEmbeddingPath_, entityEmbeddingPath_, weightEmbeddingPath_;
      std::vector<int> heads_, tails_, relations_;              // working set = test.txt, in file order
      std::vector<int> filterHeads_, filterTails_, filterRelations_;  // train.txt + valid.txt
      // one context per GPU (--gpus N: devices device .. device + N - 1); the test triples are sharded over them in
      // contiguous windows, tables and filter set are replicated (common/evaluation.cpp:221-241 shares nothing between
      // test triples); ctx_ is contexts_[0]
      std::vector<kb2e_ctx*> contexts_;
      kb2e_ctx* ctx_ = nullptr;
      EvaluationResult result_;

      void loadTriples();     // common/evaluation.cpp:41-62
      void loadEmbeddings();  // common/evaluation.cpp:74-105, transh/evaluation.cpp:20-40, transr/evaluation.cpp:34-60
      void uploadTable(int table, const double* data, long long rows, long long cols);
      void die(const char* what);
};

}  // namespace kb2e_host

#endif  // KB2E_HOST_EVALUATION_H_
