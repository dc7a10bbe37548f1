// Host driver of a training run: same public surface as the reference's common::Trainer
// (common/trainer.h:14-22: add, loadFiles, train, write) with the per-model plug-ins
// (transe|transh|transr::Trainer) folded into a `model` field, because the virtuals they override
// (bfgs, prebatch, postbatch, gradientUpdate, tripleEnergy, initialEmbeddingValue; trainer.h:59-77)
// all run on the GPU behind the C ABI (include/kb2e_b200.h).
#ifndef KB2E_HOST_TRAINER_H_
#define KB2E_HOST_TRAINER_H_

#include <string>
#include <vector>

#include "args.h"

struct kb2e_ctx;

namespace kb2e_host {

class Trainer {
   public:
      Trainer(int model, const EmbeddingArguments& args);
      ~Trainer();

      void add(int head, int tail, int relation);  // common/trainer.cpp:26-32
      void loadFiles();                            // common/trainer.cpp:151-201
      void train();                                // common/trainer.cpp:60-63: prepTrain(); bfgs();
      void write();                                // common/trainer.cpp:109-127 (+ weights file)

      // In-memory alternative to loadFiles() for embedding the trainer in another program.
      void setCounts(int numEntities, int numRelations);
      const std::vector<double>& epochLoss() const { return epochLoss_; }

   private:
      int model_;
      EmbeddingArguments args_;
      int numEntities_ = 0, numRelations_ = 0;
      std::vector<int> heads_, tails_, relations_;
      std::vector<double> headMean_, tailMean_;  // relation{Head,Tail}MeanCooccurrence_ (trainer.h:53-55)
      std::vector<double> epochLoss_;
      kb2e_ctx* ctx_ = nullptr;
      // --eval-every / --eval-after: valid.txt and test.txt (loaded on demand), ranked on the context that trains
      std::vector<int> validH_, validT_, validR_, testH_, testT_, testR_;
      bool evalSetsLoaded_ = false;
      int rankingSet_ = -1;   // which working set the context currently holds: 0 valid, 1 test

      void loadEvalSets();
      void rankOnDevice(int which, int epoch);     // prints the evaluation lines (common/evaluation.cpp:247-250)
      void seedTransRInProcess();                  // --seed-epochs: transr/trainer.cpp:88-113 without the text round trip
      void loadResumeTables();                     // --resume

      void computeBernStatistics();  // common/trainer.cpp:171-194
      void prepTrain();              // common/trainer.cpp:34-58, transh/trainer.cpp:77-88, transr/trainer.cpp:70-114
      void bfgs();                   // common/trainer.cpp:69-107
      void die(const char* what);
};

}  // namespace kb2e_host

#endif  // KB2E_HOST_TRAINER_H_
