// Command-line options of the six programs: the reference's flag set, defaults, banner and usage
// text (common/args.h:9-28, common/args.cpp:18-142, common/constants.h:28-54), plus GPU-only
// options that default to the reference's behaviour.
#ifndef KB2E_HOST_ARGS_H_
#define KB2E_HOST_ARGS_H_

#include <string>

namespace kb2e_host {

const int kMethodUnif = 0;  // common/constants.h:8-9
const int kMethodBern = 1;
const int kDistanceL1 = 0;  // common/constants.h:16-17
const int kDistanceL2 = 1;

inline const char* methodName(int method) { return method == kMethodUnif ? "unif" : "bern"; }

struct EmbeddingArguments {
   std::string dataDir = "../data";
   std::string outputDir = ".";
   int embeddingSize = 100;
   double learningRate = 0.001;
   double margin = 1.0;
   int method = kMethodBern;
   int numBatches = 100;
   int maxEpochs = 1000;
   int distanceType = kDistanceL1;
   std::string seedDataDir = ".";
   int seedMethod = kMethodUnif;
   unsigned int seed;
   // GPU-only (not in the reference; never printed in the Options banner)
   int device = 0;
   int gpus = 1;   // eval programs: shard the test triples over this many GPUs (devices device .. device + gpus - 1)
   // sampler indices: 0 = uniform (default), 1 = the index distribution of the reference's randMax
   // (common/utils.cpp:113-120), for trained-model parity with the shipped reference
   int samplerRandMax = 0;
   int deterministic = 0;   // 1: bit-reproducible training (KB2E_FLAG_DETERMINISTIC; TransE, TransH)
   // train programs, SURVEY.md 8f rows 2-4 (none of these exists in the reference; all default to its behaviour)
   int evalEvery = 0;       // > 0: filtered MeanRank / Hits@10 of valid.txt every that many epochs, on the device-resident tables
   int evalAfter = 0;       // 1: rank test.txt after training in the same process (the lines evalTrans* prints), no text round trip
   int resume = 0;          // 1: start from the tables in --outdir (a previous run's output) instead of a fresh initialisation
   int firstEpoch = 0;      // epoch number the run starts at (continues the counter-RNG stream and the "Epoch:" numbering of a resumed run)
   int checkpointEvery = 0; // > 0: write the output files every that many epochs as well
   int seedEpochs = 0;      // TransR: > 0 trains the TransE (--seedmethod) seed model for that many epochs in this process instead of reading seed files

   EmbeddingArguments();
   std::string to_string() const;  // the "Options: [...]" banner, byte-compatible with the reference
};

// Exits like the reference: --help prints the usage and exits 0; a flag without its value prints
// "Argument missing for <flag>" and exits 1; unknown flags are ignored.
EmbeddingArguments parseArgs(int argc, char** argv);
void printUsage(const char* invokedFile);

}  // namespace kb2e_host

#endif  // KB2E_HOST_ARGS_H_
