#include "args.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <vector>

namespace kb2e_host {

namespace {

// Index of "-name" / "--name" in argv (first match, scanning from 1), or -1.  When the option
// needs a value and is the last word, complain and exit(1) as common/utils.cpp:55-68 does.
int findFlag(const char* name, bool needsValue, int argc, char** argv) {
   for (int i = 1; i < argc; i++) {
      const char* a = argv[i];
      if (a[0] != '-') continue;
      const char* body = (a[1] == '-') ? a + 2 : a + 1;
      if (std::strcmp(body, name) != 0) continue;
      if (needsValue && i + 1 >= argc) {
         printf("Argument missing for %s\n", name);
         exit(1);
      }
      return i;
   }
   return -1;
}

struct Option {
   const char* name;
   std::function<void(const char*)> set;
};

}  // namespace

EmbeddingArguments::EmbeddingArguments() { seed = (unsigned int)time(NULL); }

std::string EmbeddingArguments::to_string() const {
   std::string s = "Options: [";
   s += "datadir: '" + dataDir + "', ";
   s += "outdir: '" + outputDir + "', ";
   s += "size: " + std::to_string(embeddingSize) + ", ";
   s += "rate: " + std::to_string(learningRate) + ", ";
   s += "margin: " + std::to_string(margin) + ", ";
   s += std::string("method: ") + methodName(method) + ", ";
   s += "batches: " + std::to_string(numBatches) + ", ";
   s += "epochs: " + std::to_string(maxEpochs) + ", ";
   s += "distance: " + std::to_string(distanceType) + ", ";
   s += "seeddatadir: '" + seedDataDir + "', ";
   s += std::string("seedmethod: ") + methodName(seedMethod) + ", ";
   s += "seed: " + std::to_string(seed) + "]";
   return s;
}

EmbeddingArguments parseArgs(int argc, char** argv) {
   if (findFlag("help", false, argc, argv) != -1) {
      printUsage(argv[0]);
      exit(0);
   }
   EmbeddingArguments a;
   const std::vector<Option> options = {
      {"datadir", [&](const char* v) { a.dataDir = v; }},
      {"outdir", [&](const char* v) { a.outputDir = v; }},
      {"size", [&](const char* v) { a.embeddingSize = atoi(v); }},
      {"rate", [&](const char* v) { a.learningRate = atof(v); }},
      {"margin", [&](const char* v) { a.margin = atof(v); }},
      {"method", [&](const char* v) { a.method = atoi(v); }},
      {"batches", [&](const char* v) { a.numBatches = atoi(v); }},
      {"epochs", [&](const char* v) { a.maxEpochs = atoi(v); }},
      {"distance", [&](const char* v) { a.distanceType = atoi(v); }},
      {"seeddatadir", [&](const char* v) { a.seedDataDir = v; }},
      {"seedmethod", [&](const char* v) { a.seedMethod = atoi(v); }},
      {"seed", [&](const char* v) { a.seed = (unsigned int)atoi(v); }},
      {"device", [&](const char* v) { a.device = atoi(v); }},
      {"gpus", [&](const char* v) { a.gpus = atoi(v); }},
      {"deterministic", [&](const char* v) { a.deterministic = atoi(v); }},
      {"eval-every", [&](const char* v) { a.evalEvery = atoi(v); }},
      {"eval-after", [&](const char* v) { a.evalAfter = atoi(v); }},
      {"resume", [&](const char* v) { a.resume = atoi(v); }},
      {"first-epoch", [&](const char* v) { a.firstEpoch = atoi(v); }},
      {"checkpoint-every", [&](const char* v) { a.checkpointEvery = atoi(v); }},
      {"seed-epochs", [&](const char* v) { a.seedEpochs = atoi(v); }},
      {"sampler", [&](const char* v) { a.samplerRandMax = (std::strcmp(v, "reference") == 0 || std::strcmp(v, "randmax") == 0 || atoi(v) == 1); }},
   };
   for (const Option& o : options) {
      int at = findFlag(o.name, true, argc, argv);
      if (at != -1) o.set(argv[at + 1]);
   }
   return a;
}

void printUsage(const char* invokedFile) {
   printf("USAGE: %s [option value] ...\n", invokedFile);
   printf("       %s --help\n", invokedFile);
   printf("All options require a value.\n");
   printf("Options:\n");
   printf("   --%s [%s]\n", "datadir", "../data");
   printf("   --%s [%s]\n", "outdir", ".");
   printf("   --%s [%d]\n", "size", 100);
   printf("   --%s [%f]\n", "rate", 0.001);
   printf("   --%s [%f]\n", "margin", 1.0);
   printf("   --%s [%d (%s)]\n", "method", kMethodBern, methodName(kMethodBern));
   printf("   --%s [%d]\n", "batches", 100);
   printf("   --%s [%d]\n", "epochs", 1000);
   printf("   --%s [%d]\n", "distance", kDistanceL1);
   printf("   --%s [%s] (TransR only)\n", "seeddatadir", ".");
   printf("   --%s [%d (%s)] (TransR only)\n", "seedmethod", kMethodUnif, methodName(kMethodUnif));
   printf("   --%s [now]\n", "seed");
   printf("   --%s [0] (B200 build only: CUDA device ordinal)\n", "device");
   printf("   --%s [1] (B200 build only, eval programs: shard the test triples over this many GPUs)\n", "gpus");
   printf("   --%s [0] (B200 build only, TransE / TransH training: 1 = bit-reproducible runs, fixed-point accumulation)\n", "deterministic");
   printf("   --%s [0] (B200 build only, train programs: filtered rank of valid.txt every N epochs, tables stay on the device)\n", "eval-every");
   printf("   --%s [0] (B200 build only, train programs: 1 = rank test.txt after training in the same process)\n", "eval-after");
   printf("   --%s [0] (B200 build only, train programs: 1 = start from the tables in --outdir; see --first-epoch)\n", "resume");
   printf("   --%s [0] (B200 build only, train programs: epoch number the run starts at)\n", "first-epoch");
   printf("   --%s [0] (B200 build only, train programs: write the output files every N epochs as well)\n", "checkpoint-every");
   printf("   --%s [0] (B200 build only, trainTransR: train the TransE seed model for N epochs in this process instead of reading seed files)\n", "seed-epochs");
   printf("   --%s [uniform] (B200 build only: 'reference' draws indices with the distribution of the reference's randMax)\n", "sampler");
}

}  // namespace kb2e_host
