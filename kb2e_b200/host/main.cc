// One source for the six reference-named programs (Makefile:28,33-49 of the reference):
//   -DKB2E_MAIN_MODEL={0 TransE, 1 TransH, 2 TransR}  -DKB2E_MAIN_EVAL={0 train, 1 eval}
// Same call sequence as transe/bin/trainTransE.cpp:9-20 and transe/bin/evalTransE.cpp:9-19:
// parse, print the options banner, then loadFiles/train/write or prepare/run.  srand() is not
// needed: all randomness comes from the counter RNG keyed by -seed inside the library.
#include <cstdio>

#include "args.h"
#include "evaluation.h"
#include "trainer.h"

int main(int argc, char** argv) {
   kb2e_host::EmbeddingArguments args = kb2e_host::parseArgs(argc, argv);
   printf("%s\n", args.to_string().c_str());
#if KB2E_MAIN_EVAL
   kb2e_host::EmbeddingEvaluation evaluation(KB2E_MAIN_MODEL, args);
   evaluation.prepare();
   evaluation.run();
#else
   kb2e_host::Trainer trainer(KB2E_MAIN_MODEL, args);
   trainer.loadFiles();
   trainer.train();
   trainer.write();
#endif
   return 0;
}
