// Text loaders for the reference's input files (common/loader.h:11-19, common/loader.cpp:15-62) and
// for embedding tables ("%.6lf\t" cells, common/trainer.cpp:109-127, common/evaluation.cpp:74-105).
#ifndef KB2E_HOST_LOADER_H_
#define KB2E_HOST_LOADER_H_

#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

namespace kb2e_host {

typedef std::unordered_map<std::string, int> IdMap;

// "name<ws>id" records; a later duplicate name overwrites an earlier one.  Returns false when the
// file cannot be opened (the reference dereferences NULL there).
bool loadIdFile(const std::string& path, IdMap& idMap);

// "head<ws>tail<ws>relation" NAME records; triples naming an unknown entity / relation are reported
// on stdout with the reference's messages and skipped; callback(head, tail, relation) in file order.
bool loadTripleFile(const std::string& path, const IdMap& entityIdMap, const IdMap& relationIdMap,
                    const std::function<void(int, int, int)>& callback);

// rows x cols doubles separated by white space.  Returns false if the file is missing or short.
bool loadTable(const std::string& path, size_t rows, size_t cols, std::vector<double>& out);
bool writeTable(const std::string& path, size_t rows, size_t cols, const double* data);
bool fileExists(const std::string& path);

}  // namespace kb2e_host

#endif  // KB2E_HOST_LOADER_H_
