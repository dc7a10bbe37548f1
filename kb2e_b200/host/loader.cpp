#include "loader.h"

#include <sys/stat.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace kb2e_host {

namespace {

// Whole file in memory; the parsers below walk it with pointers (the reference's fscanf loops are
// what dominates wall time once the epochs run on the GPU).
bool slurp(const std::string& path, std::string& out) {
   FILE* f = fopen(path.c_str(), "rb");
   if (!f) return false;
   fseek(f, 0, SEEK_END);
   long size = ftell(f);
   fseek(f, 0, SEEK_SET);
   out.resize(size > 0 ? (size_t)size : 0);
   size_t got = size > 0 ? fread(&out[0], 1, (size_t)size, f) : 0;
   fclose(f);
   out.resize(got);
   return true;
}

inline bool isSpace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// Next white-space delimited token in [p, end); returns false at end of input.
inline bool nextToken(const char*& p, const char* end, const char*& tok, size_t& len) {
   while (p < end && isSpace(*p)) p++;
   if (p >= end) return false;
   tok = p;
   while (p < end && !isSpace(*p)) p++;
   len = (size_t)(p - tok);
   return true;
}

}  // namespace

bool loadIdFile(const std::string& path, IdMap& idMap) {
   std::string text;
   if (!slurp(path, text)) return false;
   const char* p = text.data();
   const char* end = p + text.size();
   const char *name, *num;
   size_t nameLen, numLen;
   // fscanf("%s\t%d") stops at the first record that does not parse (common/loader.cpp:20)
   while (nextToken(p, end, name, nameLen) && nextToken(p, end, num, numLen)) {
      char* stop = NULL;
      std::string digits(num, numLen);
      long id = strtol(digits.c_str(), &stop, 10);
      if (stop == digits.c_str()) break;
      idMap[std::string(name, nameLen)] = (int)id;
   }
   return true;
}

bool loadTripleFile(const std::string& path, const IdMap& entityIdMap, const IdMap& relationIdMap,
                    const std::function<void(int, int, int)>& callback) {
   std::string text;
   if (!slurp(path, text)) return false;
   const char* p = text.data();
   const char* end = p + text.size();
   const char* tok[3];
   size_t len[3];
   while (nextToken(p, end, tok[0], len[0]) && nextToken(p, end, tok[1], len[1]) && nextToken(p, end, tok[2], len[2])) {
      std::string head(tok[0], len[0]), tail(tok[1], len[1]), relation(tok[2], len[2]);
      IdMap::const_iterator h = entityIdMap.find(head);
      IdMap::const_iterator t = entityIdMap.find(tail);
      IdMap::const_iterator r = relationIdMap.find(relation);
      bool ok = true;
      if (h == entityIdMap.end()) {
         std::cout << "Head entity found in triple file that was not found in the identity file: " << head << std::endl;
         ok = false;
      }
      if (t == entityIdMap.end()) {
         std::cout << "Tail entity found in triple file that was not found in the identity file: " << tail << std::endl;
         ok = false;
      }
      if (r == relationIdMap.end()) {
         std::cout << "Relation found in triple file that was not found in the identity file: " << relation << std::endl;
         ok = false;
      }
      if (ok) callback(h->second, t->second, r->second);
   }
   return true;
}

// Embedding tables are the bulk of the text the programs read and write (FB15k shape, size 100: 1.6 M numbers per
// file).  std::from_chars / std::to_chars are correctly rounded like strtod / printf -- the same doubles in, the same
// bytes out ("%.6lf\t" cells, common/trainer.cpp:113) -- at a fraction of the cost; anything they do not take
// (hex floats, inf / nan spellings, values too long for the cell buffer) goes through the C library as before.
bool loadTable(const std::string& path, size_t rows, size_t cols, std::vector<double>& out) {
   std::string text;
   if (!slurp(path, text)) return false;
   out.resize(rows * cols);
   const char* p = text.c_str();
   const char* end = p + text.size();
   for (size_t i = 0; i < rows * cols; i++) {
      while (p < end && isSpace(*p)) p++;
      const char* q = (p < end && *p == '+') ? p + 1 : p;   // strtod takes a leading '+', from_chars does not
      double v = 0.0;
      std::from_chars_result r = std::from_chars(q, end, v);
      if (r.ec == std::errc() && r.ptr != q && !(r.ptr < end && (*r.ptr == 'x' || *r.ptr == 'X'))) {
         out[i] = v;
         p = r.ptr;
         continue;
      }
      char* stop = NULL;
      out[i] = strtod(p, &stop);  // same conversion fscanf("%lf") performs
      if (stop == p) return false;
      p = stop;
   }
   return true;
}

bool writeTable(const std::string& path, size_t rows, size_t cols, const double* data) {
   FILE* f = fopen(path.c_str(), "w");
   if (!f) return false;
   std::string line;
   char cell[400];
   for (size_t i = 0; i < rows; i++) {
      line.clear();
      for (size_t j = 0; j < cols; j++) {
         const double v = data[i * cols + j];
         std::to_chars_result r{cell, std::errc::value_too_large};
         if (std::isfinite(v)) r = std::to_chars(cell, cell + sizeof(cell), v, std::chars_format::fixed, 6);
         if (r.ec == std::errc()) {
            line.append(cell, (size_t)(r.ptr - cell));
         } else {
            int n = snprintf(cell, sizeof(cell), "%.6lf", v);
            line.append(cell, (size_t)std::min<int>(n, (int)sizeof(cell) - 1));
         }
         line.push_back('\t');
      }
      line.push_back('\n');
      if (fwrite(line.data(), 1, line.size(), f) != line.size()) {   // disk full, quota, ...
         fclose(f);
         return false;
      }
   }
   return fclose(f) == 0;
}

bool fileExists(const std::string& path) {
   struct stat info;
   return stat(path.c_str(), &info) == 0;
}

}  // namespace kb2e_host
