// Test driver for kb2e_b200/host/loader.cpp (CPU only): doubles (binary) -> writeTable -> text -> loadTable -> doubles (binary).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "loader.h"

int main(int argc, char** argv) {
   if (argc != 6) return 2;
   const size_t rows = (size_t)atol(argv[1]), cols = (size_t)atol(argv[2]);
   std::vector<double> in(rows * cols);
   FILE* f = fopen(argv[3], "rb");
   if (!f || fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 3;
   fclose(f);
   if (!kb2e_host::writeTable(argv[4], rows, cols, in.data())) return 4;
   std::vector<double> back;
   if (!kb2e_host::loadTable(argv[4], rows, cols, back)) return 5;
   f = fopen(argv[5], "wb");
   fwrite(back.data(), sizeof(double), back.size(), f);
   fclose(f);
   return 0;
}
