// Tooling, not product code: drives the UNMODIFIED reference library through its public C++ API
// (MV_Init / MV_CreateTable<MatrixTableOption<float>> / MatrixWorkerTable::Add / Get, the same calls as
// Test/test_matrix_perf.cpp:32-171 and Test/test_matrix_table.cpp:9-99) on BASELINE.json config 2:
// a rows x cols fp32 MatrixTable, whole-table Add (server-side sgd updater) and whole-table Get, timed
// with the reference's own Timer, max over ranks taken by the caller.  One JSON line per rank.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <multiverso/multiverso.h>
#include <multiverso/table/matrix_table.h>
#include <multiverso/updater/updater.h>
#include <multiverso/util/configure.h>
#include <multiverso/util/timer.h>

int main(int argc, char* argv[]) {
  using namespace multiverso;
  int rows = argc > 1 ? atoi(argv[1]) : 1000000;
  int cols = argc > 2 ? atoi(argv[2]) : 512;
  int iters = argc > 3 ? atoi(argv[3]) : 2;
  SetCMDFlag<std::string>("updater_type", std::string("sgd"));
  int zero = 1;
  char* av[] = {argv[0], nullptr};
  MV_Init(&zero, av);
  const size_t size = static_cast<size_t>(rows) * cols;
  MatrixTableOption<float> opt(rows, cols);
  auto* table = MV_CreateTable(opt);
  std::vector<float> delta(size, 1e-3f), data(size, 0.f);
  AddOption ao;
  table->Add(delta.data(), size, &ao);          // warm-up (first touch of every buffer on both sides)
  table->Get(data.data(), size);
  MV_Barrier();
  Timer t;
  t.Start();
  for (int i = 0; i < iters; ++i) table->Add(delta.data(), size, &ao);
  double add_ms = t.elapse() / iters;
  MV_Barrier();
  t.Start();
  for (int i = 0; i < iters; ++i) table->Get(data.data(), size);
  double get_ms = t.elapse() / iters;
  MV_Barrier();
  // every worker subtracted 1e-3 (sgd: data -= delta) 1 + iters times
  double expect = -1e-3 * (1 + iters) * MV_NumWorkers();
  int ok = (data[0] - expect) < 1e-5 && (expect - data[0]) < 1e-5 && (data[size - 1] - expect) < 1e-5 &&
           (expect - data[size - 1]) < 1e-5;
  printf("{\"rank\": %d, \"size\": %d, \"rows\": %d, \"cols\": %d, \"iters\": %d, \"add_ms\": %.3f, \"get_ms\": %.3f, "
         "\"verified\": %s}\n", MV_Rank(), MV_Size(), rows, cols, iters, add_ms, get_ms, ok ? "true" : "false");
  fflush(stdout);
  MV_ShutDown();
  return ok ? 0 : 3;
}
