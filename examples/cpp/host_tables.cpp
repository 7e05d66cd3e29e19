// Minimal parameter-server program on the C++ host runtime (CPU, TCP control plane):
//
//   g++ -std=c++17 -Iinclude examples/cpp/host_tables.cpp -o host_tables \
//       -Lmultiverso_b200/_lib -lmultiverso -Wl,-rpath,$PWD/multiverso_b200/_lib -pthread -fopenmp
//   python tools/mvrun.py -n 4 -- ./host_tables -sync=true -updater_type=sgd
//
// Every rank is worker + server (the default -ps_role). Each worker pushes a gradient into a
// dense ArrayTable (the server applies the updater), a few rows into a MatrixTable, a counter
// into a KVTable, and reads everything back.
#include <cstdio>
#include <vector>

#include "multiverso/multiverso.h"
#include "multiverso/table/array_table.h"
#include "multiverso/table/kv_table.h"
#include "multiverso/table/matrix_table.h"

using namespace multiverso;

int main(int argc, char* argv[]) {
  MV_Init(&argc, argv);                                     // consumes -key=value flags
  const int W = MV_NumWorkers();

  ArrayWorker<float>* weights = MV_CreateTable(ArrayTableOption<float>(1000));   // collective
  MatrixWorkerTable<float>* emb = MV_CreateTable(MatrixTableOption<float>(100, 8));
  KVWorkerTable<int, int64_t>* counters = MV_CreateTable(KVTableOption<int, int64_t>());

  std::vector<float> grad(1000, 0.5f), w(1000);
  AddOption opt;
  opt.set_learning_rate(0.1f);
  weights->Add(grad.data(), grad.size(), &opt);             // blocking; AddAsync + Wait(id) also exist
  MV_Barrier();
  weights->Get(w.data(), w.size());                         // sgd: w = -0.5 * W, default: +0.5 * W

  std::vector<integer_t> rows = {3, 42, 99};
  std::vector<float> delta(rows.size() * 8, 1.0f), back(rows.size() * 8);
  emb->Add(delta.data(), delta.size(), rows.data(), static_cast<int>(rows.size()));
  MV_Barrier();
  emb->Get(back.data(), back.size(), rows.data(), static_cast<int>(rows.size()));

  counters->Add(7, 1000 + MV_Rank());
  MV_Barrier();
  counters->Get(7);
  const long long total = counters->raw()[7];

  printf("rank %d/%d: w[0] = %g, emb[42][0] = %g, counter = %lld\n", MV_Rank(), MV_Size(), w[0], back[8], total);
  const bool ok = (w[0] == 0.5f * W || w[0] == -0.5f * W) && (back[8] == 1.0f * W || back[8] == -1.0f * W);
  MV_Barrier();
  delete weights;
  delete emb;
  delete counters;
  MV_ShutDown();
  return ok ? 0 : 1;
}
