// mv_device_test -- scenarios of the native C++ device runtime (multiverso/device/device.h)
// on 1..8 GPUs, one process per GPU:
//
//   build/bin/mv_device_test all                                   (1 GPU)
//   python tools/mvrun.py -n 2 -- build/bin/mv_device_test all -sync=true
//   python tools/mvrun.py -n 8 -- build/bin/mv_device_test array -sync=false
//
// The exact-integer expectations are the reference's (Test/test_array_table.cpp:11-47,
// Test/test_matrix_table.cpp:9-99, Test/unittests/test_kv.cpp:25-39, Test/test_allreduce.cpp);
// the updater scenario diffs the fused kernels against the host updaters' arithmetic.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "multiverso/device/device.h"
#include "multiverso/multiverso.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"

namespace multiverso { MV_DECLARE_bool(sync); }
using namespace multiverso;
namespace dev = multiverso::device;

static int g_fail = 0;
#define EXPECT(cond)                                                                                   \
  do {                                                                                                 \
    if (!(cond)) {                                                                                     \
      fprintf(stderr, "[rank %d] EXPECT failed: %s (%s:%d)\n", dev::Rank(), #cond, __FILE__, __LINE__); \
      ++g_fail;                                                                                        \
    }                                                                                                  \
  } while (0)

// A device array with host mirror helpers.
template <typename T>
struct DeviceArray {
  explicit DeviceArray(size_t n) : n(n), ptr(static_cast<T*>(dev::DeviceAlloc(std::max<size_t>(n, 1) * sizeof(T)))) {}
  ~DeviceArray() { dev::DeviceFree(ptr); }
  void Upload(const std::vector<T>& h) { dev::CopyToDevice(ptr, h.data(), h.size() * sizeof(T)); }
  std::vector<T> Download() const {
    std::vector<T> h(n);
    dev::CopyToHost(h.data(), ptr, n * sizeof(T));
    return h;
  }
  size_t n;
  T* ptr;
};

// A whole-table Add from every worker. BSP: the collective fused Add. Async: the one-sided pushes
// of the ranks are serialised here (stateful updaters read-modify-write the owner's state
// remotely, so concurrent pushes are not bit-reproducible -- by design, like the reference's
// async server -- and the scenarios below want exact expectations).
template <typename T>
static void AddFromEveryWorker(dev::DenseTable<T>* table, const T* delta, const AddOption* opt) {
  if (MV_CONFIG(sync) || dev::Size() == 1) {
    table->Add(delta, opt);
    return;
  }
  for (int r = 0; r < dev::Size(); ++r) {
    if (dev::Rank() == r) table->Add(delta, opt);
    dev::Barrier();
  }
}

template <typename T>
static void TestArray(int64_t size) {
  const int W = MV_NumWorkers();
  dev::ArrayTable<T> table(size);
  std::vector<T> delta(size);
  for (int64_t i = 0; i < size; ++i) delta[i] = static_cast<T>(i % 97 + 1);
  DeviceArray<T> d(size), out(size);
  d.Upload(delta);
  // the reference scenario runs rank-dependent iteration counts in sync mode to exercise
  // FinishTrain; async mode fences with barriers instead
  const int iters = MV_CONFIG(sync) ? 3 + dev::Rank() : 4;
  for (int it = 1; it <= iters; ++it) {
    table.Add(d.ptr);
    if (!MV_CONFIG(sync)) dev::Barrier();
    table.Get(out.ptr);
    const std::vector<T> got = out.Download();
    // sync: every worker's it-th Get sees exactly it Adds from every still-active worker
    int active = 0;
    for (int r = 0; r < dev::Size(); ++r) active += (!MV_CONFIG(sync) || it <= 3 + r) ? 1 : 0;
    (void)active;
    bool ok = true;
    if (!MV_CONFIG(sync) || it <= 3) {
      for (int64_t i = 0; i < size && ok; ++i) ok = got[i] == static_cast<T>(delta[i] * it * W);
    } else {
      // past the common prefix only monotonic growth can be asserted rank-locally
      for (int64_t i = 0; i < size && ok; ++i) ok = got[i] >= static_cast<T>(delta[i] * 3 * W);
    }
    EXPECT(ok);
    if (!MV_CONFIG(sync)) dev::Barrier();
  }
  table.FinishTrain();
}

static void TestMatrix() {
  const int W = MV_NumWorkers();
  const int64_t rows = 64, cols = 8;
  dev::MatrixTable<float> table(rows, cols);
  std::vector<float> base(rows * cols);
  for (int64_t i = 0; i < rows * cols; ++i) base[i] = static_cast<float>(i + 1);
  const std::vector<int64_t> ids = {0, 1, 3, 7, 40, 63};
  std::vector<float> rowvals(ids.size() * cols);
  for (size_t j = 0; j < ids.size(); ++j)
    for (int64_t c = 0; c < cols; ++c) rowvals[j * cols + c] = base[ids[j] * cols + c];
  DeviceArray<float> d(base.size()), out(base.size()), dr(rowvals.size()), outr(rowvals.size());
  DeviceArray<int64_t> dids(ids.size());
  d.Upload(base);
  dr.Upload(rowvals);
  dids.Upload(ids);
  for (int count = 1; count <= 3; ++count) {
    table.Add(d.ptr);
    dev::Barrier();
    table.AddRows(dids.ptr, static_cast<int64_t>(ids.size()), dr.ptr);
    dev::Barrier();
    table.Get(out.ptr);
    const std::vector<float> got = out.Download();
    bool ok = true;
    for (int64_t r = 0; r < rows && ok; ++r) {
      const bool doubled = std::find(ids.begin(), ids.end(), r) != ids.end();
      for (int64_t c = 0; c < cols && ok; ++c)
        ok = got[r * cols + c] == base[r * cols + c] * count * W * (doubled ? 2 : 1);
    }
    EXPECT(ok);
    table.GetRows(dids.ptr, static_cast<int64_t>(ids.size()), outr.ptr);
    const std::vector<float> gr = outr.Download();
    ok = true;
    for (size_t j = 0; j < gr.size() && ok; ++j) ok = gr[j] == rowvals[j] * count * W * 2;
    EXPECT(ok);
    dev::Barrier();
  }
}

static void TestKV() {
  const int W = MV_NumWorkers();
  dev::KVTable<int64_t> counts(1024);
  dev::KVTable<float> kv(1024);
  EXPECT(kv.Get(0) == 0.0f);
  kv.Add(0, 3.0f);
  counts.Add(4, 1000 + dev::Rank());
  dev::Barrier();
  EXPECT(kv.Get(0) == 3.0f * W);
  int64_t expect = 0;
  for (int r = 0; r < dev::Size(); ++r) expect += 1000 + r;
  EXPECT(counts.Get(4) == expect);
  dev::Barrier();
  kv.Add(0, -4.0f);
  dev::Barrier();
  EXPECT(kv.Get(0) == -1.0f * W);
  const std::vector<int64_t> keys = {1, 2, 1000003, 77};
  const std::vector<float> vals = {1.0f, 2.0f, 0.5f, 0.0f};
  DeviceArray<int64_t> dk(keys.size());
  DeviceArray<float> dv(vals.size()), dout(vals.size());
  dk.Upload(keys);
  dv.Upload(vals);
  kv.Add(dk.ptr, dv.ptr, 3);
  dev::Barrier();
  kv.Get(dk.ptr, dout.ptr, 4);
  const std::vector<float> got = dout.Download();
  EXPECT(got[0] == 1.0f * W && got[1] == 2.0f * W && got[2] == 0.5f * W && got[3] == 0.0f);
  dev::Barrier();
}

static void TestAggregate() {
  const int N = dev::Size();
  for (int64_t n : {1, 1000, 300000, 3000001}) {   // fused latency path and staged two-shot path
    std::vector<float> h(n);
    for (int64_t i = 0; i < n; ++i) h[i] = static_cast<float>((i % 13) + dev::Rank());
    DeviceArray<float> d(n);
    d.Upload(h);
    for (int rep = 0; rep < 2; ++rep) {
      if (rep) d.Upload(h);
      dev::Aggregate(d.ptr, n);
      dev::StreamSync();
      dev::CheckWatchdog();
      const std::vector<float> got = d.Download();
      bool ok = true;
      for (int64_t i = 0; i < n && ok; ++i) ok = got[i] == static_cast<float>((i % 13) * N + N * (N - 1) / 2);
      EXPECT(ok);
    }
  }
  std::vector<int> ones(17, 1);
  DeviceArray<int> di(17);
  di.Upload(ones);
  dev::Aggregate(di.ptr, 17);
  dev::StreamSync();
  EXPECT(di.Download()[16] == N);
  dev::Barrier();
}

// One step of every stateful updater against the host arithmetic
// (include/multiverso/updater/*.h), whole-table Add with W workers applied in worker order.
static void TestUpdaters() {
  const int W = MV_NumWorkers();
  const int64_t n = 4099;
  for (const char* name : {"sgd", "momentum_sgd", "adagrad"}) {
    dev::ArrayTable<float> table(n, dev::TableInit::Fill(1.0), name);
    AddOption opt;
    opt.set_learning_rate(0.1f);
    opt.set_momentum(0.5f);
    opt.set_rho(0.05f);
    std::vector<float> delta(n), expect(n, 1.0f), state(n, 0.0f);
    for (int64_t i = 0; i < n; ++i) delta[i] = 0.01f * static_cast<float>(i % 31 + 1);
    DeviceArray<float> d(n), out(n);
    d.Upload(delta);
    for (int step = 0; step < 2; ++step) {
      AddFromEveryWorker(&table, d.ptr, &opt);
      for (int w = 0; w < W; ++w)
        for (int64_t i = 0; i < n; ++i) {
          if (!strcmp(name, "sgd")) {
            expect[i] -= delta[i];
          } else if (!strcmp(name, "momentum_sgd")) {
            state[i] = 0.5f * state[i] + 0.5f * delta[i];
            expect[i] -= state[i];
          }
        }
    }
    table.Get(out.ptr);
    const std::vector<float> got = out.Download();
    if (strcmp(name, "adagrad") != 0) {
      double worst = 0;
      for (int64_t i = 0; i < n; ++i) worst = std::max(worst, std::fabs(double(got[i]) - expect[i]));
      EXPECT(worst < 1e-4);
    } else {
      // per-worker history: after 2 identical steps of g = delta/lr, G^2 = 2 g^2 on every worker
      double worst = 0;
      for (int64_t i = 0; i < n; ++i) {
        const double g = delta[i] / 0.1;
        const double e = 1.0 - W * (0.05 / std::sqrt(g * g + 1e-6) * g + 0.05 / std::sqrt(2 * g * g + 1e-6) * g);
        worst = std::max(worst, std::fabs(got[i] - e));
      }
      EXPECT(worst < 1e-3);
    }
    dev::Barrier();
  }
}

// Sparse delta pull of the device MatrixTable (reference Matrix<T> with is_sparse, src/table/matrix.cpp:421-572):
// the first GetStale returns every row, later ones exactly the rows somebody added to since this worker's last pull.
static void TestSparseDeltaPull() {
  const int W = MV_NumWorkers();
  const int me = std::max(MV_WorkerId(), 0);
  const int64_t rows = 1000, cols = 8;
  dev::MatrixTable<float> table(rows, cols);
  table.EnableSparse();
  DeviceArray<int64_t> ids(rows);
  DeviceArray<float> out(rows * cols);
  EXPECT(table.GetStale(ids.ptr, out.ptr) == rows);           // everything is stale before the first pull
  EXPECT(table.GetStale(ids.ptr, out.ptr) == 0);
  dev::Barrier();
  // worker w adds 1.0 to rows w, w + 2W, w + 4W, ... (row sets of different workers are disjoint)
  std::vector<int64_t> mine;
  for (int64_t r = me; r < rows; r += 2 * W) mine.push_back(r);
  std::vector<float> vals(mine.size() * cols, 1.0f);
  DeviceArray<int64_t> dids(mine.size());
  DeviceArray<float> dvals(vals.size());
  dids.Upload(mine);
  dvals.Upload(vals);
  table.AddRows(dids.ptr, static_cast<int64_t>(mine.size()), dvals.ptr);
  dev::Barrier();
  std::vector<int64_t> expect;
  for (int64_t r = 0; r < rows; ++r)
    if (r % (2 * W) < W) expect.push_back(r);
  const int64_t n = table.GetStale(ids.ptr, out.ptr);
  EXPECT(n == static_cast<int64_t>(expect.size()));
  const std::vector<int64_t> got_ids = ids.Download();
  const std::vector<float> got = out.Download();
  bool ok = n == static_cast<int64_t>(expect.size());
  for (int64_t i = 0; ok && i < n; ++i) ok = got_ids[i] == expect[i] && got[i * cols] == 1.0f && got[i * cols + cols - 1] == 1.0f;
  EXPECT(ok);
  EXPECT(table.GetStale(ids.ptr, out.ptr) == 0);
  dev::Barrier();
  // whole-table Add with two non-zero rows: only those become stale
  std::vector<float> dense(rows * cols, 0.0f);
  if (me == 0)
    for (int64_t c = 0; c < cols; ++c) dense[5 * cols + c] = dense[999 * cols + c] = 2.0f;
  DeviceArray<float> ddense(dense.size());
  ddense.Upload(dense);
  if (MV_CONFIG(sync) || dev::Size() == 1) {
    table.AddSparse(ddense.ptr);
  } else {
    for (int r = 0; r < dev::Size(); ++r) {
      if (dev::Rank() == r) table.AddSparse(ddense.ptr);
      dev::Barrier();
    }
  }
  dev::Barrier();
  const int64_t n2 = table.GetStale(ids.ptr, out.ptr);
  const std::vector<int64_t> ids2 = ids.Download();
  EXPECT(n2 == 2 && ids2[0] == 5 && ids2[1] == 999);
  dev::Barrier();
}

// The AddOption travels with each worker's request: per-worker learning rates under AdaGrad (per-worker
// history).  Worker w adds delta = lr_w twice; the owner moves every element by rho/sqrt(1) + rho/sqrt(2) per
// worker only if it divides worker w's delta by worker w's OWN learning rate.
static void TestPerWorkerOption() {
  const int W = MV_NumWorkers();
  const int64_t n = 2053;
  dev::ArrayTable<float> table(n, dev::TableInit::Fill(0.0), "adagrad");
  const float lr = 0.01f * static_cast<float>(MV_WorkerId() + 1);
  AddOption opt;
  opt.set_learning_rate(lr);
  opt.set_rho(0.1f);
  std::vector<float> delta(n, lr);
  DeviceArray<float> d(n), out(n);
  d.Upload(delta);
  for (int step = 0; step < 2; ++step) AddFromEveryWorker(&table, d.ptr, &opt);
  dev::Barrier();
  table.Get(out.ptr);
  const std::vector<float> got = out.Download();
  const double e = -W * (0.1 / std::sqrt(1.0 + 1e-6) + 0.1 / std::sqrt(2.0 + 1e-6));
  double worst = 0;
  for (int64_t i = 0; i < n; ++i) worst = std::max(worst, std::fabs(got[i] - e));
  EXPECT(worst < 1e-3);
  dev::Barrier();
}

static void TestCheckpoint() {
  const int64_t n = 1000;
  dev::ArrayTable<float> table(n, dev::TableInit::Uniform(-1, 1, 7), "momentum_sgd");
  AddOption opt;
  opt.set_momentum(0.9f);
  std::vector<float> delta(n, 0.25f);
  DeviceArray<float> d(n), a(n), b(n);
  d.Upload(delta);
  AddFromEveryWorker(&table, d.ptr, &opt);
  dev::Barrier();
  const std::string path = "/tmp/mv_device_ckpt_" + std::to_string(dev::Rank()) + ".bin";
  {
    std::unique_ptr<Stream> s(StreamFactory::GetStream(URI(path), FileOpenMode::BinaryWrite));
    table.Store(s.get());
  }
  dev::Barrier();                      // async pushes of the next step must not race with a peer's Store
  AddFromEveryWorker(&table, d.ptr, &opt);   // continue ...
  dev::Barrier();
  table.Get(a.ptr);
  dev::Barrier();
  {
    std::unique_ptr<Stream> s(StreamFactory::GetStream(URI(path), FileOpenMode::BinaryRead));
    table.Load(s.get());               // ... roll back shard + momentum state
  }
  dev::Barrier();
  AddFromEveryWorker(&table, d.ptr, &opt);   // the replayed step must reproduce the first continuation
  dev::Barrier();
  table.Get(b.ptr);
  EXPECT(a.Download() == b.Download());
  remove(path.c_str());
  dev::Barrier();
}

// Whole-table Add / Get of the BASELINE MatrixTable (1M x 512 fp32 unless `rows` is given), timed with
// CUDA events after 3 warm-up rounds; the slowest rank's time is what counts (max over ranks: every
// rank prints its own line). Same kernels as bench/matrix_bw.py, driven from C++.
static void BenchMatrix(int64_t rows) {
  const int64_t cols = 512;
  dev::MatrixTable<float> table(rows, cols, dev::TableInit(), "sgd");
  const size_t n = static_cast<size_t>(rows) * cols;
  float* out = static_cast<float*>(dev::DeviceAlloc(n * sizeof(float)));
  dev::EventTimer timer;
  double add_ms = 0, get_ms = 0;
  const int warmup = 3, iters = 10;
  for (int i = 0; i < warmup + iters; ++i) {
    float* delta = table.Staging();                 // zero-copy Add: the delta is produced in the staging buffer
    dev::Barrier();
    timer.Start();
    table.Wait(table.AddStagedAsync());
    const float a = timer.StopMs();
    dev::Barrier();
    timer.Start();
    table.Get(out);
    const float g = timer.StopMs();
    if (i >= warmup) {
      add_ms += a;
      get_ms += g;
    }
    (void)delta;
  }
  add_ms /= iters;
  get_ms /= iters;
  const double gb = n * sizeof(float) / 1e9;
  printf("{\"bench\": \"matrix_table\", \"rank\": %d, \"ranks\": %d, \"rows\": %lld, \"cols\": %lld, \"add_ms\": %.4f, "
         "\"get_ms\": %.4f, \"add_gbs\": %.1f, \"get_gbs\": %.1f}\n",
         dev::Rank(), dev::Size(), static_cast<long long>(rows), static_cast<long long>(cols), add_ms, get_ms,
         gb / add_ms * 1e3, gb / get_ms * 1e3);
  dev::Barrier();
  dev::DeviceFree(out);
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    fprintf(stderr, "usage: mv_device_test all|array|matrix|kv|aggregate|updaters|checkpoint|bench [rows] [-flag=value ...]\n");
    return 2;
  }
  const std::string which = argv[1];
  dev::Init(&argc, argv);
  const bool all = which == "all";
  if (all || which == "array") {
    TestArray<float>(100003);
    TestArray<int>(5);          // fewer elements than servers on 8 GPUs
    TestArray<double>(4096);
  }
  if (all || which == "matrix") TestMatrix();
  if (all || which == "kv") TestKV();
  if (all || which == "aggregate") TestAggregate();
  if (all || which == "updaters") TestUpdaters();
  if (all || which == "peropt") TestPerWorkerOption();
  if (all || which == "sparse") TestSparseDeltaPull();
  if (all || which == "checkpoint") TestCheckpoint();
  if (which == "bench") BenchMatrix(argc > 2 ? atoll(argv[2]) : 1000000);
  dev::Barrier();
  printf("[mv_device_test %s rank %d/%d gpu %d] %s\n", which.c_str(), dev::Rank(), dev::Size(), dev::DeviceId(),
         g_fail ? "FAIL" : "PASS");
  dev::ShutDown();
  return g_fail ? 1 : 0;
}
