// multiverso-b200 host runtime test driver.
//
//   mv_test unit                      single-process unit tests (blob, message, node, flags,
//                                     allocator, queue, io, filters, array partition, kv, sync)
//   mv_test kv|array|net|matrix|allreduce|sparse|checkpoint [flags]
//                                     end-to-end scenarios, run with MV_RANK / MV_SIZE / MV_PORT
//                                     set per process (the reference used `mpirun -np 4
//                                     ./multiverso.test <name>`, Test/main.cpp:12-25)
//
// Scenarios mirror Test/unittests/*.cpp and Test/test_*.cpp with their exact integer
// expectations; exit code 0 = pass.
#include <atomic>
#include <thread>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <unistd.h>
#include <vector>

#include "multiverso/dashboard.h"
#include "multiverso/io/io.h"
#include "multiverso/multiverso.h"
#include "multiverso/net.h"
#include "multiverso/net/allreduce_engine.h"
#include "multiverso/table/array_table.h"
#include "multiverso/table/kv_table.h"
#include "multiverso/table/matrix.h"
#include "multiverso/table/matrix_table.h"
#include "multiverso/table/sparse_matrix_table.h"
#include "multiverso/table/sparse_table.h"
#include "multiverso/util/allocator.h"
#include "multiverso/util/async_buffer.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/util/mt_queue.h"
#include "multiverso/util/parallel_for.h"
#include "multiverso/util/quantization_util.h"
#include "multiverso/util/waiter.h"

namespace multiverso { MV_DECLARE_int(backup_worker_ratio); }
using namespace multiverso;

static int g_fail = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      fprintf(stderr, "[rank %d] EXPECT failed: %s (%s:%d)\n", MV_Rank(), #cond, __FILE__, __LINE__); \
      ++g_fail;                                                             \
    }                                                                       \
  } while (0)

// ------------------------------------------------------------------------------------- unit
static void TestBlobMessageNode() {
  Blob a(16);
  EXPECT(a.size() == 16 && a.size<int>() == 4);
  for (int i = 0; i < 4; ++i) a.As<int>(i) = i * 3;
  Blob b = a;   // shallow
  b.As<int>(1) = 42;
  EXPECT(a.As<int>(1) == 42);
  int raw[3] = {7, 8, 9};
  Blob c(raw, sizeof raw);   // copies
  raw[0] = 0;
  EXPECT(c.As<int>(0) == 7 && c[0] == 7);
  Message m;
  m.set_src(1); m.set_dst(2); m.set_type(MsgType::Request_Get); m.set_table_id(3); m.set_msg_id(4);
  m.Push(c);
  EXPECT(m.size() == 1 && m.src() == 1 && m.dst() == 2 && m.table_id() == 3 && m.msg_id() == 4);
  MessagePtr r(m.CreateReplyMessage());
  EXPECT(r->src() == 2 && r->dst() == 1 && r->type() == MsgType::Reply_Get && r->msg_id() == 4 && r->size() == 0);
  EXPECT(node::is_worker(Role::ALL) && node::is_server(Role::ALL));
  EXPECT(node::is_worker(Role::WORKER) && !node::is_server(Role::WORKER));
  EXPECT(!node::is_worker(Role::SERVER) && node::is_server(Role::SERVER));
  EXPECT(!node::is_worker(Role::NONE) && !node::is_server(Role::NONE));
}

MV_DEFINE_int(test_int_flag, 3, "unit test flag");
MV_DEFINE_double(test_double_flag, 0.5, "unit test flag");
MV_DEFINE_string(test_string_flag, "abc", "unit test flag");
MV_DEFINE_bool(test_bool_flag, false, "unit test flag");

static void TestFlagsAllocatorQueue() {
  const char* args[] = {"prog", "-test_int_flag=11", "positional", "-test_double_flag=2.25",
                        "-unknown=1", "-test_string_flag=hello", "-test_bool_flag=true", "-dash-no-eq"};
  int argc = 8;
  char* argv[8];
  for (int i = 0; i < 8; ++i) argv[i] = const_cast<char*>(args[i]);
  ParseCMDFlags(&argc, argv);
  EXPECT(argc == 4);   // prog, positional, -unknown=1, -dash-no-eq are left in place
  EXPECT(MV_CONFIG(test_int_flag) == 11 && MV_CONFIG(test_double_flag) == 2.25);
  EXPECT(MV_CONFIG(test_string_flag) == "hello" && MV_CONFIG(test_bool_flag));
  MV_SetFlag<int>("test_int_flag", 5);
  EXPECT(MV_CONFIG(test_int_flag) == 5);

  SmartAllocator sa;
  char* p = sa.Alloc(100);
  sa.Refer(p);
  sa.Free(p);
  EXPECT(sa.pooled_blocks() == 0);
  sa.Free(p);
  EXPECT(sa.pooled_blocks() == 1);
  char* q = sa.Alloc(120);   // same 128-byte class -> recycled
  EXPECT(q == p && sa.pooled_blocks() == 0);
  sa.Free(q);

  MtQueue<int> mq;
  for (int i = 0; i < 5; ++i) mq.Push(i);
  int v = -1;
  EXPECT(mq.Size() == 5 && mq.Front(v) && v == 0 && mq.Pop(v) && v == 0 && mq.TryPop(v) && v == 1);
  mq.Exit();
  EXPECT(!mq.Alive());
  while (mq.Pop(v)) {}
  EXPECT(mq.Empty());
  Waiter w(2);
  w.Notify();
  EXPECT(!w.Done());
  w.Notify();
  w.Wait();
  EXPECT(w.Done());

  std::vector<int> b0, b1;
  int fills = 0;
  ASyncBuffer<std::vector<int>> ab(&b0, &b1, [&](std::vector<int>* b) { b->assign(3, ++fills); });
  std::vector<int>* g1 = ab.Get();
  const int first = (*g1)[0];           // read before the next Get() recycles this buffer
  std::vector<int>* g2 = ab.Get();
  EXPECT(g1 != g2 && first == 1 && (*g2)[0] == 2);
  ab.Join();
}

static void TestIOAndFilters() {
  URI u1("file:///tmp/a.txt"), u2("/tmp/b.txt"), u3("hdfs://namenode:9000/dir/f");
  EXPECT(u1.scheme == "file" && u1.name == "/tmp/a.txt");
  EXPECT(u2.scheme == "file" && u2.name == "/tmp/b.txt");
  EXPECT(u3.scheme == "hdfs" && u3.host == "namenode:9000" && u3.name == "/dir/f");
  std::string path = "/tmp/mv_test_io_" + std::to_string(getpid()) + ".txt";
  {
    std::unique_ptr<Stream> s(StreamFactory::GetStream(URI(path), FileOpenMode::Write));
    EXPECT(s && s->Good());
    const char* text = "line one\nsecond\r\n\nlast";
    s->Write(text, strlen(text));
  }
  {
    TextReader rd(URI(path), 8);   // tiny buffer exercises refills
    std::string line;
    std::vector<std::string> lines;
    while (rd.GetLine(line)) lines.push_back(line);
    EXPECT(lines.size() == 4 && lines[0] == "line one" && lines[1] == "second" && lines[2].empty() && lines[3] == "last");
  }
  remove(path.c_str());

  // SparseFilter round trip: sparse blob is compressed, dense blob is passed through
  std::vector<float> sparse(100, 0.f), dense(100);
  sparse[3] = 1.5f; sparse[77] = -2.f;
  std::iota(dense.begin(), dense.end(), 1.f);
  integer_t key = 5;
  AddOption opt;
  std::vector<Blob> in = {Blob(&key, sizeof key), Blob(sparse.data(), 400), Blob(dense.data(), 400), Blob(opt.data(), opt.size())};
  std::vector<Blob> mid, out;
  SparseFilter<float, int32_t> f(0.0, true);
  f.FilterIn(in, &mid);
  EXPECT(mid.size() == 5 && mid[2].size() == 2 * 8 && mid[3].size() == 400);
  f.FilterOut(mid, &out);
  EXPECT(out.size() == 4 && out[1].size() == 400 && memcmp(out[1].data(), sparse.data(), 400) == 0);
  EXPECT(memcmp(out[2].data(), dense.data(), 400) == 0 && out[3].size() == opt.size());
  OneBitsFilter<float> ob;
  std::vector<float> g = {1.f, 3.f, -2.f, -4.f, 2.f};
  std::vector<Blob> oin = {Blob(&key, sizeof key), Blob(g.data(), 20)}, omid, oout;
  ob.FilterIn(oin, &omid);
  ob.FilterOut(omid, &oout);
  EXPECT(oout[1].size<float>() == 5 && oout[1].As<float>(0) == 2.f && oout[1].As<float>(2) == -3.f);
}

// ParallelFor / ParallelMemcpy: every index exactly once, from several caller threads at once
// (the caller, the worker actor and the server actor share the pool), odd sizes, big copies.
static void TestParallelFor() {
  EXPECT(ParallelForCapacity() >= 1);
  for (int64_t n : {0, 1, 7, 1000, 100003}) {
    std::vector<std::atomic<int>> hits(static_cast<size_t>(n));
    for (auto& h : hits) h.store(0);
    ParallelFor(n, 8, [&](int64_t lo, int64_t hi) {
      for (int64_t i = lo; i < hi; ++i) hits[static_cast<size_t>(i)].fetch_add(1);
    });
    bool once = true;
    for (auto& h : hits) once = once && h.load() == 1;
    EXPECT(once);
  }
  std::vector<std::thread> callers;
  std::atomic<int> bad{0};
  for (int t = 0; t < 4; ++t)
    callers.emplace_back([&bad, t] {
      for (int rep = 0; rep < 20; ++rep) {
        const int64_t n = 5000 + 37 * t + rep;
        std::atomic<int64_t> sum{0};
        ParallelFor(n, 4, [&](int64_t lo, int64_t hi) {
          int64_t s = 0;
          for (int64_t i = lo; i < hi; ++i) s += i;
          sum.fetch_add(s);
        });
        if (sum.load() != n * (n - 1) / 2) bad.fetch_add(1);
      }
    });
  for (auto& c : callers) c.join();
  EXPECT(bad.load() == 0);
  std::vector<char> src((24u << 20) + 12345), dst(src.size(), 0);
  for (size_t i = 0; i < src.size(); ++i) src[i] = static_cast<char>(i * 131 + 7);
  ParallelMemcpy(dst.data(), src.data(), src.size());
  EXPECT(src == dst);
  std::vector<char> small_dst(100, 0);
  ParallelMemcpy(small_dst.data(), src.data(), small_dst.size());
  EXPECT(std::equal(small_dst.begin(), small_dst.end(), src.begin()));
}

static void TestAllreduceSchedules() {
  // Bruck: total blocks received == size-1 for every rank / size
  for (int n = 1; n <= 9; ++n)
    for (int r = 0; r < n; ++r) {
      int got = 1;
      for (auto& s : BruckSchedule(r, n)) got += s.blocks;
      EXPECT(got == n);
    }
  auto steps = RecursiveHalvingSchedule(5, 8);
  EXPECT(steps.size() == 3 && steps[0].peer == 1 && steps[1].peer == 7 && steps[2].peer == 4);
  EXPECT(steps.back().keep_lo == 5 && steps.back().keep_hi == 6);
}

static void UnitWithRuntime(bool sync) {
  MV_SetFlag<bool>("sync", sync);
  MV_Init();
  EXPECT(MV_Size() == 1 && MV_Rank() == 0 && MV_NumWorkers() == 1 && MV_NumServers() == 1);
  EXPECT(MV_WorkerId() == 0 && MV_ServerId() == 0);
  {
    // test_array.cpp:26-44
    const size_t n = 100000;
    auto* t = MV_CreateTable(ArrayTableOption<float>(n));
    std::vector<float> delta(n), got(n);
    for (size_t i = 0; i < n; ++i) delta[i] = static_cast<float>(i);
    t->Add(delta.data(), n);
    t->Get(got.data(), n);
    for (size_t i = 0; i < n; ++i) if (got[i] != delta[i]) { EXPECT(false); break; }
    int a = t->AddAsync(delta.data(), n);
    int g = t->GetAsync(got.data(), n);
    t->Wait(a);
    t->Wait(g);
    for (size_t i = 0; i < n; ++i) if (got[i] != 2 * delta[i]) { EXPECT(false); break; }
    // white-box Partition (test_array.cpp:46-66): 1 server => one bucket, key -1 + the values
    integer_t whole = -1;
    std::vector<Blob> kv = {Blob(&whole, sizeof whole), Blob(delta.data(), n * sizeof(float))};
    std::unordered_map<int, std::vector<Blob>> parts;
    EXPECT(t->Partition(kv, MsgType::Request_Add, &parts) == 1);
    EXPECT(parts.count(0) && parts[0].size() == 2 && parts[0][0].As<integer_t>() == -1);
    EXPECT(parts[0][1].size<float>() == n && parts[0][1].As<float>(12345) == 12345.f);
    // two Gets in flight on the same table
    std::vector<float> g1(n), g2(n);
    int h1 = t->GetAsync(g1.data(), n), h2 = t->GetAsync(g2.data(), n);
    t->Wait(h2);
    t->Wait(h1);
    EXPECT(g1[7] == 14.f && g2[7] == 14.f);
    // size-1 array (upstream issue #69, test_multiverso.py:36-42)
    auto* one = MV_CreateTable(ArrayTableOption<int>(1));
    int v = 5, o = 0;
    one->Add(&v, 1);
    one->Add(&v, 1);
    one->Get(&o, 1);
    EXPECT(o == 10);
    delete one;
    delete t;
  }
  {
    // test_kv.cpp:25-39
    auto* kv = MV_CreateTable(KVTableOption<int, int>());
    kv->Get(0);
    EXPECT(kv->raw()[0] == 0);
    kv->Add(0, 3);
    kv->Get(0);
    EXPECT(kv->raw()[0] == 3);
    kv->Add(0, -4);
    kv->Get(0);
    EXPECT(kv->raw()[0] == -1);
    delete kv;
  }
  MV_ShutDown(false);   // keep the net so the next fixture can re-init
}

static int RunUnit() {
  TestBlobMessageNode();
  TestFlagsAllocatorQueue();
  TestIOAndFilters();
  TestParallelFor();
  TestAllreduceSchedules();
  UnitWithRuntime(false);
  UnitWithRuntime(true);    // test_sync.cpp:25-43
  MV_SetFlag<bool>("sync", false);
  return g_fail;
}

// ------------------------------------------------------------------------------ multi-process
static void TestKV() {   // Test/test_kv_table.cpp:8-34
  MV_Init();
  auto* t = MV_CreateTable(KVTableOption<int, float>());
  const int W = MV_NumWorkers();
  for (int it = 1; it <= 5; ++it) {
    if (t) {
      t->Add(std::vector<int>{0, 1, 2, 1000003}, std::vector<float>{1.f, 2.f, 3.f, 0.5f});
      MV_Barrier();
      t->Get(std::vector<int>{0, 1, 2, 1000003, 77});
      EXPECT(t->raw()[0] == 1.f * it * W && t->raw()[1] == 2.f * it * W && t->raw()[2] == 3.f * it * W);
      EXPECT(t->raw()[1000003] == 0.5f * it * W && t->raw()[77] == 0.f);
    } else {
      MV_Barrier();
    }
    MV_Barrier();
  }
  delete t;
  MV_ShutDown();
}

static void TestArray(bool sync) {   // Test/test_array_table.cpp:11-47
  MV_SetFlag<bool>("sync", sync);
  MV_Init();
  const size_t n = 500;
  auto* t = MV_CreateTable(ArrayTableOption<float>(n));
  const int W = MV_NumWorkers();
  if (t) {
    std::vector<float> delta(n), data(n);
    for (size_t i = 0; i < n; ++i) delta[i] = static_cast<float>(i);
    // uneven iteration counts per rank exercise FinishTrain / straggler logic in sync mode
    const int iters = sync ? 10 * (MV_Rank() + 2) : 20;
    for (int it = 0; it < iters; ++it) {
      t->Add(delta.data(), n);
      t->Add(delta.data(), n);
      t->Add(delta.data(), n);
      t->Get(data.data(), n);
      t->Get(data.data(), n);
      t->Get(data.data(), n);
      if (sync && it < 20 && MV_CONFIG(backup_worker_ratio) == 0) {
        // all workers' i-th Get are identical and include every worker's matching Adds
        bool ok = true;
        for (size_t i = 0; i < n; ++i) ok = ok && data[i] == delta[i] * 3 * (it + 1) * W;
        EXPECT(ok);
      }
    }
    if (!sync) {
      MV_Barrier();
      t->Get(data.data(), n);
      bool ok = true;
      for (size_t i = 0; i < n; ++i) ok = ok && data[i] == delta[i] * 3 * iters * W;
      EXPECT(ok);
    }
  } else if (!sync) {
    MV_Barrier();
  }
  MV_ShutDown();
  delete t;
}

static void TestMatrix(bool sparse_wire) {   // Test/test_matrix_table.cpp:9-99
  MV_SetFlag<bool>("sync", true);
  MV_Init();
  const integer_t R = 11, C = 10;
  MatrixWorkerTable<int>* t = MV_CreateTable(MatrixTableOption<int>(R, C));
  MatrixOption<int> so;
  so.num_row = R; so.num_col = C; so.is_sparse = true;
  MatrixWorkerTable<int>* s = sparse_wire
      ? static_cast<MatrixWorkerTable<int>*>(MV_CreateTable(SparseMatrixTableOption<int>(R, C)))
      : static_cast<MatrixWorkerTable<int>*>(MV_CreateTable(so));
  const int W = MV_NumWorkers();
  if (t) {
    std::vector<int> delta(R * C), data(R * C), sdata(R * C, 0);
    for (int i = 0; i < R * C; ++i) delta[i] = i + 1;
    std::vector<integer_t> ids = {0, 1, 3, 7};
    std::vector<int> rowbuf(ids.size() * C);
    for (size_t k = 0; k < ids.size(); ++k)
      for (int j = 0; j < C; ++j) rowbuf[k * C + j] = delta[ids[k] * C + j];
    for (int count = 1; count <= 50; ++count) {
      t->Add(delta.data(), R * C);
      t->Add(rowbuf.data(), rowbuf.size(), ids.data(), static_cast<int>(ids.size()));
      t->Get(data.data(), R * C);
      s->Add(delta.data(), R * C);
      s->Add(rowbuf.data(), rowbuf.size(), ids.data(), static_cast<int>(ids.size()));
      s->Get(sdata.data(), R * C);   // delta pull: only stale rows travel
      bool ok = true, sok = true;
      for (integer_t i = 0; i < R; ++i)
        for (integer_t j = 0; j < C; ++j) {
          int expect = static_cast<int>((i * C + j + 1) * count * W);
          if (i == 0 || i == 1 || i == 3 || i == 7) expect *= 2;
          ok = ok && data[i * C + j] == expect;
          sok = sok && sdata[i * C + j] == expect;
        }
      EXPECT(ok);
      EXPECT(sok);
      // row-set Get into scattered buffers
      std::vector<int> r0(C), r1(C);
      std::vector<int*> ptrs = {r0.data(), r1.data()};
      std::vector<integer_t> two = {3, 10};
      t->Get(two, ptrs, C);
      EXPECT(r0[0] == data[3 * C] && r1[C - 1] == data[10 * C + C - 1]);
      int single[10];
      t->Get(integer_t(5), single, C);
      EXPECT(single[4] == data[5 * C + 4]);
    }
  }
  MV_ShutDown();
  delete t;
  delete s;
}

// Test/test_matrix_perf.cpp:32-171 (TestDensePerf / TestSparsePerf; "perf" tier of SURVEY section 4):
// Get all rows -> every worker Adds its share of the first (p+1)/10 of the rows -> Get all rows,
// timed and verified.  The reference hard-codes 1,000,000 x 50 and 10 turns per percentage; the
// row count is an argument here (default 100,000, one turn) so the scenario can run in CI.
static void TestMatrixPerf(bool sparse, int num_row) {
  MV_Init();
  const int C = 50, W = MV_NumWorkers(), me = MV_WorkerId();
  const size_t size = static_cast<size_t>(num_row) * C;
  std::vector<float> data(size), delta(size);
  for (size_t i = 0; i < size; ++i) delta[i] = static_cast<float>(i % 100003);
  Timer timer, turn_timer;
  for (int percent = 0; percent < 10; ++percent) {
    turn_timer.Start();
    MatrixWorkerTable<float>* t;
    if (sparse) {
      MatrixOption<float> o;
      o.num_row = num_row; o.num_col = C; o.is_sparse = true;
      t = static_cast<MatrixWorkerTable<float>*>(MV_CreateTable(o));
    } else {
      t = MV_CreateTable(MatrixTableOption<float>(num_row, C));
    }
    if (!t) continue;                       // server-only rank
    MV_Barrier();
    if (me == 0 && getenv("MV_TEST_TIMING")) printf("    [table created after %.1f ms]\n", turn_timer.elapse());
    GetOption gopt;
    gopt.set_worker_id(me);
    timer.Start();
    t->Get(data.data(), size, &gopt);
    const double first_ms = timer.elapse();
    MV_Barrier();
    std::vector<integer_t> ids;
    std::vector<float*> ptrs;
    for (int i = 0; i < num_row; ++i)
      if (i % 10 <= percent && i % W == me) { ids.push_back(i); ptrs.push_back(delta.data() + static_cast<size_t>(i) * C); }
    AddOption aopt;
    aopt.set_worker_id(me);
    if (me == 0 && getenv("MV_TEST_TIMING")) printf("    [ids built after %.1f ms]\n", turn_timer.elapse());
    timer.Start();
    if (!ids.empty()) t->Add(ids, ptrs, C, &aopt);
    const double add_ms = timer.elapse();
    MV_Barrier();
    timer.Start();
    t->Get(data.data(), size, &gopt);
    const double get_ms = timer.elapse();
    timer.Start();
    bool ok = true;
    for (int i = 0; i < num_row && ok; ++i)
      for (int c = 0; c < C; ++c) {
        const float expect = (i % 10 <= percent) ? delta[static_cast<size_t>(i) * C + c] : 0.f;
        if (data[static_cast<size_t>(i) * C + c] != expect) { ok = false; break; }
      }
    EXPECT(ok);
    if (me == 0 && getenv("MV_TEST_TIMING")) printf("    [verify %.1f ms, since turn start %.1f ms]\n", timer.elapse(), turn_timer.elapse());
    if (me == 0)
      printf("  %s %d x %d, add %d0%% of the rows: first get %.1f ms, add %.1f ms (%zu rows), get %.1f ms (%.2f GB/s)",
             sparse ? "sparse" : "dense", num_row, C, percent + 1, first_ms, add_ms, ids.size(), get_ms,
             size * sizeof(float) / get_ms / 1e6);
    MV_Barrier();
    delete t;
    if (me == 0) printf(", turn %.1f ms\n", turn_timer.elapse());   // table creation .. verification .. teardown
  }
  if (me == 0) Dashboard::Display();
  MV_ShutDown();
}

static void TestNet() {   // Test/test_net.cpp:9-90
  NetInterface* net = NetInterface::Get();
  net->Init(nullptr, nullptr);
  const char* chunks[3] = {"hello, world", "hello, c++", "hello, multiverso"};
  if (net->rank() == 0) {
    for (int r = 1; r < net->size(); ++r) {
      MessagePtr msg(new Message());
      msg->set_src(0); msg->set_dst(r); msg->set_type(MsgType::Default); msg->set_msg_id(r);
      for (auto c : chunks) msg->Push(Blob(c, strlen(c) + 1));
      EXPECT(net->Send(msg) > 0);
    }
    for (int r = 1; r < net->size(); ++r) {
      MessagePtr reply;
      EXPECT(net->Recv(&reply) != static_cast<size_t>(-1));
      EXPECT(reply->size() == 1 && reply->data()[0].As<int>(0) == reply->src() * 7);
    }
  } else {
    MessagePtr msg;
    EXPECT(net->Recv(&msg) != static_cast<size_t>(-1));
    EXPECT(msg->src() == 0 && msg->msg_id() == net->rank() && msg->size() == 3);
    for (int i = 0; i < 3; ++i) EXPECT(strcmp(msg->data()[i].data(), chunks[i]) == 0);
    MessagePtr reply(msg->CreateReplyMessage());
    int v = net->rank() * 7;
    reply->Push(Blob(&v, sizeof v));
    net->Send(reply);
  }
  // raw SendRecv ring
  int next = (net->rank() + 1) % net->size(), prev = (net->rank() + net->size() - 1) % net->size();
  int out = net->rank() + 100, in = -1;
  if (net->size() > 1) {
    net->SendRecv(next, reinterpret_cast<char*>(&out), 4, prev, reinterpret_cast<char*>(&in), 4);
    EXPECT(in == prev + 100);
  }
  net->Finalize();
}

static void TestAllreduce() {   // Test/test_allreduce.cpp:10-19 (+ large / odd sizes)
  MV_SetFlag<bool>("ma", true);
  MV_Init();
  int a = 1;
  MV_Aggregate(&a, 1);
  EXPECT(a == MV_Size());
  for (int n : {3, 1000, 100003}) {
    std::vector<float> v(n);
    for (int i = 0; i < n; ++i) v[i] = static_cast<float>((i % 17) + MV_Rank());
    MV_Aggregate(v.data(), n);
    bool ok = true;
    const int S = MV_Size();
    for (int i = 0; i < n; ++i) ok = ok && v[i] == static_cast<float>((i % 17) * S + S * (S - 1) / 2);
    EXPECT(ok);
  }
  std::vector<double> d(5000, 0.5);
  MV_Aggregate(d.data(), 5000);
  EXPECT(d[4999] == 0.5 * MV_Size());
  MV_Barrier();
  MV_ShutDown();
  MV_SetFlag<bool>("ma", false);
}

static void TestUpdatersAndCheckpoint(const std::string& updater) {
  MV_SetFlag<std::string>("updater_type", updater);
  MV_Init();
  const size_t n = 1000;
  auto* t = MV_CreateTable(ArrayTableOption<float>(n));
  const int W = MV_NumWorkers();
  std::vector<float> d(n, 0.01f), got(n);
  AddOption opt;
  opt.set_momentum(0.5f); opt.set_learning_rate(0.01f); opt.set_rho(0.1f); opt.set_lambda(0.1f);
  if (t) t->Add(d.data(), n, &opt);
  MV_Barrier();
  if (t) t->Get(got.data(), n);
  float expect = 0.f;
  if (updater == "default") expect = 0.01f * W;
  else if (updater == "sgd") expect = -0.01f * W;
  else if (updater == "momentum_sgd") { float s = 0, x = 0; for (int w = 0; w < W; ++w) { s = 0.5f * s + 0.5f * 0.01f; x -= s; } expect = x; }
  else if (updater == "adagrad") expect = -0.1f / std::sqrt(1.f + 1e-6f) * W;
  if (t && updater != "dcasgd" && updater != "dcasgda") EXPECT(std::fabs(got[n / 2] - expect) < 1e-5f);
  if (t && (updater == "dcasgd" || updater == "dcasgda")) EXPECT(got[n / 2] < 0.f && std::isfinite(got[n / 2]));
  std::string uri = "/tmp/mv_test_ckpt_" + updater;
  EXPECT(MV_SaveTable(0, uri));
  if (t) { t->Add(d.data(), n, &opt); }
  MV_Barrier();
  EXPECT(MV_LoadTable(0, uri));
  if (t) {
    std::vector<float> again(n);
    t->Get(again.data(), n);
    EXPECT(again[n / 2] == got[n / 2]);
  }
  MV_Barrier();
  delete t;
  MV_ShutDown();
  MV_SetFlag<std::string>("updater_type", "default");
}

static void TestAppTables() {   // LogReg SparseTable / FTRLTable (sparse_table.h, ftrl_sparse_table.h)
  MV_Init();
  const int W = MV_NumWorkers();
  auto* t = MV_CreateTable(SparseTableOption<float>(1000003));
  auto* f = MV_CreateTable(FTRLTableOption<float>(5000));
  if (t) {
    std::vector<size_t> keys = {0, 7, 999, 500001, 1000002};
    std::vector<float> vals = {1.f, 2.f, 3.f, 4.f, 5.f}, got(5, -1.f);
    t->Add(keys.data(), vals.data(), keys.size());
    MV_Barrier();
    t->Get(keys.data(), keys.size(), got.data());
    for (size_t i = 0; i < keys.size(); ++i) EXPECT(got[i] == -vals[i] * W);   // server subtracts
    std::vector<size_t> ak;
    std::vector<float> av;
    t->GetAll(&ak, &av);
    EXPECT(ak.size() == keys.size() && av.size() == keys.size());
    std::vector<size_t> fk = {3, 4999};
    std::vector<FTRLEntry<float>> fv = {{1.f, 2.f}, {0.5f, 0.25f}}, fg(2);
    f->Add(fk.data(), fv.data(), 2);
    MV_Barrier();
    f->Get(fk.data(), 2, fg.data());
    EXPECT(fg[0].z == -1.f * W && fg[0].n == -2.f * W && fg[1].n == -0.25f * W);
  } else {
    MV_Barrier();
    MV_Barrier();
  }
  MV_Barrier();
  delete t;
  delete f;
  MV_ShutDown();
}

// Round trip through StreamFactory for any URI scheme: write 3 MB + a text tail, append, read
// back with Stream::Read and TextReader (used with hdfs:// against a stand-in libhdfs).
static void TestStream(const std::string& uri) {
  std::vector<int> payload(750000);
  std::iota(payload.begin(), payload.end(), 7);
  {
    std::unique_ptr<Stream> w(StreamFactory::GetStream(URI(uri), FileOpenMode::BinaryWrite));
    EXPECT(w != nullptr && w->Good());
    if (!w) return;
    w->Write(payload.data(), payload.size() * sizeof(int));
    w->Flush();
  }
  {
    std::unique_ptr<Stream> a(StreamFactory::GetStream(URI(uri), FileOpenMode::BinaryAppend));
    EXPECT(a != nullptr && a->Good());
    const char tail[] = "\nfirst line\nsecond line\n";
    if (a) a->Write(tail, sizeof tail - 1);
  }
  {
    std::unique_ptr<Stream> r(StreamFactory::GetStream(URI(uri), FileOpenMode::BinaryRead));
    EXPECT(r != nullptr && r->Good());
    if (!r) return;
    std::vector<int> back(payload.size());
    EXPECT(r->Read(back.data(), back.size() * sizeof(int)) == back.size() * sizeof(int));
    EXPECT(back == payload);
    char rest[64] = {0};
    EXPECT(r->Read(rest, sizeof rest) == 24);   // short read at the end of the file
  }
  TextReader reader(URI(uri), 1 << 12);
  std::string line, last, before_last;
  while (reader.GetLine(line)) {
    before_last = last;
    last = line;
  }
  EXPECT(before_last == "first line" && last == "second line");
  EXPECT(StreamFactory::GetStream(URI(uri + ".missing"), FileOpenMode::Read) == nullptr ||
         !std::unique_ptr<Stream>(StreamFactory::GetStream(URI(uri + ".missing"), FileOpenMode::Read))->Good());
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    fprintf(stderr, "usage: mv_test unit|kv|array|array_async|net|matrix|sparse|allreduce|apptables|dense_perf [rows]|sparse_perf [rows]|updater:<name>|stream <uri> [-flag=value ...]\n");
    return 2;
  }
  std::string which = argv[1];
  ParseCMDFlags(&argc, argv);
  if (which == "unit") g_fail = RunUnit();
  else if (which == "kv") TestKV();
  else if (which == "array") TestArray(true);
  else if (which == "array_async") TestArray(false);
  else if (which == "net") TestNet();
  else if (which == "matrix") TestMatrix(false);
  else if (which == "sparse") TestMatrix(true);
  else if (which == "allreduce") TestAllreduce();
  else if (which == "apptables") TestAppTables();
  else if (which == "dense_perf" || which == "sparse_perf")
    TestMatrixPerf(which == "sparse_perf", argc > 2 ? atoi(argv[2]) : 100000);
  else if (which.rfind("updater:", 0) == 0) TestUpdatersAndCheckpoint(which.substr(8));
  else if (which == "stream" && argc > 2) TestStream(argv[2]);
  else { fprintf(stderr, "unknown test %s\n", which.c_str()); return 2; }
  printf("[mv_test %s] %s\n", which.c_str(), g_fail ? "FAIL" : "PASS");
  return g_fail ? 1 : 0;
}
