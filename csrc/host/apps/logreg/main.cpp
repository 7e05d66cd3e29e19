// logreg -- the LogisticRegression application on the multiverso-b200 host runtime. Native
// counterpart of Applications/LogisticRegression (main.cpp:7-12, logreg.cpp:14-173):
//
//   build/bin/logreg <config file> [-mvflag=value ...]
//   python tools/mvrun.py -n 4 -- build/bin/logreg ctr.config
//
// Per epoch: reset the async reader, stream minibatches, Model::Update on each; with use_ps
// the model is pulled every sync_frequency minibatches (blocking or pipelined) and, for sparse
// input, only the keys the next window touches; loss and timing are logged every
// show_time_per_sample samples; Test() after every epoch writes the predictions to output_file
// and logs the test error; finally SaveModel. With several ranks every rank reads train_file
// and keeps the windows i with i % size == rank; output files get the "-<worker_id>" suffix.
// The GPU implementation of the same application is multiverso_b200/apps/logreg.py.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "configure.h"
#include "model.h"
#include "multiverso/apps/app_api.h"
#include "multiverso/multiverso.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/util/timer.h"

namespace multiverso {
MV_DECLARE_bool(sync);
inline bool SyncMode() { return MV_CONFIG(sync); }   // -sync=true: BSP server
}  // namespace multiverso
using multiverso::Log;
using namespace logreg;

namespace {

// Thin RAII wrapper over the native async reader.
class SampleReader {
 public:
  SampleReader(const Configure& c, const std::string& files)
      : handle_(MVA_LRReaderOpen(files.c_str(), c.reader_type.c_str(), c.sparse ? 1 : 0, c.input_size,
                                 c.read_buffer_size * 3)),
        max_nnz_(static_cast<int64_t>(c.minibatch_size) * std::min<int64_t>(c.input_size + 1, 4096)) {}
  ~SampleReader() { MVA_LRReaderClose(handle_); }
  void Reset() { MVA_LRReaderReset(handle_); }
  // false at the end of the epoch
  bool Next(int64_t max_samples, MiniBatch* b) {
    for (;;) {
      b->Reserve(max_samples, max_nnz_);
      const int64_t n = MVA_LRReaderNext(handle_, max_samples, max_nnz_, b->row_ptr.data(), b->keys.data(),
                                         b->vals.data(), b->labels.data(), b->weights.data());
      if (n < 0) {                          // one sample is wider than the buffer
        max_nnz_ = std::max(2 * max_nnz_, -n);
        continue;
      }
      b->n = n;
      return n > 0;
    }
  }

 private:
  void* handle_;
  int64_t max_nnz_;
};

// sync_frequency consecutive minibatches and the distinct feature keys they touch.
struct Window {
  std::vector<MiniBatch> batches;
  std::vector<int64_t> keys;
  bool empty() const { return batches.empty(); }
};

bool ReadWindow(SampleReader* reader, const Configure& cfg, bool need_keys, Window* w) {
  w->batches.clear();
  w->keys.clear();
  for (int i = 0; i < cfg.sync_frequency; ++i) {
    MiniBatch b;
    if (!reader->Next(cfg.minibatch_size, &b)) break;
    w->batches.push_back(std::move(b));
  }
  if (need_keys) {
    for (const MiniBatch& b : w->batches) w->keys.insert(w->keys.end(), b.keys.begin(), b.keys.begin() + b.row_ptr[b.n]);
    std::sort(w->keys.begin(), w->keys.end());
    w->keys.erase(std::unique(w->keys.begin(), w->keys.end()), w->keys.end());
  }
  return !w->empty();
}

// Next window that belongs to this rank (windows are dealt round-robin). In BSP mode every worker
// must issue the same number of table operations, so this rank's window of a round is only released
// once the whole round (one window per rank) has been read; an incomplete last round is dropped.
bool ReadMyWindow(SampleReader* reader, const Configure& cfg, bool need_keys, int rank, int size, bool full_rounds_only,
                  int64_t* index, Window* w) {
  Window scratch, mine;
  bool have_mine = false;
  while (ReadWindow(reader, cfg, need_keys, &scratch)) {
    const int64_t i = (*index)++;
    if (i % size == rank) {
      std::swap(mine, scratch);
      have_mine = true;
      if (!full_rounds_only) break;
    }
    if (full_rounds_only && (i + 1) % size == 0 && have_mine) break;
    if (full_rounds_only && (i + 1) % size == 0) have_mine = false;
  }
  if (have_mine && full_rounds_only && *index % size != 0) have_mine = false;   // the round never completed
  if (have_mine) std::swap(*w, mine);
  return have_mine;
}

double Test(const Configure& cfg, Model* model, int rank, int size) {
  if (cfg.test_file.empty()) return 0.0;
  model->PullAll();
  SampleReader reader(cfg, cfg.test_file);
  const std::string path = cfg.output_file + (size > 1 ? "-" + std::to_string(std::max(0, multiverso::MV_WorkerId())) : "");
  FILE* out = cfg.output_file.empty() ? nullptr : fopen(path.c_str(), "w");
  (void)rank;
  MiniBatch b;
  std::vector<float> pred;
  int64_t total = 0, correct = 0;
  while (reader.Next(std::max(cfg.minibatch_size, 256), &b)) {
    const BatchResult r = model->Predict(b, &pred);
    total += b.n;
    correct += r.correct;
    if (out != nullptr)
      for (int64_t i = 0; i < b.n; ++i) {
        for (int c = 0; c < model->out(); ++c) fprintf(out, c ? " %g" : "%g", pred[i * model->out() + c]);
        fputc('\n', out);
      }
  }
  if (out != nullptr) fclose(out);
  const double err = 1.0 - static_cast<double>(correct) / std::max<int64_t>(1, total);
  Log::Info("test error: %f (%lld samples)\n", err, static_cast<long long>(total));
  return err;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 2 || argv[1][0] == '-') {
    puts("usage: logreg <config file> [-mvflag=value ...]");
    return 2;
  }
  Configure cfg;
  if (!cfg.Load(argv[1])) return 2;
  // PSModel: the server applies w -= delta (ps_model.cpp:12-20); an explicit -updater_type= wins
  if (cfg.use_ps) multiverso::MV_SetFlag<std::string>("updater_type", "sgd");
  std::vector<char*> mv_args{argv[0]};
  for (int i = 2; i < argc; ++i)
    if (argv[i][0] == '-' && strchr(argv[i], '=') != nullptr) mv_args.push_back(argv[i]);
  int mv_argc = static_cast<int>(mv_args.size());
  multiverso::MV_Init(&mv_argc, mv_args.data());
  const int rank = multiverso::MV_Rank(), size = multiverso::MV_Size();

  std::unique_ptr<Model> model = Model::Create(cfg);
  if (!cfg.init_model_file.empty()) model->Load(cfg.init_model_file);
  const bool need_keys = cfg.use_ps && (cfg.sparse || cfg.ftrl());
  const bool full_rounds_only = cfg.use_ps && size > 1 && multiverso::SyncMode();
  if (need_keys && size > 1 && multiverso::SyncMode())
    Log::Fatal("logreg: the BSP server (-sync=true) needs every worker to address every server in every step; the sparse "
               "parameter-server tables only address the owners of the keys a minibatch touches. Run sparse / FTRL "
               "models in async mode (the reference's mode for them), or dense ones in BSP.\n");
  if (rank == 0)
    Log::Info("logreg: %lld inputs (+bias), %d outputs, objective %s, regular %s, updater %s, %s input, %s, "
              "minibatch %d, %d epoch(s)\n",
              static_cast<long long>(cfg.input_size), cfg.output_size, cfg.objective_type.c_str(),
              cfg.regular_type.c_str(), cfg.updater_type.c_str(), cfg.sparse ? "sparse" : "dense",
              cfg.use_ps ? (cfg.pipeline ? "parameter server (pipelined pulls)" : "parameter server") : "local model",
              cfg.minibatch_size, cfg.train_epoch);

  multiverso::Timer wall;
  SampleReader reader(cfg, cfg.train_file);
  int64_t total_samples = 0;
  double last_loss = 0, last_acc = 0, test_error = 0;
  std::string epoch_losses;
  for (int epoch = 0; epoch < cfg.train_epoch; ++epoch) {
    if (epoch > 0) reader.Reset();
    multiverso::Timer epoch_timer;
    int64_t seen = 0, shown = 0, correct = 0, window_index = 0;
    double loss = 0;
    Window cur, next;
    bool have = ReadMyWindow(&reader, cfg, need_keys, rank, size, full_rounds_only, &window_index, &cur);
    while (have) {
      const bool have_next = ReadMyWindow(&reader, cfg, need_keys, rank, size, full_rounds_only, &window_index, &next);
      model->BeginWindow(cur.keys, have_next ? &next.keys : nullptr);
      for (const MiniBatch& b : cur.batches) {
        const BatchResult r = model->Update(b);
        loss += r.loss;
        correct += r.correct;
        seen += b.n;
        if (seen - shown >= cfg.show_time_per_sample) {
          shown = seen;
          Log::Info("Sample seen %lld  train loss %.6f  learning rate %.5f  (%.0f samples/s)\n",
                    static_cast<long long>(seen), loss / seen, model->learning_rate(),
                    seen / std::max(epoch_timer.elapse() * 1e-3, 1e-9));
          model->LogTimes();
        }
      }
      std::swap(cur, next);
      have = have_next;
    }
    multiverso::MV_Barrier();
    last_loss = loss / std::max<int64_t>(1, seen);
    last_acc = static_cast<double>(correct) / std::max<int64_t>(1, seen);
    total_samples += seen;
    Log::Info("epoch %d: %lld samples in %.2fs, train loss %.6f, train accuracy %.4f\n", epoch,
              static_cast<long long>(seen), epoch_timer.elapse() * 1e-3, last_loss, last_acc);
    char buf[48];
    snprintf(buf, sizeof buf, "%s%.6f", epoch ? ", " : "", last_loss);
    epoch_losses += buf;
    test_error = Test(cfg, model.get(), rank, size);
  }
  const double seconds = wall.elapse() * 1e-3;
  if (!cfg.output_model_file.empty()) model->Store(cfg.output_model_file);
  printf("{\"app\": \"logreg\", \"rank\": %d, \"ranks\": %d, \"samples\": %lld, \"seconds\": %.3f, "
         "\"samples_per_sec\": %.1f, \"train_loss\": %.6f, \"train_acc\": %.4f, \"test_error\": %.6f, "
         "\"epoch_loss\": [%s]}\n",
         rank, size, static_cast<long long>(total_samples), seconds, total_samples / std::max(seconds, 1e-9), last_loss,
         last_acc, test_error, epoch_losses.c_str());
  fflush(stdout);
  multiverso::MV_Barrier();
  model.reset();                            // worker tables go before the runtime
  multiverso::MV_ShutDown();
  return 0;
}
