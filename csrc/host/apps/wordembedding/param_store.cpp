#include "param_store.h"

#include "multiverso/multiverso.h"
#include "multiverso/util/log.h"
#include "multiverso/util/timer.h"

namespace wordembedding {

using namespace multiverso;

ParamStore::ParamStore(const Option& option, int vocab_size) : opt_(option), dim_(option.embeding_size) {
  const float r = 0.5f / dim_;
  input_.reset(MV_CreateTable(MatrixTableOption<float>(vocab_size, dim_, -r, r)));
  output_.reset(MV_CreateTable(MatrixTableOption<float>(vocab_size, dim_)));
  if (opt_.use_adagrad) {
    input_g2_.reset(MV_CreateTable(MatrixTableOption<float>(vocab_size, dim_)));
    output_g2_.reset(MV_CreateTable(MatrixTableOption<float>(vocab_size, dim_)));
  }
  word_count_.reset(MV_CreateTable(KVTableOption<int, int64_t>()));
  if (input_ == nullptr || word_count_ == nullptr)
    Log::Fatal("wordembedding: every rank must be a worker (run with the default -ps_role)\n");
}

void ParamStore::PullRows(Table* t, RowCache* cache, FloatBuffer* dst) {
  dst->Allocate(cache->size() * dim_);
  if (cache->size() == 0) return;
  t->Get(dst->data(), dst->size(), cache->ids.data(), static_cast<int>(cache->size()));
}

void ParamStore::Pull(DataBlock* b) {
  Timer timer;
  PullRows(input_.get(), &b->input, &b->input.rows);
  PullRows(output_.get(), &b->output, &b->output.rows);
  if (opt_.use_adagrad) {
    PullRows(input_g2_.get(), &b->input, &b->input.g2);
    PullRows(output_g2_.get(), &b->output, &b->output.g2);
  }
  pull_s_ += timer.elapse() * 1e-3;
}

void ParamStore::PushRows(Table* t, const RowCache& cache, const FloatBuffer& trained) {
  if (cache.size() == 0) return;
  delta_.Allocate(trained.size());
  float* delta = delta_.data();
  const float* mine = trained.data();
  std::vector<integer_t> ids(cache.ids);
  t->Get(delta, delta_.size(), ids.data(), static_cast<int>(ids.size()));   // server now
  const float inv = 1.0f / MV_NumWorkers();
  const long long n = static_cast<long long>(delta_.size());
#pragma omp parallel for schedule(static) num_threads(opt_.thread_cnt) if (n > (1 << 16))
  for (long long i = 0; i < n; ++i) delta[i] = (mine[i] - delta[i]) * inv;
  t->Add(delta, delta_.size(), ids.data(), static_cast<int>(ids.size()));
}

void ParamStore::PushDelta(DataBlock* b) {
  Timer timer;
  PushRows(input_.get(), b->input, b->input.rows);
  PushRows(output_.get(), b->output, b->output.rows);
  if (opt_.use_adagrad) {
    PushRows(input_g2_.get(), b->input, b->input.g2);
    PushRows(output_g2_.get(), b->output, b->output.g2);
  }
  push_s_ += timer.elapse() * 1e-3;
}

void ParamStore::AddWordCount(int64_t words) { word_count_->Add(kWordCountKey, words); }

int64_t ParamStore::GlobalWordCount() {
  word_count_->Get(kWordCountKey);
  return word_count_->raw()[kWordCountKey];
}

void ParamStore::GetInputRows(integer_t begin, integer_t n, float* out) {
  std::vector<integer_t> ids(n);
  for (integer_t i = 0; i < n; ++i) ids[i] = begin + i;
  input_->Get(out, static_cast<size_t>(n) * dim_, ids.data(), static_cast<int>(n));
}

}  // namespace wordembedding
