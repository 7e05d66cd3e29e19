// Parameter-server glue of the wordembedding application (reference: Communicator,
// Applications/WordEmbedding/src/communicator.cpp:17-259). Tables, in creation order
// (ids are positional, constant.h:16-20): input embeddings (server-side random init
// U(-0.5/dim, 0.5/dim)), output embeddings (zero), with AdaGrad two more for the accumulated
// squared gradients, and a KVTable<int, int64> holding the global word count.
#ifndef MVAPP_WORDEMBEDDING_PARAM_STORE_H_
#define MVAPP_WORDEMBEDDING_PARAM_STORE_H_
#include <cstdint>
#include <memory>
#include <vector>

#include "data_block.h"
#include "multiverso/table/kv_table.h"
#include "multiverso/table/matrix_table.h"
#include "option.h"

namespace wordembedding {

constexpr int kWordCountKey = 4;

class ParamStore {
 public:
  // Collective: every rank creates the tables in the same order.
  ParamStore(const Option& option, int vocab_size);

  // RequestParameter: pull the rows listed in block->input.ids / output.ids (and their G^2 rows)
  // into the block's caches.
  void Pull(DataBlock* block);
  // AddDeltaParameter: re-read the servers' current rows and push
  // (locally trained - server now) / num_workers for every cached row.
  void PushDelta(DataBlock* block);

  void AddWordCount(int64_t words);
  int64_t GlobalWordCount();

  // Rows [begin, begin + n) of the input-embedding table (SaveEmbedding pulls in batches).
  void GetInputRows(integer_t begin, integer_t n, float* out);

  double pull_seconds() const { return pull_s_; }
  double push_seconds() const { return push_s_; }

 private:
  using Table = multiverso::MatrixWorkerTable<float>;
  void PullRows(Table* t, RowCache* cache, FloatBuffer* dst);
  void PushRows(Table* t, const RowCache& cache, const FloatBuffer& trained);
  const Option& opt_;
  int dim_;
  // the worker halves belong to the creator (the server halves to the runtime)
  std::unique_ptr<Table> input_, output_, input_g2_, output_g2_;
  std::unique_ptr<multiverso::KVWorkerTable<int, int64_t>> word_count_;
  FloatBuffer delta_;            // scratch of PushRows (only the training thread pushes)
  double pull_s_ = 0, push_s_ = 0;
};

}  // namespace wordembedding
#endif
