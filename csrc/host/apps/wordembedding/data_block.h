// One unit of work of the trainer: a slice of the corpus plus a private, contiguous copy of
// every parameter row the slice can touch (reference: DataBlock, data_block.cpp, with one
// heap row per word looked up through two vocabulary-sized pointer arrays; here two flat
// [slots x dim] arrays, and the corpus slice is translated to slot numbers once per block so
// the training loops never hash).
#ifndef MVAPP_WORDEMBEDDING_DATA_BLOCK_H_
#define MVAPP_WORDEMBEDDING_DATA_BLOCK_H_
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "multiverso/table_interface.h"

namespace wordembedding {

using multiverso::integer_t;

// Float array that is NOT value-initialised on allocation: row caches are hundreds of MB per
// block and are overwritten in full by the pull that follows, so the zero fill (and its page
// faults on the calling thread) would be pure overhead.
class FloatBuffer {
 public:
  void Allocate(size_t n) {
    if (n > capacity_) {
      data_.reset(new float[n]);
      capacity_ = n;
    }
    size_ = n;
  }
  float* data() { return data_.get(); }
  const float* data() const { return data_.get(); }
  size_t size() const { return size_; }

 private:
  std::unique_ptr<float[]> data_;
  size_t size_ = 0, capacity_ = 0;
};

// A set of table rows cached locally: sorted unique row ids, the values as pulled (trained in
// place) and, with AdaGrad, the accumulated squared gradients.
struct RowCache {
  std::vector<integer_t> ids;
  FloatBuffer rows;             // ids.size() x dim
  FloatBuffer g2;               // ids.size() x dim when AdaGrad, else empty
  size_t size() const { return ids.size(); }
};

struct DataBlock {
  std::vector<int32_t> tokens;                          // word ids, -1 between sentences
  std::vector<std::pair<int32_t, int32_t>> sentences;   // [begin, end) into tokens
  int64_t corpus_words = 0;                             // words read from the corpus (incl. dropped)
  int epoch = 0;

  RowCache input, output;
  std::vector<int32_t> in_slot;          // per token: slot in `input` (-1 at separators)
  std::vector<int32_t> out_slot;         // per token: slot in `output` (negative sampling)
  std::vector<int32_t> negative_pool;    // the block's negative samples, as slots of `output`
  std::vector<int32_t> negative_pool_ids;   // the same draws as word ids (the GPU kernel maps ids itself)
  // hierarchical softmax: per input slot the Huffman path as slots of `output` + branch codes
  std::vector<int32_t> path_begin;       // input.size() + 1
  std::vector<int32_t> path_slot;
  std::vector<int8_t> path_code;

  void IndexSentences() {
    sentences.clear();
    int32_t begin = 0;
    const int32_t n = static_cast<int32_t>(tokens.size());
    for (int32_t i = 0; i <= n; ++i) {
      if (i == n || tokens[i] < 0) {
        if (i - begin >= 2) sentences.emplace_back(begin, i);
        begin = i + 1;
      }
    }
  }
};

}  // namespace wordembedding
#endif
