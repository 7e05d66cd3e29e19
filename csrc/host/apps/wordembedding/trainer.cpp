#include "trainer.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstring>

#include "multiverso/apps/app_api.h"
#include "multiverso/util/log.h"

namespace wordembedding {

namespace {
constexpr int kMaxCodeLength = 64;
constexpr float kAdaGradEps = 1e-6f;

inline float Dot(const float* a, const float* b, int n) {
  float s = 0;
#pragma omp simd reduction(+ : s)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
inline void Axpy(float alpha, const float* x, float* y, int n) {
#pragma omp simd
  for (int i = 0; i < n; ++i) y[i] += alpha * x[i];
}
inline float Sigmoid(float x) { return 1.0f / (1.0f + std::exp(-x)); }
}  // namespace

// ------------------------------------------------------------------------------ Vocabulary
void Vocabulary::BuildHuffman() {
  max_code = kMaxCodeLength;
  points.assign(static_cast<size_t>(size) * max_code, 0);
  codes.assign(static_cast<size_t>(size) * max_code, 0);
  code_len.assign(size, 0);
  std::vector<int64_t> f(freq);
  for (auto& v : f) v = std::max<int64_t>(v, 1);
  const int longest = MVA_HuffmanBuild(f.data(), size, max_code, points.data(), codes.data(), code_len.data());
  if (longest < 0) multiverso::Log::Fatal("Huffman code longer than %d\n", max_code);
}

void Vocabulary::BuildNegativeSampler() {
  // Vose's alias method over w_i = freq_i^0.75
  const int n = size;
  alias_prob.assign(n, 1.0f);
  alias_other.resize(n);
  std::vector<double> scaled(n);
  double sum = 0;
  for (int i = 0; i < n; ++i) sum += (scaled[i] = std::pow(static_cast<double>(std::max<int64_t>(freq[i], 1)), 0.75));
  std::vector<int32_t> small, large;
  for (int i = 0; i < n; ++i) {
    scaled[i] *= n / sum;
    alias_other[i] = i;
    (scaled[i] < 1.0 ? small : large).push_back(i);
  }
  while (!small.empty() && !large.empty()) {
    const int32_t s = small.back(), l = large.back();
    small.pop_back();
    alias_prob[s] = static_cast<float>(scaled[s]);
    alias_other[s] = l;
    scaled[l] -= 1.0 - scaled[s];
    if (scaled[l] < 1.0) {
      large.pop_back();
      small.push_back(l);
    }
  }
}

int32_t Vocabulary::DrawNegative(uint64_t* rng) const {
  const uint64_t r = NextRandom(rng);
  const int32_t i = static_cast<int32_t>(r % static_cast<uint64_t>(size));
  const float u = static_cast<float>((r >> 24) & 0xFFFFFF) / 16777216.0f;
  return u < alias_prob[i] ? i : alias_other[i];
}

// --------------------------------------------------------------------------------- Prepare
void Trainer::Prepare(DataBlock* b, uint64_t seed) const {
  const int V = vocab_.size;
  b->IndexSentences();
  std::vector<int32_t> slot(V, -1);          // word / node id -> slot, reused for both caches

  // input rows: every distinct word of the block
  std::vector<integer_t> in_ids;
  for (int32_t t : b->tokens)
    if (t >= 0 && slot[t] < 0) {
      slot[t] = 0;
      in_ids.push_back(t);
    }
  std::sort(in_ids.begin(), in_ids.end());
  for (size_t i = 0; i < in_ids.size(); ++i) slot[in_ids[i]] = static_cast<int32_t>(i);
  b->in_slot.resize(b->tokens.size());
  for (size_t i = 0; i < b->tokens.size(); ++i) b->in_slot[i] = b->tokens[i] >= 0 ? slot[b->tokens[i]] : -1;
  for (integer_t id : in_ids) slot[id] = -1;
  b->input.ids = in_ids;

  std::vector<integer_t> out_ids;
  b->negative_pool.clear();
  b->negative_pool_ids.clear();
  b->path_begin.clear();
  b->path_slot.clear();
  b->path_code.clear();
  b->out_slot.clear();
  if (opt_.hs) {
    // output rows: the Huffman inner nodes on the paths of the block's words
    for (integer_t w : in_ids)
      for (int d = 0; d < vocab_.code_len[w]; ++d) {
        const int32_t node = vocab_.points[static_cast<size_t>(w) * vocab_.max_code + d];
        if (slot[node] < 0) {
          slot[node] = 0;
          out_ids.push_back(node);
        }
      }
    std::sort(out_ids.begin(), out_ids.end());
    for (size_t i = 0; i < out_ids.size(); ++i) slot[out_ids[i]] = static_cast<int32_t>(i);
    b->path_begin.reserve(in_ids.size() + 1);
    for (integer_t w : in_ids) {
      b->path_begin.push_back(static_cast<int32_t>(b->path_slot.size()));
      for (int d = 0; d < vocab_.code_len[w]; ++d) {
        const size_t k = static_cast<size_t>(w) * vocab_.max_code + d;
        b->path_slot.push_back(slot[vocab_.points[k]]);
        b->path_code.push_back(vocab_.codes[k]);
      }
    }
    b->path_begin.push_back(static_cast<int32_t>(b->path_slot.size()));
  } else {
    // output rows: the block's words plus `negative x |words|` draws from the unigram^0.75
    // distribution; the draws also form the pool negatives are taken from while training
    uint64_t rng = seed * 2862933555777941757ULL + 3037000493ULL;
    std::vector<int32_t> pool_ids(static_cast<size_t>(opt_.negative_num) * in_ids.size());
    for (auto& p : pool_ids) p = vocab_.DrawNegative(&rng);
    for (integer_t w : in_ids) {
      slot[w] = 0;
      out_ids.push_back(w);
    }
    for (int32_t p : pool_ids)
      if (slot[p] < 0) {
        slot[p] = 0;
        out_ids.push_back(p);
      }
    std::sort(out_ids.begin(), out_ids.end());
    for (size_t i = 0; i < out_ids.size(); ++i) slot[out_ids[i]] = static_cast<int32_t>(i);
    b->out_slot.resize(b->tokens.size());
    for (size_t i = 0; i < b->tokens.size(); ++i) b->out_slot[i] = b->tokens[i] >= 0 ? slot[b->tokens[i]] : -1;
    b->negative_pool.resize(pool_ids.size());
    for (size_t i = 0; i < pool_ids.size(); ++i) b->negative_pool[i] = slot[pool_ids[i]];
    b->negative_pool_ids = std::move(pool_ids);
  }
  b->output.ids = std::move(out_ids);
}

// ----------------------------------------------------------------------------------- Train
struct Trainer::Scratch {
  std::vector<float> hidden, hidden_err;
  std::vector<int32_t> context;
};

float Trainer::LearningRate(double words_seen) const {
  const double lr0 = opt_.init_learning_rate;
  const double denom = static_cast<double>(opt_.total_words) * opt_.epoch + 1.0;
  return static_cast<float>(std::max(lr0 * 1e-4, lr0 * (1.0 - words_seen / denom)));
}

TrainStats Trainer::Train(DataBlock* b, int64_t words_before, int num_workers, int threads) const {
  TrainStats total;
  const int nsent = static_cast<int>(b->sentences.size());
  if (nsent == 0) return total;
  threads = std::max(1, std::min(threads, nsent));
  const double block_words = static_cast<double>(b->corpus_words) * std::max(1, num_workers);
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
    const int tid = 0, nth = 1;
#endif
    Scratch s;
    s.hidden.resize(opt_.embeding_size);
    s.hidden_err.resize(opt_.embeding_size);
    s.context.reserve(2 * opt_.window_size + 1);
    uint64_t rng = (static_cast<uint64_t>(words_before) + 1) * 0x9E3779B97F4A7C15ULL + tid * 7919 + 1;
    TrainStats st;
    for (int i = tid; i < nsent; i += nth) {
      const float lr = LearningRate(words_before + block_words * i / nsent);
      TrainSentence(b, b->sentences[i].first, b->sentences[i].second, lr, &rng, &s, &st);
    }
#pragma omp critical
    {
      total.loss += st.loss;
      total.terms += st.terms;
      total.samples += st.samples;
      total.words += st.words;
    }
  }
  return total;
}

void Trainer::TrainSentence(DataBlock* b, int32_t begin, int32_t end, float lr, uint64_t* rng, Scratch* s,
                            TrainStats* st) const {
  const int window = opt_.window_size;
  for (int32_t pos = begin; pos < end; ++pos) {
    ++st->words;
    const int shrink = static_cast<int>(NextRandom(rng) % static_cast<uint64_t>(window));
    const int32_t lo = std::max(begin, pos - window + shrink);
    const int32_t hi = std::min(end - 1, pos + window - shrink);
    if (opt_.cbow) {
      s->context.clear();
      for (int32_t c = lo; c <= hi; ++c)
        if (c != pos) s->context.push_back(b->in_slot[c]);
      if (!s->context.empty())
        TrainSample(b, s->context.data(), static_cast<int>(s->context.size()), pos, lr, rng, s, st);
    } else {
      for (int32_t c = lo; c <= hi; ++c)
        if (c != pos) TrainSample(b, &b->in_slot[c], 1, pos, lr, rng, s, st);
    }
  }
}

void Trainer::TrainSample(DataBlock* b, const int32_t* inputs, int n_inputs, int32_t center_pos, float lr,
                          uint64_t* rng, Scratch* s, TrainStats* st) const {
  const int D = opt_.embeding_size;
  const bool adagrad = opt_.use_adagrad;
  const float lr0 = opt_.init_learning_rate;
  float* h = s->hidden.data();
  float* err = s->hidden_err.data();
  float* in_rows = b->input.rows.data();
  float* out_rows = b->output.rows.data();

  // feed forward: hidden = mean of the input rows
  if (n_inputs == 1) {
    std::memcpy(h, in_rows + static_cast<size_t>(inputs[0]) * D, sizeof(float) * D);
  } else {
    std::memset(h, 0, sizeof(float) * D);
    for (int i = 0; i < n_inputs; ++i) Axpy(1.0f, in_rows + static_cast<size_t>(inputs[i]) * D, h, D);
    const float inv = 1.0f / n_inputs;
    for (int j = 0; j < D; ++j) h[j] *= inv;
  }
  std::memset(err, 0, sizeof(float) * D);

  // one logistic unit per (output row, label)
  auto unit = [&](int32_t out_slot, float label) {
    float* w = out_rows + static_cast<size_t>(out_slot) * D;
    const float p = Sigmoid(Dot(h, w, D));
    const float g = label - p;
    st->loss -= std::log(std::max(label > 0.5f ? p : 1.0f - p, 1e-7f));
    ++st->terms;
    if (!adagrad) {
      const float gl = g * lr;
      Axpy(gl, w, err, D);
      Axpy(gl, h, w, D);
    } else {
      float* G = b->output.g2.data() + static_cast<size_t>(out_slot) * D;
      for (int j = 0; j < D; ++j) {
        const float grad = g * h[j];
        err[j] += g * w[j];
        G[j] += grad * grad;
        w[j] += lr0 * grad / std::sqrt(G[j] + kAdaGradEps);
      }
    }
  };

  if (opt_.hs) {
    const int32_t word_slot = b->in_slot[center_pos];
    for (int32_t d = b->path_begin[word_slot]; d < b->path_begin[word_slot + 1]; ++d)
      unit(b->path_slot[d], 1.0f - static_cast<float>(b->path_code[d]));
  } else {
    const int32_t target = b->out_slot[center_pos];
    unit(target, 1.0f);
    const uint64_t pool = b->negative_pool.size();
    for (int k = 0; k < opt_.negative_num && pool > 0; ++k) {
      const int32_t neg = b->negative_pool[NextRandom(rng) % pool];
      if (neg == target) continue;
      unit(neg, 0.0f);
    }
  }

  // back-propagate the hidden error into every input row
  for (int i = 0; i < n_inputs; ++i) {
    float* w = in_rows + static_cast<size_t>(inputs[i]) * D;
    if (!adagrad) {
      Axpy(1.0f, err, w, D);
    } else {
      float* G = b->input.g2.data() + static_cast<size_t>(inputs[i]) * D;
      for (int j = 0; j < D; ++j) {
        G[j] += err[j] * err[j];
        w[j] += lr0 * err[j] / std::sqrt(G[j] + kAdaGradEps);
      }
    }
  }
  ++st->samples;
}

}  // namespace wordembedding
