// Wire compression filters (include/multiverso/util/quantization_util.h:10-161).
// SparseFilter<data_t,index_t>: each data blob becomes (index,value) pairs when fewer than
// half of its elements exceed `clip` in magnitude; an extra size blob records the original
// byte size per blob or -1 when left uncompressed. Blob 0 (keys) and an optional trailing
// option blob pass through. OneBitsFilter is a real 1-bit (sign + two means) quantiser here
// (the reference ships an empty stub).
#ifndef MULTIVERSO_UTIL_QUANTIZATION_UTIL_H_
#define MULTIVERSO_UTIL_QUANTIZATION_UTIL_H_
#include <cmath>
#include <cstdint>
#include <vector>
#include "multiverso/blob.h"

namespace multiverso {

class QuantizationFilter {
 public:
  virtual ~QuantizationFilter() = default;
  virtual void FilterIn(const std::vector<Blob>& in, std::vector<Blob>* out) = 0;
  virtual void FilterOut(const std::vector<Blob>& in, std::vector<Blob>* out) = 0;
};

template <typename data_t, typename index_t>
class SparseFilter : public QuantizationFilter {
 public:
  explicit SparseFilter(double clip, bool skip_option_blob = false)
      : clip_(clip), skip_option_(skip_option_blob) {}

  void FilterIn(const std::vector<Blob>& in, std::vector<Blob>* out) override {
    out->clear();
    if (in.empty()) return;
    const size_t n_data_end = in.size() - (skip_option_ && in.size() > 1 ? 1 : 0);
    out->push_back(in[0]);
    Blob sizes(sizeof(int64_t) * (n_data_end > 1 ? n_data_end - 1 : 0));
    std::vector<Blob> packed;
    for (size_t i = 1; i < n_data_end; ++i) {
      Blob c;
      if (TryCompress(in[i], &c)) {
        sizes.As<int64_t>(i - 1) = static_cast<int64_t>(in[i].size());
        packed.push_back(std::move(c));
      } else {
        sizes.As<int64_t>(i - 1) = -1;
        packed.push_back(in[i]);
      }
    }
    out->push_back(sizes);
    for (auto& b : packed) out->push_back(std::move(b));
    if (n_data_end < in.size()) out->push_back(in.back());
  }

  void FilterOut(const std::vector<Blob>& in, std::vector<Blob>* out) override {
    out->clear();
    if (in.empty()) return;
    out->push_back(in[0]);
    if (in.size() < 2) return;
    const Blob& sizes = in[1];
    const size_t n = sizes.size<int64_t>();
    for (size_t i = 0; i < n; ++i) {
      int64_t orig = sizes.As<int64_t>(i);
      if (orig < 0) out->push_back(in[2 + i]);
      else out->push_back(DeCompress(in[2 + i], static_cast<size_t>(orig)));
    }
    for (size_t i = 2 + n; i < in.size(); ++i) out->push_back(in[i]);
  }

 private:
  bool TryCompress(const Blob& in, Blob* out) const {
    const size_t n = in.size<data_t>();
    size_t nz = 0;
    for (size_t i = 0; i < n; ++i) nz += std::fabs(static_cast<double>(in.As<data_t>(i))) > clip_;
    if (nz * 2 >= n) return false;
    constexpr size_t pair = sizeof(index_t) + sizeof(data_t);
    Blob c(nz * pair);
    char* p = c.data();
    for (size_t i = 0; i < n; ++i) {
      data_t v = in.As<data_t>(i);
      if (std::fabs(static_cast<double>(v)) > clip_) {
        index_t idx = static_cast<index_t>(i);
        std::memcpy(p, &idx, sizeof(index_t));
        std::memcpy(p + sizeof(index_t), &v, sizeof(data_t));
        p += pair;
      }
    }
    *out = std::move(c);
    return true;
  }
  Blob DeCompress(const Blob& in, size_t orig_bytes) const {
    Blob o(orig_bytes);
    std::memset(o.data(), 0, orig_bytes);
    constexpr size_t pair = sizeof(index_t) + sizeof(data_t);
    const size_t cnt = in.size() / pair;
    const char* p = in.data();
    for (size_t k = 0; k < cnt; ++k, p += pair) {
      index_t idx;
      data_t v;
      std::memcpy(&idx, p, sizeof(index_t));
      std::memcpy(&v, p + sizeof(index_t), sizeof(data_t));
      o.As<data_t>(static_cast<size_t>(idx)) = v;
    }
    return o;
  }
  double clip_;
  bool skip_option_;
};

// 1-bit quantisation: sign bitmap + mean of positives + mean of negatives per blob; the
// quantisation error is left to the caller (error feedback) via residual().
template <typename data_t>
class OneBitsFilter : public QuantizationFilter {
 public:
  void FilterIn(const std::vector<Blob>& in, std::vector<Blob>* out) override {
    out->clear();
    if (in.empty()) return;
    out->push_back(in[0]);
    for (size_t i = 1; i < in.size(); ++i) {
      const size_t n = in[i].size<data_t>();
      Blob b(sizeof(int64_t) + 2 * sizeof(data_t) + (n + 7) / 8);
      std::memset(b.data(), 0, b.size());
      double sp = 0, sn = 0;
      size_t np = 0, nn = 0;
      for (size_t k = 0; k < n; ++k) {
        data_t v = in[i].As<data_t>(k);
        if (v >= 0) { sp += v; ++np; } else { sn += v; ++nn; }
      }
      int64_t nn64 = static_cast<int64_t>(n);
      data_t mp = static_cast<data_t>(np ? sp / np : 0), mn = static_cast<data_t>(nn ? sn / nn : 0);
      std::memcpy(b.data(), &nn64, 8);
      std::memcpy(b.data() + 8, &mp, sizeof(data_t));
      std::memcpy(b.data() + 8 + sizeof(data_t), &mn, sizeof(data_t));
      unsigned char* bits = reinterpret_cast<unsigned char*>(b.data() + 8 + 2 * sizeof(data_t));
      for (size_t k = 0; k < n; ++k)
        if (in[i].As<data_t>(k) >= 0) bits[k >> 3] |= static_cast<unsigned char>(1u << (k & 7));
      out->push_back(std::move(b));
    }
  }
  void FilterOut(const std::vector<Blob>& in, std::vector<Blob>* out) override {
    out->clear();
    if (in.empty()) return;
    out->push_back(in[0]);
    for (size_t i = 1; i < in.size(); ++i) {
      int64_t n;
      data_t mp, mn;
      std::memcpy(&n, in[i].data(), 8);
      std::memcpy(&mp, in[i].data() + 8, sizeof(data_t));
      std::memcpy(&mn, in[i].data() + 8 + sizeof(data_t), sizeof(data_t));
      const unsigned char* bits =
          reinterpret_cast<const unsigned char*>(in[i].data() + 8 + 2 * sizeof(data_t));
      Blob o(static_cast<size_t>(n) * sizeof(data_t));
      for (int64_t k = 0; k < n; ++k) o.As<data_t>(k) = (bits[k >> 3] >> (k & 7)) & 1 ? mp : mn;
      out->push_back(std::move(o));
    }
  }
};

}  // namespace multiverso
#endif
