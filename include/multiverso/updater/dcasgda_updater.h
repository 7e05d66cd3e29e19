// DC-ASGD-A: DC-ASGD with an RMS-normalised compensation term.
//   g = delta/lr ; ms = m*ms + (1-m) g*g ;
//   data -= lr * (g + lambda / sqrt(ms + 1e-7) * g*g * (data - shadow[worker])) ; shadow[worker] = data
// (reference: include/multiverso/updater/dcasgd/dcasgda_updater.h:30-46)
#ifndef MULTIVERSO_UPDATER_DCASGDA_UPDATER_H_
#define MULTIVERSO_UPDATER_DCASGDA_UPDATER_H_
#include <cmath>
#include <vector>
#include "multiverso/updater/updater.h"
namespace multiverso {
int MV_NumWorkers();
template <typename T>
class DCASGDAUpdater : public Updater<T> {
 public:
  explicit DCASGDAUpdater(size_t size)
      : size_(size), state_(2 * static_cast<size_t>(MV_NumWorkers() > 0 ? MV_NumWorkers() : 1) * size, T(0)) {}
  void Update(size_t n, T* data, T* delta, AddOption* option, size_t offset) override {
    const T lr = static_cast<T>(option->learning_rate()), lam = static_cast<T>(option->lambda());
    const T m = static_cast<T>(option->momentum());
    const size_t W = state_.size() / (2 * size_);
    T* sh = state_.data() + static_cast<size_t>(option->worker_id()) * size_ + offset;
    T* ms = state_.data() + (W + static_cast<size_t>(option->worker_id())) * size_ + offset;
    T* d = data + offset;
    for (size_t i = 0; i < n; ++i) {
      T g = delta[i] / lr;
      ms[i] = m * ms[i] + (T(1) - m) * g * g;
      d[i] -= lr * (g + lam / std::sqrt(ms[i] + static_cast<T>(1e-7)) * g * g * (d[i] - sh[i]));
      sh[i] = d[i];
    }
  }
  size_t StateBytes() const override { return state_.size() * sizeof(T); }
  void StoreState(char* out) const override { std::memcpy(out, state_.data(), StateBytes()); }
  void LoadState(const char* in) override { std::memcpy(state_.data(), in, StateBytes()); }

 private:
  size_t size_;
  std::vector<T> state_;   // [W shadow slabs][W mean-square slabs]
};
}  // namespace multiverso
#endif
