// adagrad: per-worker history G2 += (delta/lr)^2 ; data -= rho / sqrt(G2 + eps) * delta/lr.
// This is the INTENDED rule: the reference copies the history vector by value and subtracts
// the squares (adagrad_updater.h:26-35), so its history never persists.
#ifndef MULTIVERSO_UPDATER_ADAGRAD_UPDATER_H_
#define MULTIVERSO_UPDATER_ADAGRAD_UPDATER_H_
#include <cmath>
#include <vector>
#include "multiverso/updater/updater.h"
namespace multiverso {
int MV_NumWorkers();
template <typename T>
class AdaGradUpdater : public Updater<T> {
 public:
  explicit AdaGradUpdater(size_t size)
      : size_(size), g2_(static_cast<size_t>(MV_NumWorkers() > 0 ? MV_NumWorkers() : 1) * size, T(0)) {}
  void Update(size_t n, T* data, T* delta, AddOption* option, size_t offset) override {
    const T lr = static_cast<T>(option->learning_rate()), rho = static_cast<T>(option->rho());
    T* h = g2_.data() + static_cast<size_t>(option->worker_id()) * size_ + offset;
    T* d = data + offset;
    for (size_t i = 0; i < n; ++i) {
      T g = delta[i] / lr;
      h[i] += g * g;
      d[i] -= rho / std::sqrt(h[i] + static_cast<T>(1e-6)) * g;
    }
  }
  size_t StateBytes() const override { return g2_.size() * sizeof(T); }
  void StoreState(char* out) const override { std::memcpy(out, g2_.data(), StateBytes()); }
  void LoadState(const char* in) override { std::memcpy(g2_.data(), in, StateBytes()); }

 private:
  size_t size_;
  std::vector<T> g2_;
};
}  // namespace multiverso
#endif
