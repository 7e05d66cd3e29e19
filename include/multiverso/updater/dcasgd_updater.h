// DC-ASGD (https://arxiv.org/abs/1609.08326): delay-compensated async SGD.
//   g = delta/lr ; data -= lr * (g + lambda * g*g * (data - shadow[worker])) ; shadow[worker] = data
// (reference: include/multiverso/updater/dcasgd/dcasgd_updater.h:29-40)
#ifndef MULTIVERSO_UPDATER_DCASGD_UPDATER_H_
#define MULTIVERSO_UPDATER_DCASGD_UPDATER_H_
#include <vector>
#include "multiverso/updater/updater.h"
namespace multiverso {
int MV_NumWorkers();
template <typename T>
class DCASGDUpdater : public Updater<T> {
 public:
  explicit DCASGDUpdater(size_t size)
      : size_(size), shadow_(static_cast<size_t>(MV_NumWorkers() > 0 ? MV_NumWorkers() : 1) * size, T(0)) {}
  void Update(size_t n, T* data, T* delta, AddOption* option, size_t offset) override {
    const T lr = static_cast<T>(option->learning_rate()), lam = static_cast<T>(option->lambda());
    T* sh = shadow_.data() + static_cast<size_t>(option->worker_id()) * size_ + offset;
    T* d = data + offset;
    for (size_t i = 0; i < n; ++i) {
      T g = delta[i] / lr;
      d[i] -= lr * (g + lam * g * g * (d[i] - sh[i]));
      sh[i] = d[i];
    }
  }
  size_t StateBytes() const override { return shadow_.size() * sizeof(T); }
  void StoreState(char* out) const override { std::memcpy(out, shadow_.data(), StateBytes()); }
  void LoadState(const char* in) override { std::memcpy(shadow_.data(), in, StateBytes()); }

 private:
  size_t size_;
  std::vector<T> shadow_;
};
}  // namespace multiverso
#endif
