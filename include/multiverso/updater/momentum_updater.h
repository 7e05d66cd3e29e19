// momentum_sgd: s = m*s + (1-m)*delta ; data -= s  -- momentum_updater.h:17-25.
#ifndef MULTIVERSO_UPDATER_MOMENTUM_UPDATER_H_
#define MULTIVERSO_UPDATER_MOMENTUM_UPDATER_H_
#include <vector>
#include "multiverso/updater/updater.h"
namespace multiverso {
template <typename T>
class MomentumUpdater : public Updater<T> {
 public:
  explicit MomentumUpdater(size_t size) : smooth_(size, T(0)) {}
  void Update(size_t n, T* data, T* delta, AddOption* option, size_t offset) override {
    const T m = static_cast<T>(option->momentum());
    T* d = data + offset;
    T* s = smooth_.data() + offset;
    for (size_t i = 0; i < n; ++i) {
      s[i] = m * s[i] + (T(1) - m) * delta[i];
      d[i] -= s[i];
    }
  }
  size_t StateBytes() const override { return smooth_.size() * sizeof(T); }
  void StoreState(char* out) const override { std::memcpy(out, smooth_.data(), StateBytes()); }
  void LoadState(const char* in) override { std::memcpy(smooth_.data(), in, StateBytes()); }

 private:
  std::vector<T> smooth_;
};
}  // namespace multiverso
#endif
