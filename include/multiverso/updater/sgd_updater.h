// sgd: data -= delta (the client pre-multiplies the learning rate) -- sgd_updater.h:14-19.
#ifndef MULTIVERSO_UPDATER_SGD_UPDATER_H_
#define MULTIVERSO_UPDATER_SGD_UPDATER_H_
#include "multiverso/updater/updater.h"
namespace multiverso {
template <typename T>
class SGDUpdater : public Updater<T> {
 public:
  explicit SGDUpdater(size_t) {}
  void Update(size_t n, T* data, T* delta, AddOption*, size_t offset) override {
    T* d = data + offset;
    for (size_t i = 0; i < n; ++i) d[i] -= delta[i];
  }
};
}  // namespace multiverso
#endif
