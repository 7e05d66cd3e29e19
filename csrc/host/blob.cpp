#include "multiverso/blob.h"
#include "multiverso/util/parallel_for.h"
#include "multiverso/util/allocator.h"

namespace multiverso {

Blob::Blob(size_t size) : data_(size ? Allocator::Get()->Alloc(size) : nullptr), size_(size) {}
Blob::Blob(const void* data, size_t size)
    : data_(size ? Allocator::Get()->Alloc(size) : nullptr), size_(size) {
  if (size) ParallelMemcpy(data_, data, size);
}
Blob::Blob(const Blob& rhs) : data_(rhs.data_), size_(rhs.size_) {
  if (data_) Allocator::Get()->Refer(data_);
}
Blob::~Blob() {
  if (data_) Allocator::Get()->Free(data_);
}
Blob& Blob::operator=(const Blob& rhs) {
  if (this == &rhs) return *this;
  if (rhs.data_) Allocator::Get()->Refer(rhs.data_);
  if (data_) Allocator::Get()->Free(data_);
  data_ = rhs.data_;
  size_ = rhs.size_;
  return *this;
}
Blob& Blob::operator=(Blob&& rhs) noexcept {
  if (this == &rhs) return *this;
  if (data_) Allocator::Get()->Free(data_);
  data_ = rhs.data_;
  size_ = rhs.size_;
  rhs.data_ = nullptr;
  rhs.size_ = 0;
  return *this;
}

}  // namespace multiverso
