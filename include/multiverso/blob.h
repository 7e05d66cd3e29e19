// Blob: ref-counted byte buffer (counterpart of include/multiverso/blob.h:13-53).
// Copy = shallow + refcount; construction from external memory copies the bytes.
#ifndef MULTIVERSO_BLOB_H_
#define MULTIVERSO_BLOB_H_
#include <cstddef>
#include <cstring>

namespace multiverso {

class Blob {
 public:
  Blob() : data_(nullptr), size_(0) {}
  explicit Blob(size_t size);
  Blob(const void* data, size_t size);   // copies
  Blob(const Blob& rhs);
  Blob(Blob&& rhs) noexcept : data_(rhs.data_), size_(rhs.size_) { rhs.data_ = nullptr; rhs.size_ = 0; }
  ~Blob();
  Blob& operator=(const Blob& rhs);
  Blob& operator=(Blob&& rhs) noexcept;

  char* data() const { return data_; }
  size_t size() const { return size_; }
  template <typename T>
  size_t size() const { return size_ / sizeof(T); }
  template <typename T>
  T& As(size_t i = 0) const { return reinterpret_cast<T*>(data_)[i]; }
  char& operator[](size_t i) const { return data_[i]; }

 private:
  char* data_;
  size_t size_;
};

}  // namespace multiverso
#endif
