// Allocators for Blob storage: ref-counted aligned blocks; SmartAllocator pools blocks in
// power-of-two size classes (counterpart of include/multiverso/util/allocator.h:14-61,
// src/util/allocator.cpp). Flags: -allocator_type=smart|plain, -allocator_alignment.
#ifndef MULTIVERSO_UTIL_ALLOCATOR_H_
#define MULTIVERSO_UTIL_ALLOCATOR_H_
#include <atomic>
#include <cstddef>
#include <mutex>
#include <vector>

namespace multiverso {

class Allocator {
 public:
  virtual ~Allocator() = default;
  virtual char* Alloc(size_t size) = 0;   // refcount = 1
  virtual void Free(char* data) = 0;      // --refcount; release at 0
  virtual void Refer(char* data) = 0;     // ++refcount
  static Allocator* Get();
};

// Header placed in front of every user block.
struct BlockHeader {
  std::atomic<int> refs;
  int size_class;          // -1 for the plain allocator
  size_t user_size;
  void* raw;               // pointer returned by malloc
};

class PlainAllocator : public Allocator {
 public:
  char* Alloc(size_t size) override;
  void Free(char* data) override;
  void Refer(char* data) override;
};

class SmartAllocator : public Allocator {
 public:
  ~SmartAllocator() override;
  char* Alloc(size_t size) override;
  void Free(char* data) override;
  void Refer(char* data) override;
  size_t pooled_blocks();

 private:
  static constexpr int kClasses = 48;
  std::mutex mu_[kClasses];
  std::vector<char*> free_[kClasses];
};

}  // namespace multiverso
#endif
