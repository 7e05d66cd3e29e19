// Allocator implementations (see include/multiverso/util/allocator.h).
#include "multiverso/util/allocator.h"
#include <cstdlib>
#include <string>
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"

namespace multiverso {

MV_DEFINE_int(allocator_alignment, 16, "alignment for the blob allocator");
MV_DEFINE_string(allocator_type, "smart", "use smart (pooling) allocator by default");

namespace {

size_t Alignment() {
  size_t a = static_cast<size_t>(MV_CONFIG(allocator_alignment));
  if (a < sizeof(void*)) a = sizeof(void*);
  size_t p = 1;
  while (p < a) p <<= 1;
  return p;
}

inline BlockHeader* HeaderOf(char* user) {
  return reinterpret_cast<BlockHeader*>(user) - 1;
}

char* RawAlloc(size_t user_size, int size_class) {
  const size_t align = Alignment();
  const size_t total = user_size + sizeof(BlockHeader) + align;
  void* raw = malloc(total);
  if (raw == nullptr) Log::Fatal("allocator: out of memory (%zu bytes)", total);
  uintptr_t p = reinterpret_cast<uintptr_t>(raw) + sizeof(BlockHeader);
  p = (p + align - 1) & ~(uintptr_t)(align - 1);
  char* user = reinterpret_cast<char*>(p);
  BlockHeader* h = HeaderOf(user);
  new (&h->refs) std::atomic<int>(1);
  h->size_class = size_class;
  h->user_size = user_size;
  h->raw = raw;
  return user;
}

}  // namespace

char* PlainAllocator::Alloc(size_t size) { return RawAlloc(size, -1); }
void PlainAllocator::Refer(char* data) { HeaderOf(data)->refs.fetch_add(1, std::memory_order_relaxed); }
void PlainAllocator::Free(char* data) {
  BlockHeader* h = HeaderOf(data);
  if (h->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) free(h->raw);
}

SmartAllocator::~SmartAllocator() {
  for (int c = 0; c < kClasses; ++c)
    for (char* p : free_[c]) free(HeaderOf(p)->raw);
}

char* SmartAllocator::Alloc(size_t size) {
  int cls = 5;   // 32-byte minimum class
  while ((static_cast<size_t>(1) << cls) < size && cls < kClasses - 1) ++cls;
  {
    std::lock_guard<std::mutex> lk(mu_[cls]);
    if (!free_[cls].empty()) {
      char* p = free_[cls].back();
      free_[cls].pop_back();
      HeaderOf(p)->refs.store(1, std::memory_order_relaxed);
      HeaderOf(p)->user_size = size;
      return p;
    }
  }
  return RawAlloc(static_cast<size_t>(1) << cls, cls);
}
void SmartAllocator::Refer(char* data) { HeaderOf(data)->refs.fetch_add(1, std::memory_order_relaxed); }
void SmartAllocator::Free(char* data) {
  BlockHeader* h = HeaderOf(data);
  if (h->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) {
    std::lock_guard<std::mutex> lk(mu_[h->size_class]);
    free_[h->size_class].push_back(data);
  }
}
size_t SmartAllocator::pooled_blocks() {
  size_t n = 0;
  for (int c = 0; c < kClasses; ++c) {
    std::lock_guard<std::mutex> lk(mu_[c]);
    n += free_[c].size();
  }
  return n;
}

Allocator* Allocator::Get() {
  static Allocator* instance = [] {
    Allocator* a;
    if (MV_CONFIG(allocator_type) == "smart") a = new SmartAllocator();
    else a = new PlainAllocator();
    return a;   // leaked: blobs may be released during static destruction
  }();
  return instance;
}

}  // namespace multiverso
