#include "multiverso/io/local_stream.h"
#include "multiverso/util/log.h"

namespace multiverso {

LocalStream::LocalStream(const URI& uri, FileOpenMode mode) : path_(uri.name) {
  static const char* kModes[] = {"w", "r", "a", "wb", "rb", "ab"};
  fp_ = fopen(path_.c_str(), kModes[static_cast<int>(mode)]);
  if (!fp_) Log::Error("LocalStream: cannot open '%s'", path_.c_str());
}
LocalStream::~LocalStream() {
  if (fp_) fclose(fp_);
}
void LocalStream::Write(const void* buf, size_t size) {
  if (!fp_ || fwrite(buf, 1, size, fp_) != size) Log::Error("LocalStream: write to '%s' failed", path_.c_str());
}
size_t LocalStream::Read(void* buf, size_t size) { return fp_ ? fread(buf, 1, size, fp_) : 0; }
void LocalStream::Flush() {
  if (fp_) fflush(fp_);
}

}  // namespace multiverso
