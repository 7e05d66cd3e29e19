// LocalStream: fopen/fread/fwrite (include/multiverso/io/local_stream.h, src/io/local_stream.cpp).
#ifndef MULTIVERSO_IO_LOCAL_STREAM_H_
#define MULTIVERSO_IO_LOCAL_STREAM_H_
#include <cstdio>
#include "multiverso/io/io.h"

namespace multiverso {

class LocalStream : public Stream {
 public:
  LocalStream(const URI& uri, FileOpenMode mode);
  ~LocalStream() override;
  void Write(const void* buf, size_t size) override;
  size_t Read(void* buf, size_t size) override;
  bool Good() override { return fp_ != nullptr; }
  void Flush() override;

 private:
  FILE* fp_ = nullptr;
  std::string path_;
};

class LocalStreamFactory : public StreamFactory {
 public:
  Stream* Open(const URI& uri, FileOpenMode mode) override { return new LocalStream(uri, mode); }
};

}  // namespace multiverso
#endif
