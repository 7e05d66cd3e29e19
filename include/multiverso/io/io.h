// IO streams (counterpart of include/multiverso/io/io.h:24-132): URI, Stream,
// StreamFactory::GetStream(uri, mode), TextReader::GetLine. "file" scheme = LocalStream;
// "hdfs" = HDFSStream over a run-time loaded libhdfs (io/hdfs_stream.h).
#ifndef MULTIVERSO_IO_IO_H_
#define MULTIVERSO_IO_IO_H_
#include <cstddef>
#include <map>
#include <memory>
#include <string>

namespace multiverso {

enum class FileOpenMode : int { Write = 0, Read = 1, Append = 2, BinaryWrite = 3, BinaryRead = 4, BinaryAppend = 5 };

struct URI {
  std::string scheme;   // default "file"
  std::string host;
  std::string name;     // path
  std::string path;     // the full original string
  URI() = default;
  explicit URI(const std::string& uri);
};

class Stream {
 public:
  virtual ~Stream() = default;
  virtual void Write(const void* buf, size_t size) = 0;
  virtual size_t Read(void* buf, size_t size) = 0;
  virtual bool Good() = 0;
  virtual void Flush() {}
  // set by a reader that found the content incomplete (e.g. a truncated table checkpoint)
  void MarkFailed() { failed_ = true; }
  bool Failed() const { return failed_; }

 private:
  bool failed_ = false;
};

class StreamFactory {
 public:
  // One factory is cached per "scheme://host" (io.cpp:8-60).
  static Stream* GetStream(const URI& uri, FileOpenMode mode);
  virtual ~StreamFactory() = default;
  virtual Stream* Open(const URI& uri, FileOpenMode mode) = 0;

 private:
  static std::map<std::string, std::unique_ptr<StreamFactory>>& instances();
};

class TextReader {
 public:
  TextReader(const URI& uri, size_t buf_size = 1 << 20);
  ~TextReader();
  // false at end of file; strips the trailing newline.
  bool GetLine(std::string& line);
  bool Good() const { return stream_ != nullptr; }

 private:
  size_t Fill();
  Stream* stream_;
  char* buf_;
  size_t buf_size_, pos_ = 0, length_ = 0;
};

}  // namespace multiverso
#endif
