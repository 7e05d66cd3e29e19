// HDFSStream: Stream over libhdfs, bound at run time (counterpart of the reference's
// include/multiverso/io/hdfs_stream.h + src/io/hdfs_stream.cpp:7-154, which need
// -DUSE_HDFS, a JDK and Hadoop at build time and do not compile as shipped).
// Here libhdfs.so is dlopen'ed when the first hdfs:// URI is opened -- $MV_LIBHDFS, then
// "libhdfs.so" / "libhdfs.so.0.0.0" on the loader path -- so the same libmultiverso.so works
// with and without Hadoop; without it, opening an hdfs:// URI fails with a clear message.
#ifndef MULTIVERSO_IO_HDFS_STREAM_H_
#define MULTIVERSO_IO_HDFS_STREAM_H_
#include <string>
#include "multiverso/io/io.h"

namespace multiverso {

class HDFSStream : public Stream {
 public:
  // `fs` / `file` are the libhdfs handles; the stream closes the file, the factory owns fs.
  HDFSStream(void* fs, void* file, const std::string& path, bool writable);
  ~HDFSStream() override;
  void Write(const void* buf, size_t size) override;
  size_t Read(void* buf, size_t size) override;
  bool Good() override { return file_ != nullptr; }
  void Flush() override;

 private:
  void* fs_;
  void* file_;
  std::string path_;
  bool writable_;
};

// One factory (= one hdfsConnect) per "hdfs://host[:port]".
class HDFSStreamFactory : public StreamFactory {
 public:
  explicit HDFSStreamFactory(const std::string& host);
  ~HDFSStreamFactory() override;
  Stream* Open(const URI& uri, FileOpenMode mode) override;
  // true when libhdfs could be loaded in this process
  static bool Available();

 private:
  std::string host_;
  void* fs_ = nullptr;
};

}  // namespace multiverso
#endif
