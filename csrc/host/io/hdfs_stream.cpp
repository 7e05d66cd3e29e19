#include "multiverso/io/hdfs_stream.h"

#include <dlfcn.h>
#include <fcntl.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "multiverso/util/log.h"

namespace multiverso {

namespace {

// The slice of the libhdfs C API the stream needs (hdfs.h: hdfsFS / hdfsFile are opaque
// pointers, tSize = int32_t, tPort = uint16_t).
struct LibHdfs {
  void* (*Connect)(const char* namenode, uint16_t port) = nullptr;
  int (*Disconnect)(void* fs) = nullptr;
  void* (*OpenFile)(void* fs, const char* path, int flags, int buffer_size, short replication,
                    int32_t block_size) = nullptr;
  int (*CloseFile)(void* fs, void* file) = nullptr;
  int32_t (*Read)(void* fs, void* file, void* buffer, int32_t length) = nullptr;
  int32_t (*Write)(void* fs, void* file, const void* buffer, int32_t length) = nullptr;
  int (*Flush)(void* fs, void* file) = nullptr;
  bool ok = false;
};

const LibHdfs& Lib() {
  static LibHdfs lib;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    if (const char* path = getenv("MV_LIBHDFS")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    for (const char* name : {"libhdfs.so", "libhdfs.so.0.0.0", "libhdfs3.so"})
      if (h == nullptr) h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    auto sym = [h](const char* name) { return dlsym(h, name); };
    lib.Connect = reinterpret_cast<decltype(lib.Connect)>(sym("hdfsConnect"));
    lib.Disconnect = reinterpret_cast<decltype(lib.Disconnect)>(sym("hdfsDisconnect"));
    lib.OpenFile = reinterpret_cast<decltype(lib.OpenFile)>(sym("hdfsOpenFile"));
    lib.CloseFile = reinterpret_cast<decltype(lib.CloseFile)>(sym("hdfsCloseFile"));
    lib.Read = reinterpret_cast<decltype(lib.Read)>(sym("hdfsRead"));
    lib.Write = reinterpret_cast<decltype(lib.Write)>(sym("hdfsWrite"));
    lib.Flush = reinterpret_cast<decltype(lib.Flush)>(sym("hdfsFlush"));
    lib.ok = lib.Connect && lib.Disconnect && lib.OpenFile && lib.CloseFile && lib.Read && lib.Write && lib.Flush;
  });
  return lib;
}

constexpr int32_t kMaxChunk = 1 << 30;   // tSize is 32 bit: move large buffers in pieces

}  // namespace

// ------------------------------------------------------------------------------- HDFSStream
HDFSStream::HDFSStream(void* fs, void* file, const std::string& path, bool writable)
    : fs_(fs), file_(file), path_(path), writable_(writable) {}

HDFSStream::~HDFSStream() {
  if (file_ == nullptr) return;
  if (writable_) Lib().Flush(fs_, file_);
  if (Lib().CloseFile(fs_, file_) != 0) Log::Error("hdfs: closing %s failed\n", path_.c_str());
}

void HDFSStream::Write(const void* buf, size_t size) {
  if (file_ == nullptr || !writable_) {
    Log::Error("hdfs: %s is not open for writing\n", path_.c_str());
    return;
  }
  const char* p = static_cast<const char*>(buf);
  while (size > 0) {
    const int32_t want = static_cast<int32_t>(std::min<size_t>(size, kMaxChunk));
    const int32_t done = Lib().Write(fs_, file_, p, want);
    if (done <= 0) {
      Log::Error("hdfs: write to %s failed\n", path_.c_str());
      return;
    }
    p += done;
    size -= static_cast<size_t>(done);
  }
}

size_t HDFSStream::Read(void* buf, size_t size) {
  if (file_ == nullptr || writable_) return 0;
  char* p = static_cast<char*>(buf);
  size_t total = 0;
  while (total < size) {   // hdfsRead may return short counts before the end of the file
    const int32_t want = static_cast<int32_t>(std::min<size_t>(size - total, kMaxChunk));
    const int32_t got = Lib().Read(fs_, file_, p + total, want);
    if (got < 0) {
      Log::Error("hdfs: read from %s failed\n", path_.c_str());
      break;
    }
    if (got == 0) break;
    total += static_cast<size_t>(got);
  }
  return total;
}

void HDFSStream::Flush() {
  if (file_ != nullptr && writable_) Lib().Flush(fs_, file_);
}

// ------------------------------------------------------------------------ HDFSStreamFactory
bool HDFSStreamFactory::Available() { return Lib().ok; }

HDFSStreamFactory::HDFSStreamFactory(const std::string& host) : host_(host) {
  if (!Lib().ok) return;
  // "namenode:port"; an empty host or "default" selects the configured default file system
  std::string name = host.empty() ? "default" : host;
  uint16_t port = 0;
  const size_t colon = name.rfind(':');
  if (colon != std::string::npos) {
    port = static_cast<uint16_t>(atoi(name.c_str() + colon + 1));
    name.resize(colon);
  }
  fs_ = Lib().Connect(name.c_str(), port);
  if (fs_ == nullptr) Log::Error("hdfs: cannot connect to %s\n", host.c_str());
}

HDFSStreamFactory::~HDFSStreamFactory() {
  if (fs_ != nullptr) Lib().Disconnect(fs_);
}

Stream* HDFSStreamFactory::Open(const URI& uri, FileOpenMode mode) {
  if (!Lib().ok) {
    Log::Error("hdfs://%s%s: libhdfs.so not found (set MV_LIBHDFS or add it to the loader path)\n",
               uri.host.c_str(), uri.name.c_str());
    return nullptr;
  }
  if (fs_ == nullptr) return nullptr;
  int flags = O_RDONLY;
  bool writable = false;
  switch (mode) {
    case FileOpenMode::Write:
    case FileOpenMode::BinaryWrite:
      flags = O_WRONLY;
      writable = true;
      break;
    case FileOpenMode::Append:
    case FileOpenMode::BinaryAppend:
      flags = O_WRONLY | O_APPEND;
      writable = true;
      break;
    default:
      break;
  }
  void* file = Lib().OpenFile(fs_, uri.name.c_str(), flags, 0, 0, 0);
  if (file == nullptr) {
    Log::Error("hdfs: cannot open %s (%s)\n", uri.path.c_str(), writable ? "write" : "read");
    return nullptr;
  }
  return new HDFSStream(fs_, file, uri.path, writable);
}

}  // namespace multiverso
