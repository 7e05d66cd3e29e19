#include "multiverso/io/io.h"
#include <cstring>
#include "multiverso/io/hdfs_stream.h"
#include "multiverso/io/local_stream.h"
#include "multiverso/util/log.h"

namespace multiverso {

URI::URI(const std::string& uri) : path(uri) {
  size_t p = uri.find("://");
  if (p == std::string::npos) {
    scheme = "file";
    name = uri;
    return;
  }
  scheme = uri.substr(0, p);
  std::string rest = uri.substr(p + 3);
  if (scheme == "file") {
    name = rest;
    return;
  }
  size_t slash = rest.find('/');
  if (slash == std::string::npos) {
    host = rest;
    name = "/";
  } else {
    host = rest.substr(0, slash);
    name = rest.substr(slash);
  }
}

std::map<std::string, std::unique_ptr<StreamFactory>>& StreamFactory::instances() {
  static auto* m = new std::map<std::string, std::unique_ptr<StreamFactory>>();
  return *m;
}

Stream* StreamFactory::GetStream(const URI& uri, FileOpenMode mode) {
  std::string key = uri.scheme + "://" + uri.host;
  auto& inst = instances();
  auto it = inst.find(key);
  if (it == inst.end()) {
    std::unique_ptr<StreamFactory> f;
    if (uri.scheme == "file") f.reset(new LocalStreamFactory());
    else if (uri.scheme == "hdfs") f.reset(new HDFSStreamFactory(uri.host));
    else {
      Log::Error("StreamFactory: unsupported scheme '%s'", uri.scheme.c_str());
      return nullptr;
    }
    it = inst.emplace(key, std::move(f)).first;
  }
  return it->second->Open(uri, mode);
}

TextReader::TextReader(const URI& uri, size_t buf_size)
    : stream_(StreamFactory::GetStream(uri, FileOpenMode::Read)), buf_(new char[buf_size]),
      buf_size_(buf_size) {
  if (stream_ && !stream_->Good()) {
    delete stream_;
    stream_ = nullptr;
  }
}
TextReader::~TextReader() {
  delete stream_;
  delete[] buf_;
}
size_t TextReader::Fill() {
  pos_ = 0;
  length_ = stream_ ? stream_->Read(buf_, buf_size_) : 0;
  return length_;
}
bool TextReader::GetLine(std::string& line) {
  line.clear();
  if (!stream_) return false;
  bool any = false;
  for (;;) {
    if (pos_ >= length_ && Fill() == 0) return any;
    char* start = buf_ + pos_;
    char* nl = static_cast<char*>(memchr(start, '\n', length_ - pos_));
    if (nl) {
      line.append(start, nl - start);
      pos_ = (nl - buf_) + 1;
      if (!line.empty() && line.back() == '\r') line.pop_back();
      return true;
    }
    line.append(start, length_ - pos_);
    pos_ = length_;
    any = true;
  }
}

}  // namespace multiverso
