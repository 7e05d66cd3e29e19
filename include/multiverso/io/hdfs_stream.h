// HDFS stream placeholder: the scheme is parsed, opening fails with a clear message
// (there is no libhdfs / network in this environment; reference: io/hdfs_stream.h, Q16).
#ifndef MULTIVERSO_IO_HDFS_STREAM_H_
#define MULTIVERSO_IO_HDFS_STREAM_H_
#include "multiverso/io/io.h"
namespace multiverso {
class HDFSStreamFactory : public StreamFactory {
 public:
  Stream* Open(const URI& uri, FileOpenMode mode) override;
};
}  // namespace multiverso
#endif
