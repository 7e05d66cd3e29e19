// NetInterface singleton + typed all-reduce (counterpart of src/net.cpp:13-35).
#include "multiverso/net.h"
#include "multiverso/net/allreduce_engine.h"
#include "multiverso/net/tcp_net.h"
#include "multiverso/util/log.h"

namespace multiverso {

NetInterface* NetInterface::Get() {
  static TcpNet* net = new TcpNet();   // leaked: must outlive every actor thread
  return net;
}

namespace net {

template <typename T>
void Allreduce(T* data, size_t count) {
  NetInterface* n = NetInterface::Get();
  CHECK(n->active());
  AllreduceEngine engine(n);
  ReduceFunction sum = [](const char* src, char* dst, int len) {
    const T* s = reinterpret_cast<const T*>(src);
    T* d = reinterpret_cast<T*>(dst);
    const int cnt = len / static_cast<int>(sizeof(T));
    for (int i = 0; i < cnt; ++i) d[i] = static_cast<T>(d[i] + s[i]);
  };
  // the engine takes int byte counts; chunk anything larger than 1 GiB
  const size_t max_elems = (size_t(1) << 30) / sizeof(T);
  for (size_t off = 0; off < count; off += max_elems) {
    size_t n_el = count - off < max_elems ? count - off : max_elems;
    engine.Allreduce(reinterpret_cast<char*>(data + off), static_cast<int>(n_el), sizeof(T), sum);
  }
}
template void Allreduce<char>(char*, size_t);
template void Allreduce<int>(int*, size_t);
template void Allreduce<float>(float*, size_t);
template void Allreduce<double>(double*, size_t);

}  // namespace net
}  // namespace multiverso
