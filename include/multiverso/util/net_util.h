// net::GetLocalIPAddress (src/util/net_util.cpp:73-96) via getifaddrs.
#ifndef MULTIVERSO_UTIL_NET_UTIL_H_
#define MULTIVERSO_UTIL_NET_UTIL_H_
#include <set>
#include <string>
namespace multiverso {
namespace net {
void GetLocalIPAddress(std::set<std::string>* result);
}
}  // namespace multiverso
#endif
