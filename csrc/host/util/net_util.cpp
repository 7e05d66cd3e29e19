#include "multiverso/util/net_util.h"
#include <arpa/inet.h>
#include <ifaddrs.h>
#include <netinet/in.h>

namespace multiverso {
namespace net {
void GetLocalIPAddress(std::set<std::string>* result) {
  result->clear();
  struct ifaddrs* ifs = nullptr;
  if (getifaddrs(&ifs) != 0) return;
  for (struct ifaddrs* it = ifs; it != nullptr; it = it->ifa_next) {
    if (it->ifa_addr == nullptr || it->ifa_addr->sa_family != AF_INET) continue;
    char buf[INET_ADDRSTRLEN];
    auto* sin = reinterpret_cast<struct sockaddr_in*>(it->ifa_addr);
    if (inet_ntop(AF_INET, &sin->sin_addr, buf, sizeof buf)) result->insert(buf);
  }
  freeifaddrs(ifs);
}
}  // namespace net
}  // namespace multiverso
