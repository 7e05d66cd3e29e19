// Role bitmask and Node POD (include/multiverso/node.h:6-27).
#ifndef MULTIVERSO_NODE_H_
#define MULTIVERSO_NODE_H_
namespace multiverso {

enum Role { NONE = 0, WORKER = 1, SERVER = 2, ALL = 3 };

struct Node {
  int rank = -1;
  int role = Role::ALL;
  int worker_id = -1;
  int server_id = -1;
};

namespace node {
inline bool is_worker(int role) { return (role & Role::WORKER) != 0; }
inline bool is_server(int role) { return (role & Role::SERVER) != 0; }
}  // namespace node

}  // namespace multiverso
#endif
