// Single-node MPI subset over a local TCP mesh (see mpi.h). Tooling for the reference arm.
#include "mpi.h"

#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Msg {
  int src;
  std::vector<char> data;
};

struct State {
  bool inited = false;
  int rank = 0, size = 1;
  std::vector<int> fds;
  std::vector<std::unique_ptr<std::mutex>> send_mu;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<Msg>> p2p;    // arrival order (point-to-point)
  std::deque<std::shared_ptr<Msg>> coll;   // collective channel
  std::thread receiver;
  bool stopping = false;
  int listen_fd = -1;
};
State g;

struct Frame {
  uint32_t kind;   // 0 p2p, 1 collective
  uint32_t src;
  uint64_t len;
};

int type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_INT: case MPI_FLOAT: return 4;
    case MPI_DOUBLE: return 8;
    default: return 1;
  }
}

bool write_all(int fd, const void* b, size_t n) {
  const char* p = static_cast<const char*>(b);
  while (n) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w <= 0) { if (errno == EINTR) continue; return false; }
    p += w; n -= w;
  }
  return true;
}
bool read_all(int fd, void* b, size_t n) {
  char* p = static_cast<char*>(b);
  while (n) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r <= 0) { if (r < 0 && errno == EINTR) continue; return false; }
    p += r; n -= r;
  }
  return true;
}

void deliver(uint32_t kind, int src, std::vector<char>&& data) {
  auto m = std::make_shared<Msg>();
  m->src = src;
  m->data = std::move(data);
  {
    std::lock_guard<std::mutex> lk(g.mu);
    (kind == 0 ? g.p2p : g.coll).push_back(std::move(m));
  }
  g.cv.notify_all();
}

void send_frame(int dst, uint32_t kind, const void* buf, size_t len) {
  if (dst == g.rank) {
    std::vector<char> d(static_cast<const char*>(buf), static_cast<const char*>(buf) + len);
    deliver(kind, g.rank, std::move(d));
    return;
  }
  Frame f{kind, static_cast<uint32_t>(g.rank), len};
  std::lock_guard<std::mutex> lk(*g.send_mu[dst]);
  if (!write_all(g.fds[dst], &f, sizeof f) || (len && !write_all(g.fds[dst], buf, len))) {
    fprintf(stderr, "[mpi_shim] send to %d failed\n", dst);
    abort();
  }
}

void receiver_loop() {
  std::vector<pollfd> pf;
  for (int i = 0; i < g.size; ++i)
    if (g.fds[i] >= 0) pf.push_back({g.fds[i], POLLIN, 0});
  size_t open = pf.size();
  while (!g.stopping && open) {
    if (poll(pf.data(), pf.size(), 100) <= 0) continue;
    for (auto& p : pf) {
      if (p.fd < 0 || !(p.revents & (POLLIN | POLLHUP | POLLERR))) continue;
      Frame f;
      if (!read_all(p.fd, &f, sizeof f)) { p.fd = -1; --open; continue; }
      std::vector<char> d(f.len);
      if (f.len && !read_all(p.fd, d.data(), f.len)) { p.fd = -1; --open; continue; }
      deliver(f.kind, static_cast<int>(f.src), std::move(d));
    }
  }
}

std::shared_ptr<Msg> pop(std::deque<std::shared_ptr<Msg>>& q, int source) {
  std::unique_lock<std::mutex> lk(g.mu);
  for (;;) {
    for (auto it = q.begin(); it != q.end(); ++it)
      if (source == MPI_ANY_SOURCE || (*it)->src == source) {
        auto m = *it;
        q.erase(it);
        return m;
      }
    g.cv.wait(lk);
  }
}

int env_int(const char* a, const char* b, int dflt) {
  const char* v = getenv(a);
  if (!v || !*v) v = b ? getenv(b) : nullptr;
  return (v && *v) ? atoi(v) : dflt;
}

}  // namespace

extern "C" {

int MPI_Initialized(int* flag) { *flag = g.inited ? 1 : 0; return MPI_SUCCESS; }

int MPI_Init_thread(int*, char***, int, int* provided) {
  if (provided) *provided = MPI_THREAD_SERIALIZED;
  if (g.inited) return MPI_SUCCESS;
  g.rank = env_int("MV_SHIM_RANK", "RANK", 0);
  g.size = env_int("MV_SHIM_SIZE", "WORLD_SIZE", 1);
  const int base = env_int("MV_SHIM_PORT", nullptr, env_int("MASTER_PORT", nullptr, 29400) + 300);
  g.fds.assign(g.size, -1);
  for (int i = 0; i < g.size; ++i) g.send_mu.emplace_back(new std::mutex());
  if (g.size > 1) {
    g.listen_fd = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(g.listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    a.sin_port = htons(static_cast<uint16_t>(base + g.rank));
    if (bind(g.listen_fd, reinterpret_cast<sockaddr*>(&a), sizeof a) || listen(g.listen_fd, 64)) {
      perror("[mpi_shim] bind/listen");
      abort();
    }
    for (int peer = 0; peer < g.rank; ++peer) {
      int fd = -1;
      for (int tries = 0; tries < 6000; ++tries) {
        fd = socket(AF_INET, SOCK_STREAM, 0);
        sockaddr_in p{};
        p.sin_family = AF_INET;
        p.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
        p.sin_port = htons(static_cast<uint16_t>(base + peer));
        if (connect(fd, reinterpret_cast<sockaddr*>(&p), sizeof p) == 0) break;
        close(fd);
        fd = -1;
        usleep(20000);
      }
      if (fd < 0) { fprintf(stderr, "[mpi_shim] cannot reach rank %d\n", peer); abort(); }
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      uint32_t me = g.rank;
      write_all(fd, &me, 4);
      g.fds[peer] = fd;
    }
    for (int k = g.rank + 1; k < g.size; ++k) {
      int fd = accept(g.listen_fd, nullptr, nullptr);
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      uint32_t who = 0;
      read_all(fd, &who, 4);
      g.fds[who] = fd;
    }
    g.receiver = std::thread(receiver_loop);
  }
  g.inited = true;
  return MPI_SUCCESS;
}

int MPI_Query_thread(int* provided) { *provided = MPI_THREAD_SERIALIZED; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* r) { *r = g.rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* s) { *s = g.size; return MPI_SUCCESS; }

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
  const size_t bytes = static_cast<size_t>(count) * type_size(type);
  if (sendbuf != MPI_IN_PLACE) memcpy(recvbuf, sendbuf, bytes);
  if (g.size == 1) return MPI_SUCCESS;
  if (g.rank == 0) {
    for (int k = 1; k < g.size; ++k) {
      auto m = pop(g.coll, MPI_ANY_SOURCE);
      for (int i = 0; i < count; ++i) {
        switch (type) {
          case MPI_INT: static_cast<int*>(recvbuf)[i] += reinterpret_cast<int*>(m->data.data())[i]; break;
          case MPI_FLOAT: static_cast<float*>(recvbuf)[i] += reinterpret_cast<float*>(m->data.data())[i]; break;
          case MPI_DOUBLE: static_cast<double*>(recvbuf)[i] += reinterpret_cast<double*>(m->data.data())[i]; break;
          default: static_cast<char*>(recvbuf)[i] += m->data[i]; break;
        }
      }
    }
    for (int k = 1; k < g.size; ++k) send_frame(k, 1, recvbuf, bytes);
  } else {
    send_frame(0, 1, recvbuf, bytes);
    auto m = pop(g.coll, 0);
    memcpy(recvbuf, m->data.data(), bytes);
  }
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm c) {
  int one = 1;
  return MPI_Allreduce(MPI_IN_PLACE, &one, 1, MPI_INT, MPI_SUM, c);
}

int MPI_Finalize(void) {
  if (!g.inited) return MPI_SUCCESS;
  if (g.size > 1) {
    MPI_Barrier(MPI_COMM_WORLD);
    g.stopping = true;
    for (int fd : g.fds) if (fd >= 0) shutdown(fd, SHUT_RDWR);
    if (g.receiver.joinable()) g.receiver.join();
    for (int fd : g.fds) if (fd >= 0) close(fd);
    if (g.listen_fd >= 0) close(g.listen_fd);
  }
  g.inited = false;
  return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int, MPI_Comm, MPI_Request* req) {
  send_frame(dest, 0, buf, static_cast<size_t>(count) * type_size(type));
  if (req) *req = 0;
  return MPI_SUCCESS;
}

int MPI_Iprobe(int source, int, MPI_Comm, int* flag, MPI_Status* status) {
  std::lock_guard<std::mutex> lk(g.mu);
  *flag = 0;
  for (auto& m : g.p2p)
    if (source == MPI_ANY_SOURCE || m->src == source) {
      *flag = 1;
      if (status) {
        status->MPI_SOURCE = m->src;
        status->MPI_TAG = 0;
        status->MPI_ERROR = 0;
        status->count_bytes = static_cast<int>(m->data.size());
      }
      break;
    }
  return MPI_SUCCESS;
}

int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int, MPI_Comm, MPI_Status* status) {
  auto m = pop(g.p2p, source);
  size_t cap = static_cast<size_t>(count) * type_size(type);
  size_t n = m->data.size() < cap ? m->data.size() : cap;
  memcpy(buf, m->data.data(), n);
  if (status) {
    status->MPI_SOURCE = m->src;
    status->MPI_TAG = 0;
    status->MPI_ERROR = 0;
    status->count_bytes = static_cast<int>(n);
  }
  return MPI_SUCCESS;
}

int MPI_Get_count(const MPI_Status* status, MPI_Datatype type, int* count) {
  *count = status->count_bytes / type_size(type);
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }
int MPI_Waitall(int, MPI_Request[], MPI_Status[]) { return MPI_SUCCESS; }
int MPI_Testall(int, MPI_Request[], int* flag, MPI_Status[]) { *flag = 1; return MPI_SUCCESS; }

}  // extern "C"
