// TcpNet: full-mesh TCP transport (see include/multiverso/net/tcp_net.h).
#include "multiverso/net/tcp_net.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>

#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/util/net_util.h"

namespace multiverso {

MV_DEFINE_string(machine_file, "", "machine file path: one ip[:port] per line");
MV_DEFINE_int(port, 55555, "base port of the TCP control plane");

namespace {

struct FrameHeader {
  uint32_t kind;   // 0 = Message, 1 = raw, 2 = bye (the peer is shutting down in order)
  uint32_t src;
  uint64_t len;
};

const char* EnvOr(const char* a, const char* b) {
  const char* v = getenv(a);
  if (v && *v) return v;
  v = b ? getenv(b) : nullptr;
  return (v && *v) ? v : nullptr;
}

bool WriteAll(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n > 0) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += w;
    n -= static_cast<size_t>(w);
  }
  return true;
}

int ConnectWithRetry(const std::string& host, int port, double timeout_s) {
  auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s);
  for (;;) {
    struct addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    std::string ps = std::to_string(port);
    if (getaddrinfo(host.c_str(), ps.c_str(), &hints, &res) == 0 && res) {
      int fd = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
      if (fd >= 0) {
        if (::connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
          freeaddrinfo(res);
          int one = 1;
          setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
          return fd;
        }
        ::close(fd);
      }
      freeaddrinfo(res);
    }
    if (std::chrono::steady_clock::now() > deadline) return -1;
    usleep(20000);
  }
}

}  // namespace

TcpNet::TcpNet() = default;
TcpNet::~TcpNet() {
  if (active_) Finalize();
}

bool TcpNet::ParseEndpoint(const std::string& ep, std::string* host, int* port) {
  std::string s = ep;
  size_t p = s.find("://");
  if (p != std::string::npos) s = s.substr(p + 3);
  size_t c = s.rfind(':');
  if (c == std::string::npos) {
    *host = s;
    return false;
  }
  *host = s.substr(0, c);
  *port = atoi(s.c_str() + c + 1);
  return true;
}

bool TcpNet::ReadExact(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n > 0) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r == 0) return false;
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    p += r;
    n -= static_cast<size_t>(r);
  }
  return true;
}

int TcpNet::Bind(int rank, char* endpoint) {
  rank_ = rank;
  bound_endpoint_ = endpoint;
  std::string host;
  int port = 0;
  if (!ParseEndpoint(bound_endpoint_, &host, &port)) {
    Log::Error("NetBind: endpoint '%s' must be ip:port", endpoint);
    return -1;
  }
  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  struct sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = INADDR_ANY;
  addr.sin_port = htons(static_cast<uint16_t>(port));
  if (::bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0 ||
      ::listen(listen_fd_, 64) != 0) {
    Log::Error("NetBind: cannot listen on %s (%s)", endpoint, strerror(errno));
    ::close(listen_fd_);
    listen_fd_ = -1;
    return -1;
  }
  return 0;
}

int TcpNet::Connect(int* ranks, char* endpoints[], int n) {
  // ranks/endpoints list every participant (own rank may be included).
  int max_rank = rank_;
  for (int i = 0; i < n; ++i) max_rank = ranks[i] > max_rank ? ranks[i] : max_rank;
  std::vector<std::string> eps(static_cast<size_t>(max_rank) + 1);
  for (int i = 0; i < n; ++i) eps[ranks[i]] = endpoints[i];
  eps[rank_] = bound_endpoint_;
  size_ = max_rank + 1;
  EstablishMesh(eps);
  return 0;
}

void TcpNet::Init(int* argc, char** argv) {
  if (active_) return;
  ParseCMDFlags(argc, argv);
  std::vector<std::string> eps;
  const std::string mf = MV_CONFIG(machine_file);
  if (!mf.empty()) {
    // ZMQ-style: one ip[:port] per line, own rank = first line matching a local NIC address
    std::ifstream in(mf);
    if (!in) Log::Fatal("cannot open machine file %s", mf.c_str());
    std::string line;
    std::set<std::string> local;
    net::GetLocalIPAddress(&local);
    int idx = 0, mine = -1;
    std::map<std::string, int> seen;
    while (std::getline(in, line)) {
      while (!line.empty() && isspace(static_cast<unsigned char>(line.back()))) line.pop_back();
      if (line.empty()) continue;
      std::string host;
      int port = MV_CONFIG(port);
      if (!ParseEndpoint(line, &host, &port)) port = MV_CONFIG(port) + seen[host];
      seen[host]++;
      if (mine < 0 && local.count(host)) {
        const char* forced = EnvOr("MV_RANK", nullptr);
        if (!forced || atoi(forced) == idx) mine = idx;
      }
      eps.push_back(host + ":" + std::to_string(port));
      ++idx;
    }
    if (mine < 0) Log::Fatal("machine file %s does not list a local address", mf.c_str());
    rank_ = mine;
    size_ = static_cast<int>(eps.size());
  } else {
    const char* r = EnvOr("MV_RANK", "RANK");
    const char* s = EnvOr("MV_SIZE", "WORLD_SIZE");
    rank_ = r ? atoi(r) : 0;
    size_ = s ? atoi(s) : 1;
    if (size_ > 1) {
      const char* host = EnvOr("MV_MASTER_ADDR", "MASTER_ADDR");
      int base = MV_CONFIG(port);
      if (const char* p = getenv("MV_PORT")) base = atoi(p);
      else if (const char* mp = getenv("MASTER_PORT")) base = atoi(mp) + 64;
      for (int i = 0; i < size_; ++i)
        eps.push_back(std::string(host ? host : "127.0.0.1") + ":" + std::to_string(base + i));
    }
  }
  if (size_ > 1) {
    std::string ep = eps[rank_];
    if (Bind(rank_, const_cast<char*>(ep.c_str())) != 0) Log::Fatal("TcpNet: bind %s failed", ep.c_str());
    EstablishMesh(eps);
  } else {
    fds_.assign(1, -1);
    send_mu_.clear();
    send_mu_.emplace_back(new std::mutex());
    raw_in_.clear();
    raw_in_.emplace_back(new MtQueue<std::shared_ptr<RawChunk>>());
    raw_partial_.assign(1, nullptr);
    active_ = true;
  }
  Log::Get().SetRank(size_ > 1 ? rank_ : -1);
  Log::Debug("TcpNet initialised: rank %d of %d", rank_, size_);
}

void TcpNet::EstablishMesh(const std::vector<std::string>& eps) {
  fds_.assign(size_, -1);
  send_mu_.clear();
  raw_in_.clear();
  for (int i = 0; i < size_; ++i) {
    send_mu_.emplace_back(new std::mutex());
    raw_in_.emplace_back(new MtQueue<std::shared_ptr<RawChunk>>());
  }
  raw_partial_.assign(size_, nullptr);
  // connect to every lower rank, accept from every higher rank
  for (int peer = 0; peer < rank_; ++peer) {
    std::string host;
    int port = 0;
    ParseEndpoint(eps[peer], &host, &port);
    int fd = ConnectWithRetry(host, port, 120.0);
    if (fd < 0) Log::Fatal("TcpNet: rank %d cannot connect to rank %d at %s", rank_, peer, eps[peer].c_str());
    uint32_t me = static_cast<uint32_t>(rank_);
    WriteAll(fd, &me, sizeof me);
    fds_[peer] = fd;
  }
  for (int k = rank_ + 1; k < size_; ++k) {
    int fd = ::accept(listen_fd_, nullptr, nullptr);
    if (fd < 0) Log::Fatal("TcpNet: accept failed (%s)", strerror(errno));
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    uint32_t who = 0;
    if (!ReadExact(fd, &who, sizeof who) || who >= static_cast<uint32_t>(size_))
      Log::Fatal("TcpNet: bad handshake");
    fds_[who] = fd;
  }
  stopping_ = false;
  active_ = true;
  receiver_ = std::thread([this] { ReceiverLoop(); });
}

void TcpNet::Finalize() {
  if (!active_) return;
  stopping_ = true;     // from here on a failed send is not an error (the peer may have closed already)
  // tell every peer that this end closes on purpose: an EOF without a preceding bye is a dead rank
  for (int i = 0; i < size_; ++i)
    if (i != rank_ && fds_[i] >= 0) WriteFrame(i, 2, {});
  for (int fd : fds_)
    if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
  if (receiver_.joinable()) receiver_.join();
  for (int& fd : fds_) {
    if (fd >= 0) ::close(fd);
    fd = -1;
  }
  if (listen_fd_ >= 0) ::close(listen_fd_);
  listen_fd_ = -1;
  inbox_.Exit();
  for (auto& q : raw_in_) q->Exit();
  active_ = false;
}

void TcpNet::ReceiverLoop() {
  std::vector<struct pollfd> pfds;
  std::vector<int> owner;
  for (int i = 0; i < size_; ++i)
    if (fds_[i] >= 0) {
      pfds.push_back({fds_[i], POLLIN, 0});
      owner.push_back(i);
    }
  size_t open = pfds.size();
  std::vector<char> said_bye(pfds.size(), 0);
  while (!stopping_ && open > 0) {
    int rc = ::poll(pfds.data(), pfds.size(), 200);
    if (rc <= 0) continue;
    for (size_t k = 0; k < pfds.size(); ++k) {
      if (pfds[k].fd < 0 || !(pfds[k].revents & (POLLIN | POLLHUP | POLLERR))) continue;
      FrameHeader h;
      if (!ReadExact(pfds[k].fd, &h, sizeof h)) {
        // an orderly shutdown closes sockets only after stopping_ is set; a peer that disappears before that is
        // dead, and every request it still owes a reply to would wait forever: say so once, loudly
        if (!stopping_ && !said_bye[k])
          Log::Error("rank %d: lost the connection to rank %d (peer exited or crashed); requests to it cannot complete\n",
                     rank_, owner[k]);
        pfds[k].fd = -1;
        --open;
        continue;
      }
      if (h.kind == 2) {
        said_bye[k] = 1;
        continue;
      }
      if (h.kind == 0) {
        MessagePtr msg(new Message());
        bool ok = ReadExact(pfds[k].fd, msg->header(), Message::kHeaderSize * sizeof(int));
        uint64_t nblobs = 0;
        ok = ok && ReadExact(pfds[k].fd, &nblobs, sizeof nblobs);
        for (uint64_t b = 0; ok && b < nblobs; ++b) {
          uint64_t len = 0;
          ok = ReadExact(pfds[k].fd, &len, sizeof len);
          Blob blob(static_cast<size_t>(len));
          if (ok && len) ok = ReadExact(pfds[k].fd, blob.data(), static_cast<size_t>(len));
          msg->Push(std::move(blob));
        }
        if (!ok) {
          pfds[k].fd = -1;
          --open;
          continue;
        }
        inbox_.Push(std::move(msg));
      } else {
        auto chunk = std::make_shared<RawChunk>();
        chunk->bytes.resize(static_cast<size_t>(h.len));
        if (h.len && !ReadExact(pfds[k].fd, chunk->bytes.data(), static_cast<size_t>(h.len))) {
          pfds[k].fd = -1;
          --open;
          continue;
        }
        raw_in_[owner[k]]->Push(std::move(chunk));
      }
    }
  }
}

void TcpNet::WriteFrame(int dst, uint32_t kind,
                        const std::vector<std::pair<const void*, size_t>>& parts) {
  uint64_t total = 0;
  for (auto& p : parts) total += p.second;
  FrameHeader h{kind, static_cast<uint32_t>(rank_), total};
  std::lock_guard<std::mutex> lk(*send_mu_[dst]);
  bool ok = WriteAll(fds_[dst], &h, sizeof h);
  for (auto& p : parts)
    if (ok && p.second) ok = WriteAll(fds_[dst], p.first, p.second);
  if (!ok && !stopping_) Log::Fatal("TcpNet: send to rank %d failed (%s)", dst, strerror(errno));
}

size_t TcpNet::Send(MessagePtr& msg) {
  const int dst = msg->dst();
  size_t bytes = Message::kHeaderSize * sizeof(int);
  for (auto& b : msg->data()) bytes += b.size();
  if (dst == rank_) {
    inbox_.Push(std::move(msg));
    return bytes;
  }
  CHECK(dst >= 0 && dst < size_);
  std::vector<std::pair<const void*, size_t>> parts;
  std::vector<uint64_t> lens(msg->data().size());
  uint64_t nblobs = msg->data().size();
  parts.emplace_back(msg->header(), Message::kHeaderSize * sizeof(int));
  parts.emplace_back(&nblobs, sizeof nblobs);
  for (size_t i = 0; i < msg->data().size(); ++i) {
    lens[i] = msg->data()[i].size();
    parts.emplace_back(&lens[i], sizeof(uint64_t));
    parts.emplace_back(msg->data()[i].data(), msg->data()[i].size());
  }
  WriteFrame(dst, 0, parts);
  return bytes;
}

size_t TcpNet::Recv(MessagePtr* msg) {
  MessagePtr m;
  if (!inbox_.Pop(m)) return static_cast<size_t>(-1);
  size_t bytes = Message::kHeaderSize * sizeof(int);
  for (auto& b : m->data()) bytes += b.size();
  *msg = std::move(m);
  return bytes;
}

void TcpNet::SendTo(int rank, const char* buf, int len) {
  if (rank == rank_) {
    auto chunk = std::make_shared<RawChunk>();
    chunk->bytes.assign(buf, buf + len);
    raw_in_[rank_]->Push(std::move(chunk));
    return;
  }
  WriteFrame(rank, 1, {{buf, static_cast<size_t>(len)}});
}

void TcpNet::RecvFrom(int rank, char* buf, int len) {
  size_t need = static_cast<size_t>(len);
  while (need > 0) {
    std::shared_ptr<RawChunk>& cur = raw_partial_[rank];
    if (!cur || cur->consumed == cur->bytes.size()) {
      cur.reset();
      std::shared_ptr<RawChunk> next;
      if (!raw_in_[rank]->Pop(next)) Log::Fatal("TcpNet: RecvFrom(%d) on a finalized net", rank);
      cur = std::move(next);
      if (cur->bytes.empty()) continue;
    }
    size_t take = cur->bytes.size() - cur->consumed;
    if (take > need) take = need;
    std::memcpy(buf, cur->bytes.data() + cur->consumed, take);
    cur->consumed += take;
    buf += take;
    need -= take;
  }
}

void TcpNet::SendRecv(int send_rank, const char* send_buf, int send_len, int recv_rank,
                      char* recv_buf, int recv_len) {
  // sends never block on the peer's receive (the receiver thread drains sockets), so a
  // plain send-then-receive cannot deadlock
  if (send_len > 0) SendTo(send_rank, send_buf, send_len);
  if (recv_len > 0) RecvFrom(recv_rank, recv_buf, recv_len);
}

}  // namespace multiverso
