// cmvs-pmvs_b200/host/distributed.cpp -- one pmvs2 process per GPU.
//
// Launch: the usual one-process-per-GPU environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT), e.g.
//   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
//          --no-python cmvs-pmvs_b200/bin/pmvs2 prefix option.txt PATCH PSET
// Every rank reads the same files and keeps the same cell bookkeeping; the candidates of each seed / expansion wave are
// cut into contiguous shards, one per GPU, and the results of the ACCEPTED candidates are exchanged once per wave: by default
// every rank stores its message straight into the other ranks' GPU memory (CUDA IPC mailboxes over NVLink, pmvsb_peer_*),
// with PMVSB_EXCHANGE=nccl through one NCCL all-gather.  Waves below PMVSB_SHARD_MIN candidates are evaluated whole on every
// rank: no exchange at all.  Rank 0 writes the models.  A persistent TCP star through rank 0 carries the bring-up (the 64-byte
// IPC handles or the 128-byte NCCL id) and -- with PMVSB_EXCHANGE=tcp -- the wave exchange itself.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <thread>

#include "pmvs_host.hpp"

namespace pmvs {

namespace {
int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}
bool send_all(int fd, const void* buf, size_t n) {
  const char* p = (const char*)buf;
  while (n) { const ssize_t k = ::send(fd, p, n, 0); if (k <= 0) return false; p += k; n -= (size_t)k; }
  return true;
}
bool recv_all(int fd, void* buf, size_t n) {
  char* p = (char*)buf;
  while (n) { const ssize_t k = ::recv(fd, p, n, 0); if (k <= 0) return false; p += k; n -= (size_t)k; }
  return true;
}
}  // namespace

Dist Dist::from_env() {
  Dist d;
  d.world = std::max(1, env_int("WORLD_SIZE", 1));
  d.rank = env_int("RANK", 0);
  d.local_rank = env_int("LOCAL_RANK", d.rank);
  if (d.rank < 0 || d.rank >= d.world) { std::cerr << "pmvs2: RANK outside [0, WORLD_SIZE)" << std::endl; std::exit(1); }
  const char* a = std::getenv("MASTER_ADDR");
  d.master_addr = a && *a ? a : "127.0.0.1";
  d.port = env_int("MASTER_PORT", 29500) + 1017;   // MASTER_PORT itself belongs to the launcher's store
  if (d.port <= 0 || d.port > 65535) { std::cerr << "pmvs2: MASTER_PORT + 1017 = " << d.port << " is not a TCP port" << std::endl; std::exit(1); }
  // PMVSB_EXCHANGE = peer (default): the ranks store their wave messages into each other's GPU memory (CUDA IPC mailboxes over
  // NVLink; also works for ranks that share one GPU); nccl: one ncclAllGather per wave (NCCL refuses two ranks on one device);
  // tcp: over the rendezvous sockets, host-staged
  const char* e = std::getenv("PMVSB_EXCHANGE");
  const std::string mode = e ? e : "";
  if (mode.empty() || mode == "peer" || mode == "p2p") d.exchange = kPeer;
  else if (mode == "nccl") d.exchange = kNccl;
  else if (mode == "tcp") d.exchange = kTcp;
  else { std::cerr << "pmvs2: PMVSB_EXCHANGE must be peer, nccl or tcp" << std::endl; std::exit(1); }
  d.tcp_exchange = d.exchange == kTcp;
  d.shard_min = std::max(0, env_int("PMVSB_SHARD_MIN", d.shard_min));
  return d;
}

// A persistent star through rank 0: every peer connects once and announces its rank (a stray connection that does not is
// dropped, it cannot take a peer's place).  A rank that exits closes its sockets, so the others fail fast instead of
// waiting in a collective.
void Dist::connect() {
  if (world == 1 || !fds.empty()) return;
  const uint32_t magic = 0x504d5653u;   // "PMVS"
  if (rank == 0) {
    const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    ::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa{};
    sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((uint16_t)port);
    if (ls < 0 || ::bind(ls, (sockaddr*)&sa, sizeof(sa)) != 0 || ::listen(ls, world + 8) != 0) {
      std::cerr << "pmvs2: cannot listen on port " << port << " for the multi-GPU rendezvous" << std::endl;
      std::exit(1);
    }
    fds.assign(world, -1);
    int have = 1;
    while (have < world) {
      const int fd = ::accept(ls, nullptr, nullptr);
      if (fd < 0) { std::cerr << "pmvs2: rendezvous accept failed" << std::endl; std::exit(1); }
      timeval tv{5, 0};
      ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
      uint32_t hello[2] = {0, 0};
      if (!recv_all(fd, hello, sizeof(hello)) || hello[0] != magic || hello[1] == 0 || hello[1] >= (uint32_t)world || fds[hello[1]] != -1) { ::close(fd); continue; }
      timeval none{0, 0};
      ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &none, sizeof(none));
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      fds[hello[1]] = fd;
      ++have;
    }
    ::close(ls);
    return;
  }
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
  const std::string ps = std::to_string(port);
  if (::getaddrinfo(master_addr.c_str(), ps.c_str(), &hints, &res) != 0 || !res) {
    std::cerr << "pmvs2: cannot resolve MASTER_ADDR " << master_addr << std::endl;
    std::exit(1);
  }
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(120);
  for (;;) {
    const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
      const uint32_t hello[2] = {magic, (uint32_t)rank};
      int one = 1;
      ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      if (send_all(fd, hello, sizeof(hello))) { fds.assign(1, fd); break; }
      ::close(fd);
    } else if (fd >= 0) ::close(fd);
    if (std::chrono::steady_clock::now() > deadline) { std::cerr << "pmvs2: rank 0 did not answer the rendezvous" << std::endl; std::exit(1); }
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
  ::freeaddrinfo(res);
}

static void peer_lost() {
  std::cerr << "pmvs2: a peer rank left the run (see its own message); stopping" << std::endl;
  std::_Exit(1);
}

// rank 0 -> everybody: `n` bytes
void Dist::broadcast_from_root(void* buf, size_t n) const {
  if (world == 1) return;
  if (rank == 0) {
    for (int i = 1; i < world; ++i)
      if (!send_all(fds[i], buf, n)) peer_lost();
  } else if (!recv_all(fds[0], buf, n)) peer_lost();
}

// n bytes from every rank, in rank order, to every rank (through rank 0)
void Dist::allgather(const void* send, size_t n, void* recv) const {
  if (world == 1) { std::memcpy(recv, send, n); return; }
  char* out = (char*)recv;
  if (rank == 0) {
    std::memcpy(out, send, n);
    for (int i = 1; i < world; ++i)
      if (!recv_all(fds[i], out + (size_t)i * n, n)) peer_lost();
    for (int i = 1; i < world; ++i)
      if (!send_all(fds[i], out, n * (size_t)world)) peer_lost();
  } else {
    if (!send_all(fds[0], send, n) || !recv_all(fds[0], out, n * (size_t)world)) peer_lost();
  }
}

}  // namespace pmvs
