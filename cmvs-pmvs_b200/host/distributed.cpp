// cmvs-pmvs_b200/host/distributed.cpp -- one pmvs2 process per GPU.
//
// Launch: the usual one-process-per-GPU environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT), e.g.
//   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
//          --no-python cmvs-pmvs_b200/bin/pmvs2 prefix option.txt PATCH PSET
// Every rank reads the same files and keeps the same cell bookkeeping; the candidates of each seed / expansion wave are
// cut into contiguous shards, one per GPU, and the per-candidate results are exchanged with one NCCL all-gather
// (pmvsb_allgather).  Rank 0 writes the models.  The 128-byte NCCL id travels over a plain TCP connection to rank 0.
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
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
  return d;
}

// rank 0 -> everybody: `n` bytes, one short-lived TCP connection per peer
void Dist::broadcast_from_root(void* buf, size_t n) const {
  if (world == 1) return;
  if (rank == 0) {
    const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    ::setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa{};
    sa.sin_family = AF_INET; sa.sin_addr.s_addr = htonl(INADDR_ANY); sa.sin_port = htons((uint16_t)port);
    if (ls < 0 || ::bind(ls, (sockaddr*)&sa, sizeof(sa)) != 0 || ::listen(ls, world) != 0) {
      std::cerr << "pmvs2: cannot listen on port " << port << " for the multi-GPU rendezvous" << std::endl;
      std::exit(1);
    }
    for (int i = 1; i < world; ++i) {
      const int fd = ::accept(ls, nullptr, nullptr);
      if (fd < 0 || !send_all(fd, buf, n)) { std::cerr << "pmvs2: rendezvous send failed" << std::endl; std::exit(1); }
      ::close(fd);
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
      const bool ok = recv_all(fd, buf, n);
      ::close(fd);
      if (ok) break;
    } else if (fd >= 0) ::close(fd);
    if (std::chrono::steady_clock::now() > deadline) { std::cerr << "pmvs2: rank 0 did not answer the rendezvous" << std::endl; std::exit(1); }
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
  }
  ::freeaddrinfo(res);
}

}  // namespace pmvs
