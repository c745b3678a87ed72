// Rendezvous for one-rank-per-process communicators: a star over an abstract-namespace Unix socket
// (rank 0 is the hub). Carries small blobs (allgather / barrier) and file descriptors (SCM_RIGHTS) —
// the cross-process GPU-buffer registry role the reference delegates to the RxDM sidecar over
// /run/tcpx (reference: gpudirect-tcpx/nccl-test-latest.yaml:62-90, gpudirect-tcpxo/best-practice.md:89-123).
// No CUDA in this file: it is unit-tested on CPU with memfd descriptors.
#pragma once
#include <stddef.h>
#include <string>
#include <vector>

namespace b200coll {

class Bootstrap {
 public:
  Bootstrap() = default;
  ~Bootstrap();
  Bootstrap(const Bootstrap&) = delete;
  Bootstrap& operator=(const Bootstrap&) = delete;

  // name: unique job id (becomes "\0b200coll-<name>"). Blocks until all ranks joined or timeout. Returns "" on success.
  std::string init(const std::string& name, int rank, int nranks, int timeout_ms);
  // out is resized to nranks*len; every rank contributes len bytes.
  std::string allgather(const void* mine, size_t len, std::vector<char>* out);
  std::string barrier();
  // Every rank contributes one fd; every rank receives nranks fds (index = source rank; own slot is a dup).
  std::string exchange_fds(int my_fd, std::vector<int>* fds);
  // Rank `root` passes fd (others pass -1); returns the received fd in *out (root gets a dup).
  std::string broadcast_fd(int root, int fd, int* out);
  void close_all();
  int rank() const { return rank_; }
  int nranks() const { return nranks_; }
  int rejected() const { return rejected_; }   // rank 0: connections dropped during init (wrong user, wrong token, malformed hello)

 private:
  int rank_ = -1, nranks_ = 0, timeout_ms_ = 0;
  int rejected_ = 0;
  int listen_fd_ = -1;
  int hub_fd_ = -1;                 // non-root: connection to rank 0
  std::vector<int> peer_fd_;        // root: connection per rank (index = rank; [0] unused)
};

// Low-level helpers (exposed for tests).
std::string uds_send_all(int fd, const void* buf, size_t len);
std::string uds_recv_all(int fd, void* buf, size_t len, int timeout_ms);
std::string uds_send_fd(int sock, int fd);
std::string uds_recv_fd(int sock, int* fd, int timeout_ms);

}  // namespace b200coll
