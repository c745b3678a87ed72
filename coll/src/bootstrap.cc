// Star-topology rendezvous over an abstract Unix socket: blob allgather, barrier, SCM_RIGHTS fd exchange (see bootstrap.h).
#include "bootstrap.h"

#include <errno.h>
#include <poll.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>

namespace b200coll {

static std::string errstr(const char* what) { return std::string(what) + ": " + strerror(errno); }

static long long now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (long long)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

static bool wait_readable(int fd, int timeout_ms) {      // timeout_ms <= 0: wait for ever
  pollfd p{fd, POLLIN, 0};
  while (true) {
    int r = poll(&p, 1, timeout_ms > 0 ? timeout_ms : -1);
    if (r > 0) return true;
    if (r == 0) return false;
    if (errno != EINTR) return false;
  }
}

std::string uds_send_all(int fd, const void* buf, size_t len) {
  const char* p = static_cast<const char*>(buf);
  while (len) {
    ssize_t n = send(fd, p, len, MSG_NOSIGNAL);
    if (n < 0) { if (errno == EINTR) continue; return errstr("send"); }
    p += n; len -= (size_t)n;
  }
  return "";
}

std::string uds_recv_all(int fd, void* buf, size_t len, int timeout_ms) {
  char* p = static_cast<char*>(buf);
  while (len) {
    if (!wait_readable(fd, timeout_ms)) return "recv: timed out waiting for peer";
    ssize_t n = recv(fd, p, len, 0);
    if (n == 0) return "recv: peer closed the bootstrap connection";
    if (n < 0) { if (errno == EINTR) continue; return errstr("recv"); }
    p += n; len -= (size_t)n;
  }
  return "";
}

std::string uds_send_fd(int sock, int fd) {
  char data = 'F';
  iovec iov{&data, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg{};
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  while (true) {
    ssize_t n = sendmsg(sock, &msg, MSG_NOSIGNAL);
    if (n == 1) return "";
    if (n < 0 && errno == EINTR) continue;
    return errstr("sendmsg(SCM_RIGHTS)");
  }
}

std::string uds_recv_fd(int sock, int* fd, int timeout_ms) {
  *fd = -1;
  if (!wait_readable(sock, timeout_ms)) return "recvmsg: timed out waiting for a file descriptor";
  char data = 0;
  iovec iov{&data, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg{};
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  ssize_t n;
  do { n = recvmsg(sock, &msg, MSG_CMSG_CLOEXEC); } while (n < 0 && errno == EINTR);
  if (n <= 0) return n == 0 ? std::string("recvmsg: peer closed") : errstr("recvmsg");
  for (cmsghdr* cm = CMSG_FIRSTHDR(&msg); cm; cm = CMSG_NXTHDR(&msg, cm)) {
    if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) {
      memcpy(fd, CMSG_DATA(cm), sizeof(int));
      return "";
    }
  }
  return "recvmsg: message carried no SCM_RIGHTS descriptor";
}

static socklen_t make_addr(const std::string& name, sockaddr_un* a) {
  memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  std::string full = "b200coll-" + name;
  size_t n = std::min(full.size(), sizeof(a->sun_path) - 2);
  a->sun_path[0] = '\0';   // abstract namespace: vanishes with the last fd, nothing to unlink
  memcpy(a->sun_path + 1, full.data(), n);
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
}

// Who is on the other end of a connected Unix socket. The rendezvous name is derived from values other processes on the host can
// guess (MASTER_ADDR:PORT), and what travels over this channel — SCM_RIGHTS descriptors of every rank's GPU arena — is the job's
// memory, so a connection only counts when it comes from this user and proves it knows the job's token (below).
static bool same_user(int fd) {
  ucred cr{};
  socklen_t len = sizeof(cr);
  if (getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cr, &len) != 0) return false;
  return cr.uid == geteuid();
}

// Second, independent hash of the rendezvous name plus an optional secret every rank of the job shares (B200COLL_RENDEZVOUS_SECRET,
// e.g. from the Job's env): sent in the hello, checked by rank 0. With the secret set, knowing MASTER_ADDR:PORT is not enough to join.
static unsigned long long job_token(const std::string& name) {
  const char* secret = getenv("B200COLL_RENDEZVOUS_SECRET");
  const std::string text = name + "|" + (secret ? secret : "");
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  for (unsigned char ch : text) { h ^= ch; h *= 0x100000001B3ull; h ^= h >> 29; }
  return h;
}

Bootstrap::~Bootstrap() { close_all(); }

void Bootstrap::close_all() {
  if (listen_fd_ >= 0) { close(listen_fd_); listen_fd_ = -1; }
  if (hub_fd_ >= 0) { close(hub_fd_); hub_fd_ = -1; }
  for (int& f : peer_fd_) if (f >= 0) { close(f); f = -1; }
}

std::string Bootstrap::init(const std::string& name, int rank, int nranks, int timeout_ms) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return "bootstrap: bad rank/nranks";
  rank_ = rank; nranks_ = nranks; timeout_ms_ = timeout_ms;
  if (nranks == 1) return "";
  sockaddr_un addr;
  socklen_t alen = make_addr(name, &addr);
  const bool forever = timeout_ms <= 0;
  const long long deadline = now_ms() + timeout_ms;
  const unsigned long long token = job_token(name);
  struct Hello { int rank, nranks; unsigned long long token; };
  if (rank == 0) {
    listen_fd_ = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (listen_fd_ < 0) return errstr("socket");
    if (bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), alen) < 0) return errstr("bind (is another job using the same id?)");
    if (listen(listen_fd_, nranks) < 0) return errstr("listen");
    peer_fd_.assign(nranks, -1);
    for (int joined = 1; joined < nranks;) {
      long long left = deadline - now_ms();
      if (!forever && left <= 0) return "bootstrap: timed out waiting for ranks to join";
      if (!wait_readable(listen_fd_, forever ? 0 : (int)left)) return "bootstrap: timed out waiting for ranks to join";
      int fd = accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
      if (fd < 0) { if (errno == EINTR) continue; return errstr("accept"); }
      // A stranger must not be able to join, and must not be able to abort the job by trying: drop the connection and keep waiting.
      if (!same_user(fd)) { close(fd); rejected_++; continue; }
      Hello hello{};
      std::string e = uds_recv_all(fd, &hello, sizeof(hello), 2000);
      if (!e.empty() || hello.token != token || hello.rank <= 0 || hello.rank >= nranks || hello.nranks != nranks || peer_fd_[hello.rank] >= 0) { close(fd); rejected_++; continue; }
      peer_fd_[hello.rank] = fd;
      joined++;
    }
    close(listen_fd_); listen_fd_ = -1;
  } else {
    while (true) {
      hub_fd_ = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
      if (hub_fd_ < 0) return errstr("socket");
      if (connect(hub_fd_, reinterpret_cast<sockaddr*>(&addr), alen) == 0) break;
      close(hub_fd_); hub_fd_ = -1;
      if (!forever && now_ms() > deadline) return "bootstrap: timed out connecting to rank 0";
      usleep(2000);
    }
    if (!same_user(hub_fd_)) return "bootstrap: the rendezvous socket is held by another user (name squatting?); refusing to hand over arena descriptors";
    Hello hello{rank, nranks, token};
    std::string e = uds_send_all(hub_fd_, &hello, sizeof(hello));
    if (!e.empty()) return e;
  }
  return barrier();
}

std::string Bootstrap::allgather(const void* mine, size_t len, std::vector<char>* out) {
  out->assign((size_t)nranks_ * len, 0);
  if (nranks_ == 1) { memcpy(out->data(), mine, len); return ""; }
  std::string e;
  if (rank_ == 0) {
    memcpy(out->data(), mine, len);
    for (int r = 1; r < nranks_; r++) if (!(e = uds_recv_all(peer_fd_[r], out->data() + (size_t)r * len, len, timeout_ms_)).empty()) return e;
    for (int r = 1; r < nranks_; r++) if (!(e = uds_send_all(peer_fd_[r], out->data(), out->size())).empty()) return e;
  } else {
    if (!(e = uds_send_all(hub_fd_, mine, len)).empty()) return e;
    if (!(e = uds_recv_all(hub_fd_, out->data(), out->size(), timeout_ms_)).empty()) return e;
  }
  return "";
}

std::string Bootstrap::barrier() {
  char b = 1;
  std::vector<char> all;
  return allgather(&b, 1, &all);
}

std::string Bootstrap::exchange_fds(int my_fd, std::vector<int>* fds) {
  fds->assign(nranks_, -1);
  (*fds)[rank_] = dup(my_fd);
  if (nranks_ == 1) return "";
  std::string e;
  if (rank_ == 0) {
    for (int r = 1; r < nranks_; r++) if (!(e = uds_recv_fd(peer_fd_[r], &(*fds)[r], timeout_ms_)).empty()) return e;
    for (int r = 1; r < nranks_; r++)
      for (int s = 0; s < nranks_; s++) {
        if (s == r) continue;
        if (!(e = uds_send_fd(peer_fd_[r], (*fds)[s])).empty()) return e;
      }
  } else {
    if (!(e = uds_send_fd(hub_fd_, my_fd)).empty()) return e;
    for (int s = 0; s < nranks_; s++) {
      if (s == rank_) continue;
      if (!(e = uds_recv_fd(hub_fd_, &(*fds)[s], timeout_ms_)).empty()) return e;
    }
  }
  return barrier();
}

std::string Bootstrap::broadcast_fd(int root, int fd, int* out) {
  *out = -1;
  if (nranks_ == 1) { *out = dup(fd); return ""; }
  std::string e;
  // star topology: route through rank 0
  if (rank_ == 0) {
    int src = fd;
    bool owned = false;
    if (root != 0) { if (!(e = uds_recv_fd(peer_fd_[root], &src, timeout_ms_)).empty()) return e; owned = true; }
    for (int r = 1; r < nranks_; r++) {
      if (r == root) continue;
      if (!(e = uds_send_fd(peer_fd_[r], src)).empty()) return e;
    }
    *out = owned ? src : dup(src);
  } else if (rank_ == root) {
    if (!(e = uds_send_fd(hub_fd_, fd)).empty()) return e;
    *out = dup(fd);
  } else {
    if (!(e = uds_recv_fd(hub_fd_, out, timeout_ms_)).empty()) return e;
  }
  return barrier();
}

}  // namespace b200coll
