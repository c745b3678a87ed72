// Minimal HTTP/2 + HPACK + gRPC framing over Unix sockets — just enough for a kubelet device plugin:
// a server for unary and server-streaming methods, and a unary client (Registration.Register, PodResourcesLister.List).
// The image has a gRPC *runtime* for Python but no C++ gRPC / protobuf headers, so the wire layer is written out here
// (RFC 7540 framing, RFC 7541 header compression incl. Huffman decoding, the gRPC length-prefixed message format).
// Not a general HTTP/2 stack: no priorities, no push, no server-side CONTINUATION emission, cleartext prior-knowledge only.
#pragma once
#include <errno.h>
#include <poll.h>
#include <stdint.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace h2 {

#include "hpack_tables.inc"

using Headers = std::vector<std::pair<std::string, std::string>>;

enum FrameType : uint8_t { DATA = 0, HEADERS = 1, PRIORITY = 2, RST_STREAM = 3, SETTINGS = 4, PUSH_PROMISE = 5, PING = 6, GOAWAY = 7, WINDOW_UPDATE = 8, CONTINUATION = 9 };
enum Flags : uint8_t { END_STREAM = 0x1, ACK = 0x1, END_HEADERS = 0x4, PADDED = 0x8, PRIORITY_FLAG = 0x20 };
static const char kPreface[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";

// ------------------------------------------------------------------------------------------------ HPACK
class HuffmanDecoder {
 public:
  HuffmanDecoder() {
    nodes_.push_back(Node{{-1, -1}, -1});
    for (int sym = 0; sym < 256; sym++) {
      int cur = 0;
      for (int b = kHuffLen[sym] - 1; b >= 0; b--) {
        const int bit = (kHuffCode[sym] >> b) & 1;
        if (nodes_[cur].child[bit] < 0) { nodes_[cur].child[bit] = (int)nodes_.size(); nodes_.push_back(Node{{-1, -1}, -1}); }
        cur = nodes_[cur].child[bit];
      }
      nodes_[cur].sym = sym;
    }
  }
  bool decode(const uint8_t* p, size_t n, std::string* out) const {
    int cur = 0, pad_bits = 0;
    bool pad_all_ones = true;
    for (size_t i = 0; i < n; i++)
      for (int b = 7; b >= 0; b--) {
        const int bit = (p[i] >> b) & 1;
        cur = nodes_[cur].child[bit];
        if (cur < 0) return false;
        pad_bits++; pad_all_ones = pad_all_ones && bit;
        if (nodes_[cur].sym >= 0) { out->push_back((char)nodes_[cur].sym); cur = 0; pad_bits = 0; pad_all_ones = true; }
      }
    return pad_bits < 8 && pad_all_ones;   // trailing bits must be a (short) prefix of EOS
  }

 private:
  struct Node { int child[2]; int sym; };
  std::vector<Node> nodes_;
};

inline const HuffmanDecoder& huffman() { static HuffmanDecoder d; return d; }

class HpackDecoder {
 public:
  bool decode(const uint8_t* p, size_t n, Headers* out) {
    size_t i = 0;
    while (i < n) {
      const uint8_t b = p[i];
      if (b & 0x80) {                                   // indexed header field
        uint64_t idx; if (!integer(p, n, &i, 7, &idx)) return false;
        std::string name, value; if (!lookup(idx, &name, &value)) return false;
        out->emplace_back(name, value);
      } else if (b & 0x40) {                            // literal with incremental indexing
        std::string name, value; if (!literal(p, n, &i, 6, &name, &value)) return false;
        out->emplace_back(name, value); add(name, value);
      } else if (b & 0x20) {                            // dynamic table size update
        uint64_t sz; if (!integer(p, n, &i, 5, &sz)) return false;
        if (sz > kSettingsTableSize) return false;      // RFC 7541 6.3: must not exceed SETTINGS_HEADER_TABLE_SIZE (we keep the 4096 default)
        max_size_ = (size_t)sz; evict();
      } else {                                          // literal without indexing / never indexed
        std::string name, value; if (!literal(p, n, &i, 4, &name, &value)) return false;
        out->emplace_back(name, value);
      }
    }
    return true;
  }

 private:
  static bool integer(const uint8_t* p, size_t n, size_t* i, int prefix, uint64_t* out) {
    if (*i >= n) return false;
    const uint64_t mask = (1u << prefix) - 1;
    uint64_t v = p[(*i)++] & mask;
    if (v < mask) { *out = v; return true; }
    int shift = 0;
    while (*i < n) {
      const uint8_t b = p[(*i)++];
      v += (uint64_t)(b & 0x7F) << shift; shift += 7;
      if (!(b & 0x80)) { *out = v; return true; }
      if (shift > 56) return false;
    }
    return false;
  }
  static bool string(const uint8_t* p, size_t n, size_t* i, std::string* out) {
    if (*i >= n) return false;
    const bool huff = p[*i] & 0x80;
    uint64_t len; if (!integer(p, n, i, 7, &len)) return false;
    if (len > n - *i) return false;                   // (not `*i + len > n`: a hostile length must not wrap)
    if (huff) { if (!huffman().decode(p + *i, (size_t)len, out)) return false; }
    else out->assign(reinterpret_cast<const char*>(p + *i), (size_t)len);
    *i += (size_t)len;
    return true;
  }
  bool literal(const uint8_t* p, size_t n, size_t* i, int prefix, std::string* name, std::string* value) {
    uint64_t idx; if (!integer(p, n, i, prefix, &idx)) return false;
    if (idx) { std::string v; if (!lookup(idx, name, &v)) return false; }
    else if (!string(p, n, i, name)) return false;
    return string(p, n, i, value);
  }
  bool lookup(uint64_t idx, std::string* name, std::string* value) const {
    if (idx == 0) return false;
    if (idx <= 61) { *name = kStaticTable[idx - 1][0]; *value = kStaticTable[idx - 1][1]; return true; }
    const size_t d = (size_t)(idx - 62);
    if (d >= dyn_.size()) return false;
    *name = dyn_[d].first; *value = dyn_[d].second;
    return true;
  }
  void add(const std::string& n, const std::string& v) {
    dyn_.insert(dyn_.begin(), {n, v}); size_ += n.size() + v.size() + 32; evict();
  }
  void evict() { while (size_ > max_size_ && !dyn_.empty()) { size_ -= dyn_.back().first.size() + dyn_.back().second.size() + 32; dyn_.pop_back(); } }
  std::vector<std::pair<std::string, std::string>> dyn_;
  static constexpr uint64_t kSettingsTableSize = 4096;
  size_t size_ = 0, max_size_ = kSettingsTableSize;
};

// Encoder: "literal header field without indexing, new name", raw strings — always valid, no shared state to get wrong.
inline void hpack_int(std::string* out, uint8_t first, int prefix, uint64_t v) {
  const uint64_t mask = (1u << prefix) - 1;
  if (v < mask) { out->push_back((char)(first | v)); return; }
  out->push_back((char)(first | mask)); v -= mask;
  while (v >= 128) { out->push_back((char)(0x80 | (v & 0x7F))); v >>= 7; }
  out->push_back((char)v);
}
inline std::string hpack_encode(const Headers& hs) {
  std::string out;
  for (const auto& h : hs) {
    out.push_back(0x00);
    hpack_int(&out, 0x00, 7, h.first.size()); out += h.first;
    hpack_int(&out, 0x00, 7, h.second.size()); out += h.second;
  }
  return out;
}

// ------------------------------------------------------------------------------------------------ socket + frames
inline bool read_exact(int fd, void* buf, size_t n) {
  char* p = static_cast<char*>(buf);
  while (n) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r == 0) return false;
    if (r < 0) { if (errno == EINTR) continue; return false; }
    p += r; n -= (size_t)r;
  }
  return true;
}
inline bool write_all(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n) {
    ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
    if (r < 0) { if (errno == EINTR) continue; return false; }
    p += r; n -= (size_t)r;
  }
  return true;
}

struct Frame { uint8_t type = 0, flags = 0; uint32_t stream = 0; std::string payload; };

inline bool read_frame(int fd, Frame* f, size_t max_len = 1 << 24) {
  uint8_t h[9];
  if (!read_exact(fd, h, 9)) return false;
  const size_t len = ((size_t)h[0] << 16) | ((size_t)h[1] << 8) | h[2];
  if (len > max_len) return false;
  f->type = h[3]; f->flags = h[4];
  f->stream = (((uint32_t)h[5] << 24) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 8) | h[8]) & 0x7FFFFFFFu;
  f->payload.resize(len);
  return len == 0 || read_exact(fd, &f->payload[0], len);
}
inline std::string frame_bytes(uint8_t type, uint8_t flags, uint32_t stream, const std::string& payload) {
  std::string out(9, '\0');
  out[0] = (char)(payload.size() >> 16); out[1] = (char)(payload.size() >> 8); out[2] = (char)payload.size();
  out[3] = (char)type; out[4] = (char)flags;
  out[5] = (char)(stream >> 24); out[6] = (char)(stream >> 16); out[7] = (char)(stream >> 8); out[8] = (char)stream;
  return out + payload;
}
inline std::string u32be(uint32_t v) { std::string s(4, '\0'); s[0] = (char)(v >> 24); s[1] = (char)(v >> 16); s[2] = (char)(v >> 8); s[3] = (char)v; return s; }
inline uint32_t get_u32be(const std::string& s, size_t off) { return ((uint32_t)(uint8_t)s[off] << 24) | ((uint32_t)(uint8_t)s[off + 1] << 16) | ((uint32_t)(uint8_t)s[off + 2] << 8) | (uint8_t)s[off + 3]; }

inline std::string grpc_message(const std::string& pb) { std::string m(1, '\0'); m += u32be((uint32_t)pb.size()); return m + pb; }
// Splits a stream's accumulated DATA bytes into complete gRPC messages; leaves a partial tail in *buf.
inline bool grpc_split(std::string* buf, std::vector<std::string>* msgs) {
  size_t off = 0;
  while (buf->size() - off >= 5) {
    if ((*buf)[off] != 0) return false;                // compressed messages are not negotiated
    const uint32_t len = get_u32be(*buf, off + 1);
    if (buf->size() - off - 5 < len) break;
    msgs->push_back(buf->substr(off + 5, len)); off += 5 + len;
  }
  buf->erase(0, off);
  return true;
}
inline std::string percent_encode(const std::string& s) {
  static const char* hex = "0123456789ABCDEF";
  std::string out;
  for (unsigned char c : s) { if (c >= 0x20 && c < 0x7F && c != '%') out.push_back((char)c); else { out.push_back('%'); out.push_back(hex[c >> 4]); out.push_back(hex[c & 15]); } }
  return out;
}

// ------------------------------------------------------------------------------------------------ server
struct Status { int code = 0; std::string message; };    // gRPC status codes: 0 OK, 2 UNKNOWN, 12 UNIMPLEMENTED, 13 INTERNAL

class ServerStream {            // handed to streaming handlers
 public:
  virtual ~ServerStream() = default;
  virtual bool send(const std::string& pb) = 0;           // false once the client went away
  virtual bool cancelled() const = 0;
};

using UnaryHandler = std::function<Status(const std::string& request, std::string* response)>;
using StreamHandler = std::function<Status(const std::string& request, ServerStream* stream)>;

class Server {
 public:
  void add_unary(const std::string& path, UnaryHandler h) { unary_[path] = std::move(h); }
  void add_stream(const std::string& path, StreamHandler h) { stream_[path] = std::move(h); }

  bool listen_unix(const std::string& path, std::string* err) {
    ::unlink(path.c_str());
    fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd_ < 0) { *err = strerror(errno); return false; }
    sockaddr_un a{}; a.sun_family = AF_UNIX;
    if (path.size() >= sizeof(a.sun_path)) { *err = "socket path too long"; return false; }
    strcpy(a.sun_path, path.c_str());
    if (::bind(fd_, reinterpret_cast<sockaddr*>(&a), sizeof(a)) < 0 || ::listen(fd_, 16) < 0) { *err = strerror(errno); ::close(fd_); fd_ = -1; return false; }
    stop_ = false;
    accept_thread_ = std::thread([this] { accept_loop(); });
    return true;
  }
  void stop() {
    stop_ = true;
    if (fd_ >= 0) ::shutdown(fd_, SHUT_RDWR);           // wakes accept4(); the descriptor stays valid until the thread is gone
    if (accept_thread_.joinable()) accept_thread_.join();
    if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
    std::vector<std::shared_ptr<Conn>> conns;
    { std::lock_guard<std::mutex> lk(mu_); conns.swap(conns_); }
    for (auto& c : conns) { c->close(); if (c->thread.joinable()) c->thread.join(); }
  }
  bool stopping() const { return stop_; }
  ~Server() { stop(); }

 private:
  struct Conn;
  struct StreamState : ServerStream {
    Conn* conn = nullptr; uint32_t id = 0; std::string path, data; std::atomic<bool> dead{false}, finished{false}; bool headers_sent = false, dispatched = false;
    int64_t send_window = 65535;
    std::thread worker;
    bool send(const std::string& pb) override;
    bool cancelled() const override;
  };
  struct Conn {
    Server* srv = nullptr; int fd = -1; std::thread thread; std::mutex wmu; std::condition_variable wcv; std::atomic<bool> closed{false}, done{false};
    ~Conn() { if (fd >= 0) ::close(fd); }     // only after `thread` was joined: close() from another thread may still shutdown(fd) until then
    int64_t conn_window = 65535; uint32_t peer_initial_window = 65535, peer_max_frame = 16384;
    HpackDecoder dec; std::map<uint32_t, std::shared_ptr<StreamState>> streams;
    void close() { closed = true; if (fd >= 0) ::shutdown(fd, SHUT_RDWR); wcv.notify_all(); }
    bool write(const std::string& bytes) { std::lock_guard<std::mutex> lk(wmu); return !closed && write_all(fd, bytes.data(), bytes.size()); }
    // DATA honouring both flow-control windows and the peer's max frame size
    bool write_data(StreamState* s, const std::string& bytes) {
      size_t off = 0;
      while (off < bytes.size()) {
        std::unique_lock<std::mutex> lk(wmu);
        if (!wcv.wait_for(lk, std::chrono::seconds(30), [&] { return closed || s->dead || (conn_window > 0 && s->send_window > 0); })) return false;
        if (closed || s->dead) return false;
        const size_t n = std::min<size_t>({bytes.size() - off, (size_t)conn_window, (size_t)s->send_window, (size_t)peer_max_frame});
        if (!write_all(fd, frame_bytes(DATA, 0, s->id, bytes.substr(off, n)).data(), 9 + n)) return false;
        conn_window -= (int64_t)n; s->send_window -= (int64_t)n; off += n;
      }
      return true;
    }
  };

  void accept_loop() {
    const int lfd = fd_;                                   // set before this thread was started; stop() closes it only after joining us
    while (!stop_) {
      int c = ::accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
      if (c < 0) { if (errno == EINTR) continue; break; }
      auto conn = std::make_shared<Conn>();
      conn->srv = this; conn->fd = c;
      std::vector<std::shared_ptr<Conn>> finished;
      {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto it = conns_.begin(); it != conns_.end();) { if ((*it)->done) { finished.push_back(*it); it = conns_.erase(it); } else ++it; }
        conns_.push_back(conn);
        conn->thread = std::thread([this, conn] { serve(conn.get()); });     // under mu_: stop() must never see a Conn without its thread
      }
      for (auto& f : finished) if (f->thread.joinable()) f->thread.join();     // connections that ended since the last accept: join, then ~Conn closes the fd
    }
  }

  static void send_headers(Conn* c, StreamState* s) {
    if (s->headers_sent) return;
    s->headers_sent = true;
    c->write(frame_bytes(HEADERS, END_HEADERS, s->id, hpack_encode({{":status", "200"}, {"content-type", "application/grpc"}})));
  }
  static void send_trailers(Conn* c, StreamState* s, const Status& st) {
    Headers h;
    if (!s->headers_sent) { h.push_back({":status", "200"}); h.push_back({"content-type", "application/grpc"}); s->headers_sent = true; }   // trailers-only
    h.push_back({"grpc-status", std::to_string(st.code)});
    if (!st.message.empty()) h.push_back({"grpc-message", percent_encode(st.message)});
    c->write(frame_bytes(HEADERS, END_HEADERS | END_STREAM, s->id, hpack_encode(h)));
  }

  // Streams that answered (unary) or whose handler returned (server-streaming) leave the table, so a kubelet connection that
  // lives for months does not accumulate one entry per Allocate.
  static void forget(Conn* c, uint32_t id) { std::lock_guard<std::mutex> lk(c->wmu); c->streams.erase(id); }
  static void reap_finished(Conn* c) {
    std::vector<std::shared_ptr<StreamState>> done;
    { std::lock_guard<std::mutex> lk(c->wmu);
      for (auto it = c->streams.begin(); it != c->streams.end();) { if (it->second->finished) { done.push_back(it->second); it = c->streams.erase(it); } else ++it; } }
    for (auto& s : done) if (s->worker.joinable()) s->worker.join();
  }

  void dispatch(Conn* c, std::shared_ptr<StreamState> s) {
    if (s->dispatched) return;                       // a second END_STREAM on the same stream: ignore (never start a second worker)
    s->dispatched = true;
    std::vector<std::string> msgs;
    if (!grpc_split(&s->data, &msgs) || msgs.size() != 1) { send_trailers(c, s.get(), {13, "malformed gRPC request"}); forget(c, s->id); return; }
    auto u = unary_.find(s->path);
    if (u != unary_.end()) {
      std::string resp;
      Status st = u->second(msgs[0], &resp);
      if (st.code == 0) { send_headers(c, s.get()); c->write_data(s.get(), grpc_message(resp)); }
      send_trailers(c, s.get(), st);
      forget(c, s->id);
      return;
    }
    auto sh = stream_.find(s->path);
    if (sh != stream_.end()) {
      StreamHandler h = sh->second;
      std::string req = msgs[0];
      s->worker = std::thread([this, c, s, h, req] {
        send_headers(c, s.get());
        Status st = h(req, s.get());
        if (!s->dead && !c->closed) send_trailers(c, s.get(), st);
        s->finished = true;                          // joined and dropped by reap_finished() or at connection end
      });
      return;
    }
    send_trailers(c, s.get(), {12, "unknown method " + s->path});
    forget(c, s->id);
  }

  void serve(Conn* c) {
    char pre[24];
    if (!read_exact(c->fd, pre, 24) || memcmp(pre, kPreface, 24) != 0) { c->close(); c->done = true; return; }
    c->write(frame_bytes(SETTINGS, 0, 0, ""));
    Frame f;
    std::string header_block; uint32_t header_stream = 0; uint8_t header_flags = 0;
    while (!c->closed && read_frame(c->fd, &f, kMaxFrame)) {
      switch (f.type) {
        case SETTINGS:
          if (f.flags & ACK) break;
          for (size_t i = 0; i + 6 <= f.payload.size(); i += 6) {
            const uint16_t id = (uint16_t)(((uint8_t)f.payload[i] << 8) | (uint8_t)f.payload[i + 1]);
            const uint32_t v = get_u32be(f.payload, i + 2);
            std::lock_guard<std::mutex> lk(c->wmu);
            if (id == 4) { for (auto& kv : c->streams) kv.second->send_window += (int64_t)v - (int64_t)c->peer_initial_window; c->peer_initial_window = v; }
            if (id == 5 && v >= 16384) c->peer_max_frame = v;
          }
          c->write(frame_bytes(SETTINGS, ACK, 0, ""));
          c->wcv.notify_all();
          break;
        case PING:
          if (!(f.flags & ACK)) c->write(frame_bytes(PING, ACK, 0, f.payload));
          break;
        case WINDOW_UPDATE: {
          if (f.payload.size() != 4) break;
          const uint32_t inc = get_u32be(f.payload, 0) & 0x7FFFFFFFu;
          { std::lock_guard<std::mutex> lk(c->wmu);
            if (f.stream == 0) c->conn_window += inc;
            else { auto it = c->streams.find(f.stream); if (it != c->streams.end()) it->second->send_window += inc; } }
          c->wcv.notify_all();
          break;
        }
        case HEADERS: {
          size_t off = 0, pad = 0;
          if (f.flags & PADDED) { if (f.payload.empty()) break; pad = (uint8_t)f.payload[0]; off = 1; }
          if (f.flags & PRIORITY_FLAG) off += 5;
          if (off + pad > f.payload.size()) break;
          header_block = f.payload.substr(off, f.payload.size() - off - pad);
          header_stream = f.stream; header_flags = f.flags;
          if (header_block.size() > kMaxHeaderBlock) { goaway(c, 11); break; }
          if (f.flags & END_HEADERS) finish_headers(c, header_stream, header_flags, &header_block);
          break;
        }
        case CONTINUATION:
          if (f.stream != header_stream || header_stream == 0) break;
          if (header_block.size() + f.payload.size() > kMaxHeaderBlock) { goaway(c, 11); break; }     // ENHANCE_YOUR_CALM
          header_block += f.payload;
          if (f.flags & END_HEADERS) finish_headers(c, header_stream, header_flags, &header_block);
          break;
        case DATA: {
          std::shared_ptr<StreamState> s;
          { std::lock_guard<std::mutex> lk(c->wmu); auto it = c->streams.find(f.stream); if (it != c->streams.end()) s = it->second; }
          size_t off = 0, pad = 0;
          if (f.flags & PADDED) { if (f.payload.empty()) break; pad = (uint8_t)f.payload[0]; off = 1; }
          if (s && s->dispatched) s.reset();                                                     // data after END_STREAM: drop
          if (s && s->data.size() + f.payload.size() > kMaxRequest) {                            // requests here are a few hundred bytes
            c->write(frame_bytes(RST_STREAM, 0, f.stream, u32be(11)));
            s->dead = true; forget(c, f.stream); s.reset();
          }
          if (s && off + pad <= f.payload.size()) s->data.append(f.payload, off, f.payload.size() - off - pad);
          if (!f.payload.empty()) {   // give the credit straight back: requests here are tiny
            c->write(frame_bytes(WINDOW_UPDATE, 0, 0, u32be((uint32_t)f.payload.size())));
            if (!(f.flags & END_STREAM)) c->write(frame_bytes(WINDOW_UPDATE, 0, f.stream, u32be((uint32_t)f.payload.size())));
          }
          if (s && (f.flags & END_STREAM)) dispatch(c, s);
          break;
        }
        case RST_STREAM: {
          std::lock_guard<std::mutex> lk(c->wmu);
          auto it = c->streams.find(f.stream);
          if (it != c->streams.end()) it->second->dead = true;
          c->wcv.notify_all();
          break;
        }
        case GOAWAY: c->close(); break;
        default: break;   // PRIORITY, unknown extension frames: ignore
      }
    }
    c->close();
    std::map<uint32_t, std::shared_ptr<StreamState>> streams;
    { std::lock_guard<std::mutex> lk(c->wmu); streams.swap(c->streams); }
    for (auto& kv : streams) { kv.second->dead = true; }
    c->wcv.notify_all();
    for (auto& kv : streams) if (kv.second->worker.joinable()) kv.second->worker.join();
    c->done = true;                                        // the descriptor is closed by ~Conn once this thread has been joined
  }

  void finish_headers(Conn* c, uint32_t stream, uint8_t flags, std::string* block) {
    Headers hs;
    if (!c->dec.decode(reinterpret_cast<const uint8_t*>(block->data()), block->size(), &hs)) { c->write(frame_bytes(GOAWAY, 0, 0, u32be(0) + u32be(9))); c->close(); return; }   // COMPRESSION_ERROR
    block->clear();
    reap_finished(c);
    auto s = std::make_shared<StreamState>();
    s->conn = c; s->id = stream;
    for (auto& h : hs) if (h.first == ":path") s->path = h.second;
    bool refuse = false, protocol_error = false;
    { std::lock_guard<std::mutex> lk(c->wmu);
      if (stream == 0 || (stream & 1) == 0 || c->streams.count(stream)) protocol_error = true;        // client streams are odd and opened once
      else if (c->streams.size() >= kMaxStreams) refuse = true;
      else { s->send_window = c->peer_initial_window; c->streams[stream] = s; } }
    if (protocol_error) { goaway(c, 1); return; }
    if (refuse) { c->write(frame_bytes(RST_STREAM, 0, stream, u32be(7))); return; }                    // REFUSED_STREAM
    if (flags & END_STREAM) dispatch(c, s);
  }
  static void goaway(Conn* c, uint32_t code) { c->write(frame_bytes(GOAWAY, 0, 0, u32be(0) + u32be(code))); c->close(); }

  static constexpr size_t kMaxFrame = 1u << 20, kMaxHeaderBlock = 64u << 10, kMaxRequest = 4u << 20, kMaxStreams = 256;

  int fd_ = -1;
  std::atomic<bool> stop_{true};
  std::thread accept_thread_;
  std::mutex mu_;
  std::vector<std::shared_ptr<Conn>> conns_;
  std::map<std::string, UnaryHandler> unary_;
  std::map<std::string, StreamHandler> stream_;
};

inline bool Server::StreamState::send(const std::string& pb) { return !dead && !conn->closed && conn->write_data(this, grpc_message(pb)); }
inline bool Server::StreamState::cancelled() const { return dead || conn->closed || conn->srv->stopping(); }

// ------------------------------------------------------------------------------------------------ unary client
// Returns the gRPC status (-1 on transport failure with *err set); *response gets the first response message.
inline int unary_call(const std::string& socket_path, const std::string& path, const std::string& request, std::string* response, std::string* err, int timeout_ms = 10000) {
  int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) { *err = strerror(errno); return -1; }
  sockaddr_un a{}; a.sun_family = AF_UNIX;
  if (socket_path.size() >= sizeof(a.sun_path)) { *err = "socket path too long"; ::close(fd); return -1; }
  strcpy(a.sun_path, socket_path.c_str());
  timeval tv{timeout_ms / 1000, (timeout_ms % 1000) * 1000};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv)); setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
  if (::connect(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) < 0) { *err = std::string("connect ") + socket_path + ": " + strerror(errno); ::close(fd); return -1; }
  std::string out(kPreface, 24);
  out += frame_bytes(SETTINGS, 0, 0, "");
  out += frame_bytes(HEADERS, END_HEADERS, 1, hpack_encode({{":method", "POST"}, {":scheme", "http"}, {":path", path}, {":authority", "localhost"}, {"content-type", "application/grpc"},
                                                           {"te", "trailers"}, {"user-agent", "b200-device-plugin/0.1"}}));
  out += frame_bytes(DATA, END_STREAM, 1, grpc_message(request));
  if (!write_all(fd, out.data(), out.size())) { *err = "write failed"; ::close(fd); return -1; }
  HpackDecoder dec;
  std::string data, block; int status = -1; std::string message; bool done = false;
  Frame f;
  while (!done && read_frame(fd, &f)) {
    if (f.type == SETTINGS && !(f.flags & ACK)) { std::string ack = frame_bytes(SETTINGS, ACK, 0, ""); write_all(fd, ack.data(), ack.size()); }
    else if (f.type == PING && !(f.flags & ACK)) { std::string pong = frame_bytes(PING, ACK, 0, f.payload); write_all(fd, pong.data(), pong.size()); }
    else if (f.type == DATA && f.stream == 1) { data += f.payload; if (f.flags & END_STREAM) done = true; }
    else if ((f.type == HEADERS || f.type == CONTINUATION) && f.stream == 1) {
      size_t off = 0, pad = 0;
      if (f.type == HEADERS) { if (f.flags & PADDED) { pad = (uint8_t)f.payload[0]; off = 1; } if (f.flags & PRIORITY_FLAG) off += 5; }
      block += f.payload.substr(off, f.payload.size() - off - pad);
      if (f.flags & END_HEADERS) {
        Headers hs;
        if (!dec.decode(reinterpret_cast<const uint8_t*>(block.data()), block.size(), &hs)) { *err = "HPACK decode failed"; ::close(fd); return -1; }
        block.clear();
        for (auto& h : hs) { if (h.first == "grpc-status") status = atoi(h.second.c_str()); if (h.first == "grpc-message") message = h.second; }
      }
      if (f.type == HEADERS && (f.flags & END_STREAM)) done = true;
    } else if (f.type == RST_STREAM && f.stream == 1) { *err = "stream reset by peer"; ::close(fd); return -1; }
    else if (f.type == GOAWAY) {
      // A server that shuts down gracefully (a kubelet restarting) announces it with GOAWAY(NO_ERROR, last-stream-id) and still answers the
      // streams up to that id: ours is stream 1, so keep reading. Anything else ends the call.
      auto be32 = [&](size_t o) { return f.payload.size() >= o + 4 ? ((uint32_t)(uint8_t)f.payload[o] << 24) | ((uint32_t)(uint8_t)f.payload[o + 1] << 16) | ((uint32_t)(uint8_t)f.payload[o + 2] << 8) | (uint8_t)f.payload[o + 3] : 0u; };
      const uint32_t last = be32(0) & 0x7FFFFFFFu, code = f.payload.size() >= 8 ? be32(4) : 2u;
      if (code != 0 || last < 1) { *err = "GOAWAY from peer (error code " + std::to_string(code) + ")"; break; }
    }
  }
  ::close(fd);
  if (status < 0) { if (err->empty()) *err = "connection closed before a grpc-status arrived"; return -1; }
  std::vector<std::string> msgs;
  grpc_split(&data, &msgs);
  if (!msgs.empty()) *response = msgs[0];
  if (status != 0) *err = message;
  return status;
}

}  // namespace h2
