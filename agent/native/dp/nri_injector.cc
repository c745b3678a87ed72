// b200-nri-device-injector — native containerd NRI plugin (static C++ binary; the reference's is a CGO-off Go binary).
//
// Contract: reference nri_device_injector/nri_device_injector.go:30-199 — plugin `device_injector_nri`,
// index `10`, socket /var/run/nri/nri.sock; on CreateContainer read the pod annotation
// `devices.gke.io/container.<name>` (a YAML list of {path, type, major, minor, file_mode, uid, gid}), honour only `path` and
// non-zero file_mode/uid/gid, re-derive type/major/minor with lstat, first duplicate path wins, any error fails the
// container creation, pod == nil is a no-op.
// Wire stack (nothing in the image speaks it): NRI multiplexes two ttrpc connections over one Unix socket
// (8-byte frames: conn id + length, big endian; conn 1 = Plugin service served here, conn 2 = Runtime service called here);
// ttrpc = 10-byte header (length, stream id, type, flags) + protobuf Request{service,method,payload} / Response{status,payload}.
// Same behaviour and tests as container_engine_accelerators_b200/agent/nri.py (tests/test_nri_injector.py runs both).
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/sysmacros.h>
#include <sys/un.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "pb.hpp"

namespace {

const char* kPluginService = "nri.pkg.api.v1alpha1.Plugin";
const char* kRuntimeService = "nri.pkg.api.v1alpha1.Runtime";
const uint32_t kPluginConn = 1, kRuntimeConn = 2;
const uint8_t kRequest = 1, kResponse = 2;

void logf(char level, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  fprintf(stderr, "%c b200-nri-device-injector] ", level); vfprintf(stderr, fmt, ap); fputc('\n', stderr);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------ annotation parsing
struct Dev { std::string path; unsigned long file_mode = 0, uid = 0, gid = 0; };

std::string trim(const std::string& s) { size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r"); return a == std::string::npos ? "" : s.substr(a, b - a + 1); }
std::string unquote(std::string v) { if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\''))) v = v.substr(1, v.size() - 2); return v; }

// The annotation is always a block sequence of flat mappings ("- path: /dev/x\n  uid: 7\n") or its flow form
// ("[{path: /dev/x}]"); anything else is an error, as a YAML type mismatch is in the reference.
bool parse_devices(const std::string& text, std::vector<Dev>* out, std::string* err) {
  std::vector<std::map<std::string, std::string>> items;
  std::string t = trim(text);
  // `key: value` with either side optionally quoted (JSON is YAML: {"path": "/dev/x"}) and a trailing " # comment" dropped
  auto strip_comment = [](const std::string& v) {
    char quote = 0;
    for (size_t i = 0; i < v.size(); i++) {
      if (quote) { if (v[i] == quote) quote = 0; continue; }
      if (v[i] == '"' || v[i] == '\'') quote = v[i];
      else if (v[i] == '#' && (i == 0 || v[i - 1] == ' ' || v[i - 1] == '\t')) return v.substr(0, i);
    }
    return v;
  };
  auto add_kv = [&](std::map<std::string, std::string>* m, const std::string& kv_in) -> bool {
    const std::string kv = trim(strip_comment(kv_in));
    size_t c = std::string::npos;
    if (!kv.empty() && (kv[0] == '"' || kv[0] == '\'')) { size_t e = kv.find(kv[0], 1); if (e != std::string::npos) c = kv.find(':', e); }   // colon after the quoted key
    else c = kv.find(':');
    if (c == std::string::npos) return false;
    (*m)[unquote(trim(kv.substr(0, c)))] = unquote(trim(kv.substr(c + 1)));
    return true;
  };
  auto to_uint = [](const std::string& v) -> unsigned long {          // 438, 0666 (YAML 1.1 octal), 0o666 (YAML 1.2), 0x1b6
    if (v.size() > 2 && v[0] == '0' && (v[1] == 'o' || v[1] == 'O')) return strtoul(v.c_str() + 2, nullptr, 8);
    return strtoul(v.c_str(), nullptr, 0);
  };
  if (t.empty()) return true;
  if (t[0] == '[') {
    if (t.back() != ']') { *err = "unterminated flow sequence"; return false; }
    std::string body = t.substr(1, t.size() - 2);
    size_t i = 0;
    while ((i = body.find('{', i)) != std::string::npos) {
      size_t j = body.find('}', i);
      if (j == std::string::npos) { *err = "unterminated flow mapping"; return false; }
      std::map<std::string, std::string> m; std::string inner = body.substr(i + 1, j - i - 1); size_t p = 0;
      while (p <= inner.size()) { size_t q = inner.find(',', p); std::string kv = trim(inner.substr(p, q == std::string::npos ? std::string::npos : q - p)); if (!kv.empty() && !add_kv(&m, kv)) { *err = "bad flow mapping entry"; return false; } if (q == std::string::npos) break; p = q + 1; }
      items.push_back(m); i = j + 1;
    }
  } else {
    size_t p = 0; bool any = false;
    while (p <= t.size()) {
      size_t q = t.find('\n', p);
      std::string line = t.substr(p, q == std::string::npos ? std::string::npos : q - p);
      p = q == std::string::npos ? t.size() + 1 : q + 1;
      std::string s = trim(line);
      if (s.empty() || s[0] == '#') continue;
      if (s[0] == '-') { items.emplace_back(); any = true; s = trim(s.substr(1)); if (s.empty()) continue; }
      else if (!any) { *err = "expected a YAML sequence of device mappings"; return false; }
      if (s.find('[') != std::string::npos && s.find(']') == std::string::npos) { *err = "unterminated flow sequence"; return false; }
      if (!add_kv(&items.back(), s)) { *err = "expected `key: value`"; return false; }
    }
    if (!any) { *err = "expected a YAML sequence of device mappings"; return false; }
  }
  std::set<std::string> seen;
  for (auto& m : items) {
    Dev d; d.path = m.count("path") ? m["path"] : "";
    if (seen.count(d.path)) continue;                        // duplicate path: first wins
    seen.insert(d.path);
    d.file_mode = m.count("file_mode") ? to_uint(m["file_mode"]) : 0;
    d.uid = m.count("uid") ? to_uint(m["uid"]) : 0;
    d.gid = m.count("gid") ? to_uint(m["gid"]) : 0;
    out->push_back(d);
  }
  return true;
}

// LinuxDevice{path=1,type=2,major=3,minor=4,file_mode=5{value=1},uid=6{value=1},gid=7{value=1}}
bool encode_device(const Dev& d, std::string* out, std::string* err) {
  struct stat st;
  if (lstat(d.path.c_str(), &st) != 0) { *err = "failed to get info from device path " + d.path + ": " + strerror(errno); return false; }
  const char* type = S_ISBLK(st.st_mode) ? "b" : S_ISCHR(st.st_mode) ? "c" : S_ISFIFO(st.st_mode) ? "p" : nullptr;
  if (!type) { *err = "invalid device type " + std::to_string(st.st_mode) + " from device path " + d.path; return false; }
  pb::put_string(out, 1, d.path); pb::put_string(out, 2, type);
  pb::put_int(out, 3, (int64_t)major(st.st_rdev)); pb::put_int(out, 4, (int64_t)minor(st.st_rdev));
  auto opt = [&](int field, unsigned long v) { if (v) { std::string o; pb::put_tag(&o, 1, 0); pb::put_varint(&o, v); pb::put_bytes(out, field, o); } };
  opt(5, d.file_mode); opt(6, d.uid); opt(7, d.gid);
  return true;
}

// CreateContainerRequest{pod=1{name=2,namespace=4,annotations=6 map},container=2{name=3}} -> CreateContainerResponse{adjust=1{linux=6{devices=1 repeated}}}
bool create_container(const std::string& req, std::string* resp, std::string* err) {
  std::vector<pb::Field> top;
  if (!pb::parse(req, &top)) { *err = "malformed CreateContainerRequest"; return false; }
  bool have_pod = false; std::map<std::string, std::string> ann; std::string ctr, pod_name;
  for (auto& f : top) {
    if (f.number == 1 && f.wire_type == 2) {
      have_pod = true;
      std::vector<pb::Field> pf; if (!pb::parse(f.bytes, &pf)) { *err = "malformed PodSandbox"; return false; }
      for (auto& p : pf) {
        if (p.number == 2) pod_name = p.bytes;
        if (p.number == 6 && p.wire_type == 2) { std::vector<pb::Field> e; if (!pb::parse(p.bytes, &e)) continue; std::string k, v; for (auto& x : e) { if (x.number == 1) k = x.bytes; if (x.number == 2) v = x.bytes; } ann[k] = v; }
      }
    } else if (f.number == 2 && f.wire_type == 2) {
      std::vector<pb::Field> cf; if (!pb::parse(f.bytes, &cf)) { *err = "malformed Container"; return false; }
      for (auto& c : cf) if (c.number == 3) ctr = c.bytes;
    }
  }
  std::string linux_adj;
  if (have_pod) {
    auto it = ann.find("devices.gke.io/container." + ctr);
    if (it != ann.end()) {
      std::vector<Dev> devs; std::string perr;
      if (!parse_devices(it->second, &devs, &perr)) { *err = "invalid device annotation \"devices.gke.io/container." + ctr + "\": " + perr; return false; }
      for (auto& d : devs) {
        logf('I', "Annotated device %s (container=%s pod=%s)", d.path.c_str(), ctr.c_str(), pod_name.c_str());
        std::string enc; if (!encode_device(d, &enc, err)) return false;
        pb::put_bytes(&linux_adj, 1, enc);
      }
    }
  }
  std::string adjust;
  if (!linux_adj.empty()) pb::put_bytes(&adjust, 6, linux_adj);
  pb::put_bytes(resp, 1, adjust);
  return true;
}

// ------------------------------------------------------------------------------------------------ mux + ttrpc
bool read_exact(int fd, void* buf, size_t n) { char* p = (char*)buf; while (n) { ssize_t r = ::recv(fd, p, n, 0); if (r <= 0) { if (r < 0 && errno == EINTR) continue; return false; } p += r; n -= (size_t)r; } return true; }
bool write_all(int fd, const void* buf, size_t n) { const char* p = (const char*)buf; while (n) { ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL); if (r < 0) { if (errno == EINTR) continue; return false; } p += r; n -= (size_t)r; } return true; }
std::string be32(uint32_t v) { std::string s(4, 0); s[0] = (char)(v >> 24); s[1] = (char)(v >> 16); s[2] = (char)(v >> 8); s[3] = (char)v; return s; }
uint32_t rd32(const char* p) { return ((uint32_t)(uint8_t)p[0] << 24) | ((uint32_t)(uint8_t)p[1] << 16) | ((uint32_t)(uint8_t)p[2] << 8) | (uint8_t)p[3]; }

class Mux {
 public:
  explicit Mux(int fd) : fd_(fd) { reader_ = std::thread([this] { loop(); }); }
  ~Mux() { close(); if (reader_.joinable()) reader_.join(); }
  void close() { mark_closed(); ::shutdown(fd_, SHUT_RDWR); }
  bool write(uint32_t conn, const std::string& data) { std::lock_guard<std::mutex> lk(wmu_); std::string f = be32(conn) + be32((uint32_t)data.size()) + data; return write_all(fd_, f.data(), f.size()); }
  bool read(uint32_t conn, size_t n, std::string* out) {   // blocking byte-stream read on one logical connection
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return closed_ || buf_[conn].size() >= n; });
    if (buf_[conn].size() < n) return false;
    out->assign(buf_[conn].begin(), buf_[conn].begin() + (long)n);
    buf_[conn].erase(buf_[conn].begin(), buf_[conn].begin() + (long)n);
    return true;
  }

 private:
  void loop() {
    char hdr[8];
    while (read_exact(fd_, hdr, 8)) {
      const uint32_t conn = rd32(hdr), len = rd32(hdr + 4);
      if (len > (1u << 24)) break;
      std::string payload(len, 0);
      if (len && !read_exact(fd_, &payload[0], len)) break;
      { std::lock_guard<std::mutex> lk(mu_); buf_[conn].insert(buf_[conn].end(), payload.begin(), payload.end()); }
      cv_.notify_all();
    }
    mark_closed();
  }
  // under mu_: a reader that has just evaluated the wait predicate must not miss the wake-up
  void mark_closed() { { std::lock_guard<std::mutex> lk(mu_); closed_ = true; } cv_.notify_all(); }
  int fd_; std::thread reader_; std::mutex mu_, wmu_; std::condition_variable cv_; bool closed_ = false;
  std::map<uint32_t, std::deque<char>> buf_;
};

bool ttrpc_write(Mux* m, uint32_t conn, uint32_t stream, uint8_t type, const std::string& payload) {
  std::string h = be32((uint32_t)payload.size()) + be32(stream); h.push_back((char)type); h.push_back(0);
  return m->write(conn, h + payload);
}
bool ttrpc_read(Mux* m, uint32_t conn, uint32_t* stream, uint8_t* type, std::string* payload) {
  std::string h; if (!m->read(conn, 10, &h)) return false;
  const uint32_t len = rd32(h.data()); *stream = rd32(h.data() + 4); *type = (uint8_t)h[8];
  payload->clear();
  return len == 0 || m->read(conn, len, payload);
}
std::string encode_response(const std::string& payload, int code, const std::string& msg) {
  std::string out;
  if (code) { std::string st; pb::put_tag(&st, 1, 0); pb::put_varint(&st, (uint64_t)code); pb::put_bytes(&st, 2, msg); pb::put_bytes(&out, 1, st); }
  pb::put_bytes(&out, 2, payload);
  return out;
}

}  // namespace

int main(int argc, char** argv) {
  std::string sock_path = "/var/run/nri/nri.sock", name = "device_injector_nri", idx = "10";
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    auto val = [&]() -> std::string { size_t eq = a.find('='); if (eq != std::string::npos) { std::string v = a.substr(eq + 1); a = a.substr(0, eq); return v; } return i + 1 < argc ? argv[++i] : ""; };
    if (a.rfind("socket", 0) == 0) sock_path = val();
    else if (a.rfind("name", 0) == 0) name = val();
    else if (a.rfind("idx", 0) == 0) idx = val();
    else if (a == "parse-annotation") {       // debugging aid: read an annotation value from stdin, print what would be injected (one `path mode uid gid` per line)
      std::string text; char buf[4096]; size_t n;
      while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) text.append(buf, n);
      std::vector<Dev> devs; std::string err;
      if (!parse_devices(text, &devs, &err)) { printf("ERR %s\n", err.c_str()); return 1; }
      for (auto& d : devs) printf("%s %lu %lu %lu\n", d.path.c_str(), d.file_mode, d.uid, d.gid);
      return 0;
    }
    else if (a == "h" || a == "help") {
      puts("b200-nri-device-injector: containerd NRI plugin. On CreateContainer it reads the pod annotation devices.gke.io/container.<name>\n"
           "(YAML/JSON list of {path, file_mode?, uid?, gid?}) and injects those device nodes; type/major/minor always come from lstat.\n"
           "  --socket PATH         NRI socket (default /var/run/nri/nri.sock)\n"
           "  --name NAME --idx NN  plugin name and index (default device_injector_nri, 10)\n"
           "  --parse-annotation    read an annotation value on stdin and print what would be injected");
      return 0;
    }
    else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
  }
  int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  sockaddr_un addr{}; addr.sun_family = AF_UNIX; strncpy(addr.sun_path, sock_path.c_str(), sizeof(addr.sun_path) - 1);
  if (fd < 0 || ::connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) < 0) { logf('E', "Failed to connect to NRI socket %s: %s", sock_path.c_str(), strerror(errno)); return 1; }
  Mux mux(fd);
  // Plugin service (conn 1): served on a thread so registration can proceed on conn 2
  std::thread server([&] {
    uint32_t stream; uint8_t type; std::string payload;
    while (ttrpc_read(&mux, kPluginConn, &stream, &type, &payload)) {
      if (type != kRequest) continue;
      std::vector<pb::Field> fs; std::string service, method, body;
      if (pb::parse(payload, &fs)) for (auto& f : fs) { if (f.number == 1) service = f.bytes; if (f.number == 2) method = f.bytes; if (f.number == 3) body = f.bytes; }
      std::string resp, err; int code = 0;
      if (service != kPluginService) { code = 12; err = "unimplemented " + service + "/" + method; }
      else if (method == "Configure") { pb::put_int(&resp, 2, 1 << 3); logf('I', "configured by the runtime; subscribed to CreateContainer"); }   // ConfigureResponse.events (field 2): CREATE_CONTAINER
      else if (method == "CreateContainer") { if (!create_container(body, &resp, &err)) { code = 2; logf('W', "CreateContainer failed: %s", err.c_str()); } }
      else if (method == "Synchronize" || method == "StateChange" || method == "Shutdown") { /* empty responses */ }
      else { code = 12; err = "unimplemented " + method; }
      ttrpc_write(&mux, kPluginConn, stream, kResponse, encode_response(resp, code, err));
    }
  });
  // Runtime.RegisterPlugin (conn 2)
  std::string reg; pb::put_string(&reg, 1, name); pb::put_string(&reg, 2, idx);
  std::string req; pb::put_string(&req, 1, kRuntimeService); pb::put_string(&req, 2, "RegisterPlugin"); pb::put_bytes(&req, 3, reg);
  int rc = 0;
  if (!ttrpc_write(&mux, kRuntimeConn, 1, kRequest, req)) { logf('E', "failed to send RegisterPlugin"); rc = 1; }
  else {
    uint32_t stream; uint8_t type; std::string payload;
    if (!ttrpc_read(&mux, kRuntimeConn, &stream, &type, &payload)) { logf('E', "connection closed during registration"); rc = 1; }
    else logf('I', "registered NRI plugin %s-%s", idx.c_str(), name.c_str());
  }
  server.join();                 // returns when the runtime closes the connection: the pod restarts the plugin
  logf('I', "NRI connection closed");
  return rc;
}
