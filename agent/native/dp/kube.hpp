// Minimal Kubernetes API client for the native node agent: HTTP/1.1 over TCP or TLS, one request per connection.
//
// What the device plugin needs from the API server is small (reference: pkg/gpu/nvidia/util/util.go:55-70 builds an
// in-cluster client-go clientset; health_check/health_checker.go:129-160,288-358,395-449 and
// version_visibility/version_visibility.go:67-86 are its only users): GET a Node, PUT its status, POST an Event and
// PATCH annotations with server-side apply. So instead of a client library this is ~300 lines:
//   * in-cluster config (KUBERNETES_SERVICE_HOST/PORT + the service-account token and CA bundle), overridable with
//     B200_KUBE_URL / B200_KUBE_TOKEN_FILE / B200_KUBE_CA_FILE (tests, out-of-cluster runs);
//   * TLS through the system's libssl.so.3, loaded with dlopen on first https use so the binary still starts (and
//     every non-Kubernetes feature still works) on an image without OpenSSL; the server certificate is verified
//     against the CA bundle and against the host name or IP address we dialled;
//   * responses: Content-Length, chunked, or read-to-close.
#pragma once
#include <arpa/inet.h>
#include <dlfcn.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/tcp.h>
#include <openssl/err.h>
#include <openssl/ssl.h>
#include <openssl/x509v3.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>

#include "json.hpp"

namespace kube {

inline std::string now_rfc3339() {
  time_t t = time(nullptr);
  struct tm tm;
  gmtime_r(&t, &tm);
  char b[32];
  strftime(b, sizeof b, "%Y-%m-%dT%H:%M:%SZ", &tm);
  return b;
}

// ---- libssl / libcrypto entry points, resolved lazily
struct Tls {
#define B200_TLS_FN(name) decltype(&::name) name = nullptr
  B200_TLS_FN(OPENSSL_init_ssl);
  B200_TLS_FN(TLS_client_method);
  B200_TLS_FN(SSL_CTX_new);
  B200_TLS_FN(SSL_CTX_free);
  B200_TLS_FN(SSL_CTX_load_verify_locations);
  B200_TLS_FN(SSL_CTX_set_default_verify_paths);
  B200_TLS_FN(SSL_CTX_set_verify);
  B200_TLS_FN(SSL_new);
  B200_TLS_FN(SSL_free);
  B200_TLS_FN(SSL_set_fd);
  B200_TLS_FN(SSL_connect);
  B200_TLS_FN(SSL_read);
  B200_TLS_FN(SSL_write);
  B200_TLS_FN(SSL_shutdown);
  B200_TLS_FN(SSL_ctrl);
  B200_TLS_FN(SSL_set1_host);
  B200_TLS_FN(SSL_get0_param);
  B200_TLS_FN(SSL_get_verify_result);
  B200_TLS_FN(X509_VERIFY_PARAM_set1_ip_asc);
  B200_TLS_FN(X509_verify_cert_error_string);
  B200_TLS_FN(ERR_get_error);
  B200_TLS_FN(ERR_error_string_n);
#undef B200_TLS_FN
  std::string error;
  bool ok = false;

  static Tls& get() {
    static Tls t;
    static std::once_flag once;
    std::call_once(once, [] { t.load(); });
    return t;
  }
  void load() {
    void* ssl = nullptr;
    void* crypto = nullptr;
    for (const char* n : {"libssl.so.3", "libssl.so.1.1", "libssl.so"}) if ((ssl = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    for (const char* n : {"libcrypto.so.3", "libcrypto.so.1.1", "libcrypto.so"}) if ((crypto = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!ssl || !crypto) { error = "TLS unavailable: libssl/libcrypto not found on this image"; return; }
    bool all = true;
#define B200_TLS_SYM(lib, name) do { name = reinterpret_cast<decltype(name)>(dlsym(lib, #name)); if (!name) { all = false; error = std::string("TLS unavailable: missing symbol ") + #name; } } while (0)
    B200_TLS_SYM(ssl, OPENSSL_init_ssl); B200_TLS_SYM(ssl, TLS_client_method); B200_TLS_SYM(ssl, SSL_CTX_new); B200_TLS_SYM(ssl, SSL_CTX_free);
    B200_TLS_SYM(ssl, SSL_CTX_load_verify_locations); B200_TLS_SYM(ssl, SSL_CTX_set_default_verify_paths); B200_TLS_SYM(ssl, SSL_CTX_set_verify);
    B200_TLS_SYM(ssl, SSL_new); B200_TLS_SYM(ssl, SSL_free); B200_TLS_SYM(ssl, SSL_set_fd); B200_TLS_SYM(ssl, SSL_connect); B200_TLS_SYM(ssl, SSL_read);
    B200_TLS_SYM(ssl, SSL_write); B200_TLS_SYM(ssl, SSL_shutdown); B200_TLS_SYM(ssl, SSL_ctrl); B200_TLS_SYM(ssl, SSL_set1_host); B200_TLS_SYM(ssl, SSL_get0_param);
    B200_TLS_SYM(ssl, SSL_get_verify_result);
    B200_TLS_SYM(crypto, X509_VERIFY_PARAM_set1_ip_asc); B200_TLS_SYM(crypto, X509_verify_cert_error_string); B200_TLS_SYM(crypto, ERR_get_error);
    B200_TLS_SYM(crypto, ERR_error_string_n);
#undef B200_TLS_SYM
    if (!all) return;
    OPENSSL_init_ssl(0, nullptr);
    ok = true;
  }
  std::string last_error() {
    char b[256] = "unknown TLS error";
    unsigned long e = ERR_get_error();
    if (e) ERR_error_string_n(e, b, sizeof b);
    return b;
  }
};

struct Response {
  int status = 0;            // 0: transport failure, `error` says why
  std::string body, error;
  bool ok() const { return status >= 200 && status < 300; }
  std::string describe() const { return status ? "HTTP " + std::to_string(status) + ": " + body.substr(0, 300) : error; }
};

class Client {
 public:
  std::string host; int port = 443; bool tls = true;
  std::string token, ca_file;
  int timeout_s = 30;
  static constexpr size_t kMaxResponseBytes = 32u << 20;   // a Node object is tens of KiB; anything near this is not one

  // B200_KUBE_URL wins; otherwise the pod's in-cluster environment. Returns "" on success, else why not.
  static std::string from_env(Client* c) {
    const char* sa = "/var/run/secrets/kubernetes.io/serviceaccount";
    std::string token_file = std::string(sa) + "/token", ca = std::string(sa) + "/ca.crt";
    if (const char* t = getenv("B200_KUBE_TOKEN_FILE")) token_file = t;
    if (const char* t = getenv("B200_KUBE_CA_FILE")) ca = t;
    const char* url = getenv("B200_KUBE_URL");
    if (url && *url) {
      std::string err = c->parse_url(url);
      if (!err.empty()) return err;
    } else {
      const char* h = getenv("KUBERNETES_SERVICE_HOST");
      if (!h || !*h) return "not running in a cluster (KUBERNETES_SERVICE_HOST unset)";
      const char* p = getenv("KUBERNETES_SERVICE_PORT");
      c->host = h; c->port = p && *p ? atoi(p) : 443; c->tls = true;
    }
    std::ifstream tf(token_file);
    if (tf) { std::stringstream ss; ss << tf.rdbuf(); c->token = ss.str(); while (!c->token.empty() && isspace((unsigned char)c->token.back())) c->token.pop_back(); }
    else if (c->tls && !(url && *url)) return "cannot read the service-account token " + token_file;
    if (access(ca.c_str(), R_OK) == 0) c->ca_file = ca;
    return "";
  }
  std::string parse_url(const std::string& url) {
    std::string rest;
    if (url.rfind("https://", 0) == 0) { tls = true; rest = url.substr(8); port = 443; }
    else if (url.rfind("http://", 0) == 0) { tls = false; rest = url.substr(7); port = 80; }
    else return "unsupported API server URL " + url;
    size_t slash = rest.find('/');
    if (slash != std::string::npos) rest = rest.substr(0, slash);
    if (!rest.empty() && rest[0] == '[') {                       // [v6]:port
      size_t rb = rest.find(']');
      if (rb == std::string::npos) return "bad IPv6 literal in " + url;
      host = rest.substr(1, rb - 1);
      if (rb + 1 < rest.size() && rest[rb + 1] == ':') port = atoi(rest.c_str() + rb + 2);
    } else {
      size_t colon = rest.rfind(':');
      if (colon != std::string::npos) { host = rest.substr(0, colon); port = atoi(rest.c_str() + colon + 1); } else host = rest;
    }
    return host.empty() ? "no host in " + url : "";
  }

  Response request(const std::string& method, const std::string& path, const std::string& content_type = "", const std::string& body = "") const {
    Response r;
    Conn c;
    std::string err = c.open(*this);
    if (!err.empty()) { r.error = err; return r; }
    std::string host_hdr = host.find(':') != std::string::npos ? "[" + host + "]" : host;
    std::string req = method + " " + path + " HTTP/1.1\r\nHost: " + host_hdr + ":" + std::to_string(port) + "\r\nUser-Agent: b200-device-plugin\r\nAccept: application/json\r\nConnection: close\r\n";
    if (!token.empty()) req += "Authorization: Bearer " + token + "\r\n";
    if (!content_type.empty()) req += "Content-Type: " + content_type + "\r\n";
    if (!body.empty() || method == "POST" || method == "PUT" || method == "PATCH") req += "Content-Length: " + std::to_string(body.size()) + "\r\n";
    req += "\r\n" + body;
    if (!c.write_all(req)) { r.error = "write to API server failed: " + c.error; return r; }
    std::string raw;
    char buf[16384];
    while (true) {
      long n = c.read_some(buf, sizeof buf);
      if (n < 0) { if (raw.empty()) { r.error = "read from API server failed: " + c.error; return r; } break; }
      if (n == 0) break;
      raw.append(buf, (size_t)n);
      if (raw.size() > kMaxResponseBytes) { r.error = "API server response exceeds " + std::to_string(kMaxResponseBytes >> 20) + " MiB"; return r; }
      if (complete(raw)) break;
    }
    parse(raw, &r);
    return r;
  }

  // ---- the handful of typed calls the agent makes
  Response get_node(const std::string& name) const { return request("GET", "/api/v1/nodes/" + name); }
  Response update_node_status(const std::string& name, const json::Value& node) const { return request("PUT", "/api/v1/nodes/" + name + "/status", "application/json", json::dump(node)); }
  Response apply_node_annotations(const std::string& name, const std::map<std::string, std::string>& annotations, const std::string& field_manager, bool force = true) const {
    json::Value patch = json::Value::object();
    patch.at("apiVersion") = json::Value::of("v1");
    patch.at("kind") = json::Value::of("Node");
    patch.at("metadata").at("name") = json::Value::of(name);
    for (auto& kv : annotations) patch.at("metadata").at("annotations").at(kv.first) = json::Value::of(kv.second);
    return request("PATCH", "/api/v1/nodes/" + name + "?fieldManager=" + field_manager + "&force=" + (force ? "true" : "false"), "application/apply-patch+yaml", json::dump(patch));
  }
  Response create_node_event(const std::string& node_name, const std::string& node_uid, const std::string& type, const std::string& reason, const std::string& message, const std::string& component) const {
    const std::string ts = now_rfc3339();
    json::Value ev = json::Value::object();
    ev.at("apiVersion") = json::Value::of("v1");
    ev.at("kind") = json::Value::of("Event");
    ev.at("metadata").at("generateName") = json::Value::of(node_name + ".");
    ev.at("metadata").at("namespace") = json::Value::of("default");
    json::Value& inv = ev.at("involvedObject");
    inv.at("kind") = json::Value::of("Node"); inv.at("name") = json::Value::of(node_name); inv.at("uid") = json::Value::of(node_uid); inv.at("apiVersion") = json::Value::of("v1");
    ev.at("type") = json::Value::of(type); ev.at("reason") = json::Value::of(reason); ev.at("message") = json::Value::of(message);
    ev.at("source").at("component") = json::Value::of(component);
    ev.at("firstTimestamp") = json::Value::of(ts); ev.at("lastTimestamp") = json::Value::of(ts); ev.at("count") = json::Value::of(1L);
    return request("POST", "/api/v1/namespaces/default/events", "application/json", json::dump(ev));
  }

  // ---- wire helpers (public so the native self-test can drive them without a socket)
  static bool response_complete(const std::string& raw) { return complete(raw); }
  static void parse_response(const std::string& raw, Response* r) { parse(raw, r); }

 private:
  struct Conn {
    int fd = -1;
    SSL_CTX* ctx = nullptr;
    SSL* ssl = nullptr;
    std::string error;
    ~Conn() {
      if (ssl) { Tls::get().SSL_shutdown(ssl); Tls::get().SSL_free(ssl); }
      if (ctx) Tls::get().SSL_CTX_free(ctx);
      if (fd >= 0) close(fd);
    }
    std::string open(const Client& c) {
      struct addrinfo hints; memset(&hints, 0, sizeof hints);
      hints.ai_socktype = SOCK_STREAM;
      struct addrinfo* res = nullptr;
      int rc = getaddrinfo(c.host.c_str(), std::to_string(c.port).c_str(), &hints, &res);
      if (rc != 0) return "cannot resolve " + c.host + ": " + gai_strerror(rc);
      std::string last = "no addresses";
      for (struct addrinfo* a = res; a; a = a->ai_next) {
        fd = socket(a->ai_family, a->ai_socktype | SOCK_CLOEXEC, a->ai_protocol);
        if (fd < 0) { last = strerror(errno); continue; }
        struct timeval tv = {c.timeout_s, 0};
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
        int one = 1; setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
        if (connect(fd, a->ai_addr, a->ai_addrlen) == 0) break;
        last = strerror(errno);
        close(fd); fd = -1;
      }
      freeaddrinfo(res);
      if (fd < 0) return "cannot connect to " + c.host + ":" + std::to_string(c.port) + ": " + last;
      if (!c.tls) return "";
      Tls& t = Tls::get();
      if (!t.ok) return t.error;
      ctx = t.SSL_CTX_new(t.TLS_client_method());
      if (!ctx) return "SSL_CTX_new: " + t.last_error();
      if (!c.ca_file.empty()) { if (t.SSL_CTX_load_verify_locations(ctx, c.ca_file.c_str(), nullptr) != 1) return "cannot load CA bundle " + c.ca_file + ": " + t.last_error(); }
      else t.SSL_CTX_set_default_verify_paths(ctx);
      t.SSL_CTX_set_verify(ctx, SSL_VERIFY_PEER, nullptr);
      ssl = t.SSL_new(ctx);
      if (!ssl) return "SSL_new: " + t.last_error();
      unsigned char addr[16];
      const bool is_ip = inet_pton(AF_INET, c.host.c_str(), addr) == 1 || inet_pton(AF_INET6, c.host.c_str(), addr) == 1;
      if (is_ip) t.X509_VERIFY_PARAM_set1_ip_asc(t.SSL_get0_param(ssl), c.host.c_str());
      else {
        t.SSL_set1_host(ssl, c.host.c_str());
        t.SSL_ctrl(ssl, SSL_CTRL_SET_TLSEXT_HOSTNAME, TLSEXT_NAMETYPE_host_name, (void*)c.host.c_str());   // SNI
      }
      t.SSL_set_fd(ssl, fd);
      if (t.SSL_connect(ssl) != 1) {
        long v = t.SSL_get_verify_result(ssl);
        return "TLS handshake with " + c.host + " failed: " + (v != X509_V_OK ? std::string(t.X509_verify_cert_error_string(v)) : t.last_error());
      }
      return "";
    }
    bool write_all(const std::string& s) {
      size_t off = 0;
      while (off < s.size()) {
        long n = ssl ? Tls::get().SSL_write(ssl, s.data() + off, (int)std::min<size_t>(s.size() - off, 1 << 20)) : ::send(fd, s.data() + off, s.size() - off, MSG_NOSIGNAL);
        if (n <= 0) { if (!ssl && errno == EINTR) continue; error = ssl ? Tls::get().last_error() : strerror(errno); return false; }
        off += (size_t)n;
      }
      return true;
    }
    long read_some(char* buf, size_t cap) {
      while (true) {
        long n = ssl ? Tls::get().SSL_read(ssl, buf, (int)cap) : ::recv(fd, buf, cap, 0);
        if (n < 0 && !ssl && errno == EINTR) continue;
        if (n < 0) error = ssl ? Tls::get().last_error() : strerror(errno);
        if (ssl && n <= 0) return 0;        // close_notify or abrupt close: hand back what we have
        return n;
      }
    }
  };

  static std::string lower(std::string s) { for (auto& ch : s) ch = (char)tolower((unsigned char)ch); return s; }
  static bool header(const std::string& head, const std::string& name, std::string* val) {
    std::istringstream in(head);
    std::string line;
    while (std::getline(in, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      size_t c = line.find(':');
      if (c == std::string::npos || lower(line.substr(0, c)) != name) continue;
      size_t b = c + 1;
      while (b < line.size() && line[b] == ' ') b++;
      *val = line.substr(b);
      return true;
    }
    return false;
  }
  // true once `raw` holds a whole response (so we need not wait for the peer to close)
  static bool complete(const std::string& raw) {
    size_t he = raw.find("\r\n\r\n");
    if (he == std::string::npos) return false;
    const std::string head = raw.substr(0, he);
    std::string v;
    if (header(head, "content-length", &v)) return raw.size() >= he + 4 + (size_t)atol(v.c_str());
    if (header(head, "transfer-encoding", &v) && lower(v).find("chunked") != std::string::npos) return raw.find("\r\n0\r\n", he + 2) != std::string::npos && raw.compare(raw.size() - 4, 4, "\r\n\r\n") == 0;
    return false;
  }
  static void parse(const std::string& raw, Response* r) {
    size_t he = raw.find("\r\n\r\n");
    if (he == std::string::npos || raw.compare(0, 5, "HTTP/") != 0) { r->error = "malformed HTTP response from API server"; return; }
    size_t sp = raw.find(' ');
    r->status = sp == std::string::npos ? 0 : atoi(raw.c_str() + sp + 1);
    const std::string head = raw.substr(0, he);
    std::string payload = raw.substr(he + 4), v;
    if (header(head, "transfer-encoding", &v) && lower(v).find("chunked") != std::string::npos) {
      std::string out; size_t i = 0;
      while (i < payload.size()) {
        size_t eol = payload.find("\r\n", i);
        if (eol == std::string::npos) break;
        size_t n = (size_t)strtoul(payload.c_str() + i, nullptr, 16);
        if (n == 0) break;
        i = eol + 2;
        if (i + n > payload.size()) { out.append(payload, i, std::string::npos); break; }
        out.append(payload, i, n);
        i += n + 2;
      }
      r->body = out;
    } else if (header(head, "content-length", &v)) r->body = payload.substr(0, (size_t)atol(v.c_str()));
    else r->body = payload;
    if (r->status == 0) r->error = "malformed HTTP status line from API server";
  }
};

}  // namespace kube
