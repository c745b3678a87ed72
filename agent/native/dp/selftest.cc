// b200-native-selftest — unit checks for the dependency-free pieces the native binaries are built from
// (json.hpp reader/writer, kube.hpp URL + HTTP response parsing, pb.hpp varints, HPACK integer/Huffman coding).
// Exit code 0 = all passed; each failure prints file:line and the expression. Run by tests/test_native_tools.py.
#include <stdio.h>

#include <string>
#include <thread>

#include "h2.hpp"
#include "json.hpp"
#include "kube.hpp"
#include "pb.hpp"

static int g_failed = 0, g_checks = 0;
#define CHECK(expr) do { g_checks++; if (!(expr)) { g_failed++; fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #expr); } } while (0)

static bool parse(const std::string& text, json::Value* v, std::string* err = nullptr) {
  std::string e;
  bool ok = json::Parser(text).parse(v, &e);
  if (err) *err = e;
  return ok;
}

static void test_json() {
  json::Value v;
  CHECK(parse(R"({"a":[1,2.5,-3e2,true,false,null],"s":"x\ny\t\"q\"\\","n":{"k":{}}})", &v));
  CHECK(v.kind == json::Value::Object && v.get("a")->arr.size() == 6);
  CHECK(v.get("a")->arr[1].num == 2.5 && v.get("a")->arr[2].num == -300.0 && v.get("a")->arr[3].b);
  CHECK(v.get_string("s") == "x\ny\t\"q\"\\");
  // numbers survive a round trip as their source text (int64 beyond 2^53, exponent forms)
  CHECK(parse(R"({"big":9223372036854775807,"e":1e-7,"neg":-0})", &v));
  CHECK(json::dump(v) == R"({"big":9223372036854775807,"e":1e-7,"neg":-0})");
  // \u escapes: BMP, surrogate pair, control characters re-escaped on output
  CHECK(parse("\"\\u00e9\\u20ac\\ud83d\\ude00\\u0001\"", &v));
  CHECK(v.str == "\xc3\xa9\xe2\x82\xac\xf0\x9f\x98\x80\x01");
  CHECK(json::dump(v) == "\"\xc3\xa9\xe2\x82\xac\xf0\x9f\x98\x80\\u0001\"");
  // a JSON document embedded as a string inside JSON (the XidCriticalError condition's reason) round-trips
  json::Value reason = json::Value::object();
  reason.at("79") = json::Value::of(true); reason.at("48") = json::Value::of(true);
  json::Value cond = json::Value::object();
  cond.at("reason") = json::Value::of(json::dump(reason));
  CHECK(json::dump(cond) == R"({"reason":"{\"48\":true,\"79\":true}"})");      // keys sorted
  json::Value back, inner;
  CHECK(parse(json::dump(cond), &back) && parse(back.get_string("reason"), &inner) && inner.get("48") && inner.get("79"));
  // malformed input is rejected, never crashes
  for (const char* bad : {"", "{", "[1,", "{\"a\"}", "{\"a\":}", "\"unterminated", "[1 2]", "{\"a\":1,}", "nul", "\"\\u12\"", "1 2"}) { json::Value x; CHECK(!parse(bad, &x)); }
  std::string deep(200, '['); deep += std::string(200, ']');
  std::string err; json::Value x;
  CHECK(!parse(deep, &x, &err) && err == "JSON nested too deeply");
  std::string ok_deep(100, '['); ok_deep += std::string(100, ']');
  CHECK(parse(ok_deep, &x));
  // at(): turns null into an object on demand, find(): no insertion
  json::Value n; n.at("status").at("conditions") = json::Value::array();
  CHECK(n.find("status") && !n.find("spec") && n.obj.size() == 1);
}

static void test_kube() {
  kube::Client c;
  CHECK(c.parse_url("https://10.0.0.1:6443").empty() && c.tls && c.host == "10.0.0.1" && c.port == 6443);
  CHECK(c.parse_url("http://localhost:8080/ignored/path").empty() && !c.tls && c.host == "localhost" && c.port == 8080);
  CHECK(c.parse_url("https://kubernetes.default.svc").empty() && c.port == 443 && c.host == "kubernetes.default.svc");
  CHECK(c.parse_url("https://[fd00::1]:443").empty() && c.host == "fd00::1" && c.port == 443);
  CHECK(!c.parse_url("ftp://x").empty() && !c.parse_url("https://").empty() && !c.parse_url("https://[fd00::1").empty());
  kube::Response r;
  const std::string cl = "HTTP/1.1 200 OK\r\nContent-Type: application/json\r\ncontent-length: 7\r\n\r\n{\"a\":1}EXTRA";
  CHECK(kube::Client::response_complete(cl));
  kube::Client::parse_response(cl, &r);
  CHECK(r.status == 200 && r.ok() && r.body == "{\"a\":1}");
  CHECK(!kube::Client::response_complete("HTTP/1.1 200 OK\r\nContent-Length: 7\r\n\r\n{\"a\""));
  const std::string chunked = "HTTP/1.1 409 Conflict\r\nTransfer-Encoding: chunked\r\n\r\n4\r\n{\"me\r\n9;ext=1\r\nssage\":1}\r\n0\r\n\r\n";
  CHECK(kube::Client::response_complete(chunked));
  r = kube::Response(); kube::Client::parse_response(chunked, &r);
  CHECK(r.status == 409 && !r.ok() && r.body == "{\"message\":1}");
  CHECK(!kube::Client::response_complete("HTTP/1.1 200 OK\r\nTransfer-Encoding: chunked\r\n\r\n4\r\n{\"me"));
  r = kube::Response(); kube::Client::parse_response("HTTP/1.0 204 No Content\r\n\r\n", &r);
  CHECK(r.status == 204 && r.body.empty());
  r = kube::Response(); kube::Client::parse_response("garbage", &r);
  CHECK(r.status == 0 && !r.error.empty());
  CHECK(kube::now_rfc3339().size() == 20 && kube::now_rfc3339()[10] == 'T' && kube::now_rfc3339().back() == 'Z');
}

static std::string hex(const char* h) {
  std::string out;
  for (size_t i = 0; h[i] && h[i + 1]; i += 2) { if (h[i] == ' ') { i--; continue; } out.push_back((char)strtol(std::string(h + i, 2).c_str(), nullptr, 16)); }
  return out;
}

static void test_hpack() {
  // RFC 7541 C.1.2: 1337 on a 5-bit prefix
  std::string enc; h2::hpack_int(&enc, 0x00, 5, 1337);
  CHECK(enc == hex("1f9a0a"));
  enc.clear(); h2::hpack_int(&enc, 0x00, 5, 10); CHECK(enc == hex("0a"));
  // RFC 7541 C.3.1 / C.3.2: literal requests sharing one decoder (dynamic table carries :authority over)
  h2::HpackDecoder d;
  h2::Headers hs;
  std::string req1 = hex("828684410f7777772e6578616d706c652e636f6d");
  CHECK(d.decode(reinterpret_cast<const uint8_t*>(req1.data()), req1.size(), &hs));
  CHECK(hs.size() == 4 && hs[0] == std::make_pair(std::string(":method"), std::string("GET")) && hs[3].first == ":authority" && hs[3].second == "www.example.com");
  hs.clear();
  std::string req2 = hex("828684be58086e6f2d6361636865");
  CHECK(d.decode(reinterpret_cast<const uint8_t*>(req2.data()), req2.size(), &hs));
  CHECK(hs.size() == 5 && hs[3].second == "www.example.com" && hs[4].first == "cache-control" && hs[4].second == "no-cache");
  // RFC 7541 C.4.1 / C.4.2: the same requests Huffman-coded
  h2::HpackDecoder dh; hs.clear();
  std::string h1 = hex("828684418cf1e3c2e5f23a6ba0ab90f4ff");
  CHECK(dh.decode(reinterpret_cast<const uint8_t*>(h1.data()), h1.size(), &hs) && hs.size() == 4 && hs[3].second == "www.example.com");
  hs.clear();
  std::string h2s = hex("828684be5886a8eb10649cbf");
  CHECK(dh.decode(reinterpret_cast<const uint8_t*>(h2s.data()), h2s.size(), &hs) && hs.size() == 5 && hs[4].second == "no-cache");
  // our encoder (literal, never indexed, raw strings) is readable by our decoder, including values > 127 bytes
  h2::Headers mine = {{":status", "200"}, {"content-type", "application/grpc"}, {"grpc-message", std::string(300, 'x')}};
  std::string wire = h2::hpack_encode(mine);
  h2::HpackDecoder d2; hs.clear();
  CHECK(d2.decode(reinterpret_cast<const uint8_t*>(wire.data()), wire.size(), &hs) && hs == mine);
  // hostile input: index 0, unknown dynamic index, string length past the end, length that would wrap, bad Huffman padding, oversized table update
  for (const char* bad : {"80", "ff00", "0005616263", "007fffffffffffffffff7f", "0083ffffff", "3fe21f"}) {
    h2::HpackDecoder x; h2::Headers o; std::string w = hex(bad);
    CHECK(!x.decode(reinterpret_cast<const uint8_t*>(w.data()), w.size(), &o));
  }
  // gRPC length-prefixed framing: split across reads
  std::string buf = h2::grpc_message("abc") + h2::grpc_message("") + h2::grpc_message("defg").substr(0, 6);
  std::vector<std::string> msgs;
  CHECK(h2::grpc_split(&buf, &msgs) && msgs.size() == 2 && msgs[0] == "abc" && msgs[1].empty() && buf.size() == 6);
}

static void test_pb() {
  std::string o; pb::put_varint(&o, 300); CHECK(o == hex("ac02"));
  o.clear(); pb::put_varint(&o, 0xFFFFFFFFFFFFFFFFull); CHECK(o.size() == 10);
  std::string req = pb::encode_register_request("v1beta1", "nvidiaGPU-1.sock", "nvidia.com/gpu");
  std::vector<pb::Field> f;
  CHECK(pb::parse(req, &f) && f.size() == 3 && f[0].number == 1 && f[0].bytes == "v1beta1" && f[1].bytes == "nvidiaGPU-1.sock" && f[2].bytes == "nvidia.com/gpu");
  // AllocateRequest{container_requests:[{devicesIDs:[a,b]},{devicesIDs:[c]}]}
  std::string c1, c2, all;
  pb::put_bytes(&c1, 1, "nvidia0"); pb::put_bytes(&c1, 1, "nvidia1"); pb::put_bytes(&c2, 1, "nvidia2/gi3");
  pb::put_bytes(&all, 1, c1); pb::put_bytes(&all, 1, c2);
  std::vector<std::vector<std::string>> ids;
  CHECK(pb::decode_allocate_request(all, &ids) && ids.size() == 2 && ids[0].size() == 2 && ids[1][0] == "nvidia2/gi3");
  // truncated / overlong input is rejected
  CHECK(!pb::parse(hex("0a05616263"), &f));               // length 5, 3 bytes present
  CHECK(!pb::parse(hex("08ffffffffffffffffffff01"), &f)); // 11-byte varint
  CHECK(!pb::parse(hex("0a"), &f));
}

// `--json-roundtrip`: parse stdin, print the re-serialised document (or "ERR <reason>"): lets the Python tests compare this
// reader/writer with the json module on generated documents.
static int json_roundtrip() {
  std::string in; char buf[65536]; size_t n;
  while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) in.append(buf, n);
  json::Value v; std::string err;
  if (!json::Parser(in).parse(&v, &err)) { printf("ERR %s\n", err.c_str()); return 0; }
  fputs(json::dump(v).c_str(), stdout);
  return 0;
}

// A scripted HTTP/2 peer on a Unix socket: swallows whatever the client sends, answers with `script`, closes.
static std::string serve_script_once(const std::string& path, const std::string& script, std::thread* t) {
  ::unlink(path.c_str());
  int lfd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  sockaddr_un a{}; a.sun_family = AF_UNIX; strncpy(a.sun_path, path.c_str(), sizeof(a.sun_path) - 1);
  if (lfd < 0 || ::bind(lfd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) < 0 || ::listen(lfd, 1) < 0) return "cannot listen on " + path;
  *t = std::thread([lfd, script] {
    int c = ::accept(lfd, nullptr, nullptr);
    if (c >= 0) {
      char buf[4096]; (void)!::read(c, buf, sizeof buf);           // preface + SETTINGS + HEADERS + DATA arrive in one write
      h2::write_all(c, script.data(), script.size());
      ::shutdown(c, SHUT_WR);
      while (::read(c, buf, sizeof buf) > 0) {}
      ::close(c);
    }
    ::close(lfd);
  });
  return "";
}

static void test_h2_client_goaway() {
  using namespace h2;
  const std::string path = "/tmp/b200-selftest-" + std::to_string(getpid()) + ".sock";
  const std::string answer = frame_bytes(HEADERS, END_HEADERS, 1, hpack_encode({{":status", "200"}, {"content-type", "application/grpc"}})) +
                             frame_bytes(DATA, 0, 1, grpc_message("ok")) + frame_bytes(HEADERS, END_HEADERS | END_STREAM, 1, hpack_encode({{"grpc-status", "0"}}));
  struct Case { std::string goaway; int want; const char* what; } cases[] = {
      {frame_bytes(GOAWAY, 0, 0, u32be(0x7FFFFFFF) + u32be(0)) + frame_bytes(GOAWAY, 0, 0, u32be(1) + u32be(0)), 0, "graceful shutdown that still covers stream 1: the answer is read"},
      {frame_bytes(GOAWAY, 0, 0, u32be(0) + u32be(0)), -1, "stream 1 will not be processed"},
      {frame_bytes(GOAWAY, 0, 0, u32be(1) + u32be(2)), -1, "INTERNAL_ERROR"},
  };
  for (const Case& k : cases) {
    std::thread t;
    std::string e = serve_script_once(path, frame_bytes(SETTINGS, 0, 0, "") + k.goaway + answer, &t);
    CHECK(e.empty());
    if (!e.empty()) continue;
    std::string resp, err;
    const int st = unary_call(path, "/v1beta1.Registration/Register", "x", &resp, &err, 3000);
    t.join();
    if (st != k.want) fprintf(stderr, "  case: %s -> status %d (%s)\n", k.what, st, err.c_str());
    CHECK(st == k.want);
    CHECK(k.want != 0 || resp == "ok");
    CHECK(k.want == 0 || err.find("GOAWAY") != std::string::npos);
  }
  ::unlink(path.c_str());
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--json-roundtrip") return json_roundtrip();
  test_h2_client_goaway();
  test_json();
  test_kube();
  test_hpack();
  test_pb();
  printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
