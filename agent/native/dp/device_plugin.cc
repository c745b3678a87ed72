// b200-device-plugin — native kubelet device plugin (static C++ binary, no Python, no gRPC library).
//
// Same contract as container_engine_accelerators_b200/agent (reference: cmd/nvidia_gpu/nvidia_gpu.go:78-186,
// pkg/gpu/nvidia/{manager,beta_plugin,mig/mig,gpusharing/gpusharing,metrics/*}.go):
//   boot: flags + gpu_config.json -> wait for nvidiactl/nvidia-uvm -> NVML -> discovery (+MIG, +MPS probe) -> serve
//   serve: DevicePlugin v1beta1 on <plugin-dir>/nvidiaGPU-<ts>.sock, Register with kubelet.sock, restart on socket removal /
//          kubelet restart / hot-added GPU
//   Allocate: requested specs + default devices + mounts + MPS envs (+ b200coll transport profile)
//   health: NVML Xid events -> Unhealthy via ListAndWatch (48 always critical + XID_CONFIG)
//   metrics: :2112/metrics, per-container (kubelet PodResources) and per-node duty cycle / memory gauges
//   Kubernetes side effects (kube.hpp: HTTP/1.1 + TLS through the system libssl): Event per critical Xid, Node condition
//          XidCriticalError with a 60 s heartbeat and boot-id based auto-clear (health_check/health_checker.go:103-160,
//          288-358), driver-version annotations by server-side apply (version_visibility/version_visibility.go:38-86)
// The Python agent (`python -m container_engine_accelerators_b200.agent.main`) implements the same rules; the same
// conformance suite drives both (tests/test_native_device_plugin.py).
#include <dirent.h>
#include <netinet/in.h>
#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>

#include <algorithm>
#include <array>
#include <deque>
#include <fstream>
#include <functional>
#include <regex>
#include <set>
#include <sstream>

#include "h2.hpp"
#include "json.hpp"
#include "kube.hpp"
#include "pb.hpp"

// ---- the NVML binding (agent/native/b200agent_nvml.cc is compiled into this binary)
extern "C" {
typedef struct { int index; int minor_number; char uuid[96]; char name[96]; char bus_id[32]; unsigned long long mem_total; unsigned long long mem_used; int mig_mode_current; int mig_mode_pending; } b200nvml_device_info;
typedef struct { char uuid[96]; unsigned long long event_type; unsigned long long event_data; unsigned int gpu_instance_id; unsigned int compute_instance_id; } b200nvml_event;
const char* b200nvml_last_error(void);
int b200nvml_init(void);
int b200nvml_device_count(int*);
int b200nvml_device_info_get(int, b200nvml_device_info*);
int b200nvml_average_usage(const char*, unsigned long long, unsigned int*);
int b200nvml_driver_version(char*, unsigned int);
int b200nvml_events_open(void**);
int b200nvml_events_register_xid(void*, int);
int b200nvml_events_wait(void*, unsigned int, b200nvml_event*);
int b200nvml_events_close(void*);
}

namespace {

const char* kResourceName = "nvidia.com/gpu";
const char* kHealthy = "Healthy";
const char* kUnhealthy = "Unhealthy";
const std::regex kNvidiaDeviceRe("^nvidia[0-9]*$");
const std::regex kVgpuDefault("nvidia([0-9]+)/vgpu([0-9]+)$");
const std::regex kVgpuMig("nvidia([0-9]+)/gi([0-9]+)/vgpu([0-9]+)$");
const std::regex kVgpuSuffix("/vgpu([0-9]+)$");
const unsigned kNotMig = 0xFFFFFFFFu;

int g_verbosity = 0;
void logf(char level, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  fprintf(stderr, "%c b200-device-plugin] ", level); vfprintf(stderr, fmt, ap); fputc('\n', stderr);
  va_end(ap);
}
#define LOGI(...) logf('I', __VA_ARGS__)
#define LOGE(...) logf('E', __VA_ARGS__)

bool exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
std::string join(const std::string& a, const std::string& b) { return a.empty() || a.back() == '/' ? a + b : a + "/" + b; }
std::string read_file(const std::string& p, bool* ok) { std::ifstream f(p); std::stringstream ss; ss << f.rdbuf(); *ok = (bool)f; return ss.str(); }
std::vector<std::string> list_dir(const std::string& p, bool* ok, bool files_only) {
  std::vector<std::string> out; *ok = false;
  DIR* d = opendir(p.c_str()); if (!d) return out;
  *ok = true;
  while (dirent* e = readdir(d)) {
    std::string n = e->d_name; if (n == "." || n == "..") continue;
    if (files_only) { struct stat st; if (::stat(join(p, n).c_str(), &st) != 0 || S_ISDIR(st.st_mode)) continue; }
    out.push_back(n);
  }
  closedir(d); std::sort(out.begin(), out.end());
  return out;
}

// ------------------------------------------------------------------------------------------------ config
struct Profile { const char* size; int id; int max_count; const char* families; };
#define MIG_PROFILE(size, id, count, fam) {size, id, count, fam},
const Profile kProfiles[] = {
#include "../mig_profiles.inc"
};
#undef MIG_PROFILE

struct Config {
  std::string partition_size, strategy, transport, lib_dir_host = "/home/kubernetes/bin/nvidia/lib64", lib_dir_container = "/usr/local/nvidia/lib64";
  long max_time_shared = 0, max_shared = 0;
  std::vector<long> xids;
  std::map<std::string, std::string> transport_env;
};

// "" on success; the caller falls back to the empty config on any error (reference nvidia_gpu.go:89-94)
std::string parse_config_text(const std::string& text, Config* cfg) {
  json::Value v; std::string err;
  if (!json::Parser(text).parse(&v, &err) || v.kind != json::Value::Object) return "gpu config must be a JSON object: " + err;
  cfg->partition_size = v.get_string("GPUPartitionSize");
  cfg->max_time_shared = v.get_int("MaxTimeSharedClientsPerGPU");
  if (const json::Value* sh = v.get("GPUSharingConfig")) { cfg->strategy = sh->get_string("GPUSharingStrategy"); cfg->max_shared = sh->get_int("MaxSharedClientsPerGPU"); }
  if (const json::Value* x = v.get("HealthCriticalXid")) for (auto& e : x->arr) if (e.kind == json::Value::Number) cfg->xids.push_back((long)e.num);
  if (const json::Value* t = v.get("Transport")) {
    if (t->kind == json::Value::String) cfg->transport = t->str;
    else { cfg->transport = t->get_string("Name"); cfg->lib_dir_host = t->get_string("LibDirHost", cfg->lib_dir_host); cfg->lib_dir_container = t->get_string("LibDirContainer", cfg->lib_dir_container);
           if (const json::Value* e = t->get("Env")) for (auto& kv : e->obj) if (kv.second.kind == json::Value::String) cfg->transport_env[kv.first] = kv.second.str; }
  }
  return "";
}
std::string add_defaults_and_validate(Config* c) {
  if (c->max_time_shared > 0) {
    if (!c->strategy.empty() || c->max_shared > 0) LOGI("Both MaxTimeSharedClientsPerGPU and GPUSharingConfig are set, use the value of MaxTimeSharedClientsPerGPU");
    c->strategy = "time-sharing"; c->max_shared = c->max_time_shared;
  } else if (c->strategy == "time-sharing" || c->strategy == "mps") {
    if (c->max_shared <= 0) return "MaxSharedClientsPerGPU should be > 0 for time-sharing or mps GPU sharing strategies";
  } else if (c->strategy.empty()) {
    if (c->max_shared > 0) return "GPU sharing strategy needs to be specified when MaxSharedClientsPerGPU > 0";
  } else return "invalid GPU Sharing strategy: " + c->strategy + ", should be one of time-sharing or mps";
  if (!c->transport.empty() && c->transport != "b200coll") return "invalid Transport: " + c->transport + ", should be empty or b200coll";
  return "";
}
Config load_config(const std::string& path) {
  Config empty;
  bool ok; std::string text = read_file(path, &ok);
  if (!ok) { LOGI("No GPU config file (%s); using defaults", path.c_str()); return empty; }
  Config c; std::string err = parse_config_text(text, &c);
  if (err.empty()) err = add_defaults_and_validate(&c);
  if (!err.empty()) { LOGE("failed to parse GPU config file %s: %s; falling back to default GPU config", path.c_str(), err.c_str()); return empty; }
  return c;
}
std::string add_health_critical_xid(Config* c) {
  const char* e = getenv("XID_CONFIG");
  if (!e || !*e) { LOGI("There is no Xid config specified"); return ""; }
  std::vector<long> out; std::stringstream ss(e); std::string tok;
  while (std::getline(ss, tok, ',')) {
    size_t a = tok.find_first_not_of(" \t"), b = tok.find_last_not_of(" \t");
    tok = a == std::string::npos ? "" : tok.substr(a, b - a + 1);
    char* end = nullptr; long v = strtol(tok.c_str(), &end, 10);
    if (tok.empty() || *end) return "Invalid HealthCriticalXid input : " + tok;
    out.push_back(v);
  }
  c->xids = out;
  return "";
}

// ------------------------------------------------------------------------------------------------ sharing
bool is_virtual_device_id(const std::string& id) { return std::regex_search(id, kVgpuDefault) || std::regex_search(id, kVgpuMig); }
std::string validate_request(const std::vector<std::string>& ids, size_t physical_count, const std::string& strategy) {
  if (ids.size() > 1 && is_virtual_device_id(ids[0])) {
    if (strategy == "time-sharing") return "invalid request for sharing GPU (time-sharing), at most 1 nvidia.com/gpu can be requested on GPU nodes";
    if (strategy == "mps" && physical_count > 1) return "invalid request for sharing GPU (MPS), at most 1 nvidia.com/gpu can be requested on multi-GPU nodes";
  }
  return "";
}
bool virtual_to_physical(const std::string& vid, std::string* out) {
  if (!is_virtual_device_id(vid)) return false;
  *out = std::regex_replace(vid, kVgpuSuffix, "");
  return true;
}

// ------------------------------------------------------------------------------------------------ manager
struct Mount { std::string host, container; bool read_only; };

class Manager {
 public:
  std::string dev_dir = "/dev", proc_dir = "/proc", pci_root = "/sys/bus/pci/devices", mps_control_bin = "/usr/local/nvidia/bin/nvidia-cuda-mps-control";
  std::vector<Mount> mounts;
  Config cfg;
  double gpu_check_interval = 10.0, socket_check_interval = 1.0;
  std::vector<std::string> default_devices;
  unsigned long long total_mem_per_gpu = 0;
  std::map<std::string, std::string> uuid_of;     // device id -> GPU uuid (health matching, metrics)
  std::map<std::string, int> index_of;            // "nvidiaN" -> NVML index

  bool numa_topology(const std::string& bus_id, long* node) const {   // false = none / unknown
    std::string bus = bus_id;
    const std::string domain = bus.substr(0, bus.find(':'));
    if (domain.size() == 8 && domain.compare(0, 4, "0000") == 0) bus = bus.substr(4);
    std::transform(bus.begin(), bus.end(), bus.begin(), ::tolower);
    bool ok; std::string text = read_file(join(join(pci_root, bus), "numa_node"), &ok);
    if (!ok) return false;
    char* end = nullptr; long v = strtol(text.c_str(), &end, 10);
    if (end == text.c_str() || v < 0) return false;
    *node = v; return true;
  }

  std::map<std::string, pb::Device> list_physical() {
    std::lock_guard<std::mutex> lk(mu_);
    return cfg.partition_size.empty() ? devices_ : partitions_;
  }
  std::map<std::string, pb::Device> list_devices() {
    auto phys = list_physical();
    if (cfg.max_shared <= 0) return phys;
    std::map<std::string, pb::Device> out;
    for (auto& kv : phys) for (long i = 0; i < cfg.max_shared; i++) { pb::Device d = kv.second; d.id = kv.first + "/vgpu" + std::to_string(i); out[d.id] = d; }   // vGPUs inherit health
    return out;
  }
  // "" on success
  std::string device_spec(std::string id, std::vector<pb::DeviceSpec>* out) {
    if (cfg.max_shared > 0) { std::string phys; if (!virtual_to_physical(id, &phys)) return "virtual device ID (" + id + ") is not valid"; id = phys; }
    std::lock_guard<std::mutex> lk(mu_);
    if (cfg.partition_size.empty()) {
      auto it = devices_.find(id);
      if (it == devices_.end()) return "invalid allocation request with non-existing device " + id;
      if (it->second.health != kHealthy) return "invalid allocation request with unhealthy device " + id;
      const std::string p = join(dev_dir, id);
      out->push_back({p, p, "mrw"});
      return "";
    }
    auto it = partition_specs_.find(id);
    if (it == partition_specs_.end()) return "invalid allocation request with non-existing GPU partition: " + id;
    out->insert(out->end(), it->second.begin(), it->second.end());
    return "";
  }
  void set_device_health(const std::string& name, const std::string& health) {
    std::lock_guard<std::mutex> lk(mu_);
    auto& m = std::regex_match(name, kNvidiaDeviceRe) ? devices_ : partitions_;
    auto it = m.find(name);
    if (it != m.end()) it->second.health = health; else { pb::Device d; d.id = name; d.health = health; m[name] = d; }
  }
  void report_unhealthy(const std::string& id) {   // never blocks the NVML listener
    std::lock_guard<std::mutex> lk(hmu_);
    if (health_q_.size() < 1024) { health_q_.push_back(id); hcv_.notify_all(); } else LOGE("health queue full; dropping update for %s", id.c_str());
  }
  bool pop_health(std::string* id, int timeout_ms) {
    std::unique_lock<std::mutex> lk(hmu_);
    if (!hcv_.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !health_q_.empty(); })) return false;
    *id = health_q_.front(); health_q_.pop_front(); return true;
  }

  std::string discover_gpus() {
    int n = 0;
    if (b200nvml_device_count(&n) != 0) return std::string("failed to get devices count: ") + b200nvml_last_error();
    for (int i = 0; i < n; i++) {
      b200nvml_device_info info;
      if (b200nvml_device_info_get(i, &info) != 0) return "failed to get the device handle for index " + std::to_string(i) + ": " + b200nvml_last_error();
      pb::Device d; d.id = "nvidia" + std::to_string(info.minor_number); d.health = kHealthy;
      long node; if (info.bus_id[0] && numa_topology(info.bus_id, &node)) { d.has_numa = true; d.numa = node; }
      std::lock_guard<std::mutex> lk(mu_);
      devices_[d.id] = d; uuid_of[d.id] = info.uuid; index_of[d.id] = i;
    }
    return "";
  }
  int discover_num_gpus() const {
    bool ok; int n = 0;
    for (auto& f : list_dir(dev_dir, &ok, true)) if (std::regex_match(f, kNvidiaDeviceRe)) n++;
    return ok ? n : -1;
  }
  bool has_additional_gpus() {
    size_t have; { std::lock_guard<std::mutex> lk(mu_); have = devices_.size(); }
    int n = discover_num_gpus();
    if (n > (int)have) { LOGI("Found %d GPUs, while only %zu are registered. Stopping device-plugin server.", n, have); return true; }
    return false;
  }
  bool check_device_paths() const { return exists(join(dev_dir, "nvidiactl")) && exists(join(dev_dir, "nvidia-uvm")); }

  std::string mig_start() {
    const Profile* prof = nullptr;
    for (auto& p : kProfiles) if (cfg.partition_size == p.size) prof = &p;
    if (!prof) return cfg.partition_size + " is not a valid GPU partition size";
    std::map<std::string, std::vector<pb::DeviceSpec>> specs; std::map<std::string, pb::Device> parts;
    const std::string cap_dir = join(proc_dir, "driver/nvidia/capabilities");
    bool ok; auto entries = list_dir(cap_dir, &ok, false);
    if (!ok) return "failed to read capabilities directory (" + cap_dir + ")";
    static const std::regex gpu_re("gpu([0-9]+)"), gi_re("gi([0-9]+)"), minor_re("DeviceFileMinor: ([0-9]+)");
    int partitioned = 0;
    for (auto& e : entries) {
      std::smatch m; if (!std::regex_search(e, m, gpu_re)) continue;
      const std::string gpu_id = m[1]; partitioned++;
      const std::string gi_base = join(join(cap_dir, e), "mig");
      auto gis = list_dir(gi_base, &ok, false);
      if (!ok) return "failed to read GPU instance capabilities dir (" + gi_base + ")";
      int count = 0;
      for (auto& gi : gis) {
        if (!std::regex_search(gi, gi_re)) continue;
        count++;
        auto minor_of = [&](const std::string& path, const char* what, int* out) -> std::string {
          bool rok; std::string text = read_file(path, &rok);
          if (!rok) return std::string("failed to read ") + what + " access file (" + path + ")";
          std::smatch mm; if (!std::regex_search(text, mm, minor_re)) return std::string("unexpected contents in ") + what + " access file(" + path + ")";
          *out = atoi(mm[1].str().c_str()); return "";
        };
        int gi_minor = 0, ci_minor = 0; std::string err;
        if (!(err = minor_of(join(join(gi_base, gi), "access"), "GPU instance", &gi_minor)).empty()) return err;
        if (!(err = minor_of(join(join(join(gi_base, gi), "ci0"), "access"), "compute instance", &ci_minor)).empty()) return err;   // only ci0 is considered
        const std::string gpu_dev = join(dev_dir, "nvidia" + gpu_id), gi_dev = join(join(dev_dir, "nvidia-caps"), "nvidia-cap" + std::to_string(gi_minor)),
                          ci_dev = join(join(dev_dir, "nvidia-caps"), "nvidia-cap" + std::to_string(ci_minor));
        for (auto& p : {gpu_dev, gi_dev, ci_dev}) if (!exists(p)) return "device (" + p + ") not found";
        const std::string id = "nvidia" + gpu_id + "/" + gi;
        LOGI("Discovered GPU partition: %s", id.c_str());
        specs[id] = {{gpu_dev, gpu_dev, "mrw"}, {gi_dev, gi_dev, "mrw"}, {ci_dev, ci_dev, "mrw"}};
        pb::Device d; d.id = id; d.health = kHealthy;
        { std::lock_guard<std::mutex> lk(mu_); auto it = devices_.find("nvidia" + gpu_id); if (it != devices_.end()) { d.has_numa = it->second.has_numa; d.numa = it->second.numa; uuid_of[id] = uuid_of["nvidia" + gpu_id]; } }
        parts[id] = d;
      }
      if (count != prof->max_count) return "Number of partitions (" + std::to_string(count) + ") for GPU " + gpu_id + " does not match expected partition count (" + std::to_string(prof->max_count) + ")";
    }
    const int gpus = discover_num_gpus();
    if (partitioned != gpus) return "Not all GPUs are partitioned as expected. Total number of GPUs: " + std::to_string(gpus) + ", number of partitioned GPUs: " + std::to_string(partitioned);
    std::lock_guard<std::mutex> lk(mu_);
    partition_specs_ = specs; partitions_ = parts;
    return "";
  }

  std::string mps_healthy() const {
    int in[2], out[2];
    if (pipe(in) != 0 || pipe(out) != 0) return "pipe failed";
    pid_t pid = fork();
    if (pid < 0) return "fork failed";
    if (pid == 0) { dup2(in[0], 0); dup2(out[1], 1); close(in[1]); close(out[0]); execl(mps_control_bin.c_str(), mps_control_bin.c_str(), (char*)nullptr); _exit(127); }
    close(in[0]); close(out[1]);
    const char* cmd = "get_default_active_thread_percentage";
    if (write(in[1], cmd, strlen(cmd)) < 0) { /* reported through the exit status */ }
    close(in[1]);
    char buf[256]; std::string text; ssize_t n;
    while ((n = read(out[0], buf, sizeof(buf))) > 0) text.append(buf, (size_t)n);
    close(out[0]);
    int st = 0; waitpid(pid, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) return "failed to health check NVIDIA MPS: exit status " + std::to_string(WIFEXITED(st) ? WEXITSTATUS(st) : -1);
    LOGI("MPS is healthy, active thread percentage = %s", text.c_str());
    return "";
  }

  std::map<std::string, std::string> envs(size_t requested) const {
    std::map<std::string, std::string> e;
    if (cfg.strategy == "mps") {
      e["CUDA_MPS_ACTIVE_THREAD_PERCENTAGE"] = std::to_string((long)requested * 100 / cfg.max_shared);
      const unsigned long long mem = (unsigned long long)requested * total_mem_per_gpu / (unsigned long long)cfg.max_shared;
      e["CUDA_MPS_PINNED_DEVICE_MEM_LIMIT"] = "0=" + std::to_string(mem / (1024ull * 1024ull)) + "M";
    }
    return e;
  }

  std::string start() {
    default_devices = {join(dev_dir, "nvidiactl"), join(dev_dir, "nvidia-uvm")};
    for (const char* extra : {"nvidia-modeset", "nvidia-uvm-tools"}) if (exists(join(dev_dir, extra))) default_devices.push_back(join(dev_dir, extra));
    std::string err = discover_gpus();
    if (!err.empty()) return err;
    if (!cfg.partition_size.empty() && !(err = mig_start()).empty()) return "failed to start mig device manager: " + err;
    if (cfg.strategy == "mps") {
      if (!(err = mps_healthy()).empty()) return "NVIDIA MPS is not running on this node: " + err;
      bool have = false; for (auto& m : mounts) have |= m.host == "/tmp/nvidia-mps";
      if (!have) mounts.push_back({"/tmp/nvidia-mps", "/tmp/nvidia-mps", false});
      b200nvml_device_info info;
      if (b200nvml_device_info_get(0, &info) != 0) return "failed to query total memory available per GPU";
      total_mem_per_gpu = info.mem_total;
    }
    return "";
  }

 private:
  std::mutex mu_, hmu_;
  std::condition_variable hcv_;
  std::deque<std::string> health_q_;
  std::map<std::string, pb::Device> devices_, partitions_;
  std::map<std::string, std::vector<pb::DeviceSpec>> partition_specs_;
};

// ------------------------------------------------------------------------------------------------ transport hook
void apply_transport(const Config& cfg, const std::vector<Mount>& mounts, pb::ContainerAllocateResponse* r) {
  if (cfg.transport != "b200coll") return;
  std::map<std::string, std::string> env = {{"B200COLL_LIB_DIR", cfg.lib_dir_container}, {"B200COLL_LIB", cfg.lib_dir_container + "/libb200coll.so"}, {"LD_LIBRARY_PATH", cfg.lib_dir_container},
                                            {"B200COLL_NVLS", "-1"}, {"B200COLL_ALGO", "auto"}, {"B200COLL_TIMEOUT_MS", "600000"}, {"B200COLL_DEBUG", "WARN"}};
  for (auto& kv : cfg.transport_env) env[kv.first] = kv.second;
  for (auto& kv : env) r->envs.insert(kv);
  bool covered = false;
  for (auto& m : mounts) covered |= cfg.lib_dir_host == m.host || cfg.lib_dir_host.compare(0, m.host.size() + 1, m.host + "/") == 0;
  if (!covered) r->mounts.push_back({cfg.lib_dir_container, cfg.lib_dir_host, true});
}

// ------------------------------------------------------------------------------------------------ preferred allocation (opt-in)
// Same rules as container_engine_accelerators_b200/agent/preferred.py (the tests drive both with one table): must-include first; stay on
// the NUMA nodes already used, else the smallest node that covers what is missing, else the fullest; `spread` takes the physical GPU
// with most free replicas and avoids ones already chosen, `packed` the opposite; ties in natural id order.
std::string g_preferred_policy = "none";

std::vector<long> natural_key(const std::string& id, std::vector<std::string>* words) {
  std::vector<long> nums; std::string cur; bool digits = false;
  auto flush = [&] { if (cur.empty()) return; if (digits) nums.push_back(atol(cur.c_str())); else words->push_back(cur); cur.clear(); };
  for (char ch : id) { const bool d = ch >= '0' && ch <= '9'; if (d != digits) flush(); digits = d; cur.push_back(ch); }
  flush();
  return nums;
}
bool natural_less(const std::string& a, const std::string& b) {
  std::vector<std::string> wa, wb;
  const std::vector<long> na = natural_key(a, &wa), nb = natural_key(b, &wb);
  if (wa != wb) return wa < wb;
  return na != nb ? na < nb : a < b;
}
std::string physical_of(const std::string& id) { return std::regex_replace(id, kVgpuSuffix, ""); }

std::vector<std::string> preferred_allocation(const pb::PreferredRequest& rq, const std::map<std::string, pb::Device>& devices, const std::string& policy) {
  auto numa_of = [&](const std::string& id) -> long { auto it = devices.find(id); return it != devices.end() && it->second.has_numa ? (long)it->second.numa : -1; };
  std::vector<std::string> chosen;
  for (auto& d : rq.must_include) if (std::find(chosen.begin(), chosen.end(), d) == chosen.end()) chosen.push_back(d);
  std::vector<std::string> pool;
  for (auto& d : rq.available) if (std::find(chosen.begin(), chosen.end(), d) == chosen.end() && std::find(pool.begin(), pool.end(), d) == pool.end()) pool.push_back(d);
  std::sort(pool.begin(), pool.end(), natural_less);
  const size_t want = (size_t)std::max<int64_t>(rq.size, 0);
  while (chosen.size() < want && !pool.empty()) {
    const long missing = (long)(want - chosen.size());
    std::set<long> used_nodes; std::set<std::string> chosen_phys;
    for (auto& d : chosen) { used_nodes.insert(numa_of(d)); chosen_phys.insert(physical_of(d)); }
    std::map<long, long> free_node; std::map<std::string, long> free_phys;
    for (auto& d : pool) { free_node[numa_of(d)]++; free_phys[physical_of(d)]++; }
    auto key = [&](const std::string& d) {
      const long node = numa_of(d), free = free_node[node];
      long n0, n1;
      if (used_nodes.count(node)) { n0 = 0; n1 = 0; } else if (free >= missing) { n0 = 1; n1 = free; } else { n0 = 2; n1 = -free; }
      const std::string phys = physical_of(d);
      long s0, s1;
      if (policy == "spread") { s0 = chosen_phys.count(phys) ? 1 : 0; s1 = -free_phys[phys]; } else { s0 = chosen_phys.count(phys) ? 0 : 1; s1 = free_phys[phys]; }
      return std::array<long, 4>{n0, n1, s0, s1};
    };
    size_t best = 0;                                   // pool is in natural order, so the first minimum is the natural-order tie-break
    for (size_t i = 1; i < pool.size(); i++) if (key(pool[i]) < key(pool[best])) best = i;
    chosen.push_back(pool[best]);
    pool.erase(pool.begin() + (long)best);
  }
  if (chosen.size() > std::max(want, rq.must_include.size())) chosen.resize(std::max(want, rq.must_include.size()));
  return chosen;
}

// ------------------------------------------------------------------------------------------------ gRPC service
std::vector<pb::Device> device_vector(Manager* m) { std::vector<pb::Device> v; for (auto& kv : m->list_devices()) v.push_back(kv.second); return v; }

void register_service(h2::Server* srv, Manager* ngm) {
  const std::string svc = "/v1beta1.DevicePlugin/";
  srv->add_unary(svc + "GetDevicePluginOptions", [](const std::string&, std::string* resp) { *resp = pb::encode_options(g_preferred_policy != "none"); return h2::Status{}; });   // empty unless opted in
  srv->add_unary(svc + "PreStartContainer", [](const std::string&, std::string* resp) { LOGE("device-plugin: PreStart should NOT be called for the B200 GPU device plugin"); resp->clear(); return h2::Status{}; });
  srv->add_unary(svc + "GetPreferredAllocation", [ngm](const std::string& req, std::string* resp) {
    resp->clear();
    if (g_preferred_policy == "none") { LOGE("device-plugin: GetPreferredAllocation should NOT be called for the B200 GPU device plugin"); return h2::Status{}; }
    std::vector<pb::PreferredRequest> rqs;
    if (!pb::decode_preferred_request(req, &rqs)) return h2::Status{13, "malformed PreferredAllocationRequest"};
    const auto devices = ngm->list_devices();
    std::vector<std::vector<std::string>> out;
    for (auto& rq : rqs) out.push_back(preferred_allocation(rq, devices, g_preferred_policy));
    *resp = pb::encode_preferred_response(out);
    return h2::Status{};
  });
  srv->add_stream(svc + "ListAndWatch", [ngm](const std::string&, h2::ServerStream* stream) {
    LOGI("device-plugin: ListAndWatch start");
    if (!stream->send(pb::encode_list_and_watch(device_vector(ngm)))) return h2::Status{};
    while (!stream->cancelled()) {
      std::string id;
      if (!ngm->pop_health(&id, 500)) continue;
      LOGI("device-plugin: %s device marked as Unhealthy", id.c_str());
      ngm->set_device_health(id, kUnhealthy);
      if (!stream->send(pb::encode_list_and_watch(device_vector(ngm)))) break;
    }
    return h2::Status{};
  });
  srv->add_unary(svc + "Allocate", [ngm](const std::string& req, std::string* resp) {
    std::vector<std::vector<std::string>> containers;
    if (!pb::decode_allocate_request(req, &containers)) return h2::Status{13, "malformed AllocateRequest"};
    std::vector<pb::ContainerAllocateResponse> out;
    for (auto& ids : containers) {
      std::string err = validate_request(ids, ngm->list_physical().size(), ngm->cfg.strategy);
      if (!err.empty()) return h2::Status{2, err};
      pb::ContainerAllocateResponse r;
      for (auto& id : ids) if (!(err = ngm->device_spec(id, &r.devices)).empty()) return h2::Status{2, err};
      for (auto& d : ngm->default_devices) r.devices.push_back({d, d, "mrw"});
      for (auto& m : ngm->mounts) r.mounts.push_back({m.container, m.host, m.read_only});
      r.envs = ngm->envs(ids.size());
      apply_transport(ngm->cfg, ngm->mounts, &r);
      out.push_back(r);
    }
    *resp = pb::encode_allocate_response(out);
    return h2::Status{};
  });
}

// ------------------------------------------------------------------------------------------------ health
const char* kXidCondition = "XidCriticalError";
const char* kEventSource = "nvidia-gpu-device-plugin";
const long kMonitorXids[] = {48, 63, 64, 79, 119, 120, 123, 140};    // Xids that put the repair condition on the Node

std::string node_name() {
  if (const char* n = getenv("NODE_NAME")) if (*n) return n;
  char b[256] = ""; gethostname(b, sizeof b - 1); return b;
}

// Node-object side of the health checker. Every method is best effort: a failing API server never stops the
// ListAndWatch path from reporting the device Unhealthy.
struct NodeStatus {
  kube::Client api;
  std::string node;

  bool fetch(json::Value* out, const char* why) {
    kube::Response r = api.get_node(node);
    std::string err;
    if (!r.ok()) { LOGE("Failed to get node %s %s: %s", node.c_str(), why, r.describe().c_str()); return false; }
    if (!json::Parser(r.body).parse(out, &err)) { LOGE("Failed to decode node %s: %s", node.c_str(), err.c_str()); return false; }
    return true;
  }
  static std::string boot_id(json::Value& n) {
    json::Value* info = n.at("status").find("nodeInfo");
    return info ? info->get_string("bootID") : "";
  }
  // GET the Node, let `mutate` edit it (0 = nothing to write, 1 = write, -1 = give up), PUT the status. The PUT carries the GET's
  // resourceVersion, so a write that raced with somebody else's (the kubelet's status updates) comes back 409: re-read and try again
  // rather than lose the condition (the reference logs the error and gives up, health_checker.go:338-343). Returns 1 written, 0 no-op, -1 failed.
  int update_status(const char* why, const std::function<int(json::Value&)>& mutate, int attempts = 4) {
    for (int attempt = 0; attempt < attempts; attempt++) {
      json::Value n;
      if (!fetch(&n, why)) return -1;
      const int verdict = mutate(n);
      if (verdict <= 0) return verdict;
      kube::Response r = api.update_node_status(node, n);
      if (r.ok()) return 1;
      if (r.status != 409 || attempt == attempts - 1) { LOGE("Failed to update node %s status %s: %s", node.c_str(), why, r.describe().c_str()); return -1; }
      LOGI("Node %s changed under us (conflict); retrying the status update", node.c_str());
    }
    return -1;
  }
  // After a reboot (bootID differs from the one stored in the condition's message) the repair happened: drop the condition.
  bool reset_condition(bool* removed) {
    *removed = false;
    const int rc = update_status("to reset the XID condition", [&](json::Value& n) {
      const std::string boot = boot_id(n);
      json::Value* conds = n.at("status").find("conditions");
      if (!conds || conds->kind != json::Value::Array) return 0;
      std::vector<json::Value> kept;
      for (auto& c : conds->arr) {
        const std::string last = c.get_string("message");
        if (c.get_string("type") == kXidCondition && c.get_string("status") == "True" && !boot.empty() && !last.empty() && boot != last) continue;
        kept.push_back(c);
      }
      if (kept.size() == conds->arr.size()) return 0;
      conds->arr = kept;
      return 1;
    });
    if (rc < 0) return false;
    if (rc == 1) { *removed = true; LOGI("Successfully removed XIDCriticalError condition from node %s.", node.c_str()); }
    else LOGI("XIDCriticalError condition doesn't exist for node %s.", node.c_str());
    return true;
  }
  void reset_condition_with_backoff(std::atomic<bool>* stop, double timeout_s = 120.0) {
    double backoff = 1.0, spent = 0;
    bool removed;
    while (!*stop) {
      if (reset_condition(&removed)) return;
      if (spent + backoff > timeout_s) { LOGE("Timeout resetting XID condition after %.0f s.", timeout_s); return; }
      LOGE("Failed to reset XID condition, will retry in %.0fs.", backoff);
      for (double t = 0; t < backoff && !*stop; t += 0.1) usleep(100000);
      spent += backoff; backoff = std::min(backoff * 2, 30.0);
    }
  }
  // Adds the Xid to the condition's reason (a JSON object used as a set), creating the condition if needed.
  void monitor_xid(long xid) {
    if (std::find(std::begin(kMonitorXids), std::end(kMonitorXids), xid) == std::end(kMonitorXids)) return;
    const std::string key = std::to_string(xid);
    const int rc = update_status("to add the XIDCriticalError condition", [&](json::Value& n) {
      json::Value& conds = n.at("status").at("conditions");
      if (conds.kind != json::Value::Array) conds = json::Value::array();
      for (auto& c : conds.arr) {
        if (c.get_string("type") != kXidCondition) continue;
        json::Value reason; std::string err, text = c.get_string("reason");
        if (text.empty()) text = "{}";
        if (!json::Parser(text).parse(&reason, &err) || reason.kind != json::Value::Object) { LOGE("Can't decode the value of condition.Reason %s", text.c_str()); return -1; }
        if (reason.get(key)) { LOGI("XIDCriticalError condition already includes this XID %ld, skip", xid); return 0; }
        reason.at(key) = json::Value::of(true);
        c.at("reason") = json::Value::of(json::dump(reason));
        return 1;
      }
      const std::string now = kube::now_rfc3339();
      json::Value reason = json::Value::object(); reason.at(key) = json::Value::of(true);
      json::Value c = json::Value::object();
      c.at("type") = json::Value::of(kXidCondition); c.at("status") = json::Value::of("True");
      c.at("lastHeartbeatTime") = json::Value::of(now); c.at("lastTransitionTime") = json::Value::of(now);
      c.at("reason") = json::Value::of(json::dump(reason)); c.at("message") = json::Value::of(boot_id(n));
      conds.arr.push_back(c);
      return 1;
    });
    if (rc == 1) LOGI("Successfully add XIDCriticalError condition on node %s.", node.c_str());
  }
  void heartbeat() {
    update_status("to update the XIDCondition heartbeat", [&](json::Value& n) {
      json::Value* conds = n.at("status").find("conditions");
      int modified = 0;
      if (conds && conds->kind == json::Value::Array)
        for (auto& c : conds->arr)
          if (c.get_string("type") == kXidCondition && c.get_string("status") == "True") { c.at("lastHeartbeatTime") = json::Value::of(kube::now_rfc3339()); modified = 1; }
      return modified;
    });
  }
  void record_event(long xid) {
    json::Value n;
    if (!fetch(&n, "to record the XID event")) return;
    json::Value* meta = n.find("metadata");
    kube::Response r = api.create_node_event(node, meta ? meta->get_string("uid") : "", "Warning", "XIDError", "Caught XID error, XID=" + std::to_string(xid), kEventSource);
    if (!r.ok()) LOGE("Failed to record XID=%ld for node %s with err %s", xid, node.c_str(), r.describe().c_str());
  }
};

void health_loop(Manager* ngm, NodeStatus* ns, double heartbeat_s, std::atomic<bool>* stop) {
  std::set<long> critical(ngm->cfg.xids.begin(), ngm->cfg.xids.end());
  critical.insert(48);                                   // double-bit ECC is always health-critical
  std::vector<std::thread> helpers;
  if (ns) {
    helpers.emplace_back([ns, stop] { ns->reset_condition_with_backoff(stop); });
    helpers.emplace_back([ns, stop, heartbeat_s] {
      while (!*stop) { ns->heartbeat(); for (double t = 0; t < heartbeat_s && !*stop; t += 0.1) usleep(100000); }
    });
  }
  struct Join { std::vector<std::thread>* t; ~Join() { for (auto& x : *t) x.join(); } } join{&helpers};
  void* set = nullptr;
  if (b200nvml_events_open(&set) != 0) { LOGE("failed to create NVML event set: %s", b200nvml_last_error()); return; }
  for (auto& kv : ngm->index_of) {
    int rc = b200nvml_events_register_xid(set, kv.second);
    if (rc == -3) LOGI("Warning: %s is too old to support healthchecking. It will always be marked healthy.", kv.first.c_str());
    else if (rc != 0) LOGE("failed to register %s for NVML events: %s", kv.first.c_str(), b200nvml_last_error());
  }
  LOGI("Starting GPU Health Checker");
  while (!*stop) {
    b200nvml_event ev;
    int rc = b200nvml_events_wait(set, 1000, &ev);
    if (rc != 0) continue;                               // timeout or transient error
    if (ev.event_type != 8) { LOGI("Skip error Xid=%llu as it is not Xid Critical", ev.event_data); continue; }
    if (ns) { ns->record_event((long)ev.event_data); ns->monitor_xid((long)ev.event_data); }
    if (!critical.count((long)ev.event_data)) { LOGI("Health checker is skipping Xid %llu error", ev.event_data); continue; }
    auto phys = ngm->list_physical();
    if (!ev.uuid[0]) { LOGE("XidCriticalError: Xid=%llu, All devices will go unhealthy.", ev.event_data); for (auto& kv : phys) ngm->report_unhealthy(kv.first); continue; }
    bool found = false;
    for (auto& kv : phys) {
      auto u = ngm->uuid_of.find(kv.first);
      if (u == ngm->uuid_of.end() || u->second != ev.uuid) continue;
      unsigned gi = kNotMig;
      std::smatch m; static const std::regex gi_re("/gi([0-9]+)");
      if (std::regex_search(kv.first, m, gi_re)) gi = (unsigned)atoi(m[1].str().c_str());
      if (gi != ev.gpu_instance_id) continue;
      LOGE("XidCriticalError: Xid=%llu on Device=%s, the device will go unhealthy.", ev.event_data, kv.first.c_str());
      ngm->report_unhealthy(kv.first); found = true;
    }
    if (!found) LOGE("XidCriticalError: Xid=%llu on unknown device.", ev.event_data);
  }
  b200nvml_events_close(set);
}

// ------------------------------------------------------------------------------------------------ driver version annotations
// "570.124.06" -> cloud.google.com/cuda.driver-version.{major,minor,revision,full}; two-part versions leave revision empty.
bool parse_driver_annotations(const std::string& version, std::map<std::string, std::string>* out) {
  std::vector<std::string> parts; std::string cur;
  for (char ch : version) { if (ch == '.') { parts.push_back(cur); cur.clear(); } else cur.push_back(ch); }
  parts.push_back(cur);
  if (parts.size() < 2 || parts.size() > 3) return false;
  for (auto& p : parts) if (p.empty() || !std::all_of(p.begin(), p.end(), [](char ch) { return ch >= '0' && ch <= '9'; })) return false;
  const std::string pre = "cloud.google.com/cuda.driver-version.";
  (*out)[pre + "major"] = parts[0]; (*out)[pre + "minor"] = parts[1]; (*out)[pre + "revision"] = parts.size() == 3 ? parts[2] : ""; (*out)[pre + "full"] = version;
  return true;
}

void publish_driver_version(NodeStatus* ns) {
  char v[96] = "";
  if (b200nvml_driver_version(v, sizeof v) != 0) { LOGE("failed to read the driver version: %s", b200nvml_last_error()); return; }
  std::map<std::string, std::string> ann;
  if (!parse_driver_annotations(v, &ann)) { LOGE("failed to publish driver version annotations: unexpected driver version format: %s", v); return; }
  kube::Response r = ns->api.apply_node_annotations(ns->node, ann, "gpu-device-plugin", true);
  if (!r.ok()) { LOGE("failed to publish driver version annotations: %s", r.describe().c_str()); return; }
  LOGI("published driver version %s on node %s", v, ns->node.c_str());
}

// ------------------------------------------------------------------------------------------------ metrics
std::string esc(const std::string& s) { std::string o; for (char c : s) { if (c == '"' || c == '\\') o.push_back('\\'); o.push_back(c); } return o; }
// libb200coll publishes one 4 KiB counters page per communicator under /dev/shm (coll/src/comm.cu stats_page_publish):
// 64-byte header {"B200COLL", version, pid, rank, nranks, device, nvls} then calls[ops], bytes[ops], algo_calls[7],
// kernel_launches, staged_calls as u64. version 1 has 4 ops, version 2 adds broadcast and reduce.
std::string g_coll_stats_dir = "/dev/shm";
void append_coll_stats(std::ostringstream& os) {
  static const char* kOps[] = {"all_reduce", "all_gather", "reduce_scatter", "alltoall", "broadcast", "reduce"};
  static const char* kAlgos[] = {"auto", "ll", "oneshot", "twoshot", "nvls", "copy", "ll2"};
  struct Page { uint32_t pid, rank; uint64_t calls[6], bytes[6], algo[7], p2p[3], ext[6]; };   // ext: host calls / bytes / zero-copy / pipelined, bulk launches, generic launches
  std::vector<Page> pages;
  if (DIR* d = opendir(g_coll_stats_dir.c_str())) {
    while (dirent* e = readdir(d)) {
      if (strncmp(e->d_name, "b200coll.", 9) != 0) continue;
      std::ifstream f(join(g_coll_stats_dir, e->d_name), std::ios::binary);
      char raw[64 + 30 * 8] = {};
      f.read(raw, sizeof raw);                                                   // pages are 4 KiB; anything shorter than the v1 payload is not one
      if (f.gcount() < 64 + 17 * 8 || memcmp(raw, "B200COLL", 8) != 0) continue;
      uint32_t hdr[6]; memcpy(hdr, raw + 8, sizeof hdr);
      const int nops = hdr[0] == 1 ? 4 : 6;
      uint64_t updated = 0; if (hdr[0] >= 2) memcpy(&updated, raw + 32, sizeof updated);
      if (updated && (uint64_t)time(nullptr) > updated + 3600) continue;      // left behind by a process that died without CommDestroy
      uint64_t v[21]; memcpy(v, raw + 64, sizeof v);
      Page p{}; p.pid = hdr[1]; p.rank = hdr[2];
      for (int i = 0; i < nops; i++) { p.calls[i] = v[i]; p.bytes[i] = v[nops + i]; }
      for (int i = 0; i < 7; i++) p.algo[i] = v[2 * nops + i];
      if (hdr[0] >= 2) memcpy(p.p2p, raw + 64 + 21 * 8, sizeof p.p2p);           // sends, recvs, bytes; zero on pages of a library without send / recv
      if (hdr[0] >= 2 && f.gcount() >= 64 + 30 * 8) memcpy(p.ext, raw + 64 + 24 * 8, sizeof p.ext);   // appended in round 2; zero on older pages
      pages.push_back(p);
    }
    closedir(d);
  }
  if (pages.empty()) return;
  auto help = [&](const char* n, const char* h) { os << "# HELP " << n << " " << h << "\n# TYPE " << n << " gauge\n"; };
  help("b200coll_calls", "Collective calls issued through libb200coll");
  for (auto& p : pages) for (int i = 0; i < 6; i++) os << "b200coll_calls{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",op=\"" << kOps[i] << "\"} " << p.calls[i] << "\n";
  help("b200coll_bytes", "Bytes moved by libb200coll collectives");
  for (auto& p : pages) for (int i = 0; i < 6; i++) os << "b200coll_bytes{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",op=\"" << kOps[i] << "\"} " << p.bytes[i] << "\n";
  help("b200coll_algo_calls", "libb200coll calls per chosen algorithm");
  for (auto& p : pages) for (int i = 0; i < 7; i++) os << "b200coll_algo_calls{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",algo=\"" << kAlgos[i] << "\"} " << p.algo[i] << "\n";
  help("b200coll_p2p_calls", "Point-to-point operations issued through libb200coll");
  for (auto& p : pages) for (int i = 0; i < 2; i++) os << "b200coll_p2p_calls{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",dir=\"" << (i ? "recv" : "send") << "\"} " << p.p2p[i] << "\n";
  help("b200coll_p2p_bytes", "Bytes sent plus received by libb200coll point-to-point operations");
  for (auto& p : pages) os << "b200coll_p2p_bytes{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\"} " << p.p2p[2] << "\n";
  help("b200coll_host_calls", "End-to-end host all-reduces (b200collAllReduceHost) by path");
  static const char* kPaths[] = {"total", "zero_copy", "pipelined"};
  for (auto& p : pages) for (int i = 0; i < 3; i++) os << "b200coll_host_calls{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",path=\"" << kPaths[i] << "\"} " << p.ext[i == 0 ? 0 : i + 1] << "\n";
  help("b200coll_host_bytes", "Input bytes of the end-to-end host all-reduces");
  for (auto& p : pages) os << "b200coll_host_bytes{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\"} " << p.ext[1] << "\n";
  help("b200coll_kernel_family_launches", "Launches of the copy-engine (bulk) and generic-reduction kernels");
  for (auto& p : pages) { os << "b200coll_kernel_family_launches{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",family=\"bulk\"} " << p.ext[4] << "\n";
                          os << "b200coll_kernel_family_launches{pid=\"" << p.pid << "\",rank=\"" << p.rank << "\",family=\"generic\"} " << p.ext[5] << "\n"; }
}

std::string collect_metrics(Manager* ngm, const std::string& pod_resources_socket) {
  std::ostringstream os;
  struct Info { std::string uuid, name; unsigned long long total, used; unsigned duty; bool ok; };
  std::map<std::string, Info> gpus;
  const unsigned long long since = (unsigned long long)((time(nullptr) - 10)) * 1000000ull;
  for (auto& kv : ngm->index_of) {
    b200nvml_device_info di; Info inf{};
    if (b200nvml_device_info_get(kv.second, &di) != 0) continue;
    inf.uuid = di.uuid; inf.name = di.name; inf.total = di.mem_total; inf.used = di.mem_used;
    inf.ok = b200nvml_average_usage(di.uuid, since, &inf.duty) == 0 && inf.duty <= 100;       // > 100 => skipped this tick
    gpus[kv.first] = inf;
  }
  auto help = [&](const char* n, const char* h) { os << "# HELP " << n << " " << h << "\n# TYPE " << n << " gauge\n"; };
  std::vector<pb::ContainerDevices> cds;
  std::string resp, err;
  // v1 first (kubelet >= 1.20; same field numbers for what is read here), then the reference's v1alpha1 (metrics/devices.go:33-34)
  int st = h2::unary_call(pod_resources_socket, "/v1.PodResourcesLister/List", "", &resp, &err, 3000);
  if (st == 12) st = h2::unary_call(pod_resources_socket, "/v1alpha1.PodResourcesLister/List", "", &resp, &err, 3000);      // 12 = UNIMPLEMENTED
  if (st == 0) pb::decode_pod_resources(resp, &cds);
  else if (g_verbosity > 0) LOGE("Failed to get devices for containers: %s", err.c_str());
  std::map<std::string, std::vector<std::string>> per_ctr;   // label prefix -> physical ids
  std::map<std::string, size_t> requests;
  for (auto& cd : cds) {
    if (cd.resource != kResourceName || cd.ids.empty()) continue;
    const std::string lab = "namespace=\"" + esc(cd.ns) + "\",pod=\"" + esc(cd.pod) + "\",container=\"" + esc(cd.container) + "\"";
    for (auto& id : cd.ids) if (!is_virtual_device_id(id)) per_ctr[lab].push_back(id);
    requests[lab] += per_ctr[lab].size();
    per_ctr[lab];
  }
  help("request", "Number of accelerator devices requested by the container");
  for (auto& kv : per_ctr) os << "request{" << kv.first << ",resource_name=\"nvidia.com/gpu\"} " << kv.second.size() << "\n";
  const char* names[3] = {"duty_cycle", "memory_total", "memory_used"};
  const char* helps[3] = {"Percent of time when the GPU was actively processing", "Total memory available on the GPU in bytes", "Allocated GPU memory in bytes"};
  for (int k = 0; k < 3; k++) {
    help(names[k], helps[k]);
    for (auto& kv : per_ctr) for (auto& id : kv.second) {
      auto g = gpus.find(id); if (g == gpus.end() || !g->second.ok) continue;
      const unsigned long long v = k == 0 ? g->second.duty : k == 1 ? g->second.total : g->second.used;
      os << names[k] << "{" << kv.first << ",make=\"nvidia\",accelerator_id=\"" << esc(g->second.uuid) << "\",model=\"" << esc(g->second.name) << "\"} " << v << "\n";
    }
  }
  for (int k = 0; k < 3; k++) {
    const std::string n = std::string(names[k]) + "_gpu_node";
    help(n.c_str(), helps[k]);
    for (auto& g : gpus) { if (!g.second.ok) continue; const unsigned long long v = k == 0 ? g.second.duty : k == 1 ? g.second.total : g.second.used;
      os << n << "{make=\"nvidia\",accelerator_id=\"" << esc(g.second.uuid) << "\",model=\"" << esc(g.second.name) << "\"} " << v << "\n"; }
  }
  append_coll_stats(os);
  return os.str();
}

void metrics_server(Manager* ngm, int port, int interval_ms, const std::string& pod_resources_socket, std::atomic<bool>* stop) {
  int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
  int one = 1; setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in a{}; a.sin_family = AF_INET; a.sin_addr.s_addr = htonl(INADDR_ANY); a.sin_port = htons((uint16_t)port);
  if (::bind(fd, reinterpret_cast<sockaddr*>(&a), sizeof(a)) < 0 || ::listen(fd, 16) < 0) { LOGE("Failed to start metric server on port %d: %s", port, strerror(errno)); ::close(fd); return; }
  std::mutex mu; std::string snapshot = collect_metrics(ngm, pod_resources_socket);
  std::thread collector([&] { while (!*stop) { for (int i = 0; i < interval_ms / 100 && !*stop; i++) usleep(100000); std::string s = collect_metrics(ngm, pod_resources_socket); std::lock_guard<std::mutex> lk(mu); snapshot = s; } });
  while (!*stop) {
    pollfd p{fd, POLLIN, 0};
    if (poll(&p, 1, 500) <= 0) continue;
    int c = ::accept4(fd, nullptr, nullptr, SOCK_CLOEXEC);
    if (c < 0) continue;
    struct timeval tv = {2, 0};                           // one silent client must not stall the (single-threaded) endpoint
    setsockopt(c, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv); setsockopt(c, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
    char req[2048]; ssize_t n = ::recv(c, req, sizeof(req) - 1, 0); req[n > 0 ? n : 0] = 0;
    std::string body, status = "200 OK";
    if (strncmp(req, "GET /metrics", 12) == 0) { std::lock_guard<std::mutex> lk(mu); body = snapshot; } else { status = "404 Not Found"; body = "not found\n"; }
    const std::string resp = "HTTP/1.1 " + status + "\r\nContent-Type: text/plain; version=0.0.4\r\nContent-Length: " + std::to_string(body.size()) + "\r\nConnection: close\r\n\r\n" + body;
    h2::write_all(c, resp.data(), resp.size());
    ::close(c);
  }
  collector.join();
  ::close(fd);
}

// ------------------------------------------------------------------------------------------------ serve loop
std::atomic<bool> g_stop{false};
void on_signal(int) { g_stop = true; }

// Identity of a socket file: inode number mixed with its change time. The inode number alone is not enough — a kubelet that removes
// and re-creates kubelet.sock between two polls usually gets the very same number back from the filesystem. 0 = no such file.
uint64_t inode_of(const std::string& p) {
  struct stat st;
  if (::stat(p.c_str(), &st) != 0) return 0;
  const uint64_t id = (uint64_t)st.st_ino ^ (((uint64_t)st.st_ctim.tv_sec * 1000000000ull + (uint64_t)st.st_ctim.tv_nsec) * 0x9E3779B97F4A7C15ull);
  return id ? id : 1;
}

int serve(Manager* ngm, const std::string& plugin_dir, const std::string& kubelet_endpoint, const std::string& plugin_endpoint) {
  const std::string kubelet_path = join(plugin_dir, kubelet_endpoint);
  bool do_register = exists(kubelet_path);
  LOGI(do_register ? "will register with the kubelet (beta API)" : "no kubelet.sock to register.");
  while (!g_stop) {
    const std::string sock = join(plugin_dir, plugin_endpoint);
    h2::Server server;
    register_service(&server, ngm);
    std::string err;
    if (!server.listen_unix(sock, &err)) { LOGE("cannot listen on %s: %s", sock.c_str(), err.c_str()); return 1; }
    LOGI("device-plugin: serving on %s", sock.c_str());
    // identity of the kubelet socket BEFORE registering: a kubelet that restarts right after our Register call then shows up as a
    // different socket in the watch loop below (taken afterwards, the new kubelet would silently become the baseline)
    const uint64_t kubelet_ino = inode_of(kubelet_path);
    if (do_register) {
      std::string resp;
      int st = h2::unary_call(kubelet_path, "/v1beta1.Registration/Register", pb::encode_register_request("v1beta1", plugin_endpoint, kResourceName, g_preferred_policy != "none"), &resp, &err);
      if (st != 0) { server.stop(); LOGE("device-plugin: cannot register to kubelet service: %s", err.c_str()); return 1; }   // pod restarts (reference: glog.Fatal)
      LOGI("device-plugin registered with the kubelet");
    }
    bool kubelet_gone = false;
    auto next_gpu_check = std::chrono::steady_clock::now() + std::chrono::milliseconds((int)(ngm->gpu_check_interval * 1000));
    bool rediscover = false;
    while (!g_stop) {
      usleep((useconds_t)(ngm->socket_check_interval * 1e6));
      if (!exists(sock)) { LOGI("plugin socket %s was removed; restarting the server", sock.c_str()); break; }
      const uint64_t ino = inode_of(kubelet_path);
      if (do_register && !ino) kubelet_gone = true;          // seen while the socket was absent: whatever appears next is a new kubelet
      if (do_register && ino && (ino != kubelet_ino || kubelet_gone)) { LOGI("kubelet socket was re-created (kubelet restart); re-registering"); break; }
      if (!do_register && ino) { do_register = true; LOGI("kubelet socket appeared; registering"); break; }
      if (std::chrono::steady_clock::now() >= next_gpu_check) {
        next_gpu_check = std::chrono::steady_clock::now() + std::chrono::milliseconds((int)(ngm->gpu_check_interval * 1000));
        if (ngm->has_additional_gpus()) { rediscover = true; break; }
      }
    }
    server.stop();
    ::unlink(sock.c_str());
    if (rediscover) {
      double backoff = 1.0;
      while (!g_stop) { std::string e = ngm->discover_gpus(); if (e.empty()) break; LOGE("rediscovery failed: %s; retrying in %.0fs", e.c_str(), backoff); usleep((useconds_t)(backoff * 1e6)); backoff = std::min(backoff * 2, 30.0); }
    }
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  std::string host_path = "/home/kubernetes/bin/nvidia", container_path = "/usr/local/nvidia", host_vulkan = "/home/kubernetes/bin/nvidia/vulkan/icd.d", container_vulkan = "/etc/vulkan/icd.d",
              plugin_dir = "/device-plugin", gpu_config = "/etc/nvidia/gpu_config.json", plugin_endpoint, pod_resources = "/var/lib/kubelet/pod-resources/kubelet.sock";
  bool enable_metrics = false, enable_health = false, publish_version = false;
  double xid_heartbeat_s = 60.0;
  int metrics_port = 2112, metrics_interval = 30000;
  Manager ngm;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    std::string val; bool has = false;
    size_t eq = a.find('=');
    if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has = true; }
    auto need = [&]() -> std::string { if (has) return val; if (i + 1 < argc) return argv[++i]; fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); };
    if (a == "host-path") host_path = need();
    else if (a == "container-path") container_path = need();
    else if (a == "host-vulkan-icd-path") host_vulkan = need();
    else if (a == "container-vulkan-icd-path") container_vulkan = need();
    else if (a == "plugin-directory") plugin_dir = need();
    else if (a == "gpu-config") gpu_config = need();
    else if (a == "gpu-metrics-port") metrics_port = atoi(need().c_str());
    else if (a == "gpu-metrics-collection-interval") metrics_interval = atoi(need().c_str());
    else if (a == "enable-container-gpu-metrics") enable_metrics = !has || val != "false";
    else if (a == "enable-health-monitoring") enable_health = !has || val != "false";
    else if (a == "publish-driver-version") publish_version = !has || val != "false";
    else if (a == "xid-heartbeat-interval") xid_heartbeat_s = atof(need().c_str());
    else if (a == "dev-directory") ngm.dev_dir = need();
    else if (a == "proc-directory") ngm.proc_dir = need();
    else if (a == "pci-root") ngm.pci_root = need();
    else if (a == "mps-control-bin") ngm.mps_control_bin = need();
    else if (a == "plugin-endpoint") plugin_endpoint = need();
    else if (a == "pod-resources-socket") pod_resources = need();
    else if (a == "coll-stats-dir") g_coll_stats_dir = need();
    else if (a == "preferred-allocation-policy") { g_preferred_policy = need(); if (g_preferred_policy != "none" && g_preferred_policy != "spread" && g_preferred_policy != "packed") { fprintf(stderr, "bad -preferred-allocation-policy %s (none|spread|packed)\n", g_preferred_policy.c_str()); return 2; } }
    else if (a == "gpu-check-interval") ngm.gpu_check_interval = atof(need().c_str());
    else if (a == "socket-check-interval") ngm.socket_check_interval = atof(need().c_str());
    else if (a == "v") g_verbosity = atoi(need().c_str());
    else if (a == "logtostderr" || a == "alsologtostderr") {}
    else if (a == "h" || a == "help") {
      puts("b200-device-plugin: kubelet device plugin for nvidia.com/gpu (whole GPUs, MIG slices, time-shared / MPS replicas).\n"
           "  -plugin-directory DIR (/device-plugin)   -gpu-config PATH (/etc/nvidia/gpu_config.json)\n"
           "  -host-path DIR -container-path DIR       driver/library mount handed to containers (/home/kubernetes/bin/nvidia -> /usr/local/nvidia)\n"
           "  -host-vulkan-icd-path DIR -container-vulkan-icd-path DIR\n"
           "  -enable-health-monitoring                NVML Xid events -> Unhealthy, Node condition XidCriticalError, Events (env XID_CONFIG=\"31,79\", NODE_NAME)\n"
           "  -enable-container-gpu-metrics            Prometheus gauges on -gpu-metrics-port (2112), every -gpu-metrics-collection-interval ms (30000)\n"
           "  -publish-driver-version                  cloud.google.com/cuda.driver-version.* Node annotations\n"
           "  -preferred-allocation-policy none|spread|packed   answer GetPreferredAllocation (default none: no plugin options, like the reference)\n"
           "  -xid-heartbeat-interval S (60)  -pod-resources-socket PATH  -coll-stats-dir DIR (/dev/shm)  -v N\n"
           "test seams: -dev-directory -proc-directory -pci-root -mps-control-bin -plugin-endpoint -gpu-check-interval -socket-check-interval;\n"
           "Kubernetes API: in-cluster config, or B200_KUBE_URL / B200_KUBE_TOKEN_FILE / B200_KUBE_CA_FILE");
      return 0;
    }
    else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
  }
  signal(SIGINT, on_signal); signal(SIGTERM, on_signal); signal(SIGPIPE, SIG_IGN);
  LOGI("device-plugin started");
  ngm.mounts = {{host_path, container_path, true}, {host_vulkan, container_vulkan, true}};
  ngm.cfg = load_config(gpu_config);
  std::string err = add_health_critical_xid(&ngm.cfg);
  if (!err.empty()) LOGE("failed to add HealthCriticalXid: %s", err.c_str());
  while (!g_stop && !ngm.check_device_paths()) { if (g_verbosity >= 3) LOGI("nvidiactl / nvidia-uvm not present yet; waiting for the driver installer"); sleep(5); }
  LOGI("Initializing nvml");
  if (b200nvml_init() != 0) { LOGE("nvml init failed: %s", b200nvml_last_error()); return 1; }
  while (!g_stop) { err = ngm.start(); if (err.empty()) break; LOGE("failed to start GPU device manager: %s", err.c_str()); sleep(5); }
  if (g_stop) return 0;
  std::vector<std::thread> side;
  if (enable_metrics) {
    if (!ngm.cfg.partition_size.empty()) LOGI("metrics are disabled when MIG partitioning is on");
    else { LOGI("Starting metrics server on port: %d, collection interval: %d", metrics_port, metrics_interval); side.emplace_back(metrics_server, &ngm, metrics_port, metrics_interval, pod_resources, &g_stop); }
  }
  NodeStatus node_status;
  bool have_kube = false;
  if (enable_health || publish_version) {
    node_status.node = node_name();
    std::string why = kube::Client::from_env(&node_status.api);
    have_kube = why.empty();
    if (!have_kube) LOGE("failed to build kube client: %s; Xid Events, the Node condition and driver-version annotations are disabled", why.c_str());
  }
  if (enable_health) side.emplace_back(health_loop, &ngm, have_kube ? &node_status : nullptr, xid_heartbeat_s, &g_stop);
  if (publish_version && have_kube) side.emplace_back(publish_driver_version, &node_status);
  if (plugin_endpoint.empty()) plugin_endpoint = "nvidiaGPU-" + std::to_string((long)time(nullptr)) + ".sock";
  int rc = serve(&ngm, plugin_dir, "kubelet.sock", plugin_endpoint);
  g_stop = true;
  for (auto& t : side) t.join();
  return rc;
}
