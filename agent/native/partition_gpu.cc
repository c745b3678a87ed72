// b200-partition-gpu — one-shot MIG reconciler (static C++ binary for a distroless initContainer).
//
// Contract: reference partition_gpu/partition_gpu.go:157-467. Command sequence, in order:
//   nvidia-smi --query-gpu=mig.mode.current --format=csv,noheader   (first line decides, HasPrefix Enabled/Disabled)
//   nvidia-smi -mig 1                      (if disabled)        + --query-gpu=gpu_name --format=csv,noheader
//   nvidia-smi mig -lgi                    compare with the desired uniform layout
//   nvidia-smi mig -dci ; mig -dgi         (tolerating "No GPU/compute instances found")
//   nvidia-smi mig -cgi id,id,...(xmax) ; mig -cci ; bare nvidia-smi
// Exit codes: 0 nothing to do / success, 1 any failure or reboot requested. A100 needs a reboot after -mig 1
// (kill(1, SIGRTMIN+5)); Hopper and Blackwell continue in-process.
// Differences from the reference: one size table shared with the device plugin (mig_profiles.inc); the
// "already in desired state" check also compares the profile id, not only the per-GPU count (the reference
// accepts e.g. 4x profile 14 when 4x profile 15 was asked for, partition_gpu.go:446-458).
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <fstream>
#include <map>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Profile { const char* size; int id; int max_count; const char* families; };
#define MIG_PROFILE(size, id, count, fam) {size, id, count, fam},
const Profile kProfiles[] = {
#include "mig_profiles.inc"
};
#undef MIG_PROFILE

std::string g_smi = "/usr/local/nvidia/bin/nvidia-smi";
std::string g_config = "/etc/nvidia/gpu_config.json";

void logi(const std::string& s) { fprintf(stderr, "I partition_gpu] %s\n", s.c_str()); }
void loge(const std::string& s) { fprintf(stderr, "E partition_gpu] %s\n", s.c_str()); }

const Profile* find_profile(const std::string& size) {
  for (const Profile& p : kProfiles) if (size == p.size) return &p;
  return nullptr;
}

struct Result { int rc; std::string out; };

Result run(const std::vector<std::string>& args) {
  int fds[2];
  if (pipe(fds) != 0) return {-1, ""};
  pid_t pid = fork();
  if (pid < 0) return {-1, ""};
  if (pid == 0) {
    dup2(fds[1], 1);
    close(fds[0]); close(fds[1]);
    std::vector<char*> argv;
    argv.push_back(const_cast<char*>(g_smi.c_str()));
    for (const std::string& a : args) argv.push_back(const_cast<char*>(a.c_str()));
    argv.push_back(nullptr);
    execv(g_smi.c_str(), argv.data());
    _exit(127);
  }
  close(fds[1]);
  std::string out;
  char buf[4096];
  ssize_t n;
  while ((n = read(fds[0], buf, sizeof(buf))) > 0) out.append(buf, (size_t)n);
  close(fds[0]);
  int st = 0;
  waitpid(pid, &st, 0);
  return {WIFEXITED(st) ? WEXITSTATUS(st) : -1, out};
}

std::string join(const std::vector<std::string>& v) { std::string s; for (auto& x : v) s += " " + x; return s; }
bool has_prefix(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }

// 1 enabled, 0 disabled, -1 error
int current_mig_mode() {
  Result r = run({"--query-gpu=mig.mode.current", "--format=csv,noheader"});
  if (r.rc != 0) { loge("nvidia-smi exited with " + std::to_string(r.rc)); return -1; }
  if (has_prefix(r.out, "Enabled")) return 1;
  if (has_prefix(r.out, "Disabled")) return 0;
  loge("nvidia-smi returned invalid output: " + r.out);
  return -1;
}

std::string check_gpu_type() {
  Result r = run({"--query-gpu=gpu_name", "--format=csv,noheader"});
  if (r.rc != 0) return "";
  // order matters: "NVIDIA GB200" before "NVIDIA B200" is irrelevant for prefixes, but keep the most specific first
  static const char* kTypes[] = {"NVIDIA GB200", "NVIDIA B200", "NVIDIA H200", "NVIDIA H100 80GB HBM3", "NVIDIA A100-SXM4-40GB", "NVIDIA A100-SXM4-80GB", "NVIDIA RTX PRO 6000"};
  for (const char* t : kTypes) if (has_prefix(r.out, t)) return t;
  loge("nvidia-smi returned invalid GPU type for MIG: " + r.out);
  return "";
}

int reboot_node() {
  if (const char* hook = getenv("B200_PARTITION_REBOOT_HOOK")) {   // tests: record instead of signalling pid 1
    std::ofstream(hook) << "reboot\n";
    return 0;
  }
  return kill(1, SIGRTMIN + 5);   // systemd: graceful reboot
}

// gpu index -> profile ids, in table order. uniform=false as soon as two rows disagree or no row parsed.
bool parse_lgi(const std::string& text, std::map<std::string, std::vector<std::string>>* by_gpu, bool* uniform) {
  static const std::regex row(R"(^\s*(\d+)\s+(MIG\s+[\w\.\+\-]+)\s+(\d+)\s+(\d+)\s+([\d:]+)\s*$)");
  std::string first;
  std::istringstream in(text);
  std::string line;
  *uniform = false;
  while (std::getline(in, line)) {
    size_t a = line.find_first_not_of(" \t\r"), b = line.find_last_not_of(" \t\r");
    if (a == std::string::npos) continue;
    line = line.substr(a, b - a + 1);
    if (line.find("====") != std::string::npos) continue;
    if (line.size() < 2 || line.front() != '|' || line.back() != '|') continue;
    std::string inner = line.substr(1, line.size() - 2);
    std::smatch m;
    if (std::regex_match(inner, m, row)) {
      const std::string gpu = m[1], pid = m[3];
      if (first.empty()) first = pid;
      else if (first != pid) return true;        // non-uniform: rebuild
      (*by_gpu)[gpu].push_back(pid);
    }
  }
  *uniform = !by_gpu->empty();
  return true;
}

bool tolerated(const std::string& out);

bool matches_desired(const Profile& want) {
  Result r = run({"mig", "-lgi"});
  if (r.rc != 0 && tolerated(r.out)) { logi("No GPU instances exist yet (nvidia-smi mig -lgi: " + r.out.substr(0, r.out.find('\n')) + ")"); return false; }   // a real B200 in fresh MIG mode answers with a non-zero status
  if (r.rc != 0) { loge("failed to execute 'nvidia-smi mig -lgi'"); return false; }
  logi("Output:\n " + r.out);
  std::map<std::string, std::vector<std::string>> by_gpu;
  bool uniform = false;
  parse_lgi(r.out, &by_gpu, &uniform);
  if (!uniform) { logi("Partitions are not uniform (or absent), partition reconstruction needed."); return false; }
  for (auto& kv : by_gpu) {
    if ((int)kv.second.size() != want.max_count) return false;
    if (kv.second.front() != std::to_string(want.id)) return false;
  }
  return true;
}

bool tolerated(const std::string& out) {
  return out.find("No GPU instances found") != std::string::npos || out.find("No compute instances found") != std::string::npos;
}

bool cleanup_all() {
  for (const char* flag : {"-dci", "-dgi"}) {
    logi(std::string("Running ") + g_smi + " mig " + flag);
    Result r = run({"mig", flag});
    if (r.rc != 0 && !tolerated(r.out)) { loge(std::string("failed to destroy instances (mig ") + flag + "), nvidia-smi output: " + r.out); return false; }
    logi("Output:\n " + r.out);
  }
  return true;
}

std::string build_partition_str(const Profile& p) {
  std::string s;
  for (int i = 0; i < p.max_count; i++) s += (i ? "," : "") + std::to_string(p.id);
  return s;
}

bool create_partitions(const Profile& p) {
  std::vector<std::string> a = {"mig", "-cgi", build_partition_str(p)};
  logi("Running " + g_smi + join(a));
  Result r = run(a);
  if (r.rc != 0) { loge("failed to create GPU Instances: output: " + r.out); return false; }
  logi("Output:\n " + r.out);
  logi("Running " + g_smi + " mig -cci");
  r = run({"mig", "-cci"});
  if (r.rc != 0) { loge("failed to create compute instances: output: " + r.out); return false; }
  logi("Output:\n " + r.out);
  return true;
}

void status() { Result r = run({}); logi("Output:\n " + r.out); }

// Returns 0 with *size filled ("" when absent), non-zero on parse failure.
int read_partition_size(const std::string& path, std::string* size) {
  std::ifstream f(path);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  size_t a = text.find_first_not_of(" \t\r\n");
  if (a == std::string::npos || text[a] != '{') return 1;
  std::smatch m;
  static const std::regex key(R"re("GPUPartitionSize"\s*:\s*"([^"]*)")re");
  *size = std::regex_search(text, m, key) ? m[1].str() : "";
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    std::string val;
    size_t eq = a.find('=');
    if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); }
    else if (i + 1 < argc && (a == "nvidia-smi-path" || a == "gpu-config")) val = argv[++i];
    if (a == "nvidia-smi-path") g_smi = val;
    else if (a == "gpu-config") g_config = val;
    else if (a == "print-table") { for (const Profile& p : kProfiles) printf("%s %d %d %s\n", p.size, p.id, p.max_count, p.families); return 0; }
    else if (a == "logtostderr" || a == "v" || a == "alsologtostderr") { /* glog-compat no-ops */ }
    else if (a == "h" || a == "help") {
      puts("b200-partition-gpu: bring every GPU of the node to a uniform MIG layout (no-op when it already matches).\n"
           "  -gpu-config PATH        JSON with GPUPartitionSize (default /etc/nvidia/gpu_config.json); missing file or empty size = nothing to do\n"
           "  -nvidia-smi-path PATH   nvidia-smi to drive (default /usr/local/nvidia/bin/nvidia-smi)\n"
           "  -print-table            list the supported sizes: size, profile id, instances per GPU, GPU families\n"
           "exit status: 0 done / nothing to do, 1 failure or reboot requested");
      return 0;
    }
    else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
  }
  struct stat st;
  if (stat(g_config.c_str(), &st) != 0) { logi("No GPU config file given, nothing to do."); return 0; }
  std::string size;
  if (read_partition_size(g_config, &size) != 0) { logi("failed to parse GPU config file, taking no action."); return 0; }
  logi("Using gpu config: {GPUPartitionSize:" + size + "}");
  if (size.empty()) { logi("No GPU partitions are required, exiting"); return 0; }
  if (stat(g_smi.c_str(), &st) != 0) { loge("nvidia-smi path " + g_smi + " not found"); return 1; }
  const Profile* want = find_profile(size);
  if (!want) { loge(size + " is not a valid partition size"); return 1; }

  int mode = current_mig_mode();
  if (mode < 0) { loge("Failed to check if MIG mode is enabled"); return 1; }
  if (mode == 0) {
    logi("MIG mode is not enabled. Enabling now.");
    std::string type = check_gpu_type();
    if (type.empty()) { loge("Failed to check GPU Type"); return 1; }
    logi("Got GPU type used: " + type);
    if (run({"-mig", "1"}).rc != 0) { loge("Failed to enable MIG mode"); return 1; }
    if (type == "NVIDIA A100-SXM4-40GB" || type == "NVIDIA A100-SXM4-80GB") {   // Ampere needs a GPU reset; Hopper/Blackwell do not
      logi("Rebooting node to enable MIG mode");
      if (reboot_node() != 0) loge("Failed to trigger node reboot after enabling MIG mode");
      return 1;
    }
  }
  logi("MIG mode is enabled on all GPUs, proceeding to create GPU partitions.");
  if (matches_desired(*want)) {
    logi("Current GPU partition configuration matches the desired state. No changes needed.");
    status();
    return 0;
  }
  logi("Current GPU partition configuration does not match the desired state. Reconfiguring partitions...");
  if (!cleanup_all()) return 1;
  logi("Creating new GPU partitions");
  if (!create_partitions(*want)) return 1;
  status();
  return 0;
}
