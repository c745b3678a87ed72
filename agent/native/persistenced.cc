// b200-persistenced — sidecar that starts nvidia-persistenced on confidential-GPU nodes, flips the GPUs to
// the conf-compute "ready" state, optionally starts nvidia-gridd on G4 machine types, then sleeps until SIGTERM.
//
// Contract: reference nvidia-persistenced-installer/nvidia_persistenced_installer.go:47-267:
//   * enabled iff <cgpu-config> (trim " \r\n\0", lower-case) is "tdx" or "sev"; a missing file means disabled
//     (not an error), any other read error is fatal
//   * driver major from /proc/driver/nvidia/version (\d+\.\d+\.\d+, 3-digit major); >= 550 adds --uvm-persistence-mode;
//     always --nvidia-cfg-path=<prefix>/lib64
//   * after -ready-delay-ms: `nvidia-smi conf-compute -srs 1`; output containing "No devices were found" => reboot
//     (kill(1, SIGRTMIN+5)); any failure => exit 1
//   * gridd only for g4-standard-{6,12,24}: wait (10 s poll) for <prefix>/bin/nvidia-gridd and the host loader, then run
//     it through <root>/lib64/ld-linux-x86-64.so.2 --library-path <prefix>/gridd-libs:<root>/lib64:<root>/usr/lib64
//   * never exits on the disabled path (a sidecar that exits restarts forever); blocks on SIGINT/SIGTERM
// Test seams (env): B200_PERSISTENCED_PROC_VERSION, B200_PERSISTENCED_LDCONF, B200_PERSISTENCED_REBOOT_HOOK,
// B200_PERSISTENCED_POLL_MS, and --oneshot to return instead of blocking.
#include <errno.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

namespace {

std::string g_prefix = "/usr/local/nvidia";
std::string g_cgpu = "/etc/nvidia/confidential_node_type.txt";
std::string g_machine = "/etc/nvidia/machine_type.txt";
long g_ready_delay_ms = 1000;
bool g_oneshot = false;

void logi(const std::string& s) { fprintf(stderr, "I persistenced] %s\n", s.c_str()); }
void loge(const std::string& s) { fprintf(stderr, "E persistenced] %s\n", s.c_str()); }
const char* env_or(const char* k, const char* d) { const char* v = getenv(k); return v && *v ? v : d; }

// 0 ok, 1 not found, 2 other error
int read_file(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return errno == ENOENT ? 1 : (access(path.c_str(), F_OK) != 0 ? 1 : 2);
  std::stringstream ss; ss << f.rdbuf(); *out = ss.str();
  return 0;
}

std::string trim_set(std::string s, const std::string& set) {
  size_t a = 0, b = s.size();
  while (a < b && set.find(s[a]) != std::string::npos) a++;
  while (b > a && set.find(s[b - 1]) != std::string::npos) b--;
  return s.substr(a, b - a);
}

struct Result { int rc; std::string out; };
Result run(const std::vector<std::string>& argv_s, bool capture) {
  int fds[2] = {-1, -1};
  if (capture && pipe(fds) != 0) return {-1, ""};
  pid_t pid = fork();
  if (pid < 0) return {-1, ""};
  if (pid == 0) {
    if (capture) { dup2(fds[1], 1); dup2(fds[1], 2); close(fds[0]); close(fds[1]); }
    std::vector<char*> argv;
    for (const std::string& a : argv_s) argv.push_back(const_cast<char*>(a.c_str()));
    argv.push_back(nullptr);
    execvp(argv[0], argv.data());
    _exit(127);
  }
  std::string out;
  if (capture) {
    close(fds[1]);
    char buf[4096]; ssize_t n;
    while ((n = read(fds[0], buf, sizeof(buf))) > 0) out.append(buf, (size_t)n);
    close(fds[0]);
  }
  int st = 0; waitpid(pid, &st, 0);
  return {WIFEXITED(st) ? WEXITSTATUS(st) : -1, out};
}

// 1 enabled, 0 disabled, -1 fatal
int confidential_enabled() {
  std::string text;
  int rc = read_file(g_cgpu, &text);
  if (rc == 1) { logi("confidential node type file not found at " + g_cgpu + ", skipping persistenced installation"); return 0; }
  if (rc != 0) { loge("cannot read " + g_cgpu); return -1; }
  std::string t = trim_set(text, std::string(" \r\n\0", 4));
  std::transform(t.begin(), t.end(), t.begin(), ::tolower);
  return (t == "tdx" || t == "sev") ? 1 : 0;
}

int driver_major() {
  std::string text;
  const std::string path = env_or("B200_PERSISTENCED_PROC_VERSION", "/proc/driver/nvidia/version");
  if (read_file(path, &text) != 0) { loge("failed to read nvidia gpu driver version at " + path); return -1; }
  std::smatch m;
  static const std::regex ver(R"((\d+)\.\d+\.\d+)");
  if (!std::regex_search(text, m, ver)) { loge("failed to read nvidia gpu driver version at " + path); return -1; }
  const std::string major = m[1];
  if (major.size() != 3) { loge("invalid nvidia gpu driver version: " + m[0].str()); return -1; }
  return atoi(major.c_str());
}

bool update_ld_cache() {
  const std::string conf = env_or("B200_PERSISTENCED_LDCONF", "/etc/ld.so.conf.d/nvidia.conf");
  std::ofstream f(conf);
  if (!f) { loge("failed to update ld cache: cannot write " + conf); return false; }
  f << g_prefix << "/lib64";
  f.close();
  if (run({env_or("B200_PERSISTENCED_LDCONFIG", "ldconfig")}, false).rc != 0) { loge("failed to update ld cache: ldconfig failed"); return false; }
  return true;
}

bool enable_persistence_mode() {
  logi("Starting NVIDIA persistence daemon.");
  int major = driver_major();
  if (major < 0) return false;
  std::vector<std::string> cmd = {g_prefix + "/bin/nvidia-persistenced"};
  if (major >= 550) { cmd.push_back("--uvm-persistence-mode"); logi("using --uvm-persistence-mode"); }   // UVM persistence exists from R550 on
  cmd.push_back("--nvidia-cfg-path=" + g_prefix + "/lib64");
  if (run(cmd, false).rc != 0) return false;
  logi("NVIDIA Persistence Mode Enabled.");
  return true;
}

int reboot_node() {
  if (const char* hook = getenv("B200_PERSISTENCED_REBOOT_HOOK")) { std::ofstream(hook) << "reboot\n"; return 0; }
  return kill(1, SIGRTMIN + 5);
}

void poll_sleep() { usleep((useconds_t)atol(env_or("B200_PERSISTENCED_POLL_MS", "10000")) * 1000); }

bool enable_gridd(const std::string& machine_type) {
  if (machine_type != "g4-standard-6" && machine_type != "g4-standard-12" && machine_type != "g4-standard-24") {
    logi("Machine type " + machine_type + " does not require nvidia-gridd.");
    return true;
  }
  const std::string gridd = g_prefix + "/bin/nvidia-gridd", libs = g_prefix + "/gridd-libs";
  struct stat st;
  logi("Waiting for " + gridd + " to appear...");
  while (stat(gridd.c_str(), &st) != 0) poll_sleep();
  const std::string root = env_or("ROOT_MOUNT_DIR", "/root");
  const std::string linker = root + "/lib64/ld-linux-x86-64.so.2";
  logi("Waiting for dynamic linker " + linker + " to appear...");
  while (stat(linker.c_str(), &st) != 0) poll_sleep();
  logi("Starting nvidia-gridd daemon via host dynamic linker: " + linker);
  Result r = run({linker, "--library-path", libs + ":" + root + "/lib64:" + root + "/usr/lib64", gridd}, true);
  if (r.rc != 0) { loge("failed to run nvidia-gridd, output: " + r.out); return false; }
  logi("nvidia-gridd daemon started.");
  return true;
}

volatile sig_atomic_t g_signal = 0;
void on_signal(int s) { g_signal = s; }

}  // namespace

int main(int argc, char** argv) {
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    std::string val;
    size_t eq = a.find('=');
    bool has_val = eq != std::string::npos;
    if (has_val) { val = a.substr(eq + 1); a = a.substr(0, eq); }
    auto need = [&]() { if (!has_val && i + 1 < argc) val = argv[++i]; };
    if (a == "container-path") { need(); g_prefix = val; }
    else if (a == "cgpu-config") { need(); g_cgpu = val; }
    else if (a == "machine-type-file") { need(); g_machine = val; }
    else if (a == "ready-delay-ms") { need(); g_ready_delay_ms = atol(val.c_str()); }
    else if (a == "oneshot") g_oneshot = true;
    else if (a == "logtostderr" || a == "v") { /* glog-compat */ }
    else if (a == "h" || a == "help") {
      puts("b200-persistenced: sidecar for confidential-GPU and vGPU nodes. Starts nvidia-persistenced (with --uvm-persistence-mode from R550),\n"
           "sets the conf-compute ready state, reboots the node on \"No devices were found\", starts nvidia-gridd on G4 machine types, then waits for SIGTERM.\n"
           "  -container-path PATH      where the driver directory is mounted (default /usr/local/nvidia)\n"
           "  -cgpu-config PATH         confidential node type file: tdx | sev enables the daemon (default /etc/nvidia/confidential_node_type.txt)\n"
           "  -machine-type-file PATH   machine type file for the gridd decision (default /etc/nvidia/machine_type.txt)\n"
           "  -ready-delay-ms N         wait before setting the ready state (default 1000)");
      return 0;
    }
    else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
  }
  int enabled = confidential_enabled();
  if (enabled < 0) { loge("parseCGPUConfig failed"); return 1; }
  if (enabled) {
    if (!update_ld_cache()) return 1;                 // so nvidia-smi resolves its libraries from the mounted install dir
    if (!enable_persistence_mode()) { loge("failed to start persistence mode"); return 1; }
    usleep((useconds_t)g_ready_delay_ms * 1000);      // starting workloads right after the daemon sometimes errors
    Result r = run({g_prefix + "/bin/nvidia-smi", "conf-compute", "-srs", "1"}, true);
    if (r.rc != 0) {
      logi("failed to set gpu to ready state, output: " + r.out);
      if (r.out.find("No devices were found") != std::string::npos) {
        logi("No devices were found, rebooting node to resolve");
        if (reboot_node() != 0) loge("Failed to trigger node reboot");
      }
      return 1;
    }
    logi("Confidential GPU is ready.");
  } else {
    logi("Confidential GPU is NOT enabled, skipping nvidia persistenced enablement.");
  }
  std::string machine;
  int rc = read_file(g_machine, &machine);
  if (rc == 1) logi("machine type file not found at " + g_machine);
  else if (rc != 0) loge("Failed to get machine type");
  else if (!enable_gridd(trim_set(machine, " \t\r\n"))) loge("Failed to enable nvidia-gridd");
  if (g_oneshot) return 0;
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  while (!g_signal) pause();
  printf("Received signal: %d. Shutting down...\n", (int)g_signal);
  return 0;
}
