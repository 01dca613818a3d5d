// scan_main.cpp — BASELINE configs[0]: a `gpud scan`-shaped one-shot on a CPU-only host (plumbing, not the hot path).
// pkg/scan/scan.go:74-102 builds every registered component, skips the unsupported ones and prints Check()'s summary; on a host
// without a GPU what is left of the reference's list (components/all/all.go:53-87) is cpu, memory, os and friends.  This driver
// runs those three checks once and prints, per component, the header line `<mark> <summary>` like printSummary (scan.go:20-28)
// and the apiv1.HealthState JSON (api/v1/types.go:50-94) a /v1/states reader would get.  No CUDA, no library: it must run where
// there is no driver.  The rules and their wording follow components/{cpu,memory,os}/component.go (file:line at each rule).
#include <dirent.h>
#include <stdio.h>
#include <string.h>
#include <sys/utsname.h>
#include <time.h>
#include <unistd.h>

#include <string>

namespace {

struct State { std::string component, health = "Healthy", reason = "ok", detail; };

void print_state(const State& s) {
  const bool ok = s.health == "Healthy";
  printf("%s %s\n%s\n", ok ? "\xe2\x9c\x94" : "\xe2\x9a\xa0", s.reason.c_str(), s.detail.c_str());    // printSummary: header + String()
  char tb[40];
  time_t t = time(nullptr);
  struct tm tmv;
  gmtime_r(&t, &tmv);
  strftime(tb, sizeof tb, "%Y-%m-%dT%H:%M:%SZ", &tmv);
  printf("{\"time\":\"%s\",\"component\":\"%s\",\"name\":\"%s\",\"health\":\"%s\",\"reason\":\"%s\"}\n\n", tb, s.component.c_str(), s.component.c_str(), s.health.c_str(),
         s.reason.c_str());
}

bool read_first_line(const char* path, char* buf, size_t cap) {
  FILE* f = fopen(path, "r");
  if (!f) return false;
  const bool ok = fgets(buf, (int)cap, f) != nullptr;
  fclose(f);
  return ok;
}

State check_cpu() {                                  // components/cpu/component.go:150-226: usage + load average, reason "ok"
  State s;
  s.component = "cpu";
  struct utsname u;
  uname(&u);
  char line[256];
  double l1 = 0, l5 = 0, l15 = 0;
  if (!read_first_line("/proc/loadavg", line, sizeof line) || sscanf(line, "%lf %lf %lf", &l1, &l5, &l15) != 3) {
    s.health = "Unhealthy"; s.reason = "error calculating load average";          // :212
    return s;
  }
  unsigned long long a[8] = {0}, b[8] = {0};
  auto stat = [&](unsigned long long* v) {
    char l[512];
    return read_first_line("/proc/stat", l, sizeof l) && sscanf(l, "cpu %llu %llu %llu %llu %llu %llu %llu %llu", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6, v + 7) >= 4;
  };
  if (!stat(a)) { s.health = "Unhealthy"; s.reason = "error calculating CPU usage"; return s; }   // :176
  usleep(100000);
  if (!stat(b)) { s.health = "Unhealthy"; s.reason = "error calculating CPU usage"; return s; }
  unsigned long long tot = 0, idle = (b[3] - a[3]) + (b[4] - a[4]);
  for (int i = 0; i < 8; ++i) tot += b[i] - a[i];
  char d[512];
  snprintf(d, sizeof d, "arch: %s, logical cores: %ld, used: %.2f %%, load avg 1/5/15 min: %.2f / %.2f / %.2f", u.machine, sysconf(_SC_NPROCESSORS_ONLN),
           tot ? 100.0 * (double)(tot - idle) / (double)tot : 0.0, l1, l5, l15);
  s.detail = d;
  return s;
}

State check_memory() {                               // components/memory/component.go:170-226: virtual memory, reason "ok"
  State s;
  s.component = "memory";
  FILE* f = fopen("/proc/meminfo", "r");
  if (!f) { s.health = "Unhealthy"; s.reason = "error getting virtual memory"; return s; }        // :182
  unsigned long long total = 0, avail = 0, free_kb = 0, v;
  char k[64], line[256];
  while (fgets(line, sizeof line, f))
    if (sscanf(line, "%63[^:]: %llu", k, &v) == 2) {
      if (!strcmp(k, "MemTotal")) total = v; else if (!strcmp(k, "MemAvailable")) avail = v; else if (!strcmp(k, "MemFree")) free_kb = v;
    }
  fclose(f);
  char d[256];
  snprintf(d, sizeof d, "total: %llu MiB, available: %llu MiB, used: %llu MiB, free: %llu MiB", total >> 10, avail >> 10, (total - avail) >> 10, free_kb >> 10);
  s.detail = d;
  return s;
}

State check_os() {                                   // components/os/component.go:330-496
  State s;
  s.component = "os";
  char line[256];
  double up = 0;
  if (!read_first_line("/proc/uptime", line, sizeof line) || sscanf(line, "%lf", &up) != 1) { s.health = "Unhealthy"; s.reason = "error getting uptime"; return s; }   // :336
  int zombies = 0, pids = 0;
  DIR* d = opendir("/proc");
  if (!d) { s.health = "Unhealthy"; s.reason = "error getting process count"; return s; }          // :352
  while (struct dirent* e = readdir(d)) {
    if (e->d_name[0] < '0' || e->d_name[0] > '9') continue;
    ++pids;
    char p[300], st[512];
    snprintf(p, sizeof p, "/proc/%s/stat", e->d_name);
    if (read_first_line(p, st, sizeof st)) { const char* r = strrchr(st, ')'); if (r && r[1] == ' ' && r[2] == 'Z') ++zombies; }
  }
  closedir(d);
  unsigned long long fh_alloc = 0, fh_unused = 0, fh_max = 0;
  if (read_first_line("/proc/sys/fs/file-nr", line, sizeof line)) sscanf(line, "%llu %llu %llu", &fh_alloc, &fh_unused, &fh_max);
  const int z_deg = 1000, z_unh = 2000;                                   // :663-664
  const double pid_max = 900000.0, fh_cap = 10000000.0;                   // DefaultMaxRunningPIDs / DefaultMaxAllocatedFileHandles, :41-45
  const double pid_pct = 100.0 * pids / pid_max, fh_pct = 100.0 * (double)fh_alloc / (fh_max && (double)fh_max < fh_cap ? (double)fh_max : fh_cap);
  char r[160];
  if (zombies > z_unh) { s.health = "Unhealthy"; snprintf(r, sizeof r, "too many zombie processes (unhealthy state threshold: %d)", z_unh); s.reason = r; }          // :369
  else if (zombies > z_deg) { s.health = "Degraded"; snprintf(r, sizeof r, "too many zombie processes (degraded state threshold: %d)", z_deg); s.reason = r; }     // :377
  else if (pid_pct > 95.0) { s.health = "Unhealthy"; snprintf(r, sizeof r, "too many running pids (unhealthy state percent threshold: %.2f %%)", 95.0); s.reason = r; }   // :453
  else if (pid_pct > 80.0) { s.health = "Degraded"; snprintf(r, sizeof r, "too many running pids (degraded state percent threshold: %.2f %%)", 80.0); s.reason = r; }     // :460
  else if (fh_pct > 95.0) { s.health = "Unhealthy"; snprintf(r, sizeof r, "too many allocated file handles (unhealthy state percent threshold: %.2f %%)", 95.0); s.reason = r; }   // :481
  else if (fh_pct > 80.0) { s.health = "Degraded"; snprintf(r, sizeof r, "too many allocated file handles (degraded state percent threshold: %.2f %%)", 80.0); s.reason = r; }     // :488
  char dd[320];
  snprintf(dd, sizeof dd, "uptime: %.0f s, processes: %d (zombies %d), running pids: %.2f %% of %.0f, allocated file handles: %llu (%.2f %%)", up, pids, zombies, pid_pct,
           pid_max, fh_alloc, fh_pct);
  s.detail = dd;
  return s;
}

}  // namespace

int main() {
  printf("\n\n\xe2\x8c\x9b scanning the host (GOOS linux)\n\n");                                     // scan.go:37
  const bool gpu = access("/dev/nvidiactl", F_OK) == 0;
  printf("%s\n\n", gpu ? "NVIDIA driver node present: the accelerator components are served by libgpud_b200.so (gpud_component_*)" :
                         "no NVIDIA driver on this host: the accelerator components report unsupported and are skipped (scan.go:95-97)");
  print_state(check_cpu());
  print_state(check_memory());
  print_state(check_os());
  printf("\n\xe2\x9c\x94 scan complete\n\n");                                                        // scan.go:104
  return 0;
}
