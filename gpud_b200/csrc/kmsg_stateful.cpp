// kmsg_stateful.cpp — the two STATEFUL kmsg matchers of the reference, driven by the line primitives the scan reports.
//   os kernel panic   createKernelPanicMatchFunc   components/os/kmsg_matcher.go:60-125  (+ :127-157 helpers)
//   memory OOM        createMatchFunc              components/memory/kmsg_matcher.go:29-109 (+ :111-208 helpers)
// The reference calls each closure once per kmsg line; every line that holds none of the six primitives
// (GPUD_KIND_OS_PANIC_START .. GPUD_KIND_MEM_OOM_KILLED_PROCESS) only advances the panic matcher's line counter, so the
// machines are run over the hit list plus the unit numbers.  No regex engine here: the captures arrive as buffer spans.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gpud_b200.h"

namespace {

bool go_atoi(const std::string& s, int64_t* out) {   // strconv.Atoi (int is 64-bit)
  size_t i = 0;
  bool neg = false;
  if (!s.empty() && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; i = 1; }
  if (i >= s.size()) return false;
  unsigned long long v = 0;
  const unsigned long long lim = neg ? 0x8000000000000000ull : 0x7fffffffffffffffull;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    const unsigned d = (unsigned)(s[i] - '0');
    if (v > (lim - d) / 10ull) return false;
    v = v * 10ull + d;
  }
  *out = neg ? (int64_t)(0ull - v) : (int64_t)v;
  return true;
}

std::string span(const uint8_t* buf, int64_t off, int32_t len) { return len > 0 ? std::string((const char*)buf + off, (size_t)len) : std::string(); }

// path.Join("/", x) = path.Clean("/" + x) (memory/kmsg_matcher.go:165-166)
std::string path_join_root(const std::string& x) {
  std::vector<std::string> parts;
  size_t i = 0;
  while (i <= x.size()) {
    size_t j = x.find('/', i);
    if (j == std::string::npos) j = x.size();
    const std::string p = x.substr(i, j - i);
    if (p == "..") { if (!parts.empty()) parts.pop_back(); }
    else if (!p.empty() && p != ".") parts.push_back(p);
    i = j + 1;
  }
  std::string o = "/";
  for (size_t k = 0; k < parts.size(); ++k) { if (k) o += "/"; o += parts[k]; }
  return o;
}

struct OOMInstance {           // memory/kmsg_matcher.go:111-122
  int64_t pid = 0;
  std::string process, container = "/", victim = "/", constraint;
  std::string summary() const {  // :124-137
    std::string m = victim == "/" ? "System OOM encountered" : "OOM encountered";
    if (!process.empty() && pid != 0) m += ", victim process: " + process + ", pid: " + std::to_string(pid);
    return m;
  }
};

constexpr int kMaxLinesAfterPanicStart = 10;     // os/kmsg_matcher.go:64
const char kPanicFallback[] = "Kernel panic detected (no CPU/PID info found)";

}  // namespace

struct gpud_kmsg_stateful {
  // kernel panic
  bool panic_reading = false;
  int64_t panic_start = 0;       // global line number of the start line
  // OOM
  bool oom_reading = false;
  OOMInstance oom;
  int64_t base = 0;              // global line number of unit 0 of the scan being fed
};

namespace {

struct Sink {
  gpud_kmsg_event* out;
  int32_t cap, n = 0;
  void emit(int64_t unit, const char* comp, const char* ev, const std::string& msg) {
    if (n < cap) {
      gpud_kmsg_event& e = out[n];
      memset(&e, 0, sizeof e);
      e.unit_index = unit;
      snprintf(e.component, sizeof e.component, "%s", comp);
      snprintf(e.event, sizeof e.event, "%s", ev);
      snprintf(e.message, sizeof e.message, "%s", msg.c_str());
    }
    ++n;
  }
};

// the panic matcher's fallback fires on the 10th line after the start, whatever that line holds: settle it as soon as
// the stream is known to have reached that line (g_upto = global number of the last line known to exist)
void panic_settle(gpud_kmsg_stateful* st, int64_t g_upto, Sink& sink) {
  if (st->panic_reading && g_upto >= st->panic_start + kMaxLinesAfterPanicStart) {
    sink.emit(st->panic_start + kMaxLinesAfterPanicStart - st->base, "os", "kernel_panic", kPanicFallback);
    st->panic_reading = false;
  }
}

}  // namespace

extern "C" int32_t gpud_kmsg_stateful_create(gpud_kmsg_stateful** out) {
  if (!out) return GPUD_E_INVALID;
  *out = new gpud_kmsg_stateful();
  return GPUD_OK;
}
extern "C" void gpud_kmsg_stateful_destroy(gpud_kmsg_stateful* st) { delete st; }

static int32_t stateful_feed(gpud_kmsg_stateful* st, const gpud_xid_hit* hits, int64_t n_hits, const uint8_t* buf, int64_t n_units,
                             gpud_kmsg_event* out, int32_t cap, int32_t* n_out) {
  if (!st || n_hits < 0 || (n_hits && (!hits || !buf)) || n_units < 0 || cap < 0 || (cap && !out)) return GPUD_E_INVALID;
  Sink sink{out, cap};
  int64_t i = 0;
  while (i < n_hits) {
    // the primitives of one unit, in the order the reference tests them
    const int64_t u = hits[i].unit_index;
    const gpud_xid_hit* prim[GPUD_KIND_COUNT] = {nullptr};
    int64_t j = i;
    for (; j < n_hits && hits[j].unit_index == u; ++j)
      if (hits[j].kind >= GPUD_KIND_OS_PANIC_START && hits[j].kind < GPUD_KIND_COUNT) prim[hits[j].kind] = &hits[j];
    i = j;
    const int64_t g = st->base + u;

    // ---- kernel panic (os/kmsg_matcher.go:66-124); the matcher sees every line, so first account for the silent ones
    panic_settle(st, g - 1, sink);
    if (prim[GPUD_KIND_OS_PANIC_START]) {
      if (st->panic_reading) sink.emit(u, "os", "kernel_panic", kPanicFallback);   // a new start while reading (:70-74)
      st->panic_reading = true;
      st->panic_start = g;
    } else if (st->panic_reading) {
      bool done = false;
      if (const gpud_xid_hit* h = prim[GPUD_KIND_OS_PANIC_CPU_PID]) {            // extractCPUandPID (:136-157)
        int64_t cpu, pid;
        if (go_atoi(span(buf, h->dev_off, h->dev_len), &cpu) && go_atoi(span(buf, h->pid_off, h->pid_len), &pid) && pid >= 0) {
          char m[440];
          snprintf(m, sizeof m, "Kernel panic detected - CPU: %lld, PID: %lld, Process: %s", (long long)cpu, (long long)pid,
                   span(buf, h->pname_off, h->pname_len).c_str());
          sink.emit(u, "os", "kernel_panic", m);
          st->panic_reading = false;
          done = true;
        }
      }
      if (!done) panic_settle(st, g, sink);
    }

    // ---- OOM (memory/kmsg_matcher.go:33-108)
    if (prim[GPUD_KIND_MEM_OOM_START]) {
      st->oom_reading = true;
      st->oom = OOMInstance();
    } else if (st->oom_reading) {
      bool container_found = false, dropped = false;
      if (const gpud_xid_hit* h = prim[GPUD_KIND_MEM_OOM_CONTAINER]) {          // getContainerName (:169-189)
        st->oom.container = span(buf, h->inj_off, h->inj_len);
        st->oom.victim = span(buf, h->unit_name_off, h->unit_name_len);
        st->oom.constraint = span(buf, h->dev_off, h->dev_len);
        int64_t pid;
        if (!go_atoi(span(buf, h->pid_off, h->pid_len), &pid)) { st->oom_reading = false; dropped = true; }
        else { st->oom.pid = pid; st->oom.process = span(buf, h->pname_off, h->pname_len); container_found = true; }
      } else if (const gpud_xid_hit* h = prim[GPUD_KIND_MEM_OOM_LEGACY_CONTAINER]) {   // getLegacyContainerName (:159-167)
        st->oom.container = path_join_root(span(buf, h->dev_off, h->dev_len));
        st->oom.victim = path_join_root(span(buf, h->unit_name_off, h->unit_name_len));
      }
      if (!dropped) {
        if (container_found && st->oom.pid != 0) {
          sink.emit(u, "memory", "OOM", st->oom.summary());
          st->oom_reading = false;
        } else if (!container_found) {
          if (const gpud_xid_hit* h = prim[GPUD_KIND_MEM_OOM_KILLED_PROCESS]) {  // getProcessNamePid (:191-208)
            int64_t pid;
            if (!go_atoi(span(buf, h->pid_off, h->pid_len), &pid)) st->oom_reading = false;
            else {
              st->oom.pid = pid;
              st->oom.process = span(buf, h->pname_off, h->pname_len);
              sink.emit(u, "memory", "OOM", st->oom.summary());
              st->oom_reading = false;
            }
          }
        }
      }
    }
  }
  panic_settle(st, st->base + n_units - 1, sink);
  st->base += n_units;
  if (n_out) *n_out = sink.n;
  return sink.n > cap ? GPUD_E_CAPACITY : GPUD_OK;
}


extern "C" int32_t gpud_kmsg_stateful_feed(gpud_kmsg_stateful* st, const gpud_xid_hit* hits, int64_t n_hits, const uint8_t* buf, int64_t n_units,
                                           gpud_kmsg_event* out, int32_t cap, int32_t* n_out) {
  return stateful_feed(st, hits, n_hits, buf, n_units, out, cap, n_out);
}

// The same with the units the kmsg watcher would have dropped as duplicates (pkg/kmsg/watcher.go:281-286, see gpud_kmsg_dedup_units)
// taken out first: the matchers never see such a line, so it neither matches nor counts towards the panic matcher's ten lines.
// Events are reported with the unit numbers of the scanned buffer.
extern "C" int32_t gpud_kmsg_stateful_feed_units(gpud_kmsg_stateful* st, const gpud_xid_hit* hits, int64_t n_hits, const uint8_t* buf, int64_t n_units,
                                                 const uint8_t* dropped, gpud_kmsg_event* out, int32_t cap, int32_t* n_out) {
  if (!dropped) return stateful_feed(st, hits, n_hits, buf, n_units, out, cap, n_out);
  if (!st || n_hits < 0 || (n_hits && (!hits || !buf)) || n_units < 0 || cap < 0 || (cap && !out)) return GPUD_E_INVALID;
  std::vector<int64_t> kept;                       // filtered unit number -> unit number in the buffer
  std::vector<int64_t> renum((size_t)n_units, -1);
  for (int64_t u = 0; u < n_units; ++u)
    if (!dropped[u]) { renum[(size_t)u] = (int64_t)kept.size(); kept.push_back(u); }
  std::vector<gpud_xid_hit> h2;
  for (int64_t i = 0; i < n_hits; ++i) {
    const int64_t u = hits[i].unit_index;
    if (u < 0 || u >= n_units || renum[(size_t)u] < 0) continue;
    h2.push_back(hits[i]);
    h2.back().unit_index = renum[(size_t)u];
  }
  int32_t n = 0;
  const int32_t rc = stateful_feed(st, h2.data(), (int64_t)h2.size(), buf, (int64_t)kept.size(), out, cap, &n);
  for (int32_t i = 0; i < n && i < cap; ++i) {
    const int64_t f = out[i].unit_index;
    if (f >= 0 && f < (int64_t)kept.size()) out[i].unit_index = kept[(size_t)f];
  }
  if (n_out) *n_out = n;
  return rc;
}
