/* oracle.c — plain-C CPU restatement used as the checker and as the timed CPU baseline.
 *
 * TEST INFRASTRUCTURE ONLY: nothing under gpud_b200/ links, loads or executes this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs use it.
 *
 * Part 1 (this section): windowed aggregates.  PARITY UNPINNED — the reference (leptonai/gpud) has no windowed
 *   min/max/mean/EMA/p99; the definitions are oracle/SPEC.md.  The only reference-derived rule is the strict `>`
 *   of n_over (components/accelerator/nvidia/temperature/component.go:228,240).
 * Part 2 (regex_bt.c + below): xid.Match / sxid.Match over the reference's verbatim regex strings
 *   (components/accelerator/nvidia/xid/kmsg.go:22-43,202-268 ; sxid/kmsg.go:17-73) with the catalog lookups of
 *   xid/xid.go:74-117,2954-3305.  Pinned by tests/golden/ (tests/test_oracle_c.py).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <sched.h>
#include <unistd.h>

static inline uint64_t f64_key(double x) {
  uint64_t b;
  memcpy(&b, &x, 8);
  return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
static inline double key_f64(uint64_t k) {
  uint64_t b = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
  double x;
  memcpy(&x, &b, 8);
  return x;
}

/* k-th smallest (0-based) by quickselect on keys; scratch is clobbered */
static uint64_t select_kth(uint64_t* a, int n, int k) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    uint64_t pivot = a[lo + (hi - lo) / 2];
    int i = lo, j = hi;
    while (i <= j) {
      while (a[i] < pivot) ++i;
      while (a[j] > pivot) --j;
      if (i <= j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; ++i; --j; }
    }
    if (k <= j) hi = j; else if (k >= i) lo = i; else return a[k];
  }
  return a[k];
}

/* One field: x[0..n) chronological.  Outputs are arrays of ceil(n/W).  `e_io` carries the EMA (in: e_{-1}). */
void orc_window_aggregates(const double* x, int64_t n, int32_t W, double thr, double alpha, int32_t q_num, int32_t q_den,
                           double* o_min, double* o_max, double* o_mean, double* o_ema, double* o_p99, uint32_t* o_nover) {
  if (alpha <= 0.0) { alpha = 2.0 / (W + 1.0); if (alpha > 0.9999) alpha = 0.9999; }
  if (q_num == 0 && q_den == 0) { q_num = 99; q_den = 100; }
  uint64_t* scratch = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)W);
  double e = n > 0 ? x[0] : 0.0;
  const int64_t nw = (n + W - 1) / W;
  for (int64_t w = 0; w < nw; ++w) {
    const double* s = x + w * W;
    const int m = (int)((n - w * W) < W ? (n - w * W) : W);
    uint64_t kmin = ~0ull, kmax = 0;
    double sum = 0.0;
    uint32_t nov = 0;
    for (int i = 0; i < m; ++i) {
      const uint64_t k = f64_key(s[i]);
      scratch[i] = k;
      if (k < kmin) kmin = k;
      if (k > kmax) kmax = k;
      sum += s[i];
      nov += s[i] > thr;
      e = alpha * s[i] + (1.0 - alpha) * e;
    }
    long long r = ((long long)m * q_num + q_den - 1) / q_den;
    if (r < 1) r = 1;
    if (r > m) r = m;
    o_min[w] = key_f64(kmin);
    o_max[w] = key_f64(kmax);
    o_mean[w] = sum / (double)m;
    o_ema[w] = e;
    o_p99[w] = key_f64(select_kth(scratch, m, (int)r - 1));
    o_nover[w] = nov;
  }
  free(scratch);
}

/* ---- tiny pthread parallel-for (the image's default gcc has no libgomp spec; pthreads is always there) ---- */
typedef void (*orc_job_fn)(void* arg, int64_t index);
typedef struct { orc_job_fn fn; void* arg; int64_t n; int64_t next; pthread_mutex_t mu; } orc_pool;
static void* orc_worker(void* p) {
  orc_pool* P = (orc_pool*)p;
  for (;;) {
    pthread_mutex_lock(&P->mu);
    const int64_t i = P->next < P->n ? P->next++ : -1;
    pthread_mutex_unlock(&P->mu);
    if (i < 0) return NULL;
    P->fn(P->arg, i);
  }
}
/* CPUs this process may actually use: the affinity mask, capped by the cgroup CPU quota (cpu.max "quota period" of cgroup v2,
 * cpu.cfs_quota_us / cpu.cfs_period_us of v1) - `nproc` of a container reports the host's logical CPUs, not its share. */
int32_t orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0 && c < n) n = c; }
  long long quota = -1, period = -1;
  FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");
  if (f) {
    char q[32] = {0};
    if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0) quota = atoll(q);
    fclose(f);
  } else {
    f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r");
    if (f) { if (fscanf(f, "%lld", &quota) != 1) quota = -1; fclose(f); }
    f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
    if (f) { if (fscanf(f, "%lld", &period) != 1) period = -1; fclose(f); }
  }
  if (quota > 0 && period > 0) { const long c = (long)((quota + period - 1) / period); if (c > 0 && c < n) n = c; }
  return n > 0 ? (int32_t)n : 1;
}
void orc_parallel_for(orc_job_fn fn, void* arg, int64_t n, int32_t threads) {
  if (threads <= 0) threads = orc_max_threads();
  if (threads > 256) threads = 256;
  if (threads > n) threads = (int32_t)(n > 0 ? n : 1);
  orc_pool P = {fn, arg, n, 0, PTHREAD_MUTEX_INITIALIZER};
  pthread_t th[256];
  for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, orc_worker, &P);
  orc_worker(&P);
  for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
}

typedef struct { const double* ring; int64_t n, nw; int32_t W; const double* thr; double alpha; int32_t q_num, q_den;
                 double *o_min, *o_max, *o_mean, *o_ema, *o_p99; uint32_t* o_nover; } orc_win_job;
static void orc_win_one(void* a, int64_t f) {
  orc_win_job* j = (orc_win_job*)a;
  orc_window_aggregates(j->ring + f * j->n, j->n, j->W, j->thr ? j->thr[f] : INFINITY, j->alpha, j->q_num, j->q_den, j->o_min + f * j->nw,
                        j->o_max + f * j->nw, j->o_mean + f * j->nw, j->o_ema + f * j->nw, j->o_p99 + f * j->nw, j->o_nover + f * j->nw);
}
/* Field-major ring [F][n] -> results [F][nw]; fields are partitioned across `threads` worker threads. */
void orc_windows_fields(const double* ring, int32_t F, int64_t n, int32_t W, const double* thr, double alpha, int32_t q_num, int32_t q_den,
                        int32_t threads, double* o_min, double* o_max, double* o_mean, double* o_ema, double* o_p99, uint32_t* o_nover) {
  orc_win_job j = {ring, n, (n + W - 1) / W, W, thr, alpha, q_num, q_den, o_min, o_max, o_mean, o_ema, o_p99, o_nover};
  orc_parallel_for(orc_win_one, &j, F, threads);
}

/* =====================================================================================================================
 * Part 2: xid.Match / sxid.Match over the reference's verbatim regex strings (engine: regex_bt.c).
 * ===================================================================================================================== */
#include "regex_bt.h"

#include <stdio.h>

#include "oracle_catalog_data.inc"   /* the oracle's own copy of the generated rows (tools/gen_catalog.py); nothing under gpud_b200/ is compiled here */

/* components/accelerator/nvidia/xid/kmsg.go:22,29,38,43 ; sxid/kmsg.go:17,20 — verbatim */
static const char* RX[28] = {
    "NVRM: Xid \\(((?:PCI:)?[0-9a-fA-F:]+)\\).*?: (\\d+),",
    "NVRM: Xid \\(PCI:([0-9a-fA-F:]+)\\): (\\d+)(?:, pid=(\\d+), name=([^,]+))?, ([A-Z_]+(?:/[A-Z_]+)?)\\s+(Nonfatal|Fatal)\\s+(XC[01])\\s+(i\\d+)\\s+Link\\s+(-?\\d+)\\s+\\((0x[0-9a-fA-F]+)\\s+(0x[0-9a-fA-F]+)(?:\\s+(0x[0-9a-fA-F]+))?(?:\\s+(0x[0-9a-fA-F]+))?(?:\\s+(0x[0-9a-fA-F]+))?(?:\\s+(0x[0-9a-fA-F]+))?",
    "(?s)NVRM:\\s+The NVIDIA GPU ((?:[0-9a-fA-F]{4}:)?[0-9a-fA-F]{2}:[0-9a-fA-F]{2})\\.0.*?fallen off the bus and is not responding to commands\\.",
    "NVRM:\\s+GPU ((?:[0-9a-fA-F]{4}:)?[0-9a-fA-F]{2}:[0-9a-fA-F]{2})\\.0:\\s+GPU has fallen off the bus\\.?",
    "SXid.*?: (\\d+),",
    "SXid \\((PCI:[0-9a-fA-F:\\.]+)\\)",
    /* next matchers on the same scanner: nccl/kmsg_matcher.go:12 ; peermem/kmsg_matcher.go:14 — verbatim */
    ".*segfault at.*in libnccl\\.so.*",
    ".*ERROR detected invalid context, skipping further processing",
    /* infiniband/kmsg_matcher.go:15,25,57 ; cpu/kmsg_matcher.go:18,30 ; os/kmsg_matcher.go:18 ; disk/kmsg_matcher.go:11-55 — verbatim */
    "Detected insufficient power on the PCIe slot \\(([0-9]+W)\\)",
    "Port module event.*High Temperature",
    "mlx5_cmd_out_err.*ACCESS_REG.*failed",
    "(?:INFO: )?task ([^:]+:[\\d]+).+blocked for more than \\d+ seconds",
    "soft lockup - CPU#\\d+ stuck for \\d+s! \\[([^:]+:[\\d]+)\\]",
    "VFS: file-max limit \\d+ reached",
    "md/raid.*: Disk failure on .* detected, failing array",
    ".*Remounting filesystem read-only",
    "block nvme.*: no available path - failing I/O",
    "nvme nvme[0-9]+: I/O .* timeout, reset controller",
    "nvme nvme[0-9]+: Disabling device after reset failure",
    "attempt to access beyond end of device",
    "Buffer I/O error on dev [^ ]+, logical block [0-9]+",
    "I/O error while writing superblock",
    /* line primitives of the two stateful matchers: os/kmsg_matcher.go:35,39 ; memory/kmsg_matcher.go:156,148,143,152 — verbatim */
    "Kernel [Pp]anic",
    "CPU: (\\d+) PID: (\\d+) Comm: (\\S+)",
    "invoked oom-killer:",
    "oom-kill:constraint=(.*),nodemask=(.*),cpuset=(.*),mems_allowed=(.*),oom_memcg=(.*),task_memcg=(.*),task=(.*),pid=(.*),uid=(.*)",
    "Task in (.*) killed as a result of limit of (.*)",
    "Killed process ([0-9]+) \\((.+)\\)"};
#define ORC_N_RX 28
#define ORC_N_EXT 22          /* RX[6 + i] is hit kind 3 + i */
static orx_prog* PR[ORC_N_RX];
static pthread_once_t rx_once = PTHREAD_ONCE_INIT;
static void rx_init(void) { for (int i = 0; i < ORC_N_RX; ++i) PR[i] = orx_compile(RX[i]); }
int32_t orc_regex_ok(void) {
  pthread_once(&rx_once, rx_init);
  for (int i = 0; i < 6; ++i) if (!PR[i]) return 0;
  return 1;
}

typedef struct {
  int64_t line, offset;
  int32_t kind, code, extended, sub_code, severity_fatal;
  int64_t link;
  uint32_t intrinfo, error_status;
  char device[64];
  char unit[64];
} orc_hit;

/* strconv.Atoi: 1 ok / 0 range or syntax error */
static int go_atoi(const char* s, int n, long long* out) {
  int i = 0, neg = 0;
  if (n > 0 && (s[0] == '-' || s[0] == '+')) { neg = s[0] == '-'; i = 1; }
  if (i >= n) return 0;
  unsigned long long v = 0, lim = neg ? 0x8000000000000000ull : 0x7fffffffffffffffull;
  for (; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return 0;
    unsigned d = (unsigned)(s[i] - '0');
    if (v > (lim - d) / 10ull) return 0;
    v = v * 10ull + d;
  }
  *out = neg ? (long long)(0ull - v) : (long long)v;
  return 1;
}
static int go_hex32(const char* s, int n, uint32_t* out) {   /* ParseUint("0x..", 0, 32) */
  unsigned long long v = 0;
  for (int i = 2; i < n; ++i) {
    int c = s[i];
    unsigned d = (c >= '0' && c <= '9') ? c - '0' : (c | 0x20) - 'a' + 10;
    v = v * 16ull + d;
    if (v > 0xffffffffull) return 0;
  }
  *out = (uint32_t)v;
  return 1;
}
static int xid_known(long long code) {
  for (int i = 0; i < GPUD_CAT_N_XID; ++i) if (GPUD_CAT_XID[i].code == code) return 1;
  return 0;
}
static int sxid_known(long long code) {
  for (int i = 0; i < GPUD_CAT_N_SXID; ++i) if (GPUD_CAT_SXID[i].sxid == code) return 1;
  return 0;
}
static void cpy(char* dst, int cap, const char* pre, const char* s, int n) {
  int k = 0;
  for (; pre && pre[k] && k < cap - 1; ++k) dst[k] = pre[k];
  for (int i = 0; i < n && k < cap - 1; ++i) dst[k++] = s[i];
  dst[k] = 0;
}

/* xid.Match (kmsg.go:202-245).  Returns 1 and fills h on a hit. */
int32_t orc_xid_match(const char* s, int32_t n, orc_hit* h) {
  pthread_once(&rx_once, rx_init);
  int c[40];
  memset(h, 0, sizeof *h);
  h->kind = 1;
  if (orx_search(PR[1], s, n, c)) {   /* ExtractNVRMXidInfoExtended (kmsg.go:116-183) */
    long long code, link;
    uint32_t intr, es;
    if (go_atoi(s + c[4], c[5] - c[4], &code) && go_hex32(s + c[20], c[21] - c[20], &intr) && go_hex32(s + c[22], c[23] - c[22], &es) &&
        go_atoi(s + c[18], c[19] - c[18], &link) && xid_known(code)) {   /* detailFromNVLinkInfo needs the base code (xid.go:2955-2958) */
      h->code = (int32_t)code; h->extended = 1; h->intrinfo = intr; h->error_status = es; h->link = link;
      h->sub_code = (int32_t)((intr >> 20) & 0x3F);
      h->severity_fatal = (c[13] - c[12]) == 5;   /* "Fatal" vs "Nonfatal" */
      cpy(h->unit, 64, NULL, s + c[10], c[11] - c[10]);
      cpy(h->device, 64, "PCI:", s + c[2], c[3] - c[2]);
      return 1;
    }
  }
  if (orx_search(PR[0], s, n, c)) {   /* ExtractNVRMXidInfo (kmsg.go:73-80) */
    long long code;
    if (go_atoi(s + c[4], c[5] - c[4], &code) && code != 0) {
      if (!xid_known(code)) return 0;
      h->code = (int32_t)code;
      cpy(h->device, 64, NULL, s + c[2], c[3] - c[2]);
      return 1;
    }
  }
  int ok = orx_search(PR[3], s, n, c);   /* extractFallenOffBusXidInfo (kmsg.go:247-257): single-line first */
  if (!ok) ok = orx_search(PR[2], s, n, c);
  if (ok) {
    if (!xid_known(79)) return 0;
    h->code = 79;
    int colons = 0;
    for (int i = c[2]; i < c[3]; ++i) colons += s[i] == ':';
    cpy(h->device, 64, colons == 1 ? "PCI:0000:" : "PCI:", s + c[2], c[3] - c[2]);   /* normalizePCIBDF (kmsg.go:259-268) */
    return 1;
  }
  return 0;
}

/* sxid.Match (sxid/kmsg.go:58-73) */
int32_t orc_sxid_match(const char* s, int32_t n, orc_hit* h) {
  pthread_once(&rx_once, rx_init);
  int c[8];
  memset(h, 0, sizeof *h);
  h->kind = 2;
  if (!orx_search(PR[4], s, n, c)) return 0;
  long long code;
  if (!go_atoi(s + c[2], c[3] - c[2], &code) || code == 0 || !sxid_known(code)) return 0;
  h->code = (int32_t)code;
  if (orx_search(PR[5], s, n, c)) cpy(h->device, 64, NULL, s + c[2], c[3] - c[2]);
  return 1;
}

/* The Has* functions of the stateless line matchers (e.g. nccl/kmsg_matcher.go:20-25, disk/kmsg_matcher.go:70-125):
 * bit i = pattern of hit kind 3 + i fires.  cap[2*i], cap[2*i+1] = span of capture group 1 (cpu patterns), -1 if none. */
/* one pattern: spans of groups 0..9 into c[20] (-1 = unset); returns 1 on match */
int32_t orc_ext_groups(int32_t kind, const char* s, int32_t n, int32_t* c20) {
  pthread_once(&rx_once, rx_init);
  int c[24];
  for (int i = 0; i < 24; ++i) c[i] = -1;
  if (kind < 3 || kind >= 3 + ORC_N_EXT || !orx_search(PR[6 + kind - 3], s, n, c)) return 0;
  for (int i = 0; i < 20; ++i) c20[i] = c[i];
  return 1;
}

int32_t orc_ext_match(const char* s, int32_t n, int32_t* cap) {
  pthread_once(&rx_once, rx_init);
  int c[24];
  int32_t m = 0;
  for (int i = 0; i < ORC_N_EXT; ++i) {
    if (cap) cap[2 * i] = cap[2 * i + 1] = -1;
    c[2] = c[3] = -1;
    if (orx_search(PR[6 + i], s, n, c)) {
      m |= 1 << i;
      if (cap && (i == 5 || i == 6)) { cap[2 * i] = c[2]; cap[2 * i + 1] = c[3]; }
    }
  }
  return m;
}

/* The reference's buffer-scan form: split on '\n', Match each line (xid/kmsg_test.go:252-267), parallel over byte ranges
 * cut at line boundaries.  Returns the number of hits (may exceed cap; only the first cap are stored, in line order). */
typedef struct { const char* buf; int64_t b, e; orc_hit* hits; int64_t n, cap, lines; int32_t ext; } orc_scan_part;
static void orc_scan_one(void* a, int64_t idx) {
  orc_scan_part* P = (orc_scan_part*)a + idx;
  int64_t ls = P->b, line = 0;
  for (int64_t i = P->b; i <= P->e; ++i) {
    if (i == P->e || P->buf[i] == '\n') {
      const int32_t n = (int32_t)(i - ls);
      /* cheap prefilter: every pattern needs "NVRM:" or "SXid" (Go's regexp does the same literal-prefix skip) */
      if (n >= 4 && (memmem(P->buf + ls, (size_t)n, "NVRM:", 5) || memmem(P->buf + ls, (size_t)n, "SXid", 4))) {
        orc_hit h;
        if (orc_xid_match(P->buf + ls, n, &h)) { h.line = line; h.offset = ls; if (P->n < P->cap) P->hits[P->n] = h; ++P->n; }
        if (orc_sxid_match(P->buf + ls, n, &h)) { h.line = line; h.offset = ls; if (P->n < P->cap) P->hits[P->n] = h; ++P->n; }
      }
      if (P->ext && n >= 7) {
        const int32_t m = orc_ext_match(P->buf + ls, n, NULL);
        for (int k = 0; k < ORC_N_EXT; ++k)
          if (m & (1 << k)) {
            orc_hit h;
            memset(&h, 0, sizeof h);
            h.kind = 3 + k; h.line = line; h.offset = ls;   /* event type Warning is fixed by pkg/kmsg/syncer.go:94 */
            if (P->n < P->cap) P->hits[P->n] = h;
            ++P->n;
          }
      }
      ls = i + 1;
      if (i < P->e) ++line;
    }
  }
  P->lines = line;   /* newlines inside [b, e) */
}

int64_t orc_scan_lines_ext(const char* buf, int64_t len, orc_hit* hits, int64_t cap, int64_t* n_lines, int32_t threads, int32_t ext);
int64_t orc_scan_lines(const char* buf, int64_t len, orc_hit* hits, int64_t cap, int64_t* n_lines, int32_t threads) {
  return orc_scan_lines_ext(buf, len, hits, cap, n_lines, threads, 0);
}
int64_t orc_scan_lines_ext(const char* buf, int64_t len, orc_hit* hits, int64_t cap, int64_t* n_lines, int32_t threads, int32_t ext) {
  if (threads <= 0) threads = orc_max_threads();
  int parts = threads * 4;
  if (parts > 1024) parts = 1024;
  if ((int64_t)parts > len / 4096 + 1) parts = (int)(len / 4096 + 1);
  orc_scan_part* P = (orc_scan_part*)calloc((size_t)parts, sizeof(orc_scan_part));
  int64_t b = 0;
  for (int i = 0; i < parts; ++i) {
    int64_t e = (i == parts - 1) ? len : (len * (i + 1)) / parts;
    if (e < b) e = b;
    while (e < len && buf[e] != '\n') ++e;       /* cut at a newline: the part is [b, e], the newline at e belongs to it */
    if (e < len && i != parts - 1) ++e;
    P[i].buf = buf; P[i].b = b; P[i].e = e; P[i].cap = cap; P[i].ext = ext;
    P[i].hits = (orc_hit*)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(orc_hit));
    b = e;
  }
  orc_parallel_for(orc_scan_one, P, parts, threads);
  int64_t total = 0, line0 = 0;
  for (int i = 0; i < parts; ++i) {
    for (int64_t k = 0; k < P[i].n && k < P[i].cap; ++k) {
      if (total < cap) { hits[total] = P[i].hits[k]; hits[total].line += line0; }
      ++total;
    }
    if (P[i].n > P[i].cap) total += P[i].n - P[i].cap;
    line0 += P[i].lines;
    free(P[i].hits);
  }
  if (n_lines) *n_lines = line0 + 1;
  free(P);
  return total;
}
