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
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
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
int32_t orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
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
