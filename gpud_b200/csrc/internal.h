// internal.h — shared by the translation units of libgpud_b200.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/gpud_b200.h"

struct gpud_scan_state;   // kmsg_scan.cu
struct gpud_comm_state;   // fabric.cu

struct gpud_ctx {
  std::vector<int> devs;
  std::mutex mu;
  std::string last_error;
  std::vector<gpud_scan_state*> scan;   // per dev slot, lazily created
  std::vector<gpud_comm_state*> comm;   // per dev slot
  std::vector<void*> fabric_tables;     // per dev slot: device table [GPUD_MAX_GPUS] x 128 B (+flags) for p2p gather
  int sm_count = 148;
  std::mutex fabric_mu;                 // gpud_fabric_gather_p2p: one gather at a time per ctx
  unsigned fabric_epoch = 0;            // gathers completed or begun on this ctx's tables (arrival counters expect epoch * n)
  bool fabric_dirty = false;            // the last gather did not finish: counters are re-zeroed before the next one
};

inline int32_t gpud_fail(gpud_ctx* ctx, int32_t code, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
inline int32_t gpud_fail(gpud_ctx* ctx, int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) {
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->last_error = buf;
  }
  return code;
}

#define GPUD_CUDA(ctx, expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return gpud_fail((ctx), GPUD_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

inline int gpud_dev_slot(gpud_ctx* ctx, int dev) {
  for (size_t i = 0; i < ctx->devs.size(); ++i)
    if (ctx->devs[i] == dev) return (int)i;
  return -1;
}

// ---- device helpers --------------------------------------------------------------------------
#ifdef __CUDACC__
// IEEE-754 totalOrder key (oracle/SPEC.md): unsigned compare of the key == totalOrder of the doubles.
__host__ __device__ __forceinline__ unsigned long long gpud_f64_key(unsigned long long bits) {
  // branch-free: m = all-ones for negatives (arithmetic shift of the sign), then flip everything / only the sign bit
  const unsigned long long m = (unsigned long long)((long long)bits >> 63);
  return bits ^ (m | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ unsigned long long gpud_key_f64bits(unsigned long long key) {
  return key ^ ((key >> 63) ? 0x8000000000000000ull : ~0ull);
}
#endif

// view of a ring range around the per-window pass (ring.cu <-> select.cu)
#define GPUD_RANGE_SAMPLED_MIN 65536   /* ranges at least this long take the sampled-pivot single pass */
#define GPUD_RANGE_LIST_CAP 65536      /* keys between the pivots kept per field */
#define GPUD_RANGE_SAMPLE_MAX 8192     /* keys per field in the sample the pivots come from */
struct gpud_range_view {
  const double* ring; int F; int64_t cap, start, n; int Wp, nw;
  const double *w_min, *w_max, *w_mean, *w_ema; const uint32_t* w_nover;
  int q_num, q_den; cudaStream_t stream; gpud_ctx* ctx; int dev, sm_count;
  int sampled;                         // 1: the pass classifies against `piv` and fills `lists`
  unsigned list_cap;
  double* piv;                         // [F][2] lo, hi
  unsigned* fill;                      // [F]
  uint4* w_cls;                        // [F][nw] {above hi, == hi, == lo, strictly inside}
  unsigned long long* lists;           // [F][list_cap]
  const double* sample;                // the ring's own sample ring [F][smp_slots] (ring.cu k_ring_append)
  int smp_shift; int64_t smp_slots;
  cudaEvent_t ev[3];
};
int32_t gpud_ring_range_prepare(gpud_ring* r, int64_t n, gpud_range_view* v);
int32_t gpud_ring_range_pass(gpud_ring* r, const gpud_range_view* v);
void gpud_ring_range_note(gpud_ring* r, bool sampled, unsigned fields_open, const int* reasons /*[GPUD_RANGE_N_OPEN_REASONS]*/);
void gpud_ring_quantile(gpud_ring* r, int* q_num, int* q_den);

void gpud_scan_state_free(gpud_scan_state*);
void gpud_comm_state_free(gpud_comm_state*);
void gpud_parallel_memcpy(void* dst, const void* src, size_t n);

// ring.cu: the ring's field count (the pollers check their row width against it before pushing)
int gpud_ring_n_fields(const gpud_ring* r);
