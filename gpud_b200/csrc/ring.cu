// ring.cu — device-resident counter-sample ring and the fused windowed aggregates (kernels K1-K4 of SURVEY.md §2.2).
//
// Data layout in HBM (per GPU):
//   ring      f64 [F][CAP]      field-major: one field's samples are contiguous in time, so a window of W samples
//                               is one 8·W-byte contiguous run -> 128-bit coalesced loads, one warp per window.
//   results   f64 [F][n_windows] x {min,max,mean,ema,p99}, u32 [F][n_windows] n_over
//   part      f64 [n_windows][F] per-window EMA partial sums (transposed so the carry scan reads coalesced)
//
// Roofline: every kernel here is HBM-bound; the fused reduce reads each sample exactly once (8 B / sample,
// SURVEY.md §8d) and writes 44 B per window.  No tensor cores: there is no contraction on this path.
//
// Reference seams replaced (the reference has no windowed aggregation at all):
//   sample sink     pkg/metrics/scraper/prometheus.go:28-81 + pkg/metrics/store/sqlite.go:108-164
//   threshold `>`   components/accelerator/nvidia/temperature/component.go:228,240
#include <math.h>
#include <stdarg.h>

#include <algorithm>

#include "internal.h"

namespace {

constexpr int kMaxWindow = 1024;      // 32 lanes x 32 elements held in registers
constexpr int kWarpsPerBlock = 8;
constexpr int kCandMax = 64;
constexpr unsigned kFull = 0xffffffffu;
constexpr size_t kStageBytes = 32u << 20;   // pinned staging buffers (x2)

struct WinParams {
  const double* ring;
  int64_t cap, start, count;
  int W, F, nw;
  int q_num, q_den;
  const double* thr;   // [F]
  const double* pw;    // [127]: (1-alpha)^e for e = -63..63
  double q64;          // (1-alpha)^64
  double alpha;
  double* out_min;
  double* out_max;
  double* out_mean;
  double* out_p99;
  uint32_t* out_nover;
  double* part;        // [nw][F]
  int do_select;       // 0: skip the order statistic (range reduce uses the radix select instead)
};

// ---------------------------------------------------------------------------------------------
// K1: append.  src [n][F] row-major (one row per poll) -> ring [F][CAP] at columns (head + i) % CAP.
// 32x32 tile transpose through shared memory: reads coalesced along F, writes coalesced along time.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ring_append(const double* __restrict__ src, double* __restrict__ ring, int64_t n, int F,
                                                      int64_t cap, int64_t head) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int64_t tiles_f = (F + 31) / 32;
  const int64_t tiles_n = (n + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_f * tiles_n; t += gridDim.x) {
    const int64_t tn = t / tiles_f, tf = t - tn * tiles_f;
    const int64_t r0 = tn * 32, f0 = tf * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t r = r0 + ty + 8 * k, f = f0 + tx;
      if (r < n && f < F) tile[ty + 8 * k][tx] = __ldcs(src + r * F + f);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t f = f0 + ty + 8 * k, r = r0 + tx;
      if (r < n && f < F) {
        int64_t col = head + r;
        col = col >= cap ? col % cap : col;
        ring[f * cap + col] = tile[tx][ty + 8 * k];
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// warp primitives on 64-bit keys
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  // two REDUX passes instead of ten shuffles: max of the high words, then max of the low words among the winners
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mh = __reduce_min_sync(kFull, hi);
  const unsigned ml = __reduce_min_sync(kFull, hi == mh ? lo : 0xffffffffu);
  return ((unsigned long long)mh << 32) | ml;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// bitonic sort across the 32 lanes, descending: afterwards lane i holds the i-th largest value
__device__ __forceinline__ unsigned long long warp_sort_desc_u64(unsigned long long v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long o = __shfl_xor_sync(kFull, v, j);
      const bool keep_max = (((lane & k) == 0) == ((lane & j) == 0));
      const bool o_gt = o > v;
      v = (keep_max == o_gt) ? o : v;
    }
  }
  return v;
}

// exact k-th largest of the 32x32 register-resident keys by MSB-first bit search (always correct, slow path)
__device__ __noinline__ unsigned long long warp_select_bits(const unsigned long long (&key)[32], int kk) {
  unsigned long long pref = 0;
  for (int b = 63; b >= 0; --b) {
    const unsigned long long trial = pref | (1ull << b);
    const unsigned long long himask = ~((1ull << b) - 1ull);
    unsigned c = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) c += ((key[i] & himask) == trial) ? 1u : 0u;
    c = __reduce_add_sync(kFull, c);
    if ((int)c >= kk) pref = trial; else kk -= (int)c;
  }
  return pref;
}

// ---------------------------------------------------------------------------------------------
// K2+K4 (+ the per-window part of K3): one warp per (field, window).
//   lane l, register pair j holds chronological elements t = 64 j + 2 l + {0,1} of the window.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 2) k_window_reduce(const WinParams p) {
  __shared__ unsigned long long s_cand[kWarpsPerBlock][kCandMax];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t n_units = (int64_t)p.F * p.nw;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int64_t u = (int64_t)blockIdx.x * kWarpsPerBlock + warp; u < n_units; u += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int f = (int)(u / p.nw);
    const int w = (int)(u - (int64_t)f * p.nw);
    const int64_t c0 = (int64_t)w * p.W;
    const int m = (int)min((int64_t)p.W, p.count - c0);
    int64_t p0 = p.start + c0;
    if (p0 >= p.cap) p0 -= p.cap;
    const double* __restrict__ base = p.ring + (int64_t)f * p.cap;
    const bool fast = (p0 + m <= p.cap) && ((p0 & 1) == 0);
    const int J = (m + 63) >> 6;       // register pairs that hold at least one valid element
    const int Jfull = m >> 6;          // register pairs in which every lane's two elements are valid
    const double thr = __ldg(p.thr + f);

    // ---- issue every load of the window before touching any of them (8 KB in flight per warp) ----
    double2 v[16];
    if (fast) {
      const double2* __restrict__ b2 = reinterpret_cast<const double2*>(base + p0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int t0 = 64 * j + 2 * lane;
        v[j] = make_double2(0.0, 0.0);
        if (t0 < m) v[j] = __ldcs(b2 + (t0 >> 1));   // 16 B aligned; element t0+1 == m is masked below (ring has slack)
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int t0 = 64 * j + 2 * lane;
        v[j] = make_double2(0.0, 0.0);
        if (t0 < m) {
          int64_t a = p0 + t0;
          if (a >= p.cap) a -= p.cap;
          v[j].x = __ldcs(base + a);
        }
        if (t0 + 1 < m) {
          int64_t a = p0 + t0 + 1;
          if (a >= p.cap) a -= p.cap;
          v[j].y = __ldcs(base + a);
        }
      }
    }

    unsigned long long key[32];
    double sum = 0.0, es0 = 0.0, es1 = 0.0;
    unsigned nov = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < Jfull) {                       // warp-uniform: whole register pair valid
        const double x0 = v[j].x, x1 = v[j].y;
        const unsigned long long k0 = gpud_f64_key((unsigned long long)__double_as_longlong(x0));
        const unsigned long long k1 = gpud_f64_key((unsigned long long)__double_as_longlong(x1));
        key[2 * j] = k0;
        key[2 * j + 1] = k1;
        sum += x0 + x1;
        es0 = fma(es0, p.q64, x0);
        es1 = fma(es1, p.q64, x1);
        nov += (x0 > thr) ? 1u : 0u;
        nov += (x1 > thr) ? 1u : 0u;
        kmax = max(kmax, max(k0, k1));
        kmin = min(kmin, min(k0, k1));
      } else if (j < J) {                    // the one partially valid pair
        const int t0 = 64 * j + 2 * lane;
        const bool a0 = t0 < m, a1 = t0 + 1 < m;
        const double x0 = a0 ? v[j].x : 0.0, x1 = a1 ? v[j].y : 0.0;
        const unsigned long long k0 = a0 ? gpud_f64_key((unsigned long long)__double_as_longlong(x0)) : 0ull;
        const unsigned long long k1 = a1 ? gpud_f64_key((unsigned long long)__double_as_longlong(x1)) : 0ull;
        key[2 * j] = k0;
        key[2 * j + 1] = k1;
        sum += x0 + x1;
        es0 = fma(es0, p.q64, x0);
        es1 = fma(es1, p.q64, x1);
        nov += (a0 && x0 > thr) ? 1u : 0u;
        nov += (a1 && x1 > thr) ? 1u : 0u;
        kmax = max(kmax, max(k0, k1));
        if (a0) kmin = min(kmin, k0);
        if (a1) kmin = min(kmin, k1);
      } else {
        key[2 * j] = 0ull;
        key[2 * j + 1] = 0ull;
      }
    }

    // ---- warp-level combines ----
    const unsigned long long wmax = warp_max_u64(kmax);
    const unsigned long long wmin = warp_min_u64(kmin);
    sum = warp_sum_f64(sum);
    nov = __reduce_add_sync(kFull, nov);
    // EMA partial: sum_t alpha (1-alpha)^(m-1-t) x_t with t = 64 j + 2 lane + h; Horner above ran over j < J with q64
    const int eb = m - 1 - 64 * (J - 1) - 2 * lane;          // exponent of this lane's h=0 element in pair J-1, in [-62, 63]
    double ep = es0 * __ldg(p.pw + (eb + 63)) + es1 * __ldg(p.pw + (eb - 1 + 63));
    ep = warp_sum_f64(ep) * p.alpha;

    // ---- exact order statistic: k-th largest, k = m - ceil(m q) + 1 ----
    long long r = ((long long)m * p.q_num + p.q_den - 1) / p.q_den;
    r = r < 1 ? 1 : (r > m ? m : r);
    const int k = m - (int)r + 1;
    unsigned long long ans = 0ull;
    if (!p.do_select) {
    } else if (k <= 32) {
      // lower bound L = k-th largest lane maximum: at least k elements are >= L, so the answer is >= L
      const unsigned long long sorted = warp_sort_desc_u64(kmax, lane);
      const unsigned long long L = __shfl_sync(kFull, sorted, k - 1);
      int cnt = 0;                                            // elements strictly above L, compacted to shared memory
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool pr = key[i] > L;
        const unsigned b = __ballot_sync(kFull, pr);
        if (b) {
          if (pr) {
            const int pos = cnt + __popc(b & lt_mask);
            if (pos < kCandMax) s_cand[warp][pos] = key[i];
          }
          cnt += __popc(b);
        }
      }
      if (cnt < k) {
        ans = L;                                              // fewer than k above L and >= k at-or-above L  =>  answer is L
      } else if (cnt <= 32) {
        __syncwarp();
        const unsigned long long c = lane < cnt ? s_cand[warp][lane] : 0ull;
        const unsigned long long cs = warp_sort_desc_u64(c, lane);
        ans = __shfl_sync(kFull, cs, k - 1);
        __syncwarp();
      } else {
        ans = warp_select_bits(key, k);
      }
    } else {
      ans = warp_select_bits(key, k);
    }

    if (lane == 0) {
      const int64_t o = (int64_t)f * p.nw + w;
      p.out_min[o] = __longlong_as_double((long long)gpud_key_f64bits(wmin));
      p.out_max[o] = __longlong_as_double((long long)gpud_key_f64bits(wmax));
      p.out_mean[o] = sum / (double)m;
      p.out_p99[o] = __longlong_as_double((long long)gpud_key_f64bits(ans));
      p.out_nover[o] = nov;
      p.part[(int64_t)w * p.F + f] = ep;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3 carry: E_w = (1-alpha)^{m_w} E_{w-1} + P_w, E_{-1} = x[0] (oldest sample).  One thread per field; the
// partials are stored [nw][F] so the loads are coalesced and independent of the FMA chain.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ema_carry(const double* __restrict__ part, const double* __restrict__ ring, int64_t cap,
                                                    int64_t start, int F, int nw, double d_full, double d_last,
                                                    double* __restrict__ out_ema) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double e = ring[(int64_t)f * cap + start];
  int w = 0;
  for (; w + 8 <= nw - 1; w += 8) {
    double pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = __ldg(part + (int64_t)(w + i) * F + f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      e = fma(d_full, e, pv[i]);
      out_ema[(int64_t)f * nw + w + i] = e;
    }
  }
  for (; w < nw; ++w) {
    const double d = (w == nw - 1) ? d_last : d_full;
    e = fma(d, e, __ldg(part + (int64_t)w * F + f));
    out_ema[(int64_t)f * nw + w] = e;
  }
}

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
struct gpud_ring {
  gpud_ctx* ctx = nullptr;
  int dev = 0;
  int F = 0, W = 0;
  int64_t cap = 0;
  double alpha = 0;
  int q_num = 99, q_den = 100;
  int64_t total = 0;
  double* d_ring = nullptr;
  double* d_thr = nullptr;
  double* d_pw = nullptr;
  int64_t nw_max = 0;
  double* d_res[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t* d_nover = nullptr;
  double* d_part = nullptr;
  int64_t reduced_nw = -1;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  double* h_stage[2] = {nullptr, nullptr};
  double* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  int64_t stage_rows = 0;
  int sm_count = 148;
  double* d_rng[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // range-reduce scratch: per-window min,max,mean,ema,partials
  uint32_t* d_rng_nover = nullptr;
  size_t range_cap = 0;
  cudaEvent_t ev_k[3] = {nullptr, nullptr, nullptr};   // around the two kernels of the last reduce (bench roofline)
};

static int64_t ring_count(const gpud_ring* r) { return std::min(r->total, r->cap); }
static int64_t ring_start(const gpud_ring* r) { return r->total <= r->cap ? 0 : r->total % r->cap; }

extern "C" int32_t gpud_ring_create(gpud_ctx* ctx, int32_t dev, const gpud_ring_cfg* cfg, gpud_ring** out) {
  if (!ctx || !cfg || !out) return GPUD_E_INVALID;
  if (gpud_dev_slot(ctx, dev) < 0) return gpud_fail(ctx, GPUD_E_INVALID, "device %d is not part of this ctx", dev);
  if (cfg->n_fields < 1 || cfg->window < 1 || cfg->window > kMaxWindow || cfg->capacity < cfg->window || (cfg->capacity & 1))
    return gpud_fail(ctx, GPUD_E_INVALID, "ring cfg: need n_fields>=1, 1<=window<=%d, capacity even and >= window", kMaxWindow);
  double alpha = cfg->ema_alpha > 0 ? cfg->ema_alpha : 2.0 / (cfg->window + 1.0);
  if (!(alpha > 0.0) || alpha > 0.9999) return gpud_fail(ctx, GPUD_E_INVALID, "ema_alpha must be in (0, 0.9999]");
  int qn = cfg->q_num, qd = cfg->q_den;
  if (qn == 0 && qd == 0) { qn = 99; qd = 100; }
  if (qn < 0 || qd <= 0 || qn > qd) return gpud_fail(ctx, GPUD_E_INVALID, "quantile q_num/q_den must satisfy 0 <= num <= den");
  GPUD_CUDA(ctx, cudaSetDevice(dev));
  gpud_ring* r = new gpud_ring();
  r->ctx = ctx; r->dev = dev; r->F = cfg->n_fields; r->W = cfg->window; r->cap = cfg->capacity; r->alpha = alpha;
  r->q_num = qn; r->q_den = qd;
  r->sm_count = ctx->sm_count;
  r->nw_max = (r->cap + r->W - 1) / r->W;
  const size_t ring_bytes = (size_t)r->F * r->cap * sizeof(double) + 64;   // slack: the 16-byte load of a window's last odd element
  const size_t res_bytes = (size_t)r->F * r->nw_max * sizeof(double);
  cudaError_t e = cudaMalloc(&r->d_ring, ring_bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->d_ring, 0, ring_bytes, 0);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_thr, r->F * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&r->d_pw, 127 * sizeof(double));
  for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaMalloc(&r->d_res[i], res_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_nover, (size_t)r->F * r->nw_max * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&r->d_part, res_bytes);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking);
  r->stage_rows = std::max<int64_t>(1, (int64_t)(kStageBytes / (sizeof(double) * r->F)));
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaMallocHost(&r->h_stage[i], (size_t)r->stage_rows * r->F * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&r->d_stage[i], (size_t)r->stage_rows * r->F * sizeof(double));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_stage[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&r->ev_k[i]);
  if (e != cudaSuccess) {
    int32_t rc = gpud_fail(ctx, e == cudaErrorMemoryAllocation ? GPUD_E_NOMEM : GPUD_E_CUDA, "ring_create: %s", cudaGetErrorString(e));
    gpud_ring_destroy(r);
    return rc;
  }
  r->stream = r->own_stream;
  std::vector<double> thr(r->F, INFINITY), pw(127);
  if (cfg->thresholds) std::copy(cfg->thresholds, cfg->thresholds + r->F, thr.begin());
  for (int ex = -63; ex <= 63; ++ex) pw[ex + 63] = pow(1.0 - alpha, (double)ex);
  GPUD_CUDA(ctx, cudaMemcpy(r->d_thr, thr.data(), r->F * sizeof(double), cudaMemcpyHostToDevice));
  GPUD_CUDA(ctx, cudaMemcpy(r->d_pw, pw.data(), 127 * sizeof(double), cudaMemcpyHostToDevice));
  GPUD_CUDA(ctx, cudaDeviceSynchronize());
  *out = r;
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_destroy(gpud_ring* r) {
  if (!r) return GPUD_E_INVALID;
  cudaSetDevice(r->dev);
  if (r->own_stream) cudaStreamSynchronize(r->own_stream);
  cudaFree(r->d_ring); cudaFree(r->d_thr); cudaFree(r->d_pw); cudaFree(r->d_nover); cudaFree(r->d_part);
  for (auto& p : r->d_res) cudaFree(p);
  for (auto& p : r->d_rng) cudaFree(p);
  cudaFree(r->d_rng_nover);
  for (int i = 0; i < 2; ++i) {
    if (r->h_stage[i]) cudaFreeHost(r->h_stage[i]);
    cudaFree(r->d_stage[i]);
    if (r->ev_stage[i]) cudaEventDestroy(r->ev_stage[i]);
  }
  for (auto& ev : r->ev_k) if (ev) cudaEventDestroy(ev);
  if (r->own_stream) cudaStreamDestroy(r->own_stream);
  delete r;
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_set_stream(gpud_ring* r, void* s) {
  if (!r) return GPUD_E_INVALID;
  r->stream = s ? (cudaStream_t)s : r->own_stream;
  return GPUD_OK;
}

static int32_t launch_append(gpud_ring* r, const double* d_rows, int64_t n) {
  // only the newest CAP rows can survive
  if (n > r->cap) {
    d_rows += (n - r->cap) * r->F;
    r->total += n - r->cap;
    n = r->cap;
  }
  const int64_t head = r->total % r->cap;
  const int64_t tiles = ((r->F + 31) / 32) * ((n + 31) / 32);
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)r->sm_count * 16);
  k_ring_append<<<grid, 256, 0, r->stream>>>(d_rows, r->d_ring, n, r->F, r->cap, head);
  GPUD_CUDA(r->ctx, cudaGetLastError());
  r->total += n;
  r->reduced_nw = -1;
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_push_device(gpud_ring* r, const double* dev_rows, int64_t n) {
  if (!r || (!dev_rows && n) || n < 0) return GPUD_E_INVALID;
  if (n == 0) return GPUD_OK;
  GPUD_CUDA(r->ctx, cudaSetDevice(r->dev));
  return launch_append(r, dev_rows, n);
}

extern "C" int32_t gpud_ring_push(gpud_ring* r, const double* host_rows, int64_t n) {
  if (!r || (!host_rows && n) || n < 0) return GPUD_E_INVALID;
  if (n == 0) return GPUD_OK;
  GPUD_CUDA(r->ctx, cudaSetDevice(r->dev));
  if (n > r->cap) {                      // rows that would be overwritten immediately never cross PCIe
    host_rows += (n - r->cap) * r->F;
    r->total += n - r->cap;
    n = r->cap;
  }
  cudaPointerAttributes attr;
  bool pinned = false;
  if (cudaPointerGetAttributes(&attr, host_rows) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
  else cudaGetLastError();
  int64_t done = 0;
  int buf = 0;
  while (done < n) {
    const int64_t rows = std::min(r->stage_rows, n - done);
    const size_t bytes = (size_t)rows * r->F * sizeof(double);
    GPUD_CUDA(r->ctx, cudaEventSynchronize(r->ev_stage[buf]));   // previous use of this staging pair has drained
    const double* src = host_rows + done * r->F;
    if (!pinned) {
      memcpy(r->h_stage[buf], src, bytes);                     // pageable caller memory: stage through pinned
      src = r->h_stage[buf];
    }
    GPUD_CUDA(r->ctx, cudaMemcpyAsync(r->d_stage[buf], src, bytes, cudaMemcpyHostToDevice, r->stream));
    int32_t rc = launch_append(r, r->d_stage[buf], rows);
    if (rc) return rc;
    GPUD_CUDA(r->ctx, cudaEventRecord(r->ev_stage[buf], r->stream));
    done += rows;
    buf ^= 1;
  }
  if (pinned) GPUD_CUDA(r->ctx, cudaStreamSynchronize(r->stream));   // caller may reuse its buffer on return
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_counts(gpud_ring* r, int64_t* total, int64_t* count, int64_t* n_windows) {
  if (!r) return GPUD_E_INVALID;
  const int64_t c = ring_count(r);
  if (total) *total = r->total;
  if (count) *count = c;
  if (n_windows) *n_windows = (c + r->W - 1) / r->W;
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_reduce(gpud_ring* r) {
  if (!r) return GPUD_E_INVALID;
  const int64_t count = ring_count(r);
  if (count == 0) return gpud_fail(r->ctx, GPUD_E_STATE, "ring is empty");
  GPUD_CUDA(r->ctx, cudaSetDevice(r->dev));
  WinParams p;
  p.ring = r->d_ring; p.cap = r->cap; p.start = ring_start(r); p.count = count;
  p.W = r->W; p.F = r->F; p.nw = (int)((count + r->W - 1) / r->W);
  p.q_num = r->q_num; p.q_den = r->q_den; p.thr = r->d_thr; p.pw = r->d_pw;
  p.q64 = pow(1.0 - r->alpha, 64.0); p.alpha = r->alpha;
  p.out_min = r->d_res[GPUD_OP_MIN]; p.out_max = r->d_res[GPUD_OP_MAX]; p.out_mean = r->d_res[GPUD_OP_MEAN];
  p.out_p99 = r->d_res[GPUD_OP_P99]; p.out_nover = r->d_nover; p.part = r->d_part; p.do_select = 1;
  const int64_t units = (int64_t)p.F * p.nw;
  // persistent grid: 2 CTAs of 8 warps per SM, a whole number of waves (148 SMs)
  const int grid = (int)std::min<int64_t>((units + kWarpsPerBlock - 1) / kWarpsPerBlock, (int64_t)r->sm_count * 2);
  cudaEventRecord(r->ev_k[0], r->stream);
  k_window_reduce<<<grid, kWarpsPerBlock * 32, 0, r->stream>>>(p);
  GPUD_CUDA(r->ctx, cudaGetLastError());
  cudaEventRecord(r->ev_k[1], r->stream);
  const int m_last = (int)(count - (int64_t)(p.nw - 1) * r->W);
  k_ema_carry<<<(r->F + 127) / 128, 128, 0, r->stream>>>(r->d_part, r->d_ring, r->cap, p.start, r->F, p.nw,
                                                        pow(1.0 - r->alpha, (double)r->W), pow(1.0 - r->alpha, (double)m_last),
                                                        r->d_res[GPUD_OP_EMA]);
  GPUD_CUDA(r->ctx, cudaGetLastError());
  cudaEventRecord(r->ev_k[2], r->stream);
  r->reduced_nw = p.nw;
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_kernel_ms(gpud_ring* r, float* reduce_ms, float* carry_ms) {
  if (!r || !reduce_ms || !carry_ms) return GPUD_E_INVALID;
  if (r->reduced_nw < 0) return gpud_fail(r->ctx, GPUD_E_STATE, "no reduce since the last push");
  GPUD_CUDA(r->ctx, cudaSetDevice(r->dev));
  GPUD_CUDA(r->ctx, cudaEventSynchronize(r->ev_k[2]));
  GPUD_CUDA(r->ctx, cudaEventElapsedTime(reduce_ms, r->ev_k[0], r->ev_k[1]));
  GPUD_CUDA(r->ctx, cudaEventElapsedTime(carry_ms, r->ev_k[1], r->ev_k[2]));
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_sync(gpud_ring* r) {
  if (!r) return GPUD_E_INVALID;
  GPUD_CUDA(r->ctx, cudaSetDevice(r->dev));
  GPUD_CUDA(r->ctx, cudaStreamSynchronize(r->stream));
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_result_ptr(gpud_ring* r, int32_t op, void** dev_ptr) {
  if (!r || !dev_ptr || op < 0 || op >= GPUD_N_OPS) return GPUD_E_INVALID;
  if (r->reduced_nw < 0) return gpud_fail(r->ctx, GPUD_E_STATE, "no reduce since the last push");
  *dev_ptr = op == GPUD_OP_NOVER ? (void*)r->d_nover : (void*)r->d_res[op];
  return GPUD_OK;
}

extern "C" int32_t gpud_ring_read(gpud_ring* r, int32_t op, void* out, int64_t out_bytes) {
  if (!r || !out || op < 0 || op >= GPUD_N_OPS) return GPUD_E_INVALID;
  if (r->reduced_nw < 0) return gpud_fail(r->ctx, GPUD_E_STATE, "no reduce since the last push");
  const size_t el = op == GPUD_OP_NOVER ? sizeof(uint32_t) : sizeof(double);
  const size_t need = (size_t)r->F * r->reduced_nw * el;
  if ((size_t)out_bytes < need) return gpud_fail(r->ctx, GPUD_E_CAPACITY, "need %zu bytes", need);
  GPUD_CUDA(r->ctx, cudaSetDevice(r->dev));
  const void* src = op == GPUD_OP_NOVER ? (const void*)r->d_nover : (const void*)r->d_res[op];
  GPUD_CUDA(r->ctx, cudaMemcpyAsync(out, src, need, cudaMemcpyDeviceToHost, r->stream));
  GPUD_CUDA(r->ctx, cudaStreamSynchronize(r->stream));
  return GPUD_OK;
}

// ---- range reduce support (select.cu): run the window kernel with W' = 1024 over the last `n` samples into scratch ----
int32_t gpud_ring_range_partials(gpud_ring* r, int64_t n, gpud_range_view* v) {
  const int64_t count = ring_count(r);
  if (n <= 0 || n > count) n = count;
  const int Wp = (int)std::min<int64_t>(kMaxWindow, n);
  const int nw = (int)((n + Wp - 1) / Wp);
  const size_t per = (size_t)r->F * nw;
  if (per > r->range_cap) {
    for (auto& q : r->d_rng) { cudaFree(q); q = nullptr; }
    cudaFree(r->d_rng_nover); r->d_rng_nover = nullptr;
    r->range_cap = 0;
    for (auto& q : r->d_rng) GPUD_CUDA(r->ctx, cudaMalloc(&q, per * sizeof(double)));
    GPUD_CUDA(r->ctx, cudaMalloc(&r->d_rng_nover, per * sizeof(uint32_t)));
    r->range_cap = per;
  }
  int64_t start = ring_start(r) + (count - n);
  if (start >= r->cap) start -= r->cap;
  WinParams p;
  p.ring = r->d_ring; p.cap = r->cap; p.start = start; p.count = n; p.W = Wp; p.F = r->F; p.nw = nw;
  p.q_num = r->q_num; p.q_den = r->q_den; p.thr = r->d_thr; p.pw = r->d_pw; p.q64 = pow(1.0 - r->alpha, 64.0); p.alpha = r->alpha;
  p.out_min = r->d_rng[0]; p.out_max = r->d_rng[1]; p.out_mean = r->d_rng[2]; p.out_p99 = r->d_rng[3]; p.out_nover = r->d_rng_nover;
  p.part = r->d_rng[4]; p.do_select = 0;
  const int64_t units = (int64_t)p.F * p.nw;
  const int grid = (int)std::min<int64_t>((units + kWarpsPerBlock - 1) / kWarpsPerBlock, (int64_t)r->sm_count * 2);
  k_window_reduce<<<grid, kWarpsPerBlock * 32, 0, r->stream>>>(p);
  GPUD_CUDA(r->ctx, cudaGetLastError());
  const int m_last = (int)(n - (int64_t)(nw - 1) * Wp);
  k_ema_carry<<<(r->F + 127) / 128, 128, 0, r->stream>>>(r->d_rng[4], r->d_ring, r->cap, start, r->F, nw, pow(1.0 - r->alpha, (double)Wp),
                                                        pow(1.0 - r->alpha, (double)m_last), r->d_rng[3]);
  GPUD_CUDA(r->ctx, cudaGetLastError());
  v->ring = r->d_ring; v->F = r->F; v->cap = r->cap; v->start = start; v->n = n; v->Wp = Wp; v->nw = nw;
  v->w_min = r->d_rng[0]; v->w_max = r->d_rng[1]; v->w_mean = r->d_rng[2]; v->w_ema = r->d_rng[3]; v->w_nover = r->d_rng_nover;
  v->q_num = r->q_num; v->q_den = r->q_den; v->stream = r->stream; v->ctx = r->ctx; v->dev = r->dev; v->sm_count = r->sm_count;
  return GPUD_OK;
}
