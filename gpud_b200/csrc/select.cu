// select.cu — whole-range aggregates of the newest n samples of every field (gpud_ring_reduce_range).
//
// min / max / mean / EMA / n_over are folded from the per-window pass of ring.cu (W' = 1024 windows over the range).
// The exact order statistic is an MSB-first radix select on the IEEE totalOrder keys that spends HBM passes only where
// they narrow the candidate set:
//   1. the bits on which the field's min and max keys agree are skipped (they are common to every key);
//   2. one 11-bit histogram pass over the first varying digit usually leaves <= kCollectMax keys in the bin that holds the
//      rank; further histogram passes run only for fields that still have more (heavy ties);
//   3. one collect pass compacts the surviving keys of every field, and a final per-field kernel selects among them
//      bit by bit in registers.
// Typical cost: 3 reads of the range (window pass, histogram, collect) instead of the 7 of a plain 6-digit radix select.
// Definitions: oracle/SPEC.md (parity unpinned: the reference has no such aggregate).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>

#include "internal.h"

namespace {

constexpr int kBins = 2048;
constexpr int kDigit = 11;
constexpr int kCollectMax = 8192;      // candidates the final kernel keeps in registers: 256 threads x 32 keys
constexpr int kMaxHistPasses = 6;      // 64 bits / 11

struct SelState {                      // one per field, device resident
  unsigned long long prefix;           // the key bits fixed so far, right-aligned
  unsigned long long kk;               // rank (1-based, from the top) among the keys that match the prefix
  unsigned long long cnt;              // how many keys match the prefix
  unsigned long long ans;              // valid when done
  int nbits;                           // number of fixed bits, from the MSB
  int done;
  unsigned collected;                  // fill counter of the candidate list
  int pad;
};

__device__ __forceinline__ bool key_matches(unsigned long long key, const SelState& s) {
  return s.nbits == 0 || (key >> (64 - s.nbits)) == s.prefix;
}

// Stream the keys of chronological positions [b, e) of one field through `fn(key)`.  When the slice is one aligned,
// non-wrapping run every thread keeps four 128-bit loads in flight (8 keys); otherwise 64-bit loads with the ring wrap.
template <typename Fn>
__device__ __forceinline__ void for_each_key(const double* __restrict__ base, int64_t cap, int64_t start, int64_t b, int64_t e, Fn fn) {
  int64_t p0 = start + b;
  if (p0 >= cap) p0 -= cap;
  const int64_t len = e - b;
  auto K = [](double x) { return gpud_f64_key((unsigned long long)__double_as_longlong(x)); };
  if (len > 0 && p0 + len <= cap && (p0 & 1) == 0) {
    const double2* __restrict__ v2 = reinterpret_cast<const double2*>(base + p0);
    const int64_t n2 = len >> 1;                            // whole pairs
    int64_t i = threadIdx.x;
    for (; i + 3 * (int64_t)blockDim.x < n2; i += 4 * (int64_t)blockDim.x) {
      const double2 a0 = __ldcs(v2 + i), a1 = __ldcs(v2 + i + blockDim.x), a2 = __ldcs(v2 + i + 2 * blockDim.x), a3 = __ldcs(v2 + i + 3 * blockDim.x);
      fn(K(a0.x)); fn(K(a0.y)); fn(K(a1.x)); fn(K(a1.y)); fn(K(a2.x)); fn(K(a2.y)); fn(K(a3.x)); fn(K(a3.y));
    }
    for (; i < n2; i += blockDim.x) { const double2 a0 = __ldcs(v2 + i); fn(K(a0.x)); fn(K(a0.y)); }
    if ((len & 1) && threadIdx.x == 0) fn(K(__ldcs(base + p0 + len - 1)));
    return;
  }
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    int64_t a = start + i;
    if (a >= cap) a -= cap;
    fn(K(__ldcs(base + a)));
  }
}

// fold the per-window partials into per-field results; initialise the select state
__global__ void k_range_fold(int F, int nw, int Wp, int64_t n, const double* __restrict__ w_min, const double* __restrict__ w_max,
                             const double* __restrict__ w_mean, const double* __restrict__ w_ema, const uint32_t* __restrict__ w_nover,
                             int q_num, int q_den, double* __restrict__ out /*[5][F]*/, uint32_t* __restrict__ out_nover, SelState* __restrict__ st) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  unsigned long long kmin = ~0ull, kmax = 0ull;
  double sum = 0.0;
  unsigned nov = 0;
  for (int w = 0; w < nw; ++w) {
    const int64_t o = (int64_t)f * nw + w;
    const int m = (int)min((int64_t)Wp, n - (int64_t)w * Wp);
    kmin = min(kmin, gpud_f64_key((unsigned long long)__double_as_longlong(w_min[o])));
    kmax = max(kmax, gpud_f64_key((unsigned long long)__double_as_longlong(w_max[o])));
    sum += w_mean[o] * (double)m;
    nov += w_nover[o];
  }
  out[0 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(kmin));
  out[1 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(kmax));
  out[2 * F + f] = sum / (double)n;
  out[3 * F + f] = w_ema[(int64_t)f * nw + nw - 1];
  out_nover[f] = nov;
  long long r = (n * q_num + q_den - 1) / q_den;
  r = r < 1 ? 1 : (r > n ? n : r);
  SelState s;
  s.kk = (unsigned long long)(n - r + 1);
  s.cnt = (unsigned long long)n;
  s.collected = 0;
  s.pad = 0;
  const unsigned long long diff = kmin ^ kmax;
  if (diff == 0ull) {                  // a constant field: every key is the answer
    s.prefix = kmax; s.nbits = 64; s.done = 1; s.ans = kmax;
  } else {
    const int cp = __clzll((long long)diff);   // leading bits common to all keys of the field
    s.nbits = cp;
    s.prefix = cp ? (kmax >> (64 - cp)) : 0ull;
    s.done = 0;
    s.ans = 0ull;
  }
  st[f] = s;
}

// grid (blocks_per_field, F): histogram of the next digit over the keys that match the field's prefix
__global__ void __launch_bounds__(256) k_sel_hist(const double* __restrict__ ring, int64_t cap, int64_t start, int64_t n,
                                                   const SelState* __restrict__ st, unsigned* __restrict__ hist) {
  __shared__ unsigned s_hist[kBins];
  const int f = blockIdx.y;
  const SelState s = st[f];
  if (s.done || s.cnt <= (unsigned long long)kCollectMax) return;   // block-uniform: nothing left to narrow for this field
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int d = min(kDigit, 64 - s.nbits), shift = 64 - s.nbits - d;
  const unsigned mask = (1u << d) - 1u;
  const int64_t per = (((n + gridDim.x - 1) / gridDim.x) + 1) & ~(int64_t)1;   // even, so aligned slices stay aligned
  const int64_t b = min(n, (int64_t)blockIdx.x * per), e = min(n, b + per);
  const double* __restrict__ base = ring + (int64_t)f * cap;
  for_each_key(base, cap, start, b, e, [&](unsigned long long key) {
    if (key_matches(key, s)) atomicAdd(&s_hist[(unsigned)(key >> shift) & mask], 1u);
  });
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[(int64_t)f * kBins + i], s_hist[i]);
}

// one block per field: walk the digit histogram from the top, fix the next digit of the k-th largest key
__global__ void __launch_bounds__(256) k_sel_pick(unsigned* __restrict__ hist, SelState* __restrict__ st) {
  __shared__ unsigned s[kBins];
  __shared__ unsigned s_sum[256];
  const int f = blockIdx.x, t = threadIdx.x;
  SelState cur = st[f];
  if (cur.done || cur.cnt <= (unsigned long long)kCollectMax) return;
  const int d = min(kDigit, 64 - cur.nbits), nb = 1 << d;
  for (int i = t; i < kBins; i += 256) { s[i] = i < nb ? hist[(int64_t)f * kBins + i] : 0; hist[(int64_t)f * kBins + i] = 0; }
  __syncthreads();
  unsigned loc = 0;                    // thread t owns bins [8t, 8t+8) counted from the TOP
  for (int j = 8 * t; j < 8 * t + 8; ++j) if (j < nb) loc += s[nb - 1 - j];
  s_sum[t] = loc;
  __syncthreads();
  if (t == 0) {
    unsigned long long acc = 0;
    int g = 0;
    while (g < 255 && acc + s_sum[g] < cur.kk) { acc += s_sum[g]; ++g; }
    int j = 8 * g;
    while (j < nb - 1 && acc + s[nb - 1 - j] < cur.kk) { acc += s[nb - 1 - j]; ++j; }
    const unsigned bin = (unsigned)(nb - 1 - j);
    cur.prefix = (cur.nbits ? (cur.prefix << d) : 0ull) | bin;
    cur.nbits += d;
    cur.kk -= acc;
    cur.cnt = s[bin];
    if (cur.nbits >= 64) { cur.done = 1; cur.ans = cur.prefix; }
    st[f] = cur;
  }
}

// grid (blocks_per_field, F): compact the keys that still match into the field's candidate list (<= kCollectMax of them)
__global__ void __launch_bounds__(256) k_sel_collect(const double* __restrict__ ring, int64_t cap, int64_t start, int64_t n,
                                                      SelState* __restrict__ st, unsigned long long* __restrict__ lists) {
  const int f = blockIdx.y;
  const SelState s = st[f];
  if (s.done) return;
  const int64_t per = (((n + gridDim.x - 1) / gridDim.x) + 1) & ~(int64_t)1;
  const int64_t b = min(n, (int64_t)blockIdx.x * per), e = min(n, b + per);
  const double* __restrict__ base = ring + (int64_t)f * cap;
  unsigned long long* __restrict__ list = lists + (int64_t)f * kCollectMax;
  unsigned* fill = &st[f].collected;
  for_each_key(base, cap, start, b, e, [&](unsigned long long key) {
    if (key_matches(key, s)) {
      const unsigned pos = atomicAdd(fill, 1u);
      if (pos < (unsigned)kCollectMax) list[pos] = key;
    }
  });
}

// one block per field: k-th largest of the (<= kCollectMax) collected keys, remaining bits fixed one at a time;
// every thread keeps its 32 keys in registers, the per-bit counts meet in shared memory
__global__ void __launch_bounds__(256) k_sel_final(const SelState* __restrict__ st, const unsigned long long* __restrict__ lists, int F,
                                                    double* __restrict__ out) {
  __shared__ unsigned s_cnt[8];
  const int f = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const SelState s = st[f];
  unsigned long long ans = s.ans;
  if (!s.done) {
    const unsigned n = (unsigned)min((unsigned long long)kCollectMax, (unsigned long long)s.collected);
    const unsigned long long* __restrict__ list = lists + (int64_t)f * kCollectMax;
    unsigned long long key[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { const unsigned idx = (unsigned)i * 256u + (unsigned)t; key[i] = idx < n ? list[idx] : 0ull; }   // 0 never matches a set bit
    unsigned long long pref = s.nbits ? (s.prefix << (64 - s.nbits)) : 0ull;
    unsigned long long kk = s.kk;
    for (int b = 63 - s.nbits; b >= 0; --b) {
      const unsigned long long trial = pref | (1ull << b);
      const unsigned long long himask = ~((1ull << b) - 1ull);
      unsigned c = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) c += ((key[i] & himask) == trial) ? 1u : 0u;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) s_cnt[wid] = c;
      __syncthreads();
      unsigned tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += s_cnt[w];
      __syncthreads();
      if ((unsigned long long)tot >= kk) pref = trial; else kk -= tot;
    }
    ans = pref;
  }
  if (t == 0) out[4 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(ans));
}

// ---- experimental bounded select (off unless GPUD_RANGE_V2=1) ---------------------------------------------------------
// When the range is a whole number of windows, the window pass can report every window's m-th largest key, m = ceil(k / nw) for the
// global rank k.  Pigeonhole: every window holds >= m keys >= its own statistic, so >= nw m >= k keys are >= t = min_w stat_w; and no
// window holds more than m - 1 keys above its statistic, so <= nw (m - 1) <= k - 1 keys are > T = max_w stat_w.  The k-th largest key
// therefore lies in [t, T], and it is the (k - #{keys > T})-th largest of the keys inside the interval: one more pass over the range
// counts the former and collects the latter (tests/test_range_bounds_model.py checks the argument on random data).  Fields whose
// interval holds more than kCollect2Max keys are handed to the histogram path above.
constexpr int kCollect2Max = 32768;

struct Sel2State {                     // one per field
  unsigned long long t, T;             // interval bounds as totalOrder keys
  unsigned long long above;            // keys > T
  unsigned collected;                  // keys in [t, T] (may exceed kCollect2Max: overflow)
  int pad;
};

// one thread per field: interval bounds from the window statistics
__global__ void k_sel2_bounds(int F, int nw, const double* __restrict__ w_stat, Sel2State* __restrict__ s2) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  unsigned long long t = ~0ull, T = 0ull;
  for (int w = 0; w < nw; ++w) {
    const unsigned long long k = gpud_f64_key((unsigned long long)__double_as_longlong(w_stat[(int64_t)f * nw + w]));
    t = min(t, k);
    T = max(T, k);
  }
  Sel2State s;
  s.t = t; s.T = T; s.above = 0ull; s.collected = 0u; s.pad = 0;
  s2[f] = s;
}

// grid (blocks_per_field, F): count the keys above the interval, collect the keys inside it
__global__ void __launch_bounds__(256) k_sel2_collect(const double* __restrict__ ring, int64_t cap, int64_t start, int64_t n, const SelState* __restrict__ st,
                                                       Sel2State* __restrict__ s2, unsigned long long* __restrict__ lists) {
  const int f = blockIdx.y;
  if (st[f].done) return;                                   // constant field: answered by the fold
  const unsigned long long t = s2[f].t, T = s2[f].T;
  const int64_t per = (((n + gridDim.x - 1) / gridDim.x) + 1) & ~(int64_t)1;
  const int64_t b = min(n, (int64_t)blockIdx.x * per), e = min(n, b + per);
  const double* __restrict__ base = ring + (int64_t)f * cap;
  unsigned long long* __restrict__ list = lists + (int64_t)f * kCollect2Max;
  unsigned* fill = &s2[f].collected;
  unsigned above = 0;
  for_each_key(base, cap, start, b, e, [&](unsigned long long key) {
    if (key > T) ++above;
    else if (key >= t) {
      const unsigned pos = atomicAdd(fill, 1u);
      if (pos < (unsigned)kCollect2Max) list[pos] = key;
    }
  });
  above = __reduce_add_sync(0xffffffffu, above);
  if ((threadIdx.x & 31) == 0 && above) atomicAdd(&s2[f].above, (unsigned long long)above);
}

// one block per field: radix select (11-bit digits, bits common to t and T skipped) over the collected list; fields that
// overflowed the list keep done = 0 and go through the histogram path
__global__ void __launch_bounds__(256) k_sel2_final(SelState* __restrict__ st, const Sel2State* __restrict__ s2, const unsigned long long* __restrict__ lists) {
  __shared__ unsigned s_hist[kBins];
  __shared__ unsigned long long s_prefix, s_kk;
  __shared__ int s_nbits;
  const int f = blockIdx.x, t = threadIdx.x;
  if (st[f].done) return;
  const Sel2State z = s2[f];
  if (z.collected > (unsigned)kCollect2Max) return;         // overflow: left to the histogram path (st[f] untouched)
  const unsigned cnt = z.collected;
  const unsigned long long* __restrict__ list = lists + (int64_t)f * kCollect2Max;
  if (t == 0) {
    const unsigned long long diff = z.t ^ z.T;
    const int cp = diff ? __clzll((long long)diff) : 64;
    s_nbits = cp;
    s_prefix = cp ? (cp == 64 ? z.T : (z.T >> (64 - cp))) : 0ull;
    s_kk = st[f].kk - z.above;                              // rank inside the interval, 1-based from the top
  }
  __syncthreads();
  while (s_nbits < 64) {
    const int nbits = s_nbits;
    const unsigned long long prefix = s_prefix;
    const int d = min(kDigit, 64 - nbits), shift = 64 - nbits - d, nb = 1 << d;
    for (int i = t; i < kBins; i += 256) s_hist[i] = 0;
    __syncthreads();
    for (unsigned i = t; i < cnt; i += 256) {
      const unsigned long long key = list[i];
      if (nbits == 0 || (key >> (64 - nbits)) == prefix) atomicAdd(&s_hist[(unsigned)(key >> shift) & (unsigned)(nb - 1)], 1u);
    }
    __syncthreads();
    if (t == 0) {
      unsigned long long acc = 0, kk = s_kk;
      int bin = nb - 1;
      while (bin > 0 && acc + s_hist[bin] < kk) { acc += s_hist[bin]; --bin; }
      s_prefix = (nbits ? (prefix << d) : 0ull) | (unsigned long long)bin;
      s_nbits = nbits + d;
      s_kk = kk - acc;
    }
    __syncthreads();
  }
  if (t == 0) {
    SelState s = st[f];
    s.done = 1;
    s.ans = s_prefix;
    s.nbits = 64;
    s.prefix = s_prefix;
    st[f] = s;
  }
}

// host side of the experimental path: 0 = not applicable (caller runs the regular path), 1 = launched
static int range_v2_applicable(int64_t n, int q_num, int q_den, int* rank_m) {
  const char* env = getenv("GPUD_RANGE_V2");
  if (!env || env[0] != '1') return 0;
  const int64_t Wp = std::min<int64_t>(1024, n);
  if (n <= kCollectMax || Wp < 1 || (n % Wp) != 0) return 0;
  const int64_t nw = n / Wp;
  long long r = (long long)((n * q_num + q_den - 1) / q_den);
  r = r < 1 ? 1 : (r > n ? n : r);
  const int64_t k = n - r + 1;
  const int64_t m = (k + nw - 1) / nw;
  if (m < 1 || m > Wp) return 0;
  *rank_m = (int)m;
  return 1;
}

}  // namespace

extern "C" int32_t gpud_ring_reduce_range(gpud_ring* ring, int64_t last_n, double* out_f64, uint32_t* out_n_over) {
  if (!ring || !out_f64 || !out_n_over || last_n < 0) return GPUD_E_INVALID;
  gpud_range_view v;
  int64_t total = 0, count = 0, nwin = 0;
  gpud_ring_counts(ring, &total, &count, &nwin);
  if (count == 0) return GPUD_E_STATE;
  int32_t rc;
  const double* w_stat = nullptr;
  int rank_m = 0;
  {
    const int64_t n_eff = (last_n <= 0 || last_n > count) ? count : last_n;
    int qn = 99, qd = 100;
    gpud_ring_quantile(ring, &qn, &qd);
    if (range_v2_applicable(n_eff, qn, qd, &rank_m)) rc = gpud_ring_range_partials_ranked(ring, last_n, rank_m, &v, &w_stat);
    else rc = gpud_ring_range_partials(ring, last_n, &v);
  }
  if (rc) return rc;
  gpud_ctx* ctx = v.ctx;
  GPUD_CUDA(ctx, cudaSetDevice(v.dev));
  double* d_out = nullptr;
  uint32_t* d_nover = nullptr;
  unsigned* d_hist = nullptr;
  SelState* d_st = nullptr;
  unsigned long long* d_lists = nullptr;
  Sel2State* d_s2 = nullptr;
  unsigned long long* d_lists2 = nullptr;
  cudaError_t e = cudaMallocAsync(&d_out, 5 * v.F * sizeof(double), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nover, v.F * sizeof(uint32_t), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_hist, (size_t)v.F * kBins * sizeof(unsigned), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_st, v.F * sizeof(SelState), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_lists, (size_t)v.F * kCollectMax * sizeof(unsigned long long), v.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_hist, 0, (size_t)v.F * kBins * sizeof(unsigned), v.stream);
  if (e == cudaSuccess) {
    k_range_fold<<<(v.F + 127) / 128, 128, 0, v.stream>>>(v.F, v.nw, v.Wp, v.n, v.w_min, v.w_max, v.w_mean, v.w_ema, v.w_nover, v.q_num, v.q_den,
                                                         d_out, d_nover, d_st);
    // blocks per field: enough CTAs to fill the machine a few times over, each streaming >= 16 Ki keys
    const int bpf = (int)std::max<int64_t>(1, std::min<int64_t>((v.n + 16383) / 16384, std::max(1, 8 * v.sm_count / v.F + 1)));
    if (w_stat) {                                                  // experimental bounded select; fields it cannot finish stay !done
      e = cudaMallocAsync(&d_s2, v.F * sizeof(Sel2State), v.stream);
      if (e == cudaSuccess) e = cudaMallocAsync(&d_lists2, (size_t)v.F * kCollect2Max * sizeof(unsigned long long), v.stream);
      if (e == cudaSuccess) {
        k_sel2_bounds<<<(v.F + 127) / 128, 128, 0, v.stream>>>(v.F, v.nw, w_stat, d_s2);
        k_sel2_collect<<<dim3(bpf, v.F), 256, 0, v.stream>>>(v.ring, v.cap, v.start, v.n, d_st, d_s2, d_lists2);
        k_sel2_final<<<v.F, 256, 0, v.stream>>>(d_st, d_s2, d_lists2);
        e = cudaGetLastError();
      }
    }
    if (e == cudaSuccess && v.n > kCollectMax) {
      for (int pass = 0; pass < kMaxHistPasses; ++pass) {          // later passes exit at once for fields that are already narrow
        k_sel_hist<<<dim3(bpf, v.F), 256, 0, v.stream>>>(v.ring, v.cap, v.start, v.n, d_st, d_hist);
        k_sel_pick<<<v.F, 256, 0, v.stream>>>(d_hist, d_st);
      }
    }
    if (e == cudaSuccess) {
      k_sel_collect<<<dim3(bpf, v.F), 256, 0, v.stream>>>(v.ring, v.cap, v.start, v.n, d_st, d_lists);
      k_sel_final<<<v.F, 256, 0, v.stream>>>(d_st, d_lists, v.F, d_out);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_f64, d_out, 5 * v.F * sizeof(double), cudaMemcpyDeviceToHost, v.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_n_over, d_nover, v.F * sizeof(uint32_t), cudaMemcpyDeviceToHost, v.stream);
  cudaFreeAsync(d_out, v.stream); cudaFreeAsync(d_nover, v.stream); cudaFreeAsync(d_hist, v.stream);
  cudaFreeAsync(d_st, v.stream); cudaFreeAsync(d_lists, v.stream);
  if (d_s2) cudaFreeAsync(d_s2, v.stream);
  if (d_lists2) cudaFreeAsync(d_lists2, v.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v.stream);
  if (e != cudaSuccess) return gpud_fail(ctx, GPUD_E_CUDA, "reduce_range: %s", cudaGetErrorString(e));
  return GPUD_OK;
}
