// select.cu — whole-range aggregates of the newest n samples of every field (gpud_ring_reduce_range).
//
// min / max / mean / EMA / n_over are folded from the per-window pass of ring.cu (W' = 1024 windows over the range).
// Long ranges (>= 64 Ki samples, i.e. the whole-ring W = CAP order statistic of BASELINE configs[3]) take ONE HBM read: pivots
// from a 0.8 % sample, classification fused into the window pass, the answer from class counts + a short list (see below).
// Short ranges, and any field the sampled pass cannot settle, use an MSB-first radix select on the IEEE totalOrder keys that
// spends HBM passes only where they narrow the candidate set:
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

// ---- sampled-pivot single pass (ranges of >= GPUD_RANGE_SAMPLED_MIN samples) -----------------------------------------------------
// One block per field reads a systematic sample of S keys (one 32-byte sector each, the stride apart), and picks two sample order
// statistics around the wanted rank as the field's pivot pair lo <= hi: for the rank k from the top among n keys the sample rank is
// about k S / n with a standard deviation of sqrt(S p (1 - p)); five of those to either side (failure odds < 1e-6 per field for
// independent samples; an autocorrelated gauge is sampled evenly over the range, which is at least as good).  The window pass
// then counts the keys above hi / equal to hi / equal to lo and parks the keys strictly between the pivots in the field's list
// (about 10 S p-deviations x n / S keys: 12 Ki of 1 Mi for p99); k_range_finish reads the answer off the counts or selects it
// from the list.  Any field for which that fails (rank outside the pivots, list overflow, +NaN in the data, NaN pivots) is
// answered by the histogram path below - results never depend on the sample.
constexpr int kSampleMax = 8192;
constexpr int kSelThreads = 512;       // block size of k_range_pivots / k_range_finish: latency-bound, one block per field

// The kkA-th and kkB-th largest (1-based) of keys[0..cnt) in shared memory, both at once; kSelThreads threads.  MSB-first 8-bit digits
// starting below the bits all keys share; a target whose chosen bin holds a single key is finished by one more sweep that
// finds that key (after two digits of an 8 Ki sample almost every bin does).  Warp 0 walks target A's histogram, warp 1 B's.
struct Sel2 { unsigned long long prefix[2]; unsigned kk[2]; int single[2]; unsigned long long ans[2]; int done[2]; };
__device__ void block_select2_smem(const unsigned long long* keys, int cnt, int kkA, int kkB, unsigned* hist /*[2][256]*/, Sel2* z,
                                   unsigned long long* red /*[16]*/, unsigned long long* outA, unsigned long long* outB) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  unsigned long long kmin = ~0ull, kmax = 0ull;
  for (int i = t; i < cnt; i += kSelThreads) { const unsigned long long k = keys[i]; kmin = min(kmin, k); kmax = max(kmax, k); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if (lane == 0) { red[wid] = kmin; red[kSelThreads / 32 + wid] = kmax; }
  __syncthreads();
  for (int w = 0; w < kSelThreads / 32; ++w) { kmin = min(kmin, red[w]); kmax = max(kmax, red[kSelThreads / 32 + w]); }
  if (kmin == kmax) { *outA = *outB = kmax; return; }
  int nfix = __clzll((long long)(kmin ^ kmax));
  if (t == 0) {
    for (int x = 0; x < 2; ++x) { z->prefix[x] = nfix ? (kmax >> (64 - nfix)) : 0ull; z->single[x] = 0; z->done[x] = 0; z->ans[x] = 0ull; }
    z->kk[0] = (unsigned)kkA; z->kk[1] = (unsigned)kkB;
  }
  __syncthreads();
  while (nfix < 64) {
    const int d = min(8, 64 - nfix), shift = 64 - nfix - d;
    const unsigned mask = (1u << d) - 1u;
    const unsigned long long pA = z->prefix[0], pB = z->prefix[1];
    const bool liveA = !z->done[0], liveB = !z->done[1];
    if (!liveA && !liveB) break;
    if (t < 512) hist[t] = 0;
    __syncthreads();
    for (int i = t; i < cnt; i += kSelThreads) {
      const unsigned long long key = keys[i];
      const unsigned long long top = nfix ? (key >> (64 - nfix)) : 0ull;
      const unsigned bin = (unsigned)(key >> shift) & mask;
      if (liveA && top == pA) atomicAdd(&hist[bin], 1u);
      if (liveB && top == pB) atomicAdd(&hist[256 + bin], 1u);
    }
    __syncthreads();
    if (wid < 2 && !z->done[wid]) {                  // lane l owns bins 255-8l .. 248-8l (descending) of target `wid`
      const unsigned* h = hist + 256 * wid;
      const unsigned kk = z->kk[wid];
      unsigned loc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) loc += h[255 - 8 * lane - j];
      unsigned incl = loc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      const unsigned before = incl - loc;            // keys in the bins above this lane's
      const unsigned long long pfx = z->prefix[wid];
      __syncwarp();                                  // every lane has read kk / prefix before the owner rewrites them
      if (before < kk && kk <= incl) {
        unsigned acc = before;
        int bin = 255 - 8 * lane;
        while (acc + h[bin] < kk) { acc += h[bin]; --bin; }
        z->prefix[wid] = (pfx << d) | (unsigned long long)bin;
        z->kk[wid] = kk - acc;
        z->single[wid] = h[bin] == 1u;
      }
    }
    __syncthreads();
    nfix += d;
    if (nfix == 64) {
      if (t < 2 && !z->done[t]) { z->ans[t] = z->prefix[t]; z->done[t] = 1; }
    } else if ((z->single[0] && !z->done[0]) || (z->single[1] && !z->done[1])) {
      const bool sA = z->single[0] && !z->done[0], sB = z->single[1] && !z->done[1];
      const unsigned long long qA = z->prefix[0], qB = z->prefix[1];
      __syncthreads();
      for (int i = t; i < cnt; i += kSelThreads) {
        const unsigned long long key = keys[i], top = key >> (64 - nfix);
        if (sA && top == qA) { z->ans[0] = key; z->done[0] = 1; }
        if (sB && top == qB) { z->ans[1] = key; z->done[1] = 1; }
      }
    }
    __syncthreads();
  }
  *outA = z->ans[0];
  *outB = z->ans[1];
}

// The sample comes from the ring's own sample ring when the range covers enough of its slots (the whole-ring case: 64 KB of contiguous
// memory per field, kept current by the append kernel), else from S strided reads of the ring itself (one 32-byte sector each).
__global__ void __launch_bounds__(kSelThreads) k_range_pivots(const double* __restrict__ ring, int64_t cap, int64_t start, int64_t n, int S, int64_t stride,
                                                       int64_t k_from_top, double* __restrict__ piv, unsigned* __restrict__ fill,
                                                       const double* __restrict__ sample, int smp_shift, int64_t smp_slots) {
  extern __shared__ __align__(16) unsigned long long s_keys[];
  __shared__ unsigned s_hist[512];
  __shared__ unsigned long long s_red[2 * kSelThreads / 32];
  __shared__ Sel2 s_z;
  __shared__ int s_cnt;
  const int f = blockIdx.x, t = threadIdx.x;
  if (sample) {
    if (t == 0) s_cnt = 0;
    __syncthreads();
    const double* __restrict__ sb = sample + (int64_t)f * smp_slots;
    const int64_t half = (1ll << smp_shift) >> 1;
    for (int64_t s0 = 0; s0 < smp_slots; s0 += kSelThreads) {             // slots compacted with one ballot + one counter bump per warp
      const int64_t sl = s0 + t;
      bool ok = false;
      if (sl < smp_slots) {
        const int64_t col = (sl << smp_shift) + half;            // the column this slot mirrors
        int64_t c = col - start;                                 // chronological index inside the range?
        if (c < 0) c += cap;
        ok = col < cap && c < n;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      int base = 0;
      if ((t & 31) == 0 && m) base = atomicAdd(&s_cnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (ok) s_keys[base + __popc(m & ((1u << (t & 31)) - 1u))] = gpud_f64_key((unsigned long long)__double_as_longlong(__ldg(sb + sl)));
    }
    __syncthreads();
    S = s_cnt;
  } else {
    const double* __restrict__ base = ring + (int64_t)f * cap;
    for (int j = t; j < S; j += kSelThreads) {
      int64_t a = start + (int64_t)j * stride + (stride >> 1);
      if (a >= cap) a -= cap;
      s_keys[j] = gpud_f64_key((unsigned long long)__double_as_longlong(__ldg(base + a)));
    }
  }
  __syncthreads();
  const double pr = (double)k_from_top / (double)n;
  const double r0 = pr * (double)S, d = 5.0 * sqrt((double)S * pr * (1.0 - pr)) + 2.0;
  const long long rh = (long long)floor(r0 - d), rl = (long long)ceil(r0 + d);
  unsigned long long hik, lok;
  block_select2_smem(s_keys, S, (int)max(1ll, min((long long)S, rh)), (int)max(1ll, min((long long)S, rl)), s_hist, &s_z, s_red, &hik, &lok);
  if (rh < 1) hik = gpud_f64_key(0x7ff0000000000000ull);      // nothing can be trusted to lie above: +inf
  if (rl > S) lok = gpud_f64_key(0xfff0000000000000ull);      // -inf
  if (t == 0) {
    piv[2 * f] = __longlong_as_double((long long)gpud_key_f64bits(lok));
    piv[2 * f + 1] = __longlong_as_double((long long)gpud_key_f64bits(hik));
    fill[f] = 0u;
  }
}

// One block per field after the window pass: fold the per-window partials into the field's min / max / mean / EMA / n_over,
// initialise the histogram path's state, and - sampled mode - settle the order statistic from the class counts and the list.
// Fields left open (done == 0) are counted in *n_open.
__global__ void __launch_bounds__(kSelThreads) k_range_finish(int F, int nw, int Wp, int64_t n, const double* __restrict__ w_min, const double* __restrict__ w_max,
                                                       const double* __restrict__ w_mean, const double* __restrict__ w_ema, const uint32_t* __restrict__ w_nover,
                                                       int q_num, int q_den, int sampled, const double* __restrict__ piv, const unsigned* __restrict__ fill,
                                                       const uint4* __restrict__ w_cls, const unsigned long long* __restrict__ lists, unsigned list_cap,
                                                       double* __restrict__ out /*[5][F]*/, uint32_t* __restrict__ out_nover, SelState* __restrict__ st,
                                                       unsigned* __restrict__ n_open, int* __restrict__ open_ids) {
  __shared__ unsigned long long s_red[kSelThreads / 32][6];
  __shared__ double s_sum[kSelThreads / 32];
  __shared__ unsigned s_hist[kBins];
  __shared__ unsigned long long s_prefix, s_kk;
  __shared__ int s_nbits, s_mode;
  __shared__ unsigned s_part[kSelThreads / 32];
  const int f = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  unsigned long long kmin = ~0ull, kmax = 0ull, c_abv = 0, c_eh = 0, c_el = 0, nov = 0;
  double sum = 0.0;
  for (int w = t; w < nw; w += kSelThreads) {
    const int64_t o = (int64_t)f * nw + w;
    const int m = (int)min((int64_t)Wp, n - (int64_t)w * Wp);
    kmin = min(kmin, gpud_f64_key((unsigned long long)__double_as_longlong(w_min[o])));
    kmax = max(kmax, gpud_f64_key((unsigned long long)__double_as_longlong(w_max[o])));
    sum += w_mean[o] * (double)m;
    nov += w_nover[o];
    if (sampled) { const uint4 c = w_cls[o]; c_abv += c.x; c_eh += c.y; c_el += c.z; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    c_abv += __shfl_xor_sync(0xffffffffu, c_abv, o);
    c_eh += __shfl_xor_sync(0xffffffffu, c_eh, o);
    c_el += __shfl_xor_sync(0xffffffffu, c_el, o);
    nov += __shfl_xor_sync(0xffffffffu, nov, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (lane == 0) { s_red[wid][0] = kmin; s_red[wid][1] = kmax; s_red[wid][2] = c_abv; s_red[wid][3] = c_eh; s_red[wid][4] = c_el; s_red[wid][5] = nov; s_sum[wid] = sum; }
  __syncthreads();
  if (t == 0) {
    for (int w = 1; w < kSelThreads / 32; ++w) {
      kmin = min(kmin, s_red[w][0]); kmax = max(kmax, s_red[w][1]);
      c_abv += s_red[w][2]; c_eh += s_red[w][3]; c_el += s_red[w][4]; nov += s_red[w][5];
      sum += s_sum[w];
    }
    out[0 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(kmin));
    out[1 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(kmax));
    out[2 * F + f] = sum / (double)n;
    out[3 * F + f] = w_ema[(int64_t)f * nw + nw - 1];
    out_nover[f] = (uint32_t)nov;
    long long r = (n * q_num + q_den - 1) / q_den;
    r = r < 1 ? 1 : (r > n ? n : r);
    SelState s;
    s.kk = (unsigned long long)(n - r + 1);
    s.cnt = (unsigned long long)n;
    s.collected = 0;
    s.pad = 0;
    const unsigned long long diff = kmin ^ kmax;
    int mode = 0;                      // 0: left to the histogram path, 1: answered, 2: select from the list
    int why = GPUD_RANGE_OPEN_SHORT;   // why the field stays open (gpud_ring_range_stats)
    if (diff == 0ull) {                // a constant field: every key is the answer
      s.prefix = kmax; s.nbits = 64; s.done = 1; s.ans = kmax;
      mode = 1;
    } else {
      const int cp = __clzll((long long)diff);   // leading bits common to all keys of the field
      s.nbits = cp;
      s.prefix = cp ? (kmax >> (64 - cp)) : 0ull;
      s.done = 0;
      s.ans = 0ull;
      if (sampled && kmax > gpud_f64_key(0x7ff0000000000000ull)) why = GPUD_RANGE_OPEN_NAN;   // a +NaN was classified as "below"
      if (sampled && kmax <= gpud_f64_key(0x7ff0000000000000ull)) {
        const unsigned long long lok = gpud_f64_key((unsigned long long)__double_as_longlong(piv[2 * f]));
        const unsigned long long hik = gpud_f64_key((unsigned long long)__double_as_longlong(piv[2 * f + 1]));
        const unsigned long long inside = fill[f];
        const unsigned long long k = s.kk;
        why = !(lok <= hik) ? GPUD_RANGE_OPEN_PIVOTS : (k <= c_abv ? GPUD_RANGE_OPEN_ABOVE : GPUD_RANGE_OPEN_BELOW);
        if (lok <= hik && k > c_abv) {
          if (k <= c_abv + c_eh) { s.done = 1; s.ans = hik; mode = 1; }
          else if (k <= c_abv + c_eh + inside) {
            why = GPUD_RANGE_OPEN_OVERFLOW;
            if (inside <= (unsigned long long)list_cap) {
              mode = 2;
              const unsigned long long d2 = lok ^ hik;     // every listed key lies between the pivots: skip their common bits
              const int cp2 = __clzll((long long)d2);
              s_nbits = cp2;
              s_prefix = cp2 ? (hik >> (64 - cp2)) : 0ull;
              s_kk = k - c_abv - c_eh;
            }
          } else if (k <= c_abv + c_eh + inside + c_el) { s.done = 1; s.ans = lok; mode = 1; }
        }
      }
    }
    if (s.done) { s.nbits = 64; s.prefix = s.ans; }
    s.pad = mode == 0 ? why : 0;
    st[f] = s;
    s_mode = mode;
    if (mode == 1) out[4 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(s.ans));
    if (mode == 0) open_ids[atomicAdd(n_open, 1u)] = f;
  }
  __syncthreads();
  if (s_mode != 2) return;
  const unsigned cnt = fill[f];
  const unsigned long long* __restrict__ list = lists + (int64_t)f * list_cap;
  while (s_nbits < 64) {               // radix select over the L2-resident list, 11-bit digits
    const int nbits = s_nbits;
    const unsigned long long prefix = s_prefix;
    const int d = min(kDigit, 64 - nbits), shift = 64 - nbits - d, nb = 1 << d;
    for (int i = t; i < kBins; i += kSelThreads) s_hist[i] = 0;
    __syncthreads();
    for (unsigned i = t; i < cnt; i += kSelThreads) {
      const unsigned long long key = list[i];
      if (nbits == 0 || (key >> (64 - nbits)) == prefix) atomicAdd(&s_hist[(unsigned)(key >> shift) & (unsigned)(nb - 1)], 1u);
    }
    __syncthreads();
    // thread t owns kBpt bins from the top (nb-1-kBpt t downwards); block-wide exclusive scan of the per-thread sums finds the owner
    constexpr int kBpt = kBins / kSelThreads;
    unsigned loc = 0;
#pragma unroll
    for (int j = 0; j < kBpt; ++j) { const int b = nb - 1 - kBpt * t - j; if (b >= 0) loc += s_hist[b]; }
    unsigned incl = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_part[wid] = incl;
    __syncthreads();
    unsigned before = incl - loc;
    for (int w = 0; w < wid; ++w) before += s_part[w];
    const unsigned long long kk = s_kk;
    __syncthreads();                                   // everyone has read s_kk / s_part before the owner rewrites them
    if ((unsigned long long)before < kk && kk <= (unsigned long long)before + loc) {
      unsigned acc = before;
      int b = nb - 1 - kBpt * t;
      while ((unsigned long long)acc + s_hist[b] < kk) { acc += s_hist[b]; --b; }
      s_prefix = (nbits ? (prefix << d) : 0ull) | (unsigned long long)b;
      s_nbits = nbits + d;
      s_kk = kk - acc;
    }
    __syncthreads();
  }
  if (t == 0) {
    SelState s = st[f];
    s.done = 1; s.ans = s_prefix; s.nbits = 64; s.prefix = s_prefix;
    st[f] = s;
    out[4 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(s_prefix));
  }
}

// grid (blocks_per_field, F): histogram of the next digit over the keys that match the field's prefix
__global__ void __launch_bounds__(256) k_sel_hist(const double* __restrict__ ring, int64_t cap, int64_t start, int64_t n,
                                                   const SelState* __restrict__ st, unsigned* __restrict__ hist, const int* __restrict__ ids) {
  __shared__ unsigned s_hist[kBins];
  const int f = ids ? ids[blockIdx.y] : blockIdx.y;
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
    if (s_hist[i]) atomicAdd(&hist[(int64_t)blockIdx.y * kBins + i], s_hist[i]);
}

// one block per field: walk the digit histogram from the top, fix the next digit of the k-th largest key
__global__ void __launch_bounds__(256) k_sel_pick(unsigned* __restrict__ hist, SelState* __restrict__ st, const int* __restrict__ ids) {
  __shared__ unsigned s[kBins];
  __shared__ unsigned s_sum[256];
  const int f = ids ? ids[blockIdx.x] : blockIdx.x, t = threadIdx.x;
  SelState cur = st[f];
  if (cur.done || cur.cnt <= (unsigned long long)kCollectMax) return;
  const int d = min(kDigit, 64 - cur.nbits), nb = 1 << d;
  for (int i = t; i < kBins; i += 256) { s[i] = i < nb ? hist[(int64_t)blockIdx.x * kBins + i] : 0; hist[(int64_t)blockIdx.x * kBins + i] = 0; }
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
                                                      SelState* __restrict__ st, unsigned long long* __restrict__ lists, const int* __restrict__ ids) {
  const int f = ids ? ids[blockIdx.y] : blockIdx.y;
  const SelState s = st[f];
  if (s.done) return;
  const int64_t per = (((n + gridDim.x - 1) / gridDim.x) + 1) & ~(int64_t)1;
  const int64_t b = min(n, (int64_t)blockIdx.x * per), e = min(n, b + per);
  const double* __restrict__ base = ring + (int64_t)f * cap;
  unsigned long long* __restrict__ list = lists + (int64_t)blockIdx.y * kCollectMax;
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
                                                    double* __restrict__ out, const int* __restrict__ ids) {
  __shared__ unsigned s_cnt[8];
  const int f = ids ? ids[blockIdx.x] : blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const SelState s = st[f];
  unsigned long long ans = s.ans;
  if (!s.done) {
    const unsigned n = (unsigned)min((unsigned long long)kCollectMax, (unsigned long long)s.collected);
    const unsigned long long* __restrict__ list = lists + (int64_t)blockIdx.x * kCollectMax;
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

}  // namespace

extern "C" int32_t gpud_ring_reduce_range(gpud_ring* ring, int64_t last_n, double* out_f64, uint32_t* out_n_over) {
  if (!ring || !out_f64 || !out_n_over || last_n < 0) return GPUD_E_INVALID;
  gpud_range_view v;
  int64_t total = 0, count = 0, nwin = 0;
  gpud_ring_counts(ring, &total, &count, &nwin);
  if (count == 0) return GPUD_E_STATE;
  { int32_t rc = gpud_ring_range_prepare(ring, last_n, &v); if (rc) return rc; }
  gpud_ctx* ctx = v.ctx;
  double* d_out = nullptr;
  uint32_t* d_nover = nullptr;
  unsigned* d_hist = nullptr;          // histogram path scratch, allocated only when a field needs it
  SelState* d_st = nullptr;
  unsigned long long* d_lists = nullptr;
  unsigned* d_open = nullptr;          // [0] = number of fields left open by k_range_finish, [1..] = their ids
  unsigned h_open = 0;
  int reasons[GPUD_RANGE_N_OPEN_REASONS] = {0};
  cudaError_t e = cudaMallocAsync(&d_out, 5 * v.F * sizeof(double), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nover, v.F * sizeof(uint32_t), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_st, v.F * sizeof(SelState), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_open, (1 + (size_t)v.F) * sizeof(unsigned), v.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_open, 0, sizeof(unsigned), v.stream);
  if (e == cudaSuccess) e = cudaEventRecord(v.ev[0], v.stream);
  if (e == cudaSuccess && v.sampled) {
    long long r = (long long)((v.n * v.q_num + v.q_den - 1) / v.q_den);
    r = r < 1 ? 1 : (r > v.n ? v.n : r);
    const int S = (int)std::min<int64_t>(kSampleMax, v.n / 8);
    const int64_t stride = v.n / S;
    const bool use_ring_sample = (v.n >> v.smp_shift) >= 2048 && v.smp_slots <= kSampleMax;   // else too few of the ring's sample slots fall into the range
    e = cudaFuncSetAttribute(k_range_pivots, cudaFuncAttributeMaxDynamicSharedMemorySize, kSampleMax * (int)sizeof(unsigned long long));
    if (e == cudaSuccess) k_range_pivots<<<v.F, kSelThreads, (size_t)kSampleMax * sizeof(unsigned long long), v.stream>>>(v.ring, v.cap, v.start, v.n, S, stride, v.n - r + 1, v.piv, v.fill,
                                                                                                          use_ring_sample ? v.sample : nullptr, v.smp_shift, v.smp_slots);
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    int32_t rc = gpud_ring_range_pass(ring, &v);
    if (rc) { cudaFreeAsync(d_out, v.stream); cudaFreeAsync(d_nover, v.stream); cudaFreeAsync(d_st, v.stream); cudaFreeAsync(d_open, v.stream); return rc; }
    e = cudaEventRecord(v.ev[1], v.stream);
  }
  if (e == cudaSuccess) {
    k_range_finish<<<v.F, kSelThreads, 0, v.stream>>>(v.F, v.nw, v.Wp, v.n, v.w_min, v.w_max, v.w_mean, v.w_ema, v.w_nover, v.q_num, v.q_den, v.sampled, v.piv, v.fill,
                                              v.w_cls, v.lists, v.list_cap, d_out, d_nover, d_st, d_open, reinterpret_cast<int*>(d_open + 1));
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaEventRecord(v.ev[2], v.stream);
  if (e == cudaSuccess && v.sampled) {   // a short range leaves every non-constant field to the radix select: no need to ask
    e = cudaMemcpyAsync(&h_open, d_open, sizeof(unsigned), cudaMemcpyDeviceToHost, v.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_f64, d_out, 5 * v.F * sizeof(double), cudaMemcpyDeviceToHost, v.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_n_over, d_nover, v.F * sizeof(uint32_t), cudaMemcpyDeviceToHost, v.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v.stream);
  }
  if (e == cudaSuccess && (h_open > 0 || !v.sampled)) {
    // radix select for the fields still open: every non-constant field of a short range, or the rare field the sampled pass
    // could not settle.  The grid covers only the open fields, with enough blocks each to fill the machine.
    const int* ids = v.sampled ? reinterpret_cast<const int*>(d_open + 1) : nullptr;
    const int n_sel = v.sampled ? (int)h_open : v.F;
    if (v.sampled) {
      std::vector<SelState> h_st(v.F);
      e = cudaMemcpyAsync(h_st.data(), d_st, v.F * sizeof(SelState), cudaMemcpyDeviceToHost, v.stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(v.stream);
      for (const SelState& q : h_st) if (q.pad > 0 && q.pad < GPUD_RANGE_N_OPEN_REASONS) ++reasons[q.pad];
    }
    if (e == cudaSuccess) e = cudaMallocAsync(&d_hist, (size_t)n_sel * kBins * sizeof(unsigned), v.stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_lists, (size_t)n_sel * kCollectMax * sizeof(unsigned long long), v.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_hist, 0, (size_t)n_sel * kBins * sizeof(unsigned), v.stream);
    if (e == cudaSuccess) {
      // blocks per field: enough CTAs to fill the machine a few times over, each streaming >= 16 Ki keys
      const int bpf = (int)std::max<int64_t>(1, std::min<int64_t>((v.n + 16383) / 16384, std::max(1, 8 * v.sm_count / n_sel + 1)));
      if (v.n > kCollectMax) {
        for (int pass = 0; pass < kMaxHistPasses; ++pass) {          // later passes exit at once for fields that are already narrow
          k_sel_hist<<<dim3(bpf, n_sel), 256, 0, v.stream>>>(v.ring, v.cap, v.start, v.n, d_st, d_hist, ids);
          k_sel_pick<<<n_sel, 256, 0, v.stream>>>(d_hist, d_st, ids);
        }
      }
      k_sel_collect<<<dim3(bpf, n_sel), 256, 0, v.stream>>>(v.ring, v.cap, v.start, v.n, d_st, d_lists, ids);
      k_sel_final<<<n_sel, 256, 0, v.stream>>>(d_st, d_lists, v.F, d_out, ids);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess && !v.sampled) {
      e = cudaMemcpyAsync(out_f64, d_out, 5 * v.F * sizeof(double), cudaMemcpyDeviceToHost, v.stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(out_n_over, d_nover, v.F * sizeof(uint32_t), cudaMemcpyDeviceToHost, v.stream);
    } else if (e == cudaSuccess) {
      e = cudaMemcpyAsync(out_f64 + 4 * (size_t)v.F, d_out + 4 * (size_t)v.F, v.F * sizeof(double), cudaMemcpyDeviceToHost, v.stream);
    }
    if (d_hist) cudaFreeAsync(d_hist, v.stream);
    if (d_lists) cudaFreeAsync(d_lists, v.stream);
  }
  cudaFreeAsync(d_out, v.stream); cudaFreeAsync(d_nover, v.stream); cudaFreeAsync(d_st, v.stream); cudaFreeAsync(d_open, v.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v.stream);
  if (e != cudaSuccess) return gpud_fail(ctx, GPUD_E_CUDA, "reduce_range: %s", cudaGetErrorString(e));
  gpud_ring_range_note(ring, v.sampled != 0, h_open, reasons);
  return GPUD_OK;
}
