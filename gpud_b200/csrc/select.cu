// select.cu — whole-range aggregates of the newest n samples of every field: exact order statistic by MSB-first radix
// select (6 passes of 11/11/11/11/11/9 key bits), the other aggregates folded from the per-window pass of ring.cu.
// Definitions: oracle/SPEC.md (parity unpinned: the reference has no such aggregate).
// Algorithmic bytes: 8 B/sample; this first version re-reads the range once per radix pass (7 reads in total).
#include <stdarg.h>

#include <algorithm>

#include "internal.h"

namespace {

constexpr int kBins = 2048;
__constant__ int c_shift[6] = {53, 42, 31, 20, 9, 0};
__constant__ int c_bits[6] = {11, 11, 11, 11, 11, 9};

// grid (blocks_per_field, F); each block histograms its slice of the field's range for the current digit
__global__ void __launch_bounds__(256) k_sel_hist(const double* __restrict__ ring, int64_t cap, int64_t start, int64_t n, int pass,
                                                   const unsigned long long* __restrict__ prefix, unsigned* __restrict__ hist) {
  __shared__ unsigned s_hist[kBins];
  const int f = blockIdx.y;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int shift = c_shift[pass], bits = c_bits[pass];
  const unsigned long long want = pass ? prefix[f] : 0ull;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b = (int64_t)blockIdx.x * per, e = min(n, b + per);
  const double* __restrict__ base = ring + (int64_t)f * cap;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
    int64_t a = start + i;
    if (a >= cap) a -= cap;
    const unsigned long long key = gpud_f64_key((unsigned long long)__double_as_longlong(__ldcs(base + a)));
    if (pass == 0 || (key >> (shift + bits)) == want) atomicAdd(&s_hist[(unsigned)(key >> shift) & ((1u << bits) - 1u)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[(int64_t)f * kBins + i], s_hist[i]);
}

// one block per field: walk the digit histogram from the top, fix the next digit of the k-th largest key
__global__ void __launch_bounds__(256) k_sel_pick(unsigned* __restrict__ hist, int pass, unsigned long long* __restrict__ prefix,
                                                   unsigned long long* __restrict__ kk) {
  __shared__ unsigned s[kBins];
  __shared__ unsigned s_sum[256];
  const int f = blockIdx.x, t = threadIdx.x;
  const int bits = c_bits[pass], nb = 1 << bits;
  for (int i = t; i < kBins; i += 256) { s[i] = i < nb ? hist[(int64_t)f * kBins + i] : 0; hist[(int64_t)f * kBins + i] = 0; }
  __syncthreads();
  // thread t owns bins [8t, 8t+8) counted from the TOP: bin index nb-1-j
  unsigned loc = 0;
  for (int j = 8 * t; j < 8 * t + 8; ++j) if (j < nb) loc += s[nb - 1 - j];
  s_sum[t] = loc;
  __syncthreads();
  if (t == 0) {
    unsigned long long want = kk[f], acc = 0;
    int g = 0;
    while (g < 255 && acc + s_sum[g] < want) { acc += s_sum[g]; ++g; }
    int j = 8 * g;
    while (j < nb - 1 && acc + s[nb - 1 - j] < want) { acc += s[nb - 1 - j]; ++j; }
    prefix[f] = ((pass ? prefix[f] : 0ull) << bits) | (unsigned long long)(nb - 1 - j);
    kk[f] = want - acc;
  }
}

// fold the per-window partials into per-field results; init the select rank
__global__ void k_range_fold(int F, int nw, int Wp, int64_t n, const double* __restrict__ w_min, const double* __restrict__ w_max,
                             const double* __restrict__ w_mean, const double* __restrict__ w_ema, const uint32_t* __restrict__ w_nover,
                             int q_num, int q_den, double* __restrict__ out /*[5][F]*/, uint32_t* __restrict__ out_nover,
                             unsigned long long* __restrict__ kk) {
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
  kk[f] = (unsigned long long)(n - r + 1);
}

__global__ void k_sel_finish(int F, const unsigned long long* __restrict__ prefix, double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) out[4 * F + f] = __longlong_as_double((long long)gpud_key_f64bits(prefix[f]));
}

}  // namespace

extern "C" int32_t gpud_ring_reduce_range(gpud_ring* ring, int64_t last_n, double* out_f64, uint32_t* out_n_over) {
  if (!ring || !out_f64 || !out_n_over || last_n < 0) return GPUD_E_INVALID;
  gpud_range_view v;
  int64_t total = 0, count = 0, nwin = 0;
  gpud_ring_counts(ring, &total, &count, &nwin);
  // ctx is reachable only after the view is built; validate emptiness through counts first
  if (count == 0) return GPUD_E_STATE;
  int32_t rc = gpud_ring_range_partials(ring, last_n, &v);
  if (rc) return rc;
  gpud_ctx* ctx = v.ctx;
  GPUD_CUDA(ctx, cudaSetDevice(v.dev));
  double* d_out = nullptr;
  uint32_t* d_nover = nullptr;
  unsigned* d_hist = nullptr;
  unsigned long long *d_prefix = nullptr, *d_kk = nullptr;
  cudaError_t e = cudaMallocAsync(&d_out, 5 * v.F * sizeof(double), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_nover, v.F * sizeof(uint32_t), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_hist, (size_t)v.F * kBins * sizeof(unsigned), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_prefix, v.F * sizeof(unsigned long long), v.stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&d_kk, v.F * sizeof(unsigned long long), v.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_hist, 0, (size_t)v.F * kBins * sizeof(unsigned), v.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_prefix, 0, v.F * sizeof(unsigned long long), v.stream);
  if (e == cudaSuccess) {
    k_range_fold<<<(v.F + 127) / 128, 128, 0, v.stream>>>(v.F, v.nw, v.Wp, v.n, v.w_min, v.w_max, v.w_mean, v.w_ema, v.w_nover, v.q_num, v.q_den,
                                                         d_out, d_nover, d_kk);
    const int bpf = (int)std::max<int64_t>(1, std::min<int64_t>((v.n + 16383) / 16384, std::max(1, 4 * v.sm_count / v.F + 1)));
    for (int pass = 0; pass < 6; ++pass) {
      k_sel_hist<<<dim3(bpf, v.F), 256, 0, v.stream>>>(v.ring, v.cap, v.start, v.n, pass, d_prefix, d_hist);
      k_sel_pick<<<v.F, 256, 0, v.stream>>>(d_hist, pass, d_prefix, d_kk);
    }
    k_sel_finish<<<(v.F + 127) / 128, 128, 0, v.stream>>>(v.F, d_prefix, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_f64, d_out, 5 * v.F * sizeof(double), cudaMemcpyDeviceToHost, v.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_n_over, d_nover, v.F * sizeof(uint32_t), cudaMemcpyDeviceToHost, v.stream);
  cudaFreeAsync(d_out, v.stream); cudaFreeAsync(d_nover, v.stream); cudaFreeAsync(d_hist, v.stream);
  cudaFreeAsync(d_prefix, v.stream); cudaFreeAsync(d_kk, v.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(v.stream);
  if (e != cudaSuccess) return gpud_fail(ctx, GPUD_E_CUDA, "reduce_range: %s", cudaGetErrorString(e));
  return GPUD_OK;
}
