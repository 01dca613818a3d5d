// Read-only ceiling of a cp.async.bulk (TMA, UBLKCP) landing ring for the window kernel's access pattern: the judge's round-1 item 5.
// Every window (8000 bytes, one contiguous run) is pulled into shared memory by ONE bulk copy that completes on an mbarrier; the
// warp then reads it with conflict-free LDS.128 and adds it up (the minimum a consumer can do).  A warp owns S stage buffers: while it
// consumes one, S - 1 copies are in flight and occupy no registers.  With 227 KB of shared memory an SM holds 28 buffers of 8 KB,
// so (warps per SM) x S <= 28 - which is the whole trade: the window kernel needs 12-16 resident warps for its post-processing, and the
// landing ring then has LESS in flight (<= 28 - warps buffers) than today's 16 warps x 8 KB of registers plus 142 KB of parked rows.
// Prints GB/s per (warps per CTA, stages) and the LDG baseline of tools/read_peak.cu for the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/bulk_read_peak tools/bulk_read_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kWin = 8000;        // bytes per window (W = 1000 doubles)
constexpr int kBuf = 8192;        // stage size

__device__ __forceinline__ void mbar_init(uint32_t a, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect(uint32_t a, int bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, int bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, int parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}

template <int WARPS, int S>
__global__ void __launch_bounds__(WARPS * 32, 1) k_bulk_read(const char* __restrict__ p, long long n_units, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[WARPS * S];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* mine = smem + (size_t)warp * S * kBuf;
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(mine);
  const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(bars + warp * S);
  if (lane == 0)
    for (int s = 0; s < S; ++s) mbar_init(bar_base + 8 * s, 1);
  __syncwarp();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long stride = (long long)gridDim.x * WARPS;
  long long u_issue = (long long)blockIdx.x * WARPS + warp, u = u_issue;
  // prologue: fill the ring
  if (lane == 0)
    for (int s = 0; s < S && u_issue < n_units; ++s, u_issue += stride) {
      mbar_expect(bar_base + 8 * s, kWin);
      bulk_g2s(sm_base + s * kBuf, p + u_issue * kWin, kWin, bar_base + 8 * s);
    }
  double acc = 0;
  int s = 0, parity = 0;
  for (; u < n_units; u += stride) {
    mbar_wait(bar_base + 8 * s, parity);
    const double2* b = reinterpret_cast<const double2*>(mine + s * kBuf) + lane;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j * 32 + lane < 500) { const double2 v = b[32 * j]; acc += v.x + v.y; }
    __syncwarp();                                  // every lane has its values in registers before the buffer is handed back
    if (lane == 0 && u_issue < n_units) {
      mbar_expect(bar_base + 8 * s, kWin);
      bulk_g2s(sm_base + s * kBuf, p + u_issue * kWin, kWin, bar_base + 8 * s);
    }
    u_issue += stride;
    if (++s == S) { s = 0; parity ^= 1; }
  }
  if (acc == 1.2345) out[0] = acc;
}

// the LDG landing of the shipped kernel (tools/read_peak.cu, mode 3), 16 warps per SM
__global__ void __launch_bounds__(256) k_ldg_read(const double2* __restrict__ p, long long n_units, double* out) {
  const int lane = threadIdx.x & 31;
  long long u = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * 8;
  double acc = 0;
  for (; u < n_units; u += stride) {
    const double2* b = p + u * 500 + lane;
    double2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] = make_double2(0, 0);
      if (j * 32 + lane < 500) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(v[j].x), "=d"(v[j].y) : "l"(b + 32 * j));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += v[j].x + v[j].y;
  }
  if (acc == 1.2345) out[0] = acc;
}

template <int WARPS, int S>
static void run(const char* d, long long n_units, double* o, cudaEvent_t e0, cudaEvent_t e1) {
  const size_t smem = (size_t)WARPS * S * kBuf;
  if (cudaFuncSetAttribute(k_bulk_read<WARPS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { printf("bulk warps %2d stages %d: smem %zu KB does not fit\n", WARPS, S, smem >> 10); cudaGetLastError(); return; }
  float best = 1e9;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0);
    k_bulk_read<WARPS, S><<<148, WARPS * 32, smem>>>(d, n_units, o);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { printf("bulk warps %d stages %d: %s\n", WARPS, S, cudaGetErrorString(cudaGetLastError())); return; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  printf("bulk  warps/SM %2d  stages %d  in flight %3d KB/SM  %.3f ms  %.1f GB/s\n", WARPS, S, WARPS * (S - 1) * 8, best, n_units * 8000.0 / best / 1e6);
}

int main() {
  const long long bytes = 4ll << 30, n_units = bytes / kWin;
  char* d; double* o;
  cudaMalloc(&d, bytes + 8192); cudaMalloc(&o, 8); cudaMemset(d, 0, bytes + 8192);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas = 2; ctas <= 4; ctas += 2) {
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
      cudaEventRecord(e0);
      k_ldg_read<<<148 * ctas, 256>>>((const double2*)d, n_units, o);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    printf("ldg   warps/SM %2d  (registers)  in flight %3d KB/SM  %.3f ms  %.1f GB/s\n", 8 * ctas, 8 * ctas * 8, best, n_units * 8000.0 / best / 1e6);
  }
  run<4, 6>(d, n_units, o, e0, e1);
  run<4, 7>(d, n_units, o, e0, e1);
  run<6, 4>(d, n_units, o, e0, e1);
  run<8, 2>(d, n_units, o, e0, e1);
  run<8, 3>(d, n_units, o, e0, e1);
  run<12, 2>(d, n_units, o, e0, e1);
  run<14, 2>(d, n_units, o, e0, e1);
  run<16, 1>(d, n_units, o, e0, e1);
  run<24, 1>(d, n_units, o, e0, e1);
  return 0;
}
