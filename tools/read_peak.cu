// Read-only streaming ceiling for the window kernel's access pattern: every warp reads 8000-byte runs with 16 x LDG.128
// per lane in flight, persistent grid.  Prints GB/s for a few CTA/SM settings.  Build: nvcc -arch=sm_100a -O3 -o build/read_peak tools/read_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ double2 ldv(const double2* p) {
  double2 r;
  if (MODE == 0) return __ldcs(p);
  if (MODE == 1) { asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r; }
  if (MODE == 2) { asm volatile("ld.global.nc.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r; }
  if (MODE == 3) { asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r; }
  if (MODE == 4) { asm volatile("ld.global.cs.L2::256B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r; }
  if (MODE == 5) { asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p)); return r; }
  return r;
}
template <int NLD, int MODE>
__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ p, long long n_units, double* out) {
  const int lane = threadIdx.x & 31;
  long long u = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * 8;
  double acc = 0;
  for (; u < n_units; u += stride) {
    const double2* b = p + u * 500 + lane;
    double2 v[NLD];
#pragma unroll
    for (int j = 0; j < NLD; ++j) v[j] = (j * 32 + lane < 500) ? ldv<MODE>(b + 32 * j) : make_double2(0, 0);
#pragma unroll
    for (int j = 0; j < NLD; ++j) acc += v[j].x + v[j].y;
  }
  if (acc == 1.2345) out[0] = acc;
}
int main() {
  const long long bytes = 4ll << 30, n_units = bytes / 8000;
  double2* d; double* o;
  cudaMalloc(&d, bytes + 8192); cudaMalloc(&o, 8); cudaMemset(d, 0, bytes + 8192);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 6; ++mode)
  for (int ctas = 2; ctas <= 4; ctas *= 2) {
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
      cudaEventRecord(e0);
      switch (mode) {
        case 0: k_read<16, 0><<<148 * ctas, 256>>>(d, n_units, o); break;
        case 1: k_read<16, 1><<<148 * ctas, 256>>>(d, n_units, o); break;
        case 2: k_read<16, 2><<<148 * ctas, 256>>>(d, n_units, o); break;
        case 3: k_read<16, 3><<<148 * ctas, 256>>>(d, n_units, o); break;
        case 4: k_read<16, 4><<<148 * ctas, 256>>>(d, n_units, o); break;
        case 5: k_read<16, 5><<<148 * ctas, 256>>>(d, n_units, o); break;
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    printf("mode %d ctas/SM %d  %.3f ms  %.1f GB/s\n", mode, ctas, best, n_units * 8000.0 / best / 1e6);
  }
  return 0;
}
