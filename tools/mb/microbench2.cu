// throughput microbenchmarks of the shared-memory crossbar / shuffle / fp64 pipes with W warps in ONE CTA (one SM)
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// 8 independent SHFL.32 per iteration
__global__ void k_shfl_tp(int* out, long long* t) {
  int x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1 + (k & 3));
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  out[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
// mode 0: every lane the same address; 1: 8-lane groups share an address, groups 48 B apart; 2: every lane its own (consecutive)
template <int BYTES, int MODE>
__global__ void k_lds_tp(double* out, long long* t) {
  __shared__ __align__(16) double s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int off;  // in doubles
  if (MODE == 0) off = 0;
  else if (MODE == 1) off = (threadIdx.x >> 3) * 6;
  else off = lane * (BYTES / 8) + (threadIdx.x >> 5) * 64;
  unsigned base = smem_u32(s + (off & 2047));
  double acc = 0.0;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (BYTES == 8) {
        double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + k * 512)); acc += v;
      } else {
        double v0, v1; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(base + k * 512)); acc += v0 + v1;
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_ddiv_tp(double* out, long long* t, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2; long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) { x0 = b / x0; x1 = b / x1; x2 = b / x2; }
  long long t1 = clock64(); out[threadIdx.x] = x0 + x1 + x2; __syncthreads(); if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_drcp_tp(double* out, long long* t, double a) {
  double x0 = a, x1 = a + 1, x2 = a + 2; long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) { x0 = __drcp_rn(x0) + 1.0; x1 = __drcp_rn(x1) + 1.0; x2 = __drcp_rn(x2) + 1.0; }
  long long t1 = clock64(); out[threadIdx.x] = x0 + x1 + x2; __syncthreads(); if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_dfma_tp(double* out, long long* t, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3; long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
  long long t1 = clock64(); out[threadIdx.x] = x0 + x1 + x2 + x3; __syncthreads(); if (threadIdx.x == 0) t[0] = t1 - t0;
}
int main() {
  double* d; long long* t; int* ii; cudaMalloc(&d, 8192 * 8); cudaMalloc(&t, 64); cudaMalloc(&ii, 8192 * 4);
  long long h;
#define RUN(name, call, ops) call; cudaDeviceSynchronize(); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); printf("%-44s %8.3f clk per warp-instruction (SM-wide)\n", name, (double)h / (ops));
  for (int w : {1, 4, 10, 13, 16}) {
    char nm[96];
    snprintf(nm, 96, "SHFL.32 x8 indep, %d warps", w); RUN(nm, (k_shfl_tp<<<1, 32 * w>>>(ii, t)), 8.0 * N * w)
    snprintf(nm, 96, "LDS.64 broadcast (1 addr), %d warps", w); RUN(nm, (k_lds_tp<8, 0><<<1, 32 * w>>>(d, t)), 8.0 * N * w)
    snprintf(nm, 96, "LDS.64 4 addr/warp (8-lane groups), %d warps", w); RUN(nm, (k_lds_tp<8, 1><<<1, 32 * w>>>(d, t)), 8.0 * N * w)
    snprintf(nm, 96, "LDS.64 32 addr/warp, %d warps", w); RUN(nm, (k_lds_tp<8, 2><<<1, 32 * w>>>(d, t)), 8.0 * N * w)
    snprintf(nm, 96, "LDS.128 broadcast (1 addr), %d warps", w); RUN(nm, (k_lds_tp<16, 0><<<1, 32 * w>>>(d, t)), 8.0 * N * w)
    snprintf(nm, 96, "LDS.128 4 addr/warp, %d warps", w); RUN(nm, (k_lds_tp<16, 1><<<1, 32 * w>>>(d, t)), 8.0 * N * w)
    snprintf(nm, 96, "LDS.128 32 addr/warp, %d warps", w); RUN(nm, (k_lds_tp<16, 2><<<1, 32 * w>>>(d, t)), 8.0 * N * w)
    snprintf(nm, 96, "DFMA x4 indep, %d warps", w); RUN(nm, (k_dfma_tp<<<1, 32 * w>>>(d, t, 1.0, 0.999)), 4.0 * N * w)
    snprintf(nm, 96, "DDIV x3 indep, %d warps", w); RUN(nm, (k_ddiv_tp<<<1, 32 * w>>>(d, t, 1.3, 2.1)), 3.0 * N * w)
    snprintf(nm, 96, "DRCP+DADD x3 indep, %d warps", w); RUN(nm, (k_drcp_tp<<<1, 32 * w>>>(d, t, 1.3)), 3.0 * N * w)
  }
  return 0;
}
