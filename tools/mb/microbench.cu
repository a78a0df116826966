// latency microbenchmarks (single warp unless stated), cycles per dependent op
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k_dfma(double* out, long long* t, double a, double b) {
  double x = a; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_dadd(double* out, long long* t, double a) {
  double x = a; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + a;
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_ddiv(double* out, long long* t, double a, double b) {
  double x = a; long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = b / x;
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* t, double a) {
  double x = a + threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1);
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* t) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 1) & 1023;
  __syncthreads();
  int x = threadIdx.x; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = s[x];
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* t) {
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) __syncthreads();
  long long t1 = clock64(); out[threadIdx.x] = 0; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_ffma(float* out, long long* t, float a, float b) {
  float x = a; long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fmaf(x, b, a);
  long long t1 = clock64(); out[threadIdx.x] = x; if (threadIdx.x == 0) t[0] = t1 - t0;
}
// throughput: many warps of independent DFMA
__global__ void k_dfma_tp(double* out, long long* t, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3; long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
  long long t1 = clock64(); out[threadIdx.x] = x0 + x1 + x2 + x3; if (threadIdx.x == 0) t[0] = t1 - t0;
}
int main() {
  double* d; long long* t; float* f; cudaMalloc(&d, 8192 * 8); cudaMalloc(&t, 64); cudaMalloc(&f, 8192 * 4);
  long long h;
#define RUN(name, call, div) call; cudaDeviceSynchronize(); cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); printf("%-28s %8.2f cycles/op\n", name, (double)h / (div));
  RUN("DFMA dependent (1 warp)", (k_dfma<<<1, 32>>>(d, t, 1.0, 0.999)), N)
  RUN("DADD dependent (1 warp)", (k_dadd<<<1, 32>>>(d, t, 1e-9)), N)
  RUN("FFMA dependent (1 warp)", (k_ffma<<<1, 32>>>(f, t, 1.f, 0.999f)), N)
  RUN("DDIV dependent (1 warp)", (k_ddiv<<<1, 32>>>(d, t, 1.3, 2.1)), N)
  RUN("SHFL.f64+DADD dependent", (k_shfl<<<1, 32>>>(d, t, 1.0)), N)
  RUN("LDS dependent (1 warp)", (k_lds<<<1, 32>>>(d, t)), N)
  RUN("BAR 10 warps", (k_bar<<<1, 320>>>(d, t)), N)
  RUN("BAR 16 warps", (k_bar<<<1, 512>>>(d, t)), N)
  RUN("DFMA x4 indep, 1 warp", (k_dfma_tp<<<1, 32>>>(d, t, 1.0, 0.999)), 4 * N)
  RUN("DFMA x4 indep, 16 warps", (k_dfma_tp<<<1, 512>>>(d, t, 1.0, 0.999)), 4 * N)
  RUN("DFMA x4 indep, 32 warps", (k_dfma_tp<<<1, 1024>>>(d, t, 1.0, 0.999)), 4 * N)
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
  return 0;
}
