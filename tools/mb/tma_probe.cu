// tools/mb/tma_probe.cu -- minimal 2-D u8 TMA tile load (cp.async.bulk.tensor.2d + mbarrier), used to pin down the descriptor /
// PTX conventions the FAST kernel relies on.  One case per process (a faulting case kills the context):
//   tma_probe <mode 0=struct-array param | 1=single __grid_constant__ param | 2=descriptor in global memory> <box 0..3> <x> <y> <w> <h> <pitch>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
struct Maps { CUtensorMap m[12]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void body(const CUtensorMap* map, int bytes, int x, int y, uint8_t* out, int* status, uint8_t* s, uint64_t* bar) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(s)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
  }
  uint32_t ok = 0;
  long long spins = 0;
  for (; spins < (1 << 22) && !ok; ++spins)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
  if (threadIdx.x == 0) { status[0] = ok; status[1] = (int)spins; }
  __syncthreads();
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = ok ? s[i] : 0xEE;
}
__global__ void probe_array(const __grid_constant__ Maps M, int l, int bytes, int x, int y, uint8_t* out, int* status) {
  __shared__ __align__(128) uint8_t s[8192];
  __shared__ __align__(8) uint64_t bar;
  body(&M.m[l], bytes, x, y, out, status, s, &bar);
}
__global__ void probe_single(const __grid_constant__ CUtensorMap M, int bytes, int x, int y, uint8_t* out, int* status) {
  __shared__ __align__(128) uint8_t s[8192];
  __shared__ __align__(8) uint64_t bar;
  body(&M, bytes, x, y, out, status, s, &bar);
}
__global__ void probe_global(const CUtensorMap* M, int bytes, int x, int y, uint8_t* out, int* status) {
  __shared__ __align__(128) uint8_t s[8192];
  __shared__ __align__(8) uint64_t bar;
  body(M, bytes, x, y, out, status, s, &bar);
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, bsel = argc > 2 ? atoi(argv[2]) : 0;
  const int x = argc > 3 ? atoi(argv[3]) : 0, y = argc > 4 ? atoi(argv[4]) : 0;
  const int w = argc > 5 ? atoi(argv[5]) : 1920, h = argc > 6 ? atoi(argv[6]) : 1080, pitch = argc > 7 ? atoi(argv[7]) : ((w + 127) & ~127);
  const int BWs[5] = {144, 64, 128, 256, 160}, BHs[5] = {40, 32, 32, 8, 40};
  const int BW = BWs[bsel % 5], BH = BHs[bsel % 5];
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { printf("no entry point\n"); return 2; }
  EncodeFn encode = (EncodeFn)fn;
  std::vector<uint8_t> img((size_t)pitch * h);
  for (int r = 0; r < h; ++r) for (int c = 0; c < pitch; ++c) img[(size_t)r * pitch + c] = (uint8_t)(c < w ? (r * 7 + c * 3 + 1) : 0xAB);
  uint8_t *d_img, *d_out; int* d_st;
  cudaMalloc(&d_img, img.size()); cudaMalloc(&d_out, 8192); cudaMalloc(&d_st, 8);
  cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemset(d_st, 0, 8);
  Maps* maps = new Maps(); memset(maps, 0, sizeof *maps);
  cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h}, gstr[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)BW, (cuuint32_t)BH}, es[2] = {1, 1};
  CUresult r = encode(&maps->m[3], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("mode %d box %dx%d img %dx%d pitch %d at (%d,%d): encode -> %d  ", mode, BW, BH, w, h, pitch, x, y, (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  if (mode == 0) probe_array<<<1, 128>>>(*maps, 3, BW * BH, x, y, d_out, d_st);
  else if (mode == 1) probe_single<<<1, 128>>>(maps->m[3], BW * BH, x, y, d_out, d_st);
  else {
    CUtensorMap* d_map; cudaMalloc(&d_map, sizeof(CUtensorMap)); cudaMemcpy(d_map, &maps->m[3], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    probe_global<<<1, 128>>>(d_map, BW * BH, x, y, d_out, d_st);
  }
  cudaError_t e = cudaDeviceSynchronize();
  int st[2] = {0, 0}; std::vector<uint8_t> out(8192);
  cudaMemcpy(st, d_st, 8, cudaMemcpyDeviceToHost); cudaMemcpy(out.data(), d_out, BW * BH, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int rr = 0; rr < BH; ++rr) for (int cc = 0; cc < BW; ++cc) {
    const int gx = x + cc, gy = y + rr;
    const uint8_t want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? (uint8_t)(gy * 7 + gx * 3 + 1) : 0;
    bad += out[rr * BW + cc] != want;
  }
  printf("sync: %s, completed %d after %d polls, mismatching bytes %d\n", cudaGetErrorString(e), st[0], st[1], bad);
  return bad != 0 || !st[0];
}
