// gslam_b200/csrc/remap.cu — frame undistortion: the bilinear LUT remap of GSLAM::Undistorter::undistort
// (GSLAM/core/Undistorter.h:271-348; tables built by UndistorterImpl::prepareReMap :120-203).  SURVEY.md section 8f-4.
//
// The table (four source pixel indices + four float weights per output pixel, remapX < 0 = outside the input image) is built by the
// REFERENCE's own prepareReMap in the plugin (camera models stay the reference's); this file applies it: one thread per output
// pixel, p_out = p[i0]*c0 + p[i1]*c1 + p[i2]*c2 + p[i3]*c3 in float with the reference's left-to-right order, no FMA contraction,
// truncation to uchar -- bit-identical to the reference wherever the reference is defined.  Two places where the reference reads or
// leaves undefined memory are given a defined value here (documented in DESIGN.md, excluded from the parity comparison):
//   * taps whose index lies beyond the input image (last row / last column: the reference reads past its buffer) contribute 0;
//   * multi-channel images: output pixels outside the input image are 0 (the reference leaves them uninitialised) and the pixel
//     test is remapX > 0 as in the reference's multi-channel branch (:318), remapX >= 0 in the 1-channel branch (:297).
#include "common.cuh"

struct gb_remap {
  int w_in = 0, h_in = 0, w_out = 0, h_out = 0;
  int32_t* d_idx = nullptr;   // [w_out*h_out][4]
  float* d_coef = nullptr;    // [w_out*h_out][4]
  float* d_x = nullptr;       // [w_out*h_out] remapX
  uint8_t* d_src = nullptr;   // staging of the frame (grow-only)
  uint8_t* d_dst = nullptr;
  size_t src_cap = 0, dst_cap = 0;
};

namespace {

__global__ void __launch_bounds__(256) remap_kernel(const int4* __restrict__ idx, const float4* __restrict__ coef, const float* __restrict__ rx, int n_out,
                                                    int n_in, int channels, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const float xx = rx[i];
  const bool inside = channels == 1 ? !(xx < 0.f) : (xx > 0.f);
  if (!inside) {
    for (int j = 0; j < channels; ++j) dst[(size_t)i * channels + j] = 0;
    return;
  }
  const int4 t = idx[i];
  const float4 c = coef[i];
  for (int j = 0; j < channels; ++j) {
    const float p0 = t.x < n_in ? (float)src[(size_t)t.x * channels + j] : 0.f, p1 = t.y < n_in ? (float)src[(size_t)t.y * channels + j] : 0.f;
    const float p2 = t.z < n_in ? (float)src[(size_t)t.z * channels + j] : 0.f, p3 = t.w < n_in ? (float)src[(size_t)t.w * channels + j] : 0.f;
    const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0, c.x), __fmul_rn(p1, c.y)), __fmul_rn(p2, c.z)), __fmul_rn(p3, c.w));
    dst[(size_t)i * channels + j] = (uint8_t)__float2int_rz(v);
  }
}

}  // namespace

extern "C" {

int gb_remap_create(gb_ctx* ctx, int w_in, int h_in, int w_out, int h_out, const int32_t* idx4, const float* coef4, const float* remap_x, gb_remap** out) {
  if (!ctx || !out || w_in < 1 || h_in < 1 || w_out < 1 || h_out < 1 || !idx4 || !coef4 || !remap_x) return GB_ERR_INVALID;
  *out = nullptr;
  CtxLock lk(ctx);
  const size_t n = (size_t)w_out * h_out;
  for (size_t k = 0; k < 4 * n; ++k)
    if (idx4[k] < 0) { gb_set_error(ctx, "gb_remap_create: negative source index at entry %zu", k); return GB_ERR_INVALID; }
  gb_remap* m = new gb_remap();
  m->w_in = w_in; m->h_in = h_in; m->w_out = w_out; m->h_out = h_out;
  cudaError_t e = cudaMalloc((void**)&m->d_idx, n * 16);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_coef, n * 16);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_x, n * 4);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_idx, idx4, n * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_coef, coef4, n * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_x, remap_x, n * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    gb_set_error(ctx, "gb_remap_create -> %s", cudaGetErrorString(e));
    cudaFree(m->d_idx); cudaFree(m->d_coef); cudaFree(m->d_x);
    delete m;
    return GB_ERR_CUDA;
  }
  *out = m;
  return GB_OK;
}

int gb_remap_destroy(gb_ctx* ctx, gb_remap* m) {
  if (!m) return GB_OK;
  if (ctx) { CtxLock lk(ctx); cudaStreamSynchronize(ctx->stream); }
  cudaFree(m->d_idx); cudaFree(m->d_coef); cudaFree(m->d_x); cudaFree(m->d_src); cudaFree(m->d_dst);
  delete m;
  return GB_OK;
}

int gb_remap_apply(gb_ctx* ctx, gb_remap* m, const uint8_t* src, int channels, uint8_t* dst) {
  if (!ctx || !m || !src || !dst || (channels != 1 && channels != 3)) return GB_ERR_INVALID;
  CtxLock lk(ctx);
  const size_t n_in = (size_t)m->w_in * m->h_in, n_out = (size_t)m->w_out * m->h_out, b_in = n_in * channels, b_out = n_out * channels;
  GB_CHECK(gb_dev_realloc(ctx, (void**)&m->d_src, &m->src_cap, b_in));
  GB_CHECK(gb_dev_realloc(ctx, (void**)&m->d_dst, &m->dst_cap, b_out));
  GB_CHECK(gb_stage_reserve(ctx, ctx->h_stage_off + b_in + b_out + 1024));
  uint8_t* hs = (uint8_t*)gb_stage_alloc(ctx, b_in);
  uint8_t* hd = (uint8_t*)gb_stage_alloc(ctx, b_out);
  memcpy(hs, src, b_in);
  GB_CUDA(ctx, cudaMemcpyAsync(m->d_src, hs, b_in, cudaMemcpyHostToDevice, ctx->stream));
  remap_kernel<<<gb_div_up((int)n_out, 256), 256, 0, ctx->stream>>>((const int4*)m->d_idx, (const float4*)m->d_coef, m->d_x, (int)n_out, (int)n_in, channels,
                                                                   m->d_src, m->d_dst);
  GB_LAUNCH_CHECK(ctx);
  GB_CUDA(ctx, cudaMemcpyAsync(hd, m->d_dst, b_out, cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(dst, hd, b_out);
  return GB_OK;
}

}  // extern "C"
