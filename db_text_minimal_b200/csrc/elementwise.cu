// elementwise.cu -- memory-bound glue kernels: layout conversion, BatchNorm statistics / apply / backward,
// ReLU + residual, max-pool, FPN nearest-upsample add / concat.  All activations are NHWC bf16 with the channel
// count a multiple of 8, so every thread moves 16-byte vectors (8 channels) and a warp covers contiguous memory.
#include "common.cuh"
#include "elementwise.h"

namespace dbb {

// ---------------------------------------------------------------------------------------------
// NCHW float32 <-> NHWC bf16   (module boundaries keep the reference's NCHW float32 layout)
// ---------------------------------------------------------------------------------------------
// tile transpose through shared memory: 32 pixels x 32 channels per step
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int c, int64_t hw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* xb = x + (int64_t)n * c * hw;
  for (int j = ty; j < 32; j += 8) {
    const int cc = c0 + j; const int64_t pp = p0 + tx;
    tile[j][tx] = (cc < c && pp < hw) ? xb[(int64_t)cc * hw + pp] : 0.f;
  }
  __syncthreads();
  bf16* yb = y + (int64_t)n * hw * c;
  for (int j = ty; j < 32; j += 8) {
    const int64_t pp = p0 + j; const int cc = c0 + tx;
    if (pp < hw && cc < c) yb[pp * c + cc] = __float2bfloat16_rn(tile[tx][j]);
  }
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16* __restrict__ x, float* __restrict__ y, int c, int64_t hw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bf16* xb = x + (int64_t)n * hw * c;
  for (int j = ty; j < 32; j += 8) {
    const int64_t pp = p0 + j; const int cc = c0 + tx;
    tile[j][tx] = (pp < hw && cc < c) ? __bfloat162float(xb[pp * c + cc]) : 0.f;
  }
  __syncthreads();
  float* yb = y + (int64_t)n * c * hw;
  for (int j = ty; j < 32; j += 8) {
    const int cc = c0 + j; const int64_t pp = p0 + tx;
    if (cc < c && pp < hw) yb[(int64_t)cc * hw + pp] = tile[tx][j];
  }
}

int nchw_f32_to_nhwc_bf16(const float* x, bf16* y, int n, int c, int64_t hw, cudaStream_t s) {
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nchw_to_nhwc_kernel<<<grid, 256, 0, s>>>(x, y, c, hw);
  DBB_CHECK_LAUNCH("nchw_to_nhwc");
  return DBB_OK;
}
int nhwc_bf16_to_nchw_f32(const bf16* x, float* y, int n, int c, int64_t hw, cudaStream_t s) {
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  nhwc_to_nchw_kernel<<<grid, 256, 0, s>>>(x, y, c, hw);
  DBB_CHECK_LAUNCH("nhwc_to_nchw");
  return DBB_OK;
}

}  // namespace dbb

using namespace dbb;

extern "C" int dbb_nchw_f32_to_nhwc_bf16(const float* x, void* y, int64_t n, int c, int64_t h, int64_t w, void* stream) {
  if (!x || !y || n <= 0 || c <= 0 || h <= 0 || w <= 0) return set_error(DBB_EINVAL, "nchw_f32_to_nhwc_bf16: bad argument");
  return nchw_f32_to_nhwc_bf16(x, (bf16*)y, (int)n, c, h * w, (cudaStream_t)stream);
}
extern "C" int dbb_nhwc_bf16_to_nchw_f32(const void* x, float* y, int64_t n, int c, int64_t h, int64_t w, void* stream) {
  if (!x || !y || n <= 0 || c <= 0 || h <= 0 || w <= 0) return set_error(DBB_EINVAL, "nhwc_bf16_to_nchw_f32: bad argument");
  return nhwc_bf16_to_nchw_f32((const bf16*)x, y, (int)n, c, h * w, (cudaStream_t)stream);
}
