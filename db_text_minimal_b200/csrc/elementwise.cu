// elementwise.cu -- memory-bound glue kernels: layout conversion, BatchNorm statistics / apply / backward,
// ReLU + residual, max-pool, FPN nearest-upsample add / concat.  All activations are NHWC bf16 with the channel
// count a multiple of 8, so every thread moves 16-byte vectors (8 channels) and a warp covers contiguous memory.
#include "common.cuh"
#include "elementwise.h"
#include "bn_fin.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace dbb {

// ---------------------------------------------------------------------------------------------
// NCHW float32 <-> NHWC bf16   (module boundaries keep the reference's NCHW float32 layout)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float act2f(float v) { return v; }
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
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int c, int64_t hw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const T* xb = x + (int64_t)n * hw * c;
  for (int j = ty; j < 32; j += 8) {
    const int64_t pp = p0 + j; const int cc = c0 + tx;
    tile[j][tx] = (pp < hw && cc < c) ? act2f(xb[pp * c + cc]) : 0.f;
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
  DBB_LAUNCH("nchw_to_nhwc", s, nchw_to_nhwc_kernel<<<grid, 256, 0, s>>>(x, y, c, hw));
  return DBB_OK;
}
template <typename T>
int nhwc_to_nchw_f32(const T* x, float* y, int n, int c, int64_t hw, cudaStream_t s) {
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
  DBB_LAUNCH("nhwc_to_nchw", s, nhwc_to_nchw_kernel<T><<<grid, 256, 0, s>>>(x, y, c, hw));
  return DBB_OK;
}
template int nhwc_to_nchw_f32<bf16>(const bf16*, float*, int, int, int64_t, cudaStream_t);
template int nhwc_to_nchw_f32<float>(const float*, float*, int, int, int64_t, cudaStream_t);
int nhwc_bf16_to_nchw_f32(const bf16* x, float* y, int n, int c, int64_t hw, cudaStream_t s) { return nhwc_to_nchw_f32<bf16>(x, y, n, c, hw, s); }


// ---------------------------------------------------------------------------------------------
// per-channel machinery: a thread owns one 8-channel group (16 B) and walks pixels
// ---------------------------------------------------------------------------------------------
constexpr int EW_THREADS = 256;

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r;
  r.v[0] = bf16lo(u.x); r.v[1] = bf16hi(u.x); r.v[2] = bf16lo(u.y); r.v[3] = bf16hi(u.y);
  r.v[4] = bf16lo(u.z); r.v[5] = bf16hi(u.z); r.v[6] = bf16lo(u.w); r.v[7] = bf16hi(u.w);
  return r;
}
__device__ __forceinline__ F8 ld8s(const bf16* p) {   // streaming variant (touched once)
  const uint4 u = ldg_stream(reinterpret_cast<const uint4*>(p));
  F8 r;
  r.v[0] = bf16lo(u.x); r.v[1] = bf16hi(u.x); r.v[2] = bf16lo(u.y); r.v[3] = bf16hi(u.y);
  r.v[4] = bf16lo(u.z); r.v[5] = bf16hi(u.z); r.v[6] = bf16lo(u.w); r.v[7] = bf16hi(u.w);
  return r;
}
__device__ __forceinline__ void st8(bf16* p, const F8& a) {
  uint4 u;
  u.x = pack_bf16(a.v[0], a.v[1]); u.y = pack_bf16(a.v[2], a.v[3]); u.z = pack_bf16(a.v[4], a.v[5]); u.w = pack_bf16(a.v[6], a.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// fp32-parity storage (DESIGN.md "fp32 mode"): the same kernels instantiated on float activations
__device__ __forceinline__ F8 ld8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  return F8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ F8 ld8s(const float* p) {
  const float4 a = ldg_stream(reinterpret_cast<const float4*>(p)), b = ldg_stream(reinterpret_cast<const float4*>(p + 4));
  return F8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void st8(float* p, const F8& a) {
  *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
// value as it will be stored in an activation tensor of type T
template <typename T> __device__ __forceinline__ float round_act(float v);
template <> __device__ __forceinline__ float round_act<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
template <> __device__ __forceinline__ float round_act<float>(float v) { return v; }
__device__ __forceinline__ F8 ldf8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  return F8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}

// Blocks of a per-channel reduction: every block ends with 2*c global fp64 atomics, so a block should own at least
// ~64 KB of the tensor (small wide tensors were spending most of their time in those atomics).
static int ew_blocks(int64_t P, int c) {
  int64_t want = (P * c * 2 + 65535) / 65536;
  if (want < 1) want = 1;
  if (want > BN_MAX_BLOCKS) want = BN_MAX_BLOCKS;
  return (int)want;
}

// block-level reduction of per-thread 8-channel accumulators over the pixel lanes; writes [2][C] (or [1][C])
template <int NACC>
__device__ __forceinline__ void reduce_groups_store(float (&acc)[NACC][8], int c, float* out /* [NACC][c] */) {
  __shared__ float sh[EW_THREADS * 8];
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int lanes = EW_THREADS / groups;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[(pl * groups + g) * 8 + j] = acc[a][j];
    __syncthreads();
    // thread t < c sums channel t over the pixel lanes
    for (int ch = threadIdx.x; ch < c; ch += EW_THREADS) {
      const int gg = ch / 8, jj = ch % 8;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += sh[(l * groups + gg) * 8 + jj];
      out[a * c + ch] = s;
    }
  }
}

__global__ void __launch_bounds__(EW_THREADS) bn_stats_kernel(const bf16* __restrict__ z, int64_t P, int c, float* __restrict__ partials) {
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = EW_THREADS / groups;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < P; p += (int64_t)gridDim.x * lanes) {
    const F8 x = ld8(z + p * c + g * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] += x.v[j]; acc[1][j] += x.v[j] * x.v[j]; }
  }
  reduce_groups_store<2>(acc, c, partials + (size_t)blockIdx.x * 2 * c);
}

// one warp per channel: lanes stride over the per-block partials (fixed order -> bitwise reproducible)
__device__ __forceinline__ void warp_sum_partials(const float* __restrict__ partials, int nblk, int c, int ch, double& s, double& q) {
  const int lane = threadIdx.x & 31;
  s = 0.0; q = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += (double)partials[(size_t)b * 2 * c + ch];
    q += (double)partials[(size_t)b * 2 * c + c + ch];
  }
  s = warp_sum(s); q = warp_sum(q);
}

__global__ void bn_finalize_train_kernel(const float* __restrict__ partials, int nblk, int c, int coff, int cn, double count,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                         float* __restrict__ stats4) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= cn) return;
  const int ch = coff + i;
  double s, q;
  warp_sum_partials(partials, nblk, c, ch, s, q);
  if ((threadIdx.x & 31) != 0) return;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + (double)eps);
  stats4[ch] = (float)((double)gamma[i] * invstd);
  stats4[c + ch] = (float)((double)beta[i] - mean * (double)gamma[i] * invstd);
  stats4[2 * c + ch] = (float)mean;
  stats4[3 * c + ch] = (float)invstd;
  if (rmean) {   // running stats: unbiased variance, momentum 0.1 (torch.nn.BatchNorm2d)
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    rmean[i] = (float)((1.0 - momentum) * (double)rmean[i] + momentum * mean);
    rvar[i] = (float)((1.0 - momentum) * (double)rvar[i] + momentum * unb);
  }
}

__global__ void bn_finalize_eval_kernel(int c, int coff, int cn, const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ rmean, const float* __restrict__ rvar, float eps,
                                        float* __restrict__ stats4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cn) return;
  const int ch = coff + i;
  const float invstd = 1.f / sqrtf(rvar[i] + eps);
  stats4[ch] = gamma[i] * invstd;
  stats4[c + ch] = beta[i] - rmean[i] * gamma[i] * invstd;
  stats4[2 * c + ch] = rmean[i];
  stats4[3 * c + ch] = invstd;
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) bn_apply_kernel(const T* __restrict__ z, int64_t P, int c, const float* __restrict__ stats4,
                                                              const T* __restrict__ res, int relu, T* __restrict__ out,
                                                              int out_ctotal, int out_coff, int rev) {
  const int groups = c / 8;
  const int lg = 31 - __clz(groups);            // groups is a power of two (check_c): shifts instead of 64-bit divides
  const int64_t total = P * groups;
  const int g = threadIdx.x & (groups - 1);     // the grid stride is a multiple of groups: one channel group per thread
  const F8 sc = ldf8(stats4 + g * 8), sh = ldf8(stats4 + c + g * 8);
  const int64_t step = (int64_t)gridDim.x * EW_THREADS;
  auto one = [&](const F8& x, const F8* r, int64_t p) {
    F8 y;
#pragma unroll
    for (int j = 0; j < 8; ++j) y.v[j] = fmaf(x.v[j], sc.v[j], sh.v[j]);
    if (r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y.v[j] += r->v[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y.v[j] = fmaxf(y.v[j], 0.f);
    }
    st8(out + p * out_ctotal + out_coff + g * 8, y);
  };
  int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x;
  for (; i + step < total; i += 2 * step) {      // two items in flight per thread
    const int64_t p0 = rev ? P - 1 - (i >> lg) : (i >> lg), p1 = rev ? P - 1 - ((i + step) >> lg) : ((i + step) >> lg);
    const F8 x0 = ld8s(z + p0 * c + g * 8), x1 = ld8s(z + p1 * c + g * 8);
    if (res) {
      const F8 r0 = ld8s(res + p0 * c + g * 8), r1 = ld8s(res + p1 * c + g * 8);
      one(x0, &r0, p0); one(x1, &r1, p1);
    } else {
      one(x0, nullptr, p0); one(x1, nullptr, p1);
    }
  }
  if (i < total) {
    const int64_t p0 = rev ? P - 1 - (i >> lg) : (i >> lg);
    const F8 x0 = ld8s(z + p0 * c + g * 8);
    if (res) { const F8 r0 = ld8s(res + p0 * c + g * 8); one(x0, &r0, p0); }
    else one(x0, nullptr, p0);
  }
}

__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_reduce_kernel(const bf16* __restrict__ dout, int dout_ctotal, int dout_coff, const bf16* __restrict__ mask_src,
                     int mask_ctotal, int mask_coff, const bf16* __restrict__ z, int64_t P, int c,
                     const float* __restrict__ stats4, float* __restrict__ partials) {
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = EW_THREADS / groups;
  const F8 mean = ldf8(stats4 + 2 * c + g * 8), inv = ldf8(stats4 + 3 * c + g * 8);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < P; p += (int64_t)gridDim.x * lanes) {
    F8 dy = ld8(dout + p * dout_ctotal + dout_coff + g * 8);
    if (mask_src) {
      const F8 m = ld8(mask_src + p * mask_ctotal + mask_coff + g * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy.v[j] = m.v[j] > 0.f ? dy.v[j] : 0.f;
    }
    const F8 x = ld8(z + p * c + g * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] += dy.v[j]; acc[1][j] += dy.v[j] * (x.v[j] - mean.v[j]) * inv.v[j]; }
  }
  reduce_groups_store<2>(acc, c, partials + (size_t)blockIdx.x * 2 * c);
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nblk, int c, int coff, int cn, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ stats4,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef3) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= cn) return;
  const int ch = coff + i;
  double s, q;
  warp_sum_partials(partials, nblk, c, ch, s, q);
  if ((threadIdx.x & 31) != 0) return;
  if (dgamma) dgamma[i] = (float)q;
  if (dbeta) dbeta[i] = (float)s;
  coef3[ch] = gamma[i] * stats4[3 * c + ch];
  coef3[c + ch] = (float)(s / count);
  coef3[2 * c + ch] = (float)(q / count);
}

template <int MASK, typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply_kernel(const T* __restrict__ dout, int dout_ctotal, int dout_coff, const T* __restrict__ mask_src,
                    int mask_ctotal, int mask_coff, const T* __restrict__ z, int64_t P, int c,
                    const float* __restrict__ stats4, const float* __restrict__ coef3, T* __restrict__ dz,
                    T* __restrict__ dsum, int rev) {
  const int groups = c / 8;
  const int lg = 31 - __clz(groups);
  const int64_t total = P * groups;
  // groups is a power of two dividing the grid stride: a thread stays on one channel group -> per-channel constants hoisted
  const int g = threadIdx.x & (groups - 1);
  const F8 mean = ldf8(stats4 + 2 * c + g * 8), inv = ldf8(stats4 + 3 * c + g * 8);
  const F8 a = ldf8(coef3 + g * 8), c1 = ldf8(coef3 + c + g * 8), c2 = ldf8(coef3 + 2 * c + g * 8);
  F8 sc, sh;
  if (MASK == 2) { sc = ldf8(stats4 + g * 8); sh = ldf8(stats4 + c + g * 8); }
  auto one = [&](F8 dy, const F8& m, const F8& x, int64_t p) {
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MASK == 1) dy.v[j] = m.v[j] > 0.f ? dy.v[j] : 0.f;
      if (MASK == 2) dy.v[j] = fmaf(x.v[j], sc.v[j], sh.v[j]) > 0.f ? dy.v[j] : 0.f;
      o.v[j] = a.v[j] * (dy.v[j] - c1.v[j] - (x.v[j] - mean.v[j]) * inv.v[j] * c2.v[j]);
    }
    st8(dz + p * c + g * 8, o);
    if (dsum) st8(dsum + p * c + g * 8, dy);
  };
  const int64_t step = (int64_t)gridDim.x * EW_THREADS;
  int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x;
  for (; i + step < total; i += 2 * step) {      // two items in flight per thread
    const int64_t p0 = rev ? P - 1 - (i >> lg) : (i >> lg), p1 = rev ? P - 1 - ((i + step) >> lg) : ((i + step) >> lg);
    const F8 dy0 = ld8s(dout + p0 * dout_ctotal + dout_coff + g * 8), dy1 = ld8s(dout + p1 * dout_ctotal + dout_coff + g * 8);
    F8 m0, m1;
    if (MASK == 1) { m0 = ld8s(mask_src + p0 * mask_ctotal + mask_coff + g * 8); m1 = ld8s(mask_src + p1 * mask_ctotal + mask_coff + g * 8); }
    const F8 x0 = ld8s(z + p0 * c + g * 8), x1 = ld8s(z + p1 * c + g * 8);
    one(dy0, m0, x0, p0); one(dy1, m1, x1, p1);
  }
  if (i < total) {
    const int64_t p0 = rev ? P - 1 - (i >> lg) : (i >> lg);
    const F8 dy0 = ld8s(dout + p0 * dout_ctotal + dout_coff + g * 8);
    F8 m0;
    if (MASK == 1) m0 = ld8s(mask_src + p0 * mask_ctotal + mask_coff + g * 8);
    const F8 x0 = ld8s(z + p0 * c + g * 8);
    one(dy0, m0, x0, p0);
  }
}

__global__ void __launch_bounds__(EW_THREADS) colsum_kernel(const bf16* __restrict__ x, int64_t P, int c, float* __restrict__ partials) {
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = EW_THREADS / groups;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < P; p += (int64_t)gridDim.x * lanes) {
    const F8 v = ld8(x + p * c + g * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += v.v[j];
  }
  reduce_groups_store<1>(acc, c, partials + (size_t)blockIdx.x * c);
}
__global__ void colsum_finalize_kernel(const float* __restrict__ partials, int nblk, int c, float* __restrict__ out) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= c) return;
  double s = 0.0;
  for (int b = threadIdx.x & 31; b < nblk; b += 32) s += (double)partials[(size_t)b * c + ch];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[ch] = (float)s;
}

// Grid of a grid-stride kernel = one resident wave: SMs x (blocks that fit per SM for THIS kernel's register use).  Launching
// more (the old fixed 16 blocks/SM) meant several waves of short-lived blocks; launching a non-multiple (592 blocks on 444
// slots) cost a second wave for a third of the work.
template <typename K>
static int resident_blocks(K kernel) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, EW_THREADS, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; }
  return DBB_NUM_SMS * occ;
}
#define DBB_RESIDENT(kernel) ([]() -> int { static const int v = resident_blocks(kernel); return v; }())
// The apply passes walk their tensors from the END: the producer (the convolution that wrote z, the reduce pass that just read
// dout and z) touched the end last, so that is the part still resident in the 126 MB L2.
static int ew_reverse() { static const int v = getenv("DBB_NO_REVERSE") ? 0 : 1; return v; }
static int fit_grid(int64_t blocks_wanted, int resident) {
  if (blocks_wanted < 1) blocks_wanted = 1;
  return (int)(blocks_wanted < resident ? blocks_wanted : resident);
}

static const char* shaped(const char* base, int64_t P, int c, int extra = -1) {
  if (!prof_enabled()) return base;
  char tmp[96];
  if (extra >= 0) snprintf(tmp, sizeof(tmp), "%s_p%lld_c%d_m%d", base, (long long)P, c, extra);
  else snprintf(tmp, sizeof(tmp), "%s_p%lld_c%d", base, (long long)P, c);
  return prof_label(tmp);
}
static int check_c(int c) { return (c % 8 == 0 && c >= 8 && c <= 2048 && EW_THREADS % (c / 8) == 0 && ((c / 8) & (c / 8 - 1)) == 0) ? 0 : 1; }


// ---------------------------------------------------------------------------------------------
// fused "reduce + finalize": blocks add their per-channel sums to global fp64 accumulators (atomics), the last block to
// arrive (ticket counter) computes the per-channel results and leaves accumulators + counter zeroed for the next layer.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool last_block_arrives(unsigned* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// per-thread accumulator type of the per-channel reductions: float on the bf16 product path; double in the fp32-parity mode
// (whose reductions must not add rounding noise of their own: its results are compared with the reference at 1e-4 / 1e-3)
template <typename T> struct AccOf { typedef float type; };
template <> struct AccOf<float> { typedef double type; };

template <int NACC, typename A>
__device__ __forceinline__ void reduce_groups_atomic(A (&acc)[NACC][8], int c, double* gacc /* [NACC][c] */) {
  __shared__ A sh[EW_THREADS * 8];
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int lanes = EW_THREADS / groups;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[(pl * groups + g) * 8 + j] = acc[a][j];
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += EW_THREADS) {
      const int gg = ch / 8, jj = ch % 8;
      A s = 0;
      for (int l = 0; l < lanes; ++l) s += sh[(l * groups + gg) * 8 + jj];
      atomicAdd(&gacc[a * c + ch], (double)s);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_stats_fin_kernel(const T* __restrict__ z, int64_t P, int c, double* __restrict__ gacc, unsigned* __restrict__ counter, BnFin fin) {
  typedef typename AccOf<T>::type A;
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = EW_THREADS / groups;
  A acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0; acc[1][j] = 0; }
  for (int64_t p = (int64_t)blockIdx.x * lanes + pl; p < P; p += (int64_t)gridDim.x * lanes) {
    const F8 x = ld8(z + p * c + g * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] += (A)x.v[j]; acc[1][j] += (A)x.v[j] * (A)x.v[j]; }
  }
  reduce_groups_atomic<2, A>(acc, c, gacc);
  if (!last_block_arrives(counter)) return;
  bn_finalize_channels(fin, c, (double)P, gacc, threadIdx.x, EW_THREADS);
  if (threadIdx.x == 0) *counter = 0u;
}

// MASK: 0 = dy = dout, 1 = dy = dout * (mask_src > 0), 2 = dy = dout * (z*scale + shift > 0)  (the layer's own ReLU output
// re-derived from z: saves reading the activation tensor)
template <int MASK, typename T>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_reduce_fin_kernel(const T* __restrict__ dout, int dout_ctotal, int dout_coff, const T* __restrict__ mask_src,
                         int mask_ctotal, int mask_coff, const T* __restrict__ z, int64_t P, int c,
                         const float* __restrict__ stats4, double* __restrict__ gacc, unsigned* __restrict__ counter, BnBwdFin fin) {
  const int groups = c / 8;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, lanes = EW_THREADS / groups;
  const F8 mean = ldf8(stats4 + 2 * c + g * 8), inv = ldf8(stats4 + 3 * c + g * 8);
  F8 sc, sh;
  if (MASK == 2) { sc = ldf8(stats4 + g * 8); sh = ldf8(stats4 + c + g * 8); }
  typedef typename AccOf<T>::type A;
  A acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0; acc[1][j] = 0; }
  const int64_t step = (int64_t)gridDim.x * lanes;
  auto body = [&](F8 dy, const F8& m, const F8& x) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MASK == 1) dy.v[j] = m.v[j] > 0.f ? dy.v[j] : 0.f;
      if (MASK == 2) dy.v[j] = fmaf(x.v[j], sc.v[j], sh.v[j]) > 0.f ? dy.v[j] : 0.f;
      acc[0][j] += (A)dy.v[j]; acc[1][j] += (A)dy.v[j] * (A)((x.v[j] - mean.v[j]) * inv.v[j]);
    }
  };
  int64_t p = (int64_t)blockIdx.x * lanes + pl;
  for (; p + step < P; p += 2 * step) {      // two pixels in flight per thread
    const int64_t p2 = p + step;
    const F8 dy0 = ld8(dout + p * dout_ctotal + dout_coff + g * 8), dy1 = ld8(dout + p2 * dout_ctotal + dout_coff + g * 8);
    F8 m0, m1;
    if (MASK == 1) { m0 = ld8(mask_src + p * mask_ctotal + mask_coff + g * 8); m1 = ld8(mask_src + p2 * mask_ctotal + mask_coff + g * 8); }
    const F8 x0 = ld8(z + p * c + g * 8), x1 = ld8(z + p2 * c + g * 8);
    body(dy0, m0, x0); body(dy1, m1, x1);
  }
  if (p < P) {
    const F8 dy0 = ld8(dout + p * dout_ctotal + dout_coff + g * 8);
    F8 m0;
    if (MASK == 1) m0 = ld8(mask_src + p * mask_ctotal + mask_coff + g * 8);
    const F8 x0 = ld8(z + p * c + g * 8);
    body(dy0, m0, x0);
  }
  reduce_groups_atomic<2, A>(acc, c, gacc);
  if (!last_block_arrives(counter)) return;
  const double count = (double)P;
  for (int sg = 0; sg < fin.nseg; ++sg) {
    const BnBwdFinSeg& S = fin.seg[sg];
    for (int i = threadIdx.x; i < S.cn; i += EW_THREADS) {
      const int ch = S.coff + i;
      const double s = __ldcg(&gacc[ch]), q = __ldcg(&gacc[c + ch]);
      gacc[ch] = 0.0; gacc[c + ch] = 0.0;
      if (S.dgamma) S.dgamma[i] = (float)q;
      if (S.dbeta) S.dbeta[i] = (float)s;
      fin.coef3[ch] = S.gamma[i] * stats4[3 * c + ch];
      fin.coef3[c + ch] = (float)(s / count);
      fin.coef3[2 * c + ch] = (float)(q / count);
    }
  }
  if (threadIdx.x == 0) *counter = 0u;
}

template <typename T>
int bn_stats_finalize(const T* z, int64_t P, int c, const BnFin& fin, double* gacc, unsigned* counter, cudaStream_t s) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bn_stats: channel count");
  DBB_LAUNCH(shaped("bn_stats_fin", P, c), s, bn_stats_fin_kernel<T><<<fit_grid(ew_blocks(P, c), DBB_RESIDENT(bn_stats_fin_kernel<T>)), EW_THREADS, 0, s>>>(z, P, c, gacc, counter, fin));
  return DBB_OK;
}
template int bn_stats_finalize<bf16>(const bf16*, int64_t, int, const BnFin&, double*, unsigned*, cudaStream_t);
template int bn_stats_finalize<float>(const float*, int64_t, int, const BnFin&, double*, unsigned*, cudaStream_t);
template <typename T>
int bn_bwd_reduce_finalize(const T* dout, int dout_ctotal, int dout_coff, const ND<T>* mask_src, int mask_ctotal, int mask_coff,
                           const ND<T>* z, int64_t P, int c, const float* stats4, const BnBwdFin& fin, double* gacc,
                           unsigned* counter, cudaStream_t s, int mask_self) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bn_bwd_reduce: channel count");
  const int want = ew_blocks(P, c);
  if (mask_self) DBB_LAUNCH(shaped("bn_bwd_reduce_fin", P, c, 2), s, bn_bwd_reduce_fin_kernel<2, T><<<fit_grid(want, DBB_RESIDENT((bn_bwd_reduce_fin_kernel<2, T>))), EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, nullptr, 0, 0, z, P, c, stats4, gacc, counter, fin));
  else if (mask_src) DBB_LAUNCH(shaped("bn_bwd_reduce_fin", P, c, 1), s, bn_bwd_reduce_fin_kernel<1, T><<<fit_grid(want, DBB_RESIDENT((bn_bwd_reduce_fin_kernel<1, T>))), EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, mask_src, mask_ctotal, mask_coff, z, P, c, stats4, gacc, counter, fin));
  else DBB_LAUNCH(shaped("bn_bwd_reduce_fin", P, c, 0), s, bn_bwd_reduce_fin_kernel<0, T><<<fit_grid(want, DBB_RESIDENT((bn_bwd_reduce_fin_kernel<0, T>))), EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, nullptr, 0, 0, z, P, c, stats4, gacc, counter, fin));
  return DBB_OK;
}
template int bn_bwd_reduce_finalize<bf16>(const bf16*, int, int, const bf16*, int, int, const bf16*, int64_t, int, const float*, const BnBwdFin&, double*, unsigned*, cudaStream_t, int);
template int bn_bwd_reduce_finalize<float>(const float*, int, int, const float*, int, int, const float*, int64_t, int, const float*, const BnBwdFin&, double*, unsigned*, cudaStream_t, int);

int bn_stats(const bf16* z, int64_t P, int c, float* partials, int* nblk, cudaStream_t s) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bn_stats: channel count");
  *nblk = ew_blocks(P, c);
  DBB_LAUNCH("bn_stats", s, bn_stats_kernel<<<*nblk, EW_THREADS, 0, s>>>(z, P, c, partials));
  return DBB_OK;
}
int bn_finalize_train(const float* partials, int nblk, int c, int coff, int cn, int64_t count, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps, float* stats4, cudaStream_t s) {
  DBB_LAUNCH("bn_finalize_train", s, bn_finalize_train_kernel<<<(cn + 7) / 8, 256, 0, s>>>(partials, nblk, c, coff, cn, (double)count, gamma, beta, running_mean, running_var, momentum, eps, stats4));
  return DBB_OK;
}
int bn_finalize_eval(int c, int coff, int cn, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     float eps, float* stats4, cudaStream_t s) {
  DBB_LAUNCH("bn_finalize_eval", s, bn_finalize_eval_kernel<<<(cn + 127) / 128, 128, 0, s>>>(c, coff, cn, gamma, beta, running_mean, running_var, eps, stats4));
  return DBB_OK;
}
static int stream_grid(int64_t total) {
  int64_t g = (total + EW_THREADS - 1) / EW_THREADS;
  if (g > DBB_NUM_SMS * 16) g = DBB_NUM_SMS * 16;
  return (int)(g < 1 ? 1 : g);
}
template <typename T>
int bn_apply(const T* z, int64_t P, int c, const float* stats4, const ND<T>* res, int relu, ND<T>* out, int out_ctotal,
             int out_coff, cudaStream_t s) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bn_apply: channel count");
  DBB_LAUNCH(shaped("bn_apply", P, c, res ? 1 : 0), s, bn_apply_kernel<T><<<fit_grid((P * (c / 8) + 2 * EW_THREADS - 1) / (2 * EW_THREADS), DBB_RESIDENT(bn_apply_kernel<T>)), EW_THREADS, 0, s>>>(z, P, c, stats4, res, relu, out, out_ctotal, out_coff, ew_reverse()));
  return DBB_OK;
}
template int bn_apply<bf16>(const bf16*, int64_t, int, const float*, const bf16*, int, bf16*, int, int, cudaStream_t);
template int bn_apply<float>(const float*, int64_t, int, const float*, const float*, int, float*, int, int, cudaStream_t);
int bn_bwd_reduce(const bf16* dout, int dout_ctotal, int dout_coff, const bf16* mask_src, int mask_ctotal, int mask_coff,
                  const bf16* z, int64_t P, int c, const float* stats4, float* partials, int* nblk, cudaStream_t s) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bn_bwd_reduce: channel count");
  *nblk = ew_blocks(P, c);
  DBB_LAUNCH("bn_bwd_reduce", s, bn_bwd_reduce_kernel<<<*nblk, EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, mask_src, mask_ctotal, mask_coff, z, P, c, stats4, partials));
  return DBB_OK;
}
int bn_bwd_finalize(const float* partials, int nblk, int c, int coff, int cn, int64_t count, const float* gamma, const float* stats4,
                    float* dgamma, float* dbeta, float* coef3, cudaStream_t s) {
  DBB_LAUNCH("bn_bwd_finalize", s, bn_bwd_finalize_kernel<<<(cn + 7) / 8, 256, 0, s>>>(partials, nblk, c, coff, cn, (double)count, gamma, stats4, dgamma, dbeta, coef3));
  return DBB_OK;
}
template <typename T>
int bn_bwd_apply(const T* dout, int dout_ctotal, int dout_coff, const ND<T>* mask_src, int mask_ctotal, int mask_coff,
                 const ND<T>* z, int64_t P, int c, const float* stats4, const float* coef3, ND<T>* dz, ND<T>* dsum,
                 cudaStream_t s, int mask_self) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bn_bwd_apply: channel count");
  const int64_t want = (P * (c / 8) + 2 * EW_THREADS - 1) / (2 * EW_THREADS);
  if (mask_self) DBB_LAUNCH(shaped("bn_bwd_apply", P, c, 2), s, bn_bwd_apply_kernel<2, T><<<fit_grid(want, DBB_RESIDENT((bn_bwd_apply_kernel<2, T>))), EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, nullptr, 0, 0, z, P, c, stats4, coef3, dz, dsum, ew_reverse()));
  else if (mask_src) DBB_LAUNCH(shaped("bn_bwd_apply", P, c, dsum ? 11 : 1), s, bn_bwd_apply_kernel<1, T><<<fit_grid(want, DBB_RESIDENT((bn_bwd_apply_kernel<1, T>))), EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, mask_src, mask_ctotal, mask_coff, z, P, c, stats4, coef3, dz, dsum, ew_reverse()));
  else DBB_LAUNCH(shaped("bn_bwd_apply", P, c, 0), s, bn_bwd_apply_kernel<0, T><<<fit_grid(want, DBB_RESIDENT((bn_bwd_apply_kernel<0, T>))), EW_THREADS, 0, s>>>(dout, dout_ctotal, dout_coff, nullptr, 0, 0, z, P, c, stats4, coef3, dz, dsum, ew_reverse()));
  return DBB_OK;
}
template int bn_bwd_apply<bf16>(const bf16*, int, int, const bf16*, int, int, const bf16*, int64_t, int, const float*, const float*, bf16*, bf16*, cudaStream_t, int);
template int bn_bwd_apply<float>(const float*, int, int, const float*, int, int, const float*, int64_t, int, const float*, const float*, float*, float*, cudaStream_t, int);
int bias_grad(const bf16* dz, int64_t P, int c, float* partials, float* dbias, cudaStream_t s) {
  if (check_c(c)) return set_error(DBB_EUNSUPPORTED, "bias_grad: channel count");
  const int nblk = ew_blocks(P, c);
  DBB_LAUNCH("colsum", s, colsum_kernel<<<nblk, EW_THREADS, 0, s>>>(dz, P, c, partials));
  DBB_LAUNCH("colsum_finalize", s, colsum_finalize_kernel<<<(c + 7) / 8, 256, 0, s>>>(partials, nblk, c, dbias));
  return DBB_OK;
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, padding 1); argmax = first maximum in (kh, kw) scan order (ATen semantics)
// ---------------------------------------------------------------------------------------------
// bn_stats4 != nullptr: x is the RAW convolution output and the pooled operand is bf16(relu(x*scale + shift)) computed on the
// fly -- the stem's BatchNorm-apply pass and its 210 MB activation tensor disappear (the backward re-derives the ReLU mask
// from x, see bn_bwd_*_kernel<2>).  The values are rounded to bf16 before the comparison, so results are bit-identical
// to pooling a materialised activation tensor.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) maxpool_fwd_kernel(const T* __restrict__ x, int n, int h, int w, int c, int oh, int ow,
                                                                 T* __restrict__ y, uint8_t* __restrict__ argmax,
                                                                 const float* __restrict__ bn_stats4) {
  const int groups = c / 8;
  const int64_t total = (int64_t)n * oh * ow * groups;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
    const unsigned u = (unsigned)i;                       // total < 2^32 (checked on the host): 32-bit divides
    const int g = (int)(u % (unsigned)groups); unsigned t = u / (unsigned)groups;
    const int x0 = (int)(t % (unsigned)ow); t /= (unsigned)ow; const int y0 = (int)(t % (unsigned)oh); const int b = (int)(t / (unsigned)oh);
    F8 best; uint8_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best.v[j] = -INFINITY; bi[j] = 0; }
    F8 bsc, bsh;
    if (bn_stats4) { bsc = ldf8(bn_stats4 + g * 8); bsh = ldf8(bn_stats4 + c + g * 8); }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = 2 * y0 - 1 + kh;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int xx = 2 * x0 - 1 + kw;
        if (xx < 0 || xx >= w) continue;
        F8 v = ld8(x + (((int64_t)b * h + yy) * w + xx) * c + g * 8);
        if (bn_stats4) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v.v[j] = round_act<T>(fmaxf(fmaf(v.v[j], bsc.v[j], bsh.v[j]), 0.f));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) if (v.v[j] > best.v[j]) { best.v[j] = v.v[j]; bi[j] = (uint8_t)(kh * 3 + kw); }
      }
    }
    const int64_t o = (((int64_t)b * oh + y0) * ow + x0) * c + g * 8;
    st8(y + o, best);
    if (argmax) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(argmax + o) = pk;
    }
  }
}
// One thread per 2x2 block of dx pixels {2a, 2a+1} x {2b, 2b+1} and 8-channel group: the block is covered by exactly the
// four pooling windows (a,b), (a,b+1), (a+1,b), (a+1,b+1), whose gradients and argmax bytes are loaded up front
// (one window load per dx pixel instead of 2.25, all in flight together).
__device__ __forceinline__ void mp_take(F8& acc, const F8& d, const uint2& pk, unsigned k) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned idx = ((j < 4 ? pk.x : pk.y) >> ((j & 3) * 8)) & 0xffu;
    acc.v[j] += (idx == k) ? d.v[j] : 0.f;
  }
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ argmax, int n, int h,
                                                                 int w, int c, int oh, int ow, T* __restrict__ dx) {
  const int groups = c / 8;
  const int hb = (h + 1) / 2, wb = (w + 1) / 2;
  const int64_t total = (int64_t)n * hb * wb * groups;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
    const unsigned u = (unsigned)i;
    const int g = (int)(u % (unsigned)groups); unsigned t = u / (unsigned)groups;
    const int b = (int)(t % (unsigned)wb); t /= (unsigned)wb; const int a = (int)(t % (unsigned)hb); const int img = (int)(t / (unsigned)hb);
    F8 d[4]; uint2 pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int y0 = a + (q >> 1), x0 = b + (q & 1);
      if (y0 < oh && x0 < ow) {
        const int64_t o = (((int64_t)img * oh + y0) * ow + x0) * c + g * 8;
        pk[q] = *reinterpret_cast<const uint2*>(argmax + o);
        d[q] = ld8(dy + o);
      } else {
        pk[q] = make_uint2(0xffffffffu, 0xffffffffu);     // matches no tap
#pragma unroll
        for (int j = 0; j < 8; ++j) d[q].v[j] = 0.f;
      }
    }
    F8 o00, o01, o10, o11;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o00.v[j] = 0.f; o01.v[j] = 0.f; o10.v[j] = 0.f; o11.v[j] = 0.f; }
    // tap index k = (yy - (2*y0 - 1)) * 3 + (xx - (2*x0 - 1)) of dx pixel (yy, xx) inside window (y0, x0)
    mp_take(o00, d[0], pk[0], 4);
    mp_take(o01, d[0], pk[0], 5); mp_take(o01, d[1], pk[1], 3);
    mp_take(o10, d[0], pk[0], 7); mp_take(o10, d[2], pk[2], 1);
    mp_take(o11, d[0], pk[0], 8); mp_take(o11, d[1], pk[1], 6); mp_take(o11, d[2], pk[2], 2); mp_take(o11, d[3], pk[3], 0);
    const int yy = 2 * a, xx = 2 * b;
    T* base = dx + (((int64_t)img * h + yy) * w + xx) * c + g * 8;
    st8(base, o00);
    if (xx + 1 < w) st8(base + c, o01);
    if (yy + 1 < h) {
      st8(base + (int64_t)w * c, o10);
      if (xx + 1 < w) st8(base + (int64_t)w * c + c, o11);
    }
  }
}
static int fits32(int64_t total, const char* who) { return total < ((int64_t)1 << 32) ? 0 : set_error(DBB_EUNSUPPORTED, who); }

template <typename T>
int maxpool_fwd(const T* x, int n, int h, int w, int c, ND<T>* y, uint8_t* argmax, cudaStream_t s, const float* bn_stats4) {
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  if (fits32((int64_t)n * h * w * (c / 8), "maxpool: tensor too large for 32-bit indexing")) return DBB_EUNSUPPORTED;
  DBB_LAUNCH("maxpool_fwd", s, maxpool_fwd_kernel<T><<<fit_grid(((int64_t)n * oh * ow * (c / 8) + EW_THREADS - 1) / EW_THREADS, DBB_RESIDENT(maxpool_fwd_kernel<T>)), EW_THREADS, 0, s>>>(x, n, h, w, c, oh, ow, y, argmax, bn_stats4));
  return DBB_OK;
}
template int maxpool_fwd<bf16>(const bf16*, int, int, int, int, bf16*, uint8_t*, cudaStream_t, const float*);
template int maxpool_fwd<float>(const float*, int, int, int, int, float*, uint8_t*, cudaStream_t, const float*);
template <typename T>
int maxpool_bwd(const T* dy, const uint8_t* argmax, int n, int h, int w, int c, ND<T>* dx, cudaStream_t s) {
  const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
  DBB_LAUNCH("maxpool_bwd", s, maxpool_bwd_kernel<T><<<fit_grid(((int64_t)n * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8) + EW_THREADS - 1) / EW_THREADS, DBB_RESIDENT(maxpool_bwd_kernel<T>)), EW_THREADS, 0, s>>>(dy, argmax, n, h, w, c, oh, ow, dx));
  return DBB_OK;
}
template int maxpool_bwd<bf16>(const bf16*, const uint8_t*, int, int, int, int, bf16*, cudaStream_t);
template int maxpool_bwd<float>(const float*, const uint8_t*, int, int, int, int, float*, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// nearest-neighbour upsampling, F.interpolate(mode='nearest'): src = min(floor(dst * fp32(in/out)), in-1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  const int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) upsample_fwd_kernel(const T* __restrict__ xs, int hs, int ws, float sch, float scw,
                                                                  const T* __restrict__ addend, int n, int h, int w, int c,
                                                                  T* __restrict__ dst, int dst_ctotal, int dst_coff) {
  const int groups = c / 8;
  const int64_t total = (int64_t)n * h * w * groups;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
    const unsigned u = (unsigned)i;
    const int g = (int)(u % (unsigned)groups); unsigned t = u / (unsigned)groups;
    const int xx = (int)(t % (unsigned)w); t /= (unsigned)w; const int yy = (int)(t % (unsigned)h); const int b = (int)(t / (unsigned)h);
    const int sy = nearest_src(yy, sch, hs), sx = nearest_src(xx, scw, ws);
    F8 v = ld8(xs + (((int64_t)b * hs + sy) * ws + sx) * c + g * 8);
    const int64_t pix = ((int64_t)b * h + yy) * w + xx;
    if (addend) {
      const F8 a = ld8s(addend + pix * c + g * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) v.v[j] += a.v[j];
    }
    st8(dst + pix * dst_ctotal + dst_coff + g * 8, v);
  }
}
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) upsample_bwd_kernel(const T* __restrict__ d_big, int big_ctotal, int big_coff, int n, int h,
                                                                  int w, int c, float sch, float scw, T* __restrict__ d_xs,
                                                                  int hs, int ws, int accumulate) {
  const int groups = c / 8;
  const int64_t total = (int64_t)n * hs * ws * groups;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
    const unsigned u = (unsigned)i;
    const int g = (int)(u % (unsigned)groups); unsigned t = u / (unsigned)groups;
    const int sx = (int)(t % (unsigned)ws); t /= (unsigned)ws; const int sy = (int)(t % (unsigned)hs); const int b = (int)(t / (unsigned)hs);
    // destination rows/cols that read (sy, sx): a contiguous range (the map is monotone)
    int y0 = (int)ceilf((float)sy / sch) - 2; if (y0 < 0) y0 = 0;
    while (y0 < h && nearest_src(y0, sch, hs) < sy) ++y0;
    int y1 = y0; while (y1 < h && nearest_src(y1, sch, hs) == sy) ++y1;
    int x0 = (int)ceilf((float)sx / scw) - 2; if (x0 < 0) x0 = 0;
    while (x0 < w && nearest_src(x0, scw, ws) < sx) ++x0;
    int x1 = x0; while (x1 < w && nearest_src(x1, scw, ws) == sx) ++x1;
    F8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) {
        const F8 d = ld8(d_big + (((int64_t)b * h + yy) * w + xx) * big_ctotal + big_coff + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] += d.v[j];
      }
    T* o = d_xs + (((int64_t)b * hs + sy) * ws + sx) * c + g * 8;
    if (accumulate) {
      const F8 old = ld8(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc.v[j] += old.v[j];
    }
    st8(o, acc);
  }
}
// c == 64 and a large ratio (FPN p4 / p5 -> p2 resolution): one WARP per source pixel, lane = (row-split rs, channel group g);
// every lane sums the window rows y0 + rs, y0 + rs + 4, ... (128 contiguous bytes per row across the 8 groups) and the four
// row-splits are combined by a fixed shuffle tree.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) upsample_bwd_warp_kernel(const T* __restrict__ d_big, int big_ctotal, int big_coff, int n, int h,
                                                                       int w, float sch, float scw, T* __restrict__ d_xs,
                                                                       int hs, int ws, int accumulate) {
  const int lane = threadIdx.x & 31, g = lane & 7, rs = lane >> 3;
  const int total = n * hs * ws;
  for (int i = blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5); i < total; i += gridDim.x * (EW_THREADS / 32)) {
    unsigned t = (unsigned)i;
    const int sx = (int)(t % (unsigned)ws); t /= (unsigned)ws; const int sy = (int)(t % (unsigned)hs); const int b = (int)(t / (unsigned)hs);
    int y0 = (int)ceilf((float)sy / sch) - 2; if (y0 < 0) y0 = 0;
    while (y0 < h && nearest_src(y0, sch, hs) < sy) ++y0;
    int y1 = y0; while (y1 < h && nearest_src(y1, sch, hs) == sy) ++y1;
    int x0 = (int)ceilf((float)sx / scw) - 2; if (x0 < 0) x0 = 0;
    while (x0 < w && nearest_src(x0, scw, ws) < sx) ++x0;
    int x1 = x0; while (x1 < w && nearest_src(x1, scw, ws) == sx) ++x1;
    F8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    for (int yy = y0 + rs; yy < y1; yy += 4)
      for (int xx = x0; xx < x1; ++xx) {
        const F8 d = ld8(d_big + (((int64_t)b * h + yy) * w + xx) * big_ctotal + big_coff + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] += d.v[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc.v[j] += __shfl_xor_sync(0xffffffffu, acc.v[j], 8);
      acc.v[j] += __shfl_xor_sync(0xffffffffu, acc.v[j], 16);
    }
    if (rs == 0) {
      T* o = d_xs + (((int64_t)b * hs + sy) * ws + sx) * 64 + g * 8;
      if (accumulate) {
        const F8 old = ld8(o);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] += old.v[j];
      }
      st8(o, acc);
    }
  }
}
template <typename T>
int upsample_add_fwd(const T* xs, int hs, int ws, const ND<T>* y, int n, int h, int w, int c, ND<T>* out, cudaStream_t s) {
  DBB_LAUNCH("upsample_add_fwd", s, upsample_fwd_kernel<T><<<fit_grid(((int64_t)n * h * w * (c / 8) + EW_THREADS - 1) / EW_THREADS, DBB_RESIDENT(upsample_fwd_kernel<T>)), EW_THREADS, 0, s>>>(xs, hs, ws, (float)hs / (float)h, (float)ws / (float)w, y, n, h, w, c, out, c, 0));
  return DBB_OK;
}
template int upsample_add_fwd<bf16>(const bf16*, int, int, const bf16*, int, int, int, int, bf16*, cudaStream_t);
template int upsample_add_fwd<float>(const float*, int, int, const float*, int, int, int, int, float*, cudaStream_t);
template <typename T>
int upsample_into(const T* xs, int hs, int ws, int n, int h, int w, int c, ND<T>* dst, int dst_ctotal, int dst_coff, cudaStream_t s) {
  DBB_LAUNCH("upsample_into", s, upsample_fwd_kernel<T><<<fit_grid(((int64_t)n * h * w * (c / 8) + EW_THREADS - 1) / EW_THREADS, DBB_RESIDENT(upsample_fwd_kernel<T>)), EW_THREADS, 0, s>>>(xs, hs, ws, (float)hs / (float)h, (float)ws / (float)w, nullptr, n, h, w, c, dst, dst_ctotal, dst_coff));
  return DBB_OK;
}
template int upsample_into<bf16>(const bf16*, int, int, int, int, int, int, bf16*, int, int, cudaStream_t);
template int upsample_into<float>(const float*, int, int, int, int, int, int, float*, int, int, cudaStream_t);
template <typename T>
int upsample_bwd(const T* d_big, int big_ctotal, int big_coff, int n, int h, int w, int c, ND<T>* d_xs, int hs, int ws,
                 int accumulate, cudaStream_t s) {
  if (c == 64 && h >= 4 * hs && (int64_t)n * hs * ws < (1 << 30)) {
    int grid = (n * hs * ws + EW_THREADS / 32 - 1) / (EW_THREADS / 32);
    if (grid > DBB_NUM_SMS * 16) grid = DBB_NUM_SMS * 16;
    DBB_LAUNCH("upsample_bwd", s, upsample_bwd_warp_kernel<T><<<grid, EW_THREADS, 0, s>>>(d_big, big_ctotal, big_coff, n, h, w, (float)hs / (float)h, (float)ws / (float)w, d_xs, hs, ws, accumulate));
    return DBB_OK;
  }
  DBB_LAUNCH("upsample_bwd", s, upsample_bwd_kernel<T><<<stream_grid((int64_t)n * hs * ws * (c / 8)), EW_THREADS, 0, s>>>(d_big, big_ctotal, big_coff, n, h, w, c, (float)hs / (float)h, (float)ws / (float)w, d_xs, hs, ws, accumulate));
  return DBB_OK;
}
template int upsample_bwd<bf16>(const bf16*, int, int, int, int, int, int, bf16*, int, int, int, cudaStream_t);
template int upsample_bwd<float>(const float*, int, int, int, int, int, int, float*, int, int, int, cudaStream_t);

// ---------------------------------------------------------------------------------------------
// conv1 staging: NCHW float32 image -> zero-padded space-to-depth bf16 [n][hs+3][ws+3][16]
// channel = (py*2+px)*3 + c for the 2x2 sub-pixel (py,px); channels 12..15 are zero
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) image_to_s2d_kernel(const float* __restrict__ img, int n, int h, int w, int hs, int ws,
                                                                  bf16* __restrict__ out) {
  const int ph = hs + 3, pw = ws + 3;
  const int64_t total = (int64_t)n * ph * pw;
  for (int64_t i = (int64_t)blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * EW_THREADS) {
    const unsigned u = (unsigned)i;
    const int q = (int)(u % (unsigned)pw); const unsigned t = u / (unsigned)pw; const int r = (int)(t % (unsigned)ph); const int b = (int)(t / (unsigned)ph);
    const int si = r - 2, sj = q - 2;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    if (si >= 0 && si < hs && sj >= 0 && sj < ws) {
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          const int yy = 2 * si + py, xx = 2 * sj + px;
          if (yy < h && xx < w) {
#pragma unroll
            for (int c = 0; c < 3; ++c) v[(py * 2 + px) * 3 + c] = __ldg(img + (((int64_t)b * 3 + c) * h + yy) * w + xx);
          }
        }
    }
    uint4 o0, o1;
    o0.x = pack_bf16(v[0], v[1]); o0.y = pack_bf16(v[2], v[3]); o0.z = pack_bf16(v[4], v[5]); o0.w = pack_bf16(v[6], v[7]);
    o1.x = pack_bf16(v[8], v[9]); o1.y = pack_bf16(v[10], v[11]); o1.z = pack_bf16(v[12], v[13]); o1.w = pack_bf16(v[14], v[15]);
    uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
    dst[0] = o0; dst[1] = o1;
  }
}
int image_to_s2d(const float* img, int n, int h, int w, bf16* s2d, cudaStream_t s) {
  const int hs = (h + 1) / 2, ws = (w + 1) / 2;
  DBB_LAUNCH("image_to_s2d", s, image_to_s2d_kernel<<<stream_grid((int64_t)n * (hs + 3) * (ws + 3)), EW_THREADS, 0, s>>>(img, n, h, w, hs, ws, s2d));
  return DBB_OK;
}
__global__ void conv1_wgrad_unpack_kernel(const float* __restrict__ dw_s2d, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over 64*3*7*7
  if (i >= 64 * 147) return;
  const int kw = i % 7, kh = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
  // kh = 2*kh2 + py - 1  ->  py = (kh+1)&1, kh2 = (kh+1)>>1
  const int py = (kh + 1) & 1, kh2 = (kh + 1) >> 1, px = (kw + 1) & 1, kw2 = (kw + 1) >> 1;
  const int nidx = kw2 * 16 + (py * 2 + px) * 3 + c;
  dw[i] = dw_s2d[((int64_t)co * 64 + nidx) * 4 + kh2];
}
int conv1_wgrad_unpack(const float* dw_s2d, float* dw, cudaStream_t s) {
  DBB_LAUNCH("conv1_wgrad_unpack", s, conv1_wgrad_unpack_kernel<<<(64 * 147 + 255) / 256, 256, 0, s>>>(dw_s2d, dw));
  return DBB_OK;
}

// ---------------------------------------------------------------------------------------------
// bilinear resize, align_corners=True (src/models.py:43-46).  Identity when sizes match (multiples of 4).
// ATen: src = dst * (in-1)/(out-1); lambda in float32
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilin_coord(int d, int in, int out, int& i0, int& i1, float& l1) {
  const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  const float src = scale * (float)d;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const float* __restrict__ x, int nc, int hi, int wi, float* __restrict__ y, int ho, int wo) {
  const int64_t total = (int64_t)nc * ho * wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int ox = (int)(i % wo); int64_t t = i / wo; const int oy = (int)(t % ho); const int m = (int)(t / ho);
    int y0, y1, x0, x1; float ly, lx;
    bilin_coord(oy, hi, ho, y0, y1, ly);
    bilin_coord(ox, wi, wo, x0, x1, lx);
    const float* p = x + (int64_t)m * hi * wi;
    const float hy = 1.f - ly, hx = 1.f - lx;
    y[i] = hy * (hx * p[y0 * wi + x0] + lx * p[y0 * wi + x1]) + ly * (hx * p[y1 * wi + x0] + lx * p[y1 * wi + x1]);
  }
}
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ dy, int nc, int hi, int wi, float* __restrict__ dx, int ho, int wo) {
  const int64_t total = (int64_t)nc * ho * wo;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int ox = (int)(i % wo); int64_t t = i / wo; const int oy = (int)(t % ho); const int m = (int)(t / ho);
    int y0, y1, x0, x1; float ly, lx;
    bilin_coord(oy, hi, ho, y0, y1, ly);
    bilin_coord(ox, wi, wo, x0, x1, lx);
    float* p = dx + (int64_t)m * hi * wi;
    const float hy = 1.f - ly, hx = 1.f - lx, g = dy[i];
    atomicAdd(p + y0 * wi + x0, hy * hx * g); atomicAdd(p + y0 * wi + x1, hy * lx * g);
    atomicAdd(p + y1 * wi + x0, ly * hx * g); atomicAdd(p + y1 * wi + x1, ly * lx * g);
  }
}
int bilinear_fwd(const float* x, int nc, int hi, int wi, float* y, int ho, int wo, cudaStream_t s) {
  DBB_LAUNCH("bilinear_fwd", s, bilinear_fwd_kernel<<<stream_grid((int64_t)nc * ho * wo), 256, 0, s>>>(x, nc, hi, wi, y, ho, wo));
  return DBB_OK;
}
int bilinear_bwd(const float* dy, int nc, int hi, int wi, float* dx, int ho, int wo, cudaStream_t s) {
  DBB_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)nc * hi * wi, s));
  DBB_LAUNCH("bilinear_bwd", s, bilinear_bwd_kernel<<<stream_grid((int64_t)nc * ho * wo), 256, 0, s>>>(dy, nc, hi, wi, dx, ho, wo));
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
