// head_tail.cu -- the memory-bound tail of DBHead, fused:  BN-apply + ReLU -> ConvTranspose2d(64->1, k2, s2) x 2
// branches -> Sigmoid -> differentiable step B = 1/(1+exp(-k(P-T))) -> packed NCHW float32 store, and its backward.
//
// Replaces src/modules/segmentation_head.py:28-29 (BatchNorm2d, ReLU, ConvTranspose2d(64,1,2,2), Sigmoid of the
// `binarize` branch), :72-76 (same for `thresh`), :39-44 and :106-108 (step_function + cat) of the reference.
//
// Input  zt  : raw output of the two ConvTranspose2d(64,64,2,2) layers, NHWC bf16 (N, H2, W2, 128);
//              channels 0..63 = binarize branch, 64..127 = thresh branch (256 B per pixel).
// Output out : (N, 3|2, 2*H2, 2*W2) float32 = [P, T, (B)].
// HBM traffic per OUTPUT pixel: 64 B read (bf16) + 12 B written forward; the 64x320^2 ReLU activations and the
// pre-sigmoid logits never touch memory.  A CTA owns a run of 64 input pixels of one row: phase 1 gives 16 lanes to
// each pixel (one 16-byte vector per lane, warp-shuffle dot products), phase 2 re-maps threads to output columns so
// the P/T/B rows leave as fully coalesced 512-byte stores.
#include "common.cuh"
#include "head_tail.h"
#include <stdlib.h>
#include "tcgen05.cuh"

namespace dbb {

using ptx::mbar_init; using ptx::mbar_wait; using ptx::mbar_arrive_expect_tx; using ptx::fence_barrier_init; using ptx::smem_u32;

// Tile staging (PIPE kernels, used when W2 is a multiple of the 64-pixel tile so that a tile is 16 KB of contiguous
// NHWC memory): a ring of HT_STAGES shared-memory buffers filled by the bulk-copy engine (cp.async.bulk, the 1-D TMA
// path) and signalled through mbarriers, so that 64 KB per CTA are in flight while the warps compute on earlier tiles.
constexpr int HT_STAGES = 4;
constexpr int HT_TILE_BYTES = 64 * 256;
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
struct HtRing {
  uint8_t* buf;
  uint64_t* full;
  const bf16* zt;
  int tiles_per_row, h2, w2;
  int64_t ntiles;
  __device__ __forceinline__ void issue(int64_t tile, int stage) const {
    const unsigned row = (unsigned)tile / (unsigned)tiles_per_row;
    const unsigned tr = (unsigned)tile - row * (unsigned)tiles_per_row;
    const bf16* src = zt + ((int64_t)row * w2 + (int64_t)tr * 64) * 128;
    mbar_arrive_expect_tx(&full[stage], HT_TILE_BYTES);
    bulk_load(buf + stage * HT_TILE_BYTES, src, HT_TILE_BYTES, &full[stage]);
  }
  __device__ __forceinline__ void start() const {
    if (threadIdx.x == 0) {
      for (int s = 0; s < HT_STAGES; ++s) mbar_init(&full[s], 1);
      fence_barrier_init();
      for (int s = 0; s < HT_STAGES; ++s) {
        const int64_t t = blockIdx.x + (int64_t)s * gridDim.x;
        if (t < ntiles) issue(t, s);
      }
    }
    __syncthreads();
  }
  __device__ __forceinline__ const bf16* wait(int64_t k) const {
    const int stage = (int)(k % HT_STAGES);
    mbar_wait(&full[stage], (uint32_t)((k / HT_STAGES) & 1));
    return reinterpret_cast<const bf16*>(buf + stage * HT_TILE_BYTES);
  }
  // call after a __syncthreads() that follows the last read of iteration k's stage
  __device__ __forceinline__ void refill(int64_t k, int64_t tile) const {
    if (threadIdx.x == 0) {
      const int64_t nxt = tile + (int64_t)HT_STAGES * gridDim.x;
      if (nxt < ntiles) issue(nxt, (int)(k % HT_STAGES));
    }
  }
};

constexpr int HT_THREADS = 256;
constexpr int HT_TILE = 64;   // input pixels per tile

struct HtF8 { float v[8]; };
__device__ __forceinline__ HtF8 ht_ld8(const bf16* p) {
  const uint4 u = ldg_stream(reinterpret_cast<const uint4*>(p));
  HtF8 r;
  r.v[0] = bf16lo(u.x); r.v[1] = bf16hi(u.x); r.v[2] = bf16lo(u.y); r.v[3] = bf16hi(u.y);
  r.v[4] = bf16lo(u.z); r.v[5] = bf16hi(u.z); r.v[6] = bf16lo(u.w); r.v[7] = bf16hi(u.w);
  return r;
}

// fp32-parity storage: the same kernels on float activations (never staged: PIPE = false)
__device__ __forceinline__ HtF8 ht_ld8(const float* p) {
  const float4 a = ldg_stream(reinterpret_cast<const float4*>(p)), b = ldg_stream(reinterpret_cast<const float4*>(p + 4));
  return HtF8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ HtF8 ht_lds8(const float* p) { return ht_ld8(p); }
__device__ __forceinline__ void ht_st8(bf16* p, const float (&o)[8]) {
  uint4 u;
  u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
  stg_stream(reinterpret_cast<uint4*>(p), u);
}
__device__ __forceinline__ void ht_st8(float* p, const float (&o)[8]) {
  stg_stream(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
  stg_stream(reinterpret_cast<float4*>(p + 4), make_float4(o[4], o[5], o[6], o[7]));
}
__device__ __forceinline__ HtF8 ht_lds8(const bf16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  HtF8 r;
  r.v[0] = bf16lo(u.x); r.v[1] = bf16hi(u.x); r.v[2] = bf16lo(u.y); r.v[3] = bf16hi(u.y);
  r.v[4] = bf16lo(u.z); r.v[5] = bf16hi(u.z); r.v[6] = bf16lo(u.w); r.v[7] = bf16hi(u.w);
  return r;
}

// packed fp32x2 FMA (sm_100 FFMA2): two taps per instruction halve the FMA issue count of these issue-bound kernels
__device__ __forceinline__ float2 fma2(float a, float2 w, float2 acc) { return __ffma2_rn(make_float2(a, a), w, acc); }

// per-lane constants: lane l = (branch = l>>3, channel group = l&7) owns channels ch0 = branch*64 + 8*(l&7) ..
struct LaneConst {
  float sc[8], sh[8], w[8][4];
};
__device__ __forceinline__ void load_lane_const(LaneConst& L, int l16, const float* __restrict__ stats4,
                                                const float* __restrict__ w2b, const float* __restrict__ w2t) {
  const int branch = l16 >> 3, cg = l16 & 7, ch0 = branch * 64 + cg * 8;
  const float* w2 = branch ? w2t : w2b;   // ConvTranspose2d weight (64, 1, 2, 2): [c][a*2+b]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    L.sc[j] = stats4[ch0 + j];
    L.sh[j] = stats4[128 + ch0 + j];
#pragma unroll
    for (int t = 0; t < 4; ++t) L.w[j][t] = w2[(cg * 8 + j) * 4 + t];
  }
}

// 32-bit index math on purpose: a 64-bit divide is a ~100-instruction software loop and these kernels are issue-bound
__device__ __forceinline__ void tile_coords(int64_t tile64, int tiles_per_row, int h2, int& n, int& i, int& j0) {
  const unsigned tile = (unsigned)tile64;
  const unsigned row = tile / (unsigned)tiles_per_row;
  const unsigned tr = tile - row * (unsigned)tiles_per_row;
  n = (int)(row / (unsigned)h2); i = (int)(row - (unsigned)n * (unsigned)h2); j0 = (int)tr * HT_TILE;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <bool PIPE, typename T>
__global__ void __launch_bounds__(HT_THREADS)
head_tail_fwd_kernel(const T* __restrict__ zt, int n_img, int h2, int w2, const float* __restrict__ stats4,
                     const float* __restrict__ w2b, const float* __restrict__ w2t, const float* __restrict__ b2b,
                     const float* __restrict__ b2t, float k, int out_c, float* __restrict__ out) {
  __shared__ float zs[2][4][HT_TILE];   // [branch][tap][pixel] pre-sigmoid logits
  __shared__ uint64_t full_bar[HT_STAGES];
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int l16 = threadIdx.x & 15, pslot = threadIdx.x >> 4;   // 16 pixel slots per pass
  LaneConst L;
  load_lane_const(L, l16, stats4, w2b, w2t);
  const float bias_b = b2b[0], bias_t = b2t[0];
  const int tiles_per_row = (w2 + HT_TILE - 1) / HT_TILE;
  const int64_t ntiles = (int64_t)n_img * h2 * tiles_per_row;
  const int H = 2 * h2, W = 2 * w2;
  const HtRing ring{ring_smem, full_bar, reinterpret_cast<const bf16*>(zt), tiles_per_row, h2, w2, ntiles};
  if (PIPE) ring.start();
  int64_t kk = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kk) {
    int n, i, j0;
    tile_coords(tile, tiles_per_row, h2, n, i, j0);
    const T* zrow = PIPE ? reinterpret_cast<const T*>(ring.wait(kk)) - (int64_t)j0 * 128 : zt + (((int64_t)n * h2 + i) * w2) * 128;
    // ---- phase 1: BN + ReLU + 4 dot products per branch
#pragma unroll
    for (int pass = 0; pass < HT_TILE / 16; ++pass) {
      const int px = pass * 16 + pslot;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (j0 + px < w2) {
        const T* zp = zrow + (int64_t)(j0 + px) * 128 + l16 * 8;
        const HtF8 z = PIPE ? ht_lds8(zp) : ht_ld8(zp);
        float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = fmaxf(fmaf(z.v[j], L.sc[j], L.sh[j]), 0.f);
          a01 = fma2(a, make_float2(L.w[j][0], L.w[j][1]), a01);
          a23 = fma2(a, make_float2(L.w[j][2], L.w[j][3]), a23);
        }
        acc[0] = a01.x; acc[1] = a01.y; acc[2] = a23.x; acc[3] = a23.y;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 1);
        acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 2);
        acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], 4);
      }
      if ((l16 & 7) == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) zs[l16 >> 3][t][px] = acc[t];
      }
    }
    __syncthreads();
    if (PIPE) ring.refill(kk, tile);
    // ---- phase 2: sigmoid, step, coalesced store.  thread -> (a = row parity, col = output column in the tile)
    {
      const int a = threadIdx.x >> 7, col = threadIdx.x & 127;
      const int px = col >> 1, tap = a * 2 + (col & 1);
      if (j0 + px < w2) {
        const float zb = zs[0][tap][px] + bias_b, ztv = zs[1][tap][px] + bias_t;
        const float Pm = 1.f / (1.f + expf(-zb));
        const float Tm = 1.f / (1.f + expf(-ztv));
        const int64_t plane = (int64_t)H * W;
        float* o = out + ((int64_t)n * out_c) * plane + (int64_t)(2 * i + a) * W + 2 * j0 + col;
        o[0] = Pm;
        o[plane] = Tm;
        if (out_c == 3) o[2 * plane] = 1.f / (1.f + expf(-k * (Pm - Tm)));
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward
//   dz_b[tap] = (dP + s) P (1-P),  dz_t[tap] = (dT - s) T (1-T),  s = dB k B^2 e,  e = exp(-k (P - T))
//   da[c] = sum_tap dz[tap] w2[c][tap];  dy[c] = da[c] [a[c] > 0]   (gradient entering the BatchNorm of ConvT1)
// pass "reduce": per-channel sums for BN backward (sum dy, sum dy*xhat), dW2[c][tap] = sum a[c] dz[tap], db2 = sum dz
// pass "apply" : d_zt = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat))  -> NHWC bf16
// ---------------------------------------------------------------------------------------------
// Phase A is software-pipelined: the six maps of tile k+1 are requested (registers) before the channel work of tile k, so
// their global-memory latency hides behind it instead of stalling every tile.
struct PhaseAIn { float P, T, B, dP, dT, dB; int valid; };
__device__ __forceinline__ void bwd_phase_a_load(PhaseAIn& r, const float* __restrict__ out, const float* __restrict__ dout,
                                                 int n, int i, int j0, int w2, int H, int W) {
  const int a = threadIdx.x >> 7, col = threadIdx.x & 127;
  const int px = col >> 1;
  r.valid = (j0 + px < w2);
  if (r.valid) {
    const int64_t plane = (int64_t)H * W;
    const int64_t o = ((int64_t)n * 3) * plane + (int64_t)(2 * i + a) * W + 2 * j0 + col;
    r.P = __ldg(out + o); r.T = __ldg(out + o + plane); r.B = __ldg(out + o + 2 * plane);
    r.dP = __ldg(dout + o); r.dT = __ldg(dout + o + plane); r.dB = __ldg(dout + o + 2 * plane);
  }
}
__device__ __forceinline__ void bwd_phase_a_store(float (&dzs)[2][4][HT_TILE], const PhaseAIn& r, float k) {
  const int a = threadIdx.x >> 7, col = threadIdx.x & 127;
  const int px = col >> 1, tap = a * 2 + (col & 1);
  float dzb = 0.f, dzt = 0.f;
  if (r.valid) {
    const float e = expf(-k * (r.P - r.T));
    const float s = r.dB * k * r.B * r.B * e;
    dzb = (r.dP + s) * r.P * (1.f - r.P);
    dzt = (r.dT - s) * r.T * (1.f - r.T);
  }
  dzs[0][tap][px] = dzb;
  dzs[1][tap][px] = dzt;
}

constexpr int HT_NACC = 6;   // per channel: dW2[4 taps], sum dy, sum dy*xhat

template <typename T> struct HtAcc { typedef float type; };
template <> struct HtAcc<float> { typedef double type; };      // fp32-parity mode: reductions add no rounding noise of their own
template <bool PIPE, typename T>
__global__ void __launch_bounds__(HT_THREADS, sizeof(T) == 2 ? 2 : 1)
head_tail_bwd_reduce_kernel(const T* __restrict__ zt, int n_img, int h2, int w2, const float* __restrict__ stats4,
                            const float* __restrict__ w2b, const float* __restrict__ w2t, const float* __restrict__ out,
                            const float* __restrict__ dout, float k, float* __restrict__ partials /* [grid][128*6 + 2] */) {
  __shared__ float dzs[2][4][HT_TILE];
  __shared__ typename HtAcc<T>::type red[HT_THREADS][9];   // padded rows
  __shared__ uint64_t full_bar[HT_STAGES];
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int l16 = threadIdx.x & 15, pslot = threadIdx.x >> 4;
  LaneConst L;
  load_lane_const(L, l16, stats4, w2b, w2t);
  float mean[8], inv[8];
  {
    const int ch0 = (l16 >> 3) * 64 + (l16 & 7) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mean[j] = stats4[256 + ch0 + j]; inv[j] = stats4[384 + ch0 + j]; }
  }
  typedef typename HtAcc<T>::type A;
  constexpr bool WIDE = sizeof(A) == 8;
  A accW[8][4], accS[8], accQ[8], accB = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { accS[j] = 0; accQ[j] = 0; for (int t = 0; t < 4; ++t) accW[j][t] = 0; }
  const int tiles_per_row = (w2 + HT_TILE - 1) / HT_TILE;
  const int64_t ntiles = (int64_t)n_img * h2 * tiles_per_row;
  const int H = 2 * h2, W = 2 * w2;
  const HtRing ring{ring_smem, full_bar, reinterpret_cast<const bf16*>(zt), tiles_per_row, h2, w2, ntiles};
  if (PIPE) ring.start();
  int64_t kk = 0;
  PhaseAIn pa;
  pa.valid = 0;
  if ((int64_t)blockIdx.x < ntiles) {
    int n, i, j0;
    tile_coords(blockIdx.x, tiles_per_row, h2, n, i, j0);
    bwd_phase_a_load(pa, out, dout, n, i, j0, w2, H, W);
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kk) {
    int n, i, j0;
    tile_coords(tile, tiles_per_row, h2, n, i, j0);
    bwd_phase_a_store(dzs, pa, k);
    if (tile + gridDim.x < ntiles) {       // request the next tile's maps now; consumed at the top of the next iteration
      int n2, i2, j2;
      tile_coords(tile + gridDim.x, tiles_per_row, h2, n2, i2, j2);
      bwd_phase_a_load(pa, out, dout, n2, i2, j2, w2, H, W);
    }
    __syncthreads();
    const T* zrow = PIPE ? reinterpret_cast<const T*>(ring.wait(kk)) - (int64_t)j0 * 128 : zt + (((int64_t)n * h2 + i) * w2) * 128;
    const int br = l16 >> 3;
#pragma unroll
    for (int pass = 0; pass < HT_TILE / 16; ++pass) {
      const int px = pass * 16 + pslot;
      if (j0 + px < w2) {
        const T* zp = zrow + (int64_t)(j0 + px) * 128 + l16 * 8;
        const HtF8 z = PIPE ? ht_lds8(zp) : ht_ld8(zp);
        float dz[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) dz[t] = dzs[br][t][px];
        if ((l16 & 7) == 0) accB += (A)dz[0] + (A)dz[1] + (A)dz[2] + (A)dz[3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = fmaxf(fmaf(z.v[j], L.sc[j], L.sh[j]), 0.f);
          if (WIDE) {
#pragma unroll
            for (int t = 0; t < 4; ++t) accW[j][t] += (A)a * (A)dz[t];
          } else {
            const float2 w01 = fma2(a, make_float2(dz[0], dz[1]), make_float2((float)accW[j][0], (float)accW[j][1]));
            const float2 w23 = fma2(a, make_float2(dz[2], dz[3]), make_float2((float)accW[j][2], (float)accW[j][3]));
            accW[j][0] = w01.x; accW[j][1] = w01.y; accW[j][2] = w23.x; accW[j][3] = w23.y;
          }
          const float2 d2 = __ffma2_rn(make_float2(dz[0], dz[1]), make_float2(L.w[j][0], L.w[j][1]),
                                       __fmul2_rn(make_float2(dz[2], dz[3]), make_float2(L.w[j][2], L.w[j][3])));
          const float da = d2.x + d2.y;
          const float dy = a > 0.f ? da : 0.f;
          accS[j] += (A)dy;
          if (WIDE) accQ[j] += (A)dy * (A)((z.v[j] - mean[j]) * inv[j]);
          else accQ[j] = fmaf(dy, (z.v[j] - mean[j]) * inv[j], (float)accQ[j]);
        }
      }
    }
    __syncthreads();
    if (PIPE) ring.refill(kk, tile);
  }
  // ---- reduce over the 16 pixel slots that share a lane role; thread (l16, pslot)
  float* my = partials + (size_t)blockIdx.x * (128 * HT_NACC + 2);
  const int ch0 = (l16 >> 3) * 64 + (l16 & 7) * 8;
#pragma unroll
  for (int q = 0; q < HT_NACC; ++q) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = (q < 4) ? accW[j][q] : (q == 4 ? accS[j] : accQ[j]);
    __syncthreads();
    if (threadIdx.x < 128) {      // thread -> channel (l = t/8, j = t%8)
      const int l = threadIdx.x >> 3, j = threadIdx.x & 7;
      A s = 0;
#pragma unroll
      for (int ps = 0; ps < 16; ++ps) s += red[ps * 16 + l][j];
      const int ch = (l >> 3) * 64 + (l & 7) * 8 + j;
      my[ch * HT_NACC + q] = (float)s;
    }
  }
  (void)ch0;
  __syncthreads();
  red[threadIdx.x][0] = accB;
  __syncthreads();
  if (threadIdx.x < 2) {
    A s = 0;
    for (int ps = 0; ps < 16; ++ps) s += red[ps * 16 + threadIdx.x * 8][0];
    my[128 * HT_NACC + threadIdx.x] = (float)s;
  }
}

// ---------------------------------------------------------------------------------------------
// backward "reduce" on the tensor cores (bf16 activations, W2 a multiple of 64: the training shapes).
//
// The CUDA-core kernel above spends ~14 FMA-class instructions per (channel, input pixel) and is issue-bound at 0.30 of the
// HBM roofline (ncu, profiles/prof_mem_r01.md).  All of its per-channel sums derive from TWO pixel reductions,
//     M[c][t] = sum_px mask[c,px] dz[t,px]           Z[c][t] = sum_px mask[c,px] z[c,px] dz[t,px]
// (mask = [z*scale + shift > 0], dz = the four tap gradients of the pixel):
//     dW2[c][t] = scale_c Z + shift_c M      sum dy = sum_t w2[c][t] M      sum dy*xhat = invstd_c sum_t w2[c][t] (Z - mean_c M)
// and M, Z are GEMMs over the pixel index: A = mask resp. mask*z (16 channels x 16 pixels, bf16: mask is 0/1 and z IS bf16, so
// both are exact), B = dz (16 pixels x 8 columns = 4 taps x {hi, lo} bf16 split of the fp32 value: relative error 2^-17),
// fp32 accumulate -- mma.sync.m16n8k16.  Per 64-pixel tile a warp owns 16 channels and issues 4 k-steps of
// {ldmatrix.x4.trans from the 128B-swizzled TMA tile, 4 packed compare / and pairs, 2 MMAs}.  The ReLU mask is evaluated as
// a packed bf16 comparison against the per-channel boundary zb = the largest bf16 value with fmaf(zb, scale, shift) <= 0
// (found by bisection in the prologue; fmaf is monotone in z, so the mask equals the forward pass's for every bf16 z).
// ---------------------------------------------------------------------------------------------
constexpr int HT_MZ = 8;                 // per channel: M[4 taps], Z[4 taps]
constexpr int HT_DZ_PITCH = 72;          // words per (branch, tap) row of the packed dz tile: conflict-free B-fragment reads

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// ordered index <-> bf16 bit pattern (monotone in the represented value; NaNs excluded by the range searched)
__device__ __forceinline__ uint32_t bf16_from_ordered(int k) {      // k in [-0x7f80, 0x7f80]: -inf .. +inf
  return k >= 0 ? (uint32_t)k : (0x8000u | (uint32_t)(-k));
}
// largest bf16 value zb (as ordered index) with fmaf(zb, sc, sh) <= 0, for sc > 0 (callers pass |sc| and flip z's sign otherwise)
__device__ __forceinline__ int relu_boundary(float sc, float sh) {
  int lo = -0x7f80, hi = 0x7f80;         // invariant: f(lo) <= 0 is assumed (f(-inf) = -inf), f(hi) > 0 (f(+inf) = +inf)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    const float z = __uint_as_float(bf16_from_ordered(mid) << 16);
    if (fmaf(z, sc, sh) > 0.f) hi = mid; else lo = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(HT_THREADS, 3)
head_tail_bwd_reduce_mma_kernel(const __grid_constant__ CUtensorMap tmap, int n_img, int h2, int w2, const float* __restrict__ stats4,
                                const float* __restrict__ out, const float* __restrict__ dout, float k,
                                float* __restrict__ partials /* [grid][128*8 + 2] */) {
  extern __shared__ uint8_t mma_smem_raw[];                               // HT_STAGES x {2 halves x [64 px][128 B], 128B swizzle}
  uint8_t* ring_smem = mma_smem_raw + ((1024u - (smem_u32(mma_smem_raw) & 1023u)) & 1023u);      // swizzled TMA tiles: 1024 B
  __shared__ uint32_t dzp[2][4][HT_DZ_PITCH];                             // packed (hi | lo << 16) bf16 split of dz
  __shared__ uint64_t full_bar[HT_STAGES];
  __shared__ float redB[HT_THREADS / 32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int tiles_per_row = w2 / HT_TILE;
  const int64_t ntiles = (int64_t)n_img * h2 * tiles_per_row;
  const int H = 2 * h2, W = 2 * w2;
  // ---- per-thread channel constants: rows g and g + 8 of this warp's 16-channel slab
  uint32_t flip[2], zb2[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int ch = warp * 16 + g + 8 * r;
    const float sc = stats4[ch], sh = stats4[128 + ch];
    uint32_t f = 0, zb;
    if (sc > 0.f) zb = bf16_from_ordered(relu_boundary(sc, sh));
    else if (sc < 0.f) { f = 0x8000u; zb = bf16_from_ordered(relu_boundary(-sc, sh)); }   // mask = [-z > zb'] with f(-z') = (-sc) z' + sh
    else zb = sh > 0.f ? 0xff80u /* -inf: every z passes */ : 0x7f80u /* +inf: none */;
    flip[r] = f | (f << 16);
    zb2[r] = zb | (zb << 16);
  }
  auto issue = [&](int64_t tile, int stage) {
    int n, i, j0;
    tile_coords(tile, tiles_per_row, h2, n, i, j0);
    uint8_t* dst = ring_smem + stage * HT_TILE_BYTES;
    mbar_arrive_expect_tx(&full_bar[stage], HT_TILE_BYTES);
    ptx::tma_load_4d(dst, &tmap, &full_bar[stage], 0, j0, i, n);
    ptx::tma_load_4d(dst + HT_TILE_BYTES / 2, &tmap, &full_bar[stage], 64, j0, i, n);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < HT_STAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_barrier_init();
    for (int s = 0; s < HT_STAGES; ++s) {
      const int64_t tl = blockIdx.x + (int64_t)s * gridDim.x;
      if (tl < ntiles) issue(tl, s);
    }
  }
  __syncthreads();
  float accM[4] = {0.f, 0.f, 0.f, 0.f}, accZ[4] = {0.f, 0.f, 0.f, 0.f};
  float accBb = 0.f, accBt = 0.f;
  PhaseAIn pa;
  pa.valid = 0;
  if ((int64_t)blockIdx.x < ntiles) {
    int n, i, j0;
    tile_coords(blockIdx.x, tiles_per_row, h2, n, i, j0);
    bwd_phase_a_load(pa, out, dout, n, i, j0, w2, H, W);
  }
  const int br = warp >> 2;                                   // channels 0..63 = binarize branch, 64..127 = thresh
  // ldmatrix row address of this lane inside a half tile: matrix mi = lane / 8 -> (pixel block, channel chunk)
  const int mi = lane >> 3, mr = lane & 7;
  const int chunk = (warp & 3) * 2 + (mi & 1);                // 16-byte chunk (8 channels) inside the 128-byte row
  const int prow = (mi >> 1) * 8 + mr;                        // pixel row inside the 16-pixel k-step
  int64_t kk = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kk) {
    // ---- phase A: dz of this tile's 256 output pixels -> packed hi/lo bf16 in shared memory
    {
      const int a = threadIdx.x >> 7, col = threadIdx.x & 127;
      const int px = col >> 1, tap = a * 2 + (col & 1);
      float dzb = 0.f, dzt = 0.f;
      if (pa.valid) {
        const float e = expf(-k * (pa.P - pa.T));
        const float sgrad = pa.dB * k * pa.B * pa.B * e;
        dzb = (pa.dP + sgrad) * pa.P * (1.f - pa.P);
        dzt = (pa.dT - sgrad) * pa.T * (1.f - pa.T);
      }
      accBb += dzb; accBt += dzt;
      const __nv_bfloat16 hb = __float2bfloat16_rn(dzb), ht = __float2bfloat16_rn(dzt);
      const __nv_bfloat16 lb = __float2bfloat16_rn(dzb - __bfloat162float(hb)), lt = __float2bfloat16_rn(dzt - __bfloat162float(ht));
      dzp[0][tap][px] = (uint32_t)__bfloat16_as_ushort(hb) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
      dzp[1][tap][px] = (uint32_t)__bfloat16_as_ushort(ht) | ((uint32_t)__bfloat16_as_ushort(lt) << 16);
    }
    if (tile + gridDim.x < ntiles) {       // request the next tile's maps now; consumed at the top of the next iteration
      int n2, i2, j2;
      tile_coords(tile + gridDim.x, tiles_per_row, h2, n2, i2, j2);
      bwd_phase_a_load(pa, out, dout, n2, i2, j2, w2, H, W);
    }
    __syncthreads();
    const int stage = (int)(kk % HT_STAGES);
    mbar_wait(&full_bar[stage], (uint32_t)((kk / HT_STAGES) & 1));
    const uint32_t half_base = smem_u32(ring_smem + stage * HT_TILE_BYTES + br * (HT_TILE_BYTES / 2));
    const uint32_t* dzrow = &dzp[br][g & 3][0];
    const uint32_t sel = (g & 4) ? 0x7632u : 0x5410u;         // PRMT selector: lo halves (bytes 2,3 of each word) or hi halves (0,1)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int p = ks * 16 + prow;
      uint32_t a[4];
      ldmatrix_x4_trans(a, half_base + p * 128 + ((chunk ^ (p & 7)) << 4));
      // B fragment: column n = g -> (tap g & 3, part g >> 2), rows k = 2t, 2t+1 | 2t+8, 2t+9
      const int kx = ks * 16 + 2 * t;
      const uint32_t b0 = __byte_perm(dzrow[kx], dzrow[kx + 1], sel), b1 = __byte_perm(dzrow[kx + 8], dzrow[kx + 9], sel);
      uint32_t am[4], az[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = q & 1;                                  // a0, a2: channel row g; a1, a3: row g + 8
        const uint32_t zf = a[q] ^ flip[r];
        const __nv_bfloat162 zz = *reinterpret_cast<const __nv_bfloat162*>(&zf), bb = *reinterpret_cast<const __nv_bfloat162*>(&zb2[r]);
        const uint32_t m = __hgt2_mask(zz, bb);               // 0xffff per element where z (sign-adjusted) > boundary
        am[q] = m & 0x3f803f80u;                              // 1.0 / 0.0
        az[q] = a[q] & m;                                     // z / 0.0
      }
      mma_bf16_16816(accM, am, b0, b1);
      mma_bf16_16816(accZ, az, b0, b1);
    }
    __syncthreads();                                          // every warp is done with this stage and with dzp
    if (threadIdx.x == 0) {
      const int64_t nxt = tile + (int64_t)HT_STAGES * gridDim.x;
      if (nxt < ntiles) issue(nxt, stage);
    }
  }
  // ---- per-CTA partials.  D fragment: c0, c1 = (row g, cols 2t, 2t+1); c2, c3 = (row g + 8, ...); cols 0..3 = hi parts of
  // taps 0..3, cols 4..7 = lo parts: lanes t and t ^ 2 hold the two halves of the same taps
  float* my = partials + (size_t)blockIdx.x * (128 * HT_MZ + 2);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    accM[q] += __shfl_xor_sync(0xffffffffu, accM[q], 2);
    accZ[q] += __shfl_xor_sync(0xffffffffu, accZ[q], 2);
  }
  if (t < 2) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int ch = warp * 16 + g + 8 * r;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        my[ch * HT_MZ + 2 * t + e] = accM[2 * r + e];
        my[ch * HT_MZ + 4 + 2 * t + e] = accZ[2 * r + e];
      }
    }
  }
  accBb = warp_sum(accBb); accBt = warp_sum(accBt);
  if (lane == 0) { redB[warp][0] = accBb; redB[warp][1] = accBt; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float sB = 0.f;
    for (int wv = 0; wv < HT_THREADS / 32; ++wv) sB += redB[wv][threadIdx.x];
    my[128 * HT_MZ + threadIdx.x] = sB;
  }
}

// one warp per channel (128 warps) + one warp per bias; lanes stride over the per-block partials in a fixed order
// layout 0: per channel {dW2[4], sum dy, sum dy*xhat} (CUDA-core reduce); layout 1: per channel {M[4], Z[4]} (tensor-core reduce)
__global__ void head_tail_bwd_finalize_kernel(const float* __restrict__ partials, int nblk, int layout, const float* __restrict__ w2b,
                                              const float* __restrict__ w2t, double count,
                                              const float* __restrict__ gamma_b, const float* __restrict__ gamma_t,
                                              const float* __restrict__ stats4, float* __restrict__ dgamma_b,
                                              float* __restrict__ dbeta_b, float* __restrict__ dgamma_t,
                                              float* __restrict__ dbeta_t, float* __restrict__ coef3,
                                              float* __restrict__ dw2b, float* __restrict__ dw2t, float* __restrict__ db2b,
                                              float* __restrict__ db2t) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // 0..129
  const int lane = threadIdx.x & 31;
  const int per = layout ? HT_MZ : HT_NACC;
  const size_t stride = 128 * per + 2;
  if (ch >= 130) return;
  if (ch >= 128) {
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += (double)partials[(size_t)b * stride + 128 * per + (ch - 128)];
    s = warp_sum(s);
    if (lane == 0) ((ch - 128) ? db2t : db2b)[0] = (float)s;
    return;
  }
  double raw[HT_MZ];
#pragma unroll
  for (int q = 0; q < HT_MZ; ++q) raw[q] = 0.0;
  for (int b = lane; b < nblk; b += 32) {
#pragma unroll
    for (int q = 0; q < HT_MZ; ++q) if (q < per) raw[q] += (double)partials[(size_t)b * stride + ch * per + q];
  }
#pragma unroll
  for (int q = 0; q < HT_MZ; ++q) raw[q] = warp_sum(raw[q]);
  if (lane != 0) return;
  const int br = ch >> 6, c = ch & 63;
  double acc[HT_NACC];
  if (layout) {        // dW2 = scale Z + shift M;  sum dy = sum_t w2 M;  sum dy*xhat = invstd sum_t w2 (Z - mean M)
    const double sc = stats4[ch], sh = stats4[128 + ch], mean = stats4[256 + ch], inv = stats4[384 + ch];
    const float* w2 = br ? w2t : w2b;
    acc[4] = 0.0; acc[5] = 0.0;
    for (int t = 0; t < 4; ++t) {
      const double M = raw[t], Z = raw[4 + t], w = (double)w2[c * 4 + t];
      acc[t] = sc * Z + sh * M;
      acc[4] += w * M;
      acc[5] += w * (Z - mean * M);
    }
    acc[5] *= inv;
  } else {
    for (int q = 0; q < HT_NACC; ++q) acc[q] = raw[q];
  }
  float* dw2 = br ? dw2t : dw2b;
  for (int t = 0; t < 4; ++t) dw2[c * 4 + t] = (float)acc[t];
  (br ? dbeta_t : dbeta_b)[c] = (float)acc[4];
  (br ? dgamma_t : dgamma_b)[c] = (float)acc[5];
  const float g = (br ? gamma_t : gamma_b)[c];
  coef3[ch] = g * stats4[384 + ch];
  coef3[128 + ch] = (float)(acc[4] / count);
  coef3[256 + ch] = (float)(acc[5] / count);
}

template <bool PIPE, typename T>
__global__ void __launch_bounds__(HT_THREADS)
head_tail_bwd_apply_kernel(const T* __restrict__ zt, int n_img, int h2, int w2, const float* __restrict__ stats4,
                           const float* __restrict__ coef3, const float* __restrict__ w2b, const float* __restrict__ w2t,
                           const float* __restrict__ out, const float* __restrict__ dout, float k, T* __restrict__ d_zt) {
  __shared__ float dzs[2][4][HT_TILE];
  __shared__ uint64_t full_bar[HT_STAGES];
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const int l16 = threadIdx.x & 15, pslot = threadIdx.x >> 4;
  LaneConst L;
  load_lane_const(L, l16, stats4, w2b, w2t);
  float mean[8], inv[8], ca[8], c1[8], c2[8];
  {
    const int ch0 = (l16 >> 3) * 64 + (l16 & 7) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = stats4[256 + ch0 + j]; inv[j] = stats4[384 + ch0 + j];
      ca[j] = coef3[ch0 + j]; c1[j] = coef3[128 + ch0 + j]; c2[j] = coef3[256 + ch0 + j];
    }
  }
  const int tiles_per_row = (w2 + HT_TILE - 1) / HT_TILE;
  const int64_t ntiles = (int64_t)n_img * h2 * tiles_per_row;
  const int H = 2 * h2, W = 2 * w2;
  const HtRing ring{ring_smem, full_bar, reinterpret_cast<const bf16*>(zt), tiles_per_row, h2, w2, ntiles};
  if (PIPE) ring.start();
  int64_t kk = 0;
  PhaseAIn pa;
  pa.valid = 0;
  if ((int64_t)blockIdx.x < ntiles) {
    int n, i, j0;
    tile_coords(blockIdx.x, tiles_per_row, h2, n, i, j0);
    bwd_phase_a_load(pa, out, dout, n, i, j0, w2, H, W);
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++kk) {
    int n, i, j0;
    tile_coords(tile, tiles_per_row, h2, n, i, j0);
    bwd_phase_a_store(dzs, pa, k);
    if (tile + gridDim.x < ntiles) {       // request the next tile's maps now; consumed at the top of the next iteration
      int n2, i2, j2;
      tile_coords(tile + gridDim.x, tiles_per_row, h2, n2, i2, j2);
      bwd_phase_a_load(pa, out, dout, n2, i2, j2, w2, H, W);
    }
    __syncthreads();
    const int64_t rowoff = (((int64_t)n * h2 + i) * w2) * 128;
    const T* ztile = PIPE ? reinterpret_cast<const T*>(ring.wait(kk)) : nullptr;
    const int br = l16 >> 3;
#pragma unroll
    for (int pass = 0; pass < HT_TILE / 16; ++pass) {
      const int px = pass * 16 + pslot;
      if (j0 + px < w2) {
        const int64_t off = rowoff + (int64_t)(j0 + px) * 128 + l16 * 8;
        const HtF8 z = PIPE ? ht_lds8(ztile + px * 128 + l16 * 8) : ht_ld8(zt + off);
        float dz[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) dz[t] = dzs[br][t][px];
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = fmaf(z.v[j], L.sc[j], L.sh[j]);
          const float2 d2 = __ffma2_rn(make_float2(dz[0], dz[1]), make_float2(L.w[j][0], L.w[j][1]),
                                       __fmul2_rn(make_float2(dz[2], dz[3]), make_float2(L.w[j][2], L.w[j][3])));
          const float da = d2.x + d2.y;
          const float dy = a > 0.f ? da : 0.f;
          o[j] = ca[j] * (dy - c1[j] - (z.v[j] - mean[j]) * inv[j] * c2[j]);
        }
        ht_st8(d_zt + off, o);
      }
    }
    __syncthreads();
    if (PIPE) ring.refill(kk, tile);
  }
}

constexpr int HT_RING_BYTES = HT_STAGES * HT_TILE_BYTES;
// persistent grid of the staged kernels: ctas_per_sm CTAs on each of the 148 SMs (64 KB ring each)
static int ht_pipe_grid(int n, int h2, int w2, int ctas_per_sm) {
  const int64_t ntiles = (int64_t)n * h2 * (w2 / HT_TILE);
  int64_t g = (int64_t)DBB_NUM_SMS * ctas_per_sm;
  if (ntiles < g) g = ntiles;
  return (int)(g < 1 ? 1 : g);
}
static int ht_grid(int n, int h2, int w2) {
  const int64_t ntiles = (int64_t)n * h2 * ((w2 + HT_TILE - 1) / HT_TILE);
  int64_t g = DBB_NUM_SMS * 8;
  if (ntiles < g) g = ntiles;
  return (int)(g < 1 ? 1 : g);
}
static int ht_reduce_grid(int n, int h2, int w2) {
  const int64_t ntiles = (int64_t)n * h2 * ((w2 + HT_TILE - 1) / HT_TILE);
  int64_t g = HT_MAX_BLOCKS;
  if (ntiles < g) g = ntiles;
  return (int)(g < 1 ? 1 : g);
}

static int ht_fwd_ctas() { static const int v = getenv("DBB_HT_FWD_CTAS") ? atoi(getenv("DBB_HT_FWD_CTAS")) : 3; return v; }   // 80 regs, 66 KB: 3 fit
template <typename T> constexpr bool ht_can_pipe() { return sizeof(T) == 2; }     // the bulk-copy ring stages 16 KB bf16 tiles
template <typename T>
int head_tail_fwd(const T* zt, int n, int h2, int w2, const float* stats4, const float* w2b, const float* w2t,
                  const float* b2b, const float* b2t, float k, int out_c, float* out, cudaStream_t s) {
  if (ht_can_pipe<T>() && w2 % HT_TILE == 0) {
    static bool attr = false;
    if (!attr) { DBB_CUDA(cudaFuncSetAttribute(head_tail_fwd_kernel<true, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_RING_BYTES)); attr = true; }
    DBB_LAUNCH("head_tail_fwd", s, head_tail_fwd_kernel<true, bf16><<<ht_pipe_grid(n, h2, w2, ht_fwd_ctas()), HT_THREADS, HT_RING_BYTES, s>>>(reinterpret_cast<const bf16*>(zt), n, h2, w2, stats4, w2b, w2t, b2b, b2t, k, out_c, out));
    return DBB_OK;
  }
  DBB_LAUNCH("head_tail_fwd", s, head_tail_fwd_kernel<false, T><<<ht_grid(n, h2, w2), HT_THREADS, 0, s>>>(zt, n, h2, w2, stats4, w2b, w2t, b2b, b2t, k, out_c, out));
  return DBB_OK;
}
template <typename T>
int head_tail_bwd_reduce(const T* zt, int n, int h2, int w2, const float* stats4, const float* w2b, const float* w2t,
                         const float* out, const float* dout, float k, float* partials, int* nblk, cudaStream_t s) {
  *nblk = ht_reduce_grid(n, h2, w2);
  static const bool no_mma = getenv("DBB_HT_NO_MMA") != nullptr;      // A/B switch
  if (ht_can_pipe<T>() && w2 % HT_TILE == 0 && !no_mma) {
    // tensor-core reduce; *nblk < 0 tells head_tail_bwd_finalize that the partials are in the {M, Z} layout
    static bool attr = false;
    if (!attr) { DBB_CUDA(cudaFuncSetAttribute(head_tail_bwd_reduce_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_RING_BYTES + 1024)); attr = true; }
    CUtensorMap tm;
    int rc = encode_tmap_nhwc(&tm, reinterpret_cast<const bf16*>(zt), n, h2, w2, 128, 0, 128, 1, 1, HT_TILE, 1, 1);
    if (rc) return rc;
    const int g = ht_pipe_grid(n, h2, w2, 3);
    DBB_LAUNCH("head_tail_bwd_reduce", s, head_tail_bwd_reduce_mma_kernel<<<g, HT_THREADS, HT_RING_BYTES + 1024, s>>>(tm, n, h2, w2, stats4, out, dout, k, partials));
    *nblk = -g;
    return DBB_OK;
  }
  if (ht_can_pipe<T>() && w2 % HT_TILE == 0) {
    static bool attr = false;
    if (!attr) { DBB_CUDA(cudaFuncSetAttribute(head_tail_bwd_reduce_kernel<true, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_RING_BYTES)); attr = true; }
    const int g = ht_pipe_grid(n, h2, w2, 2);      // 2 CTAs/SM (launch bound caps the kernel at 128 registers)
    *nblk = g < *nblk ? g : *nblk;
    DBB_LAUNCH("head_tail_bwd_reduce", s, head_tail_bwd_reduce_kernel<true, bf16><<<*nblk, HT_THREADS, HT_RING_BYTES, s>>>(reinterpret_cast<const bf16*>(zt), n, h2, w2, stats4, w2b, w2t, out, dout, k, partials));
    return DBB_OK;
  }
  DBB_LAUNCH("head_tail_bwd_reduce", s, head_tail_bwd_reduce_kernel<false, T><<<*nblk, HT_THREADS, 0, s>>>(zt, n, h2, w2, stats4, w2b, w2t, out, dout, k, partials));
  return DBB_OK;
}
int head_tail_bwd_finalize(const float* partials, int nblk, int64_t count, const float* gamma_b, const float* gamma_t,
                           const float* stats4, float* dgamma_b, float* dbeta_b, float* dgamma_t, float* dbeta_t,
                           float* coef3, float* dw2b, float* dw2t, float* db2b, float* db2t, const float* w2b, const float* w2t,
                           cudaStream_t s) {
  const int layout = nblk < 0 ? 1 : 0;
  if (nblk < 0) nblk = -nblk;
  DBB_LAUNCH("head_tail_bwd_finalize", s, head_tail_bwd_finalize_kernel<<<(130 + 7) / 8, 256, 0, s>>>(partials, nblk, layout, w2b, w2t, (double)count, gamma_b, gamma_t, stats4, dgamma_b, dbeta_b,
                                                  dgamma_t, dbeta_t, coef3, dw2b, dw2t, db2b, db2t));
  return DBB_OK;
}
template <typename T>
int head_tail_bwd_apply(const T* zt, int n, int h2, int w2, const float* stats4, const float* coef3, const float* w2b,
                        const float* w2t, const float* out, const float* dout, float k, ND<T>* d_zt, cudaStream_t s) {
  if (ht_can_pipe<T>() && w2 % HT_TILE == 0) {
    static bool attr = false;
    if (!attr) { DBB_CUDA(cudaFuncSetAttribute(head_tail_bwd_apply_kernel<true, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_RING_BYTES)); attr = true; }
    DBB_LAUNCH("head_tail_bwd_apply", s, head_tail_bwd_apply_kernel<true, bf16><<<ht_pipe_grid(n, h2, w2, 2), HT_THREADS, HT_RING_BYTES, s>>>(reinterpret_cast<const bf16*>(zt), n, h2, w2, stats4, coef3, w2b, w2t, out, dout, k, reinterpret_cast<bf16*>(d_zt)));
    return DBB_OK;
  }
  DBB_LAUNCH("head_tail_bwd_apply", s, head_tail_bwd_apply_kernel<false, T><<<ht_grid(n, h2, w2), HT_THREADS, 0, s>>>(zt, n, h2, w2, stats4, coef3, w2b, w2t, out, dout, k, d_zt));
  return DBB_OK;
}
#define DBB_HT_INST(T) \
  template int head_tail_fwd<T>(const T*, int, int, int, const float*, const float*, const float*, const float*, const float*, float, int, float*, cudaStream_t); \
  template int head_tail_bwd_reduce<T>(const T*, int, int, int, const float*, const float*, const float*, const float*, const float*, float, float*, int*, cudaStream_t); \
  template int head_tail_bwd_apply<T>(const T*, int, int, int, const float*, const float*, const float*, const float*, const float*, const float*, float, T*, cudaStream_t);
DBB_HT_INST(bf16)
DBB_HT_INST(float)

}  // namespace dbb
