// conv_tcgen05.cu -- implicit-GEMM convolutions on the 5th-generation tensor cores (sm_100a).
//
// Replaces the cuDNN/oneDNN calls under nn.Conv2d / nn.ConvTranspose2d on the reference's hot path
// (src/modules/resnet.py:73-86,232; segmentation_body.py:67-76; segmentation_head.py:25-29,64-76):
//   * igemm_persist_kernel : forward / data-gradient, persistent.  A = activation tile fetched by TMA straight from the
//                     NHWC tensor (one 4-D box per filter tap; padding = TMA out-of-bounds zero fill; stride = TMA
//                     element stride), B = packed bf16 weights (2-D TMA), 128-byte swizzle, tcgen05.mma kind::f16 with a
//                     DOUBLE-BUFFERED fp32 accumulator in TMEM, 8 epilogue warps: tcgen05.ld -> +bias -> bf16 NHWC store
//                     (+ fused training-mode BatchNorm statistics of the output).  No im2col buffer.
//   * igemm_kernel  : the first, one-tile-per-CTA version (kept behind DBB_NO_PERSIST for A/B measurements).
//   * halo64_kernel : 3x3 / stride 1 / 64 -> 64 special case (weights stationary, one halo box per tile, shifted descriptors).
//   * wgrad_kernel / wgrad_row64_kernel : weight gradient.  Both operands MN-major (pixels are the GEMM K), wave-aware
//                     split-K over pixel tiles into fp32 scratch partials, reduced in a fixed order (deterministic).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, remaining warps = epilogue (one
// TMEM lane quadrant each).  smem ring of STAGES {A,B} tiles with full/empty mbarriers.
#include "common.cuh"
#include "conv.h"
#include "bn_fin.cuh"
#include "tcgen05.cuh"
#include <mutex>
#include <stdlib.h>

namespace dbb {

using namespace ptx;

constexpr int IG_THREADS = 192;
constexpr int A_TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16

template <int BLOCK_N, int STAGES>
struct IgemmSmem {
  static constexpr int B_TILE_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;   // + barriers + alignment slack
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(IG_THREADS)
igemm_kernel(const __grid_constant__ IgemmPlan p) {
  using SM = IgemmSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int nblk = blockIdx.x % n_blocks;
  int tile = blockIdx.x / n_blocks;
  const int tw = tile % p.tiles_w; tile /= p.tiles_w;
  const int th = tile % p.tiles_h; tile /= p.tiles_h;
  const int tn = tile;
  const int n0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
  const int cin_chunks = p.cin >> 6;
  const int kiters = p.ntaps * cin_chunks;
  const uint32_t a_bytes = (uint32_t)(p.bn * p.bh * p.bw) * 128u;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_x);
    prefetch_tmap(&p.tmap_w);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tap = it / cin_chunks, cc = it - tap * cin_chunks;
        uint8_t* sa = smem + s * SM::STAGE_BYTES;
        uint8_t* sb = sa + A_TILE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], a_bytes + (uint32_t)SM::B_TILE_BYTES);
        tma_load_4d(sa, &p.tmap_x, &full_bar[s], cc * 64, w0 * p.in_sw + p.dw[tap], h0 * p.in_sh + p.dh[tap], n0);
        tma_load_2d(sb, &p.tmap_w, &full_bar[s], (int)p.wtap[tap] * p.cin + cc * 64, nblk * BLOCK_N);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint64_t da = make_desc_sw128(sa, 16, 1024);
        const uint64_t db = make_desc_sw128(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x K=16 per 64-wide chunk: +32 B inside the 128 B swizzle row
          mma_bf16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
        mma_commit(&empty_bar[s]);
      }
      mma_commit(tmem_full);
    }
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 ; one output pixel (GEMM row) per thread
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int dw_ = r % p.bw, dh_ = (r / p.bw) % p.bh, dn_ = r / (p.bw * p.bh);
    const int n = n0 + dn_, h = h0 + dh_, w = w0 + dw_;
    const bool row_ok = (dn_ < p.bn) && n < p.mn && h < p.mh && w < p.mw;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 16) {
      uint32_t v[16];
      tmem_ld_x16(taddr + (uint32_t)c, v);
      tmem_ld_wait();
      const int col0 = nblk * BLOCK_N + c;
      if (row_ok && col0 < p.cout) {
        // pixel-shuffle epilogue (ConvTranspose2d k2 s2 as ONE GEMM): column = class * cls_cols + channel,
        // class (a, b) lands on output pixel (2h + a, 2w + b)
        int ch0 = col0, oh = p.out_oh, ow = p.out_ow;
        if (p.cls_cols > 0) { const int cls = col0 / p.cls_cols; ch0 = col0 - cls * p.cls_cols; oh = cls >> 1; ow = cls & 1; }
        const int64_t opix = ((int64_t)n * p.out_h + (h * p.out_sh + oh)) * p.out_w + (w * p.out_sw + ow);
        bf16* ydst = p.y + opix * p.out_c + p.out_coff + ch0;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias + ch0 + j);
        }
        uint4* dst = reinterpret_cast<uint4*>(ydst);
        if (p.accumulate) {
          const uint4 e0 = dst[0], e1 = dst[1];
          f[0] += bf16lo(e0.x); f[1] += bf16hi(e0.x); f[2] += bf16lo(e0.y); f[3] += bf16hi(e0.y);
          f[4] += bf16lo(e0.z); f[5] += bf16hi(e0.z); f[6] += bf16lo(e0.w); f[7] += bf16hi(e0.w);
          f[8] += bf16lo(e1.x); f[9] += bf16hi(e1.x); f[10] += bf16lo(e1.y); f[11] += bf16hi(e1.y);
          f[12] += bf16lo(e1.z); f[13] += bf16hi(e1.z); f[14] += bf16lo(e1.w); f[15] += bf16hi(e1.w);
        }
        uint4 o0, o1;
        o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);   o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
        o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]); o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
        dst[0] = o0; dst[1] = o1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BLOCK_N>(tmem_d);
}


// ---------------------------------------------------------------------------------------------
// Fused BatchNorm statistics (epilogue side).
//   * register path (<= 64 columns per epilogue warp): every thread keeps sum / sum-of-squares of ITS row for each of its
//     columns across all tiles of the CTA (2 FP ops per value per tile) and the warp-level column reduction
//     (column_total16: 16 shuffles per 16 columns) runs ONCE per CTA;
//   * butterfly path (BLOCK_N = 256): column reduction per chunk, added into warp-private shared-memory slices.
// Either way the warp leaves its column totals in s_stat[quadrant][2][ncols]; stat_flush adds the four quadrants into the
// global fp64 accumulators and the last CTA finalizes.
// ---------------------------------------------------------------------------------------------
// inference epilogue on 16 consecutive channels of one output pixel (conv.h: ConvEpi); res_off = element offset of the
// pixel's channel ch0 inside the compact residual tensor
__device__ __forceinline__ void epi_apply16(float (&f)[16], const ConvEpi& e, int ch0, int64_t res_off) {
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], __ldg(e.scale + ch0 + i), __ldg(e.shift + ch0 + i));
  if (e.res) {
    const uint4* r = reinterpret_cast<const uint4*>(e.res + res_off);
    const uint4 r0 = r[0], r1 = r[1];
    f[0] += bf16lo(r0.x); f[1] += bf16hi(r0.x); f[2] += bf16lo(r0.y); f[3] += bf16hi(r0.y);
    f[4] += bf16lo(r0.z); f[5] += bf16hi(r0.z); f[6] += bf16lo(r0.w); f[7] += bf16hi(r0.w);
    f[8] += bf16lo(r1.x); f[9] += bf16hi(r1.x); f[10] += bf16lo(r1.y); f[11] += bf16hi(r1.y);
    f[12] += bf16lo(r1.z); f[13] += bf16hi(r1.z); f[14] += bf16lo(r1.w); f[15] += bf16hi(r1.w);
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
  }
}

template <int NTHREADS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" :: "n"(NTHREADS) : "memory"); }   // epilogue warps only

__device__ __forceinline__ void unpack16(const uint4& o0, const uint4& o1, float (&x)[16]) {
  x[0] = bf16lo(o0.x); x[1] = bf16hi(o0.x); x[2] = bf16lo(o0.y); x[3] = bf16hi(o0.y);
  x[4] = bf16lo(o0.z); x[5] = bf16hi(o0.z); x[6] = bf16lo(o0.w); x[7] = bf16hi(o0.w);
  x[8] = bf16lo(o1.x); x[9] = bf16hi(o1.x); x[10] = bf16lo(o1.y); x[11] = bf16hi(o1.y);
  x[12] = bf16lo(o1.z); x[13] = bf16hi(o1.z); x[14] = bf16lo(o1.w); x[15] = bf16hi(o1.w);
}

// o0/o1 = the 16 bf16 outputs this lane stored for its row (all-zero when the row is outside the tensor)
__device__ __forceinline__ void stat_accumulate_regs(const uint4& o0, const uint4& o1, float (&acc_s)[16], float (&acc_q)[16]) {
  float x[16];
  unpack16(o0, o1, x);
#pragma unroll
  for (int j = 0; j < 16; ++j) { acc_s[j] += x[j]; acc_q[j] = fmaf(x[j], x[j], acc_q[j]); }
}

// warp-level column totals of 16 columns -> the warp's shared slices (ADD = accumulate across tiles, else overwrite)
template <bool ADD>
__device__ __forceinline__ void stat_reduce_store16(const float (&xs)[16], const float (&xq)[16], int lane, float* w_sum, float* w_sq) {
  const float ts = column_total16(xs, lane), tq = column_total16(xq, lane);
  // even lanes own the sum, odd lanes the sum of squares of column (lane >> 1) & 15
  float* dst = ((lane & 1) ? w_sq : w_sum) + column_of_lane16(lane);
  const float v = (lane & 1) ? tq : ts;
  if (ADD) *dst += v; else *dst = v;
}

__device__ __forceinline__ void stat_accumulate_smem(const uint4& o0, const uint4& o1, int lane, float* w_sum, float* w_sq) {
  float x[16], q[16];
  unpack16(o0, o1, x);
#pragma unroll
  for (int j = 0; j < 16; ++j) q[j] = x[j] * x[j];
  stat_reduce_store16<true>(x, q, lane, w_sum, w_sq);
}

// After the CTA's last tile: shared partials -> global fp64 accumulators; the last CTA finalizes.  Called by all NTHREADS
// epilogue threads (e = 0..NTHREADS-1).  s_stat = [4 quadrants][2][ncols]; column j of this CTA is channel
// out_coff + (cls_cols ? col % cls_cols : col) of y.
template <int NTHREADS>
__device__ __forceinline__ void stat_flush(const ConvStats& st, const float* s_stat, int ncols, int col0, int cout, int cls_cols,
                                           int out_coff, int out_c, int e, uint32_t* s_flag) {
  epi_bar_sync<NTHREADS>();
  for (int j = e; j < ncols; j += NTHREADS) {
    const int col = col0 + j;
    if (col >= cout) break;
    const int ch = out_coff + (cls_cols > 0 ? col % cls_cols : col);
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { a += s_stat[w * 2 * ncols + j]; b += s_stat[w * 2 * ncols + ncols + j]; }
    atomicAdd(&st.gacc[ch], (double)a);
    atomicAdd(&st.gacc[out_c + ch], (double)b);
  }
  __threadfence();
  epi_bar_sync<NTHREADS>();
  if (e == 0) *s_flag = (atomicAdd(st.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  epi_bar_sync<NTHREADS>();
  if (*s_flag) {
    __threadfence();
    bn_finalize_channels(st.fin, out_c, st.count, st.gacc, e, NTHREADS);
    if (e == 0) *st.counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent implicit GEMM: one CTA per SM slot walks m-tiles (fixed n-block per CTA), the TMEM accumulator is double
// buffered so the epilogue of tile j (tcgen05.ld, bias, bf16 store, BatchNorm statistics) overlaps the MMAs of tile j+1,
// and the TMA ring keeps streaming across tile boundaries.  320 threads: warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..9 = epilogue; warp w drains TMEM lane quadrant w % 4 and column half (w - 2) / 4 of the tile.
// ---------------------------------------------------------------------------------------------
constexpr int IGP_THREADS = 320;
constexpr int IGP_EPI_THREADS = 256;

template <int BLOCK_N, int STAGES>
struct IgemmPSmem {
  static constexpr int B_TILE_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;                 // full[S], empty[S], t_full[2], t_empty[2], slot, flag
  static constexpr int STAT_OFF = BAR_OFF + 256;
  // Staged epilogue (64-wide tiles: the store-bound ConvTranspose / stem / 1x1 GEMMs).  Every epilogue warp owns a
  // [32 rows][CW bf16 + 16 B pad] slab: rows are written as they come out of TMEM (one row per lane) and read back with
  // consecutive lanes on consecutive 16-byte pieces of a row, so that a store instruction covers 8 half-lines of 64
  // contiguous bytes (full 32-byte sectors) instead of 32 lines with a 16-byte partial-sector write each.  ncu on the
  // ConvTranspose GEMM (profiles/prof_convt_r02.md): 13.1 M partial-sector writes for 6.55 M sectors of output, 119 us.
  static constexpr bool STAGED = BLOCK_N == 64;
  static constexpr int EPI_PITCH = (BLOCK_N / 2) * 2 + 16;
  static constexpr int EPI_OFF = STAT_OFF + 4 * 2 * BLOCK_N * 4;          // + [4 quadrants][2][BLOCK_N] statistics
  static constexpr int EPI_PITCH_F32 = (BLOCK_N / 2) * 4 + 16;           // lean epilogue: raw fp32 accumulators are staged
  static constexpr int EPI_BYTES = STAGED ? 8 * 32 * EPI_PITCH_F32 : 0;
  static constexpr int TOTAL = EPI_OFF + EPI_BYTES + 1024;
};

template <int BLOCK_N, int STAGES, bool STATS>
__global__ void __launch_bounds__(IGP_THREADS)
igemm_persist_kernel(const __grid_constant__ IgemmPlan p) {
  using SM = IgemmPSmem<BLOCK_N, STAGES>;
  constexpr int CW = BLOCK_N / 2;                  // columns per epilogue warp
  constexpr int NCH = CW / 16;                     // 16-column chunks per epilogue warp
  // register statistics (64 accumulators per thread) only with a deep ring, i.e. one CTA per SM: the shallow-ring variant is
  // for the store-bound launches (conv1, ConvTranspose 64->64, 1x1 laterals), where resident epilogue warps matter more than
  // the shuffles of the shared-memory reduction
  constexpr bool REG_STATS = STATS && CW <= 32 && STAGES > 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* t_full = empty_bar + STAGES;     // [2]
  uint64_t* t_empty = t_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  uint32_t* s_flag = tmem_slot + 1;
  float* s_stat = reinterpret_cast<float*>(smem + SM::STAT_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int nblk = blockIdx.x % n_blocks;
  const int mt0 = blockIdx.x / n_blocks, mt_step = gridDim.x / n_blocks;       // gridDim.x is a multiple of n_blocks
  const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
  const int cin_chunks = p.cin >> 6;
  const int kiters = p.ntaps * cin_chunks;
  const uint32_t a_bytes = (uint32_t)(p.bn * p.bh * p.bw) * 128u;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_x);
    prefetch_tmap(&p.tmap_w);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
    fence_barrier_init();
  }
  if (STATS) { for (int i = threadIdx.x; i < 8 * BLOCK_N; i += IGP_THREADS) s_stat[i] = 0.f; }
  if (warp == 1) tmem_alloc<2 * BLOCK_N>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t g = 0;                       // ring position, runs on across tiles
      for (int mt = mt0; mt < m_tiles; mt += mt_step) {
        int t = mt;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; const int tn = t / p.tiles_h;
        const int n0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
        for (int it = 0; it < kiters; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          const int tap = it / cin_chunks, cc = it - tap * cin_chunks;
          uint8_t* sa = smem + s * SM::STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], a_bytes + (uint32_t)SM::B_TILE_BYTES);
          tma_load_4d(sa, &p.tmap_x, &full_bar[s], cc * 64, w0 * p.in_sw + p.dw[tap], h0 * p.in_sh + p.dh[tap], n0);
          tma_load_2d(sb, &p.tmap_w, &full_bar[s], (int)p.wtap[tap] * p.cin + cc * 64, nblk * BLOCK_N);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      uint32_t g = 0, j = 0;
      for (int mt = mt0; mt < m_tiles; mt += mt_step, ++j) {
        const uint32_t buf = j & 1u;
        mbar_wait(&t_empty[buf], ((j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dcol = tmem_d + buf * (uint32_t)BLOCK_N;
        for (int it = 0; it < kiters; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          const uint64_t da = make_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_desc_sw128(sa + A_TILE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_bf16_ss(dcol, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&t_full[buf]);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int cbase = half * CW;                    // first column (inside the BLOCK_N tile) of this warp
    const int dw_ = r % p.bw, dh_ = (r / p.bw) % p.bh, dn_ = r / (p.bw * p.bh);
    float acc_s[REG_STATS ? NCH : 1][16], acc_q[REG_STATS ? NCH : 1][16];
    if (REG_STATS) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc_s[c][i] = 0.f; acc_q[c][i] = 0.f; }
    }
    float* w_sum = s_stat + q * 2 * BLOCK_N + cbase;
    float* w_sq = w_sum + BLOCK_N;
    // staged write-out (see IgemmPSmem): lane -> (row i*RPI + lane/PP, 16-byte piece lane%PP) for i = 0..PP-1
    constexpr int PP = CW / 8, RPI = 32 / PP;
    // (not with fused statistics: that variant runs one CTA per SM and is bound by the per-tile epilogue latency chain --
    //  measured 0.279 -> 0.375 ms for the two training-mode ConvTranspose launches with the extra shared-memory round trip)
    const bool staged = SM::STAGED && !REG_STATS && !p.accumulate && !p.no_staged_epilogue && (p.cls_cols == 0 || p.cls_cols % CW == 0);
    uint8_t* slab = smem + SM::EPI_OFF + (warp - 2) * 32 * SM::EPI_PITCH;
    int s_dn[SM::STAGED ? PP : 1], s_dh[SM::STAGED ? PP : 1], s_dw[SM::STAGED ? PP : 1];
    if (SM::STAGED) {
#pragma unroll
      for (int i = 0; i < PP; ++i) {
        const int rr = q * 32 + i * RPI + lane / PP;
        s_dw[i] = rr % p.bw; s_dh[i] = (rr / p.bw) % p.bh; s_dn[i] = rr / (p.bw * p.bh);
      }
    }
    // ---- lean epilogue of the 64-wide tiles (conv1, ConvTranspose 64->64, 1x1 laterals: store-bound launches whose epilogue
    // was ISSUE-bound -- ncu, profiles/prof_convt_r02.md: 550 warp instructions per warp and tile with register statistics,
    // 966 with the shuffle reduction, for 1,024 outputs).  The raw fp32 accumulators of the warp's 32 x 32 block go to its
    // shared-memory slab straight from TMEM (both 16-column loads in flight, one wait); they are read back with consecutive
    // lanes on consecutive 32-byte pieces of a row, so each lane always owns the SAME 8 columns: their bias lives in 8
    // registers, their BatchNorm statistics in 16 (combined across lanes once per CTA), and a store instruction writes
    // 8 x 64 contiguous bytes.
    const bool lean = SM::STAGED && !p.accumulate && !p.epi.scale && !p.no_staged_epilogue && (p.cls_cols == 0 || p.cls_cols % CW == 0);
    if (SM::STAGED && lean) {
      const int piece = lane % PP;
      const int colw = nblk * BLOCK_N + cbase;
      int valid_cols = p.cout - colw; valid_cols = valid_cols < 0 ? 0 : (valid_cols > CW ? CW : valid_cols);
      int chw = colw, oh = p.out_oh, ow = p.out_ow;
      if (p.cls_cols > 0) { const int cls = colw / p.cls_cols; chw = colw - cls * p.cls_cols; oh = cls >> 1; ow = cls & 1; }
      const bool piece_ok = piece * 8 < valid_cols;
      float bias8[8], ssum[8], ssq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { bias8[i] = (p.bias && piece_ok) ? __ldg(p.bias + chw + piece * 8 + i) : 0.f; ssum[i] = 0.f; ssq[i] = 0.f; }
      uint8_t* fslab = smem + SM::EPI_OFF + (warp - 2) * 32 * SM::EPI_PITCH_F32;
      bf16* ybase = p.y + p.out_coff + chw + piece * 8;
      uint32_t j = 0;
      for (int mt = mt0; mt < m_tiles; mt += mt_step, ++j) {
        const uint32_t buf = j & 1u;
        int t = mt;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; const int tn = t / p.tiles_h;
        mbar_wait_sleep(&t_full[buf], (j >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)BLOCK_N + (uint32_t)cbase;
        uint32_t v0[16], v1[16];
        tmem_ld_x16(taddr, v0);
        tmem_ld_x16(taddr + 16u, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[buf]);                 // the accumulator buffer is free for tile j + 2
        uint4* sl = reinterpret_cast<uint4*>(fslab + lane * SM::EPI_PITCH_F32);
        sl[0] = make_uint4(v0[0], v0[1], v0[2], v0[3]);   sl[1] = make_uint4(v0[4], v0[5], v0[6], v0[7]);
        sl[2] = make_uint4(v0[8], v0[9], v0[10], v0[11]); sl[3] = make_uint4(v0[12], v0[13], v0[14], v0[15]);
        sl[4] = make_uint4(v1[0], v1[1], v1[2], v1[3]);   sl[5] = make_uint4(v1[4], v1[5], v1[6], v1[7]);
        sl[6] = make_uint4(v1[8], v1[9], v1[10], v1[11]); sl[7] = make_uint4(v1[12], v1[13], v1[14], v1[15]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < PP; ++i) {
          const int n2 = tn * p.bn + s_dn[i], h2 = th * p.bh + s_dh[i], w2 = tw * p.bw + s_dw[i];
          const bool ok = (s_dn[i] < p.bn) && n2 < p.mn && h2 < p.mh && w2 < p.mw && piece_ok;
          if (ok) {
            const float4* src = reinterpret_cast<const float4*>(fslab + (i * RPI + lane / PP) * SM::EPI_PITCH_F32 + piece * 32);
            const float4 a = src[0], b = src[1];
            uint4 o;
            o.x = pack_bf16(a.x + bias8[0], a.y + bias8[1]); o.y = pack_bf16(a.z + bias8[2], a.w + bias8[3]);
            o.z = pack_bf16(b.x + bias8[4], b.y + bias8[5]); o.w = pack_bf16(b.z + bias8[6], b.w + bias8[7]);
            const int64_t opix = ((int64_t)n2 * p.out_h + (h2 * p.out_sh + oh)) * p.out_w + (w2 * p.out_sw + ow);
            *reinterpret_cast<uint4*>(ybase + opix * p.out_c) = o;
            if (STATS) {                                           // on the bf16-rounded values, as the next layer sees them
              const float x0 = bf16lo(o.x), x1 = bf16hi(o.x), x2 = bf16lo(o.y), x3 = bf16hi(o.y);
              const float x4 = bf16lo(o.z), x5 = bf16hi(o.z), x6 = bf16lo(o.w), x7 = bf16hi(o.w);
              ssum[0] += x0; ssum[1] += x1; ssum[2] += x2; ssum[3] += x3; ssum[4] += x4; ssum[5] += x5; ssum[6] += x6; ssum[7] += x7;
              ssq[0] = fmaf(x0, x0, ssq[0]); ssq[1] = fmaf(x1, x1, ssq[1]); ssq[2] = fmaf(x2, x2, ssq[2]); ssq[3] = fmaf(x3, x3, ssq[3]);
              ssq[4] = fmaf(x4, x4, ssq[4]); ssq[5] = fmaf(x5, x5, ssq[5]); ssq[6] = fmaf(x6, x6, ssq[6]); ssq[7] = fmaf(x7, x7, ssq[7]);
            }
          }
        }
        __syncwarp();
      }
      if (STATS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int off = PP; off < 32; off <<= 1) {                // lanes with the same piece
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], off);
            ssq[i] += __shfl_xor_sync(0xffffffffu, ssq[i], off);
          }
        }
        if (lane < PP) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { w_sum[lane * 8 + i] = ssum[i]; w_sq[lane * 8 + i] = ssq[i]; }
        }
        stat_flush<IGP_EPI_THREADS>(p.st, s_stat, BLOCK_N, nblk * BLOCK_N, p.cout, p.cls_cols, p.out_coff, p.out_c,
                                    (int)threadIdx.x - 64, s_flag);
      }
    } else {
    uint32_t j = 0;
    for (int mt = mt0; mt < m_tiles; mt += mt_step, ++j) {
      const uint32_t buf = j & 1u;
      int t = mt;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; const int tn = t / p.tiles_h;
      const int n = tn * p.bn + dn_, h = th * p.bh + dh_, w = tw * p.bw + dw_;
      const bool row_ok = (dn_ < p.bn) && n < p.mn && h < p.mh && w < p.mw;
      mbar_wait_sleep(&t_full[buf], (j >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)BLOCK_N + (uint32_t)cbase;
#pragma unroll
      for (int ci = 0; ci < NCH; ++ci) {
        const int c = ci * 16;
        uint32_t v[16];
        tmem_ld_x16(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        if (ci == NCH - 1) {               // this warp's part of the accumulator is drained
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
        const int col0 = nblk * BLOCK_N + cbase + c;
        if (col0 >= p.cout) continue;      // warp-uniform
        int ch0 = col0, oh = p.out_oh, ow = p.out_ow;
        if (p.cls_cols > 0) { const int cls = col0 / p.cls_cols; ch0 = col0 - cls * p.cls_cols; oh = cls >> 1; ow = cls & 1; }
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + ch0 + i);
        }
        uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
        if (row_ok) {
          const int64_t opix = ((int64_t)n * p.out_h + (h * p.out_sh + oh)) * p.out_w + (w * p.out_sw + ow);
          uint4* dst = reinterpret_cast<uint4*>(p.y + opix * p.out_c + p.out_coff + ch0);
          if (p.epi.scale) epi_apply16(f, p.epi, ch0, opix * p.cout + ch0);
          if (p.accumulate) {
            const uint4 e0 = dst[0], e1 = dst[1];
            f[0] += bf16lo(e0.x); f[1] += bf16hi(e0.x); f[2] += bf16lo(e0.y); f[3] += bf16hi(e0.y);
            f[4] += bf16lo(e0.z); f[5] += bf16hi(e0.z); f[6] += bf16lo(e0.w); f[7] += bf16hi(e0.w);
            f[8] += bf16lo(e1.x); f[9] += bf16hi(e1.x); f[10] += bf16lo(e1.y); f[11] += bf16hi(e1.y);
            f[12] += bf16lo(e1.z); f[13] += bf16hi(e1.z); f[14] += bf16lo(e1.w); f[15] += bf16hi(e1.w);
          }
          o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);   o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]); o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
          if (!staged) { dst[0] = o0; dst[1] = o1; }
        }
        if (SM::STAGED && staged) {
          uint4* sl = reinterpret_cast<uint4*>(slab + lane * SM::EPI_PITCH + ci * 32);
          sl[0] = o0; sl[1] = o1;
        }
        if (REG_STATS) stat_accumulate_regs(o0, o1, acc_s[REG_STATS ? ci : 0], acc_q[REG_STATS ? ci : 0]);
        else if (STATS) stat_accumulate_smem(o0, o1, lane, w_sum + c, w_sq + c);
      }
      if (SM::STAGED && staged) {
        // columns of this warp that exist (cout may end inside the tile), and their place in y
        const int colw = nblk * BLOCK_N + cbase;
        int valid_cols = p.cout - colw; valid_cols = valid_cols < 0 ? 0 : (valid_cols > CW ? CW : valid_cols);
        int chw = colw, oh = p.out_oh, ow = p.out_ow;
        if (p.cls_cols > 0) { const int cls = colw / p.cls_cols; chw = colw - cls * p.cls_cols; oh = cls >> 1; ow = cls & 1; }
        __syncwarp();
        const int piece = lane % PP;
#pragma unroll
        for (int i = 0; i < PP; ++i) {
          const int n2 = tn * p.bn + s_dn[i], h2 = th * p.bh + s_dh[i], w2 = tw * p.bw + s_dw[i];
          const bool ok = (s_dn[i] < p.bn) && n2 < p.mn && h2 < p.mh && w2 < p.mw && piece * 8 < valid_cols;
          if (ok) {
            const uint4 v = *reinterpret_cast<const uint4*>(slab + (i * RPI + lane / PP) * SM::EPI_PITCH + piece * 16);
            const int64_t opix = ((int64_t)n2 * p.out_h + (h2 * p.out_sh + oh)) * p.out_w + (w2 * p.out_sw + ow);
            *reinterpret_cast<uint4*>(p.y + opix * p.out_c + p.out_coff + chw + piece * 8) = v;
          }
        }
        __syncwarp();
      }
    }
    if (STATS) {
      if (REG_STATS) {
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci)
          stat_reduce_store16<false>(acc_s[REG_STATS ? ci : 0], acc_q[REG_STATS ? ci : 0], lane, w_sum + ci * 16, w_sq + ci * 16);
      }
      stat_flush<IGP_EPI_THREADS>(p.st, s_stat, BLOCK_N, nblk * BLOCK_N, p.cout, p.cls_cols, p.out_coff, p.out_c,
                                  (int)threadIdx.x - 64, s_flag);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<2 * BLOCK_N>(tmem_d);
}


// ---------------------------------------------------------------------------------------------
// halo64: 3x3 stride-1 64->64 convolution (fprop and dgrad), persistent + weights-stationary + shifted windows
// ---------------------------------------------------------------------------------------------
constexpr int HALO_B_BYTES = 9 * 64 * 128;      // nine 64x64 bf16 weight tiles
constexpr int HALO_A_STAGE = 224 * 128;         // up to 224 pixel rows of 128 B
constexpr int HALO_STAGES = 3;
constexpr int HALO_SMEM = HALO_B_BYTES + HALO_STAGES * HALO_A_STAGE + 256 + 4 * 2 * 64 * 4 + 1024;

// Shifted-window operands start at 128 B granularity, not on the 1024 B swizzle repeat.  Measured on B200: tcgen05 applies the
// 128 B swizzle to the ABSOLUTE shared-memory address bits (the same function TMA used when writing the tile), so the
// descriptor's base_offset field must stay 0; setting it to (addr >> 7) & 7 gives wrong products (kept as a bring-up switch).
__device__ __forceinline__ uint64_t make_desc_sw128_off(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, int base_mode) {
  uint64_t d = make_desc_sw128(smem_addr, lbo_bytes, sbo_bytes);
  if (base_mode) d |= (uint64_t)((smem_addr >> 7) & 7u) << 49;
  return d;
}

__global__ void __launch_bounds__(IGP_THREADS)
halo64_kernel(const __grid_constant__ HaloPlan p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = smem + HALO_B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sA + HALO_STAGES * HALO_A_STAGE);
  uint64_t* a_empty = a_full + HALO_STAGES;
  uint64_t* b_full = a_empty + HALO_STAGES;
  uint64_t* t_full = b_full + 1;       // [2]
  uint64_t* t_empty = t_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  uint32_t* s_flag = tmem_slot + 1;
  float* s_stat = reinterpret_cast<float*>(sA + HALO_STAGES * HALO_A_STAGE + 256);   // [4 warps][2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = (uint32_t)((p.R + 2) * p.pitch) * 128u;
  for (int i = threadIdx.x; i < 512; i += IGP_THREADS) s_stat[i] = 0.f;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_x);
    prefetch_tmap(&p.tmap_w);
    for (int s = 0; s < HALO_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(b_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  // (shared memory beyond the box is stale: it only feeds GEMM rows that are discarded, rows are independent in M)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)HALO_B_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(sB + t * 8192, &p.tmap_w, b_full, t * 64, 0);
      int k = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k) {
        const int s = k % HALO_STAGES;
        mbar_wait(&a_empty[s], ((uint32_t)(k / HALO_STAGES) & 1u) ^ 1u);
        int t = tile;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; const int n = t / p.tiles_h;
        mbar_arrive_expect_tx(&a_full[s], a_bytes);
        tma_load_4d(sA + s * HALO_A_STAGE, &p.tmap_x, &a_full[s], 0, tw * p.bw - 1, th * p.R - 1, n);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      mbar_wait(b_full, 0);
      int k = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k) {
        const int s = k % HALO_STAGES, buf = k & 1;
        mbar_wait(&t_empty[buf], ((uint32_t)(k >> 1) & 1u) ^ 1u);
        mbar_wait(&a_full[s], (uint32_t)(k / HALO_STAGES) & 1u);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * HALO_A_STAGE), b0 = smem_u32(sB);
        const uint32_t dcol = tmem_d + (uint32_t)(buf * 64);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t aa = a0 + (uint32_t)p.off[t] * 128u;
          const uint32_t bb = b0 + (uint32_t)p.wtap[t] * 8192u;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            mma_bf16_ss(dcol, make_desc_sw128_off(aa + k4 * 32, 16, 1024, p.base_offset_mode), make_desc_sw128(bb + k4 * 32, 16, 1024),
                        idesc, (t | k4) != 0);
        }
        mma_commit(&a_empty[s]);
        mma_commit(&t_full[buf]);
      }
    }
  } else {
    // 8 epilogue warps: warp w drains TMEM lane quadrant w % 4, channels [32*half, 32*half + 32) with half = (w - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int dr = r / p.pitch, dc = r - dr * p.pitch;
    float acc_s[2][16], acc_q[2][16];      // fused BatchNorm statistics: this row's running sums for its 32 channels
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc_s[c][i] = 0.f; acc_q[c][i] = 0.f; }
    int k = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++k) {
      const int buf = k & 1;
      int t = tile;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; const int n = t / p.tiles_h;
      const int hh = th * p.R + dr, ww = tw * p.bw + dc;
      const bool ok = dr < p.R && dc < p.bw && hh < p.h && ww < p.w;
      mbar_wait_sleep(&t_full[buf], (uint32_t)(k >> 1) & 1u);
      tc_fence_after();
      uint32_t v[2][16];
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 64 + half * 32);
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_x16(taddr + (uint32_t)(c * 16), v[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);       // this warp's part is drained (8 arrivals free the buffer)
      bf16* yrow = p.y + (((int64_t)n * p.h + hh) * p.w + ww) * p.out_c + p.out_coff + half * 32;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint4 o0 = make_uint4(0, 0, 0, 0), o1 = o0;
        if (ok) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[c][j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias + half * 32 + c * 16 + j);
          }
          uint4* dst = reinterpret_cast<uint4*>(yrow + c * 16);
          if (p.epi.scale)
            epi_apply16(f, p.epi, half * 32 + c * 16, (((int64_t)n * p.h + hh) * p.w + ww) * 64 + half * 32 + c * 16);
          if (p.accumulate) {
            const uint4 e0 = dst[0], e1 = dst[1];
            f[0] += bf16lo(e0.x); f[1] += bf16hi(e0.x); f[2] += bf16lo(e0.y); f[3] += bf16hi(e0.y);
            f[4] += bf16lo(e0.z); f[5] += bf16hi(e0.z); f[6] += bf16lo(e0.w); f[7] += bf16hi(e0.w);
            f[8] += bf16lo(e1.x); f[9] += bf16hi(e1.x); f[10] += bf16lo(e1.y); f[11] += bf16hi(e1.y);
            f[12] += bf16lo(e1.z); f[13] += bf16hi(e1.z); f[14] += bf16lo(e1.w); f[15] += bf16hi(e1.w);
          }
          o0.x = pack_bf16(f[0], f[1]);   o0.y = pack_bf16(f[2], f[3]);   o0.z = pack_bf16(f[4], f[5]);   o0.w = pack_bf16(f[6], f[7]);
          o1.x = pack_bf16(f[8], f[9]);   o1.y = pack_bf16(f[10], f[11]); o1.z = pack_bf16(f[12], f[13]); o1.w = pack_bf16(f[14], f[15]);
          dst[0] = o0; dst[1] = o1;
        }
        if (p.st.enabled) stat_accumulate_regs(o0, o1, acc_s[c], acc_q[c]);
      }
    }
    if (p.st.enabled) {
      float* w_sum = s_stat + q * 128 + half * 32;
#pragma unroll
      for (int c = 0; c < 2; ++c) stat_reduce_store16<false>(acc_s[c], acc_q[c], lane, w_sum + c * 16, w_sum + 64 + c * 16);
      stat_flush<IGP_EPI_THREADS>(p.st, s_stat, 64, 0, 64, 0, p.out_coff, p.out_c, (int)threadIdx.x - 64, s_flag);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_d);
}

static void halo_pick(int h, int w, int* R, int* bw) {
  double best = -1; *R = 1; *bw = 16;
  for (int b = 8; b <= 45; ++b)
    for (int r = 1; r * (b + 2) <= 128; ++r) {
      if ((r + 2) * (b + 2) > 224) continue;
      const int tw = (w + b - 1) / b, th = (h + r - 1) / r;
      const double eff = (double)h * w / ((double)tw * th * 128.0);
      if (eff > best + 1e-9) { best = eff; *R = r; *bw = b; }
    }
}
int halo64_supported(int h, int w) { return h >= 1 && w >= 1; }

int halo64_plan(HaloPlan* p, const bf16* x, int n, int h, int w, int x_ctotal, int x_coff, const bf16* wp, int dgrad) {
  static const int base_mode = getenv("DBB_HALO_BASE1") ? 1 : 0;
  p->n = n; p->h = h; p->w = w;
  halo_pick(h, w, &p->R, &p->bw);
  p->pitch = p->bw + 2;
  p->tiles_h = (h + p->R - 1) / p->R; p->tiles_w = (w + p->bw - 1) / p->bw;
  p->total_tiles = n * p->tiles_h * p->tiles_w;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int t = kh * 3 + kw;
      // box origin is (h0-1, w0-1): fprop reads x[h+kh-1, w+kw-1] -> offset (kh, kw); dgrad reads dy[h+1-kh, w+1-kw] -> (2-kh, 2-kw)
      const int oh = dgrad ? 2 - kh : kh, ow = dgrad ? 2 - kw : kw;
      p->off[t] = (int16_t)(oh * p->pitch + ow);
      p->wtap[t] = (uint8_t)t;
    }
  p->base_offset_mode = base_mode;
  int rc = encode_tmap_nhwc(&p->tmap_x, x, n, h, w, x_ctotal, x_coff, 64, 1, p->R + 2, p->pitch, 1, 1);
  if (rc) return rc;
  return encode_weights_public(&p->tmap_w, wp, 576, 64, 64);
}

int halo64_launch(const HaloPlan& p, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    DBB_CUDA(cudaFuncSetAttribute(halo64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HALO_SMEM));
    attr_done = true;
  }
  int grid = p.total_tiles < DBB_NUM_SMS ? p.total_tiles : DBB_NUM_SMS;
  const char* label = "halo64";
  if (prof_enabled()) {
    char tmp[96];
    snprintf(tmp, sizeof(tmp), "igemm_bn64_m%lld_n64_k576_halo", (long long)p.n * p.h * p.w);
    label = prof_label(tmp);
  }
  DBB_LAUNCH(label, s, halo64_kernel<<<grid, IGP_THREADS, HALO_SMEM, s>>>(p));
  return DBB_OK;
}


// ---------------------------------------------------------------------------------------------
// wgrad_row64: dW[co][ci][kh][kw] = sum_{n,i,j} dy[n,i,j,co] * x[n,i+kh-1,j+kw-1,ci]   (3x3, stride 1, 64 -> 64)
// ---------------------------------------------------------------------------------------------
constexpr int WR_XSLOTS = 4, WR_DSLOTS = 2;

__global__ void __launch_bounds__(IG_THREADS)
wgrad_row64_kernel(const __grid_constant__ WgradRowPlan p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int xbytes = (p.kp + 2) * 128, xslot = (xbytes + 1023) & ~1023, dbytes = p.kp * 128;
  uint8_t* sX = smem;
  uint8_t* sD = smem + WR_XSLOTS * xslot;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(sD + WR_DSLOTS * dbytes + 1024);   // +1 KB guard: the junk half of the
  uint64_t* dfull = xfull + WR_XSLOTS;                                              // single-tap groups reads one row past
  uint64_t* step_done = dfull + WR_DSLOTS;
  uint64_t* acc_full = step_done + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x;
  const int img = chunk / p.chunks_per_image;
  const int i0 = (chunk % p.chunks_per_image) * p.rows_per_chunk;
  const int nrows = min(p.rows_per_chunk, p.h - i0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_x);
    prefetch_tmap(&p.tmap_dy);
    for (int s = 0; s < WR_XSLOTS; ++s) mbar_init(&xfull[s], 1);
    for (int s = 0; s < WR_DSLOTS; ++s) mbar_init(&dfull[s], 1);
    for (int s = 0; s < 4; ++s) mbar_init(&step_done[s], 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      auto load_x = [&](int q) {      // x row (i0 - 1 + q) -> slot q % 4
        const int s = q % WR_XSLOTS;
        mbar_arrive_expect_tx(&xfull[s], (uint32_t)xbytes);
        tma_load_4d(sX + s * xslot, &p.tmap_x, &xfull[s], 0, -1, i0 - 1 + q, img);
      };
      auto load_d = [&](int j) {
        const int s = j % WR_DSLOTS;
        mbar_arrive_expect_tx(&dfull[s], (uint32_t)dbytes);
        tma_load_4d(sD + s * dbytes, &p.tmap_dy, &dfull[s], 0, 0, i0 + j, img);
      };
      load_x(0); load_x(1); load_x(2); load_d(0);
      for (int j = 1; j < nrows; ++j) {
        if (j >= 2) mbar_wait(&step_done[(j - 2) & 3], (uint32_t)((j - 2) >> 2) & 1u);
        load_x(j + 2);
        load_d(j);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      const int kslices = p.kp >> 4;
      for (int j = 0; j < nrows; ++j) {
        mbar_wait(&dfull[j % WR_DSLOTS], (uint32_t)(j / WR_DSLOTS) & 1u);
        for (int q = j; q <= j + 2; ++q) mbar_wait(&xfull[q % WR_XSLOTS], (uint32_t)(q / WR_XSLOTS) & 1u);
        tc_fence_after();
        const uint32_t d0 = smem_u32(sD + (j % WR_DSLOTS) * dbytes);
        for (int ks = 0; ks < kslices; ++ks) {
          const uint64_t db = make_desc_sw128(d0 + (uint32_t)ks * 2048u, 16, 1024);
#pragma unroll
          for (int g = 0; g < 6; ++g) {
            const int kh = g < 3 ? g : g - 3, kw = g < 3 ? 0 : 2;
            const uint32_t a0 = smem_u32(sX + ((j + kh) % WR_XSLOTS) * xslot) + (uint32_t)(kw + 16 * ks) * 128u;
            // M-major A: rows 0..63 = tap (kh,kw), rows 64..127 = tap (kh,kw+1) one pixel (128 B) further
            mma_bf16_ss(tmem_d + (uint32_t)(g * 64), make_desc_sw128(a0, 128, 1024), db, idesc, (j | ks) != 0);
          }
        }
        mma_commit(&step_done[j & 3]);
      }
      mma_commit(acc_full);
    }
  } else {
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
    for (int g = 0; g < 6; ++g) {
      const int kh = g < 3 ? g : g - 3;
      const int kw = g < 3 ? (m >> 6) : 2;
      const bool ok = g < 3 || m < 64;
      float* dst = p.ws + (((int64_t)chunk * 9 + kh * 3 + kw) * 64 + (m & 63)) * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[16];
        tmem_ld_x16(taddr + (uint32_t)(g * 64 + c * 16), v);
        tmem_ld_wait();
        if (ok) {
          float4* d4 = reinterpret_cast<float4*>(dst + c * 16);
          d4[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          d4[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          d4[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
          d4[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_d);
}

// dW[co][ci][tap] = sum_chunk ws[chunk][tap][ci][co]   (fixed order)
__global__ void __launch_bounds__(256) wgrad_row64_reduce_kernel(const float* __restrict__ ws, int nchunks, float* __restrict__ dw) {
  const int i = blockIdx.x * 256 + threadIdx.x;      // over [tap][ci][co]
  if (i >= 9 * 64 * 64) return;
  const int co = i & 63, ci = (i >> 6) & 63, tap = i >> 12;
  float acc = 0.f;
  for (int c = 0; c < nchunks; ++c) acc += ws[(int64_t)c * 36864 + i];
  dw[(co * 64 + ci) * 9 + tap] = acc;
}

int wgrad_row64_supported(int w) { return w >= 1 && ((w + 15) / 16 * 16) + 2 <= 256; }

int wgrad_row64(const bf16* x, int x_ctotal, int x_coff, const bf16* dy, int dy_ctotal, int dy_coff, int n, int h, int w, float* dw,
                float* scratch, size_t scratch_bytes, cudaStream_t s) {
  WgradRowPlan p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.h = h; p.w = w; p.kp = (w + 15) / 16 * 16;
  int cpi = DBB_NUM_SMS / n; if (cpi < 1) cpi = 1; if (cpi > h) cpi = h;
  p.rows_per_chunk = (h + cpi - 1) / cpi;
  p.chunks_per_image = (h + p.rows_per_chunk - 1) / p.rows_per_chunk;
  p.nchunks = n * p.chunks_per_image;
  if (!scratch || scratch_bytes < (size_t)p.nchunks * 36864 * sizeof(float)) return set_error(DBB_EWORKSPACE, "wgrad_row64: scratch too small");
  p.ws = scratch; p.dw = dw;
  int rc = encode_tmap_nhwc(&p.tmap_x, x, n, h, w, x_ctotal, x_coff, 64, 1, 1, p.kp + 2, 1, 1);
  if (rc) return rc;
  rc = encode_tmap_nhwc(&p.tmap_dy, dy, n, h, w, dy_ctotal, dy_coff, 64, 1, 1, p.kp, 1, 1);
  if (rc) return rc;
  const int xslot = ((p.kp + 2) * 128 + 1023) & ~1023;
  const int smem = WR_XSLOTS * xslot + WR_DSLOTS * p.kp * 128 + 1024 + 256 + 1024;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    DBB_CUDA(cudaFuncSetAttribute(wgrad_row64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const char* label = "wgrad_row64";
  if (prof_enabled()) {
    char tmp[96];
    snprintf(tmp, sizeof(tmp), "wgrad_nt64_m64_n64_t9_k%lld_row", (long long)n * h * w);
    label = prof_label(tmp);
  }
  DBB_LAUNCH(label, s, wgrad_row64_kernel<<<p.nchunks, IG_THREADS, smem, s>>>(p));
  DBB_LAUNCH("wgrad_reduce", s, wgrad_row64_reduce_kernel<<<(36864 + 255) / 256, 256, 0, s>>>(p.ws, p.nchunks, dw));
  return DBB_OK;
}

// ---------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------
constexpr int WG_BOX_BYTES = 64 * 128;   // 64 pixels x 64 channels bf16

template <int N_TILE, int STAGES>
struct WgradSmem {
  static constexpr int A_BYTES = 2 * WG_BOX_BYTES;              // M tile is always 128 channels (2 boxes)
  static constexpr int B_BYTES = (N_TILE / 64) * WG_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(IG_THREADS)
wgrad_kernel(const __grid_constant__ WgradPlan p) {
  using SM = WgradSmem<N_TILE, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.m_total + 127) / 128, n_tiles = (p.n_total + N_TILE - 1) / N_TILE;
  int b = blockIdx.x;
  const int split = b % p.split_k; b /= p.split_k;
  const int nt = b % n_tiles; b /= n_tiles;
  const int mt = b % m_tiles; b /= m_tiles;
  const int tap = b;
  const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
  const int t_begin = (int)((int64_t)total_tiles * split / p.split_k);
  const int t_end = (int)((int64_t)total_tiles * (split + 1) / p.split_k);
  const int kiters = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_a);
    prefetch_tmap(&p.tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<N_TILE>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (kiters > 0) {
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          int t = t_begin + it;
          const int tw = t % p.tiles_w; t /= p.tiles_w;
          const int th = t % p.tiles_h; t /= p.tiles_h;
          const int n0 = t * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
          uint8_t* sa = smem + s * SM::STAGE_BYTES;
          uint8_t* sb = sa + SM::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)SM::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sa + j * WG_BOX_BYTES, &p.tmap_a, &full_bar[s], mt * 128 + j * 64,
                        w0 * p.a_sw + p.a_dw[tap], h0 * p.a_sh + p.a_dh[tap], n0);
#pragma unroll
          for (int j = 0; j < N_TILE / 64; ++j)
            tma_load_4d(sb + j * WG_BOX_BYTES, &p.tmap_b, &full_bar[s], nt * N_TILE + j * 64,
                        w0 * p.b_sw + p.b_dw[tap], h0 * p.b_sh + p.b_dh[tap], n0);
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N_TILE, 1, 1);
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          // MN-major: 64-channel blocks LBO = 8192 B apart, 8-pixel K groups SBO = 1024 B apart
          const uint64_t da = make_desc_sw128(sa, WG_BOX_BYTES, 1024);
          const uint64_t db = make_desc_sw128(sa + SM::A_BYTES, WG_BOX_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 pixels per MMA = 2 K groups = 2048 B
            mma_bf16_ss(tmem_d, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (it | k) != 0);
          mma_commit(&empty_bar[s]);
        }
        mma_commit(tmem_full);
      }
    } else {
      const int q = warp & 3;
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < N_TILE; c += 16) {
        uint32_t v[16];
        tmem_ld_x16(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        // deterministic split-K: every CTA stores its partial tile, wgrad_reduce_kernel sums the splits in a fixed order
        float* dst = p.ws + (((int64_t)split * p.ntaps + tap) * p.m_pad + (mt * 128 + q * 32 + lane)) * p.n_pad + nt * N_TILE + c;
        float4* d4 = reinterpret_cast<float4*>(dst);
        d4[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        d4[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        d4[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
        d4[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<N_TILE>(tmem_d);
}

// ---------------------------------------------------------------------------------------------
// weight packing (fp32 parameters -> bf16 GEMM operand)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_one(int mode, const float* __restrict__ w, bf16* __restrict__ out, int co_n, int ci_n, int kh, int kw,
                                         int co_total, int co_off, int64_t start, int64_t stride) {
  const int taps = kh * kw;
  const int64_t total = (mode == 4) ? (int64_t)co_n * 256 : (int64_t)co_n * ci_n * taps;
  for (int64_t i = start; i < total; i += stride) {
    float v;
    if (mode == 0) {          // out[co][tap][ci] <- w[co][ci][tap]
      const int ci = i % ci_n; const int tap = (i / ci_n) % taps; const int co = i / ((int64_t)ci_n * taps);
      v = w[((int64_t)co * ci_n + ci) * taps + tap];
    } else if (mode == 1) {   // out[ci][tap][co] <- w[co][ci][tap]
      const int co = i % co_n; const int tap = (i / co_n) % taps; const int ci = i / ((int64_t)co_n * taps);
      v = w[((int64_t)co * ci_n + ci) * taps + tap];
      out[((int64_t)ci * taps + tap) * co_total + co_off + co] = __float2bfloat16_rn(v);
      continue;
    } else if (mode == 2) {   // out[cls][co][ci] <- w[ci][co][cls]   (ConvT weight is (ci, co, 2, 2))
      const int ci = i % ci_n; const int co = (i / ci_n) % co_n; const int cls = i / ((int64_t)ci_n * co_n);
      v = w[((int64_t)ci * co_n + co) * 4 + cls];
    } else if (mode == 3) {   // out[ci][cls][co] <- w[ci][co][cls]
      const int co = i % co_n; const int cls = (i / co_n) % 4; const int ci = i / ((int64_t)co_n * 4);
      v = w[((int64_t)ci * co_n + co) * 4 + cls];
      out[((int64_t)ci * 4 + cls) * co_total + co_off + co] = __float2bfloat16_rn(v);
      continue;
    } else {                  // conv1 7x7/2 in space-to-depth form: out[co][kh2][kw2][16], ch = (py*2+px)*3 + c (12 used)
      const int ch = i % 16; const int kw2 = (i / 16) % 4; const int kh2 = (i / 64) % 4; const int co = i / 256;
      v = 0.f;
      if (ch < 12) {
        const int c = ch % 3, px = (ch / 3) % 2, py = ch / 6;
        // s2d pixel (i+kh2-2, j+kw2-2), sub-pixel (py,px) <-> original row 2(i+kh2-2)+py = 2i + (2*kh2+py-4) = 2i - 3 + kh
        const int kh_ = 2 * kh2 + py - 1, kw_ = 2 * kw2 + px - 1;
        if (kh_ >= 0 && kh_ < 7 && kw_ >= 0 && kw_ < 7) v = w[(((int64_t)co * 3 + c) * 7 + kh_) * 7 + kw_];
      }
    }
    out[i] = __float2bfloat16_rn(v);
  }
}
__global__ void pack_weights_kernel(int mode, const float* __restrict__ w, bf16* __restrict__ out, int co_n, int ci_n, int kh, int kw,
                                    int co_total, int co_off) {
  pack_one(mode, w, out, co_n, ci_n, kh, kw, co_total, co_off, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}
// mode 5: Conv2d OIHW -> BOTH GEMM layouts from one coalesced read: a block stages a 32(co) x 32(ci) x taps tile in shared
// memory, then writes out[co][tap][ci] (mode 0 layout) and, if out2, out2[ci][tap][co_off + co] (mode 1 layout) in 64-byte runs.
constexpr int PACK_TILE_MAX_TAPS = 9;
__device__ __forceinline__ void pack_tile(const PackJob& j, int t, float* tile /* [32][32*taps + 1] */) {
  const int taps = j.kh * j.kw;
  const int row = 32 * taps + 1;
  const int tiles_ci = j.ci_n / 32;
  const int co0 = (t / tiles_ci) * 32, ci0 = (t % tiles_ci) * 32;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * 32 * taps; idx += 256) {
    const int co = idx / (32 * taps), rem = idx - co * 32 * taps;       // rem = ci * taps + tap: contiguous in w
    tile[co * row + rem] = j.w[((int64_t)(co0 + co) * j.ci_n + ci0) * taps + rem];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * 32 * taps; idx += 256) {
    const int ci = idx & 31, r = idx >> 5;
    const int tap = r % taps, co = r / taps;
    j.out[((int64_t)(co0 + co) * taps + tap) * j.ci_n + ci0 + ci] = __float2bfloat16_rn(tile[co * row + ci * taps + tap]);
  }
  if (j.out2) {
    for (int idx = threadIdx.x; idx < 32 * 32 * taps; idx += 256) {
      const int co = idx & 31, r = idx >> 5;
      const int tap = r % taps, ci = r / taps;
      j.out2[((int64_t)(ci0 + ci) * taps + tap) * j.co_total + j.co_off + co0 + co] = __float2bfloat16_rn(tile[co * row + ci * taps + tap]);
    }
  }
}
// One flat list of work units over all jobs (tile_start = prefix sums): a tiled job contributes (co/32)*(ci/32) tiles, any
// other job is one unit, so blocks are spread in proportion to the work instead of 64 per job.
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const __grid_constant__ PackBatch b) {
  __shared__ float tile[32 * (32 * PACK_TILE_MAX_TAPS + 1)];
  const int total = b.tile_start[b.njobs];
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int ji = 0;
    while (t >= b.tile_start[ji + 1]) ++ji;
    const PackJob& j = b.jobs[ji];
    if (j.mode == 5) pack_tile(j, t - b.tile_start[ji], tile);
    else pack_one(j.mode, j.w, j.out, j.co_n, j.ci_n, j.kh, j.kw, j.co_total, j.co_off, threadIdx.x, 256);
  }
}
int pack_weights_batch(const PackBatch& b_in, cudaStream_t s) {
  if (b_in.njobs <= 0) return DBB_OK;
  PackBatch b = b_in;
  b.tile_start[0] = 0;
  for (int i = 0; i < b.njobs; ++i) {
    const PackJob& j = b.jobs[i];
    int units = 1;
    if (j.mode == 5) {
      if (j.co_n % 32 || j.ci_n % 32 || j.kh * j.kw > PACK_TILE_MAX_TAPS)
        return set_error(DBB_EUNSUPPORTED, "pack_weights: tiled mode needs channels % 32 == 0 and <= 9 taps");
      units = (j.co_n / 32) * (j.ci_n / 32);
    }
    b.tile_start[i + 1] = b.tile_start[i] + units;
  }
  int grid = b.tile_start[b.njobs];
  if (grid > DBB_NUM_SMS * 6) grid = DBB_NUM_SMS * 6;
  DBB_LAUNCH("pack_weights_batch", s, pack_weights_batch_kernel<<<grid, 256, 0, s>>>(b));
  return DBB_OK;
}

int pack_weights(int mode, const float* w, bf16* out, int co_n, int ci_n, int kh, int kw, cudaStream_t s, int co_total, int co_off) {
  if (co_total <= 0) { co_total = co_n; co_off = 0; }
  const int64_t total = (mode == 4) ? (int64_t)co_n * 256 : (int64_t)co_n * ci_n * kh * kw;
  int grid = (int)((total + 255) / 256);
  if (grid > DBB_NUM_SMS * 8) grid = DBB_NUM_SMS * 8;
  DBB_LAUNCH("pack_weights", s, pack_weights_kernel<<<grid, 256, 0, s>>>(mode, w, out, co_n, ci_n, kh, kw, co_total, co_off));
  return DBB_OK;
}

// ---------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  });
  return g_encode;
}

int encode_tmap_nhwc(CUtensorMap* m, const bf16* base, int n, int h, int w, int c_total, int c_off, int c_extent,
                     int box_n, int box_h, int box_w, int stride_h, int stride_w) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(DBB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)c_extent, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c_total * 2, (cuuint64_t)w * c_total * 2, (cuuint64_t)h * w * c_total * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(box_w * stride_w), (cuuint32_t)(box_h * stride_h), (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)(base + c_off), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled(nhwc n=%d h=%d w=%d c=%d box=%d,%d,%d stride=%d,%d) failed: %d",
             n, h, w, c_total, box_n, box_h, box_w, stride_h, stride_w, (int)r);
    return DBB_ECUDA;
  }
  return DBB_OK;
}

// generic 4-D map with explicit strides (used for the overlapping conv1 space-to-depth view)
int encode_tmap_raw4(CUtensorMap* m, const bf16* base, const cuuint64_t dims[4], const cuuint64_t strides_bytes[3],
                     const cuuint32_t box[4]) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(DBB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled(raw4) failed: %d", (int)r);
    return DBB_ECUDA;
  }
  return DBB_OK;
}

static int encode_tmap_weights(CUtensorMap* m, const bf16* wp, int k_total, int rows, int block_n) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(DBB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "cuTensorMapEncodeTiled(weights k=%d rows=%d) failed: %d", k_total, rows, (int)r);
    return DBB_ECUDA;
  }
  return DBB_OK;
}

int encode_weights_public(CUtensorMap* m, const bf16* wp, int k_total, int rows, int block_n) {
  return encode_tmap_weights(m, wp, k_total, rows, block_n);
}

// pick a (bn, bh, bw) box of exactly `rows` pixels (power-of-two factors) with the least overhang
void choose_box(int rows, int mn, int mh, int mw, int* bn, int* bh, int* bw) {
  double best = 1e30; int bbn = 1, bbh = 1, bbw = rows;
  for (int w = rows; w >= 1; w >>= 1) {
    for (int h = rows / w; h >= 1; h >>= 1) {
      const int n = rows / (w * h);
      const double cover = (double)((mw + w - 1) / w * w) * ((mh + h - 1) / h * h) * ((mn + n - 1) / n * n);
      // tie-break towards wide boxes (longer contiguous runs per TMA row)
      const double score = cover * (1.0 + 1e-6 * (rows / w));
      if (score < best) { best = score; bbn = n; bbh = h; bbw = w; }
    }
  }
  *bn = bbn; *bh = bbh; *bw = bbw;
}

int igemm_plan_init(IgemmPlan* p, const bf16* x, int n, int h, int w, int c_total, int c_off, int cin,
                    const bf16* wp, int k_total, int w_rows, int block_n) {
  // caller has filled: mn, mh, mw, in_sh, in_sw, ntaps, dh, dw, cout, out_*; y, bias
  p->cin = cin;
  p->block_n = block_n;
  choose_box(128, p->mn, p->mh, p->mw, &p->bn, &p->bh, &p->bw);
  p->tiles_n = (p->mn + p->bn - 1) / p->bn;
  p->tiles_h = (p->mh + p->bh - 1) / p->bh;
  p->tiles_w = (p->mw + p->bw - 1) / p->bw;
  int rc = encode_tmap_nhwc(&p->tmap_x, x, n, h, w, c_total, c_off, cin, p->bn, p->bh, p->bw, p->in_sh, p->in_sw);
  if (rc) return rc;
  return encode_tmap_weights(&p->tmap_w, wp, k_total, w_rows, block_n);
}

template <int BLOCK_N, int STAGES>
static int igemm_launch_t(const IgemmPlan& p, cudaStream_t s) {
  using SM = IgemmSmem<BLOCK_N, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    DBB_CUDA(cudaFuncSetAttribute(igemm_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_done = true;
  }
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int64_t grid = (int64_t)p.tiles_n * p.tiles_h * p.tiles_w * n_blocks;
  if (grid <= 0 || grid > 0x7fffffff) return set_error(DBB_EINVAL, "igemm: bad grid");
  const char* label = "igemm";
  if (prof_enabled()) {
    char tmp[96];   // label carries the GEMM shape: M pixels, N channels, K = taps*cin
    snprintf(tmp, sizeof(tmp), "igemm_bn%d_m%lld_n%d_k%d", BLOCK_N, (long long)p.mn * p.mh * p.mw, p.cout, p.ntaps * p.cin);
    label = prof_label(tmp);
  }
  DBB_LAUNCH(label, s, igemm_kernel<BLOCK_N, STAGES><<<(unsigned)grid, IG_THREADS, SM::TOTAL, s>>>(p));
  return DBB_OK;
}

template <int BLOCK_N, int STAGES, bool STATS>
static int igemm_launch_p2(const IgemmPlan& p, cudaStream_t s) {
  using SM = IgemmPSmem<BLOCK_N, STAGES>;
  static bool attr_done = false;
  static int occ = 1;
  if (!attr_done) {
    DBB_CUDA(cudaFuncSetAttribute(igemm_persist_kernel<BLOCK_N, STAGES, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    // resident CTAs per SM: registers + shared memory (runtime query) and TMEM (512 columns, 2*BLOCK_N per CTA)
    // (the occupancy API answers for the default shared-memory carve-out, i.e. 1; computed from the limits instead)
    cudaFuncAttributes fa;
    DBB_CUDA(cudaFuncGetAttributes(&fa, igemm_persist_kernel<BLOCK_N, STAGES, STATS>));
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * IGP_THREADS;
    occ = (227 * 1024) / (SM::TOTAL + 1024);
    if (occ > 65536 / regs_per_cta) occ = 65536 / regs_per_cta;
    if (occ > 2048 / IGP_THREADS) occ = 2048 / IGP_THREADS;
    if (occ > 512 / (2 * BLOCK_N)) occ = 512 / (2 * BLOCK_N);
    if (occ < 1) occ = 1;
    attr_done = true;
  }
  const int n_blocks = (p.cout + BLOCK_N - 1) / BLOCK_N;
  const int64_t total = (int64_t)p.tiles_n * p.tiles_h * p.tiles_w * n_blocks;
  if (total <= 0 || total > 0x7fffffff) return set_error(DBB_EINVAL, "igemm: bad grid");
  int64_t grid = (int64_t)DBB_NUM_SMS * occ;
  grid -= grid % n_blocks;
  if (grid > total) grid = total;             // total is a multiple of n_blocks
  const char* label = "igemm";
  if (prof_enabled()) {
    char tmp[96];
    snprintf(tmp, sizeof(tmp), "igemm_bn%d_m%lld_n%d_k%d%s", BLOCK_N, (long long)p.mn * p.mh * p.mw, p.cout, p.ntaps * p.cin, STATS ? "_st" : "");
    label = prof_label(tmp);
  }
  DBB_LAUNCH(label, s, igemm_persist_kernel<BLOCK_N, STAGES, STATS><<<(unsigned)grid, IGP_THREADS, SM::TOTAL, s>>>(p));
  return DBB_OK;
}
int igemm_launch(const IgemmPlan& p_in, cudaStream_t s) {
  IgemmPlan p = p_in;
  static const bool no_staged = getenv("DBB_NO_STAGED_EPI") != nullptr;      // A/B switch
  p.no_staged_epilogue = no_staged ? 1 : 0;
  if (p.cin % 64 != 0 || p.ntaps < 1 || p.ntaps > IGEMM_MAX_TAPS) return set_error(DBB_EUNSUPPORTED, "igemm: cin must be a multiple of 64, taps <= 16");
  // The k-loop of a tile has ntaps*cin/64 iterations; short loops (1x1 convolutions, ConvTranspose, conv1) get a
  // shallow ring so that more CTAs (= more epilogue warps) are resident per SM.
  const int kiters = p.ntaps * (p.cin >> 6);
  static const bool deep_only = getenv("DBB_DEEP_RING") != nullptr;      // A/B switches
  static const bool no_persist = getenv("DBB_NO_PERSIST") != nullptr;
  const int depth = deep_only ? 4 : (kiters <= 1 ? 1 : (kiters <= 4 ? 2 : 4));
  if (!no_persist || p.st.enabled || p.epi.scale) {
    const bool st = p.st.enabled != 0;
    switch (p.block_n) {
      case 64:
        // statistics in registers (64 accumulators/thread) -> 1 CTA/SM, so that CTA gets a deep ring
        if (st) {
          static const bool deep_st = getenv("DBB_ST64_DEEP") != nullptr;                                    // A/B switch
          if (!deep_st && depth <= 2) return igemm_launch_p2<64, 2, true>(p, s);                             // lean epilogue, 2 CTAs/SM (a 3-stage ring measured slower for conv1: 0.178 vs 0.132 ms)
          return igemm_launch_p2<64, 6, true>(p, s);                                                         // 144 KB (2 CTAs/SM at 96 registers + 3 stages measured slower)
        }
        if (depth <= 2) return igemm_launch_p2<64, 2, false>(p, s);                                          // 48 KB -> 4 CTAs/SM
        return igemm_launch_p2<64, 4, false>(p, s);                                                          // 96 KB -> 2 CTAs/SM
      case 128:
        if (depth <= 2) return st ? igemm_launch_p2<128, 2, true>(p, s) : igemm_launch_p2<128, 2, false>(p, s);   // 2 CTAs/SM (TMEM)
        return st ? igemm_launch_p2<128, 3, true>(p, s) : igemm_launch_p2<128, 3, false>(p, s);
      case 256:
        if (depth <= 2) return st ? igemm_launch_p2<256, 2, true>(p, s) : igemm_launch_p2<256, 2, false>(p, s);   // 1 CTA/SM (TMEM)
        return st ? igemm_launch_p2<256, 4, true>(p, s) : igemm_launch_p2<256, 4, false>(p, s);
      default: return set_error(DBB_EUNSUPPORTED, "igemm: block_n must be 64, 128 or 256");
    }
  }
  switch (p.block_n) {
    case 64:
      if (depth == 1) return igemm_launch_t<64, 1>(p, s);     // 24 KB -> 8 CTAs/SM
      if (depth == 2) return igemm_launch_t<64, 2>(p, s);     // 48 KB -> 4 CTAs/SM
      return igemm_launch_t<64, 4>(p, s);                     // 96 KB -> 2 CTAs/SM
    case 128:
      if (depth == 1) return igemm_launch_t<128, 1>(p, s);    // 32 KB -> 4 CTAs/SM (TMEM)
      if (depth == 2) return igemm_launch_t<128, 2>(p, s);    // 64 KB -> 3 CTAs/SM
      return igemm_launch_t<128, 3>(p, s);                    // 96 KB -> 2 CTAs/SM
    case 256:
      if (depth == 1) return igemm_launch_t<256, 1>(p, s);    // 48 KB -> 2 CTAs/SM (TMEM)
      if (depth == 2) return igemm_launch_t<256, 2>(p, s);    // 96 KB -> 2 CTAs/SM
      return igemm_launch_t<256, 4>(p, s);                    // 192 KB -> 1 CTA/SM
    default: return set_error(DBB_EUNSUPPORTED, "igemm: block_n must be 64, 128 or 256");
  }
}

// dw[m][n][tap] = sum_split ws[split][tap][m][n]   (fixed summation order -> bitwise reproducible gradients)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int split_k, int ntaps, int m_pad, int n_pad,
                                                           int m_total, int n_total, int tap_stride, float* __restrict__ dw) {
  const int64_t total = (int64_t)ntaps * m_total * n_total;
  const int64_t plane = (int64_t)ntaps * m_pad * n_pad;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int n = (int)(i % n_total); int64_t t = i / n_total;
    const int m = (int)(t % m_total); const int tap = (int)(t / m_total);
    const float* src = ws + ((int64_t)tap * m_pad + m) * n_pad + n;
    float acc = 0.f;
    for (int sp = 0; sp < split_k; ++sp) acc += src[(int64_t)sp * plane];
    dw[((int64_t)m * n_total + n) * tap_stride + tap] = acc;
  }
}

// few outputs, many splits (ConvTranspose / conv1 / 1x1 gradients): a block sums 32 consecutive outputs, warp w takes
// splits w, w+8, ... and the eight partial sums are combined in a fixed order (still bitwise reproducible)
__global__ void __launch_bounds__(256) wgrad_reduce_wide_kernel(const float* __restrict__ ws, int split_k, int ntaps, int m_pad, int n_pad,
                                                                int m_total, int n_total, int tap_stride, float* __restrict__ dw) {
  __shared__ float part[8][32];
  const int64_t plane = (int64_t)ntaps * m_pad * n_pad;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;            // n_total % 32 == 0: the 32 outputs share (tap, m)
  const int n = (int)(i % n_total); const int64_t t = i / n_total;
  const int m = (int)(t % m_total); const int tap = (int)(t / m_total);
  const float* src = ws + ((int64_t)tap * m_pad + m) * n_pad + n;
  float acc = 0.f;
  for (int sp = wp; sp < split_k; sp += 8) acc += src[(int64_t)sp * plane];
  part[wp][lane] = acc;
  __syncthreads();
  if (wp == 0) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += part[k][lane];
    dw[((int64_t)m * n_total + n) * tap_stride + tap] = r;
  }
}

size_t wgrad_scratch_bytes(const WgradPlan& p) {
  return (size_t)p.split_k * p.ntaps * p.m_pad * p.n_pad * sizeof(float);
}

template <int N_TILE, int STAGES>
static int wgrad_launch_t(const WgradPlan& p, cudaStream_t s) {
  using SM = WgradSmem<N_TILE, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    DBB_CUDA(cudaFuncSetAttribute(wgrad_kernel<N_TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_done = true;
  }
  const int m_tiles = (p.m_total + 127) / 128, n_tiles = (p.n_total + N_TILE - 1) / N_TILE;
  const int64_t grid = (int64_t)p.ntaps * m_tiles * n_tiles * p.split_k;
  if (grid <= 0 || grid > 0x7fffffff) return set_error(DBB_EINVAL, "wgrad: bad grid");
  const char* label = "wgrad";
  if (prof_enabled()) {
    char tmp[96];   // M x N per tap, K = pixels reduced
    snprintf(tmp, sizeof(tmp), "wgrad_nt%d_m%d_n%d_t%d_k%lld", N_TILE, p.m_total, p.n_total, p.ntaps, (long long)p.mn * p.mh * p.mw);
    label = prof_label(tmp);
  }
  DBB_LAUNCH(label, s, wgrad_kernel<N_TILE, STAGES><<<(unsigned)grid, IG_THREADS, SM::TOTAL, s>>>(p));
  const int64_t total = (int64_t)p.ntaps * p.m_total * p.n_total;
  int rgrid = (int)((total + 255) / 256);
  if (rgrid > DBB_NUM_SMS * 8) rgrid = DBB_NUM_SMS * 8;
  const char* rlabel = "wgrad_reduce";
  if (prof_enabled()) {
    char tmp[96];
    snprintf(tmp, sizeof(tmp), "wgrad_reduce_s%d_t%d_m%d_n%d", p.split_k, p.ntaps, p.m_total, p.n_total);
    rlabel = prof_label(tmp);
  }
  if (p.split_k >= 16 && p.n_total % 32 == 0 && total / 32 <= 65535 * 16) {
    DBB_LAUNCH(rlabel, s, wgrad_reduce_wide_kernel<<<(unsigned)(total / 32), 256, 0, s>>>(p.ws, p.split_k, p.ntaps, p.m_pad, p.n_pad, p.m_total, p.n_total, p.tap_stride, p.dw));
    return DBB_OK;
  }
  DBB_LAUNCH(rlabel, s, wgrad_reduce_kernel<<<rgrid, 256, 0, s>>>(p.ws, p.split_k, p.ntaps, p.m_pad, p.n_pad, p.m_total, p.n_total, p.tap_stride, p.dw));
  return DBB_OK;
}

int wgrad_launch(const WgradPlan& p, cudaStream_t s) {
  switch (p.n_tile) {
    case 64: return wgrad_launch_t<64, 4>(p, s);     // 24 KB/stage
    case 128: return wgrad_launch_t<128, 3>(p, s);   // 32 KB/stage
    case 256: return wgrad_launch_t<256, 4>(p, s);   // 48 KB/stage
    default: return set_error(DBB_EUNSUPPORTED, "wgrad: n_tile must be 64, 128 or 256");
  }
}

}  // namespace dbb
