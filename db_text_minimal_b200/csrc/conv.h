// conv.h -- host-side interface of the tcgen05 implicit-GEMM kernels (conv_tcgen05.cu), used by net.cu
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace dbb {

typedef __nv_bfloat16 bf16;

constexpr int IGEMM_MAX_TAPS = 16;

// Training-mode BatchNorm statistics of a convolution output, fused into the convolution's epilogue: per-channel sum and
// sum of squares of the bf16-rounded outputs -> fp64 atomics into gacc ([2][c] doubles, zero on entry, left zero) -> the
// last CTA to arrive (ticket in `counter`, zero on entry, left zero) writes stats4 = [scale | shift | mean | invstd] and
// the running statistics.  Up to two parameter segments (the two head branches share one 128-channel tensor).
struct BnFinSeg { const float* gamma; const float* beta; float* rmean; float* rvar; int coff, cn; };
struct BnFin { BnFinSeg seg[2]; int nseg; float momentum, eps; float* stats4; };
struct ConvStats {
  int enabled;
  double count;            // pixels per channel of the normalised tensor (N*H*W of y)
  double* gacc;
  unsigned* counter;
  BnFin fin;               // channel indices are positions in y's channel dimension (width out_c)
};

// Inference-mode epilogue: y = [relu]((acc + bias) * scale[ch] + shift[ch] [+ res]) -- BatchNorm with fixed statistics (and
// the residual add / ReLU that follow it) applied to the fp32 accumulator, so eval-mode forward passes write activations
// directly and run no separate BatchNorm pass.  res: compact NHWC tensor of the output's shape (channel stride = cout).
struct ConvEpi { const float* scale; const float* shift; const bf16* res; int relu; };

// One implicit-GEMM launch:  Y[pixel, co] = sum_{tap, ci} X[pixel @ tap, ci] * Wp[co, tap*cin + ci]  (+ bias[co])
//
// "pixel" runs over an M-space grid (mn, mh, mw); row (n,h,w) reads X at (n, h*in_sh + dh[tap], w*in_sw + dw[tap])
// (zero outside the tensor) and writes Y at (n, h*out_sh + out_oh, w*out_sw + out_ow).  This one form covers Conv2d
// fprop (any stride), Conv2d dgrad (stride-1 directly; stride-2 as parity classes), ConvTranspose2d(k2,s2) fprop
// (4 classes) and dgrad.
struct IgemmPlan {
  CUtensorMap tmap_x;      // 4-D {C, W, H, N} bf16 NHWC, box {64, bw*in_sw, bh*in_sh, bn}, 128B swizzle
  CUtensorMap tmap_w;      // 2-D {K, Cout} bf16 K-major, box {64, block_n}, 128B swizzle
  int mn, mh, mw;          // M-space extent
  int bn, bh, bw;          // tile box, bn*bh*bw <= 128
  int tiles_n, tiles_h, tiles_w;
  int in_sh, in_sw;
  int ntaps;
  int8_t dh[IGEMM_MAX_TAPS], dw[IGEMM_MAX_TAPS];
  uint8_t wtap[IGEMM_MAX_TAPS];   // index of tap t inside the packed weight matrix (K offset = wtap[t]*cin)
  int cin;                 // channels reduced per tap (multiple of 64)
  int cout;                // valid output channels (GEMM N)
  int block_n;             // 64 / 128 / 256
  // output
  bf16* y;
  int out_h, out_w, out_c; // Y tensor extent (NHWC), out_c = channel stride
  int out_sh, out_sw, out_oh, out_ow;
  int out_coff;            // channel offset inside Y's channel dimension
  const float* bias;       // may be null
  int accumulate;          // 1: Y += result (gradient fan-in), 0: overwrite
  int cls_cols;            // > 0: pixel-shuffle epilogue, GEMM column = class*cls_cols + channel (ConvTranspose2d k2 s2)
  ConvStats st;            // optional fused BatchNorm statistics of y
  ConvEpi epi;             // optional inference epilogue (scale != nullptr)
  int no_staged_epilogue;  // A/B switch (DBB_NO_STAGED_EPI): direct per-thread stores in the 64-wide kernels
};

// weight packing: fp32 parameter -> bf16 GEMM B matrix [rows][K] (K-major)
//   mode 0: Conv2d OIHW (co,ci,kh,kw)       -> [co][(kh*KW+kw)*ci_n + ci]                 (fprop)
//   mode 1: Conv2d OIHW                      -> [ci][(kh*KW+kw)*co_n + co]                 (dgrad)
//   mode 2: ConvT   (ci,co,2,2)              -> [cls=(a*2+b)][co][ci]   4 matrices         (fprop, one per class)
//   mode 3: ConvT   (ci,co,2,2)              -> [ci][(a*2+b)*co_n + co]                    (dgrad)
//   mode 4: conv1 7x7/2 OIHW (64,3,7,7)      -> [co][kh2(4)][kw2(4)][16]  space-to-depth form, K = 256
// co_total / co_off: (modes 1 and 3) the packed matrix is co_total wide and this tensor fills columns [co_off, co_off+co_n)
int pack_weights(int mode, const float* w, bf16* out, int co_n, int ci_n, int kh, int kw, cudaStream_t s, int co_total = 0, int co_off = 0);

// batched packing: all weight tensors of the network in one or two launches
//   mode 5 (batch only): Conv2d OIHW -> out in the mode 0 layout AND (if out2) out2 in the mode 1 layout, tiled through
//           shared memory so that both the read and the two writes are coalesced (co_total / co_off apply to out2)
struct PackJob { const float* w; bf16* out; int mode, co_n, ci_n, kh, kw, co_total, co_off; bf16* out2; };
constexpr int PACK_BATCH = 48;
struct PackBatch { int njobs; PackJob jobs[PACK_BATCH]; int tile_start[PACK_BATCH + 1]; };   // tile_start: filled by pack_weights_batch
int pack_weights_batch(const PackBatch& b, cudaStream_t s);

int igemm_plan_init(IgemmPlan* p, const bf16* x, int n, int h, int w, int c_total, int c_off, int cin,
                    const bf16* wp, int k_total, int w_rows, int block_n);
int igemm_launch(const IgemmPlan& p, cudaStream_t s);

// 3x3 / stride 1 / 64 -> 64 channels: persistent, weights-stationary, shifted-window ("halo") implicit GEMM.
// One (R+2) x (bw+2) activation box per tile feeds all nine taps through descriptor start offsets; the 72 KB of packed
// weights stay in shared memory for the life of the CTA; TMEM accumulators are double buffered.
struct HaloPlan {
  CUtensorMap tmap_x;      // 4-D NHWC, box {64, bw+2, R+2, 1}
  CUtensorMap tmap_w;      // 2-D {K = 576, 64 rows}, box {64, 64}
  int n, h, w;             // output (= input) extent
  int R, bw, pitch;        // tile rows / cols, pitch = bw + 2
  int tiles_h, tiles_w, total_tiles;
  int16_t off[9];          // smem row offset of tap t inside the halo box
  uint8_t wtap[9];         // weight tile used with tap t
  bf16* y; int out_c, out_coff;
  const float* bias;
  int accumulate;
  int base_offset_mode;    // descriptor base_offset: 1 = (addr >> 7) & 7, 0 = always 0 (tuning / bring-up)
  ConvStats st;            // optional fused BatchNorm statistics of y
  ConvEpi epi;             // optional inference epilogue (scale != nullptr)
};
int halo64_supported(int h, int w);
int halo64_plan(HaloPlan* p, const bf16* x, int n, int h, int w, int x_ctotal, int x_coff, const bf16* wp, int dgrad);
int halo64_launch(const HaloPlan& p, cudaStream_t s);

// weight gradient of the 3x3 / stride 1 / 64 -> 64 layers: persistent CTAs walk down a strip of image rows keeping a ring
// of activation rows in shared memory; every row of x and dy is fetched exactly once; the nine taps are descriptor start
// offsets into the ring (two horizontally adjacent taps share one M = 128 MMA).
struct WgradRowPlan {
  CUtensorMap tmap_x;      // 4-D NHWC, box {64, KP + 2, 1, 1}
  CUtensorMap tmap_dy;     // 4-D NHWC, box {64, KP, 1, 1}
  int n, h, w, kp;         // kp = W rounded up to 16
  int rows_per_chunk, chunks_per_image, nchunks;
  float* ws;               // partials [chunk][tap][ci][co]
  float* dw;               // OIHW fp32
};
int wgrad_row64_supported(int w);
int wgrad_row64(const bf16* x, int x_ctotal, int x_coff, const bf16* dy, int dy_ctotal, int dy_coff, int n, int h, int w, float* dw,
                float* scratch, size_t scratch_bytes, cudaStream_t s);

// weight gradient:  dW[m][n][tap] (+)= sum_pixel A[pixel @ (a-side map)][m] * B[pixel @ tap][n]
struct WgradPlan {
  CUtensorMap tmap_a;      // M-side operand (dy for Conv2d, x for ConvT): 4-D NHWC, box {64, bw*a_sw, bh*a_sh, bn}
  CUtensorMap tmap_b;      // N-side operand
  int mn, mh, mw;          // pixel grid reduced over
  int bn, bh, bw;          // 64-pixel box
  int tiles_n, tiles_h, tiles_w;
  int a_sh, a_sw, b_sh, b_sw;
  int ntaps;
  int8_t a_dh[IGEMM_MAX_TAPS], a_dw[IGEMM_MAX_TAPS], b_dh[IGEMM_MAX_TAPS], b_dw[IGEMM_MAX_TAPS];
  int m_total, n_total;    // GEMM M (rows of dW), N
  int m_tile, n_tile;      // 64|128, 64|128|256
  int split_k;
  float* dw;               // fp32, layout [m][n][tap] (== OIHW for Conv2d, (ci,co,a,b) for ConvT); overwritten
  float* ws;               // split-K partials [split][tap][m_pad][n_pad] fp32
  int m_pad, n_pad;
  int tap_stride;          // ntaps of the destination layout
};
int wgrad_launch(const WgradPlan& p, cudaStream_t s);
size_t wgrad_scratch_bytes(const WgradPlan& p);
constexpr size_t WGRAD_SCRATCH_BYTES = 96ull << 20;   // shared split-K scratch the executor reserves

int encode_tmap_raw4(CUtensorMap* m, const bf16* base, const cuuint64_t dims[4], const cuuint64_t strides_bytes[3],
                     const cuuint32_t box[4]);
int encode_tmap_nhwc(CUtensorMap* m, const bf16* base, int n, int h, int w, int c_total, int c_off, int c_extent,
                     int box_n, int box_h, int box_w, int stride_h, int stride_w);
int encode_weights_public(CUtensorMap* m, const bf16* wp, int k_total, int rows, int block_n);
void choose_box(int rows, int mn, int mh, int mw, int* bn, int* bh, int* bw);

}  // namespace dbb
