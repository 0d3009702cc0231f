// conv_ops.h -- convolution flavours of the DB network expressed as igemm / wgrad plans
#pragma once
#include "conv.h"
#include <string.h>

namespace dbb {

// h, w are the spatial extent of the (forward) INPUT; for ConvTranspose2d(k2,s2) the output is (2h, 2w)
struct ConvGeom {
  int n, h, w, cin, cout, ks, stride, pad;
  int out_h() const { return (h + 2 * pad - ks) / stride + 1; }
  int out_w() const { return (w + 2 * pad - ks) / stride + 1; }
};

// wp: packed by pack_weights mode 0 ; x may be a channel slice [x_coff, x_coff+cin) of a wider NHWC tensor
// st (optional): fuse the training-mode BatchNorm statistics of y into the epilogue (conv.h: ConvStats; count is filled here)
int conv_fprop(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* wp, const float* bias, bf16* y,
               int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st = nullptr, const ConvEpi* epi = nullptr);
// wp_t: packed by pack_weights mode 1
// dy may be a channel slice [dy_coff, dy_coff+cout) of a dy_ctotal-wide tensor (0 = compact); accumulate: dx += result
int conv_dgrad(const ConvGeom& g, const bf16* dy, const bf16* wp_t, bf16* dx, cudaStream_t s, int accumulate = 0);
int conv_wgrad(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* dy, int dy_ctotal, int dy_coff, float* dw,
               float* scratch, size_t scratch_bytes, cudaStream_t s);
// ConvTranspose2d(k=2, s=2); wp_cls: mode 2, wp_t: mode 3
int convt_fprop(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* wp_cls, const float* bias, bf16* y,
                int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st = nullptr);
int convt_dgrad(const ConvGeom& g, const bf16* dy, int dy_ctotal, int dy_coff, const bf16* wp_t, bf16* dx, int dx_ctotal,
                int dx_coff, cudaStream_t s);
int convt_wgrad(const ConvGeom& g, const bf16* x, int x_ctotal, int x_coff, const bf16* dy, int dy_ctotal, int dy_coff, float* dw,
                float* scratch, size_t scratch_bytes, cudaStream_t s);
// conv1 (7x7/2, 3->64) on the space-to-depth staging buffer (elementwise.h: image_to_s2d); wp: pack mode 4
int conv1_fprop(int n, int h, int w, const bf16* s2d, const bf16* wp, bf16* y, cudaStream_t s, const ConvStats* st = nullptr);
// dw_s2d: [64][64][4] fp32 scratch, unpacked by conv1_wgrad_unpack
int conv1_wgrad(int n, int h, int w, const bf16* s2d, const bf16* dy, float* dw_s2d, float* scratch, size_t scratch_bytes, cudaStream_t s);


// ---- float32 overloads (conv_f32.cu): the executor's fp32-parity mode.  Same argument meaning, except that the weight
// operand is the RAW float32 parameter (OIHW for Conv2d, (ci,co,2,2) for ConvTranspose2d) -- there is no packing step --
// and fused statistics / inference epilogues are not available (st / epi must be null).
int conv_fprop(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* w, const float* bias, float* y,
               int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st = nullptr, const ConvEpi* epi = nullptr);
int conv_dgrad(const ConvGeom& g, const float* dy, const float* w, float* dx, cudaStream_t s, int accumulate = 0);
int conv_wgrad(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* dy, int dy_ctotal, int dy_coff, float* dw,
               float* scratch, size_t scratch_bytes, cudaStream_t s);
int convt_fprop(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* w, const float* bias, float* y,
                int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st = nullptr);
int convt_dgrad(const ConvGeom& g, const float* dy, int dy_ctotal, int dy_coff, const float* w, float* dx, int dx_ctotal,
                int dx_coff, cudaStream_t s);
int convt_wgrad(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* dy, int dy_ctotal, int dy_coff, float* dw,
                float* scratch, size_t scratch_bytes, cudaStream_t s);
// conv1 (7x7/2, 3->64) straight from the NCHW float32 image; y / dy NHWC float32; wt / dw OIHW (64,3,7,7)
int conv1_fprop_f32(int n, int h, int w, const float* img, const float* wt, float* y, cudaStream_t s);
int conv1_wgrad_f32(int n, int h, int w, const float* img, const float* dy, float* dw, float* scratch, size_t scratch_bytes, cudaStream_t s);

}  // namespace dbb
