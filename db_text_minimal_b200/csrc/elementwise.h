// elementwise.h -- host interface of the memory-bound glue kernels (elementwise.cu).
// Activations are NHWC bf16; "P" is the pixel count N*H*W; channel slices of wider tensors are (ctotal, coff).
#pragma once
#include "common.cuh"
#include "conv.h"

namespace dbb {

// ND<T>: non-deduced pointer parameters (the first activation pointer of a call fixes T; nullptr is allowed for the rest).
// T is bf16 (product path) or float (fp32-parity mode of the executor: same kernels, float activations).
template <typename T> struct NdId { typedef T type; };
template <typename T> using ND = typename NdId<T>::type;

int nchw_f32_to_nhwc_bf16(const float* x, bf16* y, int n, int c, int64_t hw, cudaStream_t s);
int nhwc_bf16_to_nchw_f32(const bf16* x, float* y, int n, int c, int64_t hw, cudaStream_t s);
template <typename T> int nhwc_to_nchw_f32(const T* x, float* y, int n, int c, int64_t hw, cudaStream_t s);

// ---- BatchNorm2d (eps 1e-5, momentum 0.1; src/modules/basic.py:34, resnet.py:74,83, segmentation_head.py:26,28,68,74)
constexpr int BN_MAX_BLOCKS = DBB_NUM_SMS * 4;
// scratch floats needed for the per-block partials of one reduction
inline size_t bn_partials_floats(int c) { return (size_t)BN_MAX_BLOCKS * 2 * c; }
// per-layer saved statistics: scale, shift, mean, invstd (4*C floats)
int bn_stats(const bf16* z, int64_t P, int c, float* partials, int* nblk, cudaStream_t s);
// finalize kernels work on the channel sub-range [coff, coff+cn) of a c-wide layout (gamma/beta/running_* have cn entries)
int bn_finalize_train(const float* partials, int nblk, int c, int coff, int cn, int64_t count, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps, float* stats4, cudaStream_t s);
int bn_finalize_eval(int c, int coff, int cn, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                     float eps, float* stats4, cudaStream_t s);
// fused statistics + finalize (one launch): fp64 atomics into `gacc` ([2][c] doubles, zero on entry, left zero) and a
// last-block-done ticket in `counter` (zero on entry, left zero).  Up to two parameter segments (the two head branches).
// (BnFinSeg / BnFin live in conv.h: the convolution epilogues fuse the same statistics + finalize step)
struct BnBwdFinSeg { const float* gamma; float* dgamma; float* dbeta; int coff, cn; };
struct BnBwdFin { BnBwdFinSeg seg[2]; int nseg; float* coef3; };
constexpr size_t BN_ACC_BYTES = 2 * 2048 * sizeof(double) + 256;
template <typename T> int bn_stats_finalize(const T* z, int64_t P, int c, const BnFin& fin, double* gacc, unsigned* counter, cudaStream_t s);
template <typename T>
int bn_bwd_reduce_finalize(const T* dout, int dout_ctotal, int dout_coff, const ND<T>* mask_src, int mask_ctotal, int mask_coff,
                           const ND<T>* z, int64_t P, int c, const float* stats4, const BnBwdFin& fin, double* gacc,
                           unsigned* counter, cudaStream_t s, int mask_self = 0);
// out = [relu](z*scale + shift [+ res])
template <typename T>
int bn_apply(const T* z, int64_t P, int c, const float* stats4, const ND<T>* res, int relu, ND<T>* out, int out_ctotal,
             int out_coff, cudaStream_t s);
// backward.  dy = dout * (mask_src > 0) if mask_src else dout;  mask_self = 1: dy = dout * (z*scale + shift > 0), i.e. the
// layer's own ReLU output re-derived from z (mask_src ignored; one tensor read less).
int bn_bwd_reduce(const bf16* dout, int dout_ctotal, int dout_coff, const bf16* mask_src, int mask_ctotal, int mask_coff,
                  const bf16* z, int64_t P, int c, const float* stats4, float* partials, int* nblk, cudaStream_t s);
// dgamma, dbeta: fp32 parameter gradients (overwritten); coef3: [a = gamma*invstd | c1 = dbeta/M | c2 = dgamma/M]
int bn_bwd_finalize(const float* partials, int nblk, int c, int coff, int cn, int64_t count, const float* gamma, const float* stats4,
                    float* dgamma, float* dbeta, float* coef3, cudaStream_t s);
// dz = a*(dy - c1 - xhat*c2);  dsum (optional) receives dy (the ReLU-masked gradient, for the residual path)
template <typename T>
int bn_bwd_apply(const T* dout, int dout_ctotal, int dout_coff, const ND<T>* mask_src, int mask_ctotal, int mask_coff,
                 const ND<T>* z, int64_t P, int c, const float* stats4, const float* coef3, ND<T>* dz, ND<T>* dsum,
                 cudaStream_t s, int mask_self = 0);
// conv bias gradient: dbias[c] = sum_px dz[px, c]  (re-uses the bn partial buffers)
int bias_grad(const bf16* dz, int64_t P, int c, float* partials, float* dbias, cudaStream_t s);

// ---- MaxPool2d(3, 2, 1)  (src/modules/resnet.py:175,235)
// bn_stats4 (optional): x is a raw conv output, pool bf16(relu(x*scale + shift)) computed on the fly (fused BatchNorm apply)
template <typename T> int maxpool_fwd(const T* x, int n, int h, int w, int c, ND<T>* y, uint8_t* argmax, cudaStream_t s, const float* bn_stats4 = nullptr);
template <typename T> int maxpool_bwd(const T* dy, const uint8_t* argmax, int n, int h, int w, int c, ND<T>* dx, cudaStream_t s);

// ---- FPN glue: F.interpolate(mode='nearest') (+ add / concat)  (src/modules/segmentation_body.py:79-87)
// out[n,h,w,:] = y[n,h,w,:] + xs[n, src(h), src(w), :]
template <typename T> int upsample_add_fwd(const T* xs, int hs, int ws, const ND<T>* y, int n, int h, int w, int c, ND<T>* out, cudaStream_t s);
// dst[n,h,w, coff:coff+c] = xs[n, src(h), src(w), :]
template <typename T> int upsample_into(const T* xs, int hs, int ws, int n, int h, int w, int c, ND<T>* dst, int dst_ctotal, int dst_coff, cudaStream_t s);
// d_xs[n,hs,ws,:] (+)= sum over the destination pixels that read it of d_big[n,h,w, coff:coff+c]
template <typename T> int upsample_bwd(const T* d_big, int big_ctotal, int big_coff, int n, int h, int w, int c, ND<T>* d_xs, int hs, int ws,
                 int accumulate, cudaStream_t s);

// ---- conv1 (7x7/2, 3->64) helpers: space-to-depth staging of the NCHW float32 image
// s2d buffer: [n][hs+3][ws+3][16] bf16, hs = ceil(h/2), ws = ceil(w/2); 2 zero rows/cols before, 1 after
int image_to_s2d(const float* img, int n, int h, int w, bf16* s2d, cudaStream_t s);
// dw_s2d [64][64][4] fp32 -> dW (64,3,7,7) fp32
int conv1_wgrad_unpack(const float* dw_s2d, float* dw, cudaStream_t s);

// ---- final bilinear resize, align_corners=True (src/models.py:43-46); NCHW float32 maps
int bilinear_fwd(const float* x, int nc, int hi, int wi, float* y, int ho, int wo, cudaStream_t s);
int bilinear_bwd(const float* dy, int nc, int hi, int wi, float* dx, int ho, int wo, cudaStream_t s);

}  // namespace dbb
