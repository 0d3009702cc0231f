// conv_f32.cu -- float32 convolution kernels of the executor's fp32-parity mode (dbb_net_create_ex(..., DBB_PRECISION_FP32)).
//
// The product path computes every convolution on the tcgen05 tensor cores with bf16 operands (conv_tcgen05.cu); one bf16
// rounding per stored activation is ~4e-3 relative, which a randomly initialised 32-BatchNorm network amplifies far above
// north_star's fp32 bar (P, T, B within 1e-4; losses and gradients within 1e-3).  This file is the other half of the
// parity argument: the SAME executor graph (net.cu, templated on the activation type) and the SAME elementwise / head-tail
// kernels run with float activations and these plain CUDA-core implicit GEMMs, so that the wiring of the whole network --
// src/models.py:34-48, src/modules/resnet.py:231-242, src/modules/segmentation_body.py:64-87,
// src/modules/segmentation_head.py:35-45 of the reference and the backward autograd derives from them -- is checked against
// the reference goldens at fp32 tolerance, while the tensor-core kernels are checked per shape on identical inputs.
// It is a verification mode: tiled SIMT FMA (64x64x16 tiles, 4x4 outputs per thread), not tuned.
//
// One kernel, three index maps (all tensors addressed through explicit strides so NHWC activations, channel slices of wider
// tensors, the NCHW image, OIHW / (ci,co,kh,kw) weights and the pixel-shuffled ConvTranspose output are the same code):
//   FPROP  y[n,oy,ox,co]   = sum_{kh,kw,ci} x[n, oy*s+kh-p, ox*s+kw-p, ci] * w[co,ci,kh,kw] (+ bias[co])
//   DGRAD  dx[n,iy,ix,ci] (+)= sum_{kh,kw,co} dy[n,(iy+p-kh)/s,(ix+p-kw)/s,co] * w[co,ci,kh,kw]   (where divisible)
//   WGRAD  dw[co,ci,kh,kw]  = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s+kh-p, ox*s+kw-p, ci]       (split-K, fp32 atomics)
#include "common.cuh"
#include "conv_ops.h"

namespace dbb {

namespace {

struct View {            // element (n, y, x, c) at p[n*sn + y*sh + x*sw + c*sc]; reads outside [0,H) x [0,W) give 0
  const float* p;
  int64_t sn, sh, sw, sc;
  int H, W;
};
struct WView { const float* p; int64_t so, si, skh, skw; };     // w[o, i, kh, kw]
struct F32Conv {
  View x;                // FPROP/WGRAD: the convolution input; DGRAD: dy
  View y;                // FPROP: output; DGRAD: dx; WGRAD: dy (read)
  WView w;               // weights (FPROP/DGRAD: read; WGRAD: written through wout)
  float* out;            // FPROP: y base; DGRAD: dx base; WGRAD: dw base
  double* acc64;         // WGRAD: split-K partial sums are added here in double (scratch), then rounded once into out
  const float* bias;
  int n, cin, cout, ks, stride, pad;
  int Ho, Wo;            // output extent of the forward convolution
  int accumulate;
  int64_t M, N, K;
  int64_t k_per_split;
};

constexpr int BM = 64, BN = 64, BK = 16, TH = 256;

__device__ __forceinline__ float view_at(const View& v, int n, int y, int x, int c) {
  if ((unsigned)y >= (unsigned)v.H || (unsigned)x >= (unsigned)v.W) return 0.f;
  return __ldg(v.p + n * v.sn + y * v.sh + x * v.sw + c * v.sc);
}

template <int MODE>      // 0 FPROP, 1 DGRAD, 2 WGRAD
__global__ void __launch_bounds__(TH) f32_conv_kernel(const F32Conv P) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * P.k_per_split;
  const int64_t kend = (kbeg + P.k_per_split < P.K) ? kbeg + P.k_per_split : P.K;
  // loader roles: thread -> (row r = tid % 64 of the A / B tile, k lanes tid/64 + 4*i)
  const int lr = tid & 63, lk = tid >> 6;
  // hoisted decode of the loader's M-side and N-side coordinates
  const int64_t am = m0 + lr, bn = n0 + lr;
  int a_n = 0, a_y = 0, a_x = 0;           // FPROP: (n, oy, ox); DGRAD: (n, iy, ix)
  int b_ci = 0, b_kh = 0, b_kw = 0;        // WGRAD: column j = (ci, kh, kw)
  if (MODE == 0) { const int64_t hw = (int64_t)P.Ho * P.Wo; a_n = (int)(am / hw); const int r = (int)(am % hw); a_y = r / P.Wo; a_x = r % P.Wo; }
  if (MODE == 1) { const int64_t hw = (int64_t)P.y.H * P.y.W; a_n = (int)(am / hw); const int r = (int)(am % hw); a_y = r / P.y.W; a_x = r % P.y.W; }
  if (MODE == 2) { const int kk = P.ks * P.ks; b_ci = (int)(bn / kk); const int r = (int)(bn % kk); b_kh = r / P.ks; b_kw = r % P.ks; }
  const int tx = tid & 15, ty = tid >> 4;
  // double accumulators: exact products of float operands, so a convolution result carries ONE float rounding (when it
  // is stored) -- the fp32 mode must not be noisier than the reference it is compared with (B200 runs FP64 FMA at half the
  // FP32 rate; this is a verification mode)
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kl = lk + 4 * i;
      const int64_t k = k0 + kl;
      float av = 0.f, bv = 0.f;
      if (k < kend) {
        if (MODE == 0) {
          const int tap = (int)(k / P.cin), ci = (int)(k % P.cin), kh = tap / P.ks, kw = tap % P.ks;
          if (am < P.M) av = view_at(P.x, a_n, a_y * P.stride + kh - P.pad, a_x * P.stride + kw - P.pad, ci);
          if (bn < P.N) bv = __ldg(P.w.p + bn * P.w.so + ci * P.w.si + kh * P.w.skh + kw * P.w.skw);
        } else if (MODE == 1) {
          const int tap = (int)(k / P.cout), co = (int)(k % P.cout), kh = tap / P.ks, kw = tap % P.ks;
          if (am < P.M) {
            const int tyy = a_y + P.pad - kh, txx = a_x + P.pad - kw;
            if (tyy >= 0 && txx >= 0 && tyy % P.stride == 0 && txx % P.stride == 0) av = view_at(P.x, a_n, tyy / P.stride, txx / P.stride, co);
          }
          if (bn < P.N) bv = __ldg(P.w.p + co * P.w.so + bn * P.w.si + kh * P.w.skh + kw * P.w.skw);
        } else {
          const int64_t hw = (int64_t)P.Ho * P.Wo;
          const int nn = (int)(k / hw); const int r = (int)(k % hw); const int oy = r / P.Wo, ox = r % P.Wo;
          if (am < P.M) av = view_at(P.y, nn, oy, ox, (int)am);
          if (bn < P.N) bv = view_at(P.x, nn, oy * P.stride + b_kh - P.pad, ox * P.stride + b_kw - P.pad, b_ci);
        }
      }
      As[kl][lr] = av;
      Bs[kl][lr] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kl = 0; kl < BK; ++kl) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = (double)As[kl][ty * 4 + i]; b[i] = (double)Bs[kl][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- store
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= P.M) continue;
    int on = 0, oy = 0, ox = 0;
    if (MODE == 0) { const int64_t hw = (int64_t)P.Ho * P.Wo; on = (int)(m / hw); const int r = (int)(m % hw); oy = r / P.Wo; ox = r % P.Wo; }
    if (MODE == 1) { const int64_t hw = (int64_t)P.y.H * P.y.W; on = (int)(m / hw); const int r = (int)(m % hw); oy = r / P.y.W; ox = r % P.y.W; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t nn = n0 + tx * 4 + j;
      if (nn >= P.N) continue;
      if (MODE == 2) {
        const int kk = P.ks * P.ks; const int ci = (int)(nn / kk); const int r = (int)(nn % kk);
        atomicAdd(P.acc64 + m * P.w.so + ci * P.w.si + (r / P.ks) * P.w.skh + (r % P.ks) * P.w.skw, acc[i][j]);
      } else {
        float* o = P.out + on * P.y.sn + oy * P.y.sh + ox * P.y.sw + nn * P.y.sc;
        double v = acc[i][j];
        if (MODE == 0 && P.bias) v += (double)P.bias[nn];
        if (P.accumulate) v += (double)*o;
        *o = (float)v;
      }
    }
  }
}

__global__ void f32_round_kernel(const double* __restrict__ a, float* __restrict__ o, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) o[i] = (float)a[i];
}
// weight gradients: zero the double scratch, accumulate the split-K partials into it, round once
int wgrad_finish(F32Conv& P, int64_t numel, float* scratch, size_t scratch_bytes, cudaStream_t s, int (*run)(int, F32Conv&, cudaStream_t)) {
  if (!scratch || scratch_bytes < (size_t)numel * sizeof(double)) return set_error(DBB_EWORKSPACE, "fp32 wgrad: scratch too small");
  P.acc64 = reinterpret_cast<double*>(scratch);
  DBB_CUDA(cudaMemsetAsync(P.acc64, 0, sizeof(double) * (size_t)numel, s));
  int rc = run(2, P, s);
  if (rc) return rc;
  DBB_LAUNCH("f32_round", s, f32_round_kernel<<<(unsigned)((numel + 255) / 256 > 1184 ? 1184 : (numel + 255) / 256), 256, 0, s>>>(P.acc64, P.out, numel));
  return DBB_OK;
}

View nhwc(const float* p, int h, int w, int ctotal, int coff) {
  return View{p + coff, (int64_t)h * w * ctotal, (int64_t)w * ctotal, ctotal, 1, h, w};
}

int launch(int mode, F32Conv& P, cudaStream_t s) {
  dim3 grid((unsigned)((P.M + BM - 1) / BM), (unsigned)((P.N + BN - 1) / BN), 1);
  P.k_per_split = P.K;
  if (mode == 2) {
    // split the pixel reduction so that ~4 waves of CTAs exist; multiples of BK keep tiles aligned
    const int64_t tiles = (int64_t)grid.x * grid.y;
    int64_t splits = (4 * DBB_NUM_SMS + tiles - 1) / tiles;
    const int64_t maxs = (P.K + 8 * BK - 1) / (8 * BK);
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    P.k_per_split = ((P.K + splits - 1) / splits + BK - 1) / BK * BK;
    grid.z = (unsigned)((P.K + P.k_per_split - 1) / P.k_per_split);
  }
  if (mode == 0) DBB_LAUNCH("f32_conv_fprop", s, f32_conv_kernel<0><<<grid, TH, 0, s>>>(P));
  else if (mode == 1) DBB_LAUNCH("f32_conv_dgrad", s, f32_conv_kernel<1><<<grid, TH, 0, s>>>(P));
  else DBB_LAUNCH("f32_conv_wgrad", s, f32_conv_kernel<2><<<grid, TH, 0, s>>>(P));
  return DBB_OK;
}

}  // namespace

int conv_fprop(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* w, const float* bias, float* y,
               int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st, const ConvEpi* epi) {
  if (st || epi) return set_error(DBB_EUNSUPPORTED, "fp32 conv_fprop: fused statistics / epilogue are bf16-path features");
  F32Conv P{};
  P.x = nhwc(x, g.h, g.w, x_ctotal, x_coff);
  P.y = nhwc(y, g.out_h(), g.out_w(), y_ctotal, y_coff);
  P.w = WView{w, (int64_t)g.cin * g.ks * g.ks, (int64_t)g.ks * g.ks, g.ks, 1};
  P.out = y + y_coff; P.bias = bias;
  P.n = g.n; P.cin = g.cin; P.cout = g.cout; P.ks = g.ks; P.stride = g.stride; P.pad = g.pad;
  P.Ho = g.out_h(); P.Wo = g.out_w();
  P.M = (int64_t)g.n * P.Ho * P.Wo; P.N = g.cout; P.K = (int64_t)g.ks * g.ks * g.cin;
  return launch(0, P, s);
}

int conv_dgrad(const ConvGeom& g, const float* dy, const float* w, float* dx, cudaStream_t s, int accumulate) {
  F32Conv P{};
  P.x = nhwc(dy, g.out_h(), g.out_w(), g.cout, 0);
  P.y = nhwc(dx, g.h, g.w, g.cin, 0);
  P.w = WView{w, (int64_t)g.cin * g.ks * g.ks, (int64_t)g.ks * g.ks, g.ks, 1};
  P.out = dx; P.accumulate = accumulate;
  P.n = g.n; P.cin = g.cin; P.cout = g.cout; P.ks = g.ks; P.stride = g.stride; P.pad = g.pad;
  P.Ho = g.out_h(); P.Wo = g.out_w();
  P.M = (int64_t)g.n * g.h * g.w; P.N = g.cin; P.K = (int64_t)g.ks * g.ks * g.cout;
  return launch(1, P, s);
}

int conv_wgrad(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* dy, int dy_ctotal, int dy_coff, float* dw,
               float* scratch, size_t scratch_bytes, cudaStream_t s) {
  F32Conv P{};
  P.x = nhwc(x, g.h, g.w, x_ctotal, x_coff);
  P.y = nhwc(dy, g.out_h(), g.out_w(), dy_ctotal, dy_coff);
  P.w = WView{nullptr, (int64_t)g.cin * g.ks * g.ks, (int64_t)g.ks * g.ks, g.ks, 1};
  P.out = dw;
  P.n = g.n; P.cin = g.cin; P.cout = g.cout; P.ks = g.ks; P.stride = g.stride; P.pad = g.pad;
  P.Ho = g.out_h(); P.Wo = g.out_w();
  P.M = g.cout; P.N = (int64_t)g.cin * g.ks * g.ks; P.K = (int64_t)g.n * P.Ho * P.Wo;
  return wgrad_finish(P, (int64_t)g.cout * g.cin * g.ks * g.ks, scratch, scratch_bytes, s, launch);
}

// ConvTranspose2d(k2, s2), weight (ci, co, 2, 2):  y[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci] W[ci,co,a,b] + bias[co]
// -> four 1x1 FPROP launches, one per output parity class, writing through doubled output strides.
int convt_fprop(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* w, const float* bias, float* y,
                int y_ctotal, int y_coff, cudaStream_t s, const ConvStats* st) {
  if (st) return set_error(DBB_EUNSUPPORTED, "fp32 convt_fprop: fused statistics are a bf16-path feature");
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      F32Conv P{};
      P.x = nhwc(x, g.h, g.w, x_ctotal, x_coff);
      const int64_t W2 = 2 * (int64_t)g.w;
      P.y = View{nullptr, 4 * (int64_t)g.h * g.w * y_ctotal, 2 * W2 * y_ctotal, 2 * (int64_t)y_ctotal, 1, g.h, g.w};
      P.out = y + y_coff + ((int64_t)a * W2 + b) * y_ctotal;
      P.w = WView{w + a * 2 + b, 4, (int64_t)g.cout * 4, 0, 0};      // w[o = co, i = ci]
      P.bias = bias;
      P.n = g.n; P.cin = g.cin; P.cout = g.cout; P.ks = 1; P.stride = 1; P.pad = 0;
      P.Ho = g.h; P.Wo = g.w;
      P.M = (int64_t)g.n * g.h * g.w; P.N = g.cout; P.K = g.cin;
      int rc = launch(0, P, s);
      if (rc) return rc;
    }
  return DBB_OK;
}

// dx[n,i,j,ci] = sum_{a,b,co} dy[n,2i+a,2j+b,co] W[ci,co,a,b]  == a 2x2 / stride 2 convolution of dy with w[o = ci, i = co]
int convt_dgrad(const ConvGeom& g, const float* dy, int dy_ctotal, int dy_coff, const float* w, float* dx, int dx_ctotal,
                int dx_coff, cudaStream_t s) {
  F32Conv P{};
  P.x = nhwc(dy, 2 * g.h, 2 * g.w, dy_ctotal, dy_coff);
  P.y = nhwc(dx, g.h, g.w, dx_ctotal, dx_coff);
  P.w = WView{w, (int64_t)g.cout * 4, 4, 2, 1};
  P.out = dx + dx_coff;
  P.n = g.n; P.cin = g.cout; P.cout = g.cin; P.ks = 2; P.stride = 2; P.pad = 0;
  P.Ho = g.h; P.Wo = g.w;
  P.M = (int64_t)g.n * g.h * g.w; P.N = g.cin; P.K = 4 * (int64_t)g.cout;
  return launch(0, P, s);
}

// dW[ci,co,a,b] = sum_{n,i,j} x[n,i,j,ci] dy[n,2i+a,2j+b,co]: WGRAD with the roles swapped ("dy" := x, "x" := dy, k2 s2)
int convt_wgrad(const ConvGeom& g, const float* x, int x_ctotal, int x_coff, const float* dy, int dy_ctotal, int dy_coff, float* dw,
                float* scratch, size_t scratch_bytes, cudaStream_t s) {
  F32Conv P{};
  P.x = nhwc(dy, 2 * g.h, 2 * g.w, dy_ctotal, dy_coff);
  P.y = nhwc(x, g.h, g.w, x_ctotal, x_coff);
  P.w = WView{nullptr, (int64_t)g.cout * 4, 4, 2, 1};
  P.out = dw;
  P.n = g.n; P.cin = g.cout; P.cout = g.cin; P.ks = 2; P.stride = 2; P.pad = 0;
  P.Ho = g.h; P.Wo = g.w;
  P.M = g.cin; P.N = (int64_t)g.cout * 4; P.K = (int64_t)g.n * g.h * g.w;
  return wgrad_finish(P, (int64_t)g.cin * g.cout * 4, scratch, scratch_bytes, s, launch);
}

// conv1: Conv2d(3, 64, 7, stride 2, pad 3, bias=False) straight from the NCHW float32 image (no space-to-depth staging)
int conv1_fprop_f32(int n, int h, int w, const float* img, const float* wt, float* y, cudaStream_t s) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  F32Conv P{};
  P.x = View{img, 3 * (int64_t)h * w, w, 1, (int64_t)h * w, h, w};
  P.y = nhwc(y, ho, wo, 64, 0);
  P.w = WView{wt, 147, 49, 7, 1};
  P.out = y;
  P.n = n; P.cin = 3; P.cout = 64; P.ks = 7; P.stride = 2; P.pad = 3;
  P.Ho = ho; P.Wo = wo;
  P.M = (int64_t)n * ho * wo; P.N = 64; P.K = 147;
  return launch(0, P, s);
}
int conv1_wgrad_f32(int n, int h, int w, const float* img, const float* dy, float* dw, float* scratch, size_t scratch_bytes, cudaStream_t s) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  F32Conv P{};
  P.x = View{img, 3 * (int64_t)h * w, w, 1, (int64_t)h * w, h, w};
  P.y = nhwc(dy, ho, wo, 64, 0);
  P.w = WView{nullptr, 147, 49, 7, 1};
  P.out = dw;
  P.n = n; P.cin = 3; P.cout = 64; P.ks = 7; P.stride = 2; P.pad = 3;
  P.Ho = ho; P.Wo = wo;
  P.M = 64; P.N = 147; P.K = (int64_t)n * ho * wo;
  return wgrad_finish(P, 64 * 147, scratch, scratch_bytes, s, launch);
}

}  // namespace dbb
