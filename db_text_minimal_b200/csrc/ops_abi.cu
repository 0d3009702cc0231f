// ops_abi.cu -- C-ABI entry points of the memory-bound operators (BatchNorm, max-pool, FPN nearest-upsample glue, fused
// head tail) so that each kernel family can be parity-tested in isolation on identical inputs.
#include "common.cuh"
#include "elementwise.h"
#include "head_tail.h"

using namespace dbb;

extern "C" size_t dbb_ops_workspace(void) {
  size_t f = bn_partials_floats(2048);
  if (head_tail_partials_floats() > f) f = head_tail_partials_floats();
  return f * sizeof(float) + 4096;
}

// BatchNorm2d (+ residual) (+ ReLU), NHWC bf16.  training != 0: batch statistics, running stats updated (may be null).
// stats4 (out): [scale | shift | mean | invstd], 4*C floats, needed by dbb_bn_bwd.
extern "C" int dbb_bn_fwd(const void* z, int64_t pixels, int c, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int training, const void* residual, int relu, void* out, float* stats4,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!z || !gamma || !beta || !out || !stats4 || !workspace) return set_error(DBB_EINVAL, "bn_fwd: null pointer");
  if (workspace_bytes < dbb_ops_workspace()) return set_error(DBB_EWORKSPACE, "bn_fwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  int rc, nblk = 0;
  if (training) {
    if ((rc = bn_stats((const bf16*)z, pixels, c, (float*)workspace, &nblk, s))) return rc;
    if ((rc = bn_finalize_train((const float*)workspace, nblk, c, 0, c, pixels, gamma, beta, running_mean, running_var, 0.1f, 1e-5f, stats4, s))) return rc;
  } else {
    if (!running_mean || !running_var) return set_error(DBB_EINVAL, "bn_fwd: eval mode needs running statistics");
    if ((rc = bn_finalize_eval(c, 0, c, gamma, beta, running_mean, running_var, 1e-5f, stats4, s))) return rc;
  }
  return bn_apply((const bf16*)z, pixels, c, stats4, (const bf16*)residual, relu, (bf16*)out, c, 0, s);
}

// backward of (BatchNorm -> [+res] -> [ReLU]): dy = dout * (act > 0) if act else dout
extern "C" int dbb_bn_bwd(const void* dout, const void* act, const void* z, int64_t pixels, int c, const float* gamma,
                          const float* stats4, void* dz, void* dres, float* dgamma, float* dbeta, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (!dout || !z || !gamma || !stats4 || !dz || !workspace) return set_error(DBB_EINVAL, "bn_bwd: null pointer");
  if (workspace_bytes < dbb_ops_workspace()) return set_error(DBB_EWORKSPACE, "bn_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  float* partials = (float*)workspace;
  float* coef3 = partials + bn_partials_floats(c);
  int rc, nblk = 0;
  if ((rc = bn_bwd_reduce((const bf16*)dout, c, 0, (const bf16*)act, c, 0, (const bf16*)z, pixels, c, stats4, partials, &nblk, s))) return rc;
  if ((rc = bn_bwd_finalize(partials, nblk, c, 0, c, pixels, gamma, stats4, dgamma, dbeta, coef3, s))) return rc;
  return bn_bwd_apply((const bf16*)dout, c, 0, (const bf16*)act, c, 0, (const bf16*)z, pixels, c, stats4, coef3, (bf16*)dz, (bf16*)dres, s);
}

extern "C" int dbb_maxpool_fwd(const void* x, int64_t n, int64_t h, int64_t w, int c, void* y, uint8_t* argmax, void* stream) {
  if (!x || !y) return set_error(DBB_EINVAL, "maxpool_fwd: null pointer");
  return maxpool_fwd((const bf16*)x, (int)n, (int)h, (int)w, c, (bf16*)y, argmax, (cudaStream_t)stream);
}
extern "C" int dbb_maxpool_bwd(const void* dy, const uint8_t* argmax, int64_t n, int64_t h, int64_t w, int c, void* dx, void* stream) {
  if (!dy || !argmax || !dx) return set_error(DBB_EINVAL, "maxpool_bwd: null pointer");
  return maxpool_bwd((const bf16*)dy, argmax, (int)n, (int)h, (int)w, c, (bf16*)dx, (cudaStream_t)stream);
}

// FPN._upsample_add (segmentation_body.py:79-80): out = nearest(xs -> (h, w)) + y
extern "C" int dbb_upsample_add_fwd(const void* xs, int64_t hs, int64_t ws, const void* y, int64_t n, int64_t h, int64_t w, int c,
                                    void* out, void* stream) {
  if (!xs || !y || !out) return set_error(DBB_EINVAL, "upsample_add_fwd: null pointer");
  return upsample_add_fwd((const bf16*)xs, (int)hs, (int)ws, (const bf16*)y, (int)n, (int)h, (int)w, c, (bf16*)out, (cudaStream_t)stream);
}
// FPN._upsample_cat (segmentation_body.py:82-87): dst[..., coff:coff+c] = nearest(xs -> (h, w))
extern "C" int dbb_upsample_into(const void* xs, int64_t hs, int64_t ws, int64_t n, int64_t h, int64_t w, int c, void* dst,
                                 int dst_ctotal, int dst_coff, void* stream) {
  if (!xs || !dst) return set_error(DBB_EINVAL, "upsample_into: null pointer");
  return upsample_into((const bf16*)xs, (int)hs, (int)ws, (int)n, (int)h, (int)w, c, (bf16*)dst, dst_ctotal, dst_coff, (cudaStream_t)stream);
}
extern "C" int dbb_upsample_bwd(const void* d_big, int big_ctotal, int big_coff, int64_t n, int64_t h, int64_t w, int c, void* d_xs,
                                int64_t hs, int64_t ws, int accumulate, void* stream) {
  if (!d_big || !d_xs) return set_error(DBB_EINVAL, "upsample_bwd: null pointer");
  return upsample_bwd((const bf16*)d_big, big_ctotal, big_coff, (int)n, (int)h, (int)w, c, (bf16*)d_xs, (int)hs, (int)ws, accumulate, (cudaStream_t)stream);
}

// fused DBHead tail.  zt (N, H2, W2, 128) bf16; gamma/beta/running_* : 128 floats ([binarize.4 | thresh.4]);
// w2b/w2t (64,1,2,2); out (N, out_c, 2*H2, 2*W2) float32; stats4: 512 floats out.
extern "C" int dbb_head_tail_fwd(const void* zt, int64_t n, int64_t h2, int64_t w2, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, int training, const float* w2b, const float* w2t,
                                 const float* b2b, const float* b2t, float k, int out_c, float* out, float* stats4,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!zt || !gamma || !beta || !w2b || !w2t || !b2b || !b2t || !out || !stats4 || !workspace) return set_error(DBB_EINVAL, "head_tail_fwd: null pointer");
  if (out_c != 2 && out_c != 3) return set_error(DBB_EINVAL, "head_tail_fwd: out_c must be 2 or 3");
  if (workspace_bytes < dbb_ops_workspace()) return set_error(DBB_EWORKSPACE, "head_tail_fwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t px = n * h2 * w2;
  int rc, nblk = 0;
  if (training) {
    if ((rc = bn_stats((const bf16*)zt, px, 128, (float*)workspace, &nblk, s))) return rc;
    if ((rc = bn_finalize_train((const float*)workspace, nblk, 128, 0, 128, px, gamma, beta, running_mean, running_var, 0.1f, 1e-5f, stats4, s))) return rc;
  } else {
    if (!running_mean || !running_var) return set_error(DBB_EINVAL, "head_tail_fwd: eval mode needs running statistics");
    if ((rc = bn_finalize_eval(128, 0, 128, gamma, beta, running_mean, running_var, 1e-5f, stats4, s))) return rc;
  }
  return head_tail_fwd((const bf16*)zt, (int)n, (int)h2, (int)w2, stats4, w2b, w2t, b2b, b2t, k, out_c, out, s);
}
// grads out: d_zt (N,H2,W2,128) bf16, dgamma/dbeta (128), dw2b/dw2t (64*4), db2b/db2t (1)
extern "C" int dbb_head_tail_bwd(const void* zt, int64_t n, int64_t h2, int64_t w2, const float* gamma, const float* stats4,
                                 const float* w2b, const float* w2t, const float* out, const float* dout, float k, void* d_zt,
                                 float* dgamma, float* dbeta, float* dw2b, float* dw2t, float* db2b, float* db2t,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!zt || !gamma || !stats4 || !w2b || !w2t || !out || !dout || !d_zt || !dgamma || !dbeta || !dw2b || !dw2t || !db2b || !db2t || !workspace)
    return set_error(DBB_EINVAL, "head_tail_bwd: null pointer");
  if (workspace_bytes < dbb_ops_workspace()) return set_error(DBB_EWORKSPACE, "head_tail_bwd: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  float* partials = (float*)workspace;
  float* coef3 = partials + head_tail_partials_floats();
  int rc, nblk = 0;
  if ((rc = head_tail_bwd_reduce((const bf16*)zt, (int)n, (int)h2, (int)w2, stats4, w2b, w2t, out, dout, k, partials, &nblk, s))) return rc;
  if ((rc = head_tail_bwd_finalize(partials, nblk, n * h2 * w2, gamma, gamma + 64, stats4, dgamma, dbeta, dgamma + 64, dbeta + 64, coef3,
                                   dw2b, dw2t, db2b, db2t, w2b, w2t, s))) return rc;
  return head_tail_bwd_apply((const bf16*)zt, (int)n, (int)h2, (int)w2, stats4, coef3, w2b, w2t, out, dout, k, (bf16*)d_zt, s);
}
