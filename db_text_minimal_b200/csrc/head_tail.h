// head_tail.h -- host interface of the fused DBHead tail (head_tail.cu)
#pragma once
#include "common.cuh"
#include "conv.h"
#include "elementwise.h"

namespace dbb {

constexpr int HT_MAX_BLOCKS = DBB_NUM_SMS * 4;
inline size_t head_tail_partials_floats() { return (size_t)HT_MAX_BLOCKS * (128 * 8 + 2); }

// zt: (N, H2, W2, 128) bf16 raw ConvT1 outputs [binarize | thresh]; stats4: BN scale/shift/mean/invstd for 128 channels
// w2b / w2t: ConvTranspose2d(64,1,2,2) weights; b2b / b2t: their biases; out: (N, out_c, 2*H2, 2*W2) float32
template <typename T>
int head_tail_fwd(const T* zt, int n, int h2, int w2, const float* stats4, const float* w2b, const float* w2t,
                  const float* b2b, const float* b2t, float k, int out_c, float* out, cudaStream_t s);
template <typename T>
int head_tail_bwd_reduce(const T* zt, int n, int h2, int w2, const float* stats4, const float* w2b, const float* w2t,
                         const float* out, const float* dout, float k, float* partials, int* nblk, cudaStream_t s);
int head_tail_bwd_finalize(const float* partials, int nblk, int64_t count, const float* gamma_b, const float* gamma_t,
                           const float* stats4, float* dgamma_b, float* dbeta_b, float* dgamma_t, float* dbeta_t,
                           float* coef3, float* dw2b, float* dw2t, float* db2b, float* db2t, const float* w2b, const float* w2t,
                           cudaStream_t s);      // nblk < 0: partials in the tensor-core reduce's {M, Z} layout
template <typename T>
int head_tail_bwd_apply(const T* zt, int n, int h2, int w2, const float* stats4, const float* coef3, const float* w2b,
                        const float* w2t, const float* out, const float* dout, float k, ND<T>* d_zt, cudaStream_t s);

}  // namespace dbb
