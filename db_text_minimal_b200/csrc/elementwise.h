// elementwise.h -- host interface of the memory-bound glue kernels (elementwise.cu)
#pragma once
#include "conv.h"

namespace dbb {

int nchw_f32_to_nhwc_bf16(const float* x, bf16* y, int n, int c, int64_t hw, cudaStream_t s);
int nhwc_bf16_to_nchw_f32(const bf16* x, float* y, int n, int c, int64_t hw, cudaStream_t s);

}  // namespace dbb
