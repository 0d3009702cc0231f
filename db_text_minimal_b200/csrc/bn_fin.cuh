// bn_fin.cuh -- device-side pieces shared by the standalone BatchNorm statistics kernel (elementwise.cu) and the fused
// statistics epilogues of the convolution kernels (conv_tcgen05.cu).
#pragma once
#include "common.cuh"
#include "conv.h"

namespace dbb {

// Run by the LAST block/CTA of a statistics pass with `nthreads` cooperating threads: gacc -> stats4 (+ running stats,
// torch.nn.BatchNorm2d: unbiased variance, momentum 0.1); leaves gacc zeroed for the next layer.
__device__ __forceinline__ void bn_finalize_channels(const BnFin& fin, int c, double count, double* __restrict__ gacc, int tid, int nthreads) {
  for (int sg = 0; sg < fin.nseg; ++sg) {
    const BnFinSeg& S = fin.seg[sg];
    for (int i = tid; i < S.cn; i += nthreads) {
      const int ch = S.coff + i;
      const double s = __ldcg(&gacc[ch]), q = __ldcg(&gacc[c + ch]);
      gacc[ch] = 0.0; gacc[c + ch] = 0.0;
      const double mean = s / count;
      double var = q / count - mean * mean;
      if (var < 0.0) var = 0.0;
      const double invstd = 1.0 / sqrt(var + (double)fin.eps);
      fin.stats4[ch] = (float)((double)S.gamma[i] * invstd);
      fin.stats4[c + ch] = (float)((double)S.beta[i] - mean * (double)S.gamma[i] * invstd);
      fin.stats4[2 * c + ch] = (float)mean;
      fin.stats4[3 * c + ch] = (float)invstd;
      if (S.rmean) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        S.rmean[i] = (float)((1.0 - fin.momentum) * (double)S.rmean[i] + fin.momentum * mean);
        S.rvar[i] = (float)((1.0 - fin.momentum) * (double)S.rvar[i] + fin.momentum * unb);
      }
    }
  }
}

// Column totals of a 32-row x 16-column register tile held one row per lane: 16 shuffles instead of 16 x 5.
// Returns, in every lane, the total of column  8*bit4 + 4*bit3 + 2*bit2 + bit1  of the lane index (lane pairs agree).
__device__ __forceinline__ float column_total16(const float (&x)[16], int lane) {
  float a[8], b[4], c2[2];
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0, h2 = (lane & 2) != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = (h16 ? x[j + 8] : x[j]) + __shfl_xor_sync(0xffffffffu, h16 ? x[j] : x[j + 8], 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = (h8 ? a[j + 4] : a[j]) + __shfl_xor_sync(0xffffffffu, h8 ? a[j] : a[j + 4], 8);
#pragma unroll
  for (int j = 0; j < 2; ++j) c2[j] = (h4 ? b[j + 2] : b[j]) + __shfl_xor_sync(0xffffffffu, h4 ? b[j] : b[j + 2], 4);
  float d = (h2 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? c2[0] : c2[1], 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}
__device__ __forceinline__ int column_of_lane16(int lane) { return (lane >> 1) & 15; }

}  // namespace dbb
