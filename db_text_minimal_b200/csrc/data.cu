// data.cu -- the loader -> device boundary of a training step.
//
// The reference's DataLoader hands the model float32 tensors: the mean-subtracted image (src/data_loaders.py:152-158:
// img.astype(np.float32); img[..., c] -= mean[c]; transpose to CHW) and four float32 ground-truth maps
// (src/train.py:163-166: prob_map, supervision_mask, thresh_map, text_area_map), 184 MB per 16 x 640 x 640 batch.  Three of
// the maps are {0, 1} and the image is 8-bit before the subtraction, so a loader feeding this library can ship them as
// uint8 (65.6 MB per batch, lossless: the threshold map stays float32) and expand on the device:
//     img_out[n][c][y][x] = float(img_u8[n][c][y][x]) - mean[c]        gts_out = [float(prob), float(mask), thresh, float(area)]
// One launch, 16-byte loads / stores; the outputs are exactly what the reference's loader would have produced.
#include "common.cuh"

namespace dbb {

__device__ __forceinline__ void u8x16_to_f32(const uint4& u, float sub, float4* dst) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    dst[k] = make_float4((float)(w[k] & 0xffu) - sub, (float)((w[k] >> 8) & 0xffu) - sub, (float)((w[k] >> 16) & 0xffu) - sub,
                         (float)(w[k] >> 24) - sub);
}

// work item = 16 consecutive pixels of one plane; planes: 3N image planes, then N planes of each of the four maps
__global__ void __launch_bounds__(256)
unpack_batch_kernel(const uint8_t* __restrict__ img, float m0, float m1, float m2, const uint8_t* __restrict__ prob,
                    const uint8_t* __restrict__ mask, const float* __restrict__ thresh, const uint8_t* __restrict__ area, int n, int64_t hw,
                    float* __restrict__ img_out, float* __restrict__ gts_out) {
  const int64_t per_plane = hw / 16;
  const int64_t planes = (int64_t)n * 7;
  const int64_t total = planes * per_plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t plane = i / per_plane, off = (i - plane * per_plane) * 16;
    if (plane < 3 * (int64_t)n) {
      const int c = (int)(plane % 3);
      const float sub = c == 0 ? m0 : (c == 1 ? m1 : m2);
      const uint4 u = ldg_stream(reinterpret_cast<const uint4*>(img + plane * hw + off));
      float4 f[4];
      u8x16_to_f32(u, sub, f);
      float4* o = reinterpret_cast<float4*>(img_out + plane * hw + off);
#pragma unroll
      for (int k = 0; k < 4; ++k) stg_stream(o + k, f[k]);
    } else {
      const int64_t q = plane - 3 * (int64_t)n;
      const int map = (int)(q / n);                        // 0 prob, 1 mask, 2 thresh, 3 area
      const int64_t im = q - (int64_t)map * n;
      float4* o = reinterpret_cast<float4*>(gts_out + ((int64_t)map * n + im) * hw + off);
      if (map == 2) {
        const float4* t = reinterpret_cast<const float4*>(thresh + im * hw + off);
#pragma unroll
        for (int k = 0; k < 4; ++k) stg_stream(o + k, ldg_stream(t + k));
      } else {
        const uint8_t* src = map == 0 ? prob : (map == 1 ? mask : area);
        const uint4 u = ldg_stream(reinterpret_cast<const uint4*>(src + im * hw + off));
        float4 f[4];
        u8x16_to_f32(u, 0.f, f);
#pragma unroll
        for (int k = 0; k < 4; ++k) stg_stream(o + k, f[k]);
      }
    }
  }
}

}  // namespace dbb

using namespace dbb;

extern "C" int dbb_unpack_batch(const uint8_t* img_u8, float mean0, float mean1, float mean2, const uint8_t* prob_u8, const uint8_t* mask_u8,
                                const float* thresh_f32, const uint8_t* area_u8, int64_t n, int64_t h, int64_t w, float* img_out,
                                float* gts_out, void* stream) {
  if (!img_u8 || !prob_u8 || !mask_u8 || !thresh_f32 || !area_u8 || !img_out || !gts_out || n <= 0 || h <= 0 || w <= 0)
    return set_error(DBB_EINVAL, "unpack_batch: bad argument");
  const int64_t hw = h * w;
  if (hw % 16) return set_error(DBB_EUNSUPPORTED, "unpack_batch: H*W must be a multiple of 16");
  if (!aligned16(img_u8) || !aligned16(prob_u8) || !aligned16(mask_u8) || !aligned16(thresh_f32) || !aligned16(area_u8) || !aligned16(img_out) ||
      !aligned16(gts_out)) return set_error(DBB_EALIGN, "unpack_batch: pointer not 16B aligned");
  const int64_t total = n * 7 * (hw / 16);
  int64_t g = (total + 255) / 256;
  if (g > DBB_NUM_SMS * 16) g = DBB_NUM_SMS * 16;
  DBB_LAUNCH("unpack_batch", (cudaStream_t)stream, unpack_batch_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(
      img_u8, mean0, mean1, mean2, prob_u8, mask_u8, thresh_f32, area_u8, (int)n, hw, img_out, gts_out));
  return DBB_OK;
}
