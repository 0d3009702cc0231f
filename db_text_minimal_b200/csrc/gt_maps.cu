// gt_maps.cu -- ground-truth border ("threshold") map of the DB data loader on the device.
//
// Replaces the distance-field part of draw_thresh_map (src/db_transforms.py:26-59, compute_distance :62-78), which the
// reference runs in numpy per text polygon inside a single-worker DataLoader (src/data_loaders.py:145-149): for every pixel
// of the dilated polygon's bounding box, canvas = fmax(canvas, 1 - min_edges clip(dist(pixel, edge) / D, 0, 1)).
// The polygon dilation itself (Clipper) stays on the host: its bounding box and D are inputs.
//
// Arithmetic is float64 with explicitly rounded operations (no FMA contraction) in the reference's order, including its
// degenerate-point behaviour (nan_to_num of 1 - cos^2, the cos < 0 branch, NaN-propagating min / NaN-ignoring fmax), so
// the result is bit-identical to numpy.  One CTA per polygon; overlapping polygons combine through an integer atomicMax on
// the float bits (the canvas is non-negative), which is order-independent.
#include "common.cuh"
#include <float.h>

namespace dbb {

__device__ __forceinline__ double seg_distance(double x, double y, double ax, double ay, double bx, double by) {
  const double dx1 = __dsub_rn(x, ax), dy1 = __dsub_rn(y, ay), dx2 = __dsub_rn(x, bx), dy2 = __dsub_rn(y, by);
  const double d1 = __dadd_rn(__dmul_rn(dx1, dx1), __dmul_rn(dy1, dy1));
  const double d2 = __dadd_rn(__dmul_rn(dx2, dx2), __dmul_rn(dy2, dy2));
  const double ex = __dsub_rn(ax, bx), ey = __dsub_rn(ay, by);
  const double d = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
  const double p12 = __dmul_rn(d1, d2);
  const double cosin = __ddiv_rn(__dsub_rn(__dsub_rn(d, d1), d2), __dmul_rn(2.0, __dsqrt_rn(p12)));
  double sq_sin = __dsub_rn(1.0, __dmul_rn(cosin, cosin));
  if (isnan(sq_sin)) sq_sin = 0.0;                               // np.nan_to_num
  else if (isinf(sq_sin)) sq_sin = sq_sin > 0 ? DBL_MAX : -DBL_MAX;
  double res = __dsqrt_rn(__ddiv_rn(__dmul_rn(p12, sq_sin), d));
  if (cosin < 0.0) res = __dsqrt_rn(fmin(d1, d2));               // (a NaN cosine compares false, as in numpy)
  return res;
}

constexpr int TM_MAX_PTS = 64;

__global__ void __launch_bounds__(256)
thresh_map_kernel(float* __restrict__ canvas, int64_t H, int64_t W, const double* __restrict__ pts, const int* __restrict__ poly_start,
                  const int* __restrict__ poly_image, const long long* __restrict__ bbox, const double* __restrict__ dist) {
  const int pi = blockIdx.x;
  const int p0 = poly_start[pi], np = poly_start[pi + 1] - p0;
  const long long xmin = bbox[4 * pi], ymin = bbox[4 * pi + 1], xmax = bbox[4 * pi + 2], ymax = bbox[4 * pi + 3];
  const double D = dist[pi];
  __shared__ double px[TM_MAX_PTS], py[TM_MAX_PTS];
  for (int i = threadIdx.x; i < np; i += 256) {                  // polygon in bounding-box coordinates (:31-32)
    px[i] = __dsub_rn(pts[2 * (p0 + i)], (double)xmin);
    py[i] = __dsub_rn(pts[2 * (p0 + i) + 1], (double)ymin);
  }
  __syncthreads();
  const long long x0 = min(max(0LL, xmin), (long long)W - 1), x1 = min(max(0LL, xmax), (long long)W - 1);
  const long long y0 = min(max(0LL, ymin), (long long)H - 1), y1 = min(max(0LL, ymax), (long long)H - 1);
  // the reference's slice pair [y0-ymin : y1-ymax+height) x [x0-xmin : x1-xmax+width) vs [y0 : y1+1) x [x0 : x1+1): both sides
  // have y1-y0+1 rows when the box overlaps the canvas; otherwise the reference raises on the shape mismatch -- skip
  if (xmax < 0 || ymax < 0 || xmin > W - 1 || ymin > H - 1) return;
  const long long bw = x1 - x0 + 1, bh = y1 - y0 + 1;
  float* cimg = canvas + (int64_t)poly_image[pi] * H * W;
  for (long long t = threadIdx.x; t < bw * bh; t += 256) {
    const long long yy = y0 + t / bw, xx = x0 + t % bw;
    const double lx = (double)(xx - xmin), ly = (double)(yy - ymin);
    float m = 0.f;
    bool nan = false;
    for (int i = 0; i < np; ++i) {
      const int j = (i + 1 == np) ? 0 : i + 1;
      double v = __ddiv_rn(seg_distance(lx, ly, px[i], py[i], px[j], py[j]), D);
      if (!isnan(v)) v = fmin(fmax(v, 0.0), 1.0);                // np.clip keeps NaN
      const float f = __double2float_rn(v);
      if (isnan(f)) nan = true;                                   // np.min propagates NaN
      m = (i == 0) ? f : fminf(m, f);
    }
    if (nan) continue;                                            // np.fmax ignores a NaN operand
    const float val = __fsub_rn(1.f, m);
    atomicMax(reinterpret_cast<int*>(cimg + yy * W + xx), __float_as_int(val));      // val in [0, 1], canvas >= 0
  }
}

}  // namespace dbb

using namespace dbb;

// canvas: (n_images, H, W) float32 device tensor, non-negative (zero it to start).  pts: xy pairs (float64) of all polygons
// back to back; poly_start[npoly + 1]; poly_image[npoly]; bbox[npoly][4] = xmin, ymin, xmax, ymax of the DILATED polygon;
// dist[npoly] = the dilation distance D.  All pointers are device memory.
extern "C" int dbb_thresh_map(float* canvas, int64_t n_images, int64_t h, int64_t w, const double* pts, const int* poly_start,
                              const int* poly_image, const long long* bbox, const double* dist, int npoly, int max_pts, void* stream) {
  if (npoly == 0) return DBB_OK;
  if (!canvas || !pts || !poly_start || !poly_image || !bbox || !dist || npoly < 0 || n_images <= 0 || h <= 0 || w <= 0)
    return set_error(DBB_EINVAL, "thresh_map: bad argument");
  if (max_pts > TM_MAX_PTS) return set_error(DBB_EUNSUPPORTED, "thresh_map: more than 64 points in a polygon");
  DBB_LAUNCH("thresh_map", (cudaStream_t)stream, thresh_map_kernel<<<npoly, 256, 0, (cudaStream_t)stream>>>(canvas, h, w, pts, poly_start, poly_image, bbox, dist));
  return DBB_OK;
}
