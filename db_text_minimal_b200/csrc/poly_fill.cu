// poly_fill.cu -- cv2.fillPoly on the device, bit-exact, for the ground-truth maps of a training batch.
//
// Replaces the three cv2.fillPoly calls of the reference's loader -- the shrink map `gt` (src/data_loaders.py:131), the
// supervision mask of ignored / too-small text (:107,125,135: fill with 0 on a canvas of ones) and the text-area map
// (src/db_transforms.py:22: the dilated polygon) -- for ALL polygons of a batch in one launch.  Every polygon of a given
// canvas writes the same constant, so the result does not depend on the order in which polygons are drawn.
//
// cv2.fillPoly(img, [pts], v) with integer vertices, line_type 8, shift 0 (OpenCV 4.x drawing.cpp: CollectPolyEdges +
// FillEdgeCollection), restated and checked bit for bit against OpenCV on 3,000 random polygons (2,470 of them reaching
// outside the canvas; tests/test_gt_maps_gpu.py does the same against the device kernel):
//   * every edge is drawn as an 8-connected Bresenham line between the endpoints CLIPPED to the canvas (cv::clipLine),
//     iterating from the left endpoint: err = dmaj - 2 dmin; step the minor axis when err < 0;
//   * scan conversion: 16.16 fixed point; an edge's x comes from the clipped endpoints (its y range from the unclipped ones;
//     the clipped y only when it is not degenerate), dx = trunc((x1 - x0) / (y1 - y0)); a scanline covers y0 <= y < y1; spans
//     between the sorted crossings pair up even-odd and cover ceil(x_left) .. floor(x_right), clipped to the canvas.
#include "common.cuh"

namespace dbb {

constexpr int PF_THREADS = 128;
constexpr int PF_MAX_V = 512;            // vertices per polygon (Clipper's round joins produce ~8 points per corner)

struct PfEdge { int y0, y1; long long x, dx; };

__device__ __forceinline__ bool pf_clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = W - 1, bottom = H - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

__global__ void __launch_bounds__(PF_THREADS)
poly_fill_kernel(const int32_t* __restrict__ verts, const int32_t* __restrict__ start, const int32_t* __restrict__ plane,
                 const float* __restrict__ value, float* __restrict__ maps, int H, int W) {
  __shared__ int vx[PF_MAX_V], vy[PF_MAX_V];
  __shared__ PfEdge edges[PF_MAX_V];
  __shared__ int n_edges, y_min, y_max;
  const int poly = blockIdx.x;
  const int v0 = start[poly], nv = start[poly + 1] - v0;
  if (nv < 1 || nv > PF_MAX_V) return;
  float* img = maps + (int64_t)plane[poly] * H * W;
  const float val = value[poly];
  for (int i = threadIdx.x; i < nv; i += PF_THREADS) { vx[i] = verts[2 * (v0 + i)]; vy[i] = verts[2 * (v0 + i) + 1]; }
  if (threadIdx.x == 0) { n_edges = 0; y_min = INT_MAX; y_max = INT_MIN; }
  __syncthreads();
  // ---- edges: boundary line + scan-conversion record
  for (int i = threadIdx.x; i < nv; i += PF_THREADS) {
    const int j = i == 0 ? nv - 1 : i - 1;                  // edge (v[j] -> v[i]) as OpenCV walks them
    const long long p0x = vx[j], p0y = vy[j], p1x = vx[i], p1y = vy[i];
    long long t0x = p0x, t0y = p0y, t1x = p1x, t1y = p1y;
    bool inside = true;
    if ((unsigned long long)p0x >= (unsigned long long)W || (unsigned long long)p1x >= (unsigned long long)W ||
        (unsigned long long)p0y >= (unsigned long long)H || (unsigned long long)p1y >= (unsigned long long)H)
      inside = pf_clip_line(W, H, t0x, t0y, t1x, t1y);
    if (inside) {                                           // Line(): Bresenham from the left endpoint
      long long x0 = t0x, y0 = t0y, x1 = t1x, y1 = t1y;
      long long dx = x1 - x0, dy = y1 - y0;
      if (dx < 0) { x0 = t1x; y0 = t1y; x1 = t0x; y1 = t0y; dx = -dx; dy = -dy; }
      int sy = 1;
      if (dy < 0) { dy = -dy; sy = -1; }
      const bool vert = dy > dx;
      const long long dmaj = vert ? dy : dx, dmin = vert ? dx : dy;
      long long err = dmaj - 2 * dmin, x = x0, y = y0;
      for (long long s = 0; s <= dmaj; ++s) {
        if ((unsigned long long)x < (unsigned long long)W && (unsigned long long)y < (unsigned long long)H) img[y * W + x] = val;
        if (err < 0) { if (vert) x += 1; else y += sy; err += 2 * dmaj; }
        err -= 2 * dmin;
        if (vert) y += sy; else x += 1;
      }
    }
    if (p0y != p1y) {
      long long c0y = p0y, c1y = p1y;
      if (t0y != t1y) { c0y = t0y; c1y = t1y; }
      const long long c0x = t0x << 16, c1x = t1x << 16;
      PfEdge e;
      e.dx = (c1x - c0x) / (c1y - c0y);                     // C++ integer division truncates toward zero, as OpenCV's
      if (p0y < p1y) { e.y0 = (int)p0y; e.y1 = (int)p1y; e.x = c0x + ((long long)e.y0 - c0y) * e.dx; }
      else { e.y0 = (int)p1y; e.y1 = (int)p0y; e.x = c1x + ((long long)e.y0 - c1y) * e.dx; }
      const int slot = atomicAdd(&n_edges, 1);
      edges[slot] = e;
      atomicMin(&y_min, e.y0); atomicMax(&y_max, e.y1);
    }
  }
  __syncthreads();
  const int ne = n_edges;
  if (ne < 2) return;
  const int ya = y_min < 0 ? 0 : y_min, yb = y_max < H ? y_max : H;
  // ---- scanlines: one thread per row; crossings sorted by insertion (a handful per row)
  for (int y = ya + threadIdx.x; y < yb; y += PF_THREADS) {
    long long xs[32];
    int k = 0;
    bool overflow = false;
    for (int e = 0; e < ne; ++e) {
      if (edges[e].y0 <= y && y < edges[e].y1) {
        const long long x = edges[e].x + (long long)(y - edges[e].y0) * edges[e].dx;
        if (k == 32) { overflow = true; break; }
        int p = k++;
        while (p > 0 && xs[p - 1] > x) { xs[p] = xs[p - 1]; --p; }
        xs[p] = x;
      }
    }
    if (overflow) continue;                                 // > 32 crossings on one row: not a text polygon (host rejects > PF_MAX_V)
    for (int p = 0; p + 1 < k; p += 2) {
      long long x1 = (xs[p] + 65535) >> 16, x2 = xs[p + 1] >> 16;
      if (x1 < W && x2 >= 0) {
        if (x1 < 0) x1 = 0;
        if (x2 >= W) x2 = W - 1;
        for (long long x = x1; x <= x2; ++x) img[(int64_t)y * W + x] = val;
      }
    }
  }
}

}  // namespace dbb

using namespace dbb;

// verts: (total, 2) int32 device; start: (npolys + 1) int32 device; plane: (npolys) canvas index of each polygon into
// maps (planes, H, W) float32; value: (npolys) constant to write.  Canvases are filled in place.
extern "C" int dbb_fill_polygons(const int32_t* verts, const int32_t* start, const int32_t* plane, const float* value, int npolys,
                                 float* maps, int64_t h, int64_t w, void* stream) {
  if (npolys == 0) return DBB_OK;
  if (!verts || !start || !plane || !value || !maps || npolys < 0 || h <= 0 || w <= 0 || h > 32767 || w > 32767)
    return set_error(DBB_EINVAL, "fill_polygons: bad argument");
  DBB_LAUNCH("poly_fill", (cudaStream_t)stream, poly_fill_kernel<<<(unsigned)npolys, PF_THREADS, 0, (cudaStream_t)stream>>>(
      verts, start, plane, value, maps, (int)h, (int)w));
  return DBB_OK;
}
