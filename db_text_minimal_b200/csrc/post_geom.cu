// post_geom.cu -- host back half of SegDetectorRepresenter.boxes_from_bitmap (box mode), in C++ (no CUDA in this file).
//
// Replaces, for the candidates the device front kept (score filter) and on the border points it emitted
// (ccl.cu: ccl_points_kernel), src/postprocess.py:119-147 of the reference:
//   :121,158-184  get_mini_boxes  = cv2.minAreaRect(contour) -> cv2.boxPoints -> corner ordering, sside = min(w, h)
//   :122          sside < min_size                                   -> dropped
//   :132,150-156  unclip          = shapely Polygon(box).area * ratio / .length, pyclipper round-join offset by that distance
//   :134-136      get_mini_boxes of the expanded polygon, sside < min_size + 2 -> dropped
//   :142-147      rescale to the destination size in float32, round half to even, clip, int16
// cv2.minAreaRect is restated from OpenCV's rotating calipers (imgproc/rotcalipers.cpp: same float32 arithmetic, same
// tie-breaking "area <= minarea"), cv2.boxPoints from RotatedRect::points.  The offset is ClipperOffset (Clipper 6.4.2,
// JT_ROUND, ET_CLOSEDPOLYGON, arc tolerance 0.25) for CONVEX input -- a min-area box always is; its arithmetic is pinned
// by nothing (pyclipper is absent from the reference tree and from this image): "parity unpinned", see DESIGN.md.
#include "common.cuh"
#include "post_geom.h"
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <climits>
#include <atomic>
#include <thread>
#include <vector>

namespace dbb {

// ---------------------------------------------------------------------------------------------- convex hull
static inline int64_t cross(const IPt& o, const IPt& a, const IPt& b) {
  return (int64_t)(a.x - o.x) * (b.y - o.y) - (int64_t)(a.y - o.y) * (b.x - o.x);
}
// Andrew's monotone chain on points sorted by (x, y) without duplicates: strictly convex vertices, counter-clockwise in a
// y-up frame, starting at pts[0]
static void hull_chain(const std::vector<IPt>& pts, std::vector<IPt>& hull) {
  const int n = (int)pts.size();
  hull.clear();
  if (n < 3) { hull = pts; return; }
  hull.resize(2 * n);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    while (k >= 2 && cross(hull[k - 2], hull[k - 1], pts[i]) <= 0) --k;
    hull[k++] = pts[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {
    while (k >= t && cross(hull[k - 2], hull[k - 1], pts[i]) <= 0) --k;
    hull[k++] = pts[i];
  }
  hull.resize(k - 1);
}
// cv2.minAreaRect runs its calipers on convexHull(points, clockwise=True) (measured against cv2 4.13: the other orientation
// never reproduces its angle / size assignment): the hull starts at the lexicographically smallest (x, then y) point and
// runs clockwise in a y-up frame; two points come largest first.
static void opencv_order(std::vector<IPt>& hull) {
  const int n = (int)hull.size();
  if (n == 2) { if (hull[0].x < hull[1].x || (hull[0].x == hull[1].x && hull[0].y < hull[1].y)) std::swap(hull[0], hull[1]); return; }
  if (n < 3) return;
  double a2 = 0;
  int lo = 0;
  for (int i = 0; i < n; ++i) {
    const IPt &p = hull[i], &q = hull[(i + 1) % n];
    a2 += (double)p.x * q.y - (double)q.x * p.y;
    if (p.x < hull[lo].x || (p.x == hull[lo].x && p.y < hull[lo].y)) lo = i;
  }
  std::rotate(hull.begin(), hull.begin() + lo, hull.end());
  if (a2 > 0) std::reverse(hull.begin() + 1, hull.end());
}
void convex_hull(std::vector<IPt>& pts, std::vector<IPt>& hull) {
  std::sort(pts.begin(), pts.end(), [](const IPt& a, const IPt& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  pts.erase(std::unique(pts.begin(), pts.end(), [](const IPt& a, const IPt& b) { return a.x == b.x && a.y == b.y; }), pts.end());
  hull_chain(pts, hull);
  opencv_order(hull);
}
// the same hull from per-row extremes (device border points bucketed by row): no sort -- rows are visited in order and the
// chain runs on (y, x)-swapped coordinates
void convex_hull_rows(const std::vector<int>& row_min, const std::vector<int>& row_max, int y0, std::vector<IPt>& scratch, std::vector<IPt>& hull) {
  scratch.clear();
  for (size_t r = 0; r < row_min.size(); ++r) {
    if (row_min[r] > row_max[r]) continue;
    scratch.push_back(IPt{y0 + (int)r, row_min[r]});                       // swapped: (x', y') = (y, x)
    if (row_max[r] != row_min[r]) scratch.push_back(IPt{y0 + (int)r, row_max[r]});
  }
  hull_chain(scratch, hull);
  for (IPt& p : hull) std::swap(p.x, p.y);
  opencv_order(hull);
}

// ---------------------------------------------------------------------------------------------- cv2.minAreaRect
// rotating calipers, float32, as OpenCV's rotatingCalipers(..., CALIPERS_MINAREARECT, out): out = {corner, vec1, vec2}
static void rotating_calipers(const FPt* points, int n, float* out) {
  float minarea = FLT_MAX;
  std::vector<float> inv_len(n);
  std::vector<FPt> vect(n);
  int left = 0, bottom = 0, right = 0, top = 0;
  int seq[4] = {-1, -1, -1, -1};
  float orientation = 0, base_a, base_b = 0;
  FPt pt0 = points[0];
  float left_x = pt0.x, right_x = pt0.x, top_y = pt0.y, bottom_y = pt0.y;
  for (int i = 0; i < n; ++i) {
    if (pt0.x < left_x) left_x = pt0.x, left = i;
    if (pt0.x > right_x) right_x = pt0.x, right = i;
    if (pt0.y > top_y) top_y = pt0.y, top = i;
    if (pt0.y < bottom_y) bottom_y = pt0.y, bottom = i;
    const FPt pt = points[(i + 1 < n) ? i + 1 : 0];
    const double dx = pt.x - pt0.x, dy = pt.y - pt0.y;
    vect[i].x = (float)dx; vect[i].y = (float)dy;
    inv_len[i] = (float)(1. / std::sqrt(dx * dx + dy * dy));
    pt0 = pt;
  }
  {
    double ax = vect[n - 1].x, ay = vect[n - 1].y;
    for (int i = 0; i < n; ++i) {
      const double bx = vect[i].x, by = vect[i].y;
      const double convexity = ax * by - ay * bx;
      if (convexity != 0) { orientation = (convexity > 0) ? 1.f : -1.f; break; }
      ax = bx; ay = by;
    }
  }
  base_a = orientation;
  seq[0] = bottom; seq[1] = right; seq[2] = top; seq[3] = left;
  int best_left = 0, best_bottom = 0;
  float best_a = 1.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
  for (int k = 0; k < n; ++k) {
    const float dp[4] = {
        +base_a * vect[seq[0]].x + base_b * vect[seq[0]].y,
        -base_b * vect[seq[1]].x + base_a * vect[seq[1]].y,
        -base_a * vect[seq[2]].x - base_b * vect[seq[2]].y,
        +base_b * vect[seq[3]].x - base_a * vect[seq[3]].y,
    };
    float maxcos = dp[0] * inv_len[seq[0]];
    int main_element = 0;
    for (int i = 1; i < 4; ++i) {
      const float cosalpha = dp[i] * inv_len[seq[i]];
      if (cosalpha > maxcos) { main_element = i; maxcos = cosalpha; }
    }
    {
      const int pindex = seq[main_element];
      const float lead_x = vect[pindex].x * inv_len[pindex], lead_y = vect[pindex].y * inv_len[pindex];
      switch (main_element) {
        case 0: base_a = lead_x; base_b = lead_y; break;
        case 1: base_a = lead_y; base_b = -lead_x; break;
        case 2: base_a = -lead_x; base_b = -lead_y; break;
        default: base_a = -lead_y; base_b = lead_x; break;
      }
    }
    seq[main_element] += 1;
    if (seq[main_element] == n) seq[main_element] = 0;
    float dx = points[seq[1]].x - points[seq[3]].x, dy = points[seq[1]].y - points[seq[3]].y;
    const float width = dx * base_a + dy * base_b;
    dx = points[seq[2]].x - points[seq[0]].x; dy = points[seq[2]].y - points[seq[0]].y;
    const float height = -dx * base_b + dy * base_a;
    const float area = width * height;
    if (area <= minarea) {
      minarea = area;
      best_left = seq[3]; best_a = base_a; best_w = width; best_b = base_b; best_h = height; best_bottom = seq[0];
    }
  }
  const float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
  const float C1 = A1 * points[best_left].x + points[best_left].y * B1;
  const float C2 = A2 * points[best_bottom].x + points[best_bottom].y * B2;
  const float idet = 1.f / (A1 * B2 - A2 * B1);
  out[0] = (C1 * B2 - C2 * B1) * idet;
  out[1] = (A1 * C2 - A2 * C1) * idet;
  out[2] = A1 * best_w; out[3] = B1 * best_w;
  out[4] = A2 * best_h; out[5] = B2 * best_h;
}

RotRect min_area_rect_hull(const std::vector<IPt>& hull) {
  const int n = (int)hull.size();
  RotRect box{0, 0, 0, 0, 0};
  if (n > 2) {
    std::vector<FPt> hp(n);
    for (int i = 0; i < n; ++i) { hp[i].x = (float)hull[i].x; hp[i].y = (float)hull[i].y; }
    float out[6];
    rotating_calipers(hp.data(), n, out);
    box.cx = out[0] + (out[2] + out[4]) * 0.5f;
    box.cy = out[1] + (out[3] + out[5]) * 0.5f;
    box.w = (float)std::sqrt((double)out[2] * out[2] + (double)out[3] * out[3]);
    box.h = (float)std::sqrt((double)out[4] * out[4] + (double)out[5] * out[5]);
    box.angle = (float)std::atan2((double)out[3], (double)out[2]);
  } else if (n == 2) {
    box.cx = ((float)hull[0].x + (float)hull[1].x) * 0.5f;
    box.cy = ((float)hull[0].y + (float)hull[1].y) * 0.5f;
    const double dx = (double)hull[1].x - hull[0].x, dy = (double)hull[1].y - hull[0].y;
    box.w = (float)std::sqrt(dx * dx + dy * dy);
    box.h = 0;
    box.angle = (float)std::atan2(dy, dx);
  } else if (n == 1) {
    box.cx = (float)hull[0].x; box.cy = (float)hull[0].y;
  }
  box.angle = (float)(box.angle * 180 / 3.1415926535897932384626433832795);
  return box;
}
RotRect min_area_rect(std::vector<IPt>& pts) {
  std::vector<IPt> hull;
  convex_hull(pts, hull);
  return min_area_rect_hull(hull);
}

// cv2.boxPoints (RotatedRect::points)
void box_points(const RotRect& r, FPt pt[4]) {
  const double ang = r.angle * 3.1415926535897932384626433832795 / 180.;
  const float b = (float)std::cos(ang) * 0.5f, a = (float)std::sin(ang) * 0.5f;
  pt[0].x = r.cx - a * r.h - b * r.w;
  pt[0].y = r.cy + b * r.h - a * r.w;
  pt[1].x = r.cx + a * r.h - b * r.w;
  pt[1].y = r.cy - b * r.h - a * r.w;
  pt[2].x = 2 * r.cx - pt[0].x;
  pt[2].y = 2 * r.cy - pt[0].y;
  pt[3].x = 2 * r.cx - pt[1].x;
  pt[3].y = 2 * r.cy - pt[1].y;
}

// src/postprocess.py:158-184: the four corners sorted by x (stable), then paired by y; returns sside = min(w, h)
static float mini_box_rect(const RotRect& r, FPt box[4]);
float mini_box(std::vector<IPt>& contour, FPt box[4]) { return mini_box_rect(min_area_rect(contour), box); }
float mini_box_hull(const std::vector<IPt>& hull, FPt box[4]) { return mini_box_rect(min_area_rect_hull(hull), box); }
static float mini_box_rect(const RotRect& r, FPt box[4]) {
  FPt p[4];
  box_points(r, p);
  std::stable_sort(p, p + 4, [](const FPt& a, const FPt& b) { return a.x < b.x; });
  int i1, i2, i3, i4;
  if (p[1].y > p[0].y) { i1 = 0; i4 = 1; } else { i1 = 1; i4 = 0; }
  if (p[3].y > p[2].y) { i2 = 2; i3 = 3; } else { i2 = 3; i3 = 2; }
  box[0] = p[i1]; box[1] = p[i2]; box[2] = p[i3]; box[3] = p[i4];
  return r.w < r.h ? r.w : r.h;
}

// ---------------------------------------------------------------------------------------------- unclip (convex)
static inline int64_t clipper_round(double v) { return v < 0 ? (int64_t)(v - 0.5) : (int64_t)(v + 0.5); }

// ClipperOffset::DoOffset / OffsetPoint / DoRound of Clipper 6.4.2 for ONE closed CONVEX path, JT_ROUND, positive delta.
// For convex input the raw offset path is simple, so the union Clipper runs afterwards only drops collinear / duplicate
// points -- which cannot change the min-area rectangle taken next.
void offset_convex_round(const IPt* in, int n_in, double delta, std::vector<IPt>& out, double arc_tolerance) {
  std::vector<IPt> pts(in, in + n_in);
  double area = 0;
  for (int i = 0; i < n_in; ++i) {
    const IPt& a = pts[i]; const IPt& b = pts[(i + 1) % n_in];
    area += ((double)a.x + b.x) * ((double)a.y - b.y);
  }
  area = -area * 0.5;
  if (area < 0) std::reverse(pts.begin(), pts.end());
  std::vector<IPt> clean;
  for (const IPt& p : pts) if (clean.empty() || p.x != clean.back().x || p.y != clean.back().y) clean.push_back(p);
  if (clean.size() > 1 && clean.front().x == clean.back().x && clean.front().y == clean.back().y) clean.pop_back();
  pts.swap(clean);
  const int n = (int)pts.size();
  out.clear();
  if (n < 3 || delta <= 0) { out = pts; return; }
  double y = arc_tolerance > 0 ? arc_tolerance : 0.25;
  if (y > std::fabs(delta) * 0.25) y = std::fabs(delta) * 0.25;
  double steps = 3.14159265358979323846 / std::acos(1 - y / std::fabs(delta));
  if (steps > std::fabs(delta) * 3.14159265358979323846) steps = std::fabs(delta) * 3.14159265358979323846;
  const double m_sin = std::sin(2 * 3.14159265358979323846 / steps), m_cos = std::cos(2 * 3.14159265358979323846 / steps);
  const double steps_per_rad = steps / (2 * 3.14159265358979323846);
  std::vector<double> nx(n), ny(n);
  for (int j = 0; j < n; ++j) {
    const IPt& a = pts[j]; const IPt& b = pts[(j + 1) % n];
    const double dx = (double)(b.x - a.x), dy = (double)(b.y - a.y);
    const double f = 1.0 / std::sqrt(dx * dx + dy * dy);
    nx[j] = dy * f; ny[j] = -dx * f;
  }
  auto push = [&](double x, double yv) { out.push_back(IPt{(int)clipper_round(x), (int)clipper_round(yv)}); };
  int k = n - 1;
  for (int j = 0; j < n; ++j) {
    double sin_a = nx[k] * ny[j] - nx[j] * ny[k];
    const double cos_a = nx[k] * nx[j] + ny[j] * ny[k];
    if (std::fabs(sin_a * delta) < 1.0 && cos_a > 0) {
      push(pts[j].x + nx[k] * delta, pts[j].y + ny[k] * delta);
      k = j;
      continue;
    }
    if (sin_a > 1.0) sin_a = 1.0; else if (sin_a < -1.0) sin_a = -1.0;
    if (sin_a * delta < 0) {
      push(pts[j].x + nx[k] * delta, pts[j].y + ny[k] * delta);
      out.push_back(pts[j]);
      push(pts[j].x + nx[j] * delta, pts[j].y + ny[j] * delta);
    } else {
      const double a = std::atan2(sin_a, cos_a);
      int st = (int)clipper_round(steps_per_rad * std::fabs(a));
      if (st < 1) st = 1;
      double X = nx[k], Y = ny[k];
      for (int q = 0; q < st; ++q) {
        push(pts[j].x + X * delta, pts[j].y + Y * delta);
        const double X2 = X;
        X = X * m_cos - m_sin * Y;
        Y = X2 * m_sin + Y * m_cos;
      }
      push(pts[j].x + nx[j] * delta, pts[j].y + ny[j] * delta);
    }
    k = j;
  }
}

// one candidate: contour points -> final box; returns false when a size filter drops it
bool box_from_hull(const std::vector<IPt>& hull, float unclip_ratio, int min_size, int src_w, int src_h, int dest_w, int dest_h,
                   int16_t out_box[8], float* sside1, float mini1[8]) {
  FPt b1[4];
  const float s1 = mini_box_hull(hull, b1);
  if (sside1) *sside1 = s1;
  if (mini1) for (int i = 0; i < 4; ++i) { mini1[2 * i] = b1[i].x; mini1[2 * i + 1] = b1[i].y; }
  if (s1 < (float)min_size) return false;
  // shapely: area (shoelace, absolute) and length of the float32 corner ring, in double
  double area = 0, length = 0;
  for (int i = 0; i < 4; ++i) {
    const FPt& a = b1[i]; const FPt& b = b1[(i + 1) & 3];
    area += (double)a.x * (double)b.y - (double)a.y * (double)b.x;
    const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
    length += std::sqrt(dx * dx + dy * dy);
  }
  area = std::fabs(area) * 0.5;
  const double distance = area * (double)unclip_ratio / length;
  IPt ip[4];
  for (int i = 0; i < 4; ++i) { ip[i].x = (int)b1[i].x; ip[i].y = (int)b1[i].y; }      // pyclipper: float -> cInt truncates
  std::vector<IPt> expanded;
  offset_convex_round(ip, 4, distance, expanded, 0.25);
  FPt b2[4];
  const float s2 = mini_box(expanded, b2);
  if (s2 < (float)(min_size + 2)) return false;
  for (int i = 0; i < 4; ++i) {
    // numpy float32 arithmetic: x / width * dest_width, np.round (half to even), np.clip, astype(int16)
    float fx = b2[i].x / (float)src_w * (float)dest_w, fy = b2[i].y / (float)src_h * (float)dest_h;
    fx = std::nearbyintf(fx); fy = std::nearbyintf(fy);
    fx = fx < 0.f ? 0.f : (fx > (float)dest_w ? (float)dest_w : fx);
    fy = fy < 0.f ? 0.f : (fy > (float)dest_h ? (float)dest_h : fy);
    out_box[2 * i] = (int16_t)fx; out_box[2 * i + 1] = (int16_t)fy;
  }
  return true;
}

}  // namespace dbb

using namespace dbb;

// Host function (no device work).  cands: (N, max_cands) records as copied back from dbb_binarize_ccl_score, n_cands (N);
// points (N, cap, 2) {slot, (y << 16) | x} / n_points (N) as copied back from dbb_ccl_border_points (cap_stride = the stride actually copied);
// dest_wh (N, 2) = (dest_width, dest_height) per image.  Outputs: boxes (N, max_cands, 4, 2) int16 and scores (N, max_cands)
// float32, zero rows for dropped candidates (src/postprocess.py:119-148); optional debug outputs sside (N, max_cands) and
// mini (N, max_cands, 4, 2) = the first get_mini_boxes result of every KEPT candidate.
extern "C" int dbb_boxes_from_border_points(const DbbCandidate* cands, const int32_t* n_cands, const int32_t* points,
                                            const int32_t* n_points, int64_t n, int max_cands, int cap_stride, int64_t h, int64_t w,
                                            const int32_t* dest_wh, float unclip_ratio, int min_size, int16_t* boxes,
                                            float* scores, float* sside_out, float* mini_out, int threads) {
  if (!cands || !n_cands || !points || !n_points || !dest_wh || !boxes || !scores || n <= 0 || max_cands <= 0)
    return set_error(DBB_EINVAL, "boxes_from_border_points: bad argument");
  for (int64_t i = 0; i < n; ++i)
    if (n_points[i] > cap_stride) return set_error(DBB_EWORKSPACE, "boxes_from_border_points: the border-point buffer overflowed (n_points > cap)");
  auto work = [&](int64_t img) {
    const int k = n_cands[img] < max_cands ? n_cands[img] : max_cands;
    const DbbCandidate* C = cands + img * max_cands;
    int16_t* B = boxes + img * max_cands * 8;
    float* Sc = scores + img * max_cands;
    std::fill(B, B + (size_t)max_cands * 8, (int16_t)0);
    std::fill(Sc, Sc + max_cands, 0.f);
    // bucket the points by candidate slot (counting sort)
    const int np = n_points[img];
    const int32_t* P = points + img * (int64_t)cap_stride * 2;
    std::vector<int> start(k + 1, 0);
    for (int q = 0; q < np; ++q) { const int s = P[2 * q]; if (s >= 0 && s < k) ++start[s + 1]; }
    for (int s = 0; s < k; ++s) start[s + 1] += start[s];
    std::vector<IPt> sorted(start[k]);
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int q = 0; q < np; ++q) {
      const int s = P[2 * q];
      if (s >= 0 && s < k) sorted[fill[s]++] = IPt{P[2 * q + 1] & 0xffff, (int)((uint32_t)P[2 * q + 1] >> 16)};
    }
    std::vector<IPt> scratch, hull;
    std::vector<int> row_min, row_max;
    for (int s = 0; s < k; ++s) {
      if (!C[s].keep || start[s + 1] == start[s]) continue;
      // per-row extremes inside the candidate's bounding box -> hull without sorting
      const int y0 = C[s].y0, rows = C[s].y1 - C[s].y0 + 1;
      row_min.assign(rows, INT32_MAX); row_max.assign(rows, INT32_MIN);
      for (int q = start[s]; q < start[s + 1]; ++q) {
        const int r = sorted[q].y - y0;
        if (r < 0 || r >= rows) continue;
        if (sorted[q].x < row_min[r]) row_min[r] = sorted[q].x;
        if (sorted[q].x > row_max[r]) row_max[r] = sorted[q].x;
      }
      convex_hull_rows(row_min, row_max, y0, scratch, hull);
      float ss = 0, mini[8];
      const bool ok = box_from_hull(hull, unclip_ratio, min_size, (int)w, (int)h, dest_wh[2 * img], dest_wh[2 * img + 1],
                                    B + (size_t)s * 8, &ss, mini);
      if (sside_out) sside_out[img * max_cands + s] = ss;
      if (mini_out) std::copy(mini, mini + 8, mini_out + ((size_t)img * max_cands + s) * 8);
      if (!ok) { std::fill(B + (size_t)s * 8, B + (size_t)s * 8 + 8, (int16_t)0); continue; }
      Sc[s] = (float)(C[s].sum / (double)C[s].count);
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt > n) nt = (int)n;
  if (nt <= 1) { for (int64_t i = 0; i < n; ++i) work(i); return DBB_OK; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back([&, t]() { for (int64_t i = t; i < n; i += nt) work(i); });
  for (auto& th : pool) th.join();
  return DBB_OK;
}

// get_mini_boxes on an explicit integer contour (tests, polygon mode): box (4, 2) float32 out; returns sside through *sside
extern "C" int dbb_mini_box(const int32_t* contour_xy, int npts, float* box8, float* sside) {
  if (!contour_xy || npts <= 0 || !box8 || !sside) return set_error(DBB_EINVAL, "mini_box: bad argument");
  std::vector<IPt> c(npts);
  for (int i = 0; i < npts; ++i) c[i] = IPt{contour_xy[2 * i], contour_xy[2 * i + 1]};
  FPt b[4];
  *sside = mini_box(c, b);
  for (int i = 0; i < 4; ++i) { box8[2 * i] = b[i].x; box8[2 * i + 1] = b[i].y; }
  return DBB_OK;
}

// =================================================================================================================
// Polygon mode (src/postprocess.py:54-104): the ordered contour of a kept candidate, cv2.arcLength, cv2.approxPolyDP
// =================================================================================================================
namespace dbb {

// packed bitmap accessor: 1 bit per pixel, 32 pixels per word, zero outside the image
struct BitImage {
  const uint32_t* bits; int h, w, wq;
  inline int at(int y, int x) const {
    if ((unsigned)y >= (unsigned)h || (unsigned)x >= (unsigned)w) return 0;
    return (bits[(size_t)y * wq + (x >> 5)] >> (x & 31)) & 1u;
  }
};

// Border following of ONE border, as OpenCV's icvFetchContour (Suzuki-Abe) with CHAIN_APPROX_SIMPLE: `start` is the border's
// start pixel (the raster-first pixel of the component for an outer border; the foreground pixel left of the hole's
// raster-first pixel for a hole border); a point is emitted whenever the chain direction changes.
void trace_border(const BitImage& im, int sx, int sy, bool is_hole, std::vector<IPt>& out) {
  static const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1}, DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  out.clear();
  int s_end = is_hole ? 0 : 4, s = s_end;
  int i1x, i1y;
  do {
    s = (s - 1) & 7;
    i1x = sx + DX[s]; i1y = sy + DY[s];
  } while (im.at(i1y, i1x) == 0 && s != s_end);
  if (s == s_end) { out.push_back(IPt{sx, sy}); return; }       // single-pixel domain
  int i3x = sx, i3y = sy, i4x = sx, i4y = sy;
  int prev_s = s ^ 4;
  int px = sx, py = sy;
  for (;;) {
    s_end = s;
    while (s < 15) {
      ++s;
      i4x = i3x + DX[s & 7]; i4y = i3y + DY[s & 7];
      if (im.at(i4y, i4x) != 0) break;
    }
    s &= 7;
    if (s != prev_s) { out.push_back(IPt{px, py}); prev_s = s; }
    px += DX[s]; py += DY[s];
    if (i4x == sx && i4y == sy && i3x == i1x && i3y == i1y) break;
    i3x = i4x; i3y = i4y;
    s = (s + 4) & 7;
  }
}

// cv2.arcLength(contour, closed=True) for integer points: float differences, float squares, sqrt, double sum
double arc_length_closed(const std::vector<IPt>& c) {
  const int n = (int)c.size();
  if (n <= 1) return 0.0;
  double perimeter = 0;
  float prevx = (float)c[n - 1].x, prevy = (float)c[n - 1].y;
  for (int i = 0; i < n; ++i) {
    const float x = (float)c[i].x, y = (float)c[i].y;
    const float dx = x - prevx, dy = y - prevy;
    perimeter += std::sqrt(dx * dx + dy * dy);
    prevx = x; prevy = y;
  }
  return perimeter;
}

// cv2.approxPolyDP(contour, eps, closed=True) for integer points (imgproc/approx.cpp: approxPolyDP_<int>)
void approx_poly_dp_closed(const std::vector<IPt>& src, double eps, std::vector<IPt>& dst) {
  struct Range { int start, end; };
  const int count0 = (int)src.size();
  dst.clear();
  if (count0 == 0) return;
  int count = count0, new_count = 0, pos = 0, i, j;
  std::vector<IPt> out(count0);
  std::vector<Range> stack;
  Range slice{0, 0}, right_slice{0, 0};
  IPt start_pt{-1000000, -1000000}, end_pt{0, 0}, pt{0, 0};
  bool le_eps = false;
  auto read_pt = [&](IPt& p, int& ps) { p = src[ps]; if (++ps >= count) ps = 0; };
  eps *= eps;
  // 1. approximately the two farthest points of the contour
  right_slice.start = 0;
  for (i = 0; i < 3; ++i) {
    double max_dist = 0;
    pos = (pos + right_slice.start) % count;
    read_pt(start_pt, pos);
    for (j = 1; j < count; ++j) {
      read_pt(pt, pos);
      const double dx = pt.x - start_pt.x, dy = pt.y - start_pt.y;
      const double dist = dx * dx + dy * dy;
      if (dist > max_dist) { max_dist = dist; right_slice.start = j; }
    }
    le_eps = max_dist <= eps;
  }
  // 2. initialise the stack
  if (!le_eps) {
    right_slice.end = slice.start = pos % count;
    slice.end = right_slice.start = (right_slice.start + slice.start) % count;
    stack.push_back(right_slice);
    stack.push_back(slice);
  } else {
    out[new_count++] = start_pt;
  }
  // 3. recursive subdivision
  while (!stack.empty()) {
    slice = stack.back(); stack.pop_back();
    end_pt = src[slice.end];
    pos = slice.start;
    read_pt(start_pt, pos);
    if (pos != slice.end) {
      // Squared distance to the SEGMENT start..end (a point that projects beyond an end is measured to that end), not to
      // the infinite line: this is what OpenCV 4.13 does.  The line version agrees on ~99.8 % of contours; on long contours a
      // farthest point that lies beyond the chord's end differs (2 of 1,275 contours in a soak; with this rule 0 of 54,844
      // contour / epsilon pairs differ from cv2.approxPolyDP).
      double max_dist = 0;
      const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y, len2 = dx * dx + dy * dy;
      while (pos != slice.end) {
        read_pt(pt, pos);
        const double px = pt.x - start_pt.x, py = pt.y - start_pt.y;
        const double t = px * dx + py * dy;
        double dist;
        if (t < 0 || len2 == 0) dist = px * px + py * py;
        else if (t > len2) { const double qx = pt.x - end_pt.x, qy = pt.y - end_pt.y; dist = qx * qx + qy * qy; }
        else { const double cr = py * dx - px * dy; dist = cr * cr / len2; }
        if (dist > max_dist) { max_dist = dist; right_slice.start = (pos + count - 1) % count; }
      }
      le_eps = max_dist <= eps;                       // eps is squared (above)
    } else {
      le_eps = true;
      start_pt = src[slice.start];
    }
    if (le_eps) {
      out[new_count++] = start_pt;
    } else {
      right_slice.end = slice.end;
      slice.end = right_slice.start;
      stack.push_back(right_slice);
      stack.push_back(slice);
    }
  }
  // 4. clean-up: points on [almost] straight lines
  count = new_count;
  auto read_dst = [&](IPt& p, int& ps) { p = out[ps]; if (++ps >= count) ps = 0; };
  pos = count - 1;
  read_dst(start_pt, pos);
  int wpos = pos;
  read_dst(pt, pos);
  for (i = 0; i < count && new_count > 2; ++i) {
    read_dst(end_pt, pos);
    const double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
    const double dist = std::fabs((pt.x - start_pt.x) * dy - (pt.y - start_pt.y) * dx);
    const double sip = (double)(pt.x - start_pt.x) * (end_pt.x - pt.x) + (double)(pt.y - start_pt.y) * (end_pt.y - pt.y);
    if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
      --new_count;
      out[wpos] = start_pt = end_pt;
      if (++wpos >= count) wpos = 0;
      read_dst(pt, pos);
      ++i;
      continue;
    }
    out[wpos] = start_pt = pt;
    if (++wpos >= count) wpos = 0;
    pt = end_pt;
  }
  dst.assign(out.begin(), out.begin() + new_count);
}

}  // namespace dbb

// HOST test entries: one border of a byte bitmap / approxPolyDP of an integer contour
extern "C" int dbb_trace_contour(const uint8_t* bitmap, int64_t h, int64_t w, int start_x, int start_y, int is_hole, int32_t* out_xy, int cap) {
  if (!bitmap || !out_xy || h <= 0 || w <= 0) return dbb::set_error(DBB_EINVAL, "trace_contour: bad argument");
  const int wq = (int)((w + 31) / 32);
  std::vector<uint32_t> bits((size_t)h * wq, 0u);
  for (int64_t y = 0; y < h; ++y)
    for (int64_t x = 0; x < w; ++x)
      if (bitmap[y * w + x]) bits[(size_t)y * wq + (x >> 5)] |= 1u << (x & 31);
  dbb::BitImage im{bits.data(), (int)h, (int)w, wq};
  std::vector<dbb::IPt> c;
  dbb::trace_border(im, start_x, start_y, is_hole != 0, c);
  if ((int)c.size() > cap) return dbb::set_error(DBB_EWORKSPACE, "trace_contour: output buffer too small");
  for (size_t i = 0; i < c.size(); ++i) { out_xy[2 * i] = c[i].x; out_xy[2 * i + 1] = c[i].y; }
  return (int)c.size();
}
extern "C" int dbb_approx_poly_dp(const int32_t* contour_xy, int npts, double eps_or_negative_ratio, int32_t* out_xy, int cap, double* arc_length) {
  if (!contour_xy || npts < 0 || !out_xy) return dbb::set_error(DBB_EINVAL, "approx_poly_dp: bad argument");
  std::vector<dbb::IPt> c(npts), d;
  for (int i = 0; i < npts; ++i) c[i] = dbb::IPt{contour_xy[2 * i], contour_xy[2 * i + 1]};
  const double len = dbb::arc_length_closed(c);
  if (arc_length) *arc_length = len;
  const double eps = eps_or_negative_ratio < 0 ? -eps_or_negative_ratio * len : eps_or_negative_ratio;   // negative: a ratio of the arc length
  dbb::approx_poly_dp_closed(c, eps, d);
  if ((int)d.size() > cap) return dbb::set_error(DBB_EWORKSPACE, "approx_poly_dp: output buffer too small");
  for (size_t i = 0; i < d.size(); ++i) { out_xy[2 * i] = d[i].x; out_xy[2 * i + 1] = d[i].y; }
  return (int)d.size();
}

// ---------------------------------------------------------------------------------------------- polygon mode, whole batch
namespace dbb {
struct LPt64 { long long x, y; };
void clipper_offset_round(const std::vector<LPt64>& in, double delta, double arc_tolerance, std::vector<std::vector<LPt64>>& out);
}

// HOST function: src/postprocess.py:54-104 (polygons_from_bitmap) for every kept candidate of a batch.  bits: the packed bitmap
// (N, H, ceil(W/32)) uint32 copied back from the device front's workspace; cands / n_cands as in dbb_boxes_from_border_points.
// Outputs per image: counts (N, max_cands) int32 = number of points of candidate s's polygon (0 = dropped), points
// (N, cap, 2) int32 = the polygons back to back in candidate order, scores (N, max_cands) float64, totals (N) int32 = points
// the image produced (> cap: call again with a larger buffer).
extern "C" int dbb_polygons_from_bitmap(const uint32_t* bits, const DbbCandidate* cands, const int32_t* n_cands, int64_t n, int max_cands,
                                        int64_t h, int64_t w, const int32_t* dest_wh, float unclip_ratio, int min_size, int32_t* counts,
                                        int32_t* points, int cap, double* scores, int32_t* totals, int threads) {
  if (!bits || !cands || !n_cands || !dest_wh || !counts || !points || !scores || !totals || n <= 0 || max_cands <= 0 || cap <= 0)
    return set_error(DBB_EINVAL, "polygons_from_bitmap: bad argument");
  const int wq = (int)((w + 31) / 32);
  // Work items are the kept candidates of the whole batch, pulled from one atomic counter: a single image spreads over all
  // host threads (batch 1: 5.3 -> ~1 ms) and images with many candidates do not hold up the others.  Each item leaves its
  // rescaled polygon in its own vector; a cheap second pass concatenates them per image in candidate order.
  struct Item { int img, s; };
  std::vector<Item> items;
  for (int64_t img = 0; img < n; ++img) {
    const int k = n_cands[img] < max_cands ? n_cands[img] : max_cands;
    const DbbCandidate* C = cands + img * max_cands;
    for (int s = 0; s < k; ++s) if (C[s].keep) items.push_back(Item{(int)img, s});
  }
  std::vector<std::vector<int32_t>> out(items.size());
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    std::vector<IPt> contour, approx, ring;
    std::vector<LPt64> path;
    std::vector<std::vector<LPt64>> res;
    for (size_t it = next.fetch_add(1); it < items.size(); it = next.fetch_add(1)) {
      const int img = items[it].img, s = items[it].s;
      const DbbCandidate& C = cands[(int64_t)img * max_cands + s];
      BitImage im{bits + (size_t)img * h * wq, (int)h, (int)w, wq};
      const double dw = (double)dest_wh[2 * img], dh = (double)dest_wh[2 * img + 1];
      trace_border(im, C.kind ? C.first_x - 1 : C.first_x, C.first_y, C.kind != 0, contour);
      approx_poly_dp_closed(contour, 0.005 * arc_length_closed(contour), approx);
      const int np_ = (int)approx.size();
      if (np_ < 4) continue;
      double area = 0, length = 0;
      for (int i = 0; i < np_; ++i) {
        const IPt& a = approx[i]; const IPt& b = approx[(i + 1) % np_];
        area += (double)a.x * b.y - (double)a.y * b.x;
        const double dx = (double)a.x - b.x, dy = (double)a.y - b.y;
        length += std::sqrt(dx * dx + dy * dy);
      }
      area = std::fabs(area) * 0.5;
      if (!(length > 0)) continue;
      const double distance = area * (double)unclip_ratio / length;
      path.resize(np_);
      for (int i = 0; i < np_; ++i) path[i] = LPt64{approx[i].x, approx[i].y};
      clipper_offset_round(path, distance, 0.25, res);
      if (res.size() != 1) continue;                            // len(box) > 1 -> dropped (an empty result cannot be reshaped either)
      ring.resize(res[0].size());
      for (size_t i = 0; i < res[0].size(); ++i) ring[i] = IPt{(int)res[0][i].x, (int)res[0][i].y};
      FPt bx[4];
      std::vector<IPt> tmp(ring);
      if (mini_box(tmp, bx) < (float)(min_size + 2)) continue;
      std::vector<int32_t>& o = out[it];
      o.resize(2 * ring.size());
      for (size_t i = 0; i < ring.size(); ++i) {
        // box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)   (float64, half to even, into an int array)
        double fx = std::nearbyint((double)ring[i].x / (double)w * dw), fy = std::nearbyint((double)ring[i].y / (double)h * dh);
        fx = fx < 0 ? 0 : (fx > dw ? dw : fx); fy = fy < 0 ? 0 : (fy > dh ? dh : fy);
        o[2 * i] = (int32_t)fx; o[2 * i + 1] = (int32_t)fy;
      }
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if ((size_t)nt > items.size()) nt = (int)items.size();
  if (nt <= 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  // ---- per image, in candidate order
  std::fill(counts, counts + n * max_cands, 0);
  std::fill(scores, scores + n * max_cands, 0.0);
  std::fill(totals, totals + n, 0);
  for (size_t it = 0; it < items.size(); ++it) {
    const std::vector<int32_t>& o = out[it];
    if (o.empty()) continue;
    const int img = items[it].img, sidx = items[it].s, m = (int)(o.size() / 2);
    const DbbCandidate& C = cands[(int64_t)img * max_cands + sidx];
    int32_t* P = points + (int64_t)img * cap * 2;
    if (totals[img] + m <= cap) std::copy(o.begin(), o.end(), P + 2 * (int64_t)totals[img]);
    totals[img] += m;
    counts[(int64_t)img * max_cands + sidx] = m;
    scores[(int64_t)img * max_cands + sidx] = C.sum / (double)C.count;
  }
  return DBB_OK;
}
