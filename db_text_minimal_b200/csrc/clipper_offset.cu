// clipper_offset.cu -- host C++ (no CUDA): pyclipper.PyclipperOffset().AddPath(path, JT_ROUND, ET_CLOSEDPOLYGON) / Execute(delta)
// for ONE closed polygon of any shape (convex or not), as the reference uses it at src/postprocess.py:150-156 (unclip of the
// approxPolyDP polygon in polygon mode, of the min-area box in box mode), src/data_loaders.py:116-122 (shrink, delta < 0)
// and src/db_transforms.py:13-21 (dilate).
//
// The arithmetic lives in a third-party dependency that is NOT in the reference tree and not installable here:
// pyclipper==1.1.0.post3 (requirements.txt:67) = Angus Johnson's Clipper 6.4.2.  What is restated from its published source:
//   * ClipperOffset::AddPath      duplicate stripping, closed-path handling
//   * FixOrientations             a polygon of negative Area() is reversed
//   * DoOffset / OffsetPoint / DoRound / GetUnitNormal: arc tolerance 0.25 (pyclipper's default), steps, m_sin / m_cos
//                                 recurrence, Round() = (cInt)(v +- 0.5), the concave-vertex triple (p + n_k d, p, p + n_j d),
//                                 and OffsetPoint's early return that does NOT advance k on near-collinear vertices
//   * Execute: the union of the raw offset path under the POSITIVE fill rule (delta > 0), resp. the bounding-rectangle /
//     pftNegative / ReverseSolution construction for delta < 0 -- both select exactly {winding number of the raw path > 0}.
// What is NOT a restatement: Clipper computes that union with its Vatti scan-line clipper; here the same region is extracted
// from the planar arrangement of the raw path (all pairwise intersections, rounded to the integer grid as Clipper rounds
// them; faces; winding numbers; the boundary between winding > 0 and <= 0).  The point SET of a result polygon can differ
// from Clipper's in start vertex, in collinear points and by the rounding of intersection points; the region is the same.
// PARITY UNPINNED: there is no pyclipper here (or in the reference tree) to generate goldens; tests check the convex case
// against the closed form, the region against winding numbers of the raw path, and area / perimeter identities.
#include "common.cuh"
#include "post_geom.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

namespace dbb {

typedef long long i64;
struct LPt { i64 x, y; };
static inline bool operator==(const LPt& a, const LPt& b) { return a.x == b.x && a.y == b.y; }
static inline bool operator!=(const LPt& a, const LPt& b) { return !(a == b); }
static inline bool operator<(const LPt& a, const LPt& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }
typedef std::vector<LPt> LPath;

static inline i64 c_round(double v) { return v < 0 ? (i64)(v - 0.5) : (i64)(v + 0.5); }
static inline __int128 crossv(i64 ax, i64 ay, i64 bx, i64 by) { return (__int128)ax * by - (__int128)ay * bx; }

// Clipper's Area(): positive for counter-clockwise paths in a y-up frame
static double clipper_area(const LPath& p) {
  const int n = (int)p.size();
  if (n < 3) return 0;
  double a = 0;
  for (int i = 0, j = n - 1; i < n; ++i) { a += ((double)p[j].x + p[i].x) * ((double)p[j].y - p[i].y); j = i; }
  return -a * 0.5;
}

// ---------------------------------------------------------------------------------------------- raw offset path
static bool raw_offset(const LPath& in, double delta, double arc_tolerance, LPath& out) {
  // AddPath: strip closing duplicates and consecutive duplicates
  int highI = (int)in.size() - 1;
  if (highI < 0) return false;
  while (highI > 0 && in[0] == in[highI]) --highI;
  LPath src;
  src.push_back(in[0]);
  for (int i = 1; i <= highI; ++i) if (src.back() != in[i]) src.push_back(in[i]);
  if ((int)src.size() < 3) return false;
  if (clipper_area(src) < 0) std::reverse(src.begin(), src.end());          // FixOrientations
  const int len = (int)src.size();
  out.clear();
  if (std::fabs(delta) < 1e-20) { out = src; return true; }
  const double pi = 3.141592653589793238, two_pi = pi * 2, def_arc = 0.25;
  double y;
  if (arc_tolerance <= 0.0) y = def_arc;
  else if (arc_tolerance > std::fabs(delta) * def_arc) y = std::fabs(delta) * def_arc;
  else y = arc_tolerance;
  double steps = pi / std::acos(1 - y / std::fabs(delta));
  if (steps > std::fabs(delta) * pi) steps = std::fabs(delta) * pi;
  double m_sin = std::sin(two_pi / steps);
  const double m_cos = std::cos(two_pi / steps), steps_per_rad = steps / two_pi;
  if (delta < 0.0) m_sin = -m_sin;
  std::vector<double> nx(len), ny(len);
  for (int j = 0; j < len; ++j) {
    const LPt& a = src[j]; const LPt& b = src[(j + 1) % len];
    double dx = (double)(b.x - a.x), dy = (double)(b.y - a.y);
    const double f = 1.0 / std::sqrt(dx * dx + dy * dy);
    dx *= f; dy *= f;
    nx[j] = dy; ny[j] = -dx;
  }
  auto push = [&](double x, double yv) { out.push_back(LPt{c_round(x), c_round(yv)}); };
  int k = len - 1;
  for (int j = 0; j < len; ++j) {
    double sinA = nx[k] * ny[j] - nx[j] * ny[k];
    if (std::fabs(sinA * delta) < 1.0) {
      const double cosA = nx[k] * nx[j] + ny[j] * ny[k];
      if (cosA > 0) {                           // angle ~ 0: one point, and k is NOT advanced (Clipper 6.4.2 returns here)
        push(src[j].x + nx[k] * delta, src[j].y + ny[k] * delta);
        continue;
      }
    } else if (sinA > 1.0) sinA = 1.0;
    else if (sinA < -1.0) sinA = -1.0;
    if (sinA * delta < 0) {                     // concave for this offset direction
      push(src[j].x + nx[k] * delta, src[j].y + ny[k] * delta);
      out.push_back(src[j]);
      push(src[j].x + nx[j] * delta, src[j].y + ny[j] * delta);
    } else {                                    // DoRound
      const double a = std::atan2(sinA, nx[k] * nx[j] + ny[k] * ny[j]);
      int st = (int)c_round(steps_per_rad * std::fabs(a));
      if (st < 1) st = 1;
      double X = nx[k], Y = ny[k];
      for (int q = 0; q < st; ++q) {
        push(src[j].x + X * delta, src[j].y + Y * delta);
        const double X2 = X;
        X = X * m_cos - m_sin * Y;
        Y = X2 * m_sin + Y * m_cos;
      }
      push(src[j].x + nx[j] * delta, src[j].y + ny[j] * delta);
    }
    k = j;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------- {winding > 0} of a closed path
namespace {

struct HalfEdge { int from, to, twin, weight, face; bool used; };

// direction ordering by angle in [0, 2pi): exact on integers
inline int half_of(i64 dx, i64 dy) { return (dy > 0 || (dy == 0 && dx > 0)) ? 0 : 1; }
inline bool angle_less(i64 ax, i64 ay, i64 bx, i64 by) {
  const int ha = half_of(ax, ay), hb = half_of(bx, by);
  if (ha != hb) return ha < hb;
  return crossv(ax, ay, bx, by) > 0;
}

void positive_region(const LPath& path, std::vector<LPath>& result) {
  result.clear();
  // ---- segments (zero-length dropped)
  LPath P;
  for (const LPt& p : path) if (P.empty() || P.back() != p) P.push_back(p);
  while (P.size() > 1 && P.front() == P.back()) P.pop_back();
  const int n = (int)P.size();
  if (n < 3) return;
  // ---- split points per segment
  std::vector<std::vector<LPt>> cuts(n);
  auto seg_a = [&](int i) -> const LPt& { return P[i]; };
  auto seg_b = [&](int i) -> const LPt& { return P[(i + 1) % n]; };
  auto strictly_inside = [](const LPt& a, const LPt& b, const LPt& q) {      // q on segment ab (collinear assumed), not an end point
    if (q == a || q == b) return false;
    return std::min(a.x, b.x) <= q.x && q.x <= std::max(a.x, b.x) && std::min(a.y, b.y) <= q.y && q.y <= std::max(a.y, b.y);
  };
  for (int i = 0; i < n; ++i) {
    const LPt &a = seg_a(i), &b = seg_b(i);
    const i64 d1x = b.x - a.x, d1y = b.y - a.y;
    for (int j = i + 1; j < n; ++j) {
      const LPt &c = seg_a(j), &d = seg_b(j);
      if (std::max(a.x, b.x) < std::min(c.x, d.x) || std::max(c.x, d.x) < std::min(a.x, b.x) ||
          std::max(a.y, b.y) < std::min(c.y, d.y) || std::max(c.y, d.y) < std::min(a.y, b.y)) continue;
      const i64 d2x = d.x - c.x, d2y = d.y - c.y;
      const __int128 den = crossv(d1x, d1y, d2x, d2y);
      const __int128 tn = crossv(c.x - a.x, c.y - a.y, d2x, d2y);      // t = tn / den along ab
      const __int128 un = crossv(c.x - a.x, c.y - a.y, d1x, d1y);      // u = un / den along cd
      if (den == 0) {
        if (un != 0) continue;                                            // parallel, not collinear
        if (strictly_inside(a, b, c)) cuts[i].push_back(c);
        if (strictly_inside(a, b, d)) cuts[i].push_back(d);
        if (strictly_inside(c, d, a)) cuts[j].push_back(a);
        if (strictly_inside(c, d, b)) cuts[j].push_back(b);
        continue;
      }
      const bool pos = den > 0;
      const bool t_in = pos ? (tn >= 0 && tn <= den) : (tn <= 0 && tn >= den);
      const bool u_in = pos ? (un >= 0 && un <= den) : (un <= 0 && un >= den);
      if (!t_in || !u_in) continue;
      LPt ip;
      if (tn == 0) ip = a; else if (tn == den) ip = b; else if (un == 0) ip = c; else if (un == den) ip = d;
      else {
        const double t = (double)tn / (double)den;
        ip.x = c_round((double)a.x + t * (double)d1x);
        ip.y = c_round((double)a.y + t * (double)d1y);
      }
      if (ip != a && ip != b) cuts[i].push_back(ip);
      if (ip != c && ip != d) cuts[j].push_back(ip);
    }
  }
  // ---- sub-edges between integer vertices, net weight per undirected pair
  std::map<LPt, int> vid;
  std::vector<LPt> V;
  auto vertex = [&](const LPt& p) { auto it = vid.find(p); if (it != vid.end()) return it->second; vid[p] = (int)V.size(); V.push_back(p); return (int)V.size() - 1; };
  std::map<std::pair<int, int>, int> net;      // (min id, max id) -> forward count in the min -> max direction
  for (int i = 0; i < n; ++i) {
    const LPt a = seg_a(i), b = seg_b(i);
    std::vector<LPt>& c = cuts[i];
    const i64 dx = b.x - a.x, dy = b.y - a.y;
    std::sort(c.begin(), c.end(), [&](const LPt& p, const LPt& q) {
      return (__int128)(p.x - a.x) * dx + (__int128)(p.y - a.y) * dy < (__int128)(q.x - a.x) * dx + (__int128)(q.y - a.y) * dy; });
    LPt prev = a;
    auto emit = [&](const LPt& q) {
      if (q == prev) return;
      const int u = vertex(prev), v = vertex(q);
      if (u < v) net[{u, v}] += 1; else net[{v, u}] -= 1;
      prev = q;
    };
    for (const LPt& q : c) emit(q);
    emit(b);
  }
  // ---- half-edges, sorted around every vertex
  std::vector<HalfEdge> H;
  for (const auto& kv : net) {
    const int u = kv.first.first, v = kv.first.second, w = kv.second;
    const int h = (int)H.size();
    H.push_back(HalfEdge{u, v, h + 1, w, -1, false});
    H.push_back(HalfEdge{v, u, h, -w, -1, false});
  }
  const int nv = (int)V.size();
  std::vector<std::vector<int>> out_of(nv);
  for (int h = 0; h < (int)H.size(); ++h) out_of[H[h].from].push_back(h);
  std::vector<int> pos_in(H.size());
  for (int v = 0; v < nv; ++v) {
    std::sort(out_of[v].begin(), out_of[v].end(), [&](int h1, int h2) {
      return angle_less(V[H[h1].to].x - V[v].x, V[H[h1].to].y - V[v].y, V[H[h2].to].x - V[v].x, V[H[h2].to].y - V[v].y); });
    for (int q = 0; q < (int)out_of[v].size(); ++q) pos_in[out_of[v][q]] = q;
  }
  // next half-edge along the face to the LEFT of h: at h.to, the outgoing edge just clockwise of twin(h)
  auto next_of = [&](int h) {
    const int t = H[h].twin, v = H[t].from;
    const int q = pos_in[t], m = (int)out_of[v].size();
    return out_of[v][(q - 1 + m) % m];
  };
  // ---- faces
  int nfaces = 0;
  std::vector<double> farea;
  for (int h = 0; h < (int)H.size(); ++h) {
    if (H[h].face >= 0) continue;
    double a2 = 0;
    int g = h;
    do {
      H[g].face = nfaces;
      const LPt &p = V[H[g].from], &q = V[H[g].to];
      a2 += (double)p.x * (double)q.y - (double)q.x * (double)p.y;
      g = next_of(g);
    } while (g != h);
    farea.push_back(a2);
    ++nfaces;
  }
  if (nfaces == 0) return;
  // the unbounded face: the (only) cycle traversed clockwise -- most negative area (the graph of one closed path is connected)
  int outer = 0;
  for (int f = 1; f < nfaces; ++f) if (farea[f] < farea[outer]) outer = f;
  // ---- winding numbers by propagation across edges: W(left of h) = W(right of h) + weight(h)
  std::vector<int> W(nfaces, 0);
  std::vector<char> seen(nfaces, 0);
  std::vector<std::vector<int>> edges_of(nfaces);
  for (int h = 0; h < (int)H.size(); ++h) edges_of[H[h].face].push_back(h);
  std::vector<int> stack{outer};
  seen[outer] = 1;
  while (!stack.empty()) {
    const int f = stack.back(); stack.pop_back();
    for (int h : edges_of[f]) {              // h has f on its left; twin(h) has the neighbour on its left
      const int g = H[H[h].twin].face;
      if (seen[g]) continue;
      W[g] = W[f] - H[h].weight;             // W(f) = W(g) + weight(h)
      seen[g] = 1;
      stack.push_back(g);
    }
  }
  // ---- boundary half-edges: winding > 0 on the left, <= 0 on the right
  auto selected = [&](int h) { return W[H[h].face] > 0 && W[H[H[h].twin].face] <= 0; };
  for (int h0 = 0; h0 < (int)H.size(); ++h0) {
    if (H[h0].used || !selected(h0)) continue;
    LPath loop;
    int h = h0;
    do {
      H[h].used = true;
      loop.push_back(V[H[h].from]);
      // continue with the first selected outgoing edge clockwise of twin(h) (keeps the region on the left, splits at touching points)
      const int t = H[h].twin, v = H[t].from, m = (int)out_of[v].size();
      int q = pos_in[t], nxt = -1;
      for (int s = 1; s <= m; ++s) {
        const int cand = out_of[v][((q - s) % m + m) % m];
        if (selected(cand) && !H[cand].used) { nxt = cand; break; }
        if (cand == h0) { nxt = h0; break; }
      }
      if (nxt < 0) break;
      h = nxt;
    } while (h != h0);
    // cleanup: collinear vertices out (Clipper's FixupOutPolygon without PreserveCollinear)
    bool changed = true;
    while (changed && loop.size() >= 3) {
      changed = false;
      for (size_t i = 0; i < loop.size() && loop.size() >= 3; ++i) {
        const LPt& a = loop[(i + loop.size() - 1) % loop.size()]; const LPt& b = loop[i]; const LPt& c = loop[(i + 1) % loop.size()];
        if (crossv(b.x - a.x, b.y - a.y, c.x - b.x, c.y - b.y) == 0) { loop.erase(loop.begin() + i); changed = true; break; }
      }
    }
    if (loop.size() >= 3 && clipper_area(loop) != 0) result.push_back(loop);
  }
  // outer polygons (positive area) first, as Clipper lists them before their holes for a single subject
  std::stable_sort(result.begin(), result.end(), [](const LPath& a, const LPath& b) { return (clipper_area(a) > 0) > (clipper_area(b) > 0); });
}

}  // namespace

void clipper_offset_round(const LPath& in, double delta, double arc_tolerance, std::vector<LPath>& out) {
  LPath raw;
  out.clear();
  if (!raw_offset(in, delta, arc_tolerance, raw)) return;
  positive_region(raw, out);
}

}  // namespace dbb

using namespace dbb;

// HOST function.  path_xy: (npts, 2) int64 (pyclipper truncates float input to its 64-bit integer type; callers pass the
// truncated values).  Writes up to max_paths result polygons, concatenated, into out_xy (capacity cap_points points);
// out_counts[i] = points of polygon i.  Returns the number of polygons (>= 0), or a negative DBB_E* code
// (DBB_EWORKSPACE when the output buffers are too small).
extern "C" int dbb_clipper_offset(const int64_t* path_xy, int npts, double delta, double arc_tolerance, int64_t* out_xy,
                                  int cap_points, int32_t* out_counts, int max_paths) {
  if (!path_xy || npts < 0 || !out_xy || !out_counts || cap_points < 0 || max_paths <= 0) return set_error(DBB_EINVAL, "clipper_offset: bad argument");
  LPath in(npts);
  for (int i = 0; i < npts; ++i) in[i] = LPt{(i64)path_xy[2 * i], (i64)path_xy[2 * i + 1]};
  std::vector<LPath> res;
  clipper_offset_round(in, delta, arc_tolerance, res);
  if ((int)res.size() > max_paths) return set_error(DBB_EWORKSPACE, "clipper_offset: more result polygons than max_paths");
  int total = 0;
  for (size_t i = 0; i < res.size(); ++i) {
    if (total + (int)res[i].size() > cap_points) return set_error(DBB_EWORKSPACE, "clipper_offset: output buffer too small");
    for (const LPt& p : res[i]) { out_xy[2 * total] = p.x; out_xy[2 * total + 1] = p.y; ++total; }
    out_counts[i] = (int)res[i].size();
  }
  return (int)res.size();
}
// the raw offset path before the union (tests: the region is checked against the winding numbers of this path)
extern "C" int dbb_clipper_offset_raw(const int64_t* path_xy, int npts, double delta, double arc_tolerance, int64_t* out_xy, int cap_points) {
  if (!path_xy || npts < 0 || !out_xy) return set_error(DBB_EINVAL, "clipper_offset_raw: bad argument");
  LPath in(npts), raw;
  for (int i = 0; i < npts; ++i) in[i] = LPt{(i64)path_xy[2 * i], (i64)path_xy[2 * i + 1]};
  if (!raw_offset(in, delta, arc_tolerance, raw)) return 0;
  if ((int)raw.size() > cap_points) return set_error(DBB_EWORKSPACE, "clipper_offset_raw: output buffer too small");
  for (size_t i = 0; i < raw.size(); ++i) { out_xy[2 * i] = raw[i].x; out_xy[2 * i + 1] = raw[i].y; }
  return (int)raw.size();
}
