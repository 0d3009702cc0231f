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
struct LPt64 { long long x, y; };
typedef LPt64 LPt;
static inline bool operator==(const LPt& a, const LPt& b) { return a.x == b.x && a.y == b.y; }
static inline bool operator!=(const LPt& a, const LPt& b) { return !(a == b); }
static inline bool operator<(const LPt& a, const LPt& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }
typedef std::vector<LPt> LPath;

static inline i64 c_round(double v) { return v < 0 ? (i64)(v - 0.5) : (i64)(v + 0.5); }
// coordinates are pixel positions (|v| < 2^30, checked on entry: Clipper's own "loRange"), so 64-bit products are exact
static inline i64 crossv(i64 ax, i64 ay, i64 bx, i64 by) { return ax * by - ay * bx; }

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

// Intersection points are snapped to a grid SUB times finer than the integer pixel grid of the input: snapping to the pixel grid
// itself (what Clipper's IntersectPoint does) moves a sub-edge by up to half a pixel, which for the 1-2 pixel arc pieces of a
// small offset creates crossings that were not there -- the arrangement stops being planar and faces / winding numbers go
// wrong (observed on 119 of 876 approxPolyDP polygons at delta ~ 4 px).  Result vertices are rounded back to pixels at the end.
constexpr i64 SUB = 1024;
void positive_region(const LPath& path, std::vector<LPath>& result) {
  result.clear();
  // ---- segments (zero-length dropped)
  LPath P;
  for (const LPt& p0 : path) { const LPt p{p0.x * SUB, p0.y * SUB}; if (P.empty() || P.back() != p) P.push_back(p); }
  while (P.size() > 1 && P.front() == P.back()) P.pop_back();
  const int n = (int)P.size();
  if (n < 3) return;
  // ---- split points per segment
  struct Cut { int seg; LPt p; };
  std::vector<Cut> cuts;
  cuts.reserve(64);
  auto seg_a = [&](int i) -> const LPt& { return P[i]; };
  auto seg_b = [&](int i) -> const LPt& { return P[(i + 1) % n]; };
  auto strictly_inside = [](const LPt& a, const LPt& b, const LPt& q) {      // q on segment ab (collinear assumed), not an end point
    if (q == a || q == b) return false;
    return std::min(a.x, b.x) <= q.x && q.x <= std::max(a.x, b.x) && std::min(a.y, b.y) <= q.y && q.y <= std::max(a.y, b.y);
  };
  // sweep and prune on x: segments sorted by their smaller x; a pair is tested only while the x ranges overlap (the raw
  // offset path consists of short arc pieces, so this is near-linear instead of the n^2 / 2 pair tests of the naive loop)
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int p, int q) { return std::min(seg_a(p).x, seg_b(p).x) < std::min(seg_a(q).x, seg_b(q).x); });
  for (int oi = 0; oi < n; ++oi) {
    const int i0 = order[oi];
    const i64 maxx_i = std::max(seg_a(i0).x, seg_b(i0).x);
    for (int oj = oi + 1; oj < n; ++oj) {
      const int j0 = order[oj];
      if (std::min(seg_a(j0).x, seg_b(j0).x) > maxx_i) break;
      const int i = i0 < j0 ? i0 : j0, j = i0 < j0 ? j0 : i0;
      const LPt &a = seg_a(i), &b = seg_b(i);
      const i64 d1x = b.x - a.x, d1y = b.y - a.y;
      const LPt &c = seg_a(j), &d = seg_b(j);
      if (std::max(a.y, b.y) < std::min(c.y, d.y) || std::max(c.y, d.y) < std::min(a.y, b.y)) continue;
      const i64 d2x = d.x - c.x, d2y = d.y - c.y;
      const i64 den = crossv(d1x, d1y, d2x, d2y);
      const i64 tn = crossv(c.x - a.x, c.y - a.y, d2x, d2y);      // t = tn / den along ab
      const i64 un = crossv(c.x - a.x, c.y - a.y, d1x, d1y);      // u = un / den along cd
      if (den == 0) {
        if (un != 0) continue;                                            // parallel, not collinear
        if (strictly_inside(a, b, c)) cuts.push_back(Cut{i, c});
        if (strictly_inside(a, b, d)) cuts.push_back(Cut{i, d});
        if (strictly_inside(c, d, a)) cuts.push_back(Cut{j, a});
        if (strictly_inside(c, d, b)) cuts.push_back(Cut{j, b});
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
      if (ip != a && ip != b) cuts.push_back(Cut{i, ip});
      if (ip != c && ip != d) cuts.push_back(Cut{j, ip});
    }
  }
  // ---- sub-edges between integer vertices, net weight per undirected pair (flat sorted arrays: no node allocations)
  std::sort(cuts.begin(), cuts.end(), [&](const Cut& p, const Cut& q) {
    if (p.seg != q.seg) return p.seg < q.seg;
    const LPt& a = seg_a(p.seg); const LPt& b = seg_b(p.seg);
    const i64 dx = b.x - a.x, dy = b.y - a.y;
    return (p.p.x - a.x) * dx + (p.p.y - a.y) * dy < (q.p.x - a.x) * dx + (q.p.y - a.y) * dy; });
  struct Sub { LPt a, b; };
  std::vector<Sub> subs;
  subs.reserve(n + cuts.size());
  {
    size_t ci = 0;
    for (int i = 0; i < n; ++i) {
      LPt prev = seg_a(i);
      const LPt b = seg_b(i);
      for (; ci < cuts.size() && cuts[ci].seg == i; ++ci) {
        if (cuts[ci].p != prev) { subs.push_back(Sub{prev, cuts[ci].p}); prev = cuts[ci].p; }
      }
      if (b != prev) subs.push_back(Sub{prev, b});
    }
  }
  std::vector<LPt> V;
  V.reserve(2 * subs.size());
  for (const Sub& e : subs) { V.push_back(e.a); V.push_back(e.b); }
  std::sort(V.begin(), V.end());
  V.erase(std::unique(V.begin(), V.end()), V.end());
  auto vertex = [&](const LPt& p) { return (int)(std::lower_bound(V.begin(), V.end(), p) - V.begin()); };
  struct Net { int u, v, w; };
  std::vector<Net> net;
  net.reserve(subs.size());
  for (const Sub& e : subs) {
    const int u = vertex(e.a), v = vertex(e.b);
    if (u < v) net.push_back(Net{u, v, 1}); else net.push_back(Net{v, u, -1});
  }
  std::sort(net.begin(), net.end(), [](const Net& p, const Net& q) { return p.u < q.u || (p.u == q.u && p.v < q.v); });
  // ---- half-edges, sorted around every vertex
  std::vector<HalfEdge> H;
  H.reserve(2 * net.size());
  for (size_t e = 0; e < net.size();) {
    size_t f = e;
    int w = 0;
    while (f < net.size() && net[f].u == net[e].u && net[f].v == net[e].v) { w += net[f].w; ++f; }
    const int h = (int)H.size();
    H.push_back(HalfEdge{net[e].u, net[e].v, h + 1, w, -1, false});
    H.push_back(HalfEdge{net[e].v, net[e].u, h, -w, -1, false});
    e = f;
  }
  const int nv = (int)V.size();
  // CSR adjacency: out_of[v] = ostart[v] .. ostart[v + 1]
  std::vector<int> ostart(nv + 1, 0), oidx(H.size()), pos_in(H.size());
  for (const HalfEdge& he : H) ++ostart[he.from + 1];
  for (int v = 0; v < nv; ++v) ostart[v + 1] += ostart[v];
  {
    std::vector<int> fillp(ostart.begin(), ostart.end() - 1);
    for (int h = 0; h < (int)H.size(); ++h) oidx[fillp[H[h].from]++] = h;
  }
  for (int v = 0; v < nv; ++v) {
    std::sort(oidx.begin() + ostart[v], oidx.begin() + ostart[v + 1], [&](int h1, int h2) {
      return angle_less(V[H[h1].to].x - V[v].x, V[H[h1].to].y - V[v].y, V[H[h2].to].x - V[v].x, V[H[h2].to].y - V[v].y); });
    for (int q = ostart[v]; q < ostart[v + 1]; ++q) pos_in[oidx[q]] = q - ostart[v];
  }
  struct OutView { const int* p; int m; int operator[](int i) const { return p[i]; } int size() const { return m; } };
  auto out_of_v = [&](int v) { return OutView{oidx.data() + ostart[v], ostart[v + 1] - ostart[v]}; };
  // next half-edge along the face to the LEFT of h: at h.to, the outgoing edge just clockwise of twin(h)
  auto next_of = [&](int h) {
    const int t = H[h].twin, v = H[t].from;
    const OutView ov = out_of_v(v);
    const int q = pos_in[t], m = ov.size();
    return ov[(q - 1 + m) % m];
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
  // (faces are few and the graph is small: relax over all half-edges until every face is labelled)
  seen[outer] = 1;
  for (int left = nfaces - 1, guard = 0; left > 0 && guard <= nfaces; ++guard) {
    for (int h = 0; h < (int)H.size(); ++h) {
      const int f = H[h].face, g = H[H[h].twin].face;      // h has f on its left; twin(h) has g on its left
      if (seen[f] && !seen[g]) { W[g] = W[f] - H[h].weight; seen[g] = 1; --left; }      // W(f) = W(g) + weight(h)
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
      const int t = H[h].twin, v = H[t].from;
      const OutView ov = out_of_v(v);
      const int m = ov.size();
      int q = pos_in[t], nxt = -1;
      for (int s = 1; s <= m; ++s) {
        const int cand = ov[((q - s) % m + m) % m];
        if (selected(cand) && !H[cand].used) { nxt = cand; break; }
        if (cand == h0) { nxt = h0; break; }
      }
      if (nxt < 0) break;
      h = nxt;
    } while (h != h0);
    // back to the pixel grid (Clipper's Round()), duplicates out
    {
      LPath px;
      for (const LPt& q : loop) {
        const LPt r{c_round((double)q.x / (double)SUB), c_round((double)q.y / (double)SUB)};
        if (px.empty() || px.back() != r) px.push_back(r);
      }
      while (px.size() > 1 && px.front() == px.back()) px.pop_back();
      loop.swap(px);
    }
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
  for (const LPt& p : in) if (p.x > (1ll << 18) || p.x < -(1ll << 18) || p.y > (1ll << 18) || p.y < -(1ll << 18)) return;   // outside the exact int64 range of the x 1024 sub-grid
  if (!(std::fabs(delta) < (double)(1ll << 18))) return;
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
