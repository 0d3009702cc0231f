// post_geom.h -- host geometry of the post-processing back half (post_geom.cu, clipper_offset.cu)
#pragma once
#include <stdint.h>
#include <vector>

namespace dbb {

struct IPt { int x, y; };
struct FPt { float x, y; };
struct RotRect { float cx, cy, w, h, angle; };

void convex_hull(std::vector<IPt>& pts, std::vector<IPt>& hull);
RotRect min_area_rect(std::vector<IPt>& pts);
void box_points(const RotRect& r, FPt pt[4]);
float mini_box(std::vector<IPt>& contour, FPt box[4]);
void offset_convex_round(const IPt* in, int n_in, double delta, std::vector<IPt>& out, double arc_tolerance);

}  // namespace dbb
