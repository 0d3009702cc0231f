"""Host-side geometry of the post-processing back half (csrc/post_geom.cu, csrc/clipper_offset.cu): plain C++ behind the C
ABI, so it runs without a GPU.

get_mini_boxes (cv2.minAreaRect + cv2.boxPoints + corner ordering) is PINNED: against the rows the reference's own
get_mini_boxes produced for every contour of tests/golden/post_cases.npz.
The ClipperOffset restatement is UNPINNED (pyclipper / Clipper 6.4.2 is neither in the reference tree nor installable here):
the tests below check what can be checked without it -- the convex case against the closed form of a round-join offset,
the result region against the winding numbers of the raw offset path, and area / perimeter identities.
"""
import ctypes as C
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def lib():
    from db_text_minimal_b200 import _lib
    return _lib.lib()


def mini_box(contour):
    c = np.ascontiguousarray(np.asarray(contour).reshape(-1, 2).astype(np.int32))
    box, ss = np.zeros(8, np.float32), C.c_float()
    assert lib().dbb_mini_box(c.ctypes.data, len(c), box.ctypes.data, C.byref(ss)) == 0
    return box.reshape(4, 2), ss.value


def test_mini_box_matches_reference_rows_on_every_golden_contour():
    """sside and the four ORDERED corners of get_mini_boxes (src/postprocess.py:158-184) for all 3,060 golden contours."""
    import cv2
    z = np.load(os.path.join(GOLD, "post_cases.npz"))
    names = sorted({k.split(":")[0] for k in z.files})
    total = 0
    for nm in names:
        contours, _ = cv2.findContours((z[nm + ":bitmap"] * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        for c, r in zip(contours[:1000], z[nm + ":cands"]):
            box, ss = mini_box(c)
            assert abs(ss - r[1]) <= 1e-3, (nm, ss, r[1])
            np.testing.assert_allclose(box.reshape(-1), r[8:16], atol=1e-3, err_msg=nm)
            total += 1
    assert total == 3060


def test_mini_box_only_needs_the_hull():
    """cv2.minAreaRect(contour) == minAreaRect(run end points of the component): what the device emits is enough."""
    import cv2
    from oracle import db_oracle as O
    bm = (O.synth_prob_map(200, 260, 5) > 0.55).astype(np.uint8)
    contours, _ = cv2.findContours(bm * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    n = 0
    for c in contours:
        mask = np.zeros_like(bm)
        cv2.drawContours(mask, [c], -1, 1, thickness=-1)
        mask &= bm
        ys, xs = np.nonzero(mask)
        ends = []
        for y in np.unique(ys):
            row = xs[ys == y]
            ends += [(row.min(), y), (row.max(), y)]
        a, sa = mini_box(c)
        b, sb = mini_box(np.array(ends))
        assert abs(sa - sb) <= 1e-4 and np.abs(a - b).max() <= 1e-3
        n += 1
    assert n > 5


def poly_area_perimeter(p):
    p = np.asarray(p, np.float64)
    q = np.roll(p, -1, 0)
    return 0.5 * abs(np.sum(p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0])), np.sqrt(((p - q) ** 2).sum(1)).sum()


def winding_number(path, pts):
    """Winding number of each point w.r.t. a closed integer path (points on the path excluded by the caller)."""
    p = np.asarray(path, np.float64)
    q = np.roll(p, -1, 0)
    wn = np.zeros(len(pts), np.int64)
    for (x0, y0), (x1, y1) in zip(p, q):
        is_left = (x1 - x0) * (pts[:, 1] - y0) - (pts[:, 0] - x0) * (y1 - y0)
        up = (y0 <= pts[:, 1]) & (y1 > pts[:, 1]) & (is_left > 0)
        dn = (y0 > pts[:, 1]) & (y1 <= pts[:, 1]) & (is_left < 0)
        wn += up.astype(np.int64) - dn.astype(np.int64)
    return wn


def dist_to_path(path, pts):
    p = np.asarray(path, np.float64)
    q = np.roll(p, -1, 0)
    d = np.full(len(pts), np.inf)
    for a, b in zip(p, q):
        ab = b - a
        t = np.clip(((pts - a) @ ab) / max(ab @ ab, 1e-30), 0, 1)
        d = np.minimum(d, np.sqrt((((a + t[:, None] * ab) - pts) ** 2).sum(1)))
    return d


@pytest.mark.parametrize("delta", [0.8, 3.0, 11.5, 40.0])
def test_unpinned_convex_offset_closed_form(delta):
    """Round-join offset of a convex polygon: area = A + P d + pi d^2, perimeter = P + 2 pi d (up to the arc tolerance 0.25 and
    integer rounding), every vertex within [d - 0.25 - 0.71, d + 0.71] of the source polygon."""
    from db_text_minimal_b200.postprocess import clipper_offset
    for poly in ([(10, 10), (110, 14), (104, 60), (8, 52)], [(0, 0), (50, 0), (50, 50), (0, 50)], [(5, 0), (45, 8), (60, 40), (30, 70), (-10, 35)]):
        res = clipper_offset(np.array(poly), delta)
        assert len(res) == 1
        A, P = poly_area_perimeter(poly)
        a, p = poly_area_perimeter(res[0])
        assert abs(a - (A + P * delta + np.pi * delta ** 2)) <= 0.02 * a + 1.5 * p * 0.5 / max(delta, 1) + 4
        assert abs(p - (P + 2 * np.pi * delta)) <= 0.02 * p + 4
        d = dist_to_path(poly, res[0].astype(np.float64))
        assert d.max() <= delta + 0.75 and d.min() >= delta - 0.25 - 0.75


def _nonconvex_cases():
    rng = np.random.RandomState(3)
    cases = {
        "L": [(0, 0), (60, 0), (60, 20), (20, 20), (20, 70), (0, 70)],
        "U": [(0, 0), (80, 0), (80, 60), (60, 60), (60, 15), (20, 15), (20, 60), (0, 60)],
        "arrow": [(0, 20), (50, 20), (50, 0), (90, 35), (50, 70), (50, 50), (0, 50)],
        "zigzag": [(0, 0), (30, 12), (60, 0), (90, 12), (120, 0), (120, 40), (90, 28), (60, 40), (30, 28), (0, 40)],
        "clockwise_L": [(0, 70), (20, 70), (20, 20), (60, 20), (60, 0), (0, 0)],
    }
    for s in range(6):      # star-shaped random polygons (simple by construction)
        k = rng.randint(6, 14)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(15, 60, k)
        cases[f"star{s}"] = [(int(100 + r * np.cos(a)), int(100 + r * np.sin(a))) for a, r in zip(ang, rad)]
    return cases


@pytest.mark.parametrize("name", sorted(_nonconvex_cases()))
@pytest.mark.parametrize("delta", [2.0, 7.5, 16.0])
def test_unpinned_nonconvex_offset_region_is_positive_winding_of_raw_path(name, delta):
    """Execute()'s union under pftPositive keeps exactly the points whose winding number w.r.t. the raw offset path is > 0.
    Checked on a point grid (points within 1 px of either boundary are skipped: intersection points are rounded to integers),
    and geometrically: the result contains the source polygon and stays within delta + 1 of it."""
    from db_text_minimal_b200.postprocess import clipper_offset, clipper_offset_raw
    poly = np.array(_nonconvex_cases()[name])
    raw = clipper_offset_raw(poly, delta)
    res = clipper_offset(poly, delta)
    assert len(raw) >= len(poly) and len(res) >= 1
    lo, hi = raw.min(0) - 3, raw.max(0) + 3
    gx, gy = np.meshgrid(np.arange(lo[0], hi[0], 1.37), np.arange(lo[1], hi[1], 1.37))
    pts = np.stack([gx.ravel() + 0.123, gy.ravel() + 0.456], 1)
    near = dist_to_path(raw, pts) < 1.0
    for r in res:
        near |= dist_to_path(r, pts) < 1.0
    want = winding_number(raw, pts) > 0
    got = np.zeros(len(pts), np.int64)
    for r in res:                     # outer polygons are counter-clockwise (+1), holes clockwise (-1)
        got += winding_number(r, pts)
    assert set(np.unique(got[~near])) <= {0, 1}
    assert np.array_equal(got[~near] > 0, want[~near]), (name, delta, int(((got > 0) != want)[~near].sum()))
    # contains the source polygon, never further than delta (+ rounding) from it
    src_ccw = poly if poly_area_perimeter(poly)[0] and winding_number(poly, poly.mean(0, keepdims=True).astype(np.float64))[0] >= 0 else poly[::-1]
    inside_src = winding_number(src_ccw, pts) != 0
    assert (got[inside_src & ~near] > 0).all()
    d = dist_to_path(poly, np.concatenate(res).astype(np.float64))
    assert d.max() <= delta + 1.0


def _blob_polygons(seed, size=1024):
    """approxPolyDP polygons of a synthetic probability map's components -- what polygons_from_bitmap hands to unclip
    (reference src/postprocess.py:62-86): many vertices, 1-3 px edges, concavities narrower than the offset."""
    ndimage = pytest.importorskip("scipy.ndimage")
    from oracle import db_oracle as O
    P = ((O.synth_prob_map(size, size, seed) - 0.45) * 8).clip(0, 1).astype(np.float32)
    bm = np.ascontiguousarray(O.binarize(P, 0.25).astype(np.uint8))
    lab, nl = ndimage.label(bm, structure=np.ones((3, 3), int))
    out = np.zeros((bm.size // 4, 2), np.int32)
    polys = []
    for sl, k in zip(ndimage.find_objects(lab), range(1, nl + 1)):
        ys, xs = np.nonzero(lab[sl] == k)
        i = np.lexsort((xs, ys))[0]
        n = lib().dbb_trace_contour(bm.ctypes.data, size, size, int(xs[i]) + sl[1].start, int(ys[i]) + sl[0].start, 0, out.ctypes.data, len(out))
        assert n > 0
        c = np.ascontiguousarray(out[:n])
        ap = np.zeros((n + 4, 2), np.int32)
        m = lib().dbb_approx_poly_dp(c.ctypes.data, n, -0.005, ap.ctypes.data, len(ap), None)
        if m >= 4:
            polys.append(ap[:m].copy())
    return polys


@pytest.mark.parametrize("seed", [100, 101, 102])
def test_unpinned_offset_region_on_realistic_polygons(seed):
    """The same positive-winding property at the unclip distance the detector uses, on the polygons it actually sees.  This is
    the case that broke a version which snapped intersection points to whole pixels (119 of 876 polygons lost their outer
    boundary): the raw path of a 4 px offset is made of 1-2 px arc pieces, so half-pixel snapping crosses neighbours."""
    from db_text_minimal_b200.postprocess import clipper_offset, clipper_offset_raw
    polys = _blob_polygons(seed)
    assert len(polys) > 40
    bad = []
    for idx, p in enumerate(polys):
        area, perim = poly_area_perimeter(p)
        d = abs(area) * 1.5 / perim
        raw, res = clipper_offset_raw(p, d), clipper_offset(p, d)
        lo, hi = raw.min(0) - 3, raw.max(0) + 3
        step = max(1.37, (hi - lo).max() / 60)
        gx, gy = np.meshgrid(np.arange(lo[0], hi[0], step), np.arange(lo[1], hi[1], step))
        pts = np.stack([gx.ravel() + 0.123, gy.ravel() + 0.456], 1)
        near = dist_to_path(raw, pts) < 1.0
        got = np.zeros(len(pts), np.int64)
        for r in res:
            near |= dist_to_path(r, pts) < 1.0
            got += winding_number(r, pts)
        want = winding_number(raw, pts) > 0
        if not (np.array_equal(got[~near] > 0, want[~near]) and set(np.unique(got[~near])) <= {0, 1}):
            bad.append((idx, len(p), round(d, 2), [len(r) for r in res]))
    assert not bad, bad[:5]


def test_unpinned_offset_with_hole_returns_two_paths():
    """A 'C' whose gap closes under the offset: the region has a hole, Execute returns 2 paths and the reference drops the
    candidate (src/postprocess.py:85-87 `if len(box) > 1: continue`)."""
    from db_text_minimal_b200.postprocess import clipper_offset
    c_shape = [(0, 0), (100, 0), (100, 40), (90, 40), (90, 10), (10, 10), (10, 90), (90, 90), (90, 46), (100, 46), (100, 100), (0, 100)]
    res = clipper_offset(np.array(c_shape), 5.0)
    assert len(res) == 2
    areas = [poly_area_perimeter(r)[0] for r in res]
    assert areas[0] > areas[1] > 0
    res_small = clipper_offset(np.array(c_shape), 1.0)       # gap (6 px) still open
    assert len(res_small) == 1


def test_unpinned_negative_delta_shrinks_and_can_vanish():
    """src/data_loaders.py:116-122 shrinks text polygons with a negative delta."""
    from db_text_minimal_b200.postprocess import clipper_offset
    sq = np.array([(0, 0), (40, 0), (40, 20), (0, 20)])
    res = clipper_offset(sq, -4.0)
    assert len(res) == 1
    a, _ = poly_area_perimeter(res[0])
    assert abs(a - 32 * 12) <= 2
    assert clipper_offset(sq, -12.0) == []
    dumbbell = np.array([(0, 0), (30, 0), (30, 12), (50, 12), (50, 0), (80, 0), (80, 30), (50, 30), (50, 18), (30, 18), (30, 30), (0, 30)])
    assert len(clipper_offset(dumbbell, -4.0)) == 2          # the 6-px bridge disappears, two pieces remain
