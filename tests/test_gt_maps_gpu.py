"""f-4: border-map distance field on the device vs the golden canvases produced by the reference's own draw_thresh_map
(oracle/make_golden.py: make_thresh_map_cases) and vs the oracle on random quads.  float64 arithmetic: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import db_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "thresh_map_cases.npz")


def test_thresh_map_matches_reference_golden():
    from db_text_minimal_b200.db_transforms import thresh_maps
    d = np.load(GOLD)
    names = [str(n) for n in d["names"]]
    H, W = d[names[0] + ":canvas_after"].shape
    polys, padded = [], []
    for n in names:                                            # one image per prefix of the polygon list: cumulative canvases
        polys.append([d[m + ":poly"] for m in names[:names.index(n) + 1]])
        padded.append([(np.array([[d[m + ":bbox"][0], d[m + ":bbox"][1]], [d[m + ":bbox"][2], d[m + ":bbox"][3]]]),
                        float(d[m + ":distance"][0])) for m in names[:names.index(n) + 1]])
    canvas, _ = thresh_maps(polys, H, W, padded=padded)
    got = canvas.cpu().numpy()
    for i, n in enumerate(names):
        assert np.array_equal(got[i], d[n + ":canvas_after"]), n      # float64 distance field: bit-exact


def test_thresh_map_random_quads_vs_oracle():
    from db_text_minimal_b200.db_transforms import thresh_maps, dilate_polygon
    rng = np.random.RandomState(4)
    H, W = 200, 240
    polys = []
    for _ in range(3):
        img = []
        for _ in range(12):
            cx, cy = rng.uniform(-10, W + 10), rng.uniform(-10, H + 10)
            w, h, a = rng.uniform(20, 90), rng.uniform(8, 30), rng.uniform(-0.6, 0.6)
            c, s = np.cos(a), np.sin(a)
            q = np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]])
            img.append(np.round(q @ np.array([[c, s], [-s, c]]) + [cx, cy]).astype(np.int64))
        polys.append(img)
    canvas, padded = thresh_maps(polys, H, W)
    got = canvas.cpu().numpy()
    for i, img in enumerate(polys):
        want = np.zeros((H, W), np.float32)
        for poly, pp in zip(img, padded[i]):
            _, dist = dilate_polygon(poly)
            bb = (pp[:, 0].min(), pp[:, 1].min(), pp[:, 0].max(), pp[:, 1].max())
            if bb[2] < 0 or bb[3] < 0 or bb[0] > W - 1 or bb[1] > H - 1:
                continue
            O.thresh_map_accumulate(want, poly, bb, dist)
        assert np.array_equal(got[i], want)
    assert float(got.max()) == 1.0 and (got > 0).mean() > 0.05


def _rand_polys(rng, count, size):
    import cv2
    polys = []
    for t in range(count):
        kind = t % 3
        if kind == 0:
            k = rng.randint(3, 10); ang = np.sort(rng.uniform(0, 2 * np.pi, k)); r = rng.uniform(3, 0.7 * size, k)
            c = rng.uniform(0, size, 2)
            polys.append(np.array([(int(c[0] + a * np.cos(q)), int(c[1] + a * np.sin(q))) for a, q in zip(r, ang)]))
        elif kind == 1:
            polys.append(rng.randint(-size // 3, size + size // 3, (rng.randint(3, 8), 2)))
        else:
            polys.append(cv2.boxPoints(((rng.uniform(-5, size + 5), rng.uniform(-5, size + 5)), (rng.uniform(2, size), rng.uniform(2, size / 2)),
                                        rng.uniform(-90, 90))).astype(np.int32))
    return polys


@pytest.mark.parametrize("size", [64, 160])
def test_fill_polygons_is_bit_exact_with_cv2_fillpoly(size):
    """a constant per canvas: every polygon alone on its own canvas, 1,200 random polygons (most reach outside the canvas:
    OpenCV clips edges before rasterising them, which the kernel reproduces)."""
    import cv2
    from db_text_minimal_b200.db_transforms import fill_polygons
    rng = np.random.RandomState(size)
    polys = _rand_polys(rng, 600, size)
    maps = torch.zeros((len(polys), size, size), dtype=torch.float32, device="cuda")
    fill_polygons(maps, polys, list(range(len(polys))), [1.0] * len(polys))
    got = maps.cpu().numpy()
    outside = 0
    for i, p in enumerate(polys):
        ref = np.zeros((size, size), np.float32)
        cv2.fillPoly(ref, [p.astype(np.int32)], 1.0)
        assert np.array_equal(got[i], ref), (i, p.tolist(), int((got[i] != ref).sum()))
        outside += bool((p < 0).any() or (p >= size).any())
    assert outside > 200


def _synthetic_annotations(rng, size, count):
    import cv2
    anns = []
    for t in range(count):
        cx, cy = rng.uniform(0, size, 2)
        w, h = rng.uniform(6, size / 2), rng.uniform(4, size / 6)
        box = cv2.boxPoints(((cx, cy), (w, h), rng.uniform(-60, 60)))
        if t % 4 == 3:          # a six-point polygon (curved text annotations have 6-14 points)
            mid = (box[0] + box[1]) / 2 + rng.uniform(-3, 3, 2)
            mid2 = (box[2] + box[3]) / 2 + rng.uniform(-3, 3, 2)
            box = np.array([box[0], mid, box[1], box[2], mid2, box[3]])
        anns.append({"poly": np.round(box).astype(np.int64), "text": "###" if t % 7 == 5 else "word"})
    return anns


def test_gt_maps_batch_matches_the_reference_loader_arithmetic_unpinned_offset():
    """f-4 end to end: the four maps of a batch against src/data_loaders.py:86-149 restated with OpenCV / numpy
    (oracle.gt_maps_reference).  Bit-exact; the Clipper offset on both sides is the product's restatement (UNPINNED)."""
    from db_text_minimal_b200.db_transforms import gt_maps
    from db_text_minimal_b200.postprocess import clipper_offset
    rng = np.random.RandomState(11)
    size = 256
    batch = [_synthetic_annotations(rng, size, rng.randint(4, 14)) for _ in range(5)]
    got = gt_maps(batch, size)
    kinds = set()
    for i, anns in enumerate(batch):
        gt, mask, thr, tmask, tags = O.gt_maps_reference(anns, size, clipper_offset)
        assert tags == got["ignore_tags"][i]
        assert np.array_equal(got["prob_map"][i].cpu().numpy(), gt)
        assert np.array_equal(got["supervision_mask"][i].cpu().numpy(), mask)
        assert np.array_equal(got["text_area_map"][i].cpu().numpy(), tmask)
        assert np.array_equal(got["thresh_map"][i].cpu().numpy(), thr), float(np.abs(got["thresh_map"][i].cpu().numpy() - thr).max())
        kinds |= set(tags)
        assert gt.sum() > 0 and (thr > 0.3).sum() > 0
    assert kinds == {True, False}
