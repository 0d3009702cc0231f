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
