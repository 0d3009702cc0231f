"""Synthetic inputs in the reference's value domains: images ~ N(0, 60^2) (mean-subtracted
0..255 pixels, src/utils.py:184-199) and the four ground-truth maps of src/data_loaders.py:87-149 (prob_map {0,1},
supervision_mask {0,1}, thresh_map [0.3, 0.7], text_area_map {0,1}) from random rotated rectangles."""
import math

import numpy as np
import torch


def gt_maps(n, h, w, seed=0):
    """(4, N, H, W) float32: prob_map, supervision_mask, thresh_map, text_area_map (src/train.py:163-166 order)."""
    import cv2
    rng = np.random.RandomState(seed)
    gts = np.zeros((4, n, h, w), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    scale = max(min(h, w) / 640.0, 0.1)
    for i in range(n):
        prob = np.zeros((h, w), np.uint8)
        area = np.zeros((h, w), np.uint8)
        thr = np.full((h, w), 0.3, np.float32)
        mask = np.ones((h, w), np.uint8)
        for _ in range(rng.randint(10, 31)):
            cx, cy = rng.uniform(0, w), rng.uniform(0, h)
            rw, rh = rng.uniform(24, 220) * scale + 8, rng.uniform(8, 48) * scale + 8
            ang = rng.uniform(-45, 45)
            d = rw * rh * (1 - 0.4 ** 2) / (2 * (rw + rh))          # shrink distance, data_loaders.py:116-117
            if min(rw, rh) - 2 * d < 1:
                continue
            inner = cv2.boxPoints(((cx, cy), (rw - 2 * d, rh - 2 * d), ang)).astype(np.int32)
            outer = cv2.boxPoints(((cx, cy), (rw + 2 * d, rh + 2 * d), ang)).astype(np.int32)
            cv2.fillPoly(prob, [inner], 1)
            tmp = np.zeros((h, w), np.uint8)
            cv2.fillPoly(tmp, [outer], 1)
            area |= tmp
            c, s = math.cos(math.radians(ang)), math.sin(math.radians(ang))
            u = (xx - cx) * c + (yy - cy) * s
            v = -(xx - cx) * s + (yy - cy) * c
            dist = np.abs(np.maximum(np.abs(u) - rw / 2, np.abs(v) - rh / 2))
            t = (0.3 + 0.4 * np.clip(1 - dist / d, 0, 1)).astype(np.float32)
            thr = np.where(tmp > 0, np.maximum(thr, t), thr)
        for _ in range(rng.randint(0, 3)):
            ign = cv2.boxPoints(((rng.uniform(0, w), rng.uniform(0, h)),
                                 (rng.uniform(20, 120) * scale + 4, rng.uniform(8, 40) * scale + 4), rng.uniform(-45, 45)))
            cv2.fillPoly(mask, [ign.astype(np.int32)], 0)
        gts[0, i], gts[1, i], gts[2, i], gts[3, i] = prob, mask, thr, area
    return gts


def images(n, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n, 3, h, w), generator=g) * 60.0


def prob_map(h, w, seed=0, sigma=6.0):
    """Blurred-noise blob map for the post-processing benchmark (about 90 candidates per 1024^2 map)."""
    import cv2
    rng = np.random.RandomState(seed)
    z = cv2.GaussianBlur(rng.uniform(0, 1, (h, w)).astype(np.float32), (0, 0), sigma)
    z = (z - z.min()) / max(float(z.max() - z.min()), 1e-12)
    return z.astype(np.float32)
