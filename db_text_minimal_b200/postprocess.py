"""SegDetectorRepresenter -- drop-in for the reference's src/postprocess.py with a GPU front.

Front half (per batch, on the device, csrc/ccl.cu): binarize, candidate extraction (the cv2.RETR_LIST contour set as
connected components: one candidate per 8-connected foreground component and per enclosed 4-connected background
region), float64 box score over the contour's fill set, score filter.  Only the per-candidate records (56 bytes each)
and, when boxes are requested, the 1-byte bitmap cross to the host -- not the two full float maps per image the
reference copies (src/postprocess.py:61-62,113-114).

Back half (host, survivors only): contour of the surviving component from the bitmap crop (cv2, same border follower
as the reference), get_mini_boxes, unclip, rescale -- src/postprocess.py:70-103,122-147,150-184.
``unclip`` needs Clipper (pyclipper 1.1.0.post3) which is neither in the reference tree nor in this image: when
pyclipper is importable it is used; otherwise a restatement of ClipperOffset's round-join arithmetic for CONVEX
input (box mode) is used and polygon mode raises.  That stage is unpinned (no pyclipper to generate goldens).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib


_PYCLIPPER = []


def _pyclipper():
    """pyclipper if it is installed, else None -- looked up once (a failing import costs ~0.3 ms per call)."""
    if not _PYCLIPPER:
        try:
            import pyclipper
            _PYCLIPPER.append(pyclipper)
        except ImportError:
            _PYCLIPPER.append(None)
    return _PYCLIPPER[0]


def _clipper_round(v):
    # Clipper's Round(): (val < 0) ? (cInt)(val - 0.5) : (cInt)(val + 0.5)
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def offset_convex_round(points, delta, arc_tolerance=0.25):
    """ClipperOffset(JT_ROUND, ET_CLOSEDPOLYGON).Execute(delta) for a CONVEX polygon, restated from Clipper 6.4.2
    (ClipperOffset::DoOffset / OffsetPoint / DoRound).  Integer coordinates in, integer coordinates out.
    For convex input the raw offset path is already simple, so the final union Clipper runs leaves the point set
    unchanged up to start vertex and collinear points -- irrelevant to the min-area rectangle taken next."""
    pts = [(int(p[0]), int(p[1])) for p in points]      # pyclipper truncates float input to cInt
    # orientation: Clipper offsets outward for positive area; reverse if needed
    area = 0.0
    n = len(pts)
    for i in range(n):
        x0, y0 = pts[i]; x1, y1 = pts[(i + 1) % n]
        area += (x0 + x1) * (y0 - y1)
    area = -area * 0.5
    if area < 0:
        pts = pts[::-1]
    # strip duplicate neighbours
    clean = [pts[0]]
    for p in pts[1:]:
        if p != clean[-1]:
            clean.append(p)
    if len(clean) > 1 and clean[0] == clean[-1]:
        clean.pop()
    pts = clean
    n = len(pts)
    if n < 3 or delta <= 0:
        return np.array(pts, dtype=np.int64).reshape(-1, 2)
    y = arc_tolerance if arc_tolerance > 0 else 0.25
    y = min(y, abs(delta) * 0.25) if y > abs(delta) * 0.25 else y
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin, m_cos = math.sin(2 * math.pi / steps), math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    normals = []
    for j in range(n):
        x0, y0 = pts[j]; x1, y1 = pts[(j + 1) % n]
        dx, dy = float(x1 - x0), float(y1 - y0)
        f = 1.0 / math.sqrt(dx * dx + dy * dy)
        normals.append((dy * f, -dx * f))
    out = []
    k = n - 1
    for j in range(n):
        sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
        cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
        if abs(sin_a * delta) < 1.0 and cos_a > 0:
            out.append((_clipper_round(pts[j][0] + normals[k][0] * delta), _clipper_round(pts[j][1] + normals[k][1] * delta)))
            k = j
            continue
        sin_a = max(-1.0, min(1.0, sin_a))
        if sin_a * delta < 0:      # concave vertex (does not occur for convex input)
            out.append((_clipper_round(pts[j][0] + normals[k][0] * delta), _clipper_round(pts[j][1] + normals[k][1] * delta)))
            out.append(pts[j])
            out.append((_clipper_round(pts[j][0] + normals[j][0] * delta), _clipper_round(pts[j][1] + normals[j][1] * delta)))
        else:                      # DoRound
            a = math.atan2(sin_a, cos_a)
            st = max(int(_clipper_round(steps_per_rad * abs(a))), 1)
            X, Y = normals[k]
            for _ in range(st):
                out.append((_clipper_round(pts[j][0] + X * delta), _clipper_round(pts[j][1] + Y * delta)))
                X2 = X
                X = X * m_cos - m_sin * Y
                Y = X2 * m_sin + Y * m_cos
            out.append((_clipper_round(pts[j][0] + normals[j][0] * delta), _clipper_round(pts[j][1] + normals[j][1] * delta)))
        k = j
    return np.array(out, dtype=np.int64).reshape(-1, 2)


class SegDetectorRepresenter():
    def __init__(self, thresh=0.3, box_thresh=0.7, max_candidates=1000, unclip_ratio=1.5):
        self.min_size = 3
        self.thresh = thresh
        self.box_thresh = box_thresh
        self.max_candidates = max_candidates
        self.unclip_ratio = unclip_ratio
        import os
        self.host_threads = min(16, os.cpu_count() or 1)     # host back half (survivors only): one image per worker

    # ------------------------------------------------------------------ GPU front
    def binarize(self, pred):
        """src/postprocess.py:51-52 -> torch.bool, same shape as pred ((N,H,W) or (H,W))."""
        _lib.require_cuda(pred)
        p = pred.detach().float().contiguous()
        shape = p.shape
        p3 = p.reshape(-1, shape[-2], shape[-1])
        out = torch.empty(p3.shape, dtype=torch.uint8, device=p.device)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().dbb_binarize(p3.data_ptr(), p3.shape[0], 1, shape[-2], shape[-1], float(self.thresh),
                                               out.data_ptr(), _lib.stream_ptr()), "dbb_binarize")
        return out.reshape(shape).bool()

    def front(self, pred, want_labels=False):
        """Device front on a (N, C, H, W) prediction (channel 0 = probability map, src/postprocess.py:33).
        Returns (bitmap uint8 (N,H,W) on device, labels int32 or None, candidate records (numpy structured), n_cands)."""
        _lib.require_cuda(pred)
        L = _lib.lib()
        p = pred.detach().float().contiguous()
        if p.dim() == 3:
            p = p[:, None]
        n, c, h, w = p.shape
        dev = p.device
        bitmap = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
        labels = torch.empty((n, h, w), dtype=torch.int32, device=dev) if want_labels else None
        csize = C.sizeof(_lib.DbbCandidate)
        cands = torch.empty((n, self.max_candidates, csize), dtype=torch.uint8, device=dev)
        ncand = torch.empty(n, dtype=torch.int32, device=dev)
        wsb = L.dbb_postprocess_workspace(n, h, w)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.dbb_binarize_ccl_score(p.data_ptr(), n, c, h, w, float(self.thresh), float(self.box_thresh),
                                                bitmap.data_ptr(), labels.data_ptr() if want_labels else None,
                                                cands.data_ptr(), ncand.data_ptr(), self.max_candidates, ws.data_ptr(), wsb,
                                                _lib.stream_ptr()), "dbb_binarize_ccl_score")
        nc = ncand.cpu().numpy()
        kmax = int(min(nc.max(), self.max_candidates)) if n else 0
        raw = cands[:, :kmax].cpu().numpy() if kmax else np.zeros((n, 0, csize), np.uint8)
        dt = np.dtype([("kind", "<i4"), ("first_y", "<i4"), ("first_x", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"),
                       ("y1", "<i4"), ("count", "<i4"), ("sum", "<f8"), ("keep", "<i4"), ("pad", "<i4")])
        assert dt.itemsize == csize
        rec = raw.reshape(n, kmax * csize).view(dt).reshape(n, kmax) if kmax else np.zeros((n, 0), dt)
        return bitmap, labels, rec, nc

    def candidates(self, pred):
        """Per image: list of dicts (kind, score, count, bbox, first, keep) in the reference's contour order."""
        _, _, rec, nc = self.front(pred)
        out = []
        for i in range(rec.shape[0]):
            k = int(min(nc[i], self.max_candidates))
            out.append([dict(kind="hole" if r["kind"] else "outer", score=float(r["sum"]) / int(r["count"]), count=int(r["count"]),
                             sum=float(r["sum"]), bbox=(int(r["x0"]), int(r["y0"]), int(r["x1"]), int(r["y1"])),
                             first=(int(r["first_y"]), int(r["first_x"])), keep=bool(r["keep"])) for r in rec[i, :k]])
        return out

    # ------------------------------------------------------------------ host back half (survivors only)
    @staticmethod
    def _contour_of(bitmap_np, r):
        """Border of one candidate, traced by OpenCV on the bitmap crop (same follower as the reference uses)."""
        import cv2
        h, w = bitmap_np.shape
        x0, y0, x1, y1 = max(int(r["x0"]) - 1, 0), max(int(r["y0"]) - 1, 0), min(int(r["x1"]) + 1, w - 1), min(int(r["y1"]) + 1, h - 1)
        crop = np.ascontiguousarray(bitmap_np[y0:y1 + 1, x0:x1 + 1])
        fy, fx = int(r["first_y"]) - y0, int(r["first_x"]) - x0
        mask = np.zeros((crop.shape[0] + 2, crop.shape[1] + 2), np.uint8)
        if r["kind"] == 0:      # component containing the first pixel, 8-connected
            cv2.floodFill(crop.copy(), mask, (fx, fy), 2, flags=8 | cv2.FLOODFILL_MASK_ONLY | (1 << 8))
            comp = mask[1:-1, 1:-1]
            cs, _ = cv2.findContours(comp * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            c = cs[0]
        else:                   # hole region containing the first pixel, 4-connected; its border lives on the parent
            inv = (1 - crop).astype(np.uint8)
            cv2.floodFill(inv.copy(), mask, (fx, fy), 2, flags=4 | cv2.FLOODFILL_MASK_ONLY | (1 << 8))
            hole = mask[1:-1, 1:-1]
            solid = (1 - hole).astype(np.uint8)          # everything but the hole is foreground -> one hole border
            cs, hier = cv2.findContours(solid * 255, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
            c = next(cs[i] for i in range(len(cs)) if hier[0][i][3] >= 0)
        return c + np.array([[[x0, y0]]], dtype=c.dtype)

    def unclip(self, box, unclip_ratio=1.5):
        """src/postprocess.py:150-156.  distance = area * ratio / perimeter (shapely Polygon.area / .length)."""
        pts = np.asarray(box, dtype=np.float64).reshape(-1, 2)
        nxt = np.concatenate([pts[1:], pts[:1]])            # (np.roll costs ~30 us per call on these 4-point arrays)
        area = 0.5 * abs(float((pts[:, 0] * nxt[:, 1] - pts[:, 1] * nxt[:, 0]).sum()))
        length = float(np.sqrt(((pts - nxt) ** 2).sum(1)).sum())
        distance = area * unclip_ratio / length
        pyclipper = _pyclipper()
        if pyclipper is not None:
            offset = pyclipper.PyclipperOffset()
            offset.AddPath(box, pyclipper.JT_ROUND, pyclipper.ET_CLOSEDPOLYGON)
            return np.array(offset.Execute(distance))
        import cv2
        hull = cv2.convexHull(pts.astype(np.float32)).reshape(-1, 2)
        if len(hull) != len(pts):
            raise _lib.DbbError("unclip of a non-convex polygon needs pyclipper (Clipper 6.4.2), which is not installed; "
                                "box mode (convex input) uses the built-in restatement")
        return offset_convex_round(pts, distance)[None]

    def get_mini_boxes(self, contour):
        """src/postprocess.py:158-184."""
        import cv2
        try:
            bounding_box = cv2.minAreaRect(contour)
            points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda x: x[0])
            i1, i4 = (0, 1) if points[1][1] > points[0][1] else (1, 0)
            i2, i3 = (2, 3) if points[3][1] > points[2][1] else (3, 2)
            return [points[i1], points[i2], points[i3], points[i4]], min(bounding_box[1])
        except Exception:
            return [], -1

    def box_score_fast(self, bitmap, _box):
        """src/postprocess.py:186-198 for callers that hold a contour: float64 mean of P over fillPoly(contour).
        (The batched path never calls this: scores come from the device front.)"""
        import cv2
        if torch.is_tensor(bitmap):
            bitmap = bitmap.detach().cpu().numpy()
        h, w = bitmap.shape[:2]
        box = np.array(_box).copy()
        xmin = int(np.clip(np.floor(box[:, 0].min()), 0, w - 1)); xmax = int(np.clip(np.ceil(box[:, 0].max()), 0, w - 1))
        ymin = int(np.clip(np.floor(box[:, 1].min()), 0, h - 1)); ymax = int(np.clip(np.ceil(box[:, 1].max()), 0, h - 1))
        mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
        box[:, 0] = box[:, 0] - xmin
        box[:, 1] = box[:, 1] - ymin
        cv2.fillPoly(mask, box.reshape(1, -1, 2).astype(np.int32), 1)
        return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]

    def _boxes(self, bitmap_np, rec, dest_width, dest_height):
        """src/postprocess.py:119-147 on the device-scored candidates."""
        height, width = bitmap_np.shape
        num = len(rec)
        boxes = np.zeros((num, 4, 2), dtype=np.int16)
        scores = np.zeros((num,), dtype=np.float32)
        for index, r in enumerate(rec):
            if not r["keep"]:           # cheap test first; the reference applies sside first, the kept set is the same
                continue
            contour = self._contour_of(bitmap_np, r).squeeze(1)
            points, sside = self.get_mini_boxes(contour)
            if sside < self.min_size:
                continue
            points = np.array(points)
            score = float(r["sum"]) / int(r["count"])
            box = self.unclip(points, unclip_ratio=self.unclip_ratio).reshape(-1, 1, 2)
            box, sside = self.get_mini_boxes(box.astype(np.float32) if box.dtype != np.int32 else box)
            if sside < self.min_size + 2:
                continue
            box = np.array(box)
            if not isinstance(dest_width, int):
                dest_width, dest_height = dest_width.item(), dest_height.item()
            box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
            box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
            boxes[index, :, :] = box.astype(np.int16)
            scores[index] = score
        return boxes, scores

    def _polygons(self, bitmap_np, rec, dest_width, dest_height):
        """src/postprocess.py:70-103."""
        import cv2
        height, width = bitmap_np.shape
        boxes, scores = [], []
        for r in rec:
            if not r["keep"]:
                continue
            contour = self._contour_of(bitmap_np, r)
            epsilon = 0.005 * cv2.arcLength(contour, True)
            approx = cv2.approxPolyDP(contour, epsilon, True)
            points = approx.reshape((-1, 2))
            if points.shape[0] < 4:
                continue
            score = float(r["sum"]) / int(r["count"])
            box = self.unclip(points, unclip_ratio=self.unclip_ratio)
            if len(box) > 1:
                continue
            box = np.asarray(box).reshape(-1, 2)
            _, sside = self.get_mini_boxes(box.reshape((-1, 1, 2)).astype(np.int32))
            if sside < self.min_size + 2:
                continue
            if not isinstance(dest_width, int):
                dest_width, dest_height = dest_width.item(), dest_height.item()
            box = box.astype(np.float64)
            box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
            box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
            boxes.append(box)
            scores.append(score)
        return boxes, scores

    def __call__(self, batch, pred, is_output_polygon=False):
        """src/postprocess.py:19-49: returns (boxes_batch, scores_batch)."""
        bitmap, _, rec, nc = self.front(pred)
        bm = bitmap.cpu().numpy()
        fn = self._polygons if is_output_polygon else self._boxes

        def one(i):
            height, width = batch['shape'][i]
            k = int(min(nc[i], self.max_candidates))
            return fn(bm[i], rec[i, :k], width, height)

        n = bm.shape[0]
        if n > 1 and self.host_threads > 1:      # images are independent; the OpenCV calls release the GIL
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(self.host_threads, n)) as pool:
                results = list(pool.map(one, range(n)))
        else:
            results = [one(i) for i in range(n)]
        return [r[0] for r in results], [r[1] for r in results]

    def boxes_from_bitmap(self, pred, _bitmap, dest_width, dest_height):
        """src/postprocess.py:106-148 for one (H, W) map; ``_bitmap`` is recomputed on the device from ``pred``."""
        assert len(_bitmap.shape) == 2
        bitmap, _, rec, nc = self.front(pred[None, None])
        return self._boxes(bitmap[0].cpu().numpy(), rec[0, :int(min(nc[0], self.max_candidates))], dest_width, dest_height)

    def polygons_from_bitmap(self, pred, _bitmap, dest_width, dest_height):
        assert len(_bitmap.shape) == 2
        bitmap, _, rec, nc = self.front(pred[None, None])
        return self._polygons(bitmap[0].cpu().numpy(), rec[0, :int(min(nc[0], self.max_candidates))], dest_width, dest_height)
