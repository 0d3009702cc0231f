"""SegDetectorRepresenter -- drop-in for the reference's src/postprocess.py with a GPU front.

Front half (per batch, on the device, csrc/ccl.cu): binarize, candidate extraction (the cv2.RETR_LIST contour set as
connected components: one candidate per 8-connected foreground component and per enclosed 4-connected background
region), float64 box score over the contour's fill set, score filter.  Only the per-candidate records (56 bytes each)
and, when boxes are requested, the 1-byte bitmap cross to the host -- not the two full float maps per image the
reference copies (src/postprocess.py:61-62,113-114).

Back half.  Box mode (src/postprocess.py:106-148): the device emits the border points of the kept candidates (run end
pixels: their convex hull is the contour's), and C++ (csrc/post_geom.cu, csrc/clipper_offset.cu) does get_mini_boxes,
unclip, the second get_mini_boxes and the rescale for the whole batch -- no bitmap crosses to the host, no Python per box.
Polygon mode (src/postprocess.py:54-104) needs the ordered contour for cv2.approxPolyDP: the packed bitmap (1 bit / pixel)
comes back from the device and C++ traces the border of every kept candidate as cv2.findContours(CHAIN_APPROX_SIMPLE) does
(Suzuki-Abe border following), then cv2.arcLength / cv2.approxPolyDP restated, the offset, the filters and the rescale.
``unclip`` is Clipper (pyclipper 1.1.0.post3), which is neither in the reference tree nor in this image; it is restated in
csrc/clipper_offset.cu for polygons of any shape.  That stage is unpinned (no pyclipper to generate goldens).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def _to_cint(points):
    """pyclipper hands Clipper 64-bit integers: float coordinates are truncated toward zero."""
    return np.ascontiguousarray(np.trunc(np.asarray(points, dtype=np.float64).reshape(-1, 2)).astype(np.int64))


def clipper_offset(points, delta, arc_tolerance=0.25):
    """pyclipper.PyclipperOffset().AddPath(points, JT_ROUND, ET_CLOSEDPOLYGON); .Execute(delta) -> list of (k, 2) int64
    arrays (one per result polygon; more than one when the offset region has holes or falls apart).  Runs in C++
    (csrc/clipper_offset.cu: a restatement of Clipper 6.4.2's ClipperOffset -- parity unpinned, pyclipper is neither in
    the reference tree nor installable here)."""
    L = _lib.lib()
    path = _to_cint(points)
    cap = max(256, 64 * len(path) + int(8 * abs(delta)) + 64)
    while True:
        out = np.empty((cap, 2), np.int64)
        counts = np.zeros(64, np.int32)
        rc = L.dbb_clipper_offset(path.ctypes.data, len(path), float(delta), float(arc_tolerance), out.ctypes.data, cap,
                                  counts.ctypes.data, len(counts))
        if rc >= 0:
            break
        if cap > (1 << 22):
            _lib.check(rc, "dbb_clipper_offset")
        cap *= 4
    res, o = [], 0
    for i in range(rc):
        res.append(out[o:o + counts[i]].copy())
        o += counts[i]
    return res


def clipper_offset_raw(points, delta, arc_tolerance=0.25):
    """The raw offset path (before the union that removes its self-intersections); for tests."""
    L = _lib.lib()
    path = _to_cint(points)
    cap = max(256, 64 * len(path) + int(8 * abs(delta)) + 64)
    out = np.empty((cap, 2), np.int64)
    rc = L.dbb_clipper_offset_raw(path.ctypes.data, len(path), float(delta), float(arc_tolerance), out.ctypes.data, cap)
    if rc < 0:
        _lib.check(rc, "dbb_clipper_offset_raw")
    return out[:rc].copy()


def offset_convex_round(points, delta, arc_tolerance=0.25):
    """Offset polygon of a CONVEX input as one (k, 2) int64 array (kept for the GT-map code and the golden generator)."""
    res = clipper_offset(points, delta, arc_tolerance)
    return res[0] if res else np.zeros((0, 2), np.int64)


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

    def _front_device(self, pred, want_labels=False):
        """Launches the device front; everything stays on the device.  Returns a dict of the tensors involved."""
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
        return dict(p=p, n=n, h=h, w=w, dev=dev, bitmap=bitmap, labels=labels, cands=cands, ncand=ncand, ws=ws, wsb=wsb, csize=csize)

    def front(self, pred, want_labels=False):
        """Device front on a (N, C, H, W) prediction (channel 0 = probability map, src/postprocess.py:33).
        Returns (bitmap uint8 (N,H,W) on device, labels int32 or None, candidate records (numpy structured), n_cands)."""
        f = self._front_device(pred, want_labels)
        n, csize, bitmap, labels, cands, ncand = f["n"], f["csize"], f["bitmap"], f["labels"], f["cands"], f["ncand"]
        nc = ncand.cpu().numpy()
        kmax = int(min(nc.max(), self.max_candidates)) if n else 0
        raw = cands[:, :kmax].cpu().numpy() if kmax else np.zeros((n, 0, csize), np.uint8)
        dt = np.dtype([("kind", "<i4"), ("first_y", "<i4"), ("first_x", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"),
                       ("y1", "<i4"), ("count", "<i4"), ("sum", "<f8"), ("keep", "<i4"), ("pad", "<i4")])
        assert dt.itemsize == csize
        rec = raw.reshape(n, kmax * csize).view(dt).reshape(n, kmax) if kmax else np.zeros((n, 0), dt)
        return bitmap, labels, rec, nc

    # ------------------------------------------------------------------ box mode, whole batch: device front + C++ back half
    def boxes_batch(self, pred, dest_wh, debug=False):
        """src/postprocess.py:106-148 for a whole batch: the device front (binarize, candidates, scores, score filter), the
        border points of the kept candidates (dbb_ccl_border_points) and the C++ back half (dbb_boxes_from_border_points:
        get_mini_boxes, unclip, get_mini_boxes, rescale).  dest_wh: (N, 2) (dest_width, dest_height).
        Returns (boxes (N, K, 4, 2) int16, scores (N, K) float32, n_cands (N,)) with K = the largest candidate count of
        the batch (capped at max_candidates); rows of dropped candidates are zero, as in the reference."""
        L = _lib.lib()
        f = self._front_device(pred)
        n, h, w, dev = f["n"], f["h"], f["w"], f["dev"]
        cap = max(8192, (h * w) // 16)
        npts_dev = torch.empty(n, dtype=torch.int32, device=dev)
        while True:
            pts_dev = torch.empty((n, cap, 2), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.dbb_ccl_border_points(f["ws"].data_ptr(), f["wsb"], n, h, w, pts_dev.data_ptr(), npts_dev.data_ptr(), cap,
                                                   _lib.stream_ptr()), "dbb_ccl_border_points")
            counts = torch.stack([f["ncand"], npts_dev]).cpu().numpy()       # one synchronising copy
            nc, npts = counts[0], counts[1]
            if int(npts.max(initial=0)) <= cap:
                break
            cap = int(npts.max())                                             # rare: rerun with the exact size
        k = int(min(max(int(nc.max(initial=0)), 1), self.max_candidates))
        pmax = max(int(npts.max(initial=0)), 1)
        # one pinned staging buffer for both device->host copies (pageable copies run at a fraction of the PCIe rate)
        nb_c, nb_p = n * k * f["csize"], n * pmax * 8
        if getattr(self, "_stage", None) is None or self._stage.numel() < nb_c + nb_p:      # (cudaHostAlloc costs milliseconds)
            self._stage = torch.empty(int(1.25 * (nb_c + nb_p)), dtype=torch.uint8, pin_memory=True)
        stage = self._stage[:nb_c + nb_p]
        stage[:nb_c].view(n, k, f["csize"]).copy_(f["cands"][:, :k], non_blocking=True)
        stage[nb_c:].view(torch.int32).view(n, pmax, 2).copy_(pts_dev[:, :pmax], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host = stage.numpy()
        cands, pts = host[:nb_c], host[nb_c:nb_c + nb_p]
        dest = np.ascontiguousarray(np.asarray(dest_wh, dtype=np.int32).reshape(n, 2))
        boxes = np.empty((n, k, 4, 2), np.int16)
        scores = np.empty((n, k), np.float32)
        sside = np.zeros((n, k), np.float32) if debug else None
        mini = np.zeros((n, k, 4, 2), np.float32) if debug else None
        nc32, np32 = np.ascontiguousarray(nc.astype(np.int32)), np.ascontiguousarray(npts.astype(np.int32))
        _lib.check(L.dbb_boxes_from_border_points(cands.ctypes.data, nc32.ctypes.data, pts.ctypes.data, np32.ctypes.data, n, k, pmax,
                                                  h, w, dest.ctypes.data, float(self.unclip_ratio), int(self.min_size),
                                                  boxes.ctypes.data, scores.ctypes.data,
                                                  sside.ctypes.data if debug else None, mini.ctypes.data if debug else None,
                                                  int(self.host_threads)), "dbb_boxes_from_border_points")
        if debug:
            return boxes, scores, nc, sside, mini, cands
        return boxes, scores, nc

    # ------------------------------------------------------------------ polygon mode, whole batch: device front + C++ back half
    def polygons_batch(self, pred, dest_wh):
        """src/postprocess.py:54-104 for a whole batch.  The device front supplies the candidates, their scores and the packed
        bitmap (1 bit / pixel); C++ (dbb_polygons_from_bitmap) traces the border of every kept candidate exactly as
        cv2.findContours(CHAIN_APPROX_SIMPLE) does, then cv2.arcLength / cv2.approxPolyDP (restated, bit-exact on the test
        contours), unclip, the filters and the rescale.  Returns (list per image of (K_i, 2) int64 arrays, list per image of
        float scores) in the reference's contour order."""
        L = _lib.lib()
        f = self._front_device(pred)
        n, h, w, dev = f["n"], f["h"], f["w"], f["dev"]
        wq = (w + 31) // 32
        nc = f["ncand"].cpu().numpy()
        k = int(min(max(int(nc.max(initial=0)), 1), self.max_candidates))
        nb_c, nb_b = n * k * f["csize"], n * h * wq * 4
        if getattr(self, "_stage", None) is None or self._stage.numel() < nb_c + nb_b:
            self._stage = torch.empty(int(1.25 * (nb_c + nb_b)), dtype=torch.uint8, pin_memory=True)
        stage = self._stage[:nb_c + nb_b]
        stage[:nb_c].view(n, k, f["csize"]).copy_(f["cands"][:, :k], non_blocking=True)
        stage[nb_c:].copy_(f["ws"][:nb_b], non_blocking=True)              # the packed bitmap sits at the start of the workspace
        torch.cuda.current_stream(dev).synchronize()
        host = stage.numpy()
        cands, bits = host[:nb_c], host[nb_c:nb_c + nb_b]
        dest = np.ascontiguousarray(np.asarray(dest_wh, dtype=np.int32).reshape(n, 2))
        nc32 = np.ascontiguousarray(nc.astype(np.int32))
        cap = 256 * max(16, min(k, 256))
        while True:
            counts = np.empty((n, k), np.int32)
            pts = np.empty((n, cap, 2), np.int32)
            scores = np.empty((n, k), np.float64)
            totals = np.empty(n, np.int32)
            _lib.check(L.dbb_polygons_from_bitmap(bits.ctypes.data, cands.ctypes.data, nc32.ctypes.data, n, k, h, w, dest.ctypes.data,
                                                  float(self.unclip_ratio), int(self.min_size), counts.ctypes.data, pts.ctypes.data, cap,
                                                  scores.ctypes.data, totals.ctypes.data, int(self.host_threads)), "dbb_polygons_from_bitmap")
            if int(totals.max(initial=0)) <= cap:
                break
            cap = int(totals.max())
        boxes_batch, scores_batch = [], []
        for i in range(n):
            kept = np.nonzero(counts[i])[0]
            offs = np.concatenate([[0], np.cumsum(counts[i][kept])])
            boxes_batch.append([pts[i, offs[j]:offs[j + 1]].astype(np.int64) for j in range(len(kept))])
            scores_batch.append([float(v) for v in scores[i][kept]])
        return boxes_batch, scores_batch

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

    # ------------------------------------------------------------------ the reference's helper methods (single contour / polygon)
    def unclip(self, box, unclip_ratio=1.5):
        """src/postprocess.py:150-156.  distance = area * ratio / perimeter (shapely Polygon.area / .length); returns the list
        of result polygons (the reference's np.array(offset.Execute(distance)); callers test len(...) > 1 and reshape)."""
        pts = np.asarray(box, dtype=np.float64).reshape(-1, 2)
        nxt = np.concatenate([pts[1:], pts[:1]])
        area = 0.5 * abs(float((pts[:, 0] * nxt[:, 1] - pts[:, 1] * nxt[:, 0]).sum()))
        length = float(np.sqrt(((pts - nxt) ** 2).sum(1)).sum())
        distance = area * unclip_ratio / length
        return clipper_offset(box, distance)

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

    def __call__(self, batch, pred, is_output_polygon=False):
        """src/postprocess.py:19-49: returns (boxes_batch, scores_batch)."""
        if not is_output_polygon:        # box mode: nothing but candidate records and border points leaves the device
            n = pred.shape[0]
            dest = [(int(batch['shape'][i][1]), int(batch['shape'][i][0])) for i in range(n)]
            boxes, scores, nc = self.boxes_batch(pred, dest)
            ks = [int(min(nc[i], self.max_candidates)) for i in range(n)]
            return [boxes[i, :ks[i]] for i in range(n)], [scores[i, :ks[i]] for i in range(n)]
        n = pred.shape[0]
        dest = [(int(batch['shape'][i][1]), int(batch['shape'][i][0])) for i in range(n)]
        return self.polygons_batch(pred, dest)

    def boxes_from_bitmap(self, pred, _bitmap, dest_width, dest_height):
        """src/postprocess.py:106-148 for one (H, W) map; ``_bitmap`` is recomputed on the device from ``pred``."""
        assert len(_bitmap.shape) == 2
        if not isinstance(dest_width, int):
            dest_width, dest_height = dest_width.item(), dest_height.item()
        boxes, scores, nc = self.boxes_batch(pred[None, None], [(int(dest_width), int(dest_height))])
        k = int(min(nc[0], self.max_candidates))
        return boxes[0, :k], scores[0, :k]

    def polygons_from_bitmap(self, pred, _bitmap, dest_width, dest_height):
        assert len(_bitmap.shape) == 2
        if not isinstance(dest_width, int):
            dest_width, dest_height = dest_width.item(), dest_height.item()
        boxes, scores = self.polygons_batch(pred[None, None], [(int(dest_width), int(dest_height))])
        return boxes[0], scores[0]
