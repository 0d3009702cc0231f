"""Ground-truth border ("threshold") maps on the device -- the distance field of the reference's draw_thresh_map
(src/db_transforms.py:8-59), SURVEY.md section 8 f-4.

The reference draws one polygon at a time into numpy canvases inside a single-worker DataLoader.  ``thresh_maps`` takes all
text polygons of a batch and produces the (N, H, W) border map with one launch (csrc/gt_maps.cu, float64, bit-identical to
the numpy arithmetic).  The polygon dilation that precedes the distance field is Clipper's (pyclipper) in the reference; here
it is the C++ ClipperOffset restatement behind postprocess.clipper_offset (csrc/clipper_offset.cu; parity of that stage is
unpinned); the dilated polygon's fill (the ``mask`` canvas) stays a host cv2.fillPoly."""
import numpy as np
import torch

from . import _lib
from .postprocess import clipper_offset


def dilate_polygon(polygon, shrink_ratio=0.4):
    """src/db_transforms.py:13-21: returns (padded integer polygon, distance) or (None, 0) for a zero-area polygon."""
    p = np.asarray(polygon, dtype=np.float64)
    q = np.concatenate([p[1:], p[:1]])
    area = 0.5 * abs(float((p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0]).sum()))
    length = float(np.sqrt(((p - q) ** 2).sum(1)).sum())
    if area <= 0:
        return None, 0.0
    distance = area * (1 - np.power(shrink_ratio, 2)) / length
    res = clipper_offset(p, distance)          # padded_polygon = np.array(padding.Execute(distance)[0])
    if not res:
        return None, 0.0
    return res[0].astype(np.int64), float(distance)


def thresh_maps(polygons_per_image, height, width, shrink_ratio=0.4, device="cuda", padded=None):
    """polygons_per_image: list (one entry per image) of lists of (K, 2) polygons.  Returns (canvas (N, H, W) float32 on the
    device, list of per-image padded polygons for the host-side mask fill).  ``padded`` (same nesting, (polygon, distance)
    pairs) overrides the dilation, e.g. with polygons dilated elsewhere."""
    n = len(polygons_per_image)
    pts, start, image, bbox, dist, padded_out = [], [0], [], [], [], []
    max_pts = 0
    for i, polys in enumerate(polygons_per_image):
        padded_out.append([])
        for k, poly in enumerate(polys):
            poly = np.asarray(poly)
            assert poly.ndim == 2 and poly.shape[1] == 2
            pp, d = padded[i][k] if padded is not None else dilate_polygon(poly, shrink_ratio)
            padded_out[i].append(pp)
            if pp is None:
                continue
            pp = np.asarray(pp)
            pts.append(poly.astype(np.float64))
            start.append(start[-1] + len(poly))
            image.append(i)
            bbox.append([pp[:, 0].min(), pp[:, 1].min(), pp[:, 0].max(), pp[:, 1].max()])
            dist.append(d)
            max_pts = max(max_pts, len(poly))
    dev = torch.device(device)
    canvas = torch.zeros((n, height, width), dtype=torch.float32, device=dev)
    if not image:
        return canvas, padded_out
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_pts, d_start, d_img = t(np.concatenate(pts), np.float64), t(start, np.int32), t(image, np.int32)
    d_bbox, d_dist = t(bbox, np.int64), t(dist, np.float64)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbb_thresh_map(canvas.data_ptr(), n, height, width, d_pts.data_ptr(), d_start.data_ptr(),
                                             d_img.data_ptr(), d_bbox.data_ptr(), d_dist.data_ptr(), len(image), max_pts,
                                             _lib.stream_ptr()), "dbb_thresh_map")
    return canvas, padded_out


def draw_thresh_map(polygon, canvas, mask, shrink_ratio=0.4):
    """Signature of the reference (numpy canvases, in place) for callers that draw one polygon at a time; the distance
    field comes from the device.  Prefer ``thresh_maps`` for a whole batch."""
    import cv2
    padded, distance = dilate_polygon(polygon, shrink_ratio)
    if padded is None:
        return
    cv2.fillPoly(mask, [padded.astype(np.int32)], 1.0)
    dev_canvas, _ = thresh_maps([[polygon]], canvas.shape[0], canvas.shape[1], shrink_ratio, padded=[[(padded, distance)]])
    np.fmax(canvas, dev_canvas[0].cpu().numpy(), out=canvas)


def fill_polygons(maps, polygons, planes, values):
    """cv2.fillPoly(maps[plane], [polygon.astype(int32)], value) for a list of polygons, on the device, bit-exact with OpenCV
    (csrc/poly_fill.cu).  maps: (P, H, W) float32 CUDA tensor, filled in place."""
    if not polygons:
        return maps
    _lib.require_cuda(maps)
    assert maps.dtype == torch.float32 and maps.dim() == 3 and maps.is_contiguous()
    verts = [np.asarray(p).astype(np.int32).reshape(-1, 2) for p in polygons]      # (poly.astype(np.int32): truncation)
    if max(len(v) for v in verts) > 512:
        raise _lib.DbbError("fill_polygons: more than 512 vertices in one polygon")
    start = np.zeros(len(verts) + 1, np.int32)
    start[1:] = np.cumsum([len(v) for v in verts])
    dev = maps.device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_v, d_s, d_p, d_val = t(np.concatenate(verts), np.int32), t(start, np.int32), t(planes, np.int32), t(values, np.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbb_fill_polygons(d_v.data_ptr(), d_s.data_ptr(), d_p.data_ptr(), d_val.data_ptr(), len(verts),
                                                maps.data_ptr(), maps.shape[1], maps.shape[2], _lib.stream_ptr()), "dbb_fill_polygons")
    return maps


def gt_maps(anns_per_image, image_size, shrink_ratio=0.4, thresh_min=0.3, thresh_max=0.7, min_text_size=8,
            ignore_tags=("*", "###"), device="cuda"):
    """The four ground-truth maps of a batch, src/data_loaders.py:86-149 (after augmentation / resize), built on the device:
    shrink map `prob_map`, `supervision_mask`, border `thresh_map` (scaled to [thresh_min, thresh_max]) and `text_area_map`
    (the reference's thresh_mask).  anns_per_image: per image a list of {'poly': (K, 2) array, 'text': str}.
    Returns a dict of (N, S, S) float32 CUDA tensors plus 'ignore_tags' (per image list of bools, data_loaders.py:102-137).

    Host side: per polygon the area / perimeter, the two Clipper offsets (C++, csrc/clipper_offset.cu -- unpinned) and the
    ignore logic; device side: every cv2.fillPoly (csrc/poly_fill.cu, bit-exact with OpenCV) and the border distance field
    (csrc/gt_maps.cu, bit-exact with the numpy arithmetic).  The reference's shapely validity tests
    (`Polygon(...).buffer(0).is_valid`, data_loaders.py:85,128) are taken as true: shapely is not a dependency."""
    from .postprocess import clipper_offset
    n, S = len(anns_per_image), int(image_size)
    dev = torch.device(device)
    # planes: [0, N) prob_map, [N, 2N) supervision_mask, [2N, 3N) text_area_map
    maps = torch.zeros((3 * n, S, S), dtype=torch.float32, device=dev)
    maps[n:2 * n] = 1.0
    fills, planes, values, thresh_polys, padded, ignore = [], [], [], [[] for _ in range(n)], [[] for _ in range(n)], [[] for _ in range(n)]
    for i, anns in enumerate(anns_per_image):
        for ann in anns:
            poly = np.array(ann["poly"])
            height = max(poly[:, 1]) - min(poly[:, 1])
            width = max(poly[:, 0]) - min(poly[:, 0])
            p = poly.astype(np.float64)
            q = np.concatenate([p[1:], p[:1]])
            area = 0.5 * abs(float((p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0]).sum()))
            length = float(np.sqrt(((p - q) ** 2).sum(1)).sum())

            def ignore_it():
                ignore[i].append(True)
                fills.append(poly); planes.append(n + i); values.append(0.0)

            if area < 1 or min(height, width) < min_text_size or ann.get("text") in ignore_tags:
                ignore_it()
                continue
            distance = area * (1 - np.power(shrink_ratio, 2)) / length
            shrinked = clipper_offset(poly, -distance)
            if len(shrinked) == 0:
                ignore_it()
                continue
            sh = np.array(shrinked[0]).reshape(-1, 2)
            if sh.shape[0] > 2:
                ignore[i].append(False)
                fills.append(sh); planes.append(i); values.append(1.0)
            else:
                ignore_it()
                continue
            # draw_thresh_map (src/db_transforms.py:8-59): dilated polygon -> text-area fill + border distance field
            pp, d = dilate_polygon(poly, shrink_ratio)
            if pp is None:
                continue
            fills.append(pp); planes.append(2 * n + i); values.append(1.0)
            thresh_polys[i].append(poly)
            padded[i].append((pp, d))
    fill_polygons(maps, fills, planes, values)
    canvas, _ = thresh_maps(thresh_polys, S, S, shrink_ratio, device=dev, padded=padded)
    thresh = canvas * (thresh_max - thresh_min) + thresh_min
    return {"prob_map": maps[:n], "supervision_mask": maps[n:2 * n], "thresh_map": thresh, "text_area_map": maps[2 * n:],
            "ignore_tags": ignore}
