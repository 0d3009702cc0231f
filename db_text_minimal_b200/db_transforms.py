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
