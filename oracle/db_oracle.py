"""CPU restatement of the DB_text_minimal hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  The model part is written with torch CPU functional ops so that
autograd supplies the reference gradients; the loss and the post-processing front are
numpy (float64 accumulation) with closed-form gradients.

Nothing here is used by the product path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default, src/modules/basic.py:34
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# parameters: same 211-key state_dict layout as the reference
# --------------------------------------------------------------------------------------

def _bn_keys(prefix: str, c: int, weight=1.0, bias=0.0) -> Dict[str, torch.Tensor]:
    return {
        prefix + ".weight": torch.full((c,), float(weight)),
        prefix + ".bias": torch.full((c,), float(bias)),
        prefix + ".running_mean": torch.zeros(c),
        prefix + ".running_var": torch.ones(c),
        prefix + ".num_batches_tracked": torch.zeros((), dtype=torch.long),
    }


def init_params(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random init following the reference's initialisers.

    backbone: src/modules/resnet.py:197-203 (normal(0, sqrt(2/(k*k*cout))), BN 1/0);
    FPN: PyTorch Conv2d defaults (kaiming_uniform(a=sqrt(5)), bias U(+-1/sqrt(fan_in)));
    head: src/modules/segmentation_head.py:47-53 (kaiming_normal, BN 1 / 1e-4).
    The draw order differs from the reference's, so values differ from
    ``DBTextModel()`` under the same seed; the key set and shapes are identical.
    """
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}

    def rn(shape, std):
        return torch.randn(shape, generator=g) * std

    def conv_resnet(name, cout, cin, k):
        p[name + ".weight"] = rn((cout, cin, k, k), math.sqrt(2.0 / (k * k * cout)))

    conv_resnet("backbone.conv1", 64, 3, 7)
    p.update(_bn_keys("backbone.bn1", 64))
    inpl = 64
    for li, planes in enumerate([64, 128, 256, 512], start=1):
        for bi in range(2):
            pre = f"backbone.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            conv_resnet(pre + ".conv1", planes, inpl, 3)
            p.update(_bn_keys(pre + ".bn1", planes))
            conv_resnet(pre + ".conv2", planes, planes, 3)
            p.update(_bn_keys(pre + ".bn2", planes))
            if stride != 1 or inpl != planes:
                conv_resnet(pre + ".downsample.0", planes, inpl, 1)
                p.update(_bn_keys(pre + ".downsample.1", planes))
            inpl = planes
    # unused-in-forward members kept for state_dict parity (src/modules/resnet.py:192-195)
    p["backbone.fc.weight"] = rn((1000, 512), 0.01)
    p["backbone.fc.bias"] = torch.zeros(1000)
    conv_resnet("backbone.smooth", 256, 2048, 1)
    p["backbone.smooth.bias"] = torch.zeros(256)

    def conv_default(name, cout, cin, k, bias=True):
        fan_in = cin * k * k
        bound = 1.0 / math.sqrt(fan_in)
        p[name + ".weight"] = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) * bound
        if bias:
            p[name + ".bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound

    for nm, cin in (("c2", 64), ("c3", 128), ("c4", 256), ("c5", 512)):
        conv_default(f"segmentation_body.reduce_conv_{nm}.conv", 64, cin, 1)
        p.update(_bn_keys(f"segmentation_body.reduce_conv_{nm}.bn", 64))
    for nm in ("p4", "p3", "p2"):
        conv_default(f"segmentation_body.smooth_{nm}.conv", 64, 64, 3)
        p.update(_bn_keys(f"segmentation_body.smooth_{nm}.bn", 64))
    conv_default("segmentation_body.conv.0", 256, 256, 3)
    p.update(_bn_keys("segmentation_body.conv.1", 256))

    def kaiming(shape, fan_in):
        return rn(shape, math.sqrt(2.0 / fan_in))

    for br in ("binarize", "thresh"):
        pre = f"segmentation_head.{br}"
        p[pre + ".0.weight"] = kaiming((64, 256, 3, 3), 256 * 9)
        if br == "binarize":   # thresh.0 has no bias: segmentation_head.py:64-68
            p[pre + ".0.bias"] = (torch.rand((64,), generator=g) * 2 - 1) / math.sqrt(256 * 9)
        p.update(_bn_keys(pre + ".1", 64, 1.0, 1e-4))
        # ConvTranspose2d weight is (Cin, Cout, 2, 2); kaiming fan_in = Cout*k*k
        p[pre + ".3.weight"] = kaiming((64, 64, 2, 2), 64 * 4)
        p[pre + ".3.bias"] = (torch.rand((64,), generator=g) * 2 - 1) / math.sqrt(64 * 4)
        p.update(_bn_keys(pre + ".4", 64, 1.0, 1e-4))
        p[pre + ".6.weight"] = kaiming((64, 1, 2, 2), 1 * 4)
        p[pre + ".6.bias"] = (torch.rand((1,), generator=g) * 2 - 1) / math.sqrt(64 * 4)
    return p


# --------------------------------------------------------------------------------------
# model restatement (a-1 .. a-6)
# --------------------------------------------------------------------------------------

class _Ctx:
    """Carries mode + collects updated BN buffers (functional restatement of nn.BatchNorm2d)."""

    def __init__(self, params, training: bool, quant=None, taps=None):
        self.p = params
        self.taps = taps          # optional dict: named intermediate activations (the executor's dbb_net_debug_read names)
        self.training = training
        self.new_buffers: Dict[str, torch.Tensor] = {}
        self.q = quant if quant is not None else (lambda t: t)

    def bn(self, x, prefix):
        """nn.BatchNorm2d (src/modules/basic.py:34 etc.) through the same ATen op, on cloned buffers."""
        w, b = self.p[prefix + ".weight"], self.p[prefix + ".bias"]
        rm = self.p[prefix + ".running_mean"].detach().clone()
        rv = self.p[prefix + ".running_var"].detach().clone()
        y = F.batch_norm(x, rm, rv, w, b, self.training, BN_MOMENTUM, BN_EPS)
        if self.training:
            self.new_buffers[prefix + ".running_mean"] = rm
            self.new_buffers[prefix + ".running_var"] = rv
            self.new_buffers[prefix + ".num_batches_tracked"] = self.p[prefix + ".num_batches_tracked"] + 1
        return y

    def tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.detach()
        return t

    def conv(self, x, prefix, stride=1, padding=0):
        w = self.q(self.p[prefix + ".weight"])
        b = self.p.get(prefix + ".bias")
        return self.q(F.conv2d(self.q(x), w, b, stride=stride, padding=padding))

    def convT(self, x, prefix):
        w = self.q(self.p[prefix + ".weight"])
        return F.conv_transpose2d(self.q(x), w, self.p[prefix + ".bias"], stride=2)


def _basic_block(c: _Ctx, x, pre: str, stride: int):
    """src/modules/resnet.py:70-91."""
    out = F.relu(c.bn(c.conv(x, pre + ".conv1", stride, 1), pre + ".bn1"))
    out = c.bn(c.conv(out, pre + ".conv2", 1, 1), pre + ".bn2")
    if (pre + ".downsample.0.weight") in c.p:
        res = c.bn(c.conv(x, pre + ".downsample.0", stride, 0), pre + ".downsample.1")
    else:
        res = x
    return F.relu(out + res)


def resnet18_forward(c: _Ctx, x):
    """src/modules/resnet.py:231-242 -> (c2, c3, c4, c5)."""
    x = F.relu(c.bn(c.conv(x, "backbone.conv1", 2, 3), "backbone.bn1"))
    x = c.tap("x1", F.max_pool2d(x, 3, 2, 1))
    feats = []
    for li in range(1, 5):
        for bi in range(2):
            x = c.tap(f"block{(li - 1) * 2 + bi}.out", _basic_block(c, x, f"backbone.layer{li}.{bi}", 2 if (li > 1 and bi == 0) else 1))
        feats.append(x)
    return tuple(feats)


def nearest_upsample(x, size):
    """F.interpolate default mode ('nearest'), src/modules/segmentation_body.py:80,84-86:
    src = min(floor(dst * fp32(in/out)), in-1)."""
    return F.interpolate(x, size=tuple(int(s) for s in size))


def fpn_forward(c: _Ctx, feats):
    """src/modules/segmentation_body.py:64-87."""
    c2, c3, c4, c5 = feats
    sb = "segmentation_body."

    def cbr(x, name, pad):
        return F.relu(c.bn(c.conv(x, sb + name + ".conv", 1, pad), sb + name + ".bn"))

    p5 = cbr(c5, "reduce_conv_c5", 0)
    l4 = cbr(c4, "reduce_conv_c4", 0)
    p4 = cbr(nearest_upsample(p5, l4.shape[2:]) + l4, "smooth_p4", 1)
    l3 = cbr(c3, "reduce_conv_c3", 0)
    p3 = cbr(nearest_upsample(p4, l3.shape[2:]) + l3, "smooth_p3", 1)
    l2 = cbr(c2, "reduce_conv_c2", 0)
    p2 = cbr(nearest_upsample(p3, l2.shape[2:]) + l2, "smooth_p2", 1)
    hw = p2.shape[2:]
    cat = torch.cat([p2, nearest_upsample(p3, hw), nearest_upsample(p4, hw), nearest_upsample(p5, hw)], 1)
    for nm, t in (("p5", p5), ("l4", l4), ("p4", p4), ("l3", l3), ("p3", p3), ("l2", l2), ("cat", cat)):
        c.tap(nm, t)
    return c.tap("af", F.relu(c.bn(c.conv(cat, sb + "conv.0", 1, 1), sb + "conv.1")))


def step_function(p, t, k=50.0):
    """src/modules/segmentation_head.py:106-108 (NOT the dead one in losses.py:5-8)."""
    return torch.reciprocal(1 + torch.exp(-k * (p - t)))


def dbhead_forward(c: _Ctx, x, k=50.0):
    """src/modules/segmentation_head.py:35-45."""
    outs = []
    ahs = []
    for br in ("binarize", "thresh"):
        pre = "segmentation_head." + br
        y = F.relu(c.bn(c.conv(x, pre + ".0", 1, 1), pre + ".1"))
        ahs.append(y)
        y = F.relu(c.bn(c.q(c.convT(y, pre + ".3")), pre + ".4"))
        outs.append(torch.sigmoid(F.conv_transpose2d(y, c.p[pre + ".6.weight"], c.p[pre + ".6.bias"], stride=2)))
    c.tap("ah", torch.cat(ahs, 1))
    if c.training:
        outs.append(step_function(outs[0], outs[1], k))
    return torch.cat(outs, 1)


def dbnet_forward(params, x, training: bool, quant=None, return_buffers=False, taps=None):
    """src/models.py:34-48.  ``quant`` (optional) is applied to conv inputs, weights and raw
    conv outputs -- used to emulate the product's bf16 rounding points."""
    c = _Ctx(params, training, quant, taps)
    H, W = x.shape[2:]
    y = dbhead_forward(c, fpn_forward(c, resnet18_forward(c, x)))
    if tuple(y.shape[2:]) != (H, W):
        y = F.interpolate(y, size=(H, W), mode="bilinear", align_corners=True)
    # (bit-exact identity when H, W are multiples of 4: SURVEY.md F6)
    return (y, c.new_buffers) if return_buffers else y


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


# --------------------------------------------------------------------------------------
# loss restatement (a-7 .. a-10), numpy, float64 accumulation
# --------------------------------------------------------------------------------------

def bce_elementwise(p: np.ndarray, g: np.ndarray) -> np.ndarray:
    """F.binary_cross_entropy(reduction='none'), src/losses.py:30-32.  ATen evaluates
    (t-1)*max(log1p(-x),-100) - t*max(log(x),-100) in float32."""
    p = p.astype(np.float32)
    g = g.astype(np.float32)
    with np.errstate(divide="ignore"):
        l0 = np.maximum(np.log(p, dtype=np.float32), np.float32(-100.0))
        l1 = np.maximum(np.log1p(-p, dtype=np.float32), np.float32(-100.0))
    return ((g - np.float32(1.0)) * l1 - g * l0).astype(np.float32)


def bce_grad(p: np.ndarray, g: np.ndarray) -> np.ndarray:
    """ATen binary_cross_entropy_backward: (p - g) / max(p (1-p), 1e-12)."""
    p = p.astype(np.float64)
    return (p - g) / np.maximum(p * (1.0 - p), 1e-12)


def ohem_counts(gt: np.ndarray, mask: np.ndarray, ratio: float) -> Tuple[int, int]:
    """src/losses.py:22-28, counted in integers (SURVEY.md section 9 'Counts')."""
    pos = (gt * mask).astype(np.float64).sum()
    neg = ((1 - gt) * mask).astype(np.float64).sum()
    n_pos = int(pos)
    n_neg = min(int(n_pos * ratio), int(neg))
    return n_pos, n_neg


def db_loss(preds: np.ndarray, gts: np.ndarray, alpha=1.0, beta=10.0, reduction="mean",
            negative_ratio=3, eps=1e-6, k=50.0, with_grad=True) -> dict:
    """DBLoss.forward, src/losses.py:105-139, plus closed-form d total / d preds.

    preds (N, 3|2, H, W) float32; gts (4, N, H, W) float32 (NOT (N,4,H,W): SURVEY F4).
    Returns a dict with the five loss terms (reference order), OHEM bookkeeping
    (n_pos, n_neg, tau, n_above, n_tie_taken) and ``grad`` = d(returned scalar)/d preds.
    """
    assert preds.ndim == 4 and gts.ndim == 4
    P = preds[:, 0].astype(np.float32)
    T = preds[:, 1].astype(np.float32)
    g = gts[0].astype(np.float32)
    m = gts[1].astype(np.float32)
    tg = gts[2].astype(np.float32)
    tm = gts[3].astype(np.float32)
    px = P.size
    out: dict = {}

    # ---- OHEM balanced BCE, src/losses.py:18-40
    pos = g * m
    neg = (1 - g) * m
    n_pos, n_neg = ohem_counts(g, m, negative_ratio)
    D = n_pos + n_neg + eps
    bce = bce_elementwise(P, g)
    bg = bce_grad(P, g)
    if reduction == "mean":
        # scalar loss -> balance = mean_bce * (sum(pos) + n_neg) / D
        mean_bce = bce.astype(np.float64).sum() / px
        pos_sum = float(pos.astype(np.float64).sum())
        prob_loss = mean_bce * (pos_sum + n_neg) / D
        # topk of {mean_bce, 0}-valued map: all n_neg picks are mean_bce (n_neg <= #neg)
        gP = bg * ((pos_sum + n_neg) / D / px)
        out.update(tau=float(mean_bce), n_above=0, n_tie=n_neg)
    elif reduction == "none":
        negl = (bce * neg).astype(np.float32).reshape(-1)
        if n_neg > 0:
            part = np.partition(negl, negl.size - n_neg)
            tau = part[negl.size - n_neg]
            above = negl > tau
            n_above = int(above.sum())
            n_tie = n_neg - n_above
            top_sum = negl[above].astype(np.float64).sum() + float(tau) * n_tie
        else:
            tau, n_above, n_tie, top_sum = np.float32(np.inf), 0, 0, 0.0
            above = np.zeros_like(negl, dtype=bool)
        prob_loss = ((bce * pos).astype(np.float64).sum() + top_sum) / D
        # selected set: all strictly above tau + the first n_tie (flat order) equal to tau.
        # torch.topk's tie order is unspecified; tests compare modulo ties (SURVEY hard part 4).
        sel = above.copy()
        if n_tie > 0:
            tie_idx = np.flatnonzero(negl == tau)[:n_tie]
            sel[tie_idx] = True
        sel = sel.reshape(P.shape)
        gP = (pos + sel * neg) * bg / D
        out.update(tau=float(tau), n_above=n_above, n_tie=n_tie, selected=sel,
                   tie_mask=(negl == tau).reshape(P.shape))
    else:
        raise ValueError(reduction)
    out.update(n_pos=n_pos, n_neg=n_neg)

    # ---- masked L1, src/losses.py:75-82
    tm_sum = tm.astype(np.float64).sum() + eps
    thr_loss = (np.abs(T - tg) * tm).astype(np.float64).sum() / tm_sum
    gT = np.sign((T - tg).astype(np.float64)) * tm / tm_sum

    if preds.shape[1] == 3:
        # ---- Dice, src/losses.py:48-66
        B = preds[:, 2].astype(np.float32)
        I = (B * g * m).astype(np.float64).sum()
        U = (B * m).astype(np.float64).sum() + (g * m).astype(np.float64).sum() + eps
        bin_loss = 1 - 2.0 * I / U
        gB = -2.0 * m * (g * U - I) / (U * U)
        pt = prob_loss + beta * thr_loss
        total = alpha * bin_loss + pt
        out.update(losses=np.array([prob_loss, thr_loss, bin_loss, pt, total], dtype=np.float64),
                   dice_I=I, dice_U=U)
        if with_grad:
            grad = np.stack([gP, beta * gT, alpha * gB], 1)
            out["grad"] = grad
    else:
        pt = prob_loss + beta * thr_loss
        out.update(losses=np.array([prob_loss, thr_loss, 0.0, pt, pt], dtype=np.float64))
        if with_grad:
            out["grad"] = np.stack([gP, beta * gT], 1)
    return out


def step_grad(P: np.ndarray, T: np.ndarray, dB: np.ndarray, k=50.0):
    """d/dP, d/dT of B = 1/(1+exp(-k(P-T))): k B^2 e (not k B (1-B))."""
    e = np.exp(-k * (P.astype(np.float64) - T.astype(np.float64)))
    B = 1.0 / (1.0 + e)
    d = dB * k * B * B * e
    return d, -d


# --------------------------------------------------------------------------------------
# post-processing front (a-11 .. a-15)
# --------------------------------------------------------------------------------------

def binarize(pred: np.ndarray, thresh: float) -> np.ndarray:
    """src/postprocess.py:51-52 -- strict '>' of float32 P against the python float thresh.
    torch compares float32 tensor with a python scalar in float32 (scalar is cast down)."""
    return pred.astype(np.float32) > np.float32(thresh)


def candidates_cv2(bitmap: np.ndarray):
    """src/postprocess.py:67-68,116-117 -- the reference's own call into OpenCV."""
    import cv2
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    return contours


def box_score_fast(pred: np.ndarray, contour_xy: np.ndarray) -> float:
    """src/postprocess.py:186-198 (np.int -> int, F9)."""
    import cv2
    h, w = pred.shape[:2]
    box = contour_xy.copy()
    xmin = int(np.clip(np.floor(box[:, 0].min()).astype(int), 0, w - 1))
    xmax = int(np.clip(np.ceil(box[:, 0].max()).astype(int), 0, w - 1))
    ymin = int(np.clip(np.floor(box[:, 1].min()).astype(int), 0, h - 1))
    ymax = int(np.clip(np.ceil(box[:, 1].max()).astype(int), 0, h - 1))
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype(np.int32), 1)
    return cv2.mean(pred[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def get_mini_boxes(contour):
    """src/postprocess.py:158-184."""
    import cv2
    try:
        bb = cv2.minAreaRect(contour)
        pts = sorted(list(cv2.boxPoints(bb)), key=lambda x: x[0])
        i1, i4 = (0, 1) if pts[1][1] > pts[0][1] else (1, 0)
        i2, i3 = (2, 3) if pts[3][1] > pts[2][1] else (3, 2)
        return [pts[i1], pts[i2], pts[i3], pts[i4]], min(bb[1])
    except Exception:
        return [], -1


def postprocess_front_cv2(pred: np.ndarray, thresh=0.25, box_thresh=0.5, max_candidates=1000, min_size=3):
    """Reference front half in box mode, src/postprocess.py:106-130 steps 1-2 (before unclip).

    Returns a list (reference order = reverse raster discovery) of dicts per candidate:
      score (float64), sside, keep (bool: survived size + score filter), fill-set stats
      (count, bbox) so the GPU path can be matched candidate by candidate.
    """
    import cv2
    bitmap = binarize(pred, thresh)
    contours = candidates_cv2(bitmap)
    res = []
    for contour in contours[:max_candidates]:
        cxy = contour.squeeze(1)
        pts, sside = get_mini_boxes(cxy)
        score = box_score_fast(pred, cxy)
        # fill set for identification
        x0, y0 = cxy[:, 0].min(), cxy[:, 1].min()
        x1, y1 = cxy[:, 0].max(), cxy[:, 1].max()
        mask = np.zeros((y1 - y0 + 1, x1 - x0 + 1), dtype=np.uint8)
        cv2.fillPoly(mask, (cxy - [x0, y0]).reshape(1, -1, 2).astype(np.int32), 1)
        keep = (sside >= min_size) and not (box_thresh > score)
        res.append(dict(score=float(score), sside=float(sside), keep=bool(keep),
                        count=int(mask.sum()), bbox=(int(x0), int(y0), int(x1), int(y1)),
                        points=np.array(pts, dtype=np.float32) if len(pts) else np.zeros((0, 2), np.float32)))
    return bitmap, res


def candidates_ccl(pred: np.ndarray, thresh: float):
    """Contour-free restatement of the candidate sets ('Contour-free score
    sets'), the formulation the GPU path implements.

    Foreground is labelled 8-connected, background 4-connected (with a virtual outside).
    Candidate fill sets (identities verified against cv2 in tests/test_oracle_golden.py):
      outer(F) = F  U  {pixels not 4-reachable from outside F's bounding complement}
               = F + all descendants in the F/G containment tree
      hole(G)  = G + descendants + pixels of parent F that are 4-adjacent to G
    Returns list of dicts(kind, count, sum(float64), bbox) in raster-discovery order of the
    candidate's first (top-left-most in raster order) border pixel.
    """
    from scipy import ndimage
    pred = pred.astype(np.float32)
    fg = binarize(pred, thresh)
    H, W = fg.shape
    s8 = np.ones((3, 3), dtype=int)
    lf, nf = ndimage.label(fg, structure=s8)
    # background with a 1-px virtual frame so everything touching the border joins 'outside'
    bgp = np.ones((H + 2, W + 2), dtype=bool)
    bgp[1:-1, 1:-1] = ~fg
    lg, ng = ndimage.label(bgp)   # default structure = 4-connectivity
    outside = lg[0, 0]
    cands = []
    p64 = pred.astype(np.float64)
    # ---- outer candidates: fill holes of each component
    for f in range(1, nf + 1):
        comp = lf == f
        ys, xs = np.nonzero(comp)
        y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
        sub = comp[y0:y1 + 1, x0:x1 + 1]
        # complement of comp, 4-connected, reachable from outside the bbox
        pad = np.ones((sub.shape[0] + 2, sub.shape[1] + 2), dtype=bool)
        pad[1:-1, 1:-1] = ~sub
        lc, _ = ndimage.label(pad)
        filled = (lc != lc[0, 0])[1:-1, 1:-1]
        first = (int(ys[0]), int(xs[0]))   # raster-first pixel of the component
        cands.append(dict(kind="outer", count=int(filled.sum()),
                          sum=float(p64[y0:y1 + 1, x0:x1 + 1][filled].sum()),
                          bbox=(int(x0), int(y0), int(x1), int(y1)), first=first, label=f))
    # ---- hole candidates
    for gi in range(1, ng + 1):
        if gi == outside:
            continue
        reg = (lg == gi)[1:-1, 1:-1]
        ys, xs = np.nonzero(reg)
        y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
        # fill8(R): R plus everything it encloses, complement taken 8-connected
        Y0, Y1, X0, X1 = max(y0 - 1, 0), min(y1 + 1, H - 1), max(x0 - 1, 0), min(x1 + 1, W - 1)
        sub = reg[Y0:Y1 + 1, X0:X1 + 1]
        pad = np.ones((sub.shape[0] + 2, sub.shape[1] + 2), dtype=bool)
        pad[1:-1, 1:-1] = ~sub
        lc, _ = ndimage.label(pad, structure=s8)
        fill8 = (lc != lc[0, 0])[1:-1, 1:-1]
        # foreground pixels 4-adjacent to R
        dil = ndimage.binary_dilation(sub, structure=ndimage.generate_binary_structure(2, 1))
        ring = dil & fg[Y0:Y1 + 1, X0:X1 + 1]
        full = fill8 | ring
        fy, fx = np.nonzero(full)
        # the hole border's first raster pixel is the foreground pixel just above R's first pixel row
        cands.append(dict(kind="hole", count=int(full.sum()),
                          sum=float(p64[Y0:Y1 + 1, X0:X1 + 1][full].sum()),
                          bbox=(int(fx.min() + X0), int(fy.min() + Y0), int(fx.max() + X0), int(fy.max() + Y0)),
                          first=(int(ys[0]), int(xs[0])), label=gi))
    return fg, cands


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------

def synth_gt_maps(n: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """(4, N, H, W) float32 GT maps in the reference's value domains
    (src/data_loaders.py:87-94,116-117,148-149): prob_map {0,1}, supervision_mask {0,1},
    thresh_map [0.3, 0.7], text_area_map {0,1}."""
    import cv2
    rng = np.random.RandomState(seed)
    gts = np.zeros((4, n, h, w), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        prob = np.zeros((h, w), np.uint8)
        area = np.zeros((h, w), np.uint8)
        thr = np.full((h, w), 0.3, np.float32)
        mask = np.ones((h, w), np.uint8)
        scale = max(min(h, w) / 640.0, 0.1)
        for _ in range(rng.randint(10, 31)):
            cx, cy = rng.uniform(0, w), rng.uniform(0, h)
            rw = rng.uniform(24, 220) * scale + 8
            rh = rng.uniform(8, 48) * scale + 8
            ang = rng.uniform(-45, 45)
            area_r, per = rw * rh, 2 * (rw + rh)
            d = area_r * (1 - 0.4 ** 2) / per
            if min(rw, rh) - 2 * d < 1:
                continue
            shrink = cv2.boxPoints(((cx, cy), (rw - 2 * d, rh - 2 * d), ang)).astype(np.int32)
            dil = cv2.boxPoints(((cx, cy), (rw + 2 * d, rh + 2 * d), ang)).astype(np.int32)
            cv2.fillPoly(prob, [shrink], 1)
            tmp = np.zeros((h, w), np.uint8)
            cv2.fillPoly(tmp, [dil], 1)
            area |= tmp
            # distance to the rectangle edge in the rectangle's frame
            c, s = math.cos(math.radians(ang)), math.sin(math.radians(ang))
            u = (xx - cx) * c + (yy - cy) * s
            v = -(xx - cx) * s + (yy - cy) * c
            dist = np.abs(np.maximum(np.abs(u) - rw / 2, np.abs(v) - rh / 2))
            t = 0.3 + 0.4 * np.clip(1 - dist / d, 0, 1)
            thr = np.where(tmp > 0, np.maximum(thr, t.astype(np.float32)), thr)
        for _ in range(rng.randint(0, 3)):
            cx, cy = rng.uniform(0, w), rng.uniform(0, h)
            ign = cv2.boxPoints(((cx, cy), (rng.uniform(20, 120) * scale + 4, rng.uniform(8, 40) * scale + 4),
                                 rng.uniform(-45, 45))).astype(np.int32)
            cv2.fillPoly(mask, [ign], 0)
        gts[0, i] = prob
        gts[1, i] = mask
        gts[2, i] = thr
        gts[3, i] = area
    return gts


def synth_prob_map(h: int, w: int, seed: int = 0, sigma: float = 6.0) -> np.ndarray:
    """Blurred-noise blob map for the post-processing tests."""
    import cv2
    rng = np.random.RandomState(seed)
    z = rng.uniform(0, 1, (h, w)).astype(np.float32)
    z = cv2.GaussianBlur(z, (0, 0), sigma)
    z = (z - z.min()) / max(float(z.max() - z.min()), 1e-12)
    return z.astype(np.float32)


def synth_images(n: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n, 3, h, w), generator=g) * 60.0


def synth_text_batch(n: int, h: int, w: int, seed: int = 0):
    """Images that DEPICT their ground truth (for conditioning a network by a few training steps and for BASELINE-size
    goldens): N(0, 25^2) noise + a per-channel offset inside the dilated text area + another inside the shrunk core.
    Returns (x float32 (N,3,H,W) torch, gts float32 (4,N,H,W) numpy).  Value range ~ +-125 like mean-subtracted pixels
    (src/utils.py:190-193)."""
    gts = synth_gt_maps(n, h, w, seed)
    g = torch.Generator().manual_seed(1000 + seed)
    area = torch.from_numpy(gts[3]).unsqueeze(1)
    core = torch.from_numpy(gts[0]).unsqueeze(1)
    img = torch.randn((n, 3, h, w), generator=g) * 25.0
    img = img + torch.tensor([70.0, -50.0, 60.0]).view(1, 3, 1, 1) * (area - 0.5)
    img = img + torch.tensor([30.0, 40.0, -35.0]).view(1, 3, 1, 1) * core
    return img.contiguous(), gts


COND_SEED = 7          # init_params seed of the conditioned network (tests/golden/cond_params.npz)


def cond_trainable(key: str, numel_dim: int) -> bool:
    """The parameter subset the golden generator lets the REFERENCE train (everything else stays at init_params(7) so the
    fixture stays small): the whole DBHead and every 1-D tensor (BatchNorm affine, conv biases) outside the unused
    backbone.fc / backbone.smooth."""
    if key.startswith("backbone.fc") or key.startswith("backbone.smooth"):
        return False
    return key.startswith("segmentation_head") or numel_dim == 1


def cond_params(golden_dir: str) -> Dict[str, torch.Tensor]:
    """init_params(COND_SEED) overlaid with the tensors the reference trained (oracle/make_golden.py:make_cond_params)."""
    import os
    z = np.load(os.path.join(golden_dir, "cond_params.npz"))
    p = init_params(COND_SEED)
    for k in z.files:
        if k.startswith("p:"):
            t = torch.from_numpy(z[k])
            assert p[k[2:]].shape == t.shape, k
            p[k[2:]] = t.to(p[k[2:]].dtype)
    return p


def strided_summary(a: np.ndarray, stride: int = 4, block: int = 16):
    """(samples a[..., ::stride, ::stride], float64 block sums over block x block tiles) of a (..., H, W) map: every pixel
    contributes to the fixture while the .npz stays small."""
    a = np.asarray(a)
    H, W = a.shape[-2:]
    hb, wb = -(-H // block) * block, -(-W // block) * block
    pad = np.zeros(a.shape[:-2] + (hb, wb), np.float64)
    pad[..., :H, :W] = a
    bs = pad.reshape(a.shape[:-2] + (hb // block, block, wb // block, block)).sum(axis=(-3, -1))
    return np.ascontiguousarray(a[..., ::stride, ::stride]), bs


# --------------------------------------------------------------------------------------
# GT border map: the distance field of draw_thresh_map
# --------------------------------------------------------------------------------------

def polygon_area_length(poly: np.ndarray):
    """shapely Polygon.area / .length of a simple polygon (shoelace, perimeter), as used at
    src/db_transforms.py:13-17 and src/data_loaders.py:107-121."""
    p = np.asarray(poly, dtype=np.float64)
    q = np.concatenate([p[1:], p[:1]])
    area = 0.5 * abs(float((p[:, 0] * q[:, 1] - p[:, 1] * q[:, 0]).sum()))
    length = float(np.sqrt(((p - q) ** 2).sum(1)).sum())
    return area, length


def segment_distance(xs, ys, a, b):
    """src/db_transforms.py:62-78 (compute_distance): distance of every grid point to segment a-b through the law of
    cosines, float64, including its behaviour at the degenerate points (nan_to_num of 1 - cos^2, the cos < 0 branch)."""
    d1 = (xs - a[0]) ** 2 + (ys - a[1]) ** 2
    d2 = (xs - b[0]) ** 2 + (ys - b[1]) ** 2
    d = float((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2)
    with np.errstate(all="ignore"):
        cosin = (d - d1 - d2) / (2 * np.sqrt(d1 * d2))
        sq_sin = np.nan_to_num(1 - cosin ** 2)
        res = np.sqrt(d1 * d2 * sq_sin / d)
        neg = cosin < 0
        res[neg] = np.sqrt(np.fmin(d1, d2))[neg]
    return res


def thresh_map_accumulate(canvas: np.ndarray, polygon, padded_bbox, distance: float):
    """src/db_transforms.py:26-59: canvas = fmax(canvas, 1 - min_edges clip(dist/distance, 0, 1)) over the bounding box
    (xmin, ymin, xmax, ymax) of the dilated polygon.  The dilation itself (pyclipper) is an input here."""
    poly = np.array(polygon).copy()
    xmin, ymin, xmax, ymax = (int(v) for v in padded_bbox)
    width, height = xmax - xmin + 1, ymax - ymin + 1
    poly[:, 0] = poly[:, 0] - xmin
    poly[:, 1] = poly[:, 1] - ymin
    xs = np.broadcast_to(np.arange(width, dtype=np.float64).reshape(1, width), (height, width))
    ys = np.broadcast_to(np.arange(height, dtype=np.float64).reshape(height, 1), (height, width))
    dm = np.zeros((poly.shape[0], height, width), dtype=np.float32)
    for i in range(poly.shape[0]):
        j = (i + 1) % poly.shape[0]
        with np.errstate(all="ignore"):
            dm[i] = np.clip(segment_distance(xs, ys, poly[i], poly[j]) / distance, 0, 1)
    dm = dm.min(axis=0)
    H, W = canvas.shape
    x0, x1 = min(max(0, xmin), W - 1), min(max(0, xmax), W - 1)
    y0, y1 = min(max(0, ymin), H - 1), min(max(0, ymax), H - 1)
    canvas[y0:y1 + 1, x0:x1 + 1] = np.fmax(1 - dm[y0 - ymin:y1 - ymax + height, x0 - xmin:x1 - xmax + width],
                                           canvas[y0:y1 + 1, x0:x1 + 1])
    return canvas


def gt_maps_reference(anns, image_size, offset_fn, shrink_ratio=0.4, thresh_min=0.3, thresh_max=0.7, min_text_size=8,
                      ignore_tags=("*", "###")):
    """src/data_loaders.py:86-149 for ONE image with OpenCV / numpy exactly as the reference runs them; the two third-party
    geometry calls are parameters: `offset_fn(poly, delta)` stands for pyclipper's Execute (list of int arrays), shapely's
    Polygon area / length are the shoelace / perimeter formulas and its validity tests are taken as true.
    Returns (gt, mask, thresh_map, thresh_mask, ignore_tags)."""
    import cv2
    S = image_size
    gt = np.zeros((S, S), np.float32)
    mask = np.ones((S, S), np.float32)
    thresh_map = np.zeros((S, S), np.float32)
    thresh_mask = np.zeros((S, S), np.float32)
    tags = []
    for ann in anns:
        poly = np.array(ann["poly"])
        height = max(poly[:, 1]) - min(poly[:, 1])
        width = max(poly[:, 0]) - min(poly[:, 0])
        area, length = polygon_area_length(poly)
        if area < 1 or min(height, width) < min_text_size or ann.get("text") in ignore_tags:
            tags.append(True)
            cv2.fillPoly(mask, poly.astype(np.int32)[np.newaxis, :, :], 0)
            continue
        distance = area * (1 - np.power(shrink_ratio, 2)) / length
        shrinked = offset_fn(poly, -distance)
        if len(shrinked) == 0:
            tags.append(True)
            cv2.fillPoly(mask, poly.astype(np.int32)[np.newaxis, :, :], 0)
            continue
        sh = np.array(shrinked[0]).reshape(-1, 2)
        if sh.shape[0] > 2:
            tags.append(False)
            cv2.fillPoly(gt, [sh.astype(np.int32)], 1)
        else:
            tags.append(True)
            cv2.fillPoly(mask, poly.astype(np.int32)[np.newaxis, :, :], 0)
            continue
        # draw_thresh_map, src/db_transforms.py:8-59
        padded = np.array(offset_fn(poly, distance)[0])
        cv2.fillPoly(thresh_mask, [padded.astype(np.int32)], 1.0)
        bbox = (padded[:, 0].min(), padded[:, 1].min(), padded[:, 0].max(), padded[:, 1].max())
        thresh_map_accumulate(thresh_map, poly.astype(np.float64), bbox, distance)
    thresh_map = thresh_map * (thresh_max - thresh_min) + thresh_min
    return gt, mask, thresh_map, thresh_mask, tags
