"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

Run in the build container only:  python -m oracle.make_golden
(TEST INFRASTRUCTURE; the fixtures are committed, the reference is not.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import db_oracle as O   # noqa: E402
from oracle import ref_import       # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def grad_summary(t: torch.Tensor):
    t = t.detach().double().reshape(-1)
    return np.array([t.norm().item(), t.sum().item(), *t[:4].tolist()] + [0.0] * max(0, 4 - t.numel()))[:6]


def make_model_case(name, seed, n, h, w):
    torch.manual_seed(seed)
    params = O.init_params(seed)
    x = O.synth_images(n, h, w, seed)
    gts = torch.from_numpy(O.synth_gt_maps(n, h, w, seed))
    _, losses, _ = ref_import.load()
    m = ref_import.build_model(params)
    out = {}
    # eval
    m.eval()
    with torch.no_grad():
        out["eval"] = m(x).numpy()
    # train + loss + backward
    m.train()
    y = m(x)
    out["train"] = y.detach().numpy()
    for red in ("mean", "none"):
        m.zero_grad()
        crit = losses.DBLoss(alpha=1.0, beta=10.0, reduction=red, negative_ratio=3)
        ls = crit(y, gts)
        ls[-1].backward(retain_graph=True)
        out[f"losses_{red}"] = np.array([float(v) for v in ls])
        keys, summ = [], []
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            keys.append(k)
            summ.append(grad_summary(p.grad))
        out[f"grad_keys_{red}"] = np.array(keys)
        out[f"grad_summary_{red}"] = np.stack(summ)
        # a few full small gradients
        out[f"grad_bn1_weight_{red}"] = m.backbone.bn1.weight.grad.numpy().copy()
        out[f"grad_head_b6w_{red}"] = m.segmentation_head.binarize[6].weight.grad.numpy().copy()
        out[f"grad_head_t6w_{red}"] = m.segmentation_head.thresh[6].weight.grad.numpy().copy()
        out[f"grad_fpn_c5_w_{red}"] = m.segmentation_body.reduce_conv_c5.conv.weight.grad.numpy().copy()
    sd = m.state_dict()
    for k in ("backbone.bn1.running_mean", "backbone.bn1.running_var",
              "segmentation_head.thresh.4.running_mean", "segmentation_head.thresh.4.running_var",
              "segmentation_body.conv.1.running_var", "backbone.bn1.num_batches_tracked"):
        out["buf:" + k] = sd[k].numpy()
    out["meta"] = np.array([seed, n, h, w])
    out["x_checksum"] = np.array([x.double().sum().item(), x.double().abs().sum().item()])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if k in ("eval", "train")}, out["losses_mean"])


def make_cond_params(steps=100):
    """Conditions a network with the UNMODIFIED reference: DBTextModel + DBLoss('mean') + torch.optim.Adam(lr 0.005), as
    src/train.py:110-117,160-172 does, for `steps` steps on 4x3x128x128 synthetic text batches.  Only the subset
    O.cond_trainable() is trained so that the fixture is ~1.3 MB; the rest stays at O.init_params(O.COND_SEED).
    A randomly initialised DB network has saturated sigmoid logits (kaiming init of the 64->1 ConvTranspose) and is too
    ill-conditioned to compare a bf16 pipeline against; this one produces text-like maps."""
    _, losses, _ = ref_import.load()
    torch.manual_seed(0)
    params = O.init_params(O.COND_SEED)
    m = ref_import.build_model(params)
    train_p = []
    for k, p in m.named_parameters():
        ok = O.cond_trainable(k, p.dim())
        p.requires_grad_(ok)
        if ok:
            train_p.append(p)
    opt = torch.optim.Adam(train_p, lr=0.005)
    crit = losses.DBLoss(alpha=1.0, beta=10.0, reduction="mean", negative_ratio=3)
    m.train()
    hist = []
    for it in range(steps):
        x, g = O.synth_text_batch(4, 128, 128, it)
        ls = crit(m(x), torch.from_numpy(g))
        opt.zero_grad()
        ls[-1].backward()
        opt.step()
        hist.append(float(ls[-1]))
    out = {}
    ref = O.init_params(O.COND_SEED)
    for k, v in m.state_dict().items():
        if not torch.equal(v, ref[k]):
            out["p:" + k] = v.detach().numpy().copy()
    out["loss_history"] = np.array(hist)
    np.savez_compressed(os.path.join(GOLD, "cond_params.npz"), **out)
    print("cond params:", len(out) - 1, "tensors,", sum(v.size for k, v in out.items() if k.startswith("p:")), "values; loss",
          hist[0], "->", hist[-1])


def make_baseline_size_cases():
    """BASELINE.json configs at their real sizes on the conditioned network: config 1 (1x3x640x640 eval + the
    post-processing front's candidate rows), config 2 (2 of the 16 images of a 640x640 training step: P, T, B, losses,
    gradient summaries), config 4 (one 1024x1024 eval image).  Maps are stored as stride-4 samples + 16x16 block sums."""
    import cv2
    _, losses, postprocess = ref_import.load()
    params = O.cond_params(GOLD)
    m = ref_import.build_model(params)

    def put(out, key, arr):
        out[key + ":samples"], out[key + ":blocks"] = O.strided_summary(arr)

    # ---- config 1
    out = {}
    x, _ = O.synth_text_batch(1, 640, 640, 901)
    m.eval()
    with torch.no_grad():
        y = m(x).numpy()
    put(out, "eval", y)
    out["meta"] = np.array([901, 1, 640, 640])
    out["x_checksum"] = np.array([x.double().sum().item(), x.double().abs().sum().item()])
    rep = postprocess.SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.5)
    P = y[0, 0]
    bitmap = rep.binarize(torch.from_numpy(P)).numpy()
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    rows = []
    for contour in contours[:rep.max_candidates]:
        c = contour.squeeze(1)
        pts, sside = rep.get_mini_boxes(c)
        score = rep.box_score_fast(P, c)
        x0, y0, x1, y1 = c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()
        mk = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
        cv2.fillPoly(mk, (c - [x0, y0]).reshape(1, -1, 2).astype(np.int32), 1)
        keep = (not (sside < rep.min_size)) and (not (rep.box_thresh > score))
        rows.append([score, sside, float(keep), float(mk.sum()), x0, y0, x1, y1] + list(np.array(pts).reshape(-1)))
    out["bitmap_packed"] = np.packbits(bitmap.astype(np.uint8))
    out["ncontours"] = np.array([len(contours)])
    out["cands"] = np.array(rows, dtype=np.float64).reshape(len(rows), 16)
    np.savez_compressed(os.path.join(GOLD, "model_c1_640_eval.npz"), **out)
    print("config 1:", len(contours), "contours,", int(sum(r[2] for r in rows)), "kept; P>0.25:", float(bitmap.mean()))

    # ---- config 2 (2 images)
    out = {}
    x, gts = O.synth_text_batch(2, 640, 640, 905)
    m.train()
    for p in m.parameters():
        p.requires_grad_(True)
    y = m(x)
    put(out, "train", y.detach().numpy())
    for red in ("mean", "none"):
        m.zero_grad()
        crit = losses.DBLoss(alpha=1.0, beta=10.0, reduction=red, negative_ratio=3)
        ls = crit(y, torch.from_numpy(gts))
        ls[-1].backward(retain_graph=True)
        out[f"losses_{red}"] = np.array([float(v) for v in ls])
        keys, summ = [], []
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            keys.append(k)
            summ.append(grad_summary(p.grad))
        out[f"grad_keys_{red}"] = np.array(keys)
        out[f"grad_summary_{red}"] = np.stack(summ)
        for k, p in m.named_parameters():      # every 1-D gradient in full (BatchNorm affine, biases): 9.9 k values
            if p.grad is not None and p.dim() == 1:
                out[f"grad_{red}:{k}"] = p.grad.numpy().copy()
        out[f"grad_{red}:backbone.conv1.weight"] = m.backbone.conv1.weight.grad.numpy().copy()
        out[f"grad_{red}:segmentation_head.binarize.6.weight"] = m.segmentation_head.binarize[6].weight.grad.numpy().copy()
        out[f"grad_{red}:segmentation_head.thresh.6.weight"] = m.segmentation_head.thresh[6].weight.grad.numpy().copy()
        out[f"grad_{red}:segmentation_head.thresh.3.weight"] = m.segmentation_head.thresh[3].weight.grad.numpy().copy()
    out["meta"] = np.array([905, 2, 640, 640])
    out["x_checksum"] = np.array([x.double().sum().item(), x.double().abs().sum().item()])
    np.savez_compressed(os.path.join(GOLD, "model_c2_640_train.npz"), **out)
    print("config 2:", out["losses_mean"], out["losses_none"])

    # ---- config 4 (one image); a fresh model: the training forward above moved the BatchNorm running statistics
    out = {}
    x, _ = O.synth_text_batch(1, 1024, 1024, 909)
    m = ref_import.build_model(params)
    m.eval()
    with torch.no_grad():
        y = m(x).numpy()
    put(out, "eval", y)
    out["meta"] = np.array([909, 1, 1024, 1024])
    out["x_checksum"] = np.array([x.double().sum().item(), x.double().abs().sum().item()])
    np.savez_compressed(os.path.join(GOLD, "model_c4_1024_eval.npz"), **out)
    print("config 4: P>0.25:", float((y[0, 0] > 0.25).mean()))


def make_loss_cases():
    _, losses, _ = ref_import.load()
    rng = np.random.RandomState(7)
    cases = {}

    def rand_preds(n, c, h, w):
        p = rng.uniform(0.02, 0.98, (n, c, h, w)).astype(np.float32)
        if c == 3:
            p[:, 2] = (1.0 / (1.0 + np.exp(-50.0 * (p[:, 0].astype(np.float64) - p[:, 1])))).astype(np.float32)
        return p

    n, h, w = 2, 64, 64
    gts = O.synth_gt_maps(n, h, w, 3)
    cases["random"] = (rand_preds(n, 3, h, w), gts)
    cases["eval2ch"] = (rand_preds(n, 2, h, w), gts)
    cases["ragged"] = (rand_preds(3, 3, 40, 52), O.synth_gt_maps(3, 40, 52, 4))
    g0 = gts.copy(); g0[0] = 0
    cases["nopos"] = (rand_preds(n, 3, h, w), g0)
    g1 = gts.copy(); g1[1] = 0
    cases["allmasked"] = (rand_preds(n, 3, h, w), g1)
    ps = rand_preds(n, 3, h, w)
    ps[:, 0].reshape(-1)[::7] = 0.0
    ps[:, 0].reshape(-1)[3::11] = 1.0
    cases["saturated"] = (ps, gts)
    # k >= number of non-zero negatives: almost everything positive
    g2 = gts.copy(); g2[0] = 1; g2[0, :, :4, :4] = 0
    cases["kbig"] = (rand_preds(n, 3, h, w), g2)
    # heavy ties in the negative losses: quantised P
    pq = rand_preds(n, 3, h, w)
    pq[:, 0] = np.round(pq[:, 0] * 8) / 8 * 0.9 + 0.05
    cases["ties"] = (pq, gts)

    out = {}
    for cname, (p, g) in cases.items():
        out[cname + ":preds"] = p
        out[cname + ":gts"] = g
        for red in ("mean", "none"):
            pt = torch.from_numpy(p).clone().requires_grad_(True)
            crit = losses.DBLoss(alpha=1.0, beta=10.0, reduction=red, negative_ratio=3)
            ls = crit(pt, torch.from_numpy(g))
            if isinstance(ls, tuple):
                ls[-1].backward()
                vals = [float(v) for v in ls]
            else:
                ls.backward()
                vals = [float(ls)]
            out[f"{cname}:{red}:losses"] = np.array(vals)
            out[f"{cname}:{red}:grad"] = pt.grad.numpy()
            pos = torch.from_numpy(g[0] * g[1]); neg = torch.from_numpy((1 - g[0]) * g[1])
            n_pos = int(pos.sum()); n_neg = min(int(n_pos * 3), int(neg.sum()))
            out[f"{cname}:{red}:counts"] = np.array([n_pos, n_neg])
    np.savez_compressed(os.path.join(GOLD, "loss_cases.npz"), **out)
    print("loss cases", list(cases))


def handmade_maps():
    """SURVEY.md section 9 examples + the section 7 step-4 list."""
    maps = {}
    a = np.full((96, 96), 0.05, np.float32)
    a[5:25, 5:45] = 0.9                     # solid block
    a[40:80, 10:60] = 0.8; a[50:70, 25:45] = 0.1    # ring with hole
    a[30:32, 70:72] = 0.95                  # 2x2 blob (sside 1)
    a[85:90, 50:85] = 0.3                   # weak strip: binarised, score < 0.5
    maps["survey96"] = a
    b = np.full((80, 80), 0.0, np.float32)
    b[4:76, 4:76] = 0.9; b[12:68, 12:68] = 0.1; b[20:60, 20:60] = 0.7; b[28:52, 28:52] = 0.2
    b[34:46, 34:46] = 0.95                  # nested rings
    maps["nested80"] = b
    c = np.full((64, 64), 0.1, np.float32)
    c[0:10, 0:20] = 0.8; c[54:64, 40:64] = 0.6; c[20:40, 0:3] = 0.9; c[0:64, 63] = 0.7   # border-touching
    c[30, 10:50] = 0.9                      # 1-px line
    c[45:50, 20] = 0.9                      # 1-px vertical line
    maps["border64"] = c
    d = np.full((48, 48), 0.1, np.float32)
    yy, xx = np.mgrid[0:48, 0:48]
    d[(yy + xx) % 2 == 0] = 0.6             # checkerboard: one 8-connected FG, many 1-px holes
    maps["checker48"] = d
    e = np.full((160, 160), 0.1, np.float32)
    e[::4, ::4] = 0.9                       # 1600 isolated pixels -> > max_candidates
    maps["many160"] = e
    f = np.full((64, 64), 0.1, np.float32)
    f[10:30, 10:30] = 0.8; f[18:22, 18:22] = 0.1; f[22:26, 22:26] = 0.1   # diagonal-touching holes
    f[40:56, 8:24] = 0.8; f[44:52, 12:20] = 0.1; f[46:50, 14:18] = 0.9    # island inside hole
    f[40:41, 40:60] = 0.7; f[41:56, 40:41] = 0.7; f[55:56, 40:60] = 0.7; f[41:56, 59:60] = 0.7  # thin frame
    maps["holes64"] = f
    maps["empty32"] = np.full((32, 32), 0.1, np.float32)
    maps["full32"] = np.full((32, 32), 0.9, np.float32)
    return maps


def make_post_cases():
    import cv2
    _, _, postprocess = ref_import.load()
    maps = handmade_maps()
    maps["blobs256"] = O.synth_prob_map(256, 256, 11)
    maps["blobs200x312"] = O.synth_prob_map(200, 312, 12)
    rng = np.random.RandomState(5)
    maps["noise128"] = rng.uniform(0, 1, (128, 128)).astype(np.float32)   # salt-and-pepper stress
    out = {}
    for name, P in maps.items():
        rep = postprocess.SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.5)
        pt = torch.from_numpy(P)
        bitmap = rep.binarize(pt).numpy()
        # src/postprocess.py:116-118
        contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
        rows = []
        for contour in contours[:rep.max_candidates]:
            c = contour.squeeze(1)
            pts, sside = rep.get_mini_boxes(c)
            score = rep.box_score_fast(P, c)
            x0, y0, x1, y1 = c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()
            m = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
            cv2.fillPoly(m, (c - [x0, y0]).reshape(1, -1, 2).astype(np.int32), 1)
            keep = (not (sside < rep.min_size)) and (not (rep.box_thresh > score))
            rows.append([score, sside, float(keep), float(m.sum()), x0, y0, x1, y1] + list(np.array(pts).reshape(-1)))
        out[name + ":P"] = P
        out[name + ":bitmap"] = bitmap
        out[name + ":ncontours"] = np.array([len(contours)])
        out[name + ":cands"] = np.array(rows, dtype=np.float64).reshape(len(rows), 16)
    np.savez_compressed(os.path.join(GOLD, "post_cases.npz"), **out)
    print("post cases", {k: int(out[k + ':ncontours'][0]) for k in maps})


def make_thresh_map_cases():
    """f-4: runs the reference's own draw_thresh_map (src/db_transforms.py:8-59) with the two absent geometry libraries
    replaced by stand-ins: shapely's Polygon -> shoelace area / perimeter, pyclipper's offset -> a dilated polygon supplied
    by the case (the arithmetic of Clipper is NOT what is pinned here; the distance field that follows it is)."""
    import types
    import sys
    ref_import.load()
    offsets = {}

    class FakeOffset:
        def AddPath(self, subject, jt, et):
            self.subject = tuple(map(tuple, subject))
        def Execute(self, distance):
            return [offsets[self.subject]]

    class FakePolygon:
        def __init__(self, pts):
            self.area, self.length = O.polygon_area_length(np.asarray(pts))

    pc = sys.modules["pyclipper"]
    pc.PyclipperOffset, pc.JT_ROUND, pc.ET_CLOSEDPOLYGON = FakeOffset, 0, 0
    sys.modules["shapely.geometry"].Polygon = FakePolygon
    sys.modules.setdefault("imgaug", types.ModuleType("imgaug"))
    import db_transforms as T
    from db_text_minimal_b200.postprocess import offset_convex_round
    rng = np.random.RandomState(9)
    polys = {
        "quad": np.array([[30, 20], [150, 28], [146, 60], [26, 52]]),
        "tall": np.array([[100, 70], [118, 72], [116, 125], [98, 122]]),
        "edge": np.array([[-6, 90], [40, 88], [42, 110], [-4, 112]]),           # dilated box leaves the canvas
        "corner": np.array([[140, 110], [170, 108], [172, 135], [138, 136]]),   # beyond the bottom-right corner
        "hex": np.array([[60, 100], [80, 92], [100, 100], [100, 120], [80, 128], [60, 120]]),
        "float": np.array([[10.5, 5.25], [60.75, 6.5], [59.5, 30.0], [9.25, 28.5]]),
    }
    H, W = 128, 160
    out = {}
    canvas = np.zeros((H, W), np.float32)
    mask = np.zeros((H, W), np.float32)
    names = list(polys)
    for name in names:
        poly = polys[name]
        area, length = O.polygon_area_length(poly)
        distance = area * (1 - 0.4 ** 2) / length
        padded = np.round(offset_convex_round(poly.astype(np.float64), distance)).astype(np.int64)
        offsets[tuple(map(tuple, poly))] = padded.tolist()
        T.draw_thresh_map(poly.copy(), canvas, mask, shrink_ratio=0.4)
        out[name + ":poly"] = poly.astype(np.float64)
        out[name + ":bbox"] = np.array([padded[:, 0].min(), padded[:, 1].min(), padded[:, 0].max(), padded[:, 1].max()], np.int64)
        out[name + ":distance"] = np.array([distance])
        out[name + ":canvas_after"] = canvas.copy()          # cumulative, in `names` order
    out["names"] = np.array(names)
    out["mask"] = mask
    np.savez_compressed(os.path.join(GOLD, "thresh_map_cases.npz"), **out)
    print("thresh map cases", names, float(canvas.max()), int((canvas > 0).sum()))


def make_metric_cases():
    """f-3: the reference's own cal_text_score / RunningScore (src/text_metrics.py:9-82), imported unmodified (its two
    unrelated imports -- iou (shapely) and utils (matplotlib) -- are stubbed), on seeded maps."""
    import types
    ref_import.load()
    for name, attr in (("iou", "DetectionIoUEvaluator"), ("utils", "to_list_tuples_coords")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, attr, object)
            sys.modules[name] = m
    import text_metrics as TM
    out = {}
    for ci, (n, h, w, seed) in enumerate([(2, 64, 64, 1), (3, 37, 53, 2), (4, 160, 160, 3)]):
        rng = np.random.RandomState(seed)
        P = rng.uniform(0, 1, (n, h, w)).astype(np.float32)
        P[0, :4, :4] = 0.25                       # exactly on the threshold ('<=' -> 0)
        gts = O.synth_gt_maps(n, h, w, seed)
        rs = TM.RunningScore(2)
        s1 = TM.cal_text_score(torch.from_numpy(P), torch.from_numpy(gts[0]), torch.from_numpy(gts[1]), rs, thresh=0.25)
        h1 = rs.confusion_matrix.copy()
        s2 = TM.cal_text_score(torch.from_numpy(P * 0.5), torch.from_numpy(gts[0]), torch.from_numpy(gts[1]), rs, thresh=0.25)
        out[f"c{ci}:meta"] = np.array([n, h, w, seed])
        out[f"c{ci}:hist1"] = h1
        out[f"c{ci}:hist2"] = rs.confusion_matrix.copy()
        keys = ["Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"]
        out[f"c{ci}:score1"] = np.array([s1[k] for k in keys])
        out[f"c{ci}:score2"] = np.array([s2[k] for k in keys])
    np.savez_compressed(os.path.join(GOLD, "metric_cases.npz"), **out)
    print("metric cases", out["c0:hist1"].tolist(), out["c0:score1"].tolist())


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    make_loss_cases()
    make_post_cases()
    make_thresh_map_cases()
    make_model_case("model_s0_64", 0, 2, 64, 64)
    make_model_case("model_s1_72x100", 1, 2, 72, 100)
    make_model_case("model_s2_54x70", 2, 1, 54, 70)
    make_cond_params()
    make_baseline_size_cases()
    make_metric_cases()
