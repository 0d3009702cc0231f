"""GPU parity of the post-processing front (binarize + CCL + float64 box score + score filter) against the fixtures
produced by the reference's own calls (cv2.findContours / box_score_fast) and against the contour-free CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import db_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["survey96", "nested80", "border64", "checker48", "many160", "holes64", "empty32", "full32", "blobs256",
         "blobs200x312", "noise128"]


def rep(**kw):
    from db_text_minimal_b200.postprocess import SegDetectorRepresenter
    return SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, unclip_ratio=1.5, **kw)


@pytest.mark.parametrize("case", CASES)
def test_front_matches_reference_contour_set(case):
    z = np.load(os.path.join(GOLD, "post_cases.npz"))
    P = z[case + ":P"]
    r = rep(max_candidates=100000)
    bitmap, labels, rec, nc = r.front(torch.from_numpy(P)[None, None].cuda(), want_labels=True)
    assert np.array_equal(bitmap[0].cpu().numpy().astype(bool), z[case + ":bitmap"])            # bit-exact
    ncont = int(z[case + ":ncontours"][0])
    assert int(nc[0]) == ncont                                                                  # candidate count: exact
    ref = z[case + ":cands"]            # rows: score, sside, keep, count, x0, y0, x1, y1, box pts (first 1000, cv2 order)
    got = rec[0]
    assert len(got) == ncont
    gs = sorted((int(g["count"]), (int(g["x0"]), int(g["y0"]), int(g["x1"]), int(g["y1"])), float(g["sum"]) / int(g["count"])) for g in got)
    if ncont <= 1000:
        ws = sorted((int(r_[3]), tuple(int(v) for v in r_[4:8]), r_[0]) for r_ in ref)
        for g, w in zip(gs, ws):
            assert g[0] == w[0] and g[1] == w[1], (g, w)                                        # integer work: exact
            assert abs(g[2] - w[2]) <= 1e-12 * max(1.0, abs(w[2]))                               # float64 mean
        # the score filter (postprocess.py:129): the kept SET is equal (bit-exact: the comparison is in float64 on both sides)
        kept_g = sorted((int(g["count"]), (int(g["x0"]), int(g["y0"]), int(g["x1"]), int(g["y1"]))) for g in got if g["keep"])
        kept_w = sorted((int(r_[3]), tuple(int(v) for v in r_[4:8])) for r_ in ref if not (0.5 > r_[0]))
        assert kept_g == kept_w
    # labels: foreground positive, background negative, consistent with the bitmap
    lab = labels[0].cpu().numpy()
    assert ((lab > 0) == z[case + ":bitmap"]).all()


@pytest.mark.parametrize("case", ["survey96", "nested80", "border64", "holes64", "blobs256", "blobs200x312"])
def test_front_order_is_cv2_order(case):
    """Candidates come back in the reference's contour order (reverse raster discovery), so contours[:max_candidates]
    truncation and output order match (thin one-pixel structures aside)."""
    z = np.load(os.path.join(GOLD, "post_cases.npz"))
    P = z[case + ":P"]
    got = rep(max_candidates=1000).candidates(torch.from_numpy(P)[None, None].cuda())[0]
    ref = z[case + ":cands"]
    assert [(c["count"], c["bbox"]) for c in got] == [(int(r_[3]), tuple(int(v) for v in r_[4:8])) for r_ in ref]
    assert [c["keep"] for c in got] == [not (0.5 > r_[0]) for r_ in ref]


def test_front_matches_oracle_on_batch_and_sizes():
    """Batched call, ragged sizes, vs the contour-free CPU restatement (exact counts / bboxes, float64 sums)."""
    maps = [O.synth_prob_map(96, 130, s) for s in range(5)]
    P = torch.from_numpy(np.stack(maps))[:, None].cuda()
    cands = rep(max_candidates=5000).candidates(P)
    for i, m in enumerate(maps):
        _, want = O.candidates_ccl(m, 0.25)
        g = sorted((c["kind"], c["count"], c["bbox"], c["first"]) for c in cands[i])
        w = sorted((c["kind"], c["count"], c["bbox"], c["first"]) for c in want)
        assert g == w
        gs = sorted(c["sum"] for c in cands[i]); ws = sorted(c["sum"] for c in want)
        np.testing.assert_allclose(gs, ws, rtol=1e-12)


@pytest.mark.parametrize("shape", [(70, 130), (33, 97), (16, 64), (17, 31), (48, 1030), (9, 2100)])
def test_front_paths_agree_with_the_oracle_on_odd_shapes(shape, monkeypatch):
    """Both labelling paths -- strips in shared memory (default) and the three global-memory kernels (rows too wide for
    shared memory; forced here with DBB_CCL_NO_STRIP) -- against the contour-free oracle, on widths that are not multiples
    of 4 / 32 and heights that are not multiples of the strip height."""
    h, w = shape
    maps = [O.synth_prob_map(h, w, 40 + s) for s in range(3)]
    P = torch.from_numpy(np.stack(maps))[:, None].cuda()
    want = [O.candidates_ccl(m, 0.25)[1] for m in maps]
    for no_strip in (False, True):
        if no_strip:
            monkeypatch.setenv("DBB_CCL_NO_STRIP", "1")
        got = rep(max_candidates=5000).candidates(P)
        for i in range(len(maps)):
            g = sorted((c["kind"], c["count"], c["bbox"], c["first"]) for c in got[i])
            wv = sorted((c["kind"], c["count"], c["bbox"], c["first"]) for c in want[i])
            assert g == wv, (shape, no_strip, i)
            np.testing.assert_allclose(sorted(c["sum"] for c in got[i]), sorted(c["sum"] for c in want[i]), rtol=1e-12)


def test_front_on_rows_too_wide_for_the_strip_kernel():
    """w = 4,000: 16 rows of labels do not fit in shared memory, the front takes the three-kernel path by itself; it must
    agree with the strip path run on the same content (two halves side by side never touch: a blank column band between)."""
    h, w = 40, 4000
    m = np.zeros((h, w), np.float32)
    a = O.synth_prob_map(h, 1900, 7)
    m[:, :1900] = a
    m[:, 2100:] = a
    got = rep(max_candidates=20000).candidates(torch.from_numpy(m)[None, None].cuda())[0]
    half = rep(max_candidates=20000).candidates(torch.from_numpy(np.pad(a, ((0, 0), (0, 100))))[None, None].cuda())[0]
    # the outer frame background is one region in both; every other region of the half map appears twice in the wide map
    key = lambda c, dx=0: (c["kind"], c["count"], (c["bbox"][0] - dx, c["bbox"][1], c["bbox"][2] - dx, c["bbox"][3]))
    left = sorted(key(c) for c in got if c["bbox"][2] < 2000)
    right = sorted(key(c, 2100) for c in got if c["bbox"][0] >= 2100)
    ref = sorted(key(c) for c in half if c["bbox"][2] < 1900)
    assert left == ref and right == ref and len(ref) > 10


def test_front_with_more_strips_than_resident_blocks():
    """640 maps of 256 x 32: 10,240 strips of 16 rows, more than the strip kernel's grid cap (148 x 64 blocks), so its
    blocks loop over several strips and reuse their shared-memory arrays; every image must come out as when run alone."""
    base = [O.synth_prob_map(256, 32, 60 + s) for s in range(8)]
    n = 640
    P = torch.from_numpy(np.stack([base[i % 8] for i in range(n)]))[:, None].cuda()
    r = rep(max_candidates=5000)
    _, _, rec, nc = r.front(P)
    single = [r.front(torch.from_numpy(b)[None, None].cuda()) for b in base]
    for i in list(range(16)) + [n // 2, n - 9, n - 1]:
        _, _, rec1, nc1 = single[i % 8]
        assert int(nc[i]) == int(nc1[0])
        a, b = rec[i][:int(nc[i])], rec1[0][:int(nc1[0])]
        for f in ("kind", "first_y", "first_x", "x0", "y0", "x1", "y1", "count", "keep"):
            assert np.array_equal(a[f], b[f]), (i, f)
        np.testing.assert_allclose(a["sum"], b["sum"], rtol=1e-12)


def test_binarize_is_strict_and_exact():
    r = rep()
    p = torch.tensor([[0.25, 0.2500001, 0.24999999, 0.3, 0.0, 1.0]]).cuda()
    assert r.binarize(p).cpu().tolist() == [[False, True, False, True, False, True]]
    x = torch.rand(3, 37, 53)
    assert torch.equal(r.binarize(x.cuda()).cpu(), x > 0.25)


def test_full_size_properties():
    """BASELINE config 4 size (1024 x 1024 maps): size-independent properties."""
    n = 4
    maps = np.stack([O.synth_prob_map(1024, 1024, 100 + s) for s in range(n)])
    maps = (maps - 0.25).clip(0) / 0.75           # ~100 separate blobs per map instead of one sheet
    P = torch.from_numpy(maps)[:, None].cuda()
    r = rep(max_candidates=100000)
    bitmap, labels, rec, nc = r.front(P, want_labels=True)
    bm = bitmap.cpu().numpy().astype(bool)
    assert np.array_equal(bm, maps > np.float32(0.25))
    from scipy import ndimage
    for i in range(n):
        nf = ndimage.label(bm[i], structure=np.ones((3, 3), int))[1]
        ri = rec[i][:int(nc[i])]
        outer = ri[ri["kind"] == 0]
        assert len(outer) == nf                                           # one outer candidate per 8-connected component
        # outer fill sets partition: total own-foreground pixel count is the bitmap sum (checksum of checksums)
        lab = labels[i].cpu().numpy()
        assert (lab > 0).sum() == bm[i].sum()
        # idempotence
    _, _, rec2, nc2 = r.front(P)
    assert np.array_equal(nc, nc2)
    for i in range(n):          # integer fields bit-exact; the float64 sums come from atomics (order varies in the last bits)
        a, b = rec[i][:int(nc[i])], rec2[i][:int(nc[i])]
        for f in ("kind", "first_y", "first_x", "x0", "y0", "x1", "y1", "count", "keep"):
            assert np.array_equal(a[f], b[f]), f
        np.testing.assert_allclose(a["sum"], b["sum"], rtol=1e-13)


def load_case(case):
    """(P, bitmap, rows) of a golden case, or -- 'kept:<h>x<w>:<seed>' -- of a blob map whose components score above
    box_thresh (the golden blob maps have none), rows computed here by the oracle's own OpenCV calls in the golden layout
    [score, sside, keep, count, x0, y0, x1, y1, 4 corner points]."""
    if not case.startswith("kept:"):
        z = np.load(os.path.join(GOLD, "post_cases.npz"))
        return z[case + ":P"], z[case + ":bitmap"], z[case + ":cands"]
    import cv2
    _, hw, seed = case.split(":")
    h, w = (int(v) for v in hw.split("x"))
    P = ((O.synth_prob_map(h, w, int(seed)) - 0.45) * 8).clip(0, 1).astype(np.float32)
    bitmap = O.binarize(P, 0.25)
    rows = []
    for contour in O.candidates_cv2(bitmap)[:1000]:
        c = contour.squeeze(1) if len(contour) > 1 else contour.reshape(-1, 2)
        pts, sside = O.get_mini_boxes(c)
        score = O.box_score_fast(P, c)
        x0, y0, x1, y1 = c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()
        m = np.zeros((y1 - y0 + 1, x1 - x0 + 1), np.uint8)
        cv2.fillPoly(m, (c - [x0, y0]).reshape(1, -1, 2).astype(np.int32), 1)
        rows.append([score, sside, float(sside >= 3 and not (0.5 > score)), float(m.sum()), x0, y0, x1, y1] + list(np.array(pts).reshape(-1)))
    return P, bitmap, np.array(rows, np.float64).reshape(len(rows), 16)


BACK_CASES = ["survey96", "nested80", "border64", "holes64", "full32", "checker48", "noise128", "kept:256x320:11", "kept:200x312:12", "kept:640x640:13"]


def _ref_back_half(r, contour, width, height, dest_w, dest_h):
    """src/postprocess.py:121-147 for one contour with OpenCV doing the geometry (cv2.minAreaRect / boxPoints) and the
    reference's arithmetic around it; the offset is the product's ClipperOffset restatement (unpinned).

    Returns (outcomes, sside, points): `outcomes` is the list of results the reference flow can produce -- each None
    (dropped) or (int16 box, float32 pre-rounding coordinates).  More than one entry appears only when a float32 corner of
    the first min-area box lies within 3e-4 of an integer: pyclipper TRUNCATES those corners to integers
    (src/postprocess.py:152), so the reference's own result flips with the last bits of OpenCV's float arithmetic there
    (OpenCV's build does not round like strict float32: the restated calipers agree with it to ~1e-5, not bit for bit)."""
    import itertools
    import cv2
    from db_text_minimal_b200.postprocess import clipper_offset

    def mini(c):
        bb = cv2.minAreaRect(c)
        pts = sorted(list(cv2.boxPoints(bb)), key=lambda x: x[0])
        i1, i4 = (0, 1) if pts[1][1] > pts[0][1] else (1, 0)
        i2, i3 = (2, 3) if pts[3][1] > pts[2][1] else (3, 2)
        return [pts[i1], pts[i2], pts[i3], pts[i4]], min(bb[1])
    points, sside = mini(contour)
    if sside < r.min_size:
        return [None], sside, points
    points = np.array(points)
    x, y = points[:, 0].astype(np.float64), points[:, 1].astype(np.float64)
    area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    length = np.sqrt((x - np.roll(x, -1)) ** 2 + (y - np.roll(y, -1)) ** 2).sum()
    flat = points.reshape(-1).astype(np.float64)
    # (the unperturbed truncation first: outcomes[0] is what the reference computes with OpenCV's own floats)
    choices = [[int(np.trunc(v))] + sorted({int(np.trunc(v - 3e-4)), int(np.trunc(v + 3e-4))} - {int(np.trunc(v))}) for v in flat]
    outcomes = []
    for combo in itertools.islice(itertools.product(*choices), 256):
        out = clipper_offset(np.array(combo, np.float64).reshape(4, 2), area * r.unclip_ratio / length)
        if len(out) != 1:
            outcomes.append(None); continue
        box, ss2 = mini(out[0].astype(np.int32).reshape(-1, 1, 2))
        if ss2 < r.min_size + 2:
            outcomes.append(None); continue
        box = np.array(box)
        fl = np.stack([box[:, 0] / width * dest_w, box[:, 1] / height * dest_h], 1)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_w), 0, dest_w)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_h), 0, dest_h)
        outcomes.append((box.astype(np.int16), fl))
    return outcomes, sside, points


def _box_matches(got, outcomes, dest_w, dest_h):
    """got (4, 2) int16 (all zero = dropped) against the possible reference outcomes; a coordinate may also sit on the other
    side of a .5 rounding boundary of the float32 rescale (|got - unrounded| <= 0.5 + 1e-3).  Returns 'exact' / 'boundary' /
    None."""
    dropped = not got.any()
    verdict = None
    for i, o in enumerate(outcomes):
        if o is None:
            if dropped:
                return "exact" if i == 0 else "boundary"
            continue
        if dropped:
            continue
        box, fl = o
        if np.array_equal(got, box):
            return "exact" if i == 0 else "boundary"
        lim = np.clip(fl, 0, [dest_w, dest_h])
        if (np.abs(got.astype(np.float64) - lim) <= 0.5 + 1e-3).all():
            verdict = "boundary"
    return verdict


@pytest.mark.parametrize("case", BACK_CASES)
def test_box_mode_back_half_matches_reference_rows(case):
    """a-14 / a-15: get_mini_boxes of every kept candidate (sside and the four ordered corners) against the rows the
    reference's own get_mini_boxes produced (tests/golden/post_cases.npz), then the rest of boxes_from_bitmap against the
    same steps driven by OpenCV.  The contour never exists here: the device emits run end points, C++ does the rest."""
    import cv2
    P, bitmap, ref = load_case(case)
    h, w = P.shape
    r = rep(max_candidates=1000)
    dest_w, dest_h = 2 * w + 3, h + 7           # a non-trivial rescale
    boxes, scores, nc, sside, mini, _ = r.boxes_batch(torch.from_numpy(P)[None, None].cuda(), [(dest_w, dest_h)], debug=True)
    k = min(int(nc[0]), 1000)
    assert k == len(ref)
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    n_kept = n_boundary = 0
    for i in range(k):
        score_keep = not (0.5 > ref[i][0])
        if not score_keep:
            assert not boxes[0, i].any() and scores[0, i] == 0
            continue
        n_kept += 1
        assert abs(sside[0, i] - ref[i][1]) <= 1e-3 * max(1.0, ref[i][1]), (i, sside[0, i], ref[i][1])
        np.testing.assert_allclose(mini[0, i].reshape(-1), ref[i][8:16], atol=1e-3, err_msg=f"candidate {i}")
        outcomes, _, _ = _ref_back_half(r, contours[i].squeeze(1) if len(contours[i]) > 1 else contours[i].reshape(-1, 2), w, h, dest_w, dest_h)
        verdict = _box_matches(boxes[0, i], outcomes, dest_w, dest_h)
        assert verdict is not None, (i, boxes[0, i].tolist(), [None if o is None else o[0].tolist() for o in outcomes[:4]])
        n_boundary += verdict == "boundary"
        if boxes[0, i].any():
            assert abs(scores[0, i] - np.float32(ref[i][0])) < 1e-6
        else:
            assert scores[0, i] == 0
    nfinal = int(boxes[0].reshape(k, -1).any(1).sum())
    print(case, "kept by score:", n_kept, "of", k, "| final boxes:", nfinal, "| decided by a truncation / rounding boundary:", n_boundary)
    if case.startswith("kept:"):
        assert nfinal >= 5 and n_boundary <= max(2, nfinal // 4)


def test_boxes_end_to_end_box_mode():
    """__call__ in box mode: same kept candidates as the reference's steps 1-2, boxes well-formed (unclip unpinned)."""
    z = np.load(os.path.join(GOLD, "post_cases.npz"))
    P = z["survey96:P"]
    r = rep()
    boxes, scores = r({"shape": [(96, 96)]}, torch.from_numpy(P)[None, None].cuda(), is_output_polygon=False)
    ref = z["survey96:cands"]
    assert boxes[0].shape == (len(ref), 4, 2) and boxes[0].dtype == np.int16 and scores[0].dtype == np.float32
    kept_ref = [i for i, r_ in enumerate(ref) if r_[2] > 0]
    kept_got = [i for i in range(len(ref)) if boxes[0][i].any()]
    assert kept_got == kept_ref              # here every survivor of steps 1-2 also survives the post-unclip size test
    for i in kept_ref:
        assert abs(scores[0][i] - np.float32(ref[i][0])) < 1e-6


@pytest.mark.parametrize("case", ["survey96", "nested80", "holes64", "kept:256x320:11", "kept:200x312:12", "kept:640x640:13"])
def test_polygon_mode_runs_and_matches_the_reference_steps_unpinned_offset(case):
    """a-16: is_output_polygon=True (the reference's shipped default, example_config.yaml:83) without pyclipper.  The flow of
    src/postprocess.py:54-104 is re-run here with OpenCV on the golden bitmap (findContours, arcLength, approxPolyDP,
    box_score_fast) and compared polygon by polygon; the offset on both sides is the product's ClipperOffset restatement
    (UNPINNED: no pyclipper to generate goldens), so this pins everything AROUND the offset: candidate order, the < 4 points
    drop, the score filter, the len(box) > 1 drop, the sside filter and the rescale."""
    import cv2
    from db_text_minimal_b200.postprocess import clipper_offset
    P, bitmap, rows = load_case(case)
    h, w = P.shape
    r = rep(max_candidates=1000)
    dest_w, dest_h = w + 11, 2 * h
    boxes, scores = r({"shape": [(dest_h, dest_w)]}, torch.from_numpy(P)[None, None].cuda(), is_output_polygon=True)
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    want_b, want_s = [], []
    for contour, row in zip(contours[:1000], rows):
        approx = cv2.approxPolyDP(contour, 0.005 * cv2.arcLength(contour, True), True)
        points = approx.reshape((-1, 2))
        if points.shape[0] < 4 or 0.5 > row[0]:
            continue
        x, y = points[:, 0].astype(np.float64), points[:, 1].astype(np.float64)
        area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        length = np.sqrt((x - np.roll(x, -1)) ** 2 + (y - np.roll(y, -1)) ** 2).sum()
        out = clipper_offset(points, area * 1.5 / length)
        if len(out) != 1:
            continue
        box = out[0].reshape(-1, 2)
        if min(cv2.minAreaRect(box.reshape(-1, 1, 2).astype(np.int32))[1]) < 5:
            continue
        box = box.astype(np.float64)
        box[:, 0] = np.clip(np.round(box[:, 0] / w * dest_w), 0, dest_w)
        box[:, 1] = np.clip(np.round(box[:, 1] / h * dest_h), 0, dest_h)
        want_b.append(box); want_s.append(row[0])
    assert len(boxes[0]) == len(want_b) and len(want_b) >= 1
    for g, wv, gs, ws in zip(boxes[0], want_b, scores[0], want_s):
        assert np.array_equal(np.asarray(g), wv)
        assert abs(gs - ws) <= 1e-12
