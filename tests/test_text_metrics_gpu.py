"""f-3: per-step pixel metric on the device vs the reference's numpy formulation (src/text_metrics.py:9-82)."""
import numpy as np
import pytest
import torch

from oracle import db_oracle as O

pytestmark = pytest.mark.gpu


def ref_hist(texts, gt, mask, thresh):
    """src/text_metrics.py:72-79 + :14-23, verbatim arithmetic in numpy."""
    pred = texts * mask
    pred = np.where(pred <= np.float32(thresh), 0, 1).astype(np.int32)
    g = (gt * mask).astype(np.int32)
    out = np.zeros((2, 2))
    for lt, lp in zip(g, pred):
        lt, lp = lt.flatten(), lp.flatten()
        m = (lt >= 0) & (lt < 2)
        out += np.bincount(2 * lt[m].astype(int) + lp[m], minlength=4).reshape(2, 2)
    return out


@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 37, 53), (16, 640, 640)])
def test_confusion_matrix_is_exact(shape):
    from db_text_minimal_b200.text_metrics import RunningScore, cal_text_score
    n, h, w = shape
    rng = np.random.RandomState(h)
    preds = rng.uniform(0, 1, (n, 3, h, w)).astype(np.float32)
    preds[0, 0, :4, :4] = 0.25                                        # exactly on the threshold: '<=' -> 0
    gts = O.synth_gt_maps(n, h, w, 2)
    want = ref_hist(preds[:, 0], gts[0], gts[1], 0.25)
    rs = RunningScore(2)
    p = torch.from_numpy(preds).cuda()
    g = torch.from_numpy(gts).cuda()
    score = cal_text_score(p[:, 0], g[0], g[1], rs, thresh=0.25)       # channel view, no copy
    assert np.array_equal(rs.confusion_matrix, want)                   # integer work: bit-exact
    hist = want
    acc = np.diag(hist).sum() / (hist.sum() + 0.0001)
    iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist) + 0.0001)
    assert score['Overall Acc'] == acc and score['Mean IoU'] == np.nanmean(iu)
    cal_text_score(p[:, 0], g[0], g[1], rs, thresh=0.25)               # accumulates like RunningScore.update
    assert np.array_equal(rs.confusion_matrix, 2 * want)


def test_matches_the_reference_function_golden():
    """tests/golden/metric_cases.npz comes from the UNMODIFIED src/text_metrics.py (cal_text_score + RunningScore), two
    accumulating calls per case: confusion matrices bit-exact, the four scores equal to the last bit of float64."""
    import os
    from db_text_minimal_b200.text_metrics import RunningScore, cal_text_score
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "metric_cases.npz"))
    keys = ["Overall Acc", "Mean Acc", "FreqW Acc", "Mean IoU"]
    for ci in range(3):
        n, h, w, seed = [int(v) for v in z[f"c{ci}:meta"]]
        rng = np.random.RandomState(seed)
        P = rng.uniform(0, 1, (n, h, w)).astype(np.float32)
        P[0, :4, :4] = 0.25
        gts = O.synth_gt_maps(n, h, w, seed)
        assert np.array_equal(ref_hist(P, gts[0], gts[1], 0.25), z[f"c{ci}:hist1"])      # the numpy restatement used above is pinned too
        rs = RunningScore(2)
        g = torch.from_numpy(gts).cuda()
        s1 = cal_text_score(torch.from_numpy(P).cuda(), g[0], g[1], rs, thresh=0.25)
        assert np.array_equal(rs.confusion_matrix, z[f"c{ci}:hist1"])
        assert [s1[k] for k in keys] == z[f"c{ci}:score1"].tolist()
        s2 = cal_text_score(torch.from_numpy(P * 0.5).cuda(), g[0], g[1], rs, thresh=0.25)
        assert np.array_equal(rs.confusion_matrix, z[f"c{ci}:hist2"])
        assert [s2[k] for k in keys] == z[f"c{ci}:score2"].tolist()
