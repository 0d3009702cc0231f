"""One-off soak of DBLoss (fused forward + backward, OHEM radix select) against the CPU oracle on random shapes, positive
fractions, mask patterns and saturated predictions -- reuses the checks of tests/test_loss_gpu.py."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_loss_gpu as T
from oracle import db_oracle as O

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.RandomState(77)
bad = 0
for ci in range(cases):
    n, h, w = int(rng.randint(1, 5)), int(rng.randint(4, 200)), int(rng.randint(4, 200))
    preds = rng.uniform(0.001, 0.999, (n, 3, h, w)).astype(np.float32)
    mode = ci % 5
    if mode == 1: preds[:, 0] = np.where(rng.rand(n, h, w) < 0.2, rng.choice([0.0, 1.0], (n, h, w)), preds[:, 0]).astype(np.float32)   # saturated P (BCE clamp)
    if mode == 2: preds[:, 0] = np.round(preds[:, 0] * 8) / 8                                                               # many equal losses: ties around tau
    preds[:, 2] = 1.0 / (1.0 + np.exp(-50.0 * (preds[:, 0].astype(np.float64) - preds[:, 1])))
    gts = O.synth_gt_maps(n, h, w, seed=1000 + ci)
    if mode == 3: gts[0][:] = 0                                   # no positives
    if mode == 4: gts[1][:] = (rng.rand(n, h, w) < 0.5)           # half masked
    for red in ("mean", "none"):
        try:
            vals, grad, st = T.run_gpu(preds, gts, red)
            orc = O.db_loss(preds, gts, reduction=red)
            assert (st.n_pos, st.n_neg) == (orc["n_pos"], orc["n_neg"]), ("counts", st.n_pos, st.n_neg, orc["n_pos"], orc["n_neg"])
            np.testing.assert_allclose(vals, orc["losses"], rtol=1e-5, atol=1e-6)
            g = orc["grad"]; diff = np.abs(grad - g)
            if red == "none":
                near = T.near_tau_mask(preds, gts, orc["tau"])
                if orc["n_neg"] > 0 and not (st.tau == 0 and orc["tau"] == 0): assert T.ulp_dist(st.tau, orc["tau"]) <= 2, ("tau", st.tau, orc["tau"])   # (+0.0 vs -0.0: the same threshold)
                mism = ((grad[:, 0] != 0) != (g[:, 0] != 0)) & ~near
                assert mism.sum() == 0, ("selected set", int(mism.sum()))
                diff[:, 0][near] = 0
            assert diff.max() <= 1e-5 * (np.abs(g).max() + 1e-30), ("grad", float(diff.max()))
        except AssertionError as e:
            bad += 1
            print("FAIL case", ci, (n, h, w), "mode", mode, red, str(e)[:200], flush=True)
print("LOSS SOAK cases", cases, "x 2 reductions, failed", bad)
