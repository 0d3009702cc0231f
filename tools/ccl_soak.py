"""Randomised soak of the post-processing front against the contour-free CPU oracle (one-off confidence run for `gpurun`,
not part of the test suite):  python tools/ccl_soak.py [cases] [seed]
Random shapes (1..220 pixels per side), random textures (blurred noise at several scales, salt-and-pepper, stripes,
checkerboards, frames), random thresholds; both labelling paths (strips in shared memory / three global-memory kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import db_oracle as O
from db_text_minimal_b200.postprocess import SegDetectorRepresenter

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def texture(h, w):
    kind = rng.integers(0, 6)
    if kind == 0:
        return O.synth_prob_map(h, w, int(rng.integers(0, 10**6)))
    if kind == 1:
        return rng.random((h, w)).astype(np.float32)
    if kind == 2:
        m = np.zeros((h, w), np.float32); p = int(rng.integers(1, 5)); m[::p] = 1.0
        return m
    if kind == 3:
        yy, xx = np.mgrid[0:h, 0:w]; p = int(rng.integers(1, 4))
        return (((yy // p) + (xx // p)) % 2).astype(np.float32)
    if kind == 4:
        m = np.zeros((h, w), np.float32)
        for _ in range(int(rng.integers(1, 6))):
            y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w)); y1, x1 = int(rng.integers(y0, h)) + 1, int(rng.integers(x0, w)) + 1
            m[y0:y1, x0:x1] = 1.0
            if y1 - y0 > 2 and x1 - x0 > 2 and rng.random() < 0.7:
                m[y0 + 1:y1 - 1, x0 + 1:x1 - 1] = 0.0
        return m
    from scipy import ndimage
    return (ndimage.gaussian_filter(rng.random((h, w)), rng.uniform(0.5, 4.0)) * 2).clip(0, 1).astype(np.float32)


bad = 0
for ci in range(cases):
    h, w = int(rng.integers(1, 221)), int(rng.integers(1, 221))
    if ci % 8 == 7:                    # rows longer than one 1,024-pixel pass of pack_row (run starts carried across passes)
        h, w = int(rng.integers(1, 40)), int(rng.integers(1025, 2700))
    n = int(rng.integers(1, 5))
    th = float(rng.choice([0.25, 0.3, 0.5, 0.45]))
    maps = [np.ascontiguousarray(texture(h, w), dtype=np.float32) for _ in range(n)]
    P = torch.from_numpy(np.stack(maps))[:, None].cuda()
    want = [O.candidates_ccl(m, th)[1] for m in maps]
    for no_strip in (False, True):
        if no_strip:
            os.environ["DBB_CCL_NO_STRIP"] = "1"
        else:
            os.environ.pop("DBB_CCL_NO_STRIP", None)
        got = SegDetectorRepresenter(thresh=th, box_thresh=0.5, max_candidates=10**6).candidates(P)
        for i in range(n):
            g = sorted((c["kind"], c["count"], c["bbox"], c["first"]) for c in got[i])
            wv = sorted((c["kind"], c["count"], c["bbox"], c["first"]) for c in want[i])
            ok = g == wv and np.allclose(sorted(c["sum"] for c in got[i]), sorted(c["sum"] for c in want[i]), rtol=1e-11, atol=1e-9)
            if not ok:
                bad += 1
                print("MISMATCH case", ci, "shape", (n, h, w), "thresh", th, "no_strip", no_strip, "image", i, len(g), len(wv), flush=True)
                np.save(f"gpurun_out/soak_bad_{ci}_{i}.npy", maps[i])
    if ci % 50 == 49:
        print("done", ci + 1, "bad", bad, flush=True)
print("SOAK cases", cases, "bad", bad)
