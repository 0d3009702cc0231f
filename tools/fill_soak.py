"""One-off soak of dbb_fill_polygons against cv2.fillPoly: many more random polygons, canvas sizes and seeds than the suite."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cv2
from test_gt_maps_gpu import _rand_polys
from db_text_minimal_b200.db_transforms import fill_polygons
tot = bad = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    rng = np.random.RandomState(1000 + seed)
    h, w = int(rng.randint(8, 300)), int(rng.randint(8, 300))
    size = max(h, w)
    polys = _rand_polys(rng, 300, size)
    maps = torch.zeros((len(polys), h, w), dtype=torch.float32, device="cuda")
    fill_polygons(maps, polys, list(range(len(polys))), [1.0] * len(polys))
    got = maps.cpu().numpy()
    for i, p in enumerate(polys):
        ref = np.zeros((h, w), np.float32)
        cv2.fillPoly(ref, [p.astype(np.int32)], 1.0)
        tot += 1
        if not np.array_equal(got[i], ref):
            bad += 1
            if bad < 5: print("MISMATCH", (h, w), p.tolist(), int((got[i] != ref).sum()), flush=True)
print("FILL SOAK polygons", tot, "mismatches", bad)
