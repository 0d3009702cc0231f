"""One-off soak of polygon mode (C++ border following / approxPolyDP / offset / filters) against the OpenCV-driven reference
steps of tests/test_postprocess_gpu.py on many synthetic maps:  python tools/poly_soak.py [cases]"""
import os, sys, io, contextlib, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_postprocess_gpu as T

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(5)
bad = 0
for i in range(cases):
    h, w = int(rng.integers(96, 700)), int(rng.integers(96, 700))
    case = f"kept:{h}x{w}:{2000 + i}"
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            T.test_polygon_mode_runs_and_matches_the_reference_steps_unpinned_offset(case)
    except AssertionError as e:
        tb = traceback.extract_tb(e.__traceback__)[-1]
        bad += 1
        print("FAIL", case, "line", tb.lineno, tb.line, str(e)[:300], flush=True)
print("POLY SOAK cases", cases, "failed", bad)
