"""One-off soak of the box-mode back half (C++ minAreaRect / unclip / rescale) against the OpenCV-driven reference steps of
tests/test_postprocess_gpu.py on many synthetic maps:  python tools/box_soak.py [cases]"""
import os, sys, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_postprocess_gpu as T

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(3)
bad = kept = boundary = 0
for i in range(cases):
    h, w = int(rng.integers(96, 700)), int(rng.integers(96, 700))
    case = f"kept:{h}x{w}:{1000 + i}"
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            T.test_box_mode_back_half_matches_reference_rows(case)
        line = buf.getvalue().strip().splitlines()[-1]
        kept += int(line.split("final boxes:")[1].split("|")[0]); boundary += int(line.rsplit(":", 1)[1])
    except AssertionError as e:
        import traceback
        tb = traceback.extract_tb(e.__traceback__)[-1]
        if "nfinal >= 5" in (tb.line or ""):            # few boxes on this map / the suite's bound on boundary cases: report, not a failure
            print("note", case, buf.getvalue().strip().splitlines()[-1] if buf.getvalue().strip() else "", flush=True)
            continue
        bad += 1
        print("FAIL", case, "line", tb.lineno, tb.line, str(e)[:300], flush=True)
print("BOX SOAK cases", cases, "failed", bad, "final boxes", kept, "decided by a truncation / rounding boundary", boundary)
