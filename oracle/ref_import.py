"""Import the UNMODIFIED reference (/root/reference/src) for golden generation.

TEST INFRASTRUCTURE.  Only usable where /root/reference exists (the build container);
the GPU box never sees it -- tests there use the committed fixtures in tests/golden/.

Stubs needed:
  * modules.resnet.model_zoo.load_url -> {}   (pretrained=True is hard-coded; no network)
  * dummy ``pyclipper`` / ``shapely.geometry.Polygon`` modules (absent wheels)
  * ``np.int = int`` (removed in NumPy >= 1.24; used at src/postprocess.py:189-192)
"""
import os
import sys
import types

REF_SRC = "/root/reference/src"
if not os.path.isdir(REF_SRC):      # GPU box: the copy staged by oracle/stage_ref.py (bench.py's reference legs only)
    _staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")
    if os.path.isfile(os.path.join(_staged, "models.py")):
        REF_SRC = _staged


def available() -> bool:
    return os.path.isdir(REF_SRC)


def load():
    """Returns (models, losses, postprocess) reference modules."""
    if not available():
        raise RuntimeError("reference not present at " + REF_SRC)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # noqa
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    if "pyclipper" not in sys.modules:
        sys.modules["pyclipper"] = types.ModuleType("pyclipper")
    if "shapely" not in sys.modules:
        sh = types.ModuleType("shapely")
        geo = types.ModuleType("shapely.geometry")
        geo.Polygon = object
        sh.geometry = geo
        sys.modules["shapely"] = sh
        sys.modules["shapely.geometry"] = geo
    import modules.resnet as R
    R.model_zoo.load_url = lambda *a, **k: {}
    import contextlib
    import io
    import models
    import losses
    import postprocess
    return models, losses, postprocess


def build_model(params=None):
    """Reference DBTextModel, optionally loaded (strict) with an oracle-style param dict."""
    import contextlib
    import io
    models, _, _ = load()
    with contextlib.redirect_stdout(io.StringIO()):   # 'load from imagenet' print
        m = models.DBTextModel()
    if params is not None:
        m.load_state_dict(params, strict=True)
    return m
