"""CPU-side checks: the C-ABI library loads without a GPU and exports exactly what include/dbb200.h declares; host
logic (sharding, gradient-bucket layout, Clipper-offset restatement, synthetic generators); gloo world_size-2 all-reduce."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from db_text_minimal_b200 import _build, _lib
    _build.build()
    return _lib.lib()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dbb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dbb_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "db_text_minimal_b200", "libdbb200.so")],
                         capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (dbb_[a-z0-9_]+)", out)))
    undeclared = [s for s in exported if s not in syms]
    assert not undeclared, undeclared


def test_no_compute_metadata_calls(lib):
    assert lib.dbb_version() >= 100
    assert lib.dbb_strerror(0) == b"ok" and b"workspace" in lib.dbb_strerror(-3)
    assert lib.dbb_net_num_params() == 115 and lib.dbb_net_num_buffers() == 64
    names = [lib.dbb_net_param_name(i).decode() for i in range(lib.dbb_net_num_params())]
    assert "segmentation_head.binarize.0.bias" in names and "segmentation_head.thresh.0.bias" not in names
    assert sum(lib.dbb_net_param_numel(i) for i in range(len(names))) == 13318474 - 2 * sum(
        lib.dbb_net_buffer_numel(i) for i in range(0, 64, 2)) - 32      # minus running stats and num_batches_tracked
    assert lib.dbb_dbloss_workspace(16, 3, 640, 640, 1) > 16 * 640 * 640 * 4
    assert lib.dbb_postprocess_workspace(1, 64, 64) > 64 * 64 * 4
    assert lib.dbb_net_num_segments() == 3
    # bad arguments fail loudly, without touching a device
    assert lib.dbb_dbloss_fwd(None, None, 1, 3, 8, 8, 1.0, 10.0, 0, 3.0, 1e-6, None, None, None, 0, None) == -1


def test_model_tree_matches_executor_tables(lib):
    from db_text_minimal_b200.models import DBTextModel
    m = DBTextModel(pretrained=False)
    assert len(m.state_dict()) == 211
    assert m._flat_numel >= 12269378 and m._segment_slices[0][0] == 0 and m._segment_slices[2][1] == m._flat_numel
    used = sum(p.numel() for k, p in m.named_parameters() if not (k.startswith("backbone.fc") or k.startswith("backbone.smooth")))
    assert used == 12269378


def test_product_refuses_cpu_tensors(lib):
    from db_text_minimal_b200 import DbbError
    from db_text_minimal_b200.losses import DBLoss
    from db_text_minimal_b200.postprocess import SegDetectorRepresenter
    with pytest.raises(DbbError):
        DBLoss()(torch.rand(1, 3, 8, 8), torch.rand(4, 1, 8, 8))
    with pytest.raises(DbbError):
        SegDetectorRepresenter().binarize(torch.rand(8, 8))
    with pytest.raises(AssertionError):
        DBLoss()(torch.rand(3, 8, 8), torch.rand(4, 1, 8, 8))          # src/losses.py:113


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "db_text_minimal_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("restat", ""), f


def test_shard_range_partitions():
    from db_text_minimal_b200.dist import shard_range
    for total in (16, 17, 64, 5):
        for world in (1, 2, 4, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total and all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def test_offset_convex_round_rectangle():
    """Clipper round-join offset restatement: a 40x10 axis-aligned box grown by d stays inside the d-expanded box,
    touches each expanded side, and its min-area rectangle is the expanded box (+-1 px integer rounding)."""
    import cv2
    from db_text_minimal_b200.postprocess import offset_convex_round
    box = np.array([[10, 20], [50, 20], [50, 30], [10, 30]], dtype=np.float32)
    d = 6.0
    out = offset_convex_round(box, d)
    assert out[:, 0].min() == 4 and out[:, 0].max() == 56 and out[:, 1].min() == 14 and out[:, 1].max() == 36
    (cx, cy), (rw, rh), _ = cv2.minAreaRect(out.astype(np.float32))
    assert abs(max(rw, rh) - 52) <= 1 and abs(min(rw, rh) - 22) <= 1
    # every output point lies within d (+ rounding) of the source polygon
    poly = box.reshape(-1, 1, 2)
    for p in out:
        dist = -cv2.pointPolygonTest(poly, (float(p[0]), float(p[1])), True)
        assert dist <= d + 0.71


def test_synth_value_domains():
    from db_text_minimal_b200 import synth
    g = synth.gt_maps(2, 64, 96, 3)
    assert g.shape == (4, 2, 64, 96) and g.dtype == np.float32
    assert set(np.unique(g[0])) <= {0.0, 1.0} and set(np.unique(g[1])) <= {0.0, 1.0} and set(np.unique(g[3])) <= {0.0, 1.0}
    assert g[2].min() >= 0.3 - 1e-6 and g[2].max() <= 0.7 + 1e-6
    assert synth.images(2, 8, 8, 0).shape == (2, 3, 8, 8)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from db_text_minimal_b200.dist import GradSync
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
class M(torch.nn.Module):            # stands in for DBTextModel: GradSync needs the hook slot, parameters and buffers
    _segment_hook = None
    def __init__(self, rank):
        super().__init__()
        self.w = torch.nn.Parameter(torch.full((5,), float(rank + 1)))
        self.register_buffer("running_mean", torch.full((3,), float(10 * (rank + 1))))
rank = int(sys.argv[3])
m = M(rank)
sync = GradSync(m)                       # broadcasts rank 0's parameters and buffers (replicas built from different seeds)
assert m._segment_hook is sync
assert torch.equal(m.w.data, torch.full((5,), 1.0)) and torch.equal(m.running_mean, torch.full((3,), 10.0))
m.running_mean.fill_(float(rank))
sync.sync_buffers(m)                     # BatchNorm running statistics averaged over the ranks
assert torch.allclose(m.running_mean, torch.full((3,), 0.5))
rank = dist.get_rank()
flat = torch.arange(100, dtype=torch.float32) * (rank + 1)
slices = [(0, 30), (30, 80), (80, 100)]
for seg, b in enumerate(slices):
    m._segment_hook(seg, flat, b)      # what _DBNetFn.backward does after each backward segment
m._segment_hook(3, flat, None)
want = torch.arange(100, dtype=torch.float32) * 1.5      # mean of the two ranks
assert torch.allclose(flat, want), (rank, flat[:5])
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_two_rank_gradient_average(tmp_path):
    """N>1 path on CPU: two gloo ranks, segment-wise asynchronous all-reduce of the flat gradient buffer = mean."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_pretrained_backbone_is_loaded_from_a_local_file_or_warns(tmp_path, monkeypatch):
    """resnet18(pretrained=True) (src/modules/resnet.py:245-255): $DBB_RESNET18_WEIGHTS is loaded with strict=False; when
    nothing can be obtained the backbone stays random AND a RuntimeWarning says so (never silently)."""
    import warnings
    from db_text_minimal_b200.modules import resnet as R
    donor = R.resnet18(pretrained=False)
    sd = {k: torch.full_like(v, 0.25) if v.dtype.is_floating_point else v for k, v in donor.state_dict().items()}
    f = tmp_path / "resnet18.pth"
    torch.save(sd, f)
    monkeypatch.setenv("DBB_RESNET18_WEIGHTS", str(f))
    m = R.resnet18(pretrained=True)
    assert float(m.conv1.weight.min()) == 0.25 and float(m.layer4[1].conv2.weight.max()) == 0.25
    # nothing local, no network: warn
    monkeypatch.delenv("DBB_RESNET18_WEIGHTS")
    monkeypatch.setattr(R, "_imagenet_state_dict", lambda: (_ for _ in ()).throw(OSError("offline")))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = R.resnet18(pretrained=True)
    assert any(issubclass(x.category, RuntimeWarning) and "RANDOMLY" in str(x.message) for x in w)
    assert float(m.conv1.weight.min()) != 0.25
