"""Loader -> device boundary (csrc/data.cu): the compact batch expands to exactly the float32 tensors of the reference's loader
(src/data_loaders.py:152-158: img.astype(np.float32); img[..., c] -= mean[c]; src/train.py:163-166 for the map order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,w", [(2, 64, 64), (3, 52, 76), (16, 640, 640)])
def test_unpack_batch_is_bit_identical_to_the_reference_loader_arithmetic(n, h, w):
    from db_text_minimal_b200 import data, synth
    rng = np.random.RandomState(3)
    img_u8 = torch.from_numpy(rng.randint(0, 256, (n, 3, h, w)).astype(np.uint8))
    gts = torch.from_numpy(synth.gt_maps(n, h, w, seed=2))
    # the reference's arithmetic, on the host in numpy float32
    want_img = img_u8.numpy().astype(np.float32)
    for c in range(3):
        want_img[:, c] -= data.REFERENCE_MEAN[c]
    packed = data.pack_batch(torch.from_numpy(want_img), gts)
    assert torch.equal(packed[0], img_u8)
    out_img, out_gts = data.unpack_batch(*[t.cuda() for t in packed])
    assert torch.equal(out_img.cpu(), torch.from_numpy(want_img))
    assert torch.equal(out_gts.cpu(), gts)
    nbytes = sum(t.numel() * t.element_size() for t in packed)
    assert nbytes * 2.7 < (want_img.size + gts.numel()) * 4


def test_pack_batch_rejects_what_is_not_representable():
    from db_text_minimal_b200 import data, synth
    gts = torch.from_numpy(synth.gt_maps(1, 32, 32, seed=1))
    with pytest.raises(ValueError):
        data.pack_batch(torch.randn(1, 3, 32, 32), gts)
    img = torch.zeros(1, 3, 32, 32) - torch.tensor(data.REFERENCE_MEAN).view(1, 3, 1, 1)
    bad = gts.clone(); bad[0] *= 0.5
    with pytest.raises(ValueError):
        data.pack_batch(img, bad)
