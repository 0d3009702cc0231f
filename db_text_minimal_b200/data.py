"""Loader -> device boundary: ship a batch as uint8 image + uint8 {0,1} maps + float32 threshold map (65.6 MB per 16x640x640
batch instead of the 184 MB of float32 tensors the reference's DataLoader produces, src/data_loaders.py:152-166) and expand
it on the device with one launch (csrc/data.cu).  Lossless: the expanded tensors are bit-identical to the float32 ones."""
import torch

from . import _lib

REFERENCE_MEAN = (103.939, 116.779, 123.68)      # src/data_loaders.py:28, subtracted per channel after the BGR->RGB flip


def pack_batch(img_f32, gts_f32, mean=REFERENCE_MEAN):
    """Inverse of unpack_batch for a float32 batch whose image is (uint8 - mean): returns (img_u8, prob_u8, mask_u8, thresh_f32,
    area_u8).  Raises if the batch is not representable (pixels not integers in 0..255 after adding the mean, maps not {0,1})."""
    m = torch.tensor(mean, dtype=torch.float32, device=img_f32.device).view(1, 3, 1, 1)
    raw = img_f32 + m
    u8 = raw.round().clamp(0, 255).to(torch.uint8)
    if not torch.equal(u8.float() - m, img_f32):
        raise ValueError("pack_batch: the image is not (uint8 - mean)")
    maps = []
    for k in (0, 1, 3):
        b = gts_f32[k].to(torch.uint8)
        if not torch.equal(b.float(), gts_f32[k]):
            raise ValueError("pack_batch: ground-truth map %d is not {0, 1}-valued" % k)
        maps.append(b.contiguous())
    return u8.contiguous(), maps[0], maps[1], gts_f32[2].contiguous(), maps[2]


def unpack_batch(img_u8, prob_u8, mask_u8, thresh_f32, area_u8, mean=REFERENCE_MEAN, out_img=None, out_gts=None):
    """Device tensors in, device tensors out: (img float32 (N,3,H,W) = uint8 - mean[c], gts float32 (4,N,H,W) in the order of
    src/train.py:163-166).  out_img / out_gts: write into existing buffers (e.g. GraphedTrainStep.img / .gts)."""
    _lib.require_cuda(img_u8, prob_u8, mask_u8, thresh_f32, area_u8)
    n, c, h, w = img_u8.shape
    assert c == 3 and img_u8.dtype == torch.uint8 and thresh_f32.dtype == torch.float32
    dev = img_u8.device
    if out_img is None:
        out_img = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
    if out_gts is None:
        out_gts = torch.empty((4, n, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbb_unpack_batch(img_u8.contiguous().data_ptr(), float(mean[0]), float(mean[1]), float(mean[2]),
                                               prob_u8.contiguous().data_ptr(), mask_u8.contiguous().data_ptr(),
                                               thresh_f32.contiguous().data_ptr(), area_u8.contiguous().data_ptr(), n, h, w,
                                               out_img.data_ptr(), out_gts.data_ptr(), _lib.stream_ptr()), "dbb_unpack_batch")
    return out_img, out_gts
