"""DBLoss and its three terms, drop-in for the reference's src/losses.py, running on libdbb200.so.

Same class names, constructor arguments, call signatures and return values as the reference
(src/losses.py:11-139); the arithmetic runs in the fused CUDA kernels of csrc/db_loss.cu.  The three
host synchronisations of the reference (``int(positive.sum())`` x2 at :25,27 and ``assert loss <= 1``
at :65) are gone: counts stay on the device.  CUDA tensors only -- there is no CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

_RED = {"mean": 0, "none": 1}


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _DBLossFn(torch.autograd.Function):
    """preds (N,C,H,W), gts (4,N,H,W) -> losses5 (5,) float32 on device."""

    @staticmethod
    def forward(ctx, preds, gts, alpha, beta, reduction, negative_ratio, eps):
        _lib.require_cuda(preds, gts)
        L = _lib.lib()
        preds_c, gts_c = _f32c(preds.detach()), _f32c(gts.detach())
        n, c, h, w = preds_c.shape
        if tuple(gts_c.shape) != (4, n, h, w):
            raise ValueError(f"gts must be (4, N, H, W) = (4, {n}, {h}, {w}) (src/train.py:163-166), got {tuple(gts_c.shape)}")
        red = _RED[reduction]
        ws_bytes = L.dbb_dbloss_workspace(n, c, h, w, red)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=preds.device)
        losses = torch.empty(5, dtype=torch.float32, device=preds.device)
        state = torch.zeros(C.sizeof(_lib.DbbLossState), dtype=torch.uint8, device=preds.device)
        with torch.cuda.device(preds.device):
            _lib.check(L.dbb_dbloss_fwd(preds_c.data_ptr(), gts_c.data_ptr(), n, c, h, w, alpha, beta, red,
                                        float(negative_ratio), eps, losses.data_ptr(), state.data_ptr(),
                                        ws.data_ptr(), ws_bytes, _lib.stream_ptr()), "dbb_dbloss_fwd")
        ctx.save_for_backward(preds_c, gts_c, state)
        ctx.cfg = (alpha, beta, red, eps)
        ctx.mark_non_differentiable(state)
        return losses, state

    @staticmethod
    def backward(ctx, g_losses, _g_state):
        preds_c, gts_c, state = ctx.saved_tensors
        alpha, beta, red, eps = ctx.cfg
        L = _lib.lib()
        n, c, h, w = preds_c.shape
        go = _f32c(g_losses)
        dpreds = torch.empty_like(preds_c)
        with torch.cuda.device(preds_c.device):
            _lib.check(L.dbb_dbloss_bwd(preds_c.data_ptr(), gts_c.data_ptr(), n, c, h, w, alpha, beta, red, eps,
                                        go.data_ptr(), state.data_ptr(), dpreds.data_ptr(), _lib.stream_ptr()),
                       "dbb_dbloss_bwd")
        return dpreds, None, None, None, None, None, None


def read_state(state_tensor):
    """Host copy of the device DbbLossState (one D2H sync; for tests / logging only)."""
    raw = bytes(state_tensor.cpu().numpy().tobytes())
    return _lib.DbbLossState.from_buffer_copy(raw)


def _pack_gts(gt, mask, thresh_gt=None, thresh_mask=None):
    z = torch.zeros_like(gt)
    return torch.stack([gt, mask, thresh_gt if thresh_gt is not None else z,
                        thresh_mask if thresh_mask is not None else z])


class OHEMBalanceCrossEntropyLoss(nn.Module):
    """src/losses.py:11-40."""

    def __init__(self, negative_ratio=3, eps=1e-6, reduction='mean'):
        super().__init__()
        self.negative_ratio = negative_ratio
        self.eps = eps
        self.reduction = reduction

    def forward(self, pred, gt, mask):
        preds = torch.stack([pred, torch.zeros_like(pred)], 1)          # (N, 2, H, W): [P, T=0]
        losses, _ = _DBLossFn.apply(preds, _pack_gts(gt, mask), 1.0, 0.0, self.reduction, self.negative_ratio, self.eps)
        return losses[0]


class DiceLoss(nn.Module):
    """src/losses.py:43-66 (the ``assert loss <= 1`` host sync is dropped; it holds by construction
    for pred, gt, mask >= 0)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.eps = eps

    def forward(self, pred, gt, mask, weights=None):
        half = torch.full_like(pred, 0.5)
        preds = torch.stack([half, half, pred], 1)
        losses, _ = _DBLossFn.apply(preds, _pack_gts(gt, mask), 1.0, 0.0, 'mean', 3, self.eps)
        return losses[2]


class L1Loss(nn.Module):
    """src/losses.py:69-82."""

    def __init__(self, eps=1e-6, reduction='mean'):
        super().__init__()
        self.eps = eps
        self.reduction = reduction

    def forward(self, pred, gt, mask):
        if mask is None:   # src/losses.py:79-81: plain torch.nn.L1Loss with self.reduction
            mask = torch.ones_like(pred)
            if self.reduction != 'mean':
                raise _lib.DbbError("L1Loss(mask=None) is only supported with reduction='mean'")
            eps = 0.0
        else:
            eps = self.eps
        half = torch.full_like(pred, 0.5)
        preds = torch.stack([half, pred], 1)
        z = torch.zeros_like(pred)
        losses, _ = _DBLossFn.apply(preds, torch.stack([z, z, gt, mask]), 1.0, 1.0, 'mean', 3, eps)
        return losses[1]


class DBLoss(nn.Module):
    """src/losses.py:85-139.  Returns the reference's 5-tuple (3-channel preds) or the single
    ``prob_threshold_loss`` tensor (2-channel preds); all are 0-dim CUDA tensors with autograd."""

    def __init__(self, alpha=1.0, beta=10.0, reduction='mean', negative_ratio=3, eps=1e-6):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.reduction = reduction
        self.negative_ratio = negative_ratio
        self.eps = eps
        self.ohem_loss = OHEMBalanceCrossEntropyLoss(self.negative_ratio, self.eps, self.reduction)
        self.dice_loss = DiceLoss(self.eps)
        self.l1_loss = L1Loss(self.eps, self.reduction)
        self.last_state = None    # device DbbLossState of the most recent call (n_pos, n_neg, tau, ...)

    def forward(self, preds, gts):
        assert preds.dim() == 4      # src/losses.py:113
        assert gts.dim() == 4        # src/losses.py:114
        losses, state = _DBLossFn.apply(preds, gts, float(self.alpha), float(self.beta), self.reduction,
                                        self.negative_ratio, float(self.eps))
        self.last_state = state
        if preds.size(1) == 3:
            return losses[0], losses[1], losses[2], losses[3], losses[4]
        return losses[3]
