"""DBTextModel -- drop-in for the reference's src/models.py on libdbb200.so.

Same registries (``backbone_dict`` / ``segmentation_body_dict`` / ``segmentation_head_dict``), same module tree and
``state_dict`` keys (211 tensors), same ``forward(x)`` contract: (N,3,H,W) float32 -> (N,3,H,W) in training mode
([P, T, B]) or (N,2,H,W) in eval mode.  The whole forward (and, through autograd, the whole backward) is ONE call
into the native executor ``dbb_net_forward`` / ``dbb_net_backward`` (csrc/net.cu).  CUDA only; no fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .modules.resnet import resnet18
from .modules.segmentation_body import FPN
from .modules.segmentation_head import DBHead

backbone_dict = {'resnet18': {'models': resnet18, 'out': [64, 128, 256, 512]}}
segmentation_body_dict = {'FPN': FPN}
segmentation_head_dict = {'DBHead': DBHead}


class _Plan:
    """One (N, H, W, training, precision) configuration of the native executor."""

    def __init__(self, n, h, w, training, precision=0):
        L = _lib.lib()
        self.key = (n, h, w, training)
        self.handle = L.dbb_net_create_ex(n, h, w, 1 if training else 0, precision)
        if not self.handle:
            raise _lib.DbbError("dbb_net_create failed: " + L.dbb_last_cuda_error().decode())
        self.ws_bytes = L.dbb_net_workspace_bytes(self.handle)
        self.out_c = L.dbb_net_out_channels(self.handle)
        self.flops_fwd = L.dbb_net_flops_fwd(self.handle)

    def __del__(self):
        try:
            _lib.lib().dbb_net_destroy(self.handle)
        except Exception:
            pass


def _aligned_workspace(nbytes, device):
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 1024
    return raw, raw.data_ptr() + off


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


class _DBNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, plan, x, *params):
        L = _lib.lib()
        dev = x.device
        n, _, h, w = x.shape
        out = torch.empty((n, plan.out_c, h, w), dtype=torch.float32, device=dev)
        training = plan.key[3]
        if training:
            ws_raw, ws_ptr = _aligned_workspace(plan.ws_bytes, dev)
        else:
            ws_raw, ws_ptr = model._eval_workspace(plan, dev)
        pa = _ptr_array([p.data_ptr() for p in params])
        ba = _ptr_array([b.data_ptr() for b in model._buffer_list()])
        with torch.cuda.device(dev):
            _lib.check(L.dbb_net_forward(plan.handle, x.data_ptr(), pa, ba, out.data_ptr(), ws_ptr, plan.ws_bytes,
                                         _lib.stream_ptr()), "dbb_net_forward")
        if training:
            ctx.plan, ctx.ws_raw, ctx.ws_ptr, ctx.model = plan, ws_raw, ws_ptr, model
            ctx.save_for_backward(out, x, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        out, x, *params = ctx.saved_tensors
        plan, model = ctx.plan, ctx.model
        dev = out.device
        dout = dout.float().contiguous()
        flat, views = model._grad_views(dev)
        model._last_flat_grad = flat          # optim.FlatAdam steps on this buffer directly
        pa = _ptr_array([p.data_ptr() for p in params])
        ga = _ptr_array([v.data_ptr() for v in views])
        hook = model._segment_hook
        nseg = L.dbb_net_num_segments()
        with torch.cuda.device(dev):
            for seg in range(nseg):
                _lib.check(L.dbb_net_backward_ex(plan.handle, x.data_ptr(), out.data_ptr(), dout.data_ptr(), pa, ga, ctx.ws_ptr,
                                                 plan.ws_bytes, seg, _lib.stream_ptr()), "dbb_net_backward")
                if hook is not None:
                    hook(seg, flat, model._segment_slices[seg])
        if hook is not None:
            hook(nseg, flat, None)     # final: wait for outstanding reductions
        ctx.ws_raw = None
        grads = [None if model._unused[i] else views[i] for i in range(len(views))]
        return (None, None, None, *grads)


PRECISIONS = {"bf16": 0, "fp32": 1}


class DBTextModel(nn.Module):
    """precision: 'bf16' (default; tcgen05 convolutions on bf16 activations, fp32 accumulate) or 'fp32' (parity mode: the
    same executor graph on float32 activations with CUDA-core convolutions -- slow, matches the reference to fp32
    round-off).  $DBB_PRECISION overrides the default.  pretrained: as the reference's hard-coded True
    (src/models.py:17); see modules/resnet.py:resnet18 for where the ImageNet weights are looked up."""

    def __init__(self, precision=None, pretrained=True):
        super().__init__()
        import os
        precision = precision or os.environ.get("DBB_PRECISION", "bf16")
        if precision not in PRECISIONS:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        backbone_name = "resnet18"
        segmentation_body_name = "FPN"
        segmentation_head_name = "DBHead"
        backbone_model, backbone_out = backbone_dict[backbone_name]['models'], backbone_dict[backbone_name]['out']
        self.backbone = backbone_model(pretrained=pretrained)
        self.segmentation_body = segmentation_body_dict[segmentation_body_name](backbone_out, inner_channels=256)
        self.segmentation_head = segmentation_head_dict[segmentation_head_name](self.segmentation_body.out_channels,
                                                                                out_channels=2)
        self.name = '{}_{}_{}'.format(backbone_name, segmentation_body_name, segmentation_head_name)
        self._plans = {}
        self._eval_ws = {}
        self._segment_hook = None
        self._bind()

    # ------------------------------------------------------------------ binding to the native parameter order
    def _bind(self):
        L = _lib.lib()
        named_p = dict(self.named_parameters())
        named_b = dict(self.named_buffers())
        self._pnames = [L.dbb_net_param_name(i).decode() for i in range(L.dbb_net_num_params())]
        self._bnames = [L.dbb_net_buffer_name(i).decode() for i in range(L.dbb_net_num_buffers())]
        missing = [k for k in self._pnames if k not in named_p] + [k for k in self._bnames if k not in named_b]
        extra = [k for k in named_p if k not in self._pnames]
        if missing or extra:
            raise _lib.DbbError(f"parameter tree does not match the native executor: missing {missing}, extra {extra}")
        for i, k in enumerate(self._pnames):
            if named_p[k].numel() != L.dbb_net_param_numel(i):
                raise _lib.DbbError(f"shape mismatch for {k}")
        self._unused = [k.startswith("backbone.fc.") or k.startswith("backbone.smooth.") for k in self._pnames]
        # flat gradient layout: [segment 0: head + FPN | segment 1: layer4, layer3 | segment 2: layer2, layer1, stem]
        def seg_of(k):
            if k.startswith("segmentation_"):
                return 0
            if k.startswith("backbone.layer4") or k.startswith("backbone.layer3"):
                return 1
            return 2
        order = sorted([i for i in range(len(self._pnames)) if not self._unused[i]], key=lambda i: (seg_of(self._pnames[i]), i))
        self._flat_offsets = {}
        off = 0
        bounds = [0, 0, 0, 0]
        for i in order:
            self._flat_offsets[i] = off
            off += (named_p[self._pnames[i]].numel() + 3) // 4 * 4      # keep 16-byte alignment
            bounds[seg_of(self._pnames[i]) + 1] = off
        self._flat_numel = off
        used = torch.zeros(off, dtype=torch.bool)
        for i in order:
            used[self._flat_offsets[i]:self._flat_offsets[i] + named_p[self._pnames[i]].numel()] = True
        self._flat_pad = torch.nonzero(~used).flatten()
        self._flat_pad_dev = {}
        self._segment_slices = [(bounds[s], bounds[s + 1]) for s in range(3)]

    def _param_list(self):
        named = dict(self.named_parameters())
        return [named[k] for k in self._pnames]

    def _buffer_list(self):
        named = dict(self.named_buffers())
        return [named[k] for k in self._bnames]

    def _grad_views(self, device):
        flat = torch.empty(self._flat_numel, dtype=torch.float32, device=device)
        if self._flat_pad.numel():            # alignment gaps (behind the two 1-element ConvTranspose biases): keep them finite
            key = str(device)
            if key not in self._flat_pad_dev:
                self._flat_pad_dev[key] = self._flat_pad.to(device)
            flat.index_fill_(0, self._flat_pad_dev[key], 0.0)
        plist = self._param_list()
        views = []
        for i, p in enumerate(plist):
            if self._unused[i]:
                views.append(p)        # placeholder pointer, never written
            else:
                o = self._flat_offsets[i]
                views.append(flat[o:o + p.numel()].view_as(p))
        return flat, views

    def _eval_workspace(self, plan, device):
        key = (plan.key, str(device))
        if key not in self._eval_ws:
            self._eval_ws = {key: _aligned_workspace(plan.ws_bytes, device)}    # keep one
        return self._eval_ws[key]

    def _plan(self, n, h, w, training):
        key = (n, h, w, training)
        if key not in self._plans:
            self._plans[key] = _Plan(n, h, w, training, PRECISIONS[self.precision])
        return self._plans[key]

    # ------------------------------------------------------------------ the reference's public surface
    def forward(self, x):
        """
        :return: TRAIN mode: prob_map, threshold_map, appro_binary_map
        :return: EVAL mode: prob_map, threshold_map
        """
        _lib.require_cuda(x)
        if x.dim() != 4 or x.size(1) != 3:
            raise ValueError("expected (N, 3, H, W) input")
        x = x.float().contiguous()
        n, _, h, w = x.shape
        plan = self._plan(n, h, w, bool(self.training))
        params = self._param_list()
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise _lib.DbbError("DBTextModel parameters must be float32 CUDA tensors (call .cuda())")
        if self.training:
            y = _DBNetFn.apply(self, plan, x, *params)
            with torch.no_grad():      # BatchNorm2d bookkeeping, one fused launch for the 32 counters
                torch._foreach_add_([m.num_batches_tracked for m in self.modules()
                                     if isinstance(m, nn.BatchNorm2d) and m.num_batches_tracked is not None], 1)
        else:
            with torch.no_grad():
                y = _DBNetFn.apply(self, plan, x, *params)
        return y

    def flops_fwd(self, n, h, w):
        """2*MACs of the convolutions of one forward pass (bench.py roofline)."""
        return self._plan(n, h, w, bool(self.training)).flops_fwd
