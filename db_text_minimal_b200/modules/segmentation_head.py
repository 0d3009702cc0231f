"""DBHead parameter tree + step function -- mirror of src/modules/segmentation_head.py:20-108."""
import torch
from torch import nn

from .. import _lib


class _StepFn(torch.autograd.Function):
    """B = 1/(1+exp(-k(P-T))) through dbb_step_fwd / dbb_step_bwd (csrc/db_loss.cu)."""

    @staticmethod
    def forward(ctx, x, y, k):
        _lib.require_cuda(x, y)
        xc, yc = x.detach().float().contiguous(), y.detach().float().contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.lib().dbb_step_fwd(xc.data_ptr(), yc.data_ptr(), out.data_ptr(), xc.numel(), float(k),
                                               _lib.stream_ptr()), "dbb_step_fwd")
        ctx.save_for_backward(xc, yc)
        ctx.k = float(k)
        return out

    @staticmethod
    def backward(ctx, g):
        xc, yc = ctx.saved_tensors
        g = g.float().contiguous()
        dx, dy = torch.empty_like(xc), torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            _lib.check(_lib.lib().dbb_step_bwd(xc.data_ptr(), yc.data_ptr(), g.data_ptr(), dx.data_ptr(), dy.data_ptr(),
                                               xc.numel(), ctx.k, _lib.stream_ptr()), "dbb_step_bwd")
        return dx, dy, None


class DBHead(nn.Module):
    def __init__(self, in_channels, out_channels, k=50):
        super().__init__()
        self.k = k
        c4 = in_channels // 4

        def branch(first_bias):
            return nn.Sequential(
                nn.Conv2d(in_channels, c4, 3, padding=1, bias=first_bias), nn.BatchNorm2d(c4), nn.ReLU(inplace=True),
                nn.ConvTranspose2d(c4, c4, 2, 2), nn.BatchNorm2d(c4), nn.ReLU(inplace=True),
                nn.ConvTranspose2d(c4, 1, 2, 2), nn.Sigmoid())

        self.binarize = branch(True)
        self.thresh = branch(False)      # first conv of `thresh` has no bias (segmentation_head.py:64-68)
        self.binarize.apply(self.weights_init)
        self.thresh.apply(self.weights_init)

    def weights_init(self, m):
        name = m.__class__.__name__
        if name.find('Conv') != -1:
            nn.init.kaiming_normal_(m.weight.data)
        elif name.find('BatchNorm') != -1:
            m.weight.data.fill_(1.)
            m.bias.data.fill_(1e-4)

    def step_function(self, x, y):
        return _StepFn.apply(x, y, self.k)

    def forward(self, x):
        """src/modules/segmentation_head.py:35-45: x (N, 256, H/4, W/4) NCHW float32 -> cat(P, T, B) in training mode,
        cat(P, T) in eval mode.  Stand-alone path (single-operator C ABI)."""
        from .. import _autograd as A
        a = A.ToNHWC.apply(x)
        zs = []
        for br in (self.binarize, self.thresh):
            h = A.conv_bn(a, br[0], br[1], relu=True)
            zs.append(A.ConvT.apply(h, br[3].weight, br[3].bias))
        zt = torch.cat(zs, dim=-1).contiguous()             # (N, H/2, W/2, 128) = [binarize | thresh]
        bnb, bnt = self.binarize[4], self.thresh[4]
        gamma, beta = torch.cat([bnb.weight, bnt.weight]), torch.cat([bnb.bias, bnt.bias])
        training = self.training
        rm, rv = torch.cat([bnb.running_mean, bnt.running_mean]), torch.cat([bnb.running_var, bnt.running_var])
        out = A.HeadTail.apply(zt, gamma, beta, rm, rv, training, self.binarize[6].weight, self.thresh[6].weight,
                               self.binarize[6].bias, self.thresh[6].bias, float(self.k))
        if training:        # the kernel updated the concatenated copies: write the running statistics back
            with torch.no_grad():
                bnb.running_mean.copy_(rm[:64]); bnt.running_mean.copy_(rm[64:])
                bnb.running_var.copy_(rv[:64]); bnt.running_var.copy_(rv[64:])
                bnb.num_batches_tracked += 1; bnt.num_batches_tracked += 1
        return out
