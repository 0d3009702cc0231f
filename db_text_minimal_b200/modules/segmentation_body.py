"""FPN parameter tree -- mirror of src/modules/segmentation_body.py:11-87."""
from torch import nn

from .basic import ConvBnRelu


class FPN(nn.Module):
    def __init__(self, backbone_out_channels, inner_channels=256):
        super().__init__()
        self.conv_out = inner_channels
        inner = inner_channels // 4
        self.reduce_conv_c2 = ConvBnRelu(backbone_out_channels[0], inner, kernel_size=1)
        self.reduce_conv_c3 = ConvBnRelu(backbone_out_channels[1], inner, kernel_size=1)
        self.reduce_conv_c4 = ConvBnRelu(backbone_out_channels[2], inner, kernel_size=1)
        self.reduce_conv_c5 = ConvBnRelu(backbone_out_channels[3], inner, kernel_size=1)
        self.smooth_p4 = ConvBnRelu(inner, inner, kernel_size=3, padding=1)
        self.smooth_p3 = ConvBnRelu(inner, inner, kernel_size=3, padding=1)
        self.smooth_p2 = ConvBnRelu(inner, inner, kernel_size=3, padding=1)
        self.conv = nn.Sequential(nn.Conv2d(self.conv_out, self.conv_out, kernel_size=3, padding=1, stride=1),
                                  nn.BatchNorm2d(self.conv_out), nn.ReLU(inplace=True))
        self.out_channels = self.conv_out

    def forward(self, x):
        """src/modules/segmentation_body.py:64-77: x = (c2, c3, c4, c5) NCHW float32 -> fused map (N, 256, H/4, W/4).
        Stand-alone path (single-operator C ABI); DBTextModel runs the same layers inside the fused executor."""
        from .. import _autograd as A
        c2, c3, c4, c5 = (A.ToNHWC.apply(t) for t in x)
        cbr = lambda m, t: A.conv_bn(t, m.conv, m.bn, relu=True)
        p5 = cbr(self.reduce_conv_c5, c5)
        p4 = cbr(self.smooth_p4, A.UpsampleAdd.apply(p5, cbr(self.reduce_conv_c4, c4)))
        p3 = cbr(self.smooth_p3, A.UpsampleAdd.apply(p4, cbr(self.reduce_conv_c3, c3)))
        p2 = cbr(self.smooth_p2, A.UpsampleAdd.apply(p3, cbr(self.reduce_conv_c2, c2)))
        cat = A.UpsampleCat.apply(p2, p3, p4, p5)
        return A.ToNCHW.apply(A.conv_bn(cat, self.conv[0], self.conv[1], relu=True))
