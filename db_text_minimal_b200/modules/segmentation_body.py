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
        from .._lib import DbbError
        raise DbbError("the FPN runs inside the fused DBTextModel graph (csrc/net.cu)")
