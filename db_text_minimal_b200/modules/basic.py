"""ConvBnRelu -- mirror of src/modules/basic.py:7-36 (Conv2d(bias=True) + BatchNorm2d + ReLU)."""
from torch import nn


class ConvBnRelu(nn.Module):
    """Holds ``conv`` and ``bn`` exactly like the reference so state_dict keys match.  Inside DBTextModel the arithmetic
    runs in the fused executor (csrc/net.cu); called on its own the block runs conv -> BatchNorm -> ReLU through the
    single-operator C ABI (``_autograd``), NCHW float32 in and out like the reference (src/modules/basic.py:31-36)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', inplace=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                              groups=groups, bias=bias, padding_mode=padding_mode)
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=inplace)

    def forward(self, x):
        from .. import _autograd as A
        return A.ToNCHW.apply(A.conv_bn(A.ToNHWC.apply(x), self.conv, self.bn, relu=True))
