"""ResNet-18 parameter tree -- mirror of src/modules/resnet.py:37-91,162-255 (only what 'resnet18' needs).

Same attribute names => same state_dict keys (conv1, bn1, layer{1-4}.{0,1}.{conv1,bn1,conv2,bn2[,downsample.{0,1}]},
plus the never-used ``fc`` and ``smooth`` members the reference carries).  Initialisation follows
src/modules/resnet.py:197-203.  ``pretrained=True`` cannot download anything here (no network): it is accepted and
ignored, and a checkpoint is loaded through load_state_dict like the reference does (src/test.py:16)."""
import math

import torch.nn as nn

__all__ = ['ResNet', 'BasicBlock', 'resnet18']


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, dcn=None):
        super().__init__()
        if dcn is not None:
            raise NotImplementedError("deformable convolutions are outside the resnet18-FPN-DBHead hot path")
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def _forward_nhwc(self, x):
        """src/modules/resnet.py:70-91 on NHWC bf16 tensors."""
        from .. import _autograd as A
        a1 = A.conv_bn(x, self.conv1, self.bn1, relu=True)
        res = x if self.downsample is None else A.conv_bn(x, self.downsample[0], self.downsample[1], relu=False)
        return A.conv_bn(a1, self.conv2, self.bn2, relu=True, residual=res)

    def forward(self, x):
        from .. import _autograd as A
        return A.ToNCHW.apply(self._forward_nhwc(A.ToNHWC.apply(x)))


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, dcn=None):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2)
        self.avgpool = nn.AvgPool2d(7, stride=1)
        self.fc = nn.Linear(512 * block.expansion, num_classes)          # unused in forward, kept for the state dict
        self.smooth = nn.Conv2d(2048, 256, kernel_size=1, stride=1, padding=1)   # idem
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        """src/modules/resnet.py:231-242: returns (c2, c3, c4, c5), NCHW float32.  Stand-alone path (single-operator
        C ABI); DBTextModel runs the same layers inside the fused executor."""
        from .. import _autograd as A
        a = A.batch_norm(A.Stem.apply(x, self.conv1.weight), self.bn1, relu=True)
        a = A.MaxPool.apply(a)
        feats = []
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                a = blk._forward_nhwc(a)
            feats.append(A.ToNCHW.apply(a))
        return tuple(feats)


model_urls = {'resnet18': 'https://download.pytorch.org/models/resnet18-5c106cde.pth'}


def _imagenet_state_dict():
    """The reference (src/modules/resnet.py:245-255) fetches the ImageNet ResNet-18 through model_zoo.load_url.  Same
    lookup here, in this order: $DBB_RESNET18_WEIGHTS (a local .pth), the torch hub checkpoint cache (where load_url
    itself would have left the file), then load_url (needs network)."""
    import os
    import torch
    cands = [os.environ.get("DBB_RESNET18_WEIGHTS")]
    try:
        cands.append(os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(model_urls['resnet18'])))
    except Exception:
        pass
    for c in cands:
        if c and os.path.exists(c):
            return torch.load(c, map_location="cpu")
    import socket
    from torch.utils import model_zoo
    old = socket.getdefaulttimeout()
    socket.setdefaulttimeout(float(os.environ.get("DBB_DOWNLOAD_TIMEOUT", "10")))     # an offline box must not hang here
    try:
        return model_zoo.load_url(model_urls['resnet18'], progress=False)
    finally:
        socket.setdefaulttimeout(old)


def resnet18(pretrained=True, **kwargs):
    """Constructs a ResNet-18 model (src/modules/resnet.py:245-255): ImageNet weights with strict=False when
    pretrained.  When the weights cannot be obtained (no network, nothing cached) the backbone keeps its random
    initialisation and a RuntimeWarning says so -- set DBB_RESNET18_WEIGHTS=<resnet18-5c106cde.pth> or load a
    checkpoint afterwards."""
    model = ResNet(BasicBlock, [2, 2, 2, 2], **kwargs)
    if pretrained:
        try:
            sd = _imagenet_state_dict()
        except Exception as e:        # offline box
            import warnings
            warnings.warn("resnet18(pretrained=True): ImageNet weights unavailable (%s: %s); the backbone is RANDOMLY "
                          "initialised. Set DBB_RESNET18_WEIGHTS or load a checkpoint." % (type(e).__name__, e),
                          RuntimeWarning, stacklevel=2)
        else:
            print('load from imagenet')
            model.load_state_dict(sd, strict=False)
    return model
