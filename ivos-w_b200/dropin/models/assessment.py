"""Drop-in for the reference's ``models/assessment.py`` (Encoder 12-63, AssessNet 66-182).

Same class names, ``state_dict`` keys (including the registered-but-unused
``conv1_m`` / ``conv1_n``, SURVEY A.Q5) and ``forward(tf, tp)`` signature and
output shape (B x 1, or ``(1,)`` for B == 1, SURVEY A.Q1).  The modules are
parameter containers; bbox, ROI crop, the 4-channel-stem ResNet-50, pooling and
the FC all run in the CUDA library.  No torchvision dependency, no download.
"""
import math

import torch
import torch.nn as nn

from ivosw.engine import get_engine


class _Bottleneck(nn.Module):
    """Parameter container with torchvision Bottleneck's attribute names."""

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))


def _stage(inplanes, planes, blocks, stride):
    layers = [_Bottleneck(inplanes, planes, stride, True)]
    layers += [_Bottleneck(planes * 4, planes, 1, False) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


class Encoder(nn.Module):
    def __init__(self):
        super(Encoder, self).__init__()
        self.conv1_m = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=True)
        self.conv1_p = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.conv1_n = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
        # the reference takes these from torchvision resnet50(pretrained=True); here they are
        # containers to be filled by load_state_dict (weights/assess_net.pt)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.res2 = _stage(64, 64, 3, 1)
        self.res3 = _stage(256, 128, 4, 2)
        self.res4 = _stage(512, 256, 6, 2)
        self.res5 = _stage(1024, 512, 3, 2)
        self.register_buffer('mean', torch.FloatTensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.FloatTensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, in_f, in_p, in_g=None):
        raise RuntimeError("Encoder is evaluated inside AssessNet.forward by the CUDA library; "
                           "it has no standalone PyTorch forward")


class AssessNet(nn.Module):
    def __init__(self):
        super(AssessNet, self).__init__()
        self.Encoder = Encoder()
        self.fc1 = nn.Linear(2048, 1)
        self.cnt = 0
        self._uploaded = None

    def _sync(self, engine):
        sd = self.state_dict()
        stamp = tuple((v.data_ptr(), v._version) for v in sd.values()) + (id(engine),)
        if stamp != self._uploaded:
            engine.load_assess(sd)
            self._uploaded = stamp

    def forward(self, tf, tp):
        """tf: B x 3 x H x W, tp: B x H x W -> B x 1 (``(1,)`` when B == 1)  (assessment.py:164-182)."""
        if not tf.is_cuda:
            raise RuntimeError("ivosw_b200 AssessNet runs on CUDA only (no CPU fallback); got a %s tensor" % tf.device)
        engine = get_engine(tf.device)
        self._sync(engine)
        scores = engine.assess_forward(tf, tp)
        return scores if scores.shape[0] == 1 else scores.unsqueeze(1)
