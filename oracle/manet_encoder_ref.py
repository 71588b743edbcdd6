"""ORACLE (test infrastructure, not product code) — CPU restatement of the MANet feature extractor
``IntVOS.extract_feature`` as ``eval_agent_manet.py:316-328`` calls it (1 x 3 x 480 x 854 frame -> 1 x 100 x 120 x 214
embedding): DeepLabv3+ ResNet-101 (output stride 16) + ASPP + shortcut decoder + semantic-embedding head.

PARITY UNPINNED — restatement from the paper / the reference's config, NOT verified against upstream code.
The network's source (lightas/CVPR2020_MANet, unpinned HEAD) and weights are absent from /root/reference and from this
image; the reference's own tests hold no vector for it (it has no tests).  This file restates the published architecture
under the hyper-parameters the reference pins (utils/config_manet/config.py:108-120; ivosw/manet_arch.py lists them) with
torch functional ops on CPU.  A CUDA result that agrees with this file agrees with THIS restatement, nothing more.

Only tests/ and scripts/ measurement helpers may import this module.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _cbr(x, sd, conv, bn, stride=1, dil=1, k=1, groups=1, relu=True):
    pad = dil * (k - 1) // 2
    y = F.conv2d(x, sd[conv + ".weight"], sd.get(conv + ".bias"), stride, pad, dil, groups)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"], False, 0.0, BN_EPS)
    return F.relu(y) if relu else y


def _bottleneck(x, sd, p, stride, dil, has_ds):
    o = _cbr(x, sd, p + "conv1", p + "bn1")
    o = _cbr(o, sd, p + "conv2", p + "bn2", stride=stride, dil=dil, k=3)
    o = _cbr(o, sd, p + "conv3", p + "bn3", relu=False)
    idt = _cbr(x, sd, p + "downsample.0", p + "downsample.1", stride=stride, relu=False) if has_ds else x
    return F.relu(o + idt)


def extract_feature(sd, frames, probes=None):
    """sd: state dict with ivosw.manet_arch's key names; frames: B x 3 x H x W (already normalised, as MANet's loader
    hands them over).  Returns the B x 100 x h4 x w4 embedding."""
    from ivosw import manet_arch as A
    x = _cbr(frames, sd, "backbone.conv1", "backbone.bn1", stride=2, k=7)
    x = F.max_pool2d(x, 3, 2, 1)
    low = None
    for name, planes, blocks, stride, dils in A.STAGES:
        for b in range(blocks):
            x = _bottleneck(x, sd, "backbone.%s.%d." % (name, b), stride if b == 0 else 1, dils[b], b == 0)
        if name == "layer1":
            low = x
        if probes is not None:
            probes[name] = x
    # ASPP
    branches = [_cbr(x, sd, "aspp.aspp1.conv", "aspp.aspp1.bn")]
    for i, r in enumerate(A.ASPP_RATES, start=2):
        branches.append(_cbr(x, sd, "aspp.aspp%d.conv" % i, "aspp.aspp%d.bn" % i, dil=r, k=3))
    g = _cbr(x.mean((2, 3), keepdim=True), sd, "aspp.gap.conv", "aspp.gap.bn")
    branches.append(F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=True))
    x = _cbr(torch.cat(branches, 1), sd, "aspp.conv1", "aspp.bn1")          # (+ dropout: identity in eval mode)
    if probes is not None:
        probes["aspp"] = x
    # decoder
    lowf = _cbr(low, sd, "decoder.conv1", "decoder.bn1")
    x = F.interpolate(x, size=lowf.shape[2:], mode="bilinear", align_corners=True)
    x = torch.cat((x, lowf), 1)
    x = _cbr(x, sd, "decoder.last_conv.0", "decoder.last_conv.1", k=3)
    x = _cbr(x, sd, "decoder.last_conv.4", "decoder.last_conv.5", k=3)
    if probes is not None:
        probes["decoder"] = x
    # MANet semantic embedding: depthwise 3x3 + BN + ReLU, 1x1 + BN + ReLU
    x = _cbr(x, sd, "embed.dw", "embed.bn1", k=3, groups=256)
    return _cbr(x, sd, "embed.pw", "embed.bn2")
