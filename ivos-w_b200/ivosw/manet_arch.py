"""Static description of the MANet feature extractor ``IntVOS.extract_feature`` (call site
``eval_agent_manet.py:316-328``): DeepLabv3+ with a ResNet-101 backbone at output stride 16, ASPP (256), the
48-channel low-level shortcut decoder, and MANet's semantic-embedding head (depthwise 3x3 + 1x1 -> 100 channels at
1/4 resolution).

RESTATEMENT, NOT VERIFIED AGAINST UPSTREAM CODE.  The MANet sources (lightas/CVPR2020_MANet, cloned at an unpinned HEAD
by the reference's README.md:39) are not part of the reference tree and not available here.  What the
reference pins are the hyper-parameters in ``utils/config_manet/config.py``: ``MODEL_BACKBONE='res101_atrous'`` (:109),
``MODEL_OUTPUT_STRIDE=16`` (:110), ``MODEL_ASPP_OUTDIM=256`` (:111), ``MODEL_SHORTCUT_DIM=48`` (:112),
``MODEL_SEMANTIC_EMBEDDING_DIM=100`` (:115); everything else below follows the published DeepLabv3+ / FEELVOS / MANet
descriptions (bottleneck [3, 4, 23, 3] with the stride on the 3x3 convolution, multi-grid (1, 2, 4) x dilation 2 in the
last stage, ASPP rates 1 / 6 / 12 / 18 plus image pooling, bilinear x4 upsampling with align_corners=True).  The key names
are THIS repository's; a checkpoint of the real network would need a key map that cannot be written without its source.

The table is the single source of truth for (a) the synthetic weight generator, (b) the CPU restatement
(oracle/manet_encoder_ref.py) and (c) the order of the flat parameter blob crossing the C ABI (ivosw_manet_encoder_load);
csrc/manet_encoder.cu mirrors it.
"""
from collections import namedtuple

BN_EPS = 1e-5
ASPP_DIM, SHORTCUT_DIM, EMBED_DIM = 256, 48, 100            # config.py:111, 112, 115
ASPP_RATES = (6, 12, 18)
STAGES = (("layer1", 64, 3, 1, (1, 1, 1)),                  # name, planes, blocks, stride, dilation per block
          ("layer2", 128, 4, 2, (1, 1, 1, 1)),
          ("layer3", 256, 23, 2, (1,) * 23),
          ("layer4", 512, 3, 1, (2, 4, 8)))                 # output stride 16: stride 1, multi-grid (1, 2, 4) x 2

# one convolution + BatchNorm: name (state-dict prefix of the conv), bn prefix, cin, cout, k, stride, dilation, groups,
# bias (a conv bias, folded into the BatchNorm shift when packing), relu
Conv = namedtuple("Conv", "name bn cin cout k stride dil groups bias relu")


def convs():
    """Every convolution of the encoder in execution order (the order of the parameter blob)."""
    out = [Conv("backbone.conv1", "backbone.bn1", 3, 64, 7, 2, 1, 1, False, True)]
    inplanes = 64
    for name, planes, blocks, stride, dils in STAGES:
        for b in range(blocks):
            s = stride if b == 0 else 1
            p = "backbone.%s.%d." % (name, b)
            out.append(Conv(p + "conv1", p + "bn1", inplanes, planes, 1, 1, 1, 1, False, True))
            out.append(Conv(p + "conv2", p + "bn2", planes, planes, 3, s, dils[b], 1, False, True))
            if b == 0:
                out.append(Conv(p + "downsample.0", p + "downsample.1", inplanes, planes * 4, 1, s, 1, 1, False, False))
            out.append(Conv(p + "conv3", p + "bn3", planes, planes * 4, 1, 1, 1, 1, False, True))
            inplanes = planes * 4
    out.append(Conv("aspp.aspp1.conv", "aspp.aspp1.bn", 2048, ASPP_DIM, 1, 1, 1, 1, False, True))
    for i, r in enumerate(ASPP_RATES, start=2):
        out.append(Conv("aspp.aspp%d.conv" % i, "aspp.aspp%d.bn" % i, 2048, ASPP_DIM, 3, 1, r, 1, False, True))
    out.append(Conv("aspp.gap.conv", "aspp.gap.bn", 2048, ASPP_DIM, 1, 1, 1, 1, False, True))
    out.append(Conv("aspp.conv1", "aspp.bn1", 5 * ASPP_DIM, ASPP_DIM, 1, 1, 1, 1, False, True))
    out.append(Conv("decoder.conv1", "decoder.bn1", 256, SHORTCUT_DIM, 1, 1, 1, 1, False, True))
    out.append(Conv("decoder.last_conv.0", "decoder.last_conv.1", ASPP_DIM + SHORTCUT_DIM, 256, 3, 1, 1, 1, False, True))
    out.append(Conv("decoder.last_conv.4", "decoder.last_conv.5", 256, 256, 3, 1, 1, 1, False, True))
    out.append(Conv("embed.dw", "embed.bn1", 256, 256, 3, 1, 1, 256, True, True))
    out.append(Conv("embed.pw", "embed.bn2", 256, EMBED_DIM, 1, 1, 1, 1, True, True))
    return out


def state_dict_keys():
    keys = []
    for c in convs():
        keys.append(c.name + ".weight")
        if c.bias:
            keys.append(c.name + ".bias")
        keys += [c.bn + "." + s for s in ("weight", "bias", "running_mean", "running_var")]
    return keys


def feature_sizes(H, W):
    """(h, w) at 1/2, 1/4, 1/8, 1/16 of an H x W frame (7x7/2 pad 3; 3x3/2 pad 1 max-pool and convolutions)."""
    def half(n):
        return (n + 2 - 3) // 2 + 1
    h2, w2 = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    h4, w4 = half(h2), half(w2)
    h8, w8 = half(h4), half(w4)
    h16, w16 = half(h8), half(w8)
    return (h2, w2), (h4, w4), (h8, w8), (h16, w16)


def gflop_per_frame(H=480, W=854):
    """Algorithmic FLOPs (2 x MAC) of one frame on the real feature-map sizes (not the padded canvases)."""
    (h2, w2), (h4, w4), (h8, w8), (h16, w16) = feature_sizes(H, W)
    size = {2: h2 * w2, 4: h4 * w4, 8: h8 * w8, 16: h16 * w16}
    total, res = 0.0, 2
    for c in convs():
        if c.name == "backbone.conv1":
            total += 2.0 * size[2] * c.cout * c.cin * 49
            res = 4
            continue
        if c.name.startswith("backbone.layer2.0.conv2") or c.name.startswith("backbone.layer2.0.downsample"):
            res = 8
        if c.name.startswith("backbone.layer3.0.conv2") or c.name.startswith("backbone.layer3.0.downsample"):
            res = 16
        if c.name.startswith("backbone.layer2.0.conv3"):
            res = 8
        if c.name.startswith("decoder") or c.name.startswith("embed"):
            res = 4
        px = 1 if c.name == "aspp.gap.conv" else size[res]
        total += 2.0 * px * c.cout * (c.cin // c.groups) * c.k * c.k
    return total / 1e9
