"""Static description of the two in-repo networks on the frame-scoring path.

The layer tables here are the single source of truth for (a) the state-dict key
names the drop-in modules must accept with ``strict=True`` (reference:
models/assessment.py:12-45, 66-71 and models/agent.py:14-31), (b) the order in
which parameters are packed into the flat blobs that cross the C ABI
(include/ivosw_b200.h), and (c) the synthetic weight generator.

AssessNet's encoder is torchvision's ResNet-50 (bottleneck [3,4,6,3], stride on
the 3x3 conv) re-exposed as ``Encoder.res2..res5`` (= ``layer1..layer4``) with a
4-channel stem: ``conv1`` (RGB) + ``conv1_p`` (probability plane), summed before
``bn1`` (models/assessment.py:47-55).
"""
from collections import namedtuple

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by torchvision resnet50

# name: state-dict prefix of the conv ("...weight"), bn: prefix of its BatchNorm
ConvSpec = namedtuple(
    "ConvSpec", "name bn cin cout k stride pad relu residual in_hw out_hw")

RES_STAGES = (  # (reference attribute, planes, blocks, stride of first block)
    ("res2", 64, 3, 1),
    ("res3", 128, 4, 2),
    ("res4", 256, 6, 2),
    ("res5", 512, 3, 2),
)


def resnet50_convs(in_hw=64):
    """The 52 convs of res2..res5 in execution order (stem excluded).

    ``residual``: "" (none), "identity" (block input is added) or "downsample"
    (output of the block's downsample conv is added).  The downsample conv of a
    block is listed *before* the block's conv3 so its output exists when conv3's
    epilogue needs it.
    """
    convs = []
    inplanes = 64
    hw = in_hw
    for stage, planes, blocks, stride in RES_STAGES:
        for b in range(blocks):
            s = stride if b == 0 else 1
            p = "Encoder.%s.%d." % (stage, b)
            out_hw = hw // s
            convs.append(ConvSpec(p + "conv1", p + "bn1", inplanes, planes, 1, 1, 0, True, "", hw, hw))
            convs.append(ConvSpec(p + "conv2", p + "bn2", planes, planes, 3, s, 1, True, "", hw, out_hw))
            if b == 0:
                convs.append(ConvSpec(p + "downsample.0", p + "downsample.1", inplanes, planes * 4,
                                      1, s, 0, False, "", hw, out_hw))
                res = "downsample"
            else:
                res = "identity"
            convs.append(ConvSpec(p + "conv3", p + "bn3", planes, planes * 4, 1, 1, 0, True, res, out_hw, out_hw))
            inplanes = planes * 4
            hw = out_hw
    return convs


def assess_state_dict_keys():
    """Every key of ``AssessNet().state_dict()`` in the reference, in order of
    registration (models/assessment.py:15-45, 69-71).  Unused-but-registered
    parameters (conv1_m, conv1_n) are part of the strict contract (SURVEY A.Q5).
    """
    keys = ["Encoder.mean", "Encoder.std",
            "Encoder.conv1_m.weight", "Encoder.conv1_m.bias",
            "Encoder.conv1_p.weight", "Encoder.conv1_n.weight",
            "Encoder.conv1.weight"]
    keys += ["Encoder.bn1." + s for s in BN_FIELDS]
    for c in resnet50_convs():
        keys.append(c.name + ".weight")
        keys += [c.bn + "." + s for s in BN_FIELDS]
    # torchvision registers conv1,bn1,conv2,bn2,conv3,bn3,downsample per block;
    # ordering is irrelevant for load_state_dict, only the set matters.
    keys += ["fc1.weight", "fc1.bias"]
    return keys


BN_FIELDS = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")

BRAIN_PARAMS = (  # (key, shape) — models/agent.py:19-29
    ("encoder_fc1.weight", (128, 2)),
    ("encoder_fc1.bias", (128,)),
    ("encoder_fc2.weight", (128, 128)),
    ("encoder_fc2.bias", (128,)),
    ("lstm_cell.weight_ih", (512, 128)),
    ("lstm_cell.weight_hh", (512, 128)),
    ("decoder_fc1.weight", (128, 256)),
    ("decoder_fc1.bias", (128,)),
    ("decoder_fc2.weight", (1, 128)),
    ("decoder_fc2.bias", (1,)),
)
BRAIN_NUM_PARAMS = sum(int(__import__("math").prod(s)) for _, s in BRAIN_PARAMS)  # 180 993

ROI_SIZE = 256  # models/assessment.py:170  dst_size=(256, 256)
ASSESS_GFLOP_PER_UNIT = 10.779  # per (frame, object); SURVEY.md §8(d), flop_counter on the reference module
