"""Seeded synthetic clips and weights.

No dataset, checkpoint or network access exists on the build or GPU boxes, so
every test / benchmark input is generated here, deterministically, from a seed:
``seed = 20210319 + clip_id``.  The generators are pure numpy / torch-CPU and
produce exactly the tensors the reference's ``recommend_frame`` receives
(utils/utils_agent.py:77-78): ``all_F`` T x 3 x H x W fp32 RGB in [0, 1] and
``all_P`` T x (O+1) x H x W fp32.

This module is data generation only; it is never on the timed product path.
"""
import math

import numpy as np
import torch

from . import arch

BASE_SEED = 20210319


def _bilinear_up(coarse, H, W):
    """coarse: (..., h, w) float32 -> (..., H, W), align_corners=True."""
    h, w = coarse.shape[-2:]
    ys = np.linspace(0.0, h - 1.0, H, dtype=np.float64)
    xs = np.linspace(0.0, w - 1.0, W, dtype=np.float64)
    y0 = np.clip(np.floor(ys).astype(np.int64), 0, h - 2)
    x0 = np.clip(np.floor(xs).astype(np.int64), 0, w - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a = coarse[..., y0[:, None], x0[None, :]]
    b = coarse[..., y0[:, None], x0[None, :] + 1]
    c = coarse[..., y0[:, None] + 1, x0[None, :]]
    d = coarse[..., y0[:, None] + 1, x0[None, :] + 1]
    out = (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + c * fy * (1 - fx) + d * fy * fx)
    return out.astype(np.float32)


def make_clip(clip_id, T, H, W, n_objects, style="manet"):
    """Returns (all_F, all_P, annotated_frames_list) as numpy arrays / list.

    style="manet": all_P = softmax over (O+1) channels (utils/utils_manet.py:161).
    style="atnet": channel 0 all-zero, channels 1..O independent sigmoids
                   (utils/utils_atnet.py:158-159) — rows do not sum to 1.
    One frame per clip has object 1 empty and one has it smaller than 128 px so
    the bbox special cases (models/assessment.py:119-136) are always exercised.
    """
    rng = np.random.default_rng(BASE_SEED + clip_id)
    O = n_objects
    # --- frames: smooth colour field, translated 4 px per frame -------------
    pad = 4 * T
    cw = max(14, int(math.ceil(14.0 * (W + pad) / W)))
    coarse = rng.random((3, 8, cw), dtype=np.float32)
    wide = _bilinear_up(coarse, H, W + pad)  # 3 x H x (W+pad)
    all_F = np.empty((T, 3, H, W), dtype=np.float32)
    for t in range(T):
        all_F[t] = wide[:, :, 4 * t:4 * t + W]
    all_F += (rng.random((T, 3, H, W), dtype=np.float32) - 0.5) * 0.04
    np.clip(all_F, 0.0, 1.0, out=all_F)
    # --- masks: moving ellipses ---------------------------------------------
    yy = np.arange(H, dtype=np.float32)[:, None]
    xx = np.arange(W, dtype=np.float32)[None, :]
    logits = np.empty((T, O + 1, H, W), dtype=np.float32)
    logits[:, 0] = 0.0
    t_empty = (3 + clip_id) % T
    t_small = (5 + 2 * clip_id) % T
    if t_small == t_empty:
        t_small = (t_small + 1) % T
    for o in range(O):
        cy0, cx0 = rng.uniform(0.3, 0.7) * H, rng.uniform(0.25, 0.75) * W
        vy, vx = rng.uniform(-1.5, 1.5), rng.uniform(-3.0, 3.0)
        ry0, rx0 = rng.uniform(20, 90), rng.uniform(20, 90)
        for t in range(T):
            ry = ry0 * (1.0 + 0.2 * math.sin(0.3 * t + o))
            rx = rx0 * (1.0 + 0.2 * math.cos(0.2 * t + o))
            if o == 0 and t == t_small:
                ry, rx = 9.0, 14.0
            cy, cx = cy0 + vy * t, cx0 + vx * t
            # signed distance-like field of the ellipse (px), >0 outside
            d = (np.sqrt(((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) - 1.0) * min(ry, rx)
            logits[t, o + 1] = -d / 8.0
            if o == 0 and t == t_empty:
                logits[t, o + 1] = -6.0
    logits += 0.5 * rng.standard_normal(logits.shape, dtype=np.float32)
    if style == "manet":
        m = logits.max(1, keepdims=True)
        e = np.exp(logits - m)
        all_P = (e / e.sum(1, keepdims=True)).astype(np.float32)
    elif style == "atnet":
        all_P = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
        all_P[:, 0] = 0.0
    else:
        raise ValueError(style)
    n_ann = 1 + clip_id % 3
    annotated = [int(v) for v in rng.integers(0, T, size=n_ann)]
    return all_F, all_P, annotated


def annotated_counts(annotated_frames_list, T):
    """utils/utils_agent.py:112-113: histogram of annotated frame indices (float64)."""
    c = np.zeros(T, dtype=np.float64)
    for i in annotated_frames_list:
        c[i] += 1
    return c


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------

def brain_state_dict(seed=0):
    """Seeded parameters for models.agent.Brain with torch's default init
    distributions (Linear: U(+-1/sqrt(fan_in)); LSTMCell: U(+-1/sqrt(hidden)))."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}
    for key, shape in arch.BRAIN_PARAMS:
        if key.startswith("lstm_cell"):
            bound = 1.0 / math.sqrt(128)
        else:
            layer = key.rsplit(".", 1)[0]
            fan_in = dict(arch.BRAIN_PARAMS)[layer + ".weight"][1]
            bound = 1.0 / math.sqrt(fan_in)
        sd[key] = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound * _BRAIN_GAIN.get(key, 1.0)
    return sd


# With torch's plain default init the Q-values of a random Brain are dominated by the
# frame position (argmax is frame 0 for every input), which would make an argmax-parity
# test vacuous.  These gains make Q depend on the quality scores: 57 distinct argmax
# positions over 200 random T=64 states, median top-1/top-2 gap 1.3e-2, min 4e-5.
_BRAIN_GAIN = {"encoder_fc1.weight": 10.0, "lstm_cell.weight_ih": 2.5, "lstm_cell.weight_hh": 2.5,
               "decoder_fc1.weight": 2.0}
# A random-weight ResNet-50 has nearly input-independent pooled features (score spread
# 3e-3); the FC gain widens the spread to ~0.3 and the bias recentres it (constant found
# once for seed 0 with tests/golden/make_golden.py) so scores look like IoUs in [0.3, 0.9].
_ASSESS_FC_GAIN = 40.0
_ASSESS_FC_BIAS = 2.194


def assess_state_dict(seed=0):
    """Seeded parameters for models.assessment.AssessNet (all keys of the strict
    state-dict contract).  Conv weights: N(0, 2/fan_out) as in torchvision's
    ResNet init; BatchNorm running stats / affine parameters are randomised
    around (0, 1) so that the BN epilogue is genuinely exercised; the last BN of
    every bottleneck is scaled by 0.25 so activations stay O(1) through the 16
    residual blocks and the quality scores land in a range the Q-network is
    sensitive to."""
    g = torch.Generator().manual_seed(2000 + seed)

    def conv(cout, cin, k):
        std = math.sqrt(2.0 / (k * k * cout))
        return torch.randn((cout, cin, k, k), generator=g, dtype=torch.float32) * std

    def bn(prefix, c, gain=1.0):
        return {
            prefix + ".weight": (0.8 + 0.4 * torch.rand(c, generator=g)) * gain,
            prefix + ".bias": 0.05 * torch.randn(c, generator=g),
            prefix + ".running_mean": 0.05 * torch.randn(c, generator=g),
            prefix + ".running_var": 0.8 + 0.4 * torch.rand(c, generator=g),
            prefix + ".num_batches_tracked": torch.tensor(0, dtype=torch.long),
        }

    sd = {
        "Encoder.mean": torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1),
        "Encoder.std": torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1),
        "Encoder.conv1_m.weight": conv(64, 1, 7),
        "Encoder.conv1_m.bias": torch.zeros(64),
        "Encoder.conv1_p.weight": conv(64, 1, 7),
        "Encoder.conv1_n.weight": conv(64, 1, 7),
        "Encoder.conv1.weight": conv(64, 3, 7),
    }
    sd.update(bn("Encoder.bn1", 64))
    for c in arch.resnet50_convs():
        sd[c.name + ".weight"] = conv(c.cout, c.cin, c.k)
        sd.update(bn(c.bn, c.cout, gain=0.25 if c.bn.endswith("bn3") else 1.0))
    bound = 1.0 / math.sqrt(2048)
    sd["fc1.weight"] = (torch.rand((1, 2048), generator=g) * 2 - 1) * bound * _ASSESS_FC_GAIN
    sd["fc1.bias"] = torch.tensor([_ASSESS_FC_BIAS])
    return sd


def manet_encoder_state_dict(seed=0):
    """Seeded parameters for the MANet feature extractor restatement (ivosw/manet_arch.py): He-normal convolutions,
    BatchNorm statistics / affine parameters randomised around (0, 1), the last BatchNorm of every bottleneck scaled by
    0.25 so that activations stay O(1) through 33 residual blocks."""
    from . import manet_arch
    g = torch.Generator().manual_seed(3000 + seed)
    sd = {}
    for c in manet_arch.convs():
        fan = c.k * c.k * c.cout
        sd[c.name + ".weight"] = torch.randn((c.cout, c.cin // c.groups, c.k, c.k), generator=g) * math.sqrt(2.0 / fan)
        if c.bias:
            sd[c.name + ".bias"] = 0.05 * torch.randn(c.cout, generator=g)
        gain = 0.25 if c.bn.endswith("bn3") else 1.0
        sd[c.bn + ".weight"] = (0.8 + 0.4 * torch.rand(c.cout, generator=g)) * gain
        sd[c.bn + ".bias"] = 0.05 * torch.randn(c.cout, generator=g)
        sd[c.bn + ".running_mean"] = 0.05 * torch.randn(c.cout, generator=g)
        sd[c.bn + ".running_var"] = 0.8 + 0.4 * torch.rand(c.cout, generator=g)
    return sd


def manet_frames(seed, B, H, W):
    """B x 3 x H x W frames as MANet's loader hands them to extract_feature (ImageNet-normalised RGB)."""
    all_F, _, _ = make_clip(seed, B, H, W, 1)
    mean = np.array([0.485, 0.456, 0.406], np.float32)[None, :, None, None]
    std = np.array([0.229, 0.224, 0.225], np.float32)[None, :, None, None]
    return ((all_F - mean) / std).astype(np.float32)
