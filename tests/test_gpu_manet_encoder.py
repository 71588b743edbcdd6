"""GPU (-m gpu): the MANet feature extractor (csrc/manet_encoder.cu: ResNet-101 os16 + ASPP + decoder + embedding head on
the tcgen05 convolution kernel) against oracle/manet_encoder_ref.py.

PARITY UNPINNED: both sides restate the PUBLISHED architecture under the hyper-parameters the reference pins
(utils/config_manet/config.py:108-120); the upstream MANet source is not available, so agreement here means agreement
with that restatement.  Tolerance: 1e-4 + 1e-4 |ref| on the embedding (fp32-grade split-fp16
arithmetic through 111 convolutions), the bar north_star states for mask logits."""
import numpy as np
import pytest
import torch

from ivosw import manet_arch, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ivosw.engine import Engine
    e = Engine(0)
    e.load_manet_encoder(synth.manet_encoder_state_dict(0))
    yield e
    e.close()


@pytest.mark.parametrize("B,H,W", [(2, 96, 160), (1, 128, 224), (3, 120, 214)])
def test_encoder_vs_restatement(eng, B, H, W):
    """(120 x 214: odd feature-map sizes at every stride — 60 x 107, 30 x 54, 15 x 27, 8 x 14 — on canvases with unused
    columns and rows, stride-2 layers reading an odd-width map through the parity view)"""
    from oracle import manet_encoder_ref
    torch.set_num_threads(16)
    sd = synth.manet_encoder_state_dict(0)
    frames = synth.manet_frames(60 + B, B, H, W)
    ref = manet_encoder_ref.extract_feature(sd, torch.from_numpy(frames)).numpy()
    got = eng.manet_extract_feature(torch.from_numpy(frames).cuda()).cpu().numpy()
    (_, _), (h4, w4), _, _ = manet_arch.feature_sizes(H, W)
    assert got.shape == ref.shape == (B, manet_arch.EMBED_DIM, h4, w4)
    err = np.abs(got - ref)
    assert float(ref.max()) > 0.1                            # the head is alive (ReLU output)
    assert (err <= 1e-4 + 1e-4 * np.abs(ref)).all(), float(err.max())
    assert eng.saturation_count() == 0


def test_encoder_is_batch_invariant_and_deterministic(eng):
    frames = torch.from_numpy(synth.manet_frames(70, 11, 96, 160)).cuda()       # 11 frames: two passes (8 + 3)
    a = eng.manet_extract_feature(frames)
    b = eng.manet_extract_feature(frames)
    assert torch.equal(a, b)
    one = eng.manet_extract_feature(frames[9:10])
    assert torch.equal(one[0], a[9])
