"""Measurement helper (B200 box): a few MANet-encoder passes over 8 frames of 480x854 for ncu launch lists."""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
import torch  # noqa: E402
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402
eng = Engine(0)
eng.load_manet_encoder(synth.manet_encoder_state_dict(0))
x = torch.from_numpy(synth.manet_frames(80, 8, 480, 854)).cuda()
for _ in range(3):
    eng.manet_extract_feature(x)
torch.cuda.synchronize()
