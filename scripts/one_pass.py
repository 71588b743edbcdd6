"""Measurement helper (B200 box): a few scoring rounds over T frames (O = 2) for ncu launch lists.  usage: one_pass.py T"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
import torch  # noqa: E402
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W, O = 480, 854, 2
all_F, all_P, annotated = synth.make_clip(0, T, H, W, O)
F, P = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
ann = synth.annotated_counts(annotated, T)
eng = Engine(0)
eng.load_assess(synth.assess_state_dict(0)); eng.load_brain(synth.brain_state_dict(0))
for _ in range(int(os.environ.get("ROUNDS", "3"))):
    eng.round_device(F, P, ann)
torch.cuda.synchronize()
