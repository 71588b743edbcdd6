"""Measurement helper (B200 box): per-layer time attribution inside conv_stack_kernel (IVOSW_STACK_PROFILE=1)."""
import os
import sys
os.environ["IVOSW_STACK_PROFILE"] = "1"
os.environ["IVOSW_GRAPHS"] = "0"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
import torch  # noqa: E402
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402

T, H, W, O = 64, 480, 854, 2
all_F, all_P, annotated = synth.make_clip(0, T, H, W, O)
eng = Engine(0)
eng.load_assess(synth.assess_state_dict(0)); eng.load_brain(synth.brain_state_dict(0))
F, P = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
ann = synth.annotated_counts(annotated, T)
for i in range(3):
    if i == 2:
        sys.stderr.write("==== run %d\n" % i)
    eng.round_device(F, P, ann)
