"""Measurement helper (B200 box): device time of one scoring pass over a small unit batch (an 8-GPU frame shard of
config C2 is 8 frames x 2 objects = 16 units) — per-layer kernels (IVOSW_STACK=0) vs the persistent stack kernel."""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
import torch  # noqa: E402
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402

H, W, O = 480, 854, 2
for T in (4, 8, 16, 32):
    all_F, all_P, annotated = synth.make_clip(0, T, H, W, O)
    F, P = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    ann = synth.annotated_counts(annotated, T)
    out = []
    for stack in ("0", "1"):
        os.environ["IVOSW_STACK"] = stack
        eng = Engine(0)
        eng.load_assess(synth.assess_state_dict(0)); eng.load_brain(synth.brain_state_dict(0))
        for _ in range(4):
            eng.round_device(F, P, ann)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            eng.round_device(F, P, ann)
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 20)
        eng.close()
    print("T=%2d (%3d units): per-layer kernels %.3f ms   stack kernel %.3f ms" % (T, T * O, out[0], out[1]))
