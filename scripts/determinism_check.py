"""GPU debug helper: repeats kernels on identical inputs and reports bitwise differences."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import arch, synth
from ivosw.engine import Engine

eng = Engine(0, "tc_fp16x3")
eng.load_assess(synth.assess_state_dict(0))
eng.load_brain(synth.brain_state_dict(0))
g = torch.Generator(device="cuda").manual_seed(1)
specs = arch.resnet50_convs()
B = 32
for li, sp in enumerate(specs):
    if li not in (2, 3, 6, 12, 13, 16, 25, 26, 29, 44, 45, 48, 1, 4, 11, 15):
        continue
    x = torch.randn((B, sp.in_hw, sp.in_hw, sp.cin), device="cuda", generator=g).relu_()
    res = torch.randn((B, sp.out_hw, sp.out_hw, sp.cout), device="cuda", generator=g) if sp.residual else None
    ref = eng.debug_conv(li, x, res, "tc_fp16x3").clone()
    bad = 0
    for it in range(30):
        y = eng.debug_conv(li, x, res, "tc_fp16x3")
        nd = int((y != ref).sum())
        if nd:
            bad += 1
            if bad == 1:
                idx = (y != ref).nonzero()[:5].tolist()
                print("  layer", li, sp.name, "iter", it, "diff elems", nd, "first", idx)
    print("layer %2d %-28s staged=%s  nondeterministic runs: %d/30" % (li, sp.name[8:], sp.k == 1 and sp.cout >= 256 and (sp.residual or "downsample" in sp.name), bad))

# whole network, probes
T, H, W, O = 64, 480, 854, 2
all_F, all_P, ann = synth.make_clip(0, T, H, W, O)
F_d, P_d = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
eng.enable_probes(True)
refs = None
for it in range(8):
    s = eng.assess_forward(F_d, P_d[:, 1])
    cur = [eng.probe(i).clone() for i in range(6)] + [s.clone()]
    if refs is None:
        refs = cur
    else:
        print("iter", it, "diff counts crop/pool/r2/r3/r4/r5/score:", [int((a != b).sum()) for a, b in zip(cur, refs)])
