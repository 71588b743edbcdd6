"""GPU helper: run-to-run behaviour of chosen conv layers (tensor-core path) against the fp32 CUDA-core kernel."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import arch, synth
from ivosw.engine import Engine
eng = Engine(0, "tc_fp16x3")
eng.load_assess(synth.assess_state_dict(0))
specs = arch.resnet50_convs()
B = int(os.environ.get("PD_B", "48"))
for li in [int(a) for a in sys.argv[1].split(",")]:
    sp = specs[li]
    for use_res in ((True, False) if sp.residual else (False,)):
        g = torch.Generator(device="cuda").manual_seed(7 + li)
        x = torch.randn((B, sp.in_hw, sp.in_hw, sp.cin), device="cuda", generator=g).relu_()
        res = torch.randn((B, sp.out_hw, sp.out_hw, sp.cout), device="cuda", generator=g) if use_res else None
        simt = eng.debug_conv(li, x, res, "simt_fp32").clone()
        scale = float(simt.abs().max())
        errs = []
        for _ in range(8):
            y = eng.debug_conv(li, x, res, "tc_fp16x3")
            d = (y - simt).abs()
            bad = (d > 1e-3 * scale)
            first_bad = int(bad.flatten().nonzero()[0]) if bad.any() else -1
            errs.append("%.2g/%d@%d" % (float(d.max()) / scale, int(bad.sum()), first_bad))
        print("layer %2d %-28s B=%d res=%d  rel.err/bad count@first bad index per run: %s" % (li, sp.name[8:], B, use_res, " ".join(errs)), flush=True)
