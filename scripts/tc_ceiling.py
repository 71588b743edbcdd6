"""GPU helper: time selected conv layers with IVOSW_TC_DEBUG = 0 (normal), 1 (no TMA loads), 2 (no MMAs)."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import arch, synth
from ivosw.engine import Engine
B = 128
eng = Engine(0, "tc_fp16x3")
eng.load_assess(synth.assess_state_dict(0))
specs = arch.resnet50_convs()
g = torch.Generator(device="cuda").manual_seed(1)
for li in (15, 28, 47, 27, 6):
    sp = specs[li]
    x = torch.randn((B, sp.in_hw, sp.in_hw, sp.cin), device="cuda", generator=g).relu_()
    res = torch.randn((B, sp.out_hw, sp.out_hw, sp.cout), device="cuda", generator=g) if sp.residual else None
    for _ in range(2):
        eng.debug_conv(li, x, res, "tc_fp16x3")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.debug_conv(li, x, res, "tc_fp16x3")
    e1.record(); torch.cuda.synchronize()
    print("dbg=%s layer %2d %-22s k%d cin%4d cout%4d hw%2d : %.1f us per call (incl. split/merge helpers)" %
          (os.environ.get("IVOSW_TC_DEBUG", "0"), li, sp.name[8:], sp.k, sp.cin, sp.cout, sp.out_hw, e0.elapsed_time(e1) * 100))
