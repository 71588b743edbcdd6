"""GPU helper: per-layer-class time of the conv stack from stage-free direct timing (CUDA events per layer)."""
import os, sys, collections
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import arch, synth
from ivosw.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
eng = Engine(0, "tc_fp16x3")
eng.load_assess(synth.assess_state_dict(0))
g = torch.Generator(device="cuda").manual_seed(1)
specs = arch.resnet50_convs()
seen = collections.OrderedDict()
tot = 0.0
for li, sp in enumerate(specs):
    key = (sp.cin, sp.cout, sp.k, sp.stride, sp.out_hw, bool(sp.residual))
    if key in seen:
        seen[key][0] += 1
        continue
    x = torch.randn((B, sp.in_hw, sp.in_hw, sp.cin), device="cuda", generator=g).relu_()
    res = torch.randn((B, sp.out_hw, sp.out_hw, sp.cout), device="cuda", generator=g) if sp.residual else None
    # time only the conv kernel: debug_conv also splits/merges, so measure via library stage timing (stage 2 not used); use events around repeated calls minus split/merge baseline
    for _ in range(2):
        eng.debug_conv(li, x, res, "tc_fp16x3")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.debug_conv(li, x, res, "tc_fp16x3")
    e1.record(); torch.cuda.synchronize()
    seen[key] = [1, e0.elapsed_time(e1) / 5 * 1e3]
for key, (n, us) in seen.items():
    tot += n * us
    print("%4d %4d k%d s%d hw%2d res%d | n=%d | %8.1f us (incl. split/merge)" % (key + (n, us)))
print("sum over layers (incl. split/merge overhead): %.2f ms" % (tot / 1e3))
