"""GPU helper: kernel-only time of selected conv layers under the IVOSW_TC_DEBUG measurement switches
(0 normal, 1 no TMA operand loads, 2 no MMAs, 4 no staged-epilogue work; bits combine).
The kernel time is the difference between 11 and 1 back-to-back launches inside ivosw_debug_conv."""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import arch, synth
from ivosw.engine import Engine
B = int(os.environ.get("TC_CEILING_B", "128"))
LAYERS = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [15, 28, 47, 27, 6, 16, 29, 48, 12, 25, 44]
FLAGS = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
eng = Engine(0, "tc_fp16x3")
eng.load_assess(synth.assess_state_dict(0))
specs = arch.resnet50_convs()
g = torch.Generator(device="cuda").manual_seed(1)


def timed(li, x, res, reps):
    os.environ["IVOSW_DEBUG_CONV_REPEAT"] = str(reps)
    for _ in range(2):
        eng.debug_conv(li, x, res, "tc_fp16x3")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.debug_conv(li, x, res, "tc_fp16x3")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1000.0


print("| layer | " + " | ".join("dbg %d" % f for f in FLAGS) + " |")
print("|---|" + "---|" * len(FLAGS))
for li in LAYERS:
    sp = specs[li]
    x = torch.randn((B, sp.in_hw, sp.in_hw, sp.cin), device="cuda", generator=g).relu_()
    res = torch.randn((B, sp.out_hw, sp.out_hw, sp.cout), device="cuda", generator=g) if sp.residual else None
    cells = []
    for f in FLAGS:
        os.environ["IVOSW_TC_DEBUG"] = str(f)
        cells.append("%.1f" % ((timed(li, x, res, 11) - timed(li, x, res, 1)) / 10))
    print("| %2d %s %dx%d %d->%d @%d^2%s | " % (li, sp.name[8:], sp.k, sp.k, sp.cin, sp.cout, sp.out_hw, " +res" if sp.residual else "")
          + " | ".join(cells) + " |", flush=True)
