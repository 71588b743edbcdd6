"""Measurement helper (B200 box): does a small unit batch (an 8-GPU frame shard, a chunk of the host-buffer path) finish
sooner as K concurrent sub-batches on K streams (one library context each) than as one pass?  Few-tile layers leave SMs
idle and a tile's K loop is a latency chain; a second stream can fill both."""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
import torch  # noqa: E402
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402

H, W, O = 480, 854, 2
KMAX = 4
engs = []
for _ in range(KMAX):
    e = Engine(0)
    e.load_assess(synth.assess_state_dict(0)); e.load_brain(synth.brain_state_dict(0))
    engs.append(e)
streams = [torch.cuda.Stream() for _ in range(KMAX)]
for T in (4, 8, 16, 32):
    all_F, all_P, annotated = synth.make_clip(0, T, H, W, O)
    F, P = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    mq = torch.zeros(T, dtype=torch.float64, device="cuda")
    ref = None
    line = "T=%2d (%3d units):" % (T, T * O)
    for K in (1, 2, 4):
        if T // K < 1:
            continue
        per = T // K

        def run():
            ev = torch.cuda.Event()
            ev.record()
            for k in range(K):
                streams[k].wait_event(ev)
                with torch.cuda.stream(streams[k]):
                    engs[k].score_shard(F, P, k * per, (k + 1) * per, mq[k * per:(k + 1) * per])
            for k in range(K):
                torch.cuda.current_stream().wait_stream(streams[k])
        for _ in range(4):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record(); torch.cuda.synchronize()
        got = mq.clone()
        if ref is None:
            ref = got
        line += "  K=%d %.3f ms%s" % (K, e0.elapsed_time(e1) / 20, "" if torch.equal(got, ref) else " (DIFFERENT)")
    print(line, flush=True)
