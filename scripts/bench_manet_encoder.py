"""Measurement helper (B200 box): the MANet feature extractor (K8, restatement — see ivosw/manet_arch.py) at the clip
resolution of config C2 (480 x 854), csrc/manet_encoder.cu on the tcgen05 convolution kernel, next to the same
restatement in stock PyTorch / cuDNN on the same GPU (oracle/manet_encoder_ref.py run on CUDA tensors, fp32 and TF32).

    python scripts/bench_manet_encoder.py [--frames 16] [--reps 5]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "ivos-w_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ivosw import manet_arch, synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402
from oracle import manet_encoder_ref  # noqa: E402


def timed(fn, reps):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--H", type=int, default=480)
    ap.add_argument("--W", type=int, default=854)
    args = ap.parse_args()
    sd = synth.manet_encoder_state_dict(0)
    frames = torch.from_numpy(synth.manet_frames(80, args.frames, args.H, args.W)).cuda()
    gflop = manet_arch.gflop_per_frame(args.H, args.W)
    out = {"config": "MANet extract_feature (restatement), %d frames %dx%d" % (args.frames, args.H, args.W), "gflop_per_frame": gflop}
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    torch.backends.cudnn.deterministic = True
    ref = None
    with torch.no_grad():
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            # batch 1 per call, as eval_agent_manet.py:316-326 runs it (DataLoader batch_size=1) ...
            ms1 = timed(lambda: [manet_encoder_ref.extract_feature(sd_gpu, frames[i:i + 1]) for i in range(args.frames)], max(1, args.reps // 2))
            # ... and batched, the best stock PyTorch can do
            msb = timed(lambda: manet_encoder_ref.extract_feature(sd_gpu, frames), args.reps)
            r = manet_encoder_ref.extract_feature(sd_gpu, frames)
            if not tf32:
                ref = r
            out["torch_cudnn_%s" % ("tf32" if tf32 else "fp32")] = {
                "frames_per_s_batch1_calls": args.frames / ms1 * 1e3, "frames_per_s_batched": args.frames / msb * 1e3,
                "max_abs_diff_vs_fp32": float((r - ref).abs().max())}
    for mode in ("tc_fp16x3", "tc_fp16x1"):
        eng = Engine(0, mode)
        eng.load_manet_encoder(sd)
        ms = timed(lambda: eng.manet_extract_feature(frames), args.reps)
        got = eng.manet_extract_feature(frames)
        out[mode] = {"ms_per_frame": ms / args.frames, "frames_per_s": args.frames / ms * 1e3,
                     "algorithmic_tflops": gflop * args.frames / ms, "max_abs_diff_vs_cudnn_fp32": float((got - ref).abs().max()),
                     "max_ref": float(ref.abs().max())}
        eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
