"""Measurement helper (B200 box): BASELINE config C5 — one AssessNet optimisation step (quality_assessment.py::train
:240-269) on a batch of synthetic 480x854 samples, csrc/train.cu, next to the same step in stock PyTorch on the box's
host cores (the oracle restatement oracle/assess_train_ref.py, pinned to the reference by tests/golden/assess_train.npz).

    python scripts/bench_train.py [--batch 16] [--steps 5] [--cpu-batch 4]

With torchrun (N ranks): data-parallel form — every rank steps on its own batch with apply_update = 0, the raw gradients
(94 MB fp32) are all-reduced with NCCL, every rank applies the identical clamp + SGD update.  NB per-GPU BatchNorm
statistics: the reference is single-GPU, so at N > 1 the batch statistics are those of a rank's samples."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "ivos-w_b200")):
    sys.path.insert(0, p)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402

HP = dict(lr=5e-6, momentum=0.9, weight_decay=5e-4)        # configs/config.yaml:25-28


def batch(seed, n, H=480, W=854):
    all_F, all_P, _ = synth.make_clip(seed, n, H, W, 1)
    rng = np.random.default_rng(9000 + seed)
    return all_F, all_P[:, 1], rng.random(n).astype(np.float32), np.ones(n, bool)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU (config C5: 128 over 8 GPUs)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-batch", type=int, default=4)
    args = ap.parse_args()
    world, rank, lr_ = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr_)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
    eng = Engine(lr_)
    sd = synth.assess_state_dict(0)
    eng.train_begin(sd)
    imgs, probs, tg, valid = batch(rank, args.batch)
    F, P = torch.from_numpy(imgs).cuda(), torch.from_numpy(probs).cuda()
    from ivosw import dist as ivdist

    def step():
        if world == 1:
            return eng.train_step(F, P, tg, valid, **HP)[0]
        return ivdist.assess_train_step_data_parallel(eng, F, P, tg, valid, **HP)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / args.steps * 1e3
    out = {"config": "C5: AssessNet optimisation step, %d x 480x854 samples per GPU, %d GPU(s)" % (args.batch, world),
           "ms_per_step": ms, "samples_per_s": args.batch * world / ms * 1e3, "loss": loss,
           "algorithmic_gflop_per_sample": 3 * 10.779, "achieved_tflops": 3 * 10.779 * args.batch / ms}
    if rank == 0 and world == 1 and args.cpu_batch > 0:
        from oracle import assess_train_ref
        torch.set_num_threads(os.cpu_count() or 1)
        st = assess_train_ref.TrainState(sd)
        ci, cp, ct, cv = batch(0, args.cpu_batch)
        assess_train_ref.train_step(st, ci, cp, ct, cv, **HP)
        t0 = time.perf_counter()
        assess_train_ref.train_step(st, ci, cp, ct, cv, **HP)
        cs = time.perf_counter() - t0
        out["cpu_port"] = {"samples_per_s": args.cpu_batch / cs, "cores": os.cpu_count(), "batch": args.cpu_batch,
                           "what": "oracle/assess_train_ref.py (torch-CPU fp32 autograd), one step"}
        out["speedup_vs_cpu_port"] = out["samples_per_s"] / out["cpu_port"]["samples_per_s"]
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
