#!/usr/bin/env python
"""bench.py — frames/sec per interaction round of the IVOS-W frame-scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--conv-mode M]

A "step" is one scoring round (utils/utils_agent.py::recommend_frame, wild/ours) over one
synthetic 64-frame 480p clip with 2 objects (BASELINE.json configs[1], "C2"): bbox -> ROI crop ->
4-channel-stem ResNet-50 -> pool/FC for all T*O (frame, object) units, float64 object mean,
bi-LSTM Q-network, argmax.  The VOS backbone forward that north_star also names is NOT part of the
step: its source is absent from the reference tree (parity unpinned), so the metric
is the fully specified scoring round, as SURVEY.md §8(d)(i) defines it.

N > 1 (launched by torchrun, one rank per GPU): the clip's frames are sharded across ranks
(strong scaling: total work fixed), one NCCL all-gather of T float64 quality values, Brain replicated.

One JSON line on stdout (rank 0).  --impl reference times the CPU oracle port of the reference's
algorithm (oracle/, torch CPU fp32, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line.  Native libraries write to file descriptor 1 on their own (NCCL prints a
# "NCCL version ..." banner there when its first communicator is created), so fd 1 is pointed at stderr for the whole run
# and the JSON line goes to a private duplicate of the original stdout.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "ivos-w_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

T_FRAMES, HEIGHT, WIDTH, N_OBJ = 64, 480, 854, 2
WORKLOAD = "C2: 64-frame 480x854 synthetic clip (DAVIS-val shape), O=2, scoring round wild/ours"
GFLOP_PER_UNIT = 10.779          # AssessNet per (frame, object)
STEM_GFLOP_PER_UNIT = 0.411      # the 7x7 stem (own kernel, not part of the conv-stack roofline line)
CPU_SAMPLE_FRAMES = 16
# `dtype` = the arithmetic the encoder computes in: fp32 values carried as two fp16 planes, three kind::f16 tensor-core
# products per multiply, fp32 accumulation in TMEM (fp32-grade: 2^-22 relative); everything else fp32/fp64
DTYPE_NOTE = {"tc_fp16x3": "f16x3 (split-fp16 tensor-core products, f32 accumulate; f32-grade)",
              "tc_fp16x1": "f16 (single tensor-core product, f32 accumulate)", "simt_fp32": "f32"}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "which": "measured (MEASURED_PEAKS.json, bf16 sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "which": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load (B200_PROFILING.md).
    Started before the warm-up; samples are attributed to the timed region by wall-clock window and,
    when the timed region is shorter than a few sampling periods, the warm-up samples (same load)
    are used as well."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t_mark = [None, None]

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self, which):
        self.t_mark[which] = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(rows):
            sm, mx, reasons, pw = [], [], set(), []
            for _, r in rows:
                try:
                    sm.append(float(r[2])); mx.append(float(r[3])); pw.append(float(r[4]))
                except Exception:
                    continue
                for nm, v in zip(names, r[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons, pw

        t0, t1 = self.t_mark
        timed = [x for x in self.rows if t0 is not None and t1 is not None and t0 <= x[0] <= t1 + 0.03]
        window = "timed region"
        if len(timed) < 3:
            timed = [x for x in self.rows if t1 is None or x[0] <= t1 + 0.03]
            window = "warm-up + timed region"
        sm, mx, reasons, pw = collect(timed)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window,
                "power_w_max": max(pw) if pw else None}


def make_inputs(n_clips):
    from ivosw import synth
    clips = []
    for cid in range(n_clips):
        all_F, all_P, annotated = synth.make_clip(cid, T_FRAMES, HEIGHT, WIDTH, N_OBJ)
        clips.append((all_F, all_P, synth.annotated_counts(annotated, T_FRAMES), annotated))
    return clips


def cpu_oracle_round(assess_sd, brain_sd, clip, n_frames):
    """One oracle round on the first n_frames of a clip (bounded sample).  Returns seconds."""
    from oracle import round_ref
    all_F, all_P, _, annotated = clip
    ann = [a for a in annotated if a < n_frames] or [0]
    t0 = time.perf_counter()
    round_ref.recommend_frame_wild_ours(assess_sd, brain_sd, all_F[:n_frames], all_P[:n_frames], ann)
    return time.perf_counter() - t0


PARITY_SAMPLE_FRAMES = 16


def parity_block(eng, clips, dev_clips, world, rank, dev, assess_sd, brain_sd):
    """Correctness of THIS run, printed with its numbers.  Collective: every rank calls it.
      ranks_agree      every rank's recommended frame / gathered quality vector / Q are bit-identical
      vs_single_gpu    rank 0's sharded result against a 1-GPU round of the same clip on rank 0's device
      oracle           the CPU oracle (oracle/round_ref.py, pinned to the reference) on the first 16 frames of
                       clip 0: recommended frame identical, max|d score| <= 1e-4, max|d Q| <= 1e-5
    Untimed; runs after the timed regions."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from ivosw import dist as ivdist
    Fd, Pd, ann = dev_clips[0]
    if world == 1:
        r = eng.round_device(Fd, Pd, ann)
        nf, q, mq = r["next_frame"], r["q"], r["mask_quality"]
    else:
        nf, q, mq_t = ivdist.sharded_round(eng, Fd, Pd, ann)
        mq = mq_t.cpu().numpy()
    out = {}
    if world > 1:
        mine = torch.cat([torch.tensor([float(nf)], dtype=torch.float64), torch.from_numpy(np.asarray(mq, np.float64)),
                          torch.from_numpy(np.asarray(q, np.float64))]).to(dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        out["ranks_agree"] = bool(all(torch.equal(a, allr[0]) for a in allr))
    if rank != 0:
        return None
    all_F, all_P, ann_np, annotated = clips[0]
    if world > 1:       # rank 0 holds only its shard of the clip on the device: upload the rest for the 1-GPU round
        full = eng.round_device(torch.from_numpy(all_F).to(dev), torch.from_numpy(all_P).to(dev), ann)
        out["vs_single_gpu"] = {"next_frame_equal": int(full["next_frame"]) == int(nf),
                                "mask_quality_bitwise_equal": bool(np.array_equal(full["mask_quality"], mq)),
                                "q_bitwise_equal": bool(np.array_equal(full["q"], q))}
    n = PARITY_SAMPLE_FRAMES
    from oracle import round_ref
    from ivosw import synth
    torch.set_num_threads(os.cpu_count() or 1)
    ann_s = [a for a in annotated if a < n] or [0]
    ref = round_ref.recommend_frame_wild_ours(assess_sd, brain_sd, all_F[:n], all_P[:n], ann_s)
    got = eng.round_device(torch.from_numpy(all_F[:n]).to(dev), torch.from_numpy(all_P[:n]).to(dev),
                           synth.annotated_counts(ann_s, n), want_scores=True)
    qs = np.sort(ref["q"])[::-1]
    out["oracle"] = {"sample": "first %d frames of clip 0, O=%d" % (n, N_OBJ),
                     "next_frame_equal": int(got["next_frame"]) == int(ref["next_frame"]),
                     "next_frame": int(got["next_frame"]),
                     "max_abs_dscore": float(np.abs(got["scores"] - ref["scores"]).max()),
                     "max_abs_dq": float(np.abs(got["q"] - ref["q"]).max()),
                     "top2_q_gap": float(qs[0] - qs[1]), "tol": {"score": 1e-4, "q": 1e-5}}
    out["ok"] = bool(out["oracle"]["next_frame_equal"] and out["oracle"]["max_abs_dscore"] <= 1e-4 and
                     out["oracle"]["max_abs_dq"] <= 1e-5 and out.get("ranks_agree", True) and
                     all(out.get("vs_single_gpu", {"x": True}).values()))
    return out


def ref_gpu_pytorch_leg(clip, ours_ms, ours_e2e_ms):
    """The reference's single-GPU PyTorch path on this box (the '>= 5x' denominator):
    stock eager torch / cuDNN fp32 in the reference's call structure (scripts/ref_gpu_path.py restates it from
    utils/utils_agent.py:111-122, models/assessment.py:110-182, models/agent.py:33-64 because /root/reference does
    not travel): per-round pageable H2D of all_F, per-object AssessNet with the mask round trip and numpy bbox loop,
    Python-loop Brain.  cudnn.deterministic=True, allow_tf32=False (fp32, like the CPU oracle)."""
    import statistics as st
    import torch
    sys.path.insert(0, os.path.join(REPO, "scripts"))
    import ref_gpu_path as rg
    from ivosw import synth
    all_F, all_P, _, annotated = clip
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    res = {}
    assess, brain = rg.AssessNet(), rg.Brain()
    assess.load_state_dict(synth.assess_state_dict(0), strict=True)
    brain.load_state_dict(synth.brain_state_dict(0), strict=True)
    assess.cuda().eval(); brain.cuda().eval()
    F_cpu = torch.from_numpy(all_F)
    P_gpu = torch.from_numpy(all_P).cuda()
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(2):
            nf, mq = rg.reference_round(assess, brain, F_cpu, P_gpu, annotated, T_FRAMES, N_OBJ)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            nf, mq = rg.reference_round(assess, brain, F_cpu, P_gpu, annotated, T_FRAMES, N_OBJ)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ms = 1e3 * st.median(ts)
        res["tf32" if tf32 else "fp32"] = {"ms_per_round": ms, "frames_per_s": T_FRAMES * 1e3 / ms, "next_frame": int(nf)}
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    del assess, brain, P_gpu
    torch.cuda.empty_cache()
    res["what"] = ("stock eager PyTorch/cuDNN in the reference's call structure, wall clock around the round "
                   "(pageable H2D of all_F every round included, as utils_agent.py:114 does), median of 5")
    res["speedup_device_resident_vs_fp32"] = res["fp32"]["ms_per_round"] / ours_ms
    res["speedup_e2e_vs_fp32"] = res["fp32"]["ms_per_round"] / ours_e2e_ms
    return res


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the reference is Python and does not
    travel to the GPU box; oracle/ restates it on torch-CPU fp32 and is pinned to it by goldens)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from ivosw import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clips = make_inputs(1)
    assess_sd, brain_sd = synth.assess_state_dict(0), synth.brain_state_dict(0)
    # the stated config, whole: every step is one full round over all 64 frames x 2 objects (~3-4 s on 16 cores);
    # --ref-sample-frames N bounds it for slow hosts (then `same_config` is false and the line says so)
    n = args.ref_sample_frames if args.ref_sample_frames > 0 else T_FRAMES
    n = min(n, T_FRAMES)
    for _ in range(args.warmup):
        cpu_oracle_round(assess_sd, brain_sd, clips[0], n)
    times = [cpu_oracle_round(assess_sd, brain_sd, clips[0], n) for _ in range(args.steps)]
    total = sum(times)
    fps = n * args.steps / total
    sample = ("full workload: %d frames x %d objects per step" % (T_FRAMES, N_OBJ) if n == T_FRAMES else
              "%d of %d frames x %d objects per step" % (n, T_FRAMES, N_OBJ)) + \
        " (oracle port, torch-CPU fp32, %d threads)" % cores
    line = {
        "impl": "reference", "metric": "frames/sec per interaction round", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "same_config": n == T_FRAMES},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from ivosw import synth
    from ivosw import dist as ivdist
    from ivosw.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # clocks / throttle reasons: nvidia-smi needs a few hundred ms before its first sample, so it starts streaming now
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    eng = Engine(local_rank, args.conv_mode)
    peer_gather = ivdist.setup_peer_gather(eng) if world > 1 else False
    assess_sd, brain_sd = synth.assess_state_dict(0), synth.brain_state_dict(0)
    eng.load_assess(assess_sd)
    eng.load_brain(brain_sd)

    clips = make_inputs(2)
    a, b = ivdist.shard_range(T_FRAMES, world, rank)
    dev_clips, pin_clips = [], []
    for all_F, all_P, ann, _ in clips:
        # device-resident inputs for `value`; only this rank's frame shard is uploaded
        Fd = torch.empty((T_FRAMES, 3, HEIGHT, WIDTH), device=dev)
        Pd = torch.empty((T_FRAMES, N_OBJ + 1, HEIGHT, WIDTH), device=dev)
        Fd[a:b] = torch.from_numpy(all_F[a:b]).to(dev)
        Pd[a:b] = torch.from_numpy(all_P[a:b]).to(dev)
        dev_clips.append((Fd, Pd, ann))
        pin_clips.append((torch.from_numpy(all_F).pin_memory(), torch.from_numpy(all_P).pin_memory(), ann))

    def step_device(i):
        Fd, Pd, ann = dev_clips[i % len(dev_clips)]
        if world == 1:
            return eng.round_device(Fd, Pd, ann)["next_frame"]
        return ivdist.sharded_round(eng, Fd, Pd, ann)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up, then the timed region with clocks sampled under load
    # Per-stage CUDA-event timing (the roofline leg).  N == 1: recorded live inside the timed region (the
    # library syncs once per round, so the events of every CUDA-graph replay are read back).  N > 1: the
    # shard path is fully asynchronous, so the timed region runs untimed graphs and the stage times come
    # from a second pass of the same steps.
    live_stage_timing = world == 1
    eng.stage_timing(live_stage_timing)
    n_warm = max(4, args.warmup)          # every (clip, mode) CUDA graph is captured on its 2nd call
    for i in range(n_warm):
        step_device(i)
    launches0 = eng.launch_count
    eng.stage_times(reset=True)
    if sampler:
        sampler.mark(0)
    total_ms = timed(step_device, args.steps)
    if sampler:
        sampler.mark(1)
    launches = eng.launch_count - launches0
    if not live_stage_timing:
        eng.stage_timing(True)
        for i in range(2):
            step_device(i)
        eng.stage_times(reset=True)
        for i in range(args.steps):
            step_device(i)
        torch.cuda.synchronize()
    stage_ms, n_conv = eng.stage_times(reset=True)
    eng.stage_timing(False)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    value = T_FRAMES * 1e3 / ms_per_step

    # ---- end to end through the public host-buffer API (pinned host inputs, H2D inside the timed region)
    sent_bytes = []

    def step_e2e(i):
        Fh, Ph, ann = pin_clips[i % len(pin_clips)]
        nf = eng.round_host(Fh, Ph, ann)["next_frame"] if world == 1 else ivdist.sharded_round(eng, Fh, Ph, ann)[0]
        sent_bytes.append(eng.last_h2d_bytes())      # counted by the library from the copies it issued
        return nf

    for i in range(2):
        step_e2e(i)
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = timed(step_e2e, e2e_steps) / e2e_steps
    # the reference's caller holds all_F in PAGEABLE memory (torch.Tensor(np.stack(...)), eval_agent_manet.py:296-299):
    # the same call on pageable tensors, reported next to the pinned number
    page_clips = [(torch.from_numpy(F.copy()), torch.from_numpy(P.copy()), ann) for F, P, ann, _ in clips]

    def step_e2e_pageable(i):
        Fh, Ph, ann = page_clips[i % len(page_clips)]
        return eng.round_host(Fh, Ph, ann)["next_frame"] if world == 1 else ivdist.sharded_round(eng, Fh, Ph, ann)[0]

    step_e2e_pageable(0)
    e2e_page_ms = timed(step_e2e_pageable, e2e_steps) / e2e_steps

    # ---- parity inside the run: every N, every rank
    parity = None if args.no_parity else parity_block(eng, clips, dev_clips, world, rank, dev, assess_sd, brain_sd)
    # probability channel 0 (background) and the frame rows no ROI can touch are never read and not transferred:
    # this rank's bytes as counted by the library, mean over the timed steps (+ the annotated-count vector)
    h2d = int(sum(sent_bytes[-e2e_steps:]) / e2e_steps) + T_FRAMES * 8
    d2h = T_FRAMES * (8 + 4) + 4

    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel family: the res2..res5 convolution stack (tensor-bound)
    pk = peaks()
    units_local = (b - a) * N_OBJ
    conv_flops = units_local * (GFLOP_PER_UNIT - STEM_GFLOP_PER_UNIT) * 1e9 * args.steps
    conv_s = stage_ms["conv_stack"] / 1e3
    achieved = conv_flops / conv_s / 1e12 if conv_s > 0 else 0.0
    # DRAM bytes per launch come from an `ncu --set full` capture of this command at N=1 (128 units per launch,
    # profiles/conv_traffic.json names the capture); a shard moves fewer bytes per launch and has no capture of its
    # own, so the field is null at N>1 rather than a number that was not measured for this run
    traffic, traffic_source = None, None
    tpath = os.path.join(REPO, "profiles", "conv_traffic.json")
    if world == 1 and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_source = tj.get(args.conv_mode), tj.get("source")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": traffic_source,
                "kernel": "conv stack res2..res5 (%s), %d launches/step, avg %.1f us/launch" %
                          (args.conv_mode, n_conv // max(1, args.steps), 1e3 * stage_ms["conv_stack"] / max(1, n_conv)),
                "peak_source": pk["which"],
                # achieved / frac count ALGORITHMIC flops; the fp32-grade split-fp16 mode issues three MMA terms per
                # product, so the tensor pipe executes 3x that (what ncu's tensor-pipe % sees)
                "mma_terms": {"tc_fp16x3": 3, "tc_fp16x1": 1}.get(args.conv_mode),
                "executed_frac": ({"tc_fp16x3": 3, "tc_fp16x1": 1}.get(args.conv_mode, 0) * achieved / pk["tflops"]) or None,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()}}

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample (rank 0, N=1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_oracle_round(assess_sd, brain_sd, clips[0], 4)                      # warm-up
        reps = [cpu_oracle_round(assess_sd, brain_sd, clips[0], CPU_SAMPLE_FRAMES) for _ in range(3)]
        cpu_baseline = {"value": CPU_SAMPLE_FRAMES / statistics.median(reps), "unit": "frames/s", "cores": cores,
                        "kind": "port",
                        "sample": "%d of %d frames x %d objects, median of 3 (oracle port, torch-CPU fp32)" %
                                  (CPU_SAMPLE_FRAMES, T_FRAMES, N_OBJ)}

    line = {
        "metric": "frames/sec per interaction round", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(4, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": DTYPE_NOTE.get(args.conv_mode, "f32"), "data": "synthetic",
        "config": {"workload": WORKLOAD, "conv_mode": args.conv_mode, "frames_per_gpu": b - a,
                   "parallelism": ("frame-shard x%d + one exchange of T float64 (%s)" %
                                   (world, "own kernels over NVLink peer memory, csrc/gather.cu" if peer_gather else
                                    "NCCL all-gather")) if world > 1 else "single GPU",
                   "l2": "inputs (630 MB per clip, 2 clips alternating) exceed the 126 MB L2",
                   "launch": "CUDA-graph replay of the round (captured on the 2nd call per clip)",
                   "stage_timing": "live in the timed region" if world == 1 else "separate pass of the same steps"},
        "clocks": clocks,
        "e2e": {"value": T_FRAMES * 1e3 / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "host_memory": "pinned",
                "pageable": {"value": T_FRAMES * 1e3 / e2e_page_ms, "ms_per_step": e2e_page_ms,
                             "note": "same call on pageable host tensors (what the reference's caller holds)"}},
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_ref_gpu:
        try:
            line["ref_gpu_pytorch"] = ref_gpu_pytorch_leg(clips[0], ms_per_step, e2e_ms)
        except Exception as ex:            # the leg is a comparison, never a reason to lose the bench line
            line["ref_gpu_pytorch"] = {"unavailable": "%s: %s" % (type(ex).__name__, ex)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--conv-mode", default=os.environ.get("IVOSW_CONV_MODE", "tc_fp16x3"),
                    choices=["simt_fp32", "tc_fp16x3", "tc_fp16x1"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-style single-GPU PyTorch/cuDNN leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity block")
    ap.add_argument("--ref-sample-frames", type=int, default=0,
                    help="--impl reference: frames per step (0 = the full 64-frame clip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
