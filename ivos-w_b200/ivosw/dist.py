"""Frame-sharded scoring round over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink; gloo on CPU for the
host-logic tests).  Rank r owns the contiguous frame range ``shard_range(T, G, r)``
and scores it for every object; ONE exchange of the per-frame float64 mean quality
follows — the Q-network is a bidirectional LSTM over all frames, so the gather
must precede it — and every rank then runs Brain + argmax redundantly on the
identical vector (deterministic kernel -> identical index on every rank).

The exchange: after ``setup_peer_gather`` (one start-up all_gather of 64-byte IPC
handles through torch.distributed) it is two small kernels of the CUDA library that
write each rank's slice straight into every rank's buffer over NVLink peer memory
and wait on flags there (csrc/gather.cu) — no collective call and no host
synchronisation on the critical path.  Without it (or with ``IVOSW_PEER_GATHER=0``)
the same bytes go through ``all_gather_into_tensor`` (NCCL), which is also what the
gloo tests of the protocol use.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_range(T, world, rank):
    """Contiguous frame range of ``rank``: ceil(T / world) frames each, last shards may be short/empty."""
    per = -(-T // world)
    a = min(T, rank * per)
    return a, min(T, a + per)


def gather_layout(T, world):
    """(per, padded_T): the all-gather moves ``per`` doubles per rank."""
    per = -(-T // world)
    return per, per * world


def setup_peer_gather(engine, group=None, capacity=1024):
    """Start-up, once per process: allocate this rank's gather buffer, exchange the IPC handles, map the peers.
    Returns True when the peer path is active (CUDA engine, world > 1, not disabled)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1 or os.environ.get("IVOSW_PEER_GATHER", "1") == "0":
        return False
    mine = torch.frombuffer(bytearray(engine.gather_create(world, rank, capacity)), dtype=torch.uint8).to(engine.device)
    allh = torch.empty(64 * world, dtype=torch.uint8, device=engine.device)
    dist.all_gather_into_tensor(allh, mine, group=group)
    engine.gather_open(bytes(allh.cpu().numpy().tobytes()))
    dist.barrier(group)
    return True


def sharded_round(engine, all_F, all_P, annotated_counts, group=None):
    """all_F / all_P: this rank's FULL-clip tensors are not needed — only rows [a, b) are read, so callers
    may pass tensors whose other rows are uninitialised.  Returns (next_frame, q[T], mask_quality[T])."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    T = all_F.shape[0]
    per, padded = gather_layout(T, world)
    a, b = shard_range(T, world, rank)
    if engine.gather_ready:
        # peer-memory exchange: score into a local vector, post it into every rank's buffer, wait for all slices + Brain
        local = _local_mq(engine, per)
        if b > a:
            if all_F.is_cuda:
                engine.score_shard(all_F, all_P, a, b, local[: b - a])
            else:
                engine.score_shard_host(all_F, all_P, a, b, local[: b - a])
        engine.gather_post(local[: b - a] if b > a else None, a)
        nf, q, mq = engine.agent_action_gathered(annotated_counts)
        return nf, q, torch.from_numpy(mq)
    buf = torch.zeros(padded, dtype=torch.float64, device=engine.device)
    if b > a:
        if all_F.is_cuda:
            engine.score_shard(all_F, all_P, a, b, buf[rank * per: rank * per + (b - a)])
        else:   # host-resident clip: chunked upload of this rank's shard, overlapped with its scoring
            engine.score_shard_host(all_F, all_P, a, b, buf[rank * per: rank * per + (b - a)])
    if world > 1:
        dist.all_gather_into_tensor(buf, buf[rank * per:(rank + 1) * per].clone(), group=group)
    nf, q = engine.agent_action_dev(buf[:T], annotated_counts)
    return nf, q, buf[:T]


_LOCAL_MQ = {}


def _local_mq(engine, per):
    """Stable per-engine scratch for this rank's slice (a stable address keeps the CUDA-graph replay of score_shard)."""
    t = _LOCAL_MQ.get(id(engine))
    if t is None or t.numel() < per:
        t = torch.zeros(max(per, 64), dtype=torch.float64, device=engine.device)
        _LOCAL_MQ[id(engine)] = t
    return t


def host_gather_round(local_mq, T, annotated_counts, action_fn, group=None):
    """Backend-agnostic form of the same protocol on host vectors (used by the gloo CPU tests of
    the sharding logic): local_mq is this rank's float64 slice."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per, padded = gather_layout(T, world)
    send = torch.zeros(per, dtype=torch.float64)
    send[: len(local_mq)] = torch.as_tensor(np.asarray(local_mq), dtype=torch.float64)
    out = torch.zeros(padded, dtype=torch.float64)
    if world > 1:
        dist.all_gather_into_tensor(out, send, group=group)
    else:
        out[:per] = send
    mq = out[:T].numpy()
    return action_fn(mq, annotated_counts), mq


def dqn_update_data_parallel(engine, state, new_state, action, reward_step, reward_done, gamma=0.95, lr=5e-6,
                             weight_decay=5e-4, group=None):
    """BASELINE config C4, data-parallel: every rank passes ITS slice of the replay batch; raw gradients of
    the slice-mean losses are averaged with ONE all-reduce (724 KB), then every rank clamps and takes the
    identical Adam step.  Equal slice sizes reproduce the single-GPU full-batch update (agent.py:103-166).
    Returns the global mean loss."""
    loss, grads = engine.dqn_update(state, new_state, action, reward_step, reward_done, gamma=gamma, lr=lr,
                                    weight_decay=weight_decay, apply=False)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        lt = torch.tensor([loss], device=engine.device, dtype=torch.float32)
        dist.all_reduce(grads, group=group)
        dist.all_reduce(lt, group=group)
        grads /= world
        loss = float(lt.item()) / world
    engine.dqn_apply(grads, lr=lr, weight_decay=weight_decay)
    return loss


def assess_train_step_data_parallel(engine, imgs, probs, targets, valid, lr=5e-6, momentum=0.9, weight_decay=5e-4, group=None):
    """BASELINE config C5, data-parallel: every rank runs forward + backward of quality_assessment.py::train's loop body
    on ITS samples (csrc/train.cu, apply_update = 0), the raw gradients (23.5 M fp32 = 94 MB, blob order) are averaged with
    ONE NCCL all-reduce in place in the library's buffer, then every rank applies the identical accumulate + clamp + SGD
    update.  BatchNorm statistics are per rank (the reference is single-GPU: SURVEY.md §8(f) flags the difference).
    Returns the mean loss over ranks (None when no rank had a valid sample)."""
    loss, _ = engine.train_step(imgs, probs, targets, valid, lr=lr, momentum=momentum, weight_decay=weight_decay, apply=False)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    g = engine.train_grads_tensor()
    stat = torch.tensor([0.0 if loss is None else loss, 0.0 if loss is None else 1.0], device=engine.device, dtype=torch.float64)
    if world > 1:
        if loss is None:
            g.zero_()                       # a rank that skipped backward contributes nothing
        dist.all_reduce(g, group=group)
        dist.all_reduce(stat, group=group)
        g /= world
    if float(stat[1]) == 0.0:
        return None
    engine.train_apply(lr=lr, momentum=momentum, weight_decay=weight_decay)
    return float(stat[0] / stat[1])
