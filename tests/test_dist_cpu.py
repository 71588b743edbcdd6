"""CPU: the frame-sharding protocol of ivosw.dist with world_size = 2 (and 3) over gloo — shard
ranges cover the clip exactly once, the single all-gather reassembles the float64 quality vector,
and every rank derives the identical recommended frame."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "ivos-w_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def test_shard_ranges_cover():
    from ivosw.dist import gather_layout, shard_range
    for T in (1, 5, 8, 64, 127, 128):
        for G in (1, 2, 3, 4, 8):
            seen = []
            for r in range(G):
                a, b = shard_range(T, G, r)
                assert 0 <= a <= b <= T
                seen += list(range(a, b))
            assert seen == list(range(T))
            per, padded = gather_layout(T, G)
            assert padded >= T and per * G == padded


def _worker(rank, world, port, T, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ivosw import synth
    from ivosw.dist import host_gather_round, shard_range
    from oracle import brain_ref
    rng = np.random.default_rng(5)
    mq_full = rng.uniform(0.2, 0.9, T)                      # stands in for the per-frame scores
    ann = synth.annotated_counts([1, T // 2], T)
    sd = {k: v.numpy() for k, v in synth.brain_state_dict(0).items()}
    a, b = shard_range(T, world, rank)

    def action(mq, ann_):
        return brain_ref.agent_action_greedy(sd, np.stack([mq, ann_], 1))[0]

    nf, mq = host_gather_round(mq_full[a:b], T, ann, action)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([[nf], mq]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 64), (2, 7), (3, 8)])
def test_gloo_sharded_round(tmp_path, world, T):
    port = 29500 + (os.getpid() + 17 * world + T) % 2000
    mp.spawn(_worker, args=(world, port, T, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    for r in res[1:]:
        np.testing.assert_array_equal(r, res[0])            # identical index and vector on every rank
    rng = np.random.default_rng(5)
    np.testing.assert_array_equal(res[0][1:], rng.uniform(0.2, 0.9, T))
