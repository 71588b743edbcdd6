"""GPU (-m gpu, needs >= 2 devices; skipped otherwise): the frame-sharded round over NCCL gives, on every rank,
exactly the index / Q / quality vector of the single-GPU round; BASELINE config C3 shape
(ATNet-style probabilities: channel 0 all-zero, rows not summing to 1)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, T, H, W, O, peer):
    for p in (REPO, os.path.join(REPO, "ivos-w_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from ivosw import dist as ivdist
    from ivosw import synth
    from ivosw.engine import Engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = Engine(rank)
    os.environ["IVOSW_PEER_GATHER"] = "1" if peer else "0"
    assert ivdist.setup_peer_gather(eng) == bool(peer)         # peer: csrc/gather.cu over cudaIpc-mapped NVLink memory
    eng.load_assess(synth.assess_state_dict(0))
    eng.load_brain(synth.brain_state_dict(0))
    all_F, all_P, annotated = synth.make_clip(21, T, H, W, O, "atnet")
    ann = synth.annotated_counts(annotated, T)
    a, b = ivdist.shard_range(T, world, rank)
    Fd = torch.zeros((T, 3, H, W), device="cuda"); Pd = torch.zeros((T, O + 1, H, W), device="cuda")
    Fd[a:b] = torch.from_numpy(all_F[a:b]).cuda(); Pd[a:b] = torch.from_numpy(all_P[a:b]).cuda()   # only the shard is resident
    res = []
    for _ in range(3):      # eager, graph capture, graph replay
        nf, q, mq = ivdist.sharded_round(eng, Fd, Pd, ann)
        res.append((nf, q.copy(), mq.cpu().numpy().copy()))
    nf_h, q_h, mq_h = ivdist.sharded_round(eng, torch.from_numpy(all_F).pin_memory(), torch.from_numpy(all_P).pin_memory(), ann)
    res.append((nf_h, q_h.copy(), mq_h.cpu().numpy().copy()))
    if rank == 0:
        full = eng.round_device(torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda(), ann)
        np.savez(os.path.join(out_dir, "full.npz"), nf=full["next_frame"], q=full["q"], mq=full["mask_quality"])
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), nf=np.array([r[0] for r in res]), q=np.stack([r[1] for r in res]),
             mq=np.stack([r[2] for r in res]))
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("T,peer", [(16, 1), (13, 1), (16, 0), (1, 1)])
def test_sharded_round_matches_single_gpu(tmp_path, T, peer):
    """T = 13: ragged shards; T = 1: the second rank's shard is empty (it still has to post its flag)."""
    world = 2
    port = 29700 + (os.getpid() + 7 * T + peer) % 200
    mp.spawn(_worker, args=(world, port, str(tmp_path), T, 240, 432, 2, peer), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npz")
    for r in range(world):
        d = np.load(tmp_path / ("r%d.npz" % r))
        for i in range(d["nf"].shape[0]):
            assert int(d["nf"][i]) == int(full["nf"])
            np.testing.assert_array_equal(d["mq"][i], full["mq"])      # scoring is per-unit: bit-identical
            np.testing.assert_array_equal(d["q"][i], full["q"])
