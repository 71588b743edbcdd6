"""GPU helper: Double-DQN update of BASELINE config C4 (batch 256 x T 25), data-parallel over the ranks of a torchrun launch
(256 / world samples per rank, one all-reduce of the 724 KB gradient between dqn_update(apply=False) and dqn_apply):
    python scripts/bench_dqn_dp.py                                   # one GPU, full batch
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29660 scripts/bench_dqn_dp.py
Device-timed (CUDA events), max over ranks; rank 0 prints one JSON line and checks that all ranks end with identical weights."""
import json
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200")); sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ivosw import dist as ivdist, synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402
from test_gpu_dqn import _dqn_batch  # noqa: E402

world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
N, T, K = 256, 25, 20
per = N // world
eng = Engine(lrank)
eng.load_brain(synth.brain_state_dict(0)); eng.load_target(synth.brain_state_dict(1)); eng.reset_optimizer()
s, ns, act, rs, rd = _dqn_batch(100, N, T)
sl = slice(rank * per, (rank + 1) * per)
dev = [torch.from_numpy(s[sl]).float().cuda(), torch.from_numpy(ns[sl]).float().cuda(), torch.from_numpy(act[sl]).cuda(),
       torch.from_numpy(rs[sl]).float().cuda(), torch.from_numpy(rd[sl]).float().cuda()]
for _ in range(3):
    ivdist.dqn_update_data_parallel(eng, *dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    loss = ivdist.dqn_update_data_parallel(eng, *dev)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
w = eng.brain_params("policy").clone()
same = torch.ones(1, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    w0 = w.clone(); dist.broadcast(w0, 0)
    same = torch.tensor([float(torch.equal(w, w0))], device="cuda"); dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"workload": "C4 DQN step, batch 256 x T 25, data-parallel", "n_gpus": world, "samples_per_gpu": per,
                      "ms_per_step": float(ms.item()), "samples_per_s": N / float(ms.item()) * 1e3, "loss": loss,
                      "weights_identical_on_all_ranks": bool(same.item() == 1.0)}))
if world > 1:
    dist.destroy_process_group()
