"""GPU helper: time of one Double-DQN update (BASELINE config C4: batch 256 x T 25) vs the CPU oracle."""
import os, sys, time, json
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200")); sys.path.insert(0, REPO)
from ivosw import synth
from ivosw.engine import Engine
from oracle import dqn_ref
sys.path.insert(0, os.path.join(REPO, "tests"))
from test_gpu_dqn import _dqn_batch

N, T = 256, 25
eng = Engine(0)
eng.load_brain(synth.brain_state_dict(0)); eng.load_target(synth.brain_state_dict(1)); eng.reset_optimizer()
s, ns, act, rs, rd = _dqn_batch(100, N, T)
dev = [torch.from_numpy(s).float().cuda(), torch.from_numpy(ns).float().cuda(), torch.from_numpy(act).cuda(),
       torch.from_numpy(rs).float().cuda(), torch.from_numpy(rd).float().cuda()]
for _ in range(3):
    eng.dqn_update(*dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 20
for _ in range(K):
    eng.dqn_update(*dev)
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / K
torch.set_num_threads(os.cpu_count())
st = dqn_ref.DqnState(synth.brain_state_dict(0), synth.brain_state_dict(1))
cpu = [torch.from_numpy(s).float(), torch.from_numpy(ns).float(), torch.from_numpy(act), torch.from_numpy(rs).float(), torch.from_numpy(rd).float()]
dqn_ref.update_agent(st, *cpu)
t0 = time.perf_counter()
for _ in range(3):
    dqn_ref.update_agent(st, *cpu)
cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
print(json.dumps({"workload": "C4 DQN step, batch 256 x T 25", "gpu_ms_per_step": gpu_ms, "gpu_samples_per_s": N / gpu_ms * 1e3,
                  "cpu_oracle_ms_per_step": cpu_ms, "cpu_cores": os.cpu_count()}))
