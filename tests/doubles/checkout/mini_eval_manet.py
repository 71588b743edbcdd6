"""TEST DOUBLE of an IVOS-W entry script (NOT the reference's eval_agent_manet.py): the same import block for the
in-repo modules (eval_agent_manet.py:24-36) and one `while sess.next()` style loop of two interaction rounds
(:270-425: rough_ROI on the first round, get_results, recommend_frame) over a synthetic clip with a stand-in
session and a stand-in IntVOS.  Started by tests through `python -m ivosw.run`; writes what happened as JSON.
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.append(os.path.join('utils', 'config_manet'))
sys.path.append(os.path.join('VOS', 'MANet'))
from utils.misc import (set_random_seed, load_agent_checkpoint, load_network_checkpoint)
from utils.utils_agent import recommend_frame
from models.agent import Agent
from models.assessment import AssessNet

from config import cfg
from utils.utils_manet import (load_network, rough_ROI, preprocess, get_results)
from networks.IntVOS import IntVOS


def main(out_path):
    from ivosw import synth                                    # synthetic clip + weights (data generation only)
    T, H, W, O = 8, 128, 224, 2
    h, w = H // 4, W // 4
    device = torch.device("cuda:0")
    set_random_seed(0)
    cfg_yl = SimpleNamespace(phase="eval", setting="wild", method="ours",
                             agent=SimpleNamespace(memory_size=10, gamma=0.95, eps_start=0.7, eps_end=0.25, eps_decay=500,
                                                   update_rate=0.05, lr=5e-6, weight_decay=5e-4),
                             data=SimpleNamespace(subset="val"))
    agent = Agent(device, cfg_yl)
    assert load_agent_checkpoint(agent, synth.brain_state_dict(0), device, strict=True)
    assess_net = load_network_checkpoint(synth.assess_state_dict(0), AssessNet(), device='cpu').to(device).eval()
    all_F_np, all_P_np, annotated = synth.make_clip(5, T, H, W, O)
    all_F = torch.Tensor(all_F_np)                             # CPU, as eval_agent_manet.py:297-300 builds it
    # stand-in network output: log-probabilities of the synthetic masks at embedding resolution
    logits = torch.log(torch.from_numpy(all_P_np[:, :, ::4, ::4]).clamp_min(1e-6)).contiguous().to(device)
    model = IntVOS(cfg, None, logits=logits)
    load_network(torch.nn.Linear(2, 2), {})                    # the checkout's own helper must still be there
    assert preprocess('', ['synthetic'], os.devnull) == {'synthetic': [1]}
    embedding_memory = torch.zeros((T, 4, h, w), device=device)
    prev_label_storage = torch.zeros((T, H, W), device=device)
    mask_quality = np.zeros(T)
    next_frame, annotated_frames_list, rounds = int(annotated[0]), [], []
    for n_interaction in (1, 2):                               # two rounds of the stand-in session
        first = n_interaction == 1
        annotated_frames_list.append(next_frame)
        scribble_label = -torch.ones((1, 1, h, w), device=device)
        scribble_label[0, 0, 5:12, 8:20] = 1.0
        if first:
            scribble_label = rough_ROI(scribble_label)
        final_masks, all_P = get_results(model, embedding_memory[next_frame:next_frame + 1], scribble_label, None, {},
                                         ({}, {}), n_interaction, 'synthetic', O, next_frame, first, H, W,
                                         prev_label_storage, T, embedding_memory)
        nf = recommend_frame(cfg_yl, assess_net, agent, device, n_frame=T, n_objects=O, all_F=all_F, all_P=all_P,
                             new_masks_quality=np.zeros(T), prev_frames=[next_frame],
                             annotated_frames_list=annotated_frames_list, mask_quality=mask_quality, first_frame=0,
                             max_nb_interactions=8)
        rounds.append({"annotated": list(annotated_frames_list), "next_frame": int(nf),
                       "mask_quality": [float(v) for v in mask_quality],
                       "masks_sum": float(final_masks.sum().item()), "all_P_shape": list(all_P.shape)})
        next_frame = int(nf)
    from ivosw import hook
    with open(out_path, "w") as f:
        json.dump({"rounds": rounds, "modules": hook.report(), "steps_done": agent.steps_done,
                   "misc_file": sys.modules["utils.misc"].__file__,
                   "momory_pool_file": sys.modules["models.momory_pool"].__file__,
                   "argv0": sys.argv[0], "name": __name__}, f, default=str)


if __name__ == "__main__":
    main(sys.argv[1])
