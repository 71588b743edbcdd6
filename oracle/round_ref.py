"""ORACLE (test infrastructure, not product code) — CPU restatement of the
per-round glue ``utils/utils_agent.py::recommend_frame``, setting='wild',
method='ours' (lines 111-122), on top of assess_ref / brain_ref.

Parity status: PINNED (tests/golden/*.npz hold mask_quality, Q and next_frame
from the reference's own recommend_frame + AssessNet + Agent run on CPU).
"""
import numpy as np
import torch

from . import assess_ref, brain_ref


def recommend_frame_wild_ours(assess_sd, brain_sd, all_F, all_P, annotated_frames_list,
                              mask_quality=None, dtype=torch.float32):
    """all_F: T x 3 x H x W, all_P: T x (O+1) x H x W (numpy or torch, fp32).
    Returns dict(next_frame, q, mask_quality, scores[T,O]).

    :112-113  annotated histogram (float64)
    :116-119  per object i: scores[:, i] = assess_net(all_F, all_P[:, i+1])
    :120      mask_quality[:] = scores.mean(1)            (float64 mean of fp32)
    :121      state = stack([mask_quality, annotated], 1) (T x 2 float64)
    :122      agent.action(state) -> fp32 cast, Brain, first-max argmax
    """
    all_F = torch.as_tensor(all_F)
    all_P = torch.as_tensor(all_P)
    T = all_F.shape[0]
    O = all_P.shape[1] - 1
    ann = np.zeros(T)
    for i in annotated_frames_list:
        ann[i] += 1
    pred = np.zeros((T, O))
    for i in range(O):
        s = assess_ref.assess_forward(assess_sd, all_F, all_P[:, i + 1], dtype)
        pred[:, i] = s.to(torch.float32).numpy() if dtype == torch.float32 else s.numpy()
    mq = pred.mean(1)
    if mask_quality is not None:
        mask_quality[:] = mq
    state = np.stack([mq, ann], 1)
    if dtype == torch.float64:
        q = brain_ref.brain_forward({k: v.numpy() for k, v in brain_sd.items()}, state[None], np.float64)[0]
        nxt = int(q.argmax())
    else:
        nxt, q = brain_ref.agent_action_greedy({k: v.numpy() for k, v in brain_sd.items()}, state)
    return {"next_frame": nxt, "q": q, "mask_quality": mq, "scores": pred}
