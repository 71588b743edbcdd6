"""Drop-in for the per-round glue of the reference's ``utils/utils_agent.py``:
``recommend_frame`` (77-128) with its helpers ``select_next_frame`` (38-74) and
``gen_subseq`` (131-157).  Same signature and return value; the 'wild' branches
that score frames ('ours', 'worst') run as ONE device-resident round in the CUDA
library instead of O AssessNet calls with host round-trips:

  * ``all_F`` (a CPU tensor in the reference, re-uploaded every round,
    utils_agent.py:104,114) is uploaded once per clip and cached (SURVEY A.Q8);
  * bbox / ROI / ResNet-50 / pooling / FC / float64 object-mean / Brain / argmax
    stay on the GPU; one small device->host copy returns mask_quality (which the
    caller's array must receive in place, A.Q9), Q and the index.

The RL training bookkeeping of the reference file (goal_only_reward,
agent_business, ...) is outside the scoring path and not reproduced here.
"""
import random
import weakref

import numpy as np
import torch

from ivosw.engine import get_engine

_CLIP_CACHE = {}   # device index -> (weakref to the caller's CPU tensor, its _version, device copy)


def _frames_on_device(all_F, device):
    """The clip is uploaded once and reused for the 8 rounds of a (sequence, scribble) visit.  The cache entry is
    tied to the IDENTITY of the caller's tensor object (weak reference) and its in-place version counter — not to its
    address: a new clip that the allocator places at a recycled address must never be served the old frames."""
    dev = torch.device(device)
    if all_F.is_cuda:
        return all_F
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    hit = _CLIP_CACHE.get(idx)
    if hit is not None and hit[0]() is all_F and hit[1] == all_F._version:
        return hit[2]
    t = all_F.to(dev, torch.float32, non_blocking=False)

    def _drop(ref, idx=idx):               # the source clip was collected: free its device copy too
        cur = _CLIP_CACHE.get(idx)
        if cur is not None and cur[0] is ref:
            del _CLIP_CACHE[idx]

    _CLIP_CACHE[idx] = (weakref.ref(all_F, _drop), all_F._version, t)
    return t


def select_next_frame(frame_value, metric='min', prev_frames=None):
    """utils_agent.py:38-74 ('prob' branch omitted: it references an un-imported name in the
    reference, SURVEY A.Q21).  Unhandled keywords such as 'worst' fall through to
    "argmin, skipping prev_frames", exactly as in the reference."""
    nb_frames = len(frame_value)
    if metric == 'random':
        return int(np.random.randint(nb_frames, size=1))
    if metric == 'uniform':
        assert prev_frames is not None
    if metric == 'max':
        frame_value = -frame_value
    if prev_frames is not None:
        value_idx = frame_value.argsort()
        i = 0
        while i < nb_frames and value_idx[i] in prev_frames:
            i += 1
        if i == nb_frames:
            return frame_value.argmin()
        return value_idx[i]
    return frame_value.argmin()


def gen_subseq(first_frame, n_frame, len_subseq, subseq_style='consecutive'):
    """utils_agent.py:131-157."""
    if subseq_style == 'consecutive':
        assert n_frame >= len_subseq
        i_start = max(0, (first_frame - len_subseq + 1))
        i_end = first_frame - max((first_frame + len_subseq) - n_frame, 0)
        i = int((i_start + i_end) / 2)
        return list(range(i, i + len_subseq))
    if subseq_style == 'equal':
        if n_frame < len_subseq + 1:
            return list(np.array(range(len_subseq)))
        subseq = np.linspace(0, n_frame - 1, num=len_subseq + 1).astype(int)
        while first_frame not in list(subseq):
            subseq += 1
        return list(subseq[:-1]) if first_frame != subseq[-1] else list(subseq[1:])
    raise NotImplementedError


def _score_round(assess_net, agent, device, all_F, all_P, annotated_counts, mask_quality, want_action):
    engine = get_engine(device)
    assess_net._sync(engine)
    if want_action:
        agent.policy_net._sync(engine)
    frames = _frames_on_device(all_F, device)
    probs = all_P if all_P.is_cuda else all_P.to(device)
    r = engine.round_device(frames, probs, annotated_counts, want_action=want_action)
    mask_quality[:] = r["mask_quality"]            # in-place contract (utils_agent.py:120)
    return r


def recommend_frame(cfg_yl, assess_net, agent, device, n_frame, n_objects, all_F, all_P, new_masks_quality,
                    prev_frames, annotated_frames_list, mask_quality, first_frame, max_nb_interactions):
    if cfg_yl.setting == 'oracle':
        if cfg_yl.method == 'worst':
            next_frame = select_next_frame(new_masks_quality, metric='worst', prev_frames=prev_frames)
        elif cfg_yl.method == 'ours':
            ann = np.zeros(len(new_masks_quality))
            for i in annotated_frames_list:
                ann[i] += 1
            next_frame = agent.action(np.stack([new_masks_quality, ann], 1))
        else:
            raise NotImplementedError
    elif cfg_yl.setting == 'wild':
        if cfg_yl.method == 'random':
            next_frame = select_next_frame(new_masks_quality, metric='random')
        elif cfg_yl.method == 'linspace':
            next_frame = prev_frames[0]
            len_subseq = min(max_nb_interactions, n_frame)
            for i in gen_subseq(first_frame, n_frame, len_subseq, 'equal'):
                if i not in prev_frames:
                    next_frame = i
                    break
        elif cfg_yl.method == 'worst':
            ann = np.zeros(len(new_masks_quality))
            _score_round(assess_net, agent, device, all_F, all_P, ann, mask_quality, want_action=False)
            next_frame = select_next_frame(mask_quality, metric='worst', prev_frames=prev_frames)
        elif cfg_yl.method == 'ours':
            ann = np.zeros(len(new_masks_quality))
            for i in annotated_frames_list:
                ann[i] += 1
            if agent.cfg.phase == 'train':
                # epsilon-greedy exploration is training-time behaviour: score on device, then the
                # generic Agent.action (which draws the RNG and may pick a random frame)
                _score_round(assess_net, agent, device, all_F, all_P, ann, mask_quality, want_action=False)
                next_frame = agent.action(np.stack([mask_quality, ann], 1))
            else:
                # eval: eps_threshold = 0 (agent.py:170-171).  Keep Agent.action's side effects — step
                # counter, exactly one RNG draw, the log line (agent.py:169,178-181; SURVEY A.Q6).
                agent.steps_done += 1
                rand_flag = random.random()
                greedy = rand_flag > 0
                print(f"step:{agent.steps_done}, rand_flag:{rand_flag:.4f}, eps_threshold:{0:.4f}, "
                      f"frame index was selected {'by agent' if greedy else 'randomly'}")
                r = _score_round(assess_net, agent, device, all_F, all_P, ann, mask_quality, want_action=greedy)
                if greedy:
                    next_frame = np.int64(r["next_frame"])
                else:   # rand_flag == 0.0 exactly: the reference's random branch (agent.py:194-196)
                    next_frame = random.choice(np.array(range(n_frame)))
        else:
            raise NotImplementedError
    else:
        raise NotImplementedError
    return next_frame
