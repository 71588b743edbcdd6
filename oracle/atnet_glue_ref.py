"""ORACLE (test infrastructure, not product code) — CPU restatement of the in-repo arithmetic of the ATNet round
wrapper ``utils/utils_atnet.py::run_VOS_singleiact`` (the ATNet networks themselves are external and absent,
SURVEY.md §8(c): parity for them is unpinned; this file covers only what the reference's own file computes).

Parity status: PINNED for the wrapper arithmetic — tests/golden/atnet_round.npz holds prob_map_of_frames, all_P and
output_masks produced by the reference's own ``run_VOS_singleiact`` (imported from /root/reference by
tests/golden/make_golden_atnet.py) driven with the labelled stand-in network / loader of tests/doubles;
tests/test_oracle_golden.py replays the same rounds through this restatement and requires bit equality on CPU.

Only tests/ may import this module.
"""
import numpy as np
import torch


def reflect_pad(planes, pad_info):
    """:95-96  torch.nn.ReflectionPad2d(pad_info[1] + pad_info[0]) — (left, right, top, bottom)."""
    (top, bottom), (left, right) = pad_info[0], pad_info[1]
    x = torch.as_tensor(planes)
    H, W = x.shape[-2:]
    ys = np.abs(np.arange(-top, H + bottom))
    ys = np.where(ys >= H, 2 * (H - 1) - ys, ys)
    xs = np.abs(np.arange(-left, W + right))
    xs = np.where(xs >= W, 2 * (W - 1) - xs, xs)
    return x[..., torch.as_tensor(ys)[:, None], torch.as_tensor(xs)[None, :]]


def blend_alpha(annotated_frames, annotated_now, frame, backward):
    """:127-147"""
    a = np.array(annotated_frames)
    if backward:
        side = a[a < annotated_now]
        if len(side) == 0:
            return 1
        c = np.max(side)
        return 0.5 + (1 - 0.5) * ((frame - c) / (annotated_now - c))
    side = a[a > annotated_now]
    if len(side) == 0:
        return 1
    c = np.min(side)
    return 0.5 + (1 - 0.5) * ((c - frame) / (c - annotated_now))


def sigmoid_blend(logit, prev, alpha):
    """:124-126, 146-147: prob = sigmoid(logit); blended = alpha*prob[:, 0] + (1-alpha)*prev."""
    prob = torch.sigmoid(torch.as_tensor(logit))
    p0 = prob[:, 0]
    if prev is None:
        return prob, p0
    return prob, (alpha * p0) + ((1 - alpha) * torch.as_tensor(prev))


def assemble_all_p(prob_map, hpad1, hpad2, wpad1, wpad2):
    """:157-159"""
    pm = torch.as_tensor(prob_map)
    return torch.cat([torch.zeros_like(pm[:, 0:1]), pm], 1)[:, :, hpad1:-hpad2, wpad1:-wpad2]


def run_round(net, frames_fn, planes, prop_list, annotated_frames, prob_map, pad_info, r5_3_list, r5_6_list):
    """The loop of :72-150 over an explicit frame source: frames_fn(idx) -> n_obj x 3 x P_H x P_W image tensor.
    ``planes`` = the un-padded n_obj x 3 x H x W scribble planes (:31-52, built by external helpers).
    prob_map (T x n_obj x P_H x P_W) is updated in place, as the reference does."""
    annotated_now = annotated_frames[-1]
    flag, adjacent = 0, False
    prob_anno = prob_prop = r2_prev = r2_anno = None
    planes = torch.as_tensor(planes)
    for frame in prop_list:
        image = frames_fn(frame)
        if frame == annotated_now:
            if flag == 0:
                flag, adjacent = 1, True
            elif flag == 1:
                flag, adjacent = 2, True
                continue
            else:
                raise NotImplementedError
            planes = reflect_pad(planes, pad_info)
            logit, r5_6 = net.forward_ANet(torch.cat([image, planes], 1))
            prob_anno, p0 = sigmoid_blend(logit, None, 1)
            r5_3, _, _, r2_anno = net.encoder_3ch.forward(image)
            r5_6_list.append(r5_6)
            r5_3_list.append(r5_3)
        else:
            if adjacent:
                r2_prev, pm_prev = r2_anno, prob_anno
            else:
                pm_prev = prob_prop
            adjacent = False
            logit, r2_prev = net.forward_TNet(r5_3_list, image, r5_6_list, r2_prev, pm_prev)
            alpha = blend_alpha(annotated_frames, annotated_now, frame, backward=(flag == 1))
            prob_prop, p0 = sigmoid_blend(logit, prob_map[frame], alpha)
        prob_map[frame] = p0
    return prob_map
