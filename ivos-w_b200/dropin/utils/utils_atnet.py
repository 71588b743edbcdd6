"""Drop-in for ``utils/utils_atnet.py::run_VOS_singleiact`` (14-161) — the ATNet round wrapper (config C3).

The ATNet networks (``net.forward_ANet`` / ``net.forward_TNet`` / ``net.encoder_3ch``), ``libs.utils`` /
``libs.utils_torch`` and the ``DAVIS2017`` loader are external (yuk6heo/IVOS-ATNet, not part of the reference
tree) and are called exactly as the reference calls them.  What changes is the wrapper's own arithmetic, which
runs in the CUDA library (csrc/atnet_glue.cu) instead of as a chain of ATen temporaries:

  * the scribble planes' ``ReflectionPad2d`` (95-96)                     -> ``ivosw_atnet_reflect_pad``
  * ``sigmoid`` of the logits and the alpha-blend against the previous round's probabilities, written straight
    into ``prob_map_of_frames[frame]`` (101-102, 124-126, 128-150)       -> ``ivosw_atnet_sigmoid_blend``
  * ``all_P = cat([zeros, prob_map_of_frames], 1)`` un-padded (157-159)  -> ``ivosw_atnet_assemble`` (one pass,
    contiguous result; the reference returns a strided view of a T x (O+1) x P_H x P_W concatenation)

This module is never imported under its own name by the launcher: ``ivosw.hook`` lets the reference's
``utils.utils_atnet`` import normally (so ``datasets`` / ``libs`` / ``DataLoader`` resolve as they always did),
stores that module here as ``_ref`` and replaces its ``run_VOS_singleiact`` attribute with the one below.
"""
import numpy as np
import torch

from ivosw.engine import get_engine

_ref = None     # the reference's own utils.utils_atnet module (set by ivosw.hook, or a stand-in namespace in tests)


def _blend_alpha(annotated_frames_np, annotated_now, frame, backward):
    """utils_atnet.py:128-147: 1 when no other annotated frame lies on this side, else 0.5 .. 1 linearly in the
    distance from that frame (Python float arithmetic, as the reference)."""
    smallest_alpha = 0.5
    if backward:
        side = annotated_frames_np[annotated_frames_np < annotated_now]
        if len(side) == 0:
            return 1
        closest = np.max(side)
        return smallest_alpha + (1 - smallest_alpha) * ((frame - closest) / (annotated_now - closest))
    side = annotated_frames_np[annotated_frames_np > annotated_now]
    if len(side) == 0:
        return 1
    closest = np.min(side)
    return smallest_alpha + (1 - smallest_alpha) * ((closest - frame) / (closest - annotated_now))


def _scribble_planes(R, config, scribbles_list, annotated_now, final_masks, n_objects, n_interaction):
    """utils_atnet.py:31-52: n_obj x 3 x H x W (previous-round mask, positive, negative scribble images)."""
    planes = []
    for obj_id in range(1, n_objects + 1):
        if n_interaction == 1:
            pos = R.utils.scribble_to_image(scribbles_list, annotated_now, obj_id, dilation=config.scribble_dilation_param,
                                            prev_mask=final_masks[annotated_now])
            planes.append(np.stack([np.ones_like(pos) / 2, pos, np.zeros_like(pos)], axis=0))
        else:
            prev_round = (final_masks[annotated_now] == obj_id).astype(np.float32)
            pos, neg = R.utils.scribble_to_image(scribbles_list, annotated_now, obj_id,
                                                 dilation=config.scribble_dilation_param,
                                                 prev_mask=final_masks[annotated_now], blur=True, singleimg=False,
                                                 seperate_pos_neg=True)
            planes.append(np.stack([prev_round, pos, neg], axis=0))
    return np.stack(planes, axis=0)


def run_VOS_singleiact(net, config, split, scribbles_data, annotated_frames, final_masks, num_frames, n_objects,
                       n_interaction, subseq, pad_info, anno_3chEnc_r5_list, anno_6chEnc_r5_list, prob_map_of_frames,
                       hpad1, hpad2, wpad1, wpad2):
    R = _ref
    if R is None:
        raise RuntimeError("ivosw drop-in utils_atnet: start the entry script through `python -m ivosw.run` "
                           "(ivosw.hook binds the reference's utils.utils_atnet here)")
    if not prob_map_of_frames.is_cuda:
        raise RuntimeError("ivosw_b200 ATNet glue runs on CUDA only (no CPU fallback)")
    engine = get_engine(prob_map_of_frames.device)
    device = prob_map_of_frames.device
    annotated_frames_np = np.array(annotated_frames)
    annotated_now = annotated_frames[-1]
    output_masks = final_masks.copy().astype(np.float64)

    prop_list = R.utils.get_prop_list(annotated_frames, annotated_now, num_frames,
                                      proportion=config.test_propagation_proportion)
    prop_fore, prop_rear = sorted(prop_list)[0], sorted(prop_list)[-1]
    planes = torch.from_numpy(_scribble_planes(R, config, scribbles_data['scribbles'], annotated_now, final_masks,
                                               n_objects, n_interaction)).to(device)
    if (prop_list[0] != annotated_now) and (prop_list.count(annotated_now) != 2):
        raise NotImplementedError               # :55-57 (the list must run backward first, then forward)

    tfm = R.transforms.Compose([R.tr.Normalize_ApplymeanvarImage(config.mean, config.var), R.tr.ToTensor()])
    db_test = R.DAVIS2017(split=split, subseq=subseq, transform=tfm, root=config.davis_dataset_dir,
                          custom_frames=prop_list, seq_name=scribbles_data['sequence'], rgb=True, obj_id=None, no_gt=True,
                          retname=True, prev_round_masks=final_masks)
    loader = R.DataLoader(db_test, batch_size=1, shuffle=False, num_workers=4, pin_memory=True)

    (top, bottom), (left, right) = pad_info[0], pad_info[1]
    visits = 0                                   # times the annotated frame came up: 1 = backward pass, 2 = forward
    adjacent_to_anno = False
    prob_anno = prob_prop = r2_prev = r2_from_anno = None
    for batched in loader:
        frame = int(batched['meta']['frame_id'][0])
        image = batched['image'].to(device).expand(n_objects, -1, -1, -1)
        if frame == annotated_now:
            if visits >= 2:
                raise NotImplementedError
            visits += 1
            adjacent_to_anno = True
            if visits == 2:
                continue
            planes = engine.reflect_pad(planes, left, right, top, bottom)                     # :95-96
            logit, r5_6ch = net.forward_ANet(torch.cat([image, planes], dim=1))              # :97-99
            prob_anno = engine.sigmoid_blend(logit)                                          # :101
            prob_map_of_frames[frame] = prob_anno[:, 0].detach()                             # :102, 150
            r5_3ch, _, _, r2_from_anno = net.encoder_3ch.forward(image)                      # :103-104
            anno_6chEnc_r5_list.append(r5_6ch)
            anno_3chEnc_r5_list.append(r5_3ch)
            if len(anno_6chEnc_r5_list) != len(annotated_frames):
                raise NotImplementedError
            continue
        if adjacent_to_anno:
            r2_prev, predmask_prev = r2_from_anno, prob_anno
        else:
            predmask_prev = prob_prop
        adjacent_to_anno = False
        logit, r2_prev = net.forward_TNet(anno_3chEnc_r5_list, image, anno_6chEnc_r5_list, r2_prev, predmask_prev)
        alpha = _blend_alpha(annotated_frames_np, annotated_now, frame, backward=(visits == 1))
        # n_obj x 1 x P_H x P_W logits against the n_obj x P_H x P_W slice of the probability map: same element order
        prob_prop = engine.sigmoid_blend(logit, prev_inplace=prob_map_of_frames[frame], alpha=alpha)   # :124-150

    merged = R.utils_torch.combine_masks_with_batch(prob_map_of_frames[prop_fore:prop_rear + 1], n_obj=n_objects,
                                                    th=config.test_propth)
    H_all, W_all = prob_map_of_frames.shape[-2:]
    ys, xs = range(H_all)[hpad1:-hpad2], range(W_all)[wpad1:-wpad2]        # Python slice semantics, as the reference
    output_masks[prop_fore:prop_rear + 1] = merged[:, 0, hpad1:-hpad2, wpad1:-wpad2].cpu().numpy().astype(float)   # :152-155
    torch.cuda.empty_cache()
    if len(ys) == 0 or len(xs) == 0:            # a zero pad makes the reference's slice empty; keep that behaviour
        return output_masks, prob_map_of_frames.new_zeros((prob_map_of_frames.shape[0], n_objects + 1, len(ys), len(xs)))
    all_P = engine.atnet_assemble(prob_map_of_frames, ys[0], xs[0], len(ys), len(xs))        # :157-159
    return output_masks, all_P
