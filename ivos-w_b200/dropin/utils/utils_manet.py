"""Drop-in for ``utils/utils_manet.py::get_results`` (59-163) — the MANet round wrapper.

The MANet network itself (``model.int_seghead`` / ``model.prop_seghead``, external
lightas/CVPR2020_MANet, not part of the reference tree) is called exactly as the
reference calls it.  What changes is the wrapper's own arithmetic: per frame the
bilinear upsample + argmax (76-81, 109-117, 146-154) and, at the end, the
``cat`` + channel softmax over T x (O+1) x H x W (160-161) are one CUDA kernel
per frame that writes the mask, ``prev_label_storage`` and the frame's slice of
``all_P`` directly — no T separate logit tensors, no concatenation pass.
"""
import torch

from ivosw.engine import get_engine

try:                                    # the reference reads cfg.KNNS from MANet's config module
    from config import cfg
except Exception:                       # pragma: no cover - MANet checkout absent
    class cfg:                          # utils/config_manet/config.py:90
        KNNS = 1


def rough_ROI(ref_scribble_labels):
    """utils/utils_manet.py:22-39 on the device (bbox reduction + masked copy in one kernel)."""
    return get_engine(ref_scribble_labels.device).rough_roi(ref_scribble_labels, dist=20).to(ref_scribble_labels.dtype)


def _tail(engine, logits, h, w, masks, all_P, idx):
    engine.manet_tail(logits, h, w, masks_out=masks[idx:idx + 1], all_p_out=all_P[idx:idx + 1])
    return masks[idx:idx + 1]


def get_results(model, ref_frame_embedding, scribble_label, prev_label, eval_global_map_tmp_dic, local_map_dics,
                n_interaction, sequence, obj_nums, next_frame, first_scribble, h, w, prev_label_storage,
                total_frame_num, embedding_memory):
    device = ref_frame_embedding.device
    engine = get_engine(device)
    T = total_frame_num
    final_masks = None
    all_P = None

    def finish(tmp_dic, frame):
        nonlocal final_masks, all_P
        logits = tmp_dic[sequence]
        if final_masks is None:
            final_masks = torch.empty((T, h, w), device=device, dtype=torch.float32)
            all_P = torch.empty((T, logits.shape[1], h, w), device=device, dtype=torch.float32)
        m = _tail(engine, logits, h, w, final_masks, all_P, frame)
        pred_label = m.long()
        prev_label_storage[frame] = pred_label
        return pred_label

    tmp_dic, local_map_dics = model.int_seghead(ref_frame_embedding=ref_frame_embedding,
                                                ref_scribble_label=scribble_label,
                                                prev_round_label=prev_label,
                                                global_map_tmp_dic=eval_global_map_tmp_dic,
                                                local_map_dics=local_map_dics,
                                                interaction_num=n_interaction,
                                                seq_names=[sequence],
                                                gt_ids=torch.Tensor([obj_nums]),
                                                frame_num=[next_frame],
                                                first_inter=first_scribble)
    pred_label = finish(tmp_dic, next_frame)
    ref_prev_label = pred_label.unsqueeze(0)

    def propagate(frames):
        nonlocal eval_global_map_tmp_dic, local_map_dics
        prev_label = ref_prev_label
        prev_embedding = ref_frame_embedding
        for ii in frames:
            current_embedding = embedding_memory[ii].unsqueeze(0)
            tmp_dic, eval_global_map_tmp_dic, local_map_dics = model.prop_seghead(
                ref_frame_embedding, prev_embedding, current_embedding, scribble_label, prev_label,
                normalize_nearest_neighbor_distances=True, use_local_map=True, seq_names=[sequence],
                gt_ids=torch.Tensor([obj_nums]), k_nearest_neighbors=cfg.KNNS,
                global_map_tmp_dic=eval_global_map_tmp_dic, local_map_dics=local_map_dics,
                interaction_num=n_interaction, start_annotated_frame=next_frame, frame_num=[ii],
                dynamic_seghead=model.dynamic_seghead)
            prev_label = finish(tmp_dic, ii).unsqueeze(0)
            prev_embedding = current_embedding

    propagate(range(next_frame + 1, T))           # propagation ->   (87-117)
    propagate(range(next_frame - 1, -1, -1))      # propagation <-   (123-154)
    return final_masks, all_P
