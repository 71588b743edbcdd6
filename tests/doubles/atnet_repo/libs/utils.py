"""TEST DOUBLE for ATNet's libs/utils.py (external repo yuk6heo/IVOS-ATNet, absent): synthetic, deterministic
behaviour with the call signatures utils/utils_atnet.py uses (:25-26, 34-36, 43-46)."""
import numpy as np

IS_TEST_DOUBLE = True


def get_prop_list(annotated_frames, annotated_now, num_frames, proportion=1.0):
    """backward from the annotated frame to 0, then the annotated frame again and forward to the end"""
    return list(range(annotated_now, -1, -1)) + list(range(annotated_now, num_frames))


def scribble_to_image(scribbles_list, frame, obj_id, dilation=5, prev_mask=None, blur=False, singleimg=True,
                      seperate_pos_neg=False):
    """scribbles_list[frame] is a list of dicts {'object_id', 'path': [[x, y] in 0..1]}: draws dilated dots"""
    h, w = prev_mask.shape
    pos, neg = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    for stroke in scribbles_list[frame]:
        img = pos if stroke['object_id'] == obj_id else neg
        for x, y in stroke['path']:
            cy, cx = int(round(y * (h - 1))), int(round(x * (w - 1)))
            img[max(0, cy - dilation):cy + dilation + 1, max(0, cx - dilation):cx + dilation + 1] = 1.0
    if blur:        # deterministic 3-tap smoothing in both directions
        for img in (pos, neg):
            img[:] = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0) + np.roll(img, 1, 1) + np.roll(img, -1, 1)) / 5.0
    if seperate_pos_neg:
        return pos, neg
    return pos


def apply_pad(img, padinfo=None):
    """pads H, W up to the next multiple of 32, at least 2 px on each side (so that x[h1:-h2] is never empty)"""
    h, w = img.shape[:2]

    def split(n):
        total = (-n) % 32
        if total < 4:
            total += 32
        return (total // 2, total - total // 2)
    pad = (split(h), split(w))
    out = np.pad(img, (pad[0], pad[1]) + ((0, 0),) * (img.ndim - 2), mode='reflect')
    return out, pad
