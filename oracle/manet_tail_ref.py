"""ORACLE (test infrastructure, not product code) — CPU restatement of the
fully-specified tail of the MANet round wrapper
``utils/utils_manet.py::get_results``: per frame bilinear upsample
(align_corners=True) of the (O+1)-channel logits to H x W (lines 76-77,
109-110, 146-147), per-pixel argmax -> mask (78-79, 113-114), and the final
channel softmax of the stacked upsampled logits (161).

Parity status: PINNED for this tail (golden produced by running the
reference's own get_results on CPU with a labelled stand-in IntVOS whose
logits are synthetic, tests/golden/make_golden.py).  The MANet network itself
(IntVOS / DeepLab) is absent from the reference tree: parity unpinned, not
restated here.
"""
import torch
import torch.nn.functional as F


def manet_tail(logits, H, W):
    """logits: T x (O+1) x h x w fp32.  Returns (final_masks T x H x W fp32,
    all_P T x (O+1) x H x W fp32)."""
    up = F.interpolate(torch.as_tensor(logits), size=(H, W), mode="bilinear", align_corners=True)
    masks = torch.argmax(up, dim=1).float()
    return masks, torch.softmax(up, 1)
