"""TEST DOUBLE for ATNet's libs/utils_torch.py (utils/utils_atnet.py:153)."""
import torch

IS_TEST_DOUBLE = True


def combine_masks_with_batch(prob, n_obj, th=0.5):
    """prob: f x n_obj x H x W -> f x 1 x H x W labels: argmax object where its probability exceeds th, else 0"""
    best, idx = prob.max(1, keepdim=True)
    return torch.where(best > th, (idx + 1).to(prob.dtype), torch.zeros_like(best))
