"""Generates tests/golden/atnet_round.npz by running the REFERENCE's own utils/utils_atnet.py::run_VOS_singleiact
(imported, unmodified, from /root/reference) for three consecutive interaction rounds on CPU.

ATNet itself (yuk6heo/IVOS-ATNet) is not in the reference tree, so the harness supplies LABELLED stand-ins
(tests/doubles/atnet_repo: libs.utils / libs.utils_torch / libs.custom_transforms, networks.atnet.ATnet) and replaces
the DAVIS2017 loader by a synthetic one (tests/doubles/checkout/datasets).  Harness-side shims, none of which touches the
reference file: ``np.float = float`` (removed from NumPy >= 1.24, used at utils_atnet.py:155, SURVEY A.Q14),
``Tensor.cuda`` / ``Module.cuda`` identity (the wrapper hard-codes .cuda()), ``torch.cuda.empty_cache`` no-op.

    python tests/golden/make_golden_atnet.py        (needs /root/reference; run in the build container only)
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
DOUBLES = os.path.join(REPO, "tests", "doubles")

T, N_OBJ = 9, 2
ROUNDS = ([4], [4, 1], [4, 1, 7])          # annotated_frames after each interaction


def scribbles_for(frame, rnd):
    rng = np.random.default_rng(100 * rnd + frame)
    strokes = []
    for obj in range(1, N_OBJ + 1):
        strokes.append({'object_id': obj, 'path': rng.uniform(0.15, 0.85, size=(4, 2)).tolist()})
    lst = [[] for _ in range(T)]
    lst[frame] = strokes
    return {'scribbles': lst, 'sequence': 'synthetic'}


def config():
    return SimpleNamespace(test_propagation_proportion=1.0, scribble_dilation_param=3, mean=0.45, var=0.22,
                           davis_dataset_dir='', test_propth=0.5)


def main():
    np.float = float                                                    # shim (A.Q14)
    torch.Tensor.cuda = lambda self, *a, **k: self                      # shim: CPU run of a .cuda()-hard-coded wrapper
    torch.cuda.empty_cache = lambda: None
    sys.path.insert(0, os.path.join(DOUBLES, "atnet_repo"))             # what sys.path.append('VOS/ATNet') provides
    sys.path.insert(0, REF)
    import utils.utils_atnet as ref                                     # the reference's own module
    assert ref.__file__.startswith(REF)
    sys.path.insert(0, os.path.join(DOUBLES, "checkout"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("double_davis", os.path.join(DOUBLES, "checkout", "datasets", "davis_dataset.py"))
    dd = importlib.util.module_from_spec(spec); spec.loader.exec_module(dd)
    ref.DAVIS2017 = dd.DAVIS2017                                        # shim: synthetic loader instead of JPEG reading
    _DL = ref.DataLoader
    ref.DataLoader = lambda ds, **kw: _DL(ds, batch_size=1, shuffle=False, num_workers=0)
    from libs import utils as lutils
    from networks.atnet import ATnet

    H, W = dd.FRAME_HW
    final_masks = np.zeros((T, H, W))
    pad_info = lutils.apply_pad(final_masks[0])[1]
    assert pad_info == dd.PAD
    (h1, h2), (w1, w2) = pad_info
    prob_map = torch.zeros((T, N_OBJ, H + h1 + h2, W + w1 + w2))
    net = ATnet()
    r5_3, r5_6 = [], []
    out = {"meta": np.array([T, N_OBJ, H, W, h1, h2, w1, w2])}
    for rnd, annotated in enumerate(ROUNDS, start=1):
        sd = scribbles_for(annotated[-1], rnd)
        masks, all_P = ref.run_VOS_singleiact(net, config(), 'val', sd, list(annotated), final_masks, T, N_OBJ, rnd, None,
                                              pad_info, r5_3, r5_6, prob_map, h1, h2, w1, w2)
        final_masks = masks
        out["r%d_masks" % rnd] = masks.astype(np.float32)
        out["r%d_all_P" % rnd] = all_P.numpy().copy()
        out["r%d_prob_map" % rnd] = prob_map.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "atnet_round.npz"), **out)
    print("wrote atnet_round.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
