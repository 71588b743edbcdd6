"""GPU (-m gpu): the AssessNet optimisation step (csrc/train.cu, BASELINE config C5) against the fixture produced by the
reference's own AssessNet in train mode + torch.optim.SGD (tests/golden/assess_train.npz: two consecutive iterations of
quality_assessment.py::train's loop body, no zero_grad between them).

Tolerances: predictions / loss 1e-4 (fp32 forward through 54 train-mode BatchNorms at batch 4); gradients 3e-2 of a
tensor's largest entry — the band tests/test_oracle_golden.py measured for ANY fp32-grade re-implementation whose ROI crop
differs from ATen's by ~5e-6 (53 batch-statistics layers at batch 4 amplify it); parameters after the steps 1e-7 (the
update is lr = 5e-6 times a clamped gradient); BatchNorm running statistics 1e-4."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mg(golden_dir):
    sys.path.insert(0, golden_dir)
    try:
        import make_golden as mg
    finally:
        sys.path.pop(0)
    return mg


def test_train_step_vs_reference_golden(golden_dir):
    from ivosw.engine import Engine
    mg = _mg(golden_dir)
    g = np.load(os.path.join(golden_dir, "assess_train.npz"))
    e = Engine(0)
    sd0 = mg.train_state_dict()
    e.train_begin(sd0)
    for step in range(2):
        imgs, probs, targets, valid = mg.synth_train_batch(step)
        loss, pred = e.train_step(torch.from_numpy(imgs).cuda(), torch.from_numpy(probs).cuda(), targets, valid, **mg.TRAIN_HP)
        np.testing.assert_allclose(pred, g["pred%d" % step], rtol=1e-4, atol=1e-4)
        assert abs(loss - float(g["loss%d" % step])) <= 1e-4 * abs(float(g["loss%d" % step]))
        sd, grads = e.train_export(want_grads=True)
        for k in mg.TRAIN_KEYS:
            stride = 101 if sd[k].numel() > 10000 else 1
            ref = g["grad%d_%s" % (step, k)]
            np.testing.assert_allclose(grads[k].numpy().reshape(-1)[::stride], ref, rtol=0,
                                       atol=3e-2 * float(np.abs(ref).max()) + 1e-12, err_msg="grad %d %s" % (step, k))
            np.testing.assert_allclose(sd[k].numpy().reshape(-1)[::stride], g["param%d_%s" % (step, k)], rtol=0, atol=1e-7,
                                       err_msg="param %d %s" % (step, k))
    for k in mg.TRAIN_BUFFERS:
        if k.endswith("num_batches_tracked"):
            continue                                  # a host-side counter in the reference; the blob does not carry it
        np.testing.assert_allclose(sd[k].numpy().reshape(-1).astype(np.float64), g["buf_" + k].astype(np.float64), rtol=1e-4,
                                   atol=1e-5, err_msg=k)
    # the exported state loads straight back into the inference path
    e.load_assess(sd)
    imgs, probs, _, _ = mg.synth_train_batch(0)
    s = e.assess_forward(torch.from_numpy(imgs).cuda(), torch.from_numpy(probs).cuda())
    assert torch.isfinite(s).all()
    e.close()


def test_train_step_no_valid_sample_and_determinism(golden_dir):
    """`if counter == 0: continue` (:259): no backward, no update.  Two engines stepping on the same data agree bit for bit
    (fixed-order reductions, no atomics)."""
    from ivosw.engine import Engine
    mg = _mg(golden_dir)
    imgs, probs, targets, valid = mg.synth_train_batch(0)
    F, P = torch.from_numpy(imgs).cuda(), torch.from_numpy(probs).cuda()
    outs = []
    for _ in range(2):
        e = Engine(0)
        e.train_begin(mg.train_state_dict())
        loss0, _ = e.train_step(F, P, targets, np.zeros(len(valid), bool), **mg.TRAIN_HP)
        assert loss0 is None
        sd_a = e.train_export()
        np.testing.assert_array_equal(sd_a["fc1.weight"].numpy(), mg.train_state_dict()["fc1.weight"].numpy())
        loss, pred = e.train_step(F, P, targets, valid, **mg.TRAIN_HP)
        sd, grads = e.train_export(want_grads=True)
        outs.append((loss, pred, sd["Encoder.res3.1.conv2.weight"].numpy(), grads["Encoder.conv1.weight"].numpy()))
        e.close()
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1:], outs[1][1:]):
        np.testing.assert_array_equal(a, b)


def test_train_step_split_then_apply_equals_fused(golden_dir):
    """The data-parallel form on one rank: train_step(apply=False) + the aliased gradient tensor + train_apply gives the
    same parameters as the fused step (ivosw.dist.assess_train_step_data_parallel without a process group)."""
    from ivosw import dist as ivdist
    from ivosw.engine import Engine
    mg = _mg(golden_dir)
    imgs, probs, targets, valid = mg.synth_train_batch(0)
    F, P = torch.from_numpy(imgs).cuda(), torch.from_numpy(probs).cuda()
    a, b = Engine(0), Engine(0)
    a.train_begin(mg.train_state_dict()); b.train_begin(mg.train_state_dict())
    la, _ = a.train_step(F, P, targets, valid, **mg.TRAIN_HP)
    lb = ivdist.assess_train_step_data_parallel(b, F, P, targets, valid, **mg.TRAIN_HP)
    assert la == lb
    g = b.train_grads_tensor()
    assert g.is_cuda and g.numel() == 23566343 and float(g.abs().max()) > 0
    sa, sb = a.train_export(), b.train_export()
    for k in ("fc1.weight", "Encoder.conv1.weight", "Encoder.res4.3.conv3.weight", "Encoder.res2.0.bn1.bias"):
        np.testing.assert_array_equal(sa[k].numpy(), sb[k].numpy())
    a.close(); b.close()


def test_train_step_tensor_core_path_matches_cuda_core_path(golden_dir, monkeypatch):
    """Forward convolutions and data gradients run on the tcgen05 kernel (split-fp16, fp32-grade), or — IVOSW_TRAIN_TC=0 — on
    the fp32 CUDA-core kernel.  Same batch, both paths: predictions / loss within the fp32 band, gradients within the 3 % of a
    tensor's scale that any two fp32-grade paths differ by at batch 4 (module docstring).  With 3 samples the 8x8 layers (3 * 64 pixels: not a multiple of the 128-pixel GEMM tile) take the
    CUDA-core kernel inside the tensor-core run, so the mixed dispatch is exercised too."""
    from ivosw.engine import Engine
    mg = _mg(golden_dir)
    imgs, probs, targets, valid = mg.synth_train_batch(0)
    outs = {}
    for n in (4, 3):
        F, P = torch.from_numpy(imgs[:n]).cuda(), torch.from_numpy(probs[:n]).cuda()
        for tc in ("1", "0"):
            monkeypatch.setenv("IVOSW_TRAIN_TC", tc)
            e = Engine(0)
            e.train_begin(mg.train_state_dict())
            loss, pred = e.train_step(F, P, targets[:n], np.ones(n, bool), **mg.TRAIN_HP)
            _, grads = e.train_export(want_grads=True)
            outs[(n, tc)] = (loss, pred, {k: grads[k].numpy() for k in ("Encoder.conv1.weight", "Encoder.res2.0.conv1.weight",
                                                                         "Encoder.res3.1.conv2.weight", "Encoder.res5.2.conv3.weight",
                                                                         "Encoder.res4.0.downsample.0.weight", "fc1.weight")})
            e.close()
        (la, pa, ga), (lb, pb, gb) = outs[(n, "1")], outs[(n, "0")]
        assert abs(la - lb) <= 1e-4 * abs(lb), (n, la, lb)
        np.testing.assert_allclose(pa, pb, rtol=1e-4, atol=1e-4)
        for k in ga:
            scale = float(np.abs(gb[k]).max()) + 1e-12
            assert float(np.abs(ga[k] - gb[k]).max()) <= 3e-2 * scale, (n, k, float(np.abs(ga[k] - gb[k]).max()), scale)
