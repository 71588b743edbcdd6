"""CPU: pins the oracle restatements (oracle/*.py) against fixtures generated
from the reference's own modules (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from ivosw import synth
from oracle import assess_ref, brain_ref, manet_tail_ref, round_ref

torch.set_num_threads(max(1, os.cpu_count() or 1))


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_boxes_known_answers(golden_dir):
    g = _load(golden_dir, "boxes")
    masks = np.unpackbits(g["masks"], axis=-1)[..., :854].astype(np.float32)
    got = assess_ref.all2yxhw(masks, scale=1.5)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, g["boxes"])          # bit-exact: integer/float64 arithmetic
    # SURVEY §8(a) a3 verified cases
    np.testing.assert_array_equal(got[0], [199.5, 349.5, 300, 450])
    np.testing.assert_array_equal(got[1], [204.5, 402, 192, 193.5])
    np.testing.assert_array_equal(got[2], [240, 427, 491, 865])
    np.testing.assert_array_equal(got[3], [240, 427, 491, 865])
    np.testing.assert_array_equal(assess_ref.all2yxhw(g["small"], 1.5), g["boxes_small"])


def test_brain_matches_reference(golden_dir):
    g = _load(golden_dir, "brain")
    for seed in (0, 1):
        sd = {k: v.numpy() for k, v in synth.brain_state_dict(seed).items()}
        for (N, T) in ((1, 1), (1, 8), (1, 64), (1, 128), (4, 25)):
            x = g["s%d_N%d_T%d_x" % (seed, N, T)]
            q64 = brain_ref.brain_forward(sd, x, np.float64)
            np.testing.assert_allclose(q64, g["s%d_N%d_T%d_f64_q" % (seed, N, T)], rtol=0, atol=1e-12)
            q32 = brain_ref.brain_forward(sd, x, np.float32)
            ref32 = g["s%d_N%d_T%d_f32_q" % (seed, N, T)]
            np.testing.assert_allclose(q32, ref32, rtol=0, atol=2e-6)
            assert (q32.argmax(1) == ref32.argmax(1)).all()


def test_grid_sample_restatement_matches_torch():
    g = torch.Generator().manual_seed(3)
    img = torch.rand((3, 4, 37, 53), generator=g)
    roi = torch.tensor([[18., 26., 50., 70.], [5., 5., 20., 20.], [30., 50., 30., 30.]])
    gx, gy = assess_ref.roi_grid(roi, (37, 53), 64)
    mine = assess_ref.grid_sample_bilinear(img, gx, gy)
    theta = torch.zeros(3, 2, 3)
    t00, t02, t11, t12 = assess_ref.roi_theta(roi, (37, 53))
    theta[:, 0, 0], theta[:, 0, 2], theta[:, 1, 1], theta[:, 1, 2] = t00, t02, t11, t12
    grid = torch.nn.functional.affine_grid(theta, (3, 1, 64, 64), align_corners=True)
    np.testing.assert_array_equal(torch.stack([gx, gy], -1).numpy(), grid.numpy())   # bit-exact grid
    ref = torch.nn.functional.grid_sample(img, grid, align_corners=True)
    np.testing.assert_allclose(mine.numpy(), ref.numpy(), atol=1e-6)


@pytest.mark.parametrize("name", ["round_c1", "round_atnet_small", "round_single", "round_t16"])
def test_round_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    clip_id, T, H, W, O, seed = [int(v) for v in g["meta"]]
    all_F, all_P, annotated = synth.make_clip(clip_id, T, H, W, O, str(g["style"]))
    assert annotated == [int(v) for v in g["annotated"]]
    assess_sd, brain_sd = synth.assess_state_dict(seed), synth.brain_state_dict(seed)
    # stage probes
    for o in range(O):
        pr = {}
        s = assess_ref.assess_forward(assess_sd, torch.from_numpy(all_F), torch.from_numpy(all_P[:, o + 1]), probes=pr)
        np.testing.assert_array_equal(pr["boxes"].numpy(), g["f32_boxes"][o])
        np.testing.assert_allclose(pr["tf_roi"][:, ::1, ::8, ::8].numpy(), g["roi_f"][o], atol=5e-6)
        np.testing.assert_allclose(pr["tp_roi"][:, ::8, ::8].numpy(), g["roi_p"][o], atol=5e-6)
        for k, step in (("pool", 8), ("r2", 8), ("r3", 4), ("r4", 2), ("r5", 1)):
            t = pr[k]
            probe = t[:, ::max(1, t.shape[1] // 8), ::step, ::step].numpy()
            np.testing.assert_allclose(probe, g[k][o], atol=2e-5, rtol=1e-5)
        np.testing.assert_allclose(s.numpy(), g["f32_scores"][:, o], atol=2e-5)
    mq = np.zeros(T)
    r = round_ref.recommend_frame_wild_ours(assess_sd, brain_sd, all_F, all_P, annotated, mask_quality=mq)
    np.testing.assert_allclose(mq, g["f32_mask_quality"], atol=2e-5)
    np.testing.assert_allclose(r["q"], g["f32_q"], atol=2e-6)
    assert r["next_frame"] == int(g["f32_next_frame"]) == int(g["f64_next_frame"])


def test_round_fp64_arbiter(golden_dir):
    g = _load(golden_dir, "round_atnet_small")
    clip_id, T, H, W, O, seed = [int(v) for v in g["meta"]]
    all_F, all_P, annotated = synth.make_clip(clip_id, T, H, W, O, str(g["style"]))
    r = round_ref.recommend_frame_wild_ours(synth.assess_state_dict(seed), synth.brain_state_dict(seed),
                                            all_F, all_P, annotated, dtype=torch.float64)
    np.testing.assert_allclose(r["mask_quality"], g["f64_mask_quality"], atol=1e-9)
    np.testing.assert_allclose(r["q"], g["f64_q"], atol=1e-10)


def test_manet_tail(golden_dir):
    g = _load(golden_dir, "manet_tail")
    H, W = [int(v) for v in g["hw"]]
    masks, all_P = manet_tail_ref.manet_tail(g["logits"], H, W)
    np.testing.assert_array_equal(masks.numpy().astype(np.uint8), g["masks"])
    np.testing.assert_allclose(all_P.numpy(), g["all_P"], atol=1e-7)


def _dqn_batch(seed, N=256, T=25):
    """Same generator as tests/golden/make_golden.py::synth_dqn_batch."""
    rng = np.random.default_rng(seed)

    def ann():
        a = np.zeros((N, T))
        for n in range(N):
            for i in rng.integers(0, T, size=rng.integers(1, 6)):
                a[n, i] += 1
        return a
    a0 = ann()
    a1 = a0.copy()
    act = rng.integers(0, T, size=N)
    a1[np.arange(N), act] += 1
    old_iou = rng.uniform(0.3, 0.95, (N, T)); new_iou = rng.uniform(0.3, 0.95, (N, T))
    rs = rng.choice([-1.0, 1.0], N); rd = rng.standard_normal(N)
    return (np.stack([old_iou, a0], 2), np.stack([new_iou, a1], 2), act, rs, rd)


def test_dqn_step_matches_reference(golden_dir):
    from oracle import dqn_ref
    g = _load(golden_dir, "dqn_step")
    st = dqn_ref.DqnState(synth.brain_state_dict(0), synth.brain_state_dict(1))
    for step in range(2):
        s, ns, act, rs, rd = _dqn_batch(100 + step)
        loss, grads = dqn_ref.update_agent(st, torch.from_numpy(s).float(), torch.from_numpy(ns).float(),
                                           torch.from_numpy(act), torch.from_numpy(rs).float(), torch.from_numpy(rd).float())
        assert abs(loss - float(g["f32_loss%d" % step])) < 1e-7
        for k, gr in grads.items():
            ref = g["grad%d_%s" % (step, k)]
            mine = gr.numpy().reshape(-1)
            mine = mine if step == 0 else mine[::8]
            np.testing.assert_allclose(mine, ref, atol=1e-7, rtol=1e-4, err_msg=k)
    for k, v in st.p.items():
        np.testing.assert_allclose(v.detach().numpy().reshape(-1)[::8], g["param_final_" + k], atol=2e-7, err_msg=k)


@pytest.mark.parametrize("sampler", ["torch", "restated"])
def test_assess_train_step(golden_dir, sampler, monkeypatch):
    """SURVEY §8(f) rank 3 (config C5): two consecutive iterations of quality_assessment.py::train's loop body
    (train-mode BatchNorm, masked MSE, clamp, SGD with momentum and weight decay, gradients accumulating because the
    loop never zeroes them) — oracle restatement vs the reference's own AssessNet + torch.optim.SGD.

    sampler="torch": the ROI crop comes from F.grid_sample itself, so the step sees bit-identical inputs and must
    reproduce the reference's gradients to round-off — that pins the semantics.  sampler="restated": the crop comes
    from the oracle's own bilinear sampler (<= 5e-6 away); 53 layers of train-mode BatchNorm at batch 4 amplify that to
    ~1 % of a gradient tensor's scale, which is the band any fp32-grade re-implementation of this step will land in."""
    import sys
    sys.path.insert(0, os.path.join(golden_dir))
    import torch.nn.functional as F
    from oracle import assess_train_ref
    import make_golden as mg                     # only its seeded batch generator and constants (no reference import)
    if sampler == "torch":
        monkeypatch.setattr(assess_ref, "grid_sample_bilinear",
                            lambda img, gx, gy: F.grid_sample(img, torch.stack([gx, gy], -1), align_corners=True))
    grad_tol = 1e-5 if sampler == "torch" else 3e-2
    g = _load(golden_dir, "assess_train")
    st = assess_train_ref.TrainState(mg.train_state_dict())
    for step in range(2):
        imgs, probs, targets, valid = mg.synth_train_batch(step)
        loss, pred = assess_train_ref.train_step(st, imgs, probs, targets, valid, **mg.TRAIN_HP)
        np.testing.assert_allclose(pred.numpy().reshape(-1), g["pred%d" % step], rtol=1e-5, atol=1e-5)
        assert abs(float(loss) - float(g["loss%d" % step])) <= 1e-5 * abs(float(g["loss%d" % step]))
        for k in mg.TRAIN_KEYS:
            p = st.params[k]
            stride = 101 if p.numel() > 10000 else 1
            ref = g["grad%d_%s" % (step, k)]
            np.testing.assert_allclose(p.grad.numpy().reshape(-1)[::stride], ref, rtol=0,
                                       atol=grad_tol * float(np.abs(ref).max()) + 1e-12, err_msg="grad %d %s" % (step, k))
            np.testing.assert_allclose(p.detach().numpy().reshape(-1)[::stride], g["param%d_%s" % (step, k)],
                                       rtol=0, atol=1e-7, err_msg="param %d %s" % (step, k))
    for k in mg.TRAIN_BUFFERS:
        np.testing.assert_allclose(st.buffers[k].numpy().reshape(-1).astype(np.float64), g["buf_" + k].astype(np.float64),
                                   rtol=1e-4, atol=1e-5, err_msg=k)
    no_grad = sorted(k for k, v in st.params.items() if v.grad is None)
    assert no_grad == [str(k) for k in g["no_grad_params"]]


def test_atnet_wrapper_oracle_vs_reference_golden(golden_dir):
    """oracle/atnet_glue_ref.py (restatement of utils/utils_atnet.py:72-159) replays the three interaction rounds the
    reference's own run_VOS_singleiact produced tests/golden/atnet_round.npz from — same stand-in network and loader
    (tests/doubles) — and must reproduce prob_map_of_frames and all_P bit for bit on CPU."""
    import importlib.util
    import sys
    from oracle import atnet_glue_ref as og
    here = os.path.dirname(os.path.abspath(__file__))
    dbl = os.path.join(here, "doubles")

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    mk = load("mk_atnet", os.path.join(golden_dir, "make_golden_atnet.py"))
    lutils = load("dbl_lutils", os.path.join(dbl, "atnet_repo", "libs", "utils.py"))
    dd = load("dbl_davis", os.path.join(dbl, "checkout", "datasets", "davis_dataset.py"))
    atnet = load("dbl_atnet", os.path.join(dbl, "atnet_repo", "networks", "atnet.py"))
    g = np.load(os.path.join(golden_dir, "atnet_round.npz"))
    T, n_obj, H, W, h1, h2, w1, w2 = [int(v) for v in g["meta"]]
    cfg = mk.config()
    pad_info = ((h1, h2), (w1, w2))

    def frames_fn(idx):
        img = (dd.synthetic_frame(idx) - cfg.mean) / cfg.var                       # libs.custom_transforms doubles
        return torch.from_numpy(img.transpose(2, 0, 1).copy()).float()[None].expand(n_obj, -1, -1, -1)

    prob_map = torch.zeros((T, n_obj, H + h1 + h2, W + w1 + w2))
    final_masks = np.zeros((T, H, W))
    r5_3, r5_6 = [], []
    net = atnet.ATnet()
    for rnd, annotated in enumerate(mk.ROUNDS, start=1):
        now = annotated[-1]
        sl = mk.scribbles_for(now, rnd)['scribbles']
        planes = []
        for obj in range(1, n_obj + 1):                                            # utils_atnet.py:31-52
            if rnd == 1:
                pos = lutils.scribble_to_image(sl, now, obj, dilation=cfg.scribble_dilation_param, prev_mask=final_masks[now])
                planes.append(np.stack([np.ones_like(pos) / 2, pos, np.zeros_like(pos)], 0))
            else:
                pos, neg = lutils.scribble_to_image(sl, now, obj, dilation=cfg.scribble_dilation_param,
                                                    prev_mask=final_masks[now], blur=True, singleimg=False,
                                                    seperate_pos_neg=True)
                planes.append(np.stack([(final_masks[now] == obj).astype(np.float32), pos, neg], 0))
        prop_list = lutils.get_prop_list(list(annotated), now, T)
        og.run_round(net, frames_fn, np.stack(planes, 0), prop_list, list(annotated), prob_map, pad_info, r5_3, r5_6)
        np.testing.assert_array_equal(prob_map.numpy(), g["r%d_prob_map" % rnd])
        np.testing.assert_array_equal(og.assemble_all_p(prob_map, h1, h2, w1, w2).numpy(), g["r%d_all_P" % rnd])
        final_masks = g["r%d_masks" % rnd].astype(np.float64)
    # pieces: reflection pad against torch's module
    x = torch.randn(2, 3, 9, 11)
    np.testing.assert_array_equal(og.reflect_pad(x, ((3, 2), (4, 1))).numpy(), torch.nn.ReflectionPad2d((4, 1, 3, 2))(x).numpy())
