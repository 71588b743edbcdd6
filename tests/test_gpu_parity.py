"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, against the CPU oracle and the
committed golden fixtures (generated from the reference's own modules).

Tolerances (stated by north_star: argmax bit-exact, fp within tolerance):
  boxes           bit-exact (integer / float64 arithmetic)
  ROI crop        |d| <= 2e-6          (same rounding sequence as ATen's CPU kernels)
  r2..r5 probes   |d| <= 1e-4 + 1e-4*|ref|
  quality scores  |d| <= 1e-4
  Q-values        |d| <= 1e-5
  next_frame      identical (unless the fp64 arbiter's top-2 gap is below 1e-6, reported)
"""
import os

import numpy as np
import pytest
import torch

from ivosw import synth

pytestmark = pytest.mark.gpu

CONV_MODE = os.environ.get("IVOSW_CONV_MODE", "tc_fp16x3")


@pytest.fixture(scope="module")
def eng():
    from ivosw.engine import Engine
    e = Engine(0, CONV_MODE)
    e.load_assess(synth.assess_state_dict(0))
    e.load_brain(synth.brain_state_dict(0))
    yield e
    e.close()


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def test_brain_vs_golden_and_oracle(eng, golden_dir):
    from oracle import brain_ref
    g = _g(golden_dir, "brain")
    for seed in (0, 1):
        sd = synth.brain_state_dict(seed)
        eng.load_brain(sd)
        for (N, T) in ((1, 1), (1, 8), (1, 64), (1, 128), (4, 25)):
            x = g["s%d_N%d_T%d_x" % (seed, N, T)]
            q, am = eng.brain_forward(torch.from_numpy(x).float().cuda(), want_argmax=True)
            q = q.cpu().numpy()
            ref32 = g["s%d_N%d_T%d_f32_q" % (seed, N, T)]
            ref64 = g["s%d_N%d_T%d_f64_q" % (seed, N, T)]
            np.testing.assert_allclose(q, ref32, rtol=0, atol=1e-5)
            np.testing.assert_allclose(q, ref64, rtol=0, atol=1e-5)
            np.testing.assert_array_equal(am.cpu().numpy(), ref32.argmax(1))
    # random states, longer clips, vs the oracle
    sd = synth.brain_state_dict(0)
    eng.load_brain(sd)
    sdn = {k: v.numpy() for k, v in sd.items()}
    rng = np.random.default_rng(3)
    flips = 0
    for trial in range(24):
        T = int(rng.integers(2, 200))
        x = np.stack([rng.uniform(0.2, 0.95, T), rng.integers(0, 3, T).astype(np.float64)], 1)[None]
        q, am = eng.brain_forward(torch.from_numpy(x).float().cuda(), want_argmax=True)
        ref = brain_ref.brain_forward(sdn, x, np.float32)
        np.testing.assert_allclose(q.cpu().numpy(), ref, rtol=0, atol=1e-5)
        flips += int(am.item() != ref[0].argmax())
    assert flips == 0


def test_boxes_bit_exact(eng, golden_dir):
    g = _g(golden_dir, "boxes")
    masks = np.unpackbits(g["masks"], axis=-1)[..., :854].astype(np.float32)
    B, H, W = masks.shape
    tf = torch.zeros((B, 3, H, W), device="cuda")
    tp = torch.from_numpy(masks * 0.8 + 0.1).cuda()           # probabilities 0.9 / 0.1
    _, boxes = eng.assess_forward(tf, tp, want_boxes=True)
    np.testing.assert_array_equal(boxes.cpu().numpy(), g["boxes"])
    small = g["small"]
    _, boxes = eng.assess_forward(torch.zeros((3, 3, 37, 53), device="cuda"), torch.from_numpy(small * 0.9).cuda(),
                                  want_boxes=True)
    np.testing.assert_array_equal(boxes.cpu().numpy(), g["boxes_small"])
    # threshold is strict: exactly 0.5 is background (tp > 0.5)
    tp = torch.full((1, 200, 300), 0.5, device="cuda")
    _, b = eng.assess_forward(torch.zeros((1, 3, 200, 300), device="cuda"), tp, want_boxes=True)
    from oracle import assess_ref
    np.testing.assert_array_equal(b.cpu().numpy(), assess_ref.all2yxhw(np.zeros((1, 200, 300), np.float32), 1.5))


@pytest.mark.parametrize("name", ["round_c1", "round_atnet_small", "round_single", "round_t16"])
def test_round_vs_golden(eng, golden_dir, name):
    g = _g(golden_dir, name)
    clip_id, T, H, W, O, seed = [int(v) for v in g["meta"]]
    all_F, all_P, annotated = synth.make_clip(clip_id, T, H, W, O, str(g["style"]))
    eng.load_assess(synth.assess_state_dict(seed))
    eng.load_brain(synth.brain_state_dict(seed))
    F_d, P_d = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    eng.enable_probes(True)
    for o in range(O):
        s, boxes = eng.assess_forward(F_d, P_d[:, o + 1], want_boxes=True)
        np.testing.assert_array_equal(boxes.cpu().numpy(), g["f32_boxes"][o])
        crop = eng.probe(0).cpu().numpy()                        # T x 4 x 256 x 256, RGB normalised
        mean = np.array([0.485, 0.456, 0.406], np.float32)[None, :, None, None]
        std = np.array([0.229, 0.224, 0.225], np.float32)[None, :, None, None]
        np.testing.assert_allclose(crop[:, :3, ::8, ::8] * std + mean, g["roi_f"][o], atol=2e-6)
        np.testing.assert_allclose(crop[:, 3, ::8, ::8], g["roi_p"][o], atol=2e-6)
        for which, (k, step) in enumerate((("pool", 8), ("r2", 8), ("r3", 4), ("r4", 2), ("r5", 1)), start=1):
            t = eng.probe(which).cpu().numpy()
            probe = t[:, ::max(1, t.shape[1] // 8), ::step, ::step]
            np.testing.assert_allclose(probe, g[k][o], atol=1e-4, rtol=1e-4, err_msg=k)
        np.testing.assert_allclose(s.cpu().numpy(), g["f32_scores"][:, o], atol=1e-4)
    eng.enable_probes(False)
    r = eng.round_device(F_d, P_d, synth.annotated_counts(annotated, T), want_scores=True)
    np.testing.assert_allclose(r["scores"], g["f32_scores"], atol=1e-4)
    np.testing.assert_allclose(r["mask_quality"], g["f32_mask_quality"], atol=1e-4)
    np.testing.assert_allclose(r["q"], g["f32_q"], atol=1e-5)
    q64 = np.sort(g["f64_q"])[::-1]
    if T == 1 or q64[0] - q64[1] > 1e-6:
        assert r["next_frame"] == int(g["f32_next_frame"])
    # host-buffer entry point gives the same answer; only the frame rows an ROI can touch are uploaded, so the
    # staging buffer is poisoned with NaNs first: reading a row that was not sent would show
    os.environ["IVOSW_E2E_POISON"] = "1"
    try:
        rh = eng.round_host(torch.from_numpy(all_F), torch.from_numpy(all_P), synth.annotated_counts(annotated, T),
                            want_scores=True)
    finally:
        del os.environ["IVOSW_E2E_POISON"]
    full_bytes = T * (3 + O) * H * W * 4
    assert 0 < eng.last_h2d_bytes() <= full_bytes
    assert rh["next_frame"] == r["next_frame"]
    np.testing.assert_array_equal(rh["mask_quality"], r["mask_quality"])
    np.testing.assert_array_equal(rh["scores"], r["scores"])
    np.testing.assert_array_equal(rh["q"], r["q"])


def test_round_vs_oracle_480p(eng):
    """One clip at the headline resolution (T kept small so the CPU oracle finishes in seconds)."""
    from oracle import round_ref
    T, H, W, O = 6, 480, 854, 2
    all_F, all_P, annotated = synth.make_clip(11, T, H, W, O)
    assess_sd, brain_sd = synth.assess_state_dict(0), synth.brain_state_dict(0)
    eng.load_assess(assess_sd)
    eng.load_brain(brain_sd)
    r = eng.round_device(torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda(),
                         synth.annotated_counts(annotated, T), want_scores=True)
    ref = round_ref.recommend_frame_wild_ours(assess_sd, brain_sd, all_F, all_P, annotated)
    np.testing.assert_allclose(r["scores"], ref["scores"], atol=1e-4)
    np.testing.assert_allclose(r["q"], ref["q"], atol=1e-5)
    assert r["next_frame"] == ref["next_frame"]


def test_round_properties_full_size(eng):
    """BASELINE config C2 size (T=64, 480p, O=2): size-independent properties instead of the oracle.
    (1) sharding invariance: scoring frame ranges separately gives the same mask_quality bit for bit;
    (2) chunking invariance; (3) determinism; (4) frame permutation permutes the scores."""
    T, H, W, O = 64, 480, 854, 2
    all_F, all_P, annotated = synth.make_clip(0, T, H, W, O)
    eng.load_assess(synth.assess_state_dict(0))
    eng.load_brain(synth.brain_state_dict(0))
    F_d, P_d = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    ann = synth.annotated_counts(annotated, T)
    full = eng.round_device(F_d, P_d, ann, want_scores=True)
    again = eng.round_device(F_d, P_d, ann, want_scores=True)
    np.testing.assert_array_equal(full["scores"], again["scores"])
    assert full["next_frame"] == again["next_frame"]
    parts = [eng.round_device(F_d, P_d, ann, t_begin=a, t_end=b)["mask_quality"] for a, b in ((0, 8), (8, 40), (40, 64))]
    np.testing.assert_array_equal(np.concatenate(parts), full["mask_quality"])
    nf, q = eng.agent_action(full["mask_quality"], ann)
    assert nf == full["next_frame"]
    np.testing.assert_array_equal(q, full["q"])
    # (5) host-buffer path at full size: chunked upload, only the ROI row bands of every frame are sent (poisoned
    #     staging buffer: an unsent row that is read would turn the scores into NaN)
    os.environ["IVOSW_E2E_POISON"] = "1"
    try:
        rh = eng.round_host(torch.from_numpy(all_F), torch.from_numpy(all_P), ann, want_scores=True)
    finally:
        del os.environ["IVOSW_E2E_POISON"]
    np.testing.assert_array_equal(rh["scores"], full["scores"])
    assert rh["next_frame"] == full["next_frame"]
    assert eng.last_h2d_bytes() < T * (3 + O) * H * W * 4
    perm = np.random.default_rng(0).permutation(T)
    pr = eng.round_device(F_d[perm].contiguous(), P_d[perm].contiguous(), ann, want_scores=True)
    np.testing.assert_array_equal(pr["scores"], full["scores"][perm])
    assert np.isfinite(full["q"]).all() and 0 <= full["next_frame"] < T


def test_manet_tail(eng, golden_dir):
    g = _g(golden_dir, "manet_tail")
    H, W = [int(v) for v in g["hw"]]
    logits = torch.from_numpy(g["logits"]).cuda()
    masks, all_P = eng.manet_tail(logits, H, W)
    np.testing.assert_allclose(all_P.cpu().numpy(), g["all_P"], atol=2e-6)
    # masks: identical except where the top-2 upsampled logits tie to fp32 rounding
    up = torch.nn.functional.interpolate(torch.from_numpy(g["logits"]).double(), size=(H, W), mode="bilinear",
                                         align_corners=True)
    top2 = up.topk(2, dim=1).values
    near_tie = ((top2[:, 0] - top2[:, 1]) < 1e-6).numpy()
    diff = masks.cpu().numpy().astype(np.uint8) != g["masks"]
    assert not (diff & ~near_tie).any()
    assert diff.sum() <= near_tie.sum()


def test_multi_chunk_invariance():
    """T*O larger than the per-pass chunk (IVOSW_CHUNK): odd chunk sizes, ragged last chunk — same bits."""
    from ivosw.engine import Engine
    T, H, W, O = 7, 200, 300, 2
    all_F, all_P, annotated = synth.make_clip(31, T, H, W, O)
    ann = synth.annotated_counts(annotated, T)
    F_d, P_d = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    res = []
    for chunk in ("128", "5", "1"):
        os.environ["IVOSW_CHUNK"] = chunk
        e = Engine(0, CONV_MODE)
        e.load_assess(synth.assess_state_dict(0)); e.load_brain(synth.brain_state_dict(0))
        res.append(e.round_device(F_d, P_d, ann, want_scores=True))
        res.append(e.round_device(F_d, P_d, ann, want_scores=True))      # graph replay path
        e.close()
    os.environ.pop("IVOSW_CHUNK")
    for r in res[1:]:
        np.testing.assert_array_equal(r["scores"], res[0]["scores"])
        assert r["next_frame"] == res[0]["next_frame"]


def test_odd_frame_size_vs_oracle(eng):
    """Odd H, W (H*W not a multiple of 4): the scalar bbox path, unaligned plane starts and odd row pitches in the
    strided host uploads.  Device-resident and host-buffer rounds against the CPU oracle."""
    from oracle import round_ref
    T, H, W, O = 3, 201, 303, 2
    all_F, all_P, annotated = synth.make_clip(17, T, H, W, O)
    assess_sd, brain_sd = synth.assess_state_dict(0), synth.brain_state_dict(0)
    eng.load_assess(assess_sd)
    eng.load_brain(brain_sd)
    ref = round_ref.recommend_frame_wild_ours(assess_sd, brain_sd, all_F, all_P, annotated)
    ann = synth.annotated_counts(annotated, T)
    r = eng.round_device(torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda(), ann, want_scores=True)
    np.testing.assert_allclose(r["scores"], ref["scores"], atol=1e-4)
    np.testing.assert_allclose(r["q"], ref["q"], atol=1e-5)
    assert r["next_frame"] == ref["next_frame"]
    os.environ["IVOSW_E2E_POISON"] = "1"
    try:
        rh = eng.round_host(torch.from_numpy(all_F), torch.from_numpy(all_P), ann, want_scores=True)
    finally:
        del os.environ["IVOSW_E2E_POISON"]
    np.testing.assert_array_equal(rh["scores"], r["scores"])
    assert rh["next_frame"] == r["next_frame"]


def test_out_of_memory_surfaces_as_runtime_error(eng):
    """eval_agent_manet.py:391-396 retries on RuntimeError containing 'out of memory'."""
    B = 30000                                    # ~22 MB of workspace per unit -> far beyond 180 GB
    tf = torch.zeros((B, 3, 8, 8), device="cuda")
    tp = torch.zeros((B, 8, 8), device="cuda")
    os.environ["IVOSW_CHUNK"] = "30000"
    from ivosw.engine import Engine
    e = Engine(0, CONV_MODE)
    e.load_assess(synth.assess_state_dict(0))
    with pytest.raises(RuntimeError, match="out of memory"):
        e.assess_forward(tf, tp)
    os.environ.pop("IVOSW_CHUNK")
    # the context stays usable afterwards
    s = e.assess_forward(tf[:2], tp[:2])
    assert torch.isfinite(s).all()
    e.close()


def test_stack_kernel_matches_per_layer_kernels():
    """conv_stack.cu (IVOSW_STACK=1: all 52 layers in one persistent launch, tile-level dependencies) against the default
    conv_tc.cu (one launch per layer): the tile arithmetic is the same, so scores must agree bit for bit — for several group schedules
    (units per group in res2 / res3 / res4 / res5, including groups that leave a ragged tail), an odd unit count
    (8x8 tiles that span two images, the second one missing), and on graph replay."""
    from ivosw.engine import Engine
    T, H, W, O = 7, 160, 288, 3                      # 21 units
    all_F, all_P, annotated = synth.make_clip(33, T, H, W, O)
    ann = synth.annotated_counts(annotated, T)
    F_d, P_d = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    keys = ("IVOSW_STACK", "IVOSW_STACK_G2", "IVOSW_STACK_G3", "IVOSW_STACK_G4", "IVOSW_STACK_G5", "IVOSW_FUSE_DS")
    saved = {k: os.environ.get(k) for k in keys}
    res = []
    try:
        # (the per-layer reference runs with IVOSW_FUSE_DS=0: the default path evaluates conv3 + downsample of a stage's first
        #  block as one GEMM with folded BatchNorm scales — same mathematics, different rounding; checked at the end)
        for cfg in ({"IVOSW_FUSE_DS": "0"}, {"IVOSW_STACK": "1"},
                    {"IVOSW_STACK": "1", "IVOSW_STACK_G2": "2", "IVOSW_STACK_G3": "4", "IVOSW_STACK_G4": "6", "IVOSW_STACK_G5": "10"},
                    {"IVOSW_STACK": "1", "IVOSW_STACK_G2": "0", "IVOSW_STACK_G3": "0"},
                    {"IVOSW_STACK": "1", "IVOSW_STACK_G2": "5", "IVOSW_STACK_G3": "3"}):
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(cfg)
            e = Engine(0, CONV_MODE)
            e.load_assess(synth.assess_state_dict(0)); e.load_brain(synth.brain_state_dict(0))
            for _ in range(3):                        # eager, capture, replay
                res.append(e.round_device(F_d, P_d, ann, want_scores=True))
            res.append(e.round_device(F_d[:5], P_d[:5], ann[:5], want_scores=True, want_action=False))   # 15 units (odd)
            assert e.saturation_count() == 0
            e.close()
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    for i, r in enumerate(res):
        ref = res[3] if r["scores"].shape[0] == 5 else res[0]
        np.testing.assert_array_equal(r["scores"], ref["scores"], err_msg="run %d" % i)
        if r["next_frame"] is not None:
            assert r["next_frame"] == res[0]["next_frame"]
    # default path (fused conv3 + downsample tails): equal to the unfused arithmetic far inside the parity tolerance
    e = Engine(0, CONV_MODE)
    e.load_assess(synth.assess_state_dict(0)); e.load_brain(synth.brain_state_dict(0))
    fused = e.round_device(F_d, P_d, ann, want_scores=True)
    e.close()
    assert float(np.abs(fused["scores"] - res[0]["scores"]).max()) < 2e-5
    assert fused["next_frame"] == res[0]["next_frame"]


def test_wide_dynamic_range_and_saturation_counter():
    """Activations far from O(1): the split-fp16 planes carry x = hi + lo/2048 with fp16's exponent range.  Trunks scaled
    by 300x and by 1/300 (the stem's BatchNorm) must still agree with the fp32 oracle; a trunk scaled until it leaves the
    fp16 range must be REPORTED (ivosw_conv_saturation_count), not silently clamped."""
    from ivosw.engine import Engine
    from oracle import round_ref
    T, H, W, O = 4, 160, 256, 1
    all_F, all_P, annotated = synth.make_clip(51, T, H, W, O)
    ann = synth.annotated_counts(annotated, T)
    F_d, P_d = torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda()
    brain_sd = synth.brain_state_dict(0)
    e = Engine(0, CONV_MODE)
    e.load_brain(brain_sd)
    for gain in (300.0, 1.0 / 300.0):
        sd = {k: v.clone() for k, v in synth.assess_state_dict(0).items()}
        sd["Encoder.bn1.weight"] *= gain
        sd["Encoder.bn1.bias"] *= gain
        sd["fc1.weight"] /= max(gain, 1.0)                    # keep the scores O(1) for the large trunk
        e.load_assess(sd)
        r = e.round_device(F_d, P_d, ann, want_scores=True)
        ref = round_ref.recommend_frame_wild_ours(sd, brain_sd, all_F, all_P, annotated)
        scale = max(1.0, float(np.abs(ref["scores"]).max()))
        assert float(np.abs(r["scores"] - ref["scores"]).max()) <= 1e-4 * scale, (gain, r["scores"], ref["scores"])
        assert e.saturation_count() == 0
    sd = {k: v.clone() for k, v in synth.assess_state_dict(0).items()}
    sd["Encoder.bn1.weight"] *= 3e4
    sd["Encoder.bn1.bias"] *= 3e4
    e.load_assess(sd)
    e.round_device(F_d, P_d, ann, want_scores=True)
    assert e.saturation_count() > 0                           # loud: the caller can see that the range guard fired
    from ivosw._lib import lib
    assert b"fp16 range" in lib.ivosw_last_error()
    assert e.saturation_count() == 0                          # the read above reset the counter
    e.close()
