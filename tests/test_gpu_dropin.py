"""GPU (-m gpu): the module-substitution drop-ins (ivos-w_b200/dropin) exercised the way the reference's entry
scripts use them (eval_agent_manet.py:170-190, 386-389, 411-418), against the fixtures generated from the
reference's own modules."""
import os
import random
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from ivosw import synth

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dropin():
    """the drop-in modules under the reference's names, through ivosw.hook over the test-double checkout"""
    from tests import doubles
    d = doubles.load_dropin(with_manet=True, with_atnet=True)
    assert d.M.IS_CHECKOUT_ORIGINAL and d.M.__ivosw_patched__ == ("get_results", "rough_ROI")
    assert d.M.load_network.__module__ == "utils.utils_manet"          # the checkout's own helper survives
    yield d
    doubles.deactivate()


def _cfg():
    return SimpleNamespace(phase="eval", agent=SimpleNamespace(memory_size=10, gamma=0.95, eps_start=0.7, eps_end=0.25,
                           eps_decay=500, update_rate=0.05, lr=5e-6, weight_decay=5e-4), data=SimpleNamespace(subset="train"))


@pytest.mark.parametrize("name", ["round_c1", "round_t16", "round_single"])
def test_recommend_frame_dropin(dropin, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    clip_id, T, H, W, O, seed = [int(v) for v in g["meta"]]
    all_F, all_P, annotated = synth.make_clip(clip_id, T, H, W, O, str(g["style"]))
    device = torch.device("cuda:0")
    # as eval_agent_manet.py:183-190 / 170-176 do it: build on CPU, load_state_dict(strict=True), .to(device), .eval()
    assess_net = dropin.S.AssessNet()
    assess_net.load_state_dict(synth.assess_state_dict(seed), strict=True)
    assess_net = assess_net.to(device).eval()
    agent = dropin.A.Agent(device, _cfg())
    agent.policy_net.load_state_dict(synth.brain_state_dict(seed), strict=True)
    mask_quality = np.zeros(T)
    random.seed(0)
    steps0 = agent.steps_done
    all_F_cpu = torch.from_numpy(all_F)                       # the reference keeps all_F on the CPU
    all_P_dev = torch.from_numpy(all_P).to(device)            # ... and all_P on the GPU
    for rnd in range(2):                                      # second round hits the per-clip frame cache
        nxt = dropin.U.recommend_frame(SimpleNamespace(setting="wild", method="ours"), assess_net, agent, device,
                                       n_frame=T, n_objects=O, all_F=all_F_cpu, all_P=all_P_dev,
                                       new_masks_quality=np.zeros(T), prev_frames=[], annotated_frames_list=annotated,
                                       mask_quality=mask_quality, first_frame=0, max_nb_interactions=8)
        assert int(nxt) == int(g["f32_next_frame"])
        assert isinstance(nxt, (np.integer, int))
        np.testing.assert_allclose(mask_quality, g["f32_mask_quality"], atol=1e-4)
    assert agent.steps_done == steps0 + 2                     # Agent.action's side effects are kept
    random.seed(0); random.random(); random.random()
    expect_next_draw = random.random()
    random.seed(0)
    for _ in range(2):
        random.random()
    assert random.random() == expect_next_draw
    # AssessNet.forward keeps the reference's output shapes (B x 1; (1,) for B == 1)
    out = assess_net(all_F_cpu.to(device), all_P_dev[:, 1])
    assert tuple(out.shape) == ((T, 1) if T > 1 else (1,))
    np.testing.assert_allclose(out.reshape(-1).cpu().numpy(), g["f32_scores"][:, 0], atol=1e-4)
    # 'worst' branch: argmin of predicted quality, skipping previously chosen frames
    mq2 = np.zeros(T)
    nxt_w = dropin.U.recommend_frame(SimpleNamespace(setting="wild", method="worst"), assess_net, None, device,
                                     n_frame=T, n_objects=O, all_F=all_F_cpu, all_P=all_P_dev,
                                     new_masks_quality=np.zeros(T), prev_frames=[], annotated_frames_list=annotated,
                                     mask_quality=mq2, first_frame=0, max_nb_interactions=8)
    assert int(nxt_w) == int(np.argsort(g["f32_mask_quality"])[0]) or T == 1
    # Brain.forward drop-in
    state = torch.from_numpy(np.stack([g["f32_mask_quality"], synth.annotated_counts(annotated, T)], 1)[None]).float().to(device)
    np.testing.assert_allclose(agent.policy_net(state).cpu().numpy()[0], g["f32_q"], atol=1e-5)


def test_get_results_dropin(dropin, golden_dir):
    """utils.utils_manet.get_results drop-in with the same stand-in IntVOS the fixture was generated with."""
    g = np.load(os.path.join(golden_dir, "manet_tail.npz"))
    logits = torch.from_numpy(g["logits"]).cuda()
    T, C, h, w = logits.shape
    H, W = [int(v) for v in g["hw"]]

    class StandInIntVOS:      # LABELLED STAND-IN: returns synthetic logits, not MANet
        dynamic_seghead = None

        def int_seghead(self, **kw):
            return {kw["seq_names"][0]: logits[kw["frame_num"][0]][None]}, kw["local_map_dics"]

        def prop_seghead(self, *a, **kw):
            return ({kw["seq_names"][0]: logits[kw["frame_num"][0]][None]}, kw["global_map_tmp_dic"], kw["local_map_dics"])

    emb = torch.zeros(T, 4, h, w, device="cuda")
    storage = torch.zeros(T, H, W, device="cuda")
    masks, all_P = dropin.M.get_results(StandInIntVOS(), emb[2:3], torch.zeros(1, 1, h, w, device="cuda"), None, {},
                                        ({}, {}), 1, "seq", C - 1, 2, True, H, W, storage, T, emb)
    assert tuple(masks.shape) == (T, H, W) and tuple(all_P.shape) == (T, C, H, W)
    np.testing.assert_allclose(all_P.cpu().numpy(), g["all_P"], atol=2e-6)
    up = torch.nn.functional.interpolate(torch.from_numpy(g["logits"]).double(), size=(H, W), mode="bilinear", align_corners=True)
    top2 = up.topk(2, dim=1).values
    near_tie = ((top2[:, 0] - top2[:, 1]) < 1e-6).numpy()
    diff = masks.cpu().numpy().astype(np.uint8) != g["masks"]
    assert not (diff & ~near_tie).any()
    np.testing.assert_array_equal(storage.cpu().numpy().astype(np.uint8)[~near_tie], g["storage"][~near_tie])


def test_rough_roi_dropin(dropin):
    """utils.utils_manet.rough_ROI: same result as the reference's algorithm restated with torch ops."""
    g = torch.Generator().manual_seed(3)
    lab = -torch.ones((3, 1, 120, 214))
    lab[0, 0, 40:60, 100:130] = torch.randint(0, 3, (20, 30), generator=g).float()
    lab[1, 0, 0:5, 0:7] = 1.0
    lab[1, 0, 110:120, 200:214] = 2.0
    lab[2, 0, 119, 213] = 0.0

    def ref(x):    # utils_manet.py:22-39
        b, _, h, w = x.shape
        filt = torch.zeros_like(x)
        for i in range(b):
            nb = (x[i] != -1).squeeze(0).nonzero()
            (hmin, wmin), _ = torch.min(nb, 0)
            (hmax, wmax), _ = torch.max(nb, 0)
            filt[i, 0, max(hmin - 20, 0):min(hmax + 20, h - 1), max(wmin - 20, 0):min(wmax + 20, w - 1)] = 1
        return torch.where(filt.bool(), x, torch.zeros_like(x))
    out = dropin.M.rough_ROI(lab.cuda())
    assert torch.equal(out.cpu(), ref(lab))
    with pytest.raises(ValueError):
        dropin.M.rough_ROI(-torch.ones((1, 1, 8, 8)).cuda())


# ------------------------------------------------------------------------------------------------------------
# ATNet round wrapper (utils/utils_atnet.py::run_VOS_singleiact, config C3) — against the reference's own function
# ------------------------------------------------------------------------------------------------------------
def _atnet_round_inputs():
    """The three interaction rounds tests/golden/make_golden_atnet.py drove the reference with (same stand-ins)."""
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    try:
        import make_golden_atnet as mk
    finally:
        sys.path.pop(0)
    return mk


def test_run_vos_singleiact_dropin_vs_reference_golden(dropin, golden_dir):
    g = np.load(os.path.join(golden_dir, "atnet_round.npz"))
    T, n_obj, H, W, h1, h2, w1, w2 = [int(v) for v in g["meta"]]
    mk = _atnet_round_inputs()
    from networks.atnet import ATnet                         # LABELLED STAND-IN (tests/doubles/atnet_repo)
    pad_info = ((h1, h2), (w1, w2))
    prob_map = torch.zeros((T, n_obj, H + h1 + h2, W + w1 + w2), device="cuda")
    final_masks = np.zeros((T, H, W))
    r5_3, r5_6 = [], []
    # the double checkout's DataLoader import is torch's own: 4 workers + pin_memory, as the reference asks for
    for rnd, annotated in enumerate(mk.ROUNDS, start=1):
        masks, all_P = dropin.AT.run_VOS_singleiact(ATnet(), mk.config(), 'val', mk.scribbles_for(annotated[-1], rnd),
                                                    list(annotated), final_masks, T, n_obj, rnd, None, pad_info, r5_3, r5_6,
                                                    prob_map, h1, h2, w1, w2)
        final_masks = masks
        assert all_P.is_cuda and tuple(all_P.shape) == (T, n_obj + 1, H, W)
        assert float(all_P[:, 0].abs().max()) == 0.0
        np.testing.assert_allclose(prob_map.cpu().numpy(), g["r%d_prob_map" % rnd], atol=2e-6, err_msg="round %d" % rnd)
        np.testing.assert_allclose(all_P.cpu().numpy(), g["r%d_all_P" % rnd], atol=2e-6)
        # masks come from the (stand-in) external combine_masks_with_batch thresholding at 0.5: identical except where a
        # probability sits within 2e-6 of the threshold or of the other object's
        diff = masks.astype(np.float32) != g["r%d_masks" % rnd]
        pm = g["r%d_prob_map" % rnd][:, :, h1:-h2, w1:-w2]
        near = (np.abs(pm - 0.5).min(1) < 4e-6) | (np.abs(pm[:, 0] - pm[:, 1]) < 4e-6)
        assert not (diff & ~near).any()


def test_atnet_glue_kernels_vs_torch():
    """reflect_pad / sigmoid_blend / assemble one by one against torch on the same device (odd sizes, unaligned tails)."""
    from ivosw.engine import get_engine
    eng = get_engine("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 3, 37, 53), device="cuda", generator=g)
    for (l, r, t, b) in ((14, 13, 12, 11), (0, 0, 0, 0), (52, 1, 36, 0)):
        assert torch.equal(eng.reflect_pad(x, l, r, t, b), torch.nn.ReflectionPad2d((l, r, t, b))(x))
    logit = 6 * torch.randn((3, 1, 45, 67), device="cuda", generator=g)
    prev = torch.rand((3, 45, 67), device="cuda", generator=g)
    for alpha in (1, 0.5, 0.5 + 0.5 * (3 / 7)):
        want_p = torch.sigmoid(logit)
        want_b = (alpha * want_p[:, 0]) + ((1 - alpha) * prev)
        mine = prev.clone()
        p = eng.sigmoid_blend(logit, prev_inplace=mine, alpha=alpha)
        assert float((p - want_p).abs().max()) <= 2.5e-7        # expf vs torch's sigmoid: a couple of ulp at most
        assert float((mine - want_b).abs().max()) <= 2.5e-7
    pm = torch.rand((5, 2, 96, 128), device="cuda", generator=g)
    want = torch.cat([torch.zeros_like(pm[:, 0:1]), pm], 1)[:, :, 12:-12, 14:-14]
    assert torch.equal(eng.atnet_assemble(pm, 12, 14, 72, 100), want)


# ------------------------------------------------------------------------------------------------------------
# the launcher, end to end: an entry script started with `python -m ivosw.run` runs two interaction rounds
# ------------------------------------------------------------------------------------------------------------
def test_launcher_runs_an_entry_script_on_the_b200_path(tmp_path):
    import json
    import subprocess
    from tests import doubles
    from oracle import manet_tail_ref, round_ref
    out = tmp_path / "mini.json"
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "ivos-w_b200"), os.path.join(doubles.HERE, "manet_repo"),
                                         os.path.join(doubles.HERE, "third_party")])
    script = os.path.join(doubles.CHECKOUT, "mini_eval_manet.py")
    r = subprocess.run([sys.executable, "-m", "ivosw.run", script, str(out)], cwd=doubles.CHECKOUT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    dropin_dir = os.path.join(REPO, "ivos-w_b200", "dropin")
    for m in ("models.agent", "models.assessment", "utils.utils_agent"):
        assert res["modules"][m].startswith(dropin_dir), res["modules"]
    assert res["modules"]["utils.utils_manet"]["file"].startswith(doubles.CHECKOUT)
    assert res["misc_file"].startswith(doubles.CHECKOUT) and res["momory_pool_file"].startswith(doubles.CHECKOUT)
    assert res["name"] == "__main__" and res["steps_done"] == 2
    # the same two rounds through the CPU oracle (MANet tail -> all_P -> recommend_frame)
    T, H, W, O = 8, 128, 224, 2
    all_F, all_P_np, annotated = synth.make_clip(5, T, H, W, O)
    logits = torch.log(torch.from_numpy(all_P_np[:, :, ::4, ::4]).clamp_min(1e-6))
    _, all_P = manet_tail_ref.manet_tail(logits, H, W)
    assess_sd, brain_sd = synth.assess_state_dict(0), synth.brain_state_dict(0)
    ann_list, nxt = [], int(annotated[0])
    for rnd in res["rounds"]:
        ann_list.append(nxt)
        assert rnd["annotated"] == ann_list
        ref = round_ref.recommend_frame_wild_ours(assess_sd, brain_sd, all_F, all_P, ann_list)
        np.testing.assert_allclose(rnd["mask_quality"], ref["mask_quality"], atol=1e-4)
        assert rnd["next_frame"] == ref["next_frame"]
        nxt = rnd["next_frame"]


def test_ipn_style_all_p_view_and_clip_cache_identity(dropin):
    """eval_agent_ipn.py:248,261 hands recommend_frame a TRANSPOSED view (variables['probs'][0].transpose(1, 0));
    and a new clip tensor must never be served the previous clip's cached frames, whatever its address."""
    T, H, W, O = 5, 96, 160, 2
    device = torch.device("cuda:0")
    assess_net = dropin.S.AssessNet()
    assess_net.load_state_dict(synth.assess_state_dict(0), strict=True)
    assess_net = assess_net.to(device).eval()
    agent = dropin.A.Agent(device, _cfg())
    agent.policy_net.load_state_dict(synth.brain_state_dict(0), strict=True)
    res = []
    for cid in (40, 41):
        all_F, all_P, annotated = synth.make_clip(cid, T, H, W, O)
        probs_ipn = torch.from_numpy(all_P).to(device).transpose(1, 0).contiguous()      # (O+1) x T x H x W, as IPN keeps it
        mq_view, mq_contig = np.zeros(T), np.zeros(T)
        kw = dict(n_frame=T, n_objects=O, new_masks_quality=np.zeros(T), prev_frames=[], annotated_frames_list=annotated,
                  first_frame=0, max_nb_interactions=8)
        all_F_cpu = torch.Tensor(all_F)
        a = dropin.U.recommend_frame(SimpleNamespace(setting="wild", method="ours"), assess_net, agent, device,
                                     all_F=all_F_cpu, all_P=probs_ipn.transpose(1, 0), mask_quality=mq_view, **kw)
        b = dropin.U.recommend_frame(SimpleNamespace(setting="wild", method="ours"), assess_net, agent, device,
                                     all_F=all_F_cpu, all_P=torch.from_numpy(all_P).to(device), mask_quality=mq_contig, **kw)
        assert int(a) == int(b)
        np.testing.assert_array_equal(mq_view, mq_contig)
        res.append(mq_view.copy())
        del all_F_cpu                                          # the next clip may well land at the same address
    assert not np.array_equal(res[0], res[1])
    from ivosw.engine import Engine
    e = Engine(0)
    e.load_assess(synth.assess_state_dict(0)); e.load_brain(synth.brain_state_dict(0))
    all_F, all_P, annotated = synth.make_clip(41, T, H, W, O)
    fresh = e.round_device(torch.from_numpy(all_F).cuda(), torch.from_numpy(all_P).cuda(), synth.annotated_counts(annotated, T))
    np.testing.assert_array_equal(res[1], fresh["mask_quality"])        # clip 41 was scored on clip 41's frames
    e.close()
