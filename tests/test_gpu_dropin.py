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
    p = os.path.join(REPO, "ivos-w_b200", "dropin")
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[m]
    sys.path.insert(0, p)
    import models.agent as A
    import models.assessment as S
    import utils.utils_agent as U
    import utils.utils_manet as M
    sys.path.remove(p)
    return SimpleNamespace(A=A, S=S, U=U, M=M)


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
