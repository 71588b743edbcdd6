"""Generates tests/golden/*.npz by running the REFERENCE's own modules
(imported, unmodified, from /root/reference) on CPU over the seeded synthetic
inputs of ivosw.synth.  Run in the build container only (the GPU box has no
/root/reference); the resulting small fixtures are committed.

    python tests/golden/make_golden.py

Harness-side shims (none of them edits the reference):
  * models.assessment.resnet50 is replaced after import by a weights=None
    constructor (the reference asks for pretrained=True -> download, no network);
  * for utils.utils_manet.get_results: stub modules ``config`` (cfg.KNNS = 1)
    and ``davisinteractive.dataset.davis`` are pre-inserted in sys.modules,
    ``torch.Tensor.cuda`` is an identity on this CPU run, and a labelled
    stand-in IntVOS returns synthetic logits.  Only the wrapper's tail
    (upsample / argmax / softmax) is therefore pinned.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import synth  # noqa: E402

sys.path.insert(0, REF)
torch.set_num_threads(8)


def load_reference():
    import torchvision
    import models.assessment as ref_assess
    ref_assess.resnet50 = lambda pretrained=True: torchvision.models.resnet50(weights=None)
    import models.agent as ref_agent
    import utils.utils_agent as ref_glue
    return ref_assess, ref_agent, ref_glue


def agent_cfg():
    return SimpleNamespace(
        phase="eval",
        agent=SimpleNamespace(memory_size=10, gamma=0.95, eps_start=0.7, eps_end=0.25, eps_decay=500,
                              update_rate=0.05, lr=5e-6, weight_decay=5e-4),
        data=SimpleNamespace(subset="train"))


def sub(t, step):
    """Strided probe of a B x C x H x W tensor (keeps fixtures small)."""
    return t[:, ::max(1, t.shape[1] // 8), ::step, ::step].contiguous().numpy()


def golden_round(name, clip_id, T, H, W, O, style, ref_assess, ref_agent, ref_glue, seed=0):
    all_F, all_P, annotated = synth.make_clip(clip_id, T, H, W, O, style)
    assess_sd = synth.assess_state_dict(seed)
    brain_sd = synth.brain_state_dict(seed)
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        net = ref_assess.AssessNet()
        net.load_state_dict(assess_sd, strict=True)
        net = net.to(dt).eval()
        # fp64 arbiter leg only: the reference builds theta with float32 zeros through
        # ToCudaVariable (assessment.py:86); promote that helper so the grid is fp64 too.
        orig_tcv = ref_assess.ToCudaVariable
        if dt == torch.float64:
            ref_assess.ToCudaVariable = lambda xs, requires_grad=False: [x.double() for x in xs]
            # ... and the two hard-coded .float() casts (assessment.py:165,174) must not demote
            orig_float = torch.Tensor.float
            torch.Tensor.float = lambda self, *a, **k: self if self.dtype == torch.float64 else orig_float(self).double()
        agent = ref_agent.Agent("cpu", agent_cfg())
        agent.policy_net.load_state_dict(brain_sd, strict=True)
        agent.policy_net.to(dt)
        # probes through hooks on the reference modules
        probes = {}
        enc = net.Encoder
        hs = [enc.register_forward_pre_hook(lambda m, a: probes.setdefault("roi", []).append((a[0].clone(), a[1].clone())))]
        for nm, attr in (("r2", "res2"), ("r3", "res3"), ("r4", "res4"), ("r5", "res5"), ("pool", "maxpool")):
            hs.append(getattr(enc, attr).register_forward_hook(
                lambda m, a, o, nm=nm: probes.setdefault(nm, []).append(o.clone())))
        boxes = []
        orig = net.all2yxhw
        net.all2yxhw = lambda mask, scale=1.0: boxes.append(orig(mask, scale).clone()) or boxes[-1]
        mask_quality = np.zeros(T)
        q_holder = {}
        fwd = agent.policy_net.forward
        agent.policy_net.forward = lambda x: q_holder.setdefault("q", fwd(x.to(dt)))
        F_t = torch.from_numpy(all_F).to(dt)
        P_t = torch.from_numpy(all_P).to(dt)
        import random
        random.seed(0)
        cfg = SimpleNamespace(setting="wild", method="ours")
        nxt = ref_glue.recommend_frame(cfg, net, agent, "cpu", n_frame=T, n_objects=O, all_F=F_t, all_P=P_t,
                                       new_masks_quality=np.zeros(T), prev_frames=[], annotated_frames_list=annotated,
                                       mask_quality=mask_quality, first_frame=0, max_nb_interactions=8)
        for h in hs:
            h.remove()
        out[tag + "_next_frame"] = np.int64(nxt)
        out[tag + "_mask_quality"] = mask_quality.copy()
        out[tag + "_q"] = q_holder["q"].detach().numpy()[0]
        out[tag + "_boxes"] = torch.stack(boxes, 0).numpy()                 # O x T x 4
        if tag == "f32":
            # per-object scores are recoverable only via a second direct call
            sc = np.stack([net(F_t, P_t[:, i + 1]).detach().numpy().reshape(-1) for i in range(O)], 1)
            out["f32_scores"] = sc
            out["roi_f"] = np.stack([sub(p[0], 8) for p in probes["roi"][:O]], 0)       # O x T x 3 x 32 x 32
            out["roi_p"] = np.stack([p[1][:, ::8, ::8].numpy() for p in probes["roi"][:O]], 0)
            out["pool"] = np.stack([sub(p, 8) for p in probes["pool"][:O]], 0)
            out["r2"] = np.stack([sub(p, 8) for p in probes["r2"][:O]], 0)
            out["r3"] = np.stack([sub(p, 4) for p in probes["r3"][:O]], 0)
            out["r4"] = np.stack([sub(p, 2) for p in probes["r4"][:O]], 0)
            out["r5"] = np.stack([sub(p, 1) for p in probes["r5"][:O]], 0)
        else:
            sc = np.stack([net(F_t, P_t[:, i + 1]).detach().numpy().reshape(-1) for i in range(O)], 1)
            out["f64_scores"] = sc
        ref_assess.ToCudaVariable = orig_tcv
        if dt == torch.float64:
            torch.Tensor.float = orig_float
    out["meta"] = np.array([clip_id, T, H, W, O, seed], dtype=np.int64)
    out["style"] = np.array(style)
    out["annotated"] = np.array(annotated, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "next_frame f32/f64:", out["f32_next_frame"], out["f64_next_frame"],
          "mq:", np.round(out["f32_mask_quality"], 4))


def golden_brain(ref_agent):
    out = {}
    rng = np.random.default_rng(7)
    for seed in (0, 1):
        sd = synth.brain_state_dict(seed)
        for (N, T) in ((1, 1), (1, 8), (1, 64), (1, 128), (4, 25)):
            iou = rng.uniform(0.2, 0.95, size=(N, T))
            ann = np.zeros((N, T))
            for n in range(N):
                for i in rng.integers(0, T, size=rng.integers(1, 6)):
                    ann[n, i] += 1
            x = np.stack([iou, ann], 2)
            for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
                b = ref_agent.Brain()
                b.load_state_dict(sd, strict=True)
                b = b.to(dt).eval()
                with torch.no_grad():
                    q = b(torch.from_numpy(x).to(dt)).numpy()
                out["s%d_N%d_T%d_%s_q" % (seed, N, T, tag)] = q
            out["s%d_N%d_T%d_x" % (seed, N, T)] = x
    np.savez_compressed(os.path.join(HERE, "brain.npz"), **out)
    print("brain.npz:", len(out), "arrays")


def golden_boxes(ref_assess):
    """Known-answer cases of all2yxhw + random masks."""
    net = ref_assess.AssessNet()
    cases = []
    m = np.zeros((480, 854), np.float32); m[100:300, 200:500] = 1; cases.append(m)
    m = np.zeros((480, 854), np.float32); m[200:210, 400:405] = 1; cases.append(m)
    cases.append(np.zeros((480, 854), np.float32))
    cases.append(np.ones((480, 854), np.float32))
    m = np.zeros((480, 854), np.float32); m[0, 0] = 1; cases.append(m)
    m = np.zeros((480, 854), np.float32); m[479, 853] = 1; cases.append(m)
    m = np.zeros((480, 854), np.float32); m[0:3, :] = 1; cases.append(m)
    m = np.zeros((480, 854), np.float32); m[:, 850:] = 1; cases.append(m)
    rng = np.random.default_rng(11)
    for _ in range(24):
        m = np.zeros((480, 854), np.float32)
        y0, x0 = rng.integers(0, 470), rng.integers(0, 840)
        hh, ww = rng.integers(1, 480 - y0 + 1), rng.integers(1, 854 - x0 + 1)
        m[y0:y0 + hh, x0:x0 + ww] = (rng.random((hh, ww)) < 0.3)
        cases.append(m)
    masks = np.stack(cases, 0)
    boxes = net.all2yxhw(torch.from_numpy(masks), scale=1.5).numpy()
    small = np.zeros((3, 37, 53), np.float32); small[1, 5:9, 7:40] = 1; small[2] = 1
    boxes_small = net.all2yxhw(torch.from_numpy(small), scale=1.5).numpy()
    np.savez_compressed(os.path.join(HERE, "boxes.npz"), masks=np.packbits(masks.astype(np.uint8), axis=-1),
                        boxes=boxes, small=small, boxes_small=boxes_small)
    print("boxes.npz: first four", boxes[:4])


def golden_manet_tail():
    """utils/utils_manet.py::get_results on CPU with a stand-in IntVOS."""
    cfgmod = types.ModuleType("config"); cfgmod.cfg = SimpleNamespace(KNNS=1)
    sys.modules["config"] = cfgmod
    for nm in ("davisinteractive", "davisinteractive.dataset", "davisinteractive.dataset.davis"):
        sys.modules[nm] = types.ModuleType(nm)
    sys.modules["davisinteractive.dataset.davis"].Davis = object
    torch.Tensor.cuda = lambda self, *a, **k: self
    import utils.utils_manet as ref_manet

    T, O, h, w, H, W = 5, 2, 15, 27, 60, 107
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((T, O + 1, h, w)).astype(np.float32) * 2

    class StandInIntVOS:  # LABELLED STAND-IN: returns synthetic logits, not MANet
        dynamic_seghead = None

        def int_seghead(self, **kw):
            return {kw["seq_names"][0]: torch.from_numpy(logits[kw["frame_num"][0]])[None]}, kw["local_map_dics"]

        def prop_seghead(self, *a, **kw):
            return ({kw["seq_names"][0]: torch.from_numpy(logits[kw["frame_num"][0]])[None]},
                    kw["global_map_tmp_dic"], kw["local_map_dics"])

    emb = torch.zeros(T, 4, h, w)
    storage = torch.zeros(T, H, W)
    masks, all_P = ref_manet.get_results(StandInIntVOS(), emb[2:3], torch.zeros(1, 1, h, w), None, {}, ({}, {}),
                                         1, "seq", O, 2, True, H, W, storage, T, emb)
    np.savez_compressed(os.path.join(HERE, "manet_tail.npz"), logits=logits, masks=masks.numpy().astype(np.uint8),
                        all_P=all_P.numpy(), storage=storage.numpy().astype(np.uint8), hw=np.array([H, W]))
    print("manet_tail.npz", masks.shape, all_P.shape)


def synth_dqn_batch(seed, N=256, T=25):
    """BASELINE config C4 inputs: IoU ~ U[0.3, 0.95], annotated counts = histogram of
    k in [1, 5] random frames, action ~ U{0..T-1}, reward_step in {-1, +1}, reward_done ~ N(0, 1)."""
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
    return {
        "old_state_iou": torch.from_numpy(rng.uniform(0.3, 0.95, (N, T))),
        "new_state_iou": torch.from_numpy(rng.uniform(0.3, 0.95, (N, T))),
        "annotated_frames": torch.from_numpy(a0), "next_annotated_frames": torch.from_numpy(a1),
        "action": torch.from_numpy(act), "reward_step": torch.from_numpy(rng.choice([-1.0, 1.0], N)),
        "reward_done": torch.from_numpy(rng.standard_normal(N)), "done": torch.zeros(N),
    }


def golden_dqn(ref_agent):
    """Two consecutive Agent.update_agent steps of the reference on CPU (target net = a second seed)."""
    out = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        cfg = agent_cfg()
        cfg.phase = "train"
        agent = ref_agent.Agent("cpu", cfg)
        agent.policy_net.load_state_dict(synth.brain_state_dict(0))
        agent.target_net.load_state_dict(synth.brain_state_dict(1))
        agent.policy_net.to(dt); agent.target_net.to(dt)
        agent.optimizer = torch.optim.Adam(agent.policy_net.parameters(), lr=cfg.agent.lr,
                                           weight_decay=cfg.agent.weight_decay)
        np.random.seed(12345)       # the stochastic target sync must not fire (update_rate 0.05)
        for step in range(2):
            batch = synth_dqn_batch(100 + step, 64 if dt == torch.float64 else 256, 25)
            if dt == torch.float64:
                # the reference casts inputs with .float(); keep the arbiter in fp64 by pre-casting its tensors
                orig_float = torch.Tensor.float
                torch.Tensor.float = lambda self, *a, **k: self.double() if self.is_floating_point() else orig_float(self).double()
            before = np.random.get_state()[2]
            loss = agent.update_agent(batch)
            if dt == torch.float64:
                torch.Tensor.float = orig_float
            out["%s_loss%d" % (tag, step)] = np.float64(loss)
            if tag == "f32":    # fixtures kept small: full gradients of step 0, every 8th value afterwards
                for k, v in agent.policy_net.named_parameters():
                    g = v.grad.detach().numpy().reshape(-1)
                    out["grad%d_%s" % (step, k)] = g.copy() if step == 0 else g[::8].copy()
                    if step == 1:
                        out["param_final_%s" % k] = v.detach().numpy().reshape(-1)[::8].copy()
        tgt_same = all(torch.equal(a, b.to(dt)) for a, b in zip(agent.target_net.state_dict().values(),
                                                                 synth.brain_state_dict(1).values()))
        assert tgt_same, "target sync fired; pick another numpy seed"
    np.savez_compressed(os.path.join(HERE, "dqn_step.npz"), **out)
    print("dqn_step.npz: loss f32", out["f32_loss0"], out["f32_loss1"], "f64", out["f64_loss0"], out["f64_loss1"])


TRAIN_KEYS = ("fc1.weight", "fc1.bias", "Encoder.conv1.weight", "Encoder.conv1_p.weight", "Encoder.bn1.weight",
              "Encoder.bn1.bias", "Encoder.res2.0.conv1.weight", "Encoder.res2.0.downsample.0.weight",
              "Encoder.res3.1.conv2.weight", "Encoder.res3.1.bn2.weight", "Encoder.res4.3.conv3.weight",
              "Encoder.res5.2.conv3.weight", "Encoder.res5.2.bn3.bias")
TRAIN_BUFFERS = ("Encoder.bn1.running_mean", "Encoder.bn1.running_var", "Encoder.bn1.num_batches_tracked",
                 "Encoder.res3.1.bn2.running_mean", "Encoder.res5.2.bn3.running_var")
TRAIN_HP = dict(lr=5e-6, momentum=0.9, weight_decay=5e-4)        # config.yaml:25-28


def synth_train_batch(step, N=4, H=160, W=256):
    """imgs N x 3 x H x W, probs N x H x W (object 1 of a synthetic clip), targets U[0,1] standing in for the J&F
    metric (davisinteractive is not installed), valid[n] standing in for `union[n] > 0` (one sample skipped)."""
    all_F, all_P, _ = synth.make_clip(40 + step, N, H, W, 1)
    rng = np.random.default_rng(7000 + step)
    targets = rng.random(N).astype(np.float32)
    valid = np.ones(N, dtype=bool)
    valid[(step + 2) % N] = False
    return all_F, all_P[:, 1], targets, valid


def train_state_dict():
    """synth.assess_state_dict(0) with fc1.weight scaled by 1e-3: most encoder gradients then stay inside the
    [-1, 1] clamp (their fp32 round-off, amplified by 53 layers of train-mode BatchNorm, is ~1e-3 relative — clamped
    values would hide everything but the sign), while the fc1 gradients still exceed it and exercise the clamp."""
    sd = synth.assess_state_dict(0)
    sd["fc1.weight"] = sd["fc1.weight"] * 1e-3
    return sd


def golden_assess_train(ref_assess):
    """Two consecutive iterations of quality_assessment.py::train's loop body (:240-269) with the reference's own
    AssessNet in train mode and torch.optim.SGD — no zero_grad between them, exactly like the loop."""
    import torch.nn.functional as F
    out = {}
    net = ref_assess.AssessNet()
    net.load_state_dict(train_state_dict(), strict=True)
    net.train()                                                              # :213
    opt = torch.optim.SGD(net.parameters(), **TRAIN_HP)                      # :309-310
    params = dict(net.named_parameters())
    for step in range(2):
        imgs, probs, targets, valid = synth_train_batch(step)
        iou_pred = net(torch.from_numpy(imgs), torch.from_numpy(probs))      # :240
        metric_gt = torch.from_numpy(targets)
        loss, counter = 0., 0
        for n in range(imgs.shape[0]):                                       # :251-257
            if valid[n]:
                loss += F.mse_loss(iou_pred[n], metric_gt[n])
                counter += 1
        loss /= counter
        loss.backward()                                                      # :265
        for param in net.parameters():                                       # :266-268
            if param.grad is not None:
                param.grad.data.clamp_(-1, 1)
        opt.step()                                                           # :269
        out["pred%d" % step] = iou_pred.detach().numpy().reshape(-1)
        out["loss%d" % step] = np.float64(loss.item())
        for k in TRAIN_KEYS:
            stride = 101 if params[k].numel() > 10000 else 1         # fixtures kept small
            out["grad%d_%s" % (step, k)] = params[k].grad.detach().numpy().reshape(-1)[::stride].copy()
            out["param%d_%s" % (step, k)] = params[k].detach().numpy().reshape(-1)[::stride].copy()
    sd = net.state_dict()
    for k in TRAIN_BUFFERS:
        out["buf_" + k] = sd[k].numpy().reshape(-1).copy()
    out["no_grad_params"] = np.array(sorted(k for k, v in params.items() if v.grad is None))
    np.savez_compressed(os.path.join(HERE, "assess_train.npz"), **out)
    print("assess_train.npz: loss", out["loss0"], out["loss1"], "params without grad:", list(out["no_grad_params"]))


if __name__ == "__main__":
    ref_assess, ref_agent, ref_glue = load_reference()
    with torch.no_grad():
        golden_boxes(ref_assess)
        golden_brain(ref_agent)
        golden_round("round_c1", 0, 8, 256, 448, 2, "manet", ref_assess, ref_agent, ref_glue)
        golden_round("round_atnet_small", 1, 5, 192, 320, 1, "atnet", ref_assess, ref_agent, ref_glue)
        golden_round("round_single", 2, 1, 160, 288, 3, "manet", ref_assess, ref_agent, ref_glue)
        golden_round("round_t16", 4, 16, 128, 224, 2, "manet", ref_assess, ref_agent, ref_glue)
        golden_manet_tail()
    golden_dqn(ref_agent)
    golden_assess_train(ref_assess)
