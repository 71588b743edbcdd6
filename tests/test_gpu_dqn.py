"""GPU (-m gpu): the Double-DQN training step (BASELINE config C4: batch 256 x T 25) through the C ABI against
the fixture produced by the reference's own Agent.update_agent (tests/golden/dqn_step.npz) and the oracle.

Tolerances: loss 2e-6 relative; clamped gradients |d| <= 2e-6 + 2e-3 |ref| (fp32 accumulation over 6400 rows in a
different order); parameters after two Adam steps |d| <= 2e-6 (Adam's first steps move every weight by ~lr = 5e-6,
so a sign error in any gradient would show as 1e-5)."""
import os

import numpy as np
import pytest
import torch

from ivosw import arch, synth

pytestmark = pytest.mark.gpu


def _dqn_batch(seed, N=256, T=25):
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


def test_dqn_two_steps_vs_reference(golden_dir):
    from ivosw.engine import Engine
    g = np.load(os.path.join(golden_dir, "dqn_step.npz"))
    eng = Engine(0)
    eng.load_brain(synth.brain_state_dict(0))
    eng.load_target(synth.brain_state_dict(1))
    eng.reset_optimizer()
    for step in range(2):
        s, ns, act, rs, rd = _dqn_batch(100 + step)
        loss, grads = eng.dqn_update(torch.from_numpy(s).float().cuda(), torch.from_numpy(ns).float().cuda(),
                                     torch.from_numpy(act).cuda(), torch.from_numpy(rs).float().cuda(),
                                     torch.from_numpy(rd).float().cuda(), gamma=0.95, lr=5e-6, weight_decay=5e-4,
                                     want_grads=True)
        ref_loss = float(g["f32_loss%d" % step])
        assert abs(loss - ref_loss) <= 2e-6 * abs(ref_loss) + 1e-9, (loss, ref_loss)
        gd = eng.unpack_brain(grads)
        for k, _ in arch.BRAIN_PARAMS:
            mine = gd[k].cpu().numpy().reshape(-1)
            mine = mine if step == 0 else mine[::8]
            np.testing.assert_allclose(mine, g["grad%d_%s" % (step, k)], atol=2e-6, rtol=2e-3, err_msg="%s step %d" % (k, step))
    final = eng.unpack_brain(eng.brain_params("policy"))
    for k, _ in arch.BRAIN_PARAMS:
        np.testing.assert_allclose(final[k].cpu().numpy().reshape(-1)[::8], g["param_final_" + k], atol=2e-6, err_msg=k)
    # the target network is untouched until synced
    tgt = eng.unpack_brain(eng.brain_params("target"))
    for k, v in synth.brain_state_dict(1).items():
        np.testing.assert_array_equal(tgt[k].cpu().numpy(), v.numpy())
    eng.sync_target()
    np.testing.assert_array_equal(eng.brain_params("target").cpu().numpy(), eng.brain_params("policy").cpu().numpy())
    # the inference path sees the updated policy
    x = torch.from_numpy(_dqn_batch(5, 3, 25)[0]).float().cuda()
    from oracle import brain_ref
    q = eng.brain_forward(x).cpu().numpy()
    ref = brain_ref.brain_forward({k: v.cpu().numpy() for k, v in final.items()}, x.cpu().numpy())
    np.testing.assert_allclose(q, ref, atol=1e-5)
    eng.close()


def test_dqn_vs_oracle_other_shapes():
    from ivosw.engine import Engine
    from oracle import dqn_ref
    eng = Engine(0)
    for (N, T, seed) in ((7, 5, 1), (33, 40, 2)):
        eng.load_brain(synth.brain_state_dict(seed))
        eng.load_target(synth.brain_state_dict(seed + 1))
        eng.reset_optimizer()
        st = dqn_ref.DqnState(synth.brain_state_dict(seed), synth.brain_state_dict(seed + 1))
        s, ns, act, rs, rd = _dqn_batch(seed, N, T)
        ref_loss, ref_g = dqn_ref.update_agent(st, torch.from_numpy(s).float(), torch.from_numpy(ns).float(),
                                               torch.from_numpy(act), torch.from_numpy(rs).float(), torch.from_numpy(rd).float())
        loss, grads = eng.dqn_update(torch.from_numpy(s).float().cuda(), torch.from_numpy(ns).float().cuda(),
                                     torch.from_numpy(act).cuda(), torch.from_numpy(rs).float().cuda(),
                                     torch.from_numpy(rd).float().cuda(), want_grads=True)
        assert abs(loss - ref_loss) <= 5e-6 * abs(ref_loss) + 1e-9
        gd = eng.unpack_brain(grads)
        for k, _ in arch.BRAIN_PARAMS:
            np.testing.assert_allclose(gd[k].cpu().numpy(), ref_g[k].numpy(), atol=3e-6, rtol=3e-3, err_msg=k)
    eng.close()


def test_dqn_data_parallel_arithmetic():
    """Two engines standing in for two ranks: raw gradients of the half batches, averaged, then clamp + Adam,
    must reproduce the single full-batch update."""
    from ivosw.engine import Engine
    full, r0, r1 = Engine(0), Engine(0), Engine(0)
    for e in (full, r0, r1):
        e.load_brain(synth.brain_state_dict(0)); e.load_target(synth.brain_state_dict(1)); e.reset_optimizer()
    s, ns, act, rs, rd = _dqn_batch(100)
    t = lambda a, sl: torch.from_numpy(a[sl]).float().cuda()
    loss_full, g_full = full.dqn_update(t(s, slice(None)), t(ns, slice(None)), torch.from_numpy(act).cuda(), t(rs, slice(None)),
                                        t(rd, slice(None)), want_grads=True)
    halves = []
    for e, sl in ((r0, slice(0, 128)), (r1, slice(128, 256))):
        halves.append(e.dqn_update(t(s, sl), t(ns, sl), torch.from_numpy(act[sl]).cuda(), t(rs, sl), t(rd, sl), apply=False))
    loss_dp = 0.5 * (halves[0][0] + halves[1][0])
    g = 0.5 * (halves[0][1] + halves[1][1])
    assert abs(loss_dp - loss_full) <= 2e-6 * abs(loss_full)
    for e in (r0, r1):
        gg = g.clone()
        e.dqn_apply(gg)
        np.testing.assert_allclose(gg.cpu().numpy(), g_full.cpu().numpy(), atol=1e-6, rtol=1e-3)
        np.testing.assert_allclose(e.brain_params().cpu().numpy(), full.brain_params().cpu().numpy(), atol=1e-6)
    for e in (full, r0, r1):
        e.close()


def test_dropin_agent_update_matches_reference(golden_dir):
    """The drop-in Agent.update_agent (same signature / sample dict as agent.py:103) against the reference fixture."""
    import sys
    from types import SimpleNamespace
    from tests import doubles
    A = doubles.load_dropin().A
    g = np.load(os.path.join(golden_dir, "dqn_step.npz"))
    cfg = SimpleNamespace(phase="train", agent=SimpleNamespace(memory_size=10, gamma=0.95, eps_start=0.7, eps_end=0.25,
                          eps_decay=500, update_rate=0.05, lr=5e-6, weight_decay=5e-4), data=SimpleNamespace(subset="train"))
    agent = A.Agent("cuda:0", cfg)
    agent.policy_net.load_state_dict(synth.brain_state_dict(0))
    agent.target_net.load_state_dict(synth.brain_state_dict(1))
    from ivosw.engine import get_engine
    get_engine("cuda:0").reset_optimizer()
    np.random.seed(12345)
    for step in range(2):
        s, ns, act, rs, rd = _dqn_batch(100 + step)
        sample = {"old_state_iou": torch.from_numpy(s[..., 0]), "annotated_frames": torch.from_numpy(s[..., 1]),
                  "new_state_iou": torch.from_numpy(ns[..., 0]), "next_annotated_frames": torch.from_numpy(ns[..., 1]),
                  "action": torch.from_numpy(act), "reward_step": torch.from_numpy(rs), "reward_done": torch.from_numpy(rd),
                  "done": torch.zeros(len(act))}
        loss = agent.update_agent(sample)
        assert abs(loss - float(g["f32_loss%d" % step])) <= 2e-6 * abs(loss)
    for k, v in agent.policy_net.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy().reshape(-1)[::8], g["param_final_" + k], atol=2e-6, err_msg=k)
    assert abs(agent.get_avg_loss() - 0.5 * (float(g["f32_loss0"]) + float(g["f32_loss1"]))) < 1e-6
