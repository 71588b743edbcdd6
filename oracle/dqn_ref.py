"""ORACLE (test infrastructure, not product code) — CPU restatement of the Double-DQN training step
``models/agent.py::Agent.update_agent`` (lines 103-166) for BASELINE config C4
(batch 256 x T 25 replay samples).

Parity status: PINNED.  tests/golden/dqn_step.npz holds loss, clamped gradients and the parameters
after one and two steps produced by the reference's own ``Agent.update_agent`` on CPU
(tests/golden/make_golden.py); tests/test_oracle_golden.py checks this restatement against it.

The forward is brain_ref's numpy restatement re-expressed in torch so autograd supplies the
backward; optimiser arithmetic (element-wise grad clamp to +-1, then torch.optim.Adam with L2
weight decay folded into the gradient, agent.py:101,157-160) is restated explicitly.
"""
import numpy as np
import torch
import torch.nn.functional as F


def brain_forward_torch(p, x):
    """p: dict of tensors (Brain.state_dict keys), x: N x T x 2 -> N x T  (agent.py:33-64)."""
    N, T, _ = x.shape
    Hd = 128

    def enc(xt):
        return F.linear(F.relu(F.linear(xt, p["encoder_fc1.weight"], p["encoder_fc1.bias"])),
                        p["encoder_fc2.weight"], p["encoder_fc2.bias"])

    def cell(e, h, c):
        g = F.linear(e, p["lstm_cell.weight_ih"]) + F.linear(h, p["lstm_cell.weight_hh"])
        i, f, gg, o = g[:, :Hd], g[:, Hd:2 * Hd], g[:, 2 * Hd:3 * Hd], g[:, 3 * Hd:]
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        return torch.sigmoid(o) * torch.tanh(c2), c2

    z = x.new_zeros((N, Hd))
    h, c, hf = z, z, []
    for t in range(T):
        h, c = cell(enc(x[:, t]), h, c)
        hf.append(h)
    h, c, hb = z, z, [None] * T
    for t in range(T - 1, -1, -1):
        h, c = cell(enc(x[:, t]), h, c)
        hb[t] = h
    q = []
    for t in range(T):
        s = F.relu(torch.cat([hf[t], hb[t]], 1))
        q.append(F.linear(F.relu(F.linear(s, p["decoder_fc1.weight"], p["decoder_fc1.bias"])),
                          p["decoder_fc2.weight"], p["decoder_fc2.bias"]))
    return torch.cat(q, 1)


class DqnState:
    """Policy / target parameters and Adam moments (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8)."""

    def __init__(self, policy_sd, target_sd=None, lr=5e-6, weight_decay=5e-4, gamma=0.95, dtype=torch.float32):
        self.p = {k: v.clone().to(dtype).requires_grad_(True) for k, v in policy_sd.items()}
        tsd = target_sd if target_sd is not None else policy_sd
        self.t = {k: v.clone().to(dtype) for k, v in tsd.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.step = 0
        self.lr, self.wd, self.gamma = lr, weight_decay, gamma


def update_agent(st, state, new_state, action, reward_step, reward_done):
    """One update.  state/new_state: N x T x 2, action: N (int64), rewards: N.  Returns (loss, grads dict)
    where grads are the CLAMPED gradients (what the optimiser consumed).  Target sync is the caller's
    decision (agent.py:163-165 draws np.random.random())."""
    dt = next(iter(st.p.values())).dtype
    state, new_state = state.to(dt), new_state.to(dt)
    N = state.shape[0]
    action = action.view(N, 1).long()
    rs, rd = reward_step.to(dt).view(N, 1), reward_done.to(dt).view(N, 1)
    with torch.no_grad():                                                     # :131-143
        out = brain_forward_torch({k: v.detach() for k, v in st.p.items()}, new_state)
        next_action = out.max(1)[1].view(N, 1)
        q_next = brain_forward_torch(st.t, new_state).gather(1, next_action)
        y_step = q_next * torch.tensor(st.gamma, dtype=torch.float32).to(dt) + rs * 0.1
        y_done = rd * 0.1
    q_sa = brain_forward_torch(st.p, state).gather(1, action)                 # :146-147
    loss = F.mse_loss(q_sa, y_step) + F.mse_loss(q_sa, y_done)                # :151-153
    for v in st.p.values():
        v.grad = None
    loss.backward()
    grads = {}
    st.step += 1
    b1, b2, eps = 0.9, 0.999, 1e-8
    with torch.no_grad():
        for k, v in st.p.items():
            g = v.grad.clamp(-1, 1)                                           # :157-159
            grads[k] = g.clone()
            g = g + st.wd * v                                                 # Adam(weight_decay): L2 in the gradient
            st.m[k].mul_(b1).add_(g, alpha=1 - b1)
            st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1, bc2 = 1 - b1 ** st.step, 1 - b2 ** st.step
            denom = (st.v[k].sqrt() / np.sqrt(bc2)).add_(eps)
            v.addcdiv_(st.m[k], denom, value=-st.lr / bc1)
    return float(loss.detach()), grads
