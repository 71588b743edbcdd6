"""ORACLE (test infrastructure, not product code) — CPU restatement of the
reference Q-network ``models/agent.py::Brain.forward`` (lines 33-64) and of the
greedy branch of ``Agent.action`` (lines 168-189).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product path never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
against tests/golden/*.npz, which were produced by importing the reference's own
``models.agent.Brain`` from /root/reference (tests/golden/make_golden.py).

Plain numpy; ``dtype`` selects fp32 (the oracle) or fp64 (the arbiter used to
judge near-ties, SURVEY.md §8(c) "Oracle precision policy").
"""
import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def brain_forward(sd, x, dtype=np.float32):
    """sd: dict key -> ndarray (the 10 tensors of Brain.state_dict()).
    x: N x T x 2.  Returns Q: N x T.

    agent.py:46-47  e_t = fc2(relu(fc1(x_t)))          (no ReLU after fc2)
    agent.py:48-49  one shared LSTMCell(128,128,bias=False) run forward over
                    t = 0..T-1 and backward over t = T-1..0, zero initial state;
                    torch gate order i, f, g, o.
    agent.py:55-60  Q_t = fc_d2(relu(fc_d1(relu([h_fw_t ; h_bw_t])))).
    """
    p = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    x = np.asarray(x, dtype=dtype)
    N, T, _ = x.shape
    Hd = 128

    def enc(xt):
        a = np.maximum(xt @ p["encoder_fc1.weight"].T + p["encoder_fc1.bias"], 0)
        return a @ p["encoder_fc2.weight"].T + p["encoder_fc2.bias"]

    def cell(e, h, c):
        g = e @ p["lstm_cell.weight_ih"].T + h @ p["lstm_cell.weight_hh"].T
        i, f, gg, o = g[:, :Hd], g[:, Hd:2 * Hd], g[:, 2 * Hd:3 * Hd], g[:, 3 * Hd:]
        c2 = _sigmoid(f) * c + _sigmoid(i) * np.tanh(gg)
        h2 = _sigmoid(o) * np.tanh(c2)
        return h2.astype(dtype), c2.astype(dtype)

    h_fw = np.zeros((T, N, Hd), dtype)
    h_bw = np.zeros((T, N, Hd), dtype)
    h = np.zeros((N, Hd), dtype); c = np.zeros((N, Hd), dtype)
    for t in range(T):
        h, c = cell(enc(x[:, t]), h, c)
        h_fw[t] = h
    h = np.zeros((N, Hd), dtype); c = np.zeros((N, Hd), dtype)
    for t in range(T - 1, -1, -1):
        h, c = cell(enc(x[:, t]), h, c)
        h_bw[t] = h
    q = np.empty((N, T), dtype)
    for t in range(T):
        s = np.maximum(np.concatenate([h_fw[t], h_bw[t]], 1), 0)
        d = np.maximum(s @ p["decoder_fc1.weight"].T + p["decoder_fc1.bias"], 0)
        q[:, t] = (d @ p["decoder_fc2.weight"].T + p["decoder_fc2.bias"])[:, 0]
    return q


def agent_action_greedy(sd, state):
    """agent.py:176,185-189 with eps_threshold = 0 (eval): state (T x 2, float64)
    is cast to fp32, Q = policy_net(state[None]); returns (argmax, Q[T]).
    numpy argmax: the first maximum wins."""
    q = brain_forward(sd, np.asarray(state, dtype=np.float32)[None], np.float32)[0]
    return int(q.argmax()), q
