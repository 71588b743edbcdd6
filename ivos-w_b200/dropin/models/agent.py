"""Drop-in for the reference's ``models/agent.py`` (Brain 13-64, Agent 67-237).

Same class names, constructor arguments, ``state_dict`` keys and ``forward`` /
``action`` signatures; the Q-network arithmetic runs in the CUDA library
(ivos-w_b200/csrc/brain.cu) through the C ABI.  The modules hold ordinary
``nn.Parameter``s only so that ``load_state_dict(strict=True)`` and
``utils.misc.load_agent_checkpoint`` keep working; the parameters are re-packed
and uploaded to the library whenever they change.  There is no PyTorch forward:
calling these modules without a CUDA device raises.
"""
import math
import random

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from ivosw import arch
from ivosw.engine import get_engine


class Brain(nn.Module):
    def __init__(self, lstm_input_channels=128, hidden_channels=128, num_fc_concat=128):
        super(Brain, self).__init__()
        if (lstm_input_channels, hidden_channels, num_fc_concat) != (128, 128, 128):
            raise NotImplementedError("the CUDA Q-network is built for the reference's 128/128/128 Brain")
        self.input_channels = lstm_input_channels
        self.hidden_channels = hidden_channels
        self.num_fc_concat = num_fc_concat
        # parameter containers with the reference's names and default initialisation (agent.py:19-29)
        self.encoder_fc1 = nn.Linear(2, 128)
        self.encoder_fc2 = nn.Linear(128, self.input_channels)
        self.lstm_cell = nn.LSTMCell(self.input_channels, self.hidden_channels, False)
        self.decoder_fc1 = nn.Linear(2 * self.hidden_channels, self.num_fc_concat)
        self.decoder_fc2 = nn.Linear(self.num_fc_concat, 1)
        self._uploaded = None

    def _sync(self, engine):
        sd = self.state_dict()
        stamp = tuple((sd[k].data_ptr(), sd[k]._version) for k, _ in arch.BRAIN_PARAMS) + (id(engine),)
        if stamp != self._uploaded:
            engine.load_brain(sd)
            self._uploaded = stamp

    def forward(self, input):
        """input: N x T x 2 -> N x T  (agent.py:33-64)."""
        if not input.is_cuda:
            raise RuntimeError("ivosw_b200 Brain runs on CUDA only (no CPU fallback); got a %s tensor" % input.device)
        engine = get_engine(input.device)
        self._sync(engine)
        return engine.brain_forward(input)

    def q_and_argmax(self, input):
        engine = get_engine(input.device)
        self._sync(engine)
        return engine.brain_forward(input, want_argmax=True)


# attribute <- cfg.agent field: the public attributes of the reference's Agent (agent.py:71-80)
_AGENT_CFG = (("memory_size", "memory_size"), ("GAMMA", "gamma"), ("EPS_START", "eps_start"), ("EPS_END", "eps_end"),
              ("EPS_DECAY", "eps_decay"), ("update_rate", "update_rate"))
_LOSS_WINDOW = 32            # running mean of the last 32 update losses (agent.py:94-101, 198-207)


class Agent(nn.Module):
    def __init__(self, device, cfg):
        super().__init__()
        self.cfg, self.device = cfg, device
        for attr, field in _AGENT_CFG:
            setattr(self, attr, getattr(cfg.agent, field))
        self.subset = cfg.data.subset
        self.steps_done = 0
        from models.momory_pool import ReplayMemory
        self.memory_pool = ReplayMemory(self.memory_size)

        # two Q-networks with identical initial weights, both on `device` (agent.py:84-91)
        self.policy_net, self.target_net = Brain(), Brain()
        self.target_net.load_state_dict(self.policy_net.state_dict())
        for net in (self.policy_net, self.target_net):
            net.to(device)

        self.loss, self.loss_position, self.loss_capacity, self.loss_avg = [], 0, _LOSS_WINDOW, 0
        self.optimizer = optim.Adam(self.policy_net.parameters(), lr=cfg.agent.lr,
                                    weight_decay=cfg.agent.weight_decay)

    def _epsilon(self):
        """Exploration threshold: 0 outside training, else an exponential decay in steps_done (agent.py:171-175)."""
        if self.cfg.phase != 'train':
            return 0
        return self.EPS_END + (self.EPS_START - self.EPS_END) * math.exp(-0.5 * self.steps_done / self.EPS_DECAY)

    def action(self, state, verbose=True):
        """agent.py:168-196.  Side effects kept: steps_done += 1 and exactly one random.random()
        draw per call (SURVEY A.Q6) so the global RNG stream stays aligned with the reference."""
        self.steps_done += 1
        eps_threshold = self._epsilon()
        state = np.asarray(state)
        rand_flag = random.random()
        greedy = rand_flag > eps_threshold
        if verbose:
            print(f"step:{self.steps_done}, rand_flag:{rand_flag:.4f}, eps_threshold:{eps_threshold:.4f}, "
                  f"frame index was selected {'by agent' if greedy else 'randomly'}")
        if not greedy:
            return random.choice(np.array(range(state.shape[0])))
        engine = get_engine(self.device)
        self.policy_net._sync(engine)
        action, _ = engine.agent_action(state[:, 0], state[:, 1])
        return np.int64(action)

    def _sync_target(self, engine):
        sd = self.target_net.state_dict()
        stamp = tuple((sd[k].data_ptr(), sd[k]._version) for k, _ in arch.BRAIN_PARAMS) + (id(engine),)
        if stamp != getattr(self, "_target_uploaded", None):
            engine.load_target(sd)
            self._target_uploaded = stamp

    def update_agent(self, sample):
        """Double-DQN step (agent.py:103-166) as one device-side update in the CUDA library: no-grad target
        computation, forward/backward of the bi-LSTM Q-network, element-wise gradient clamp, Adam with L2 weight
        decay.  The Adam moments live inside the library (one training agent per device); the updated
        parameters are copied back into ``policy_net`` so checkpoints keep working."""
        if sample is None:
            print('no input')
            return
        N = sample['action'].shape[0]

        def col(k):
            return sample[k].float().view(N, -1).to(self.device)

        state = torch.stack([col('old_state_iou'), col('annotated_frames')], 2)
        new_state = torch.stack([col('new_state_iou'), col('next_annotated_frames')], 2)
        engine = get_engine(self.device)
        self.policy_net._sync(engine)
        self._sync_target(engine)
        group = self.optimizer.param_groups[0]
        loss = engine.dqn_update(state, new_state, sample['action'].view(N), col('reward_step').view(N),
                                 col('reward_done').view(N), gamma=float(self.GAMMA), lr=float(group['lr']),
                                 weight_decay=float(group['weight_decay']))
        new_sd = engine.unpack_brain(engine.brain_params("policy"))
        with torch.no_grad():
            for k, p in self.policy_net.named_parameters():
                p.copy_(new_sd[k])
        sd = self.policy_net.state_dict()     # the library already holds these values: refresh the stamp, no re-upload
        self.policy_net._uploaded = tuple((sd[k].data_ptr(), sd[k]._version) for k, _ in arch.BRAIN_PARAMS) + (id(engine),)
        self._update_avg_loss(loss)
        if np.random.random() < self.update_rate:                      # stochastic hard target sync (:163-165)
            print("target_net updated!")
            self.target_net.load_state_dict(self.policy_net.state_dict())
        return loss

    def _update_avg_loss(self, loss):
        """Ring buffer of the last `loss_capacity` losses and their mean (agent.py:198-207)."""
        if len(self.loss) < self.loss_capacity:
            self.loss.append(float(loss))
        else:
            self.loss[self.loss_position] = float(loss)
        self.loss_position = (self.loss_position + 1) % self.loss_capacity
        self.loss_avg = sum(self.loss) / len(self.loss)

    def set_train(self):
        for net in (self.policy_net, self.target_net):
            net.train()

    def set_eval(self):
        for net in (self.policy_net, self.target_net):
            net.eval()

    def memory(self, *args):
        self.memory_pool.push(*args[:-1])

    def get_avg_loss(self):
        return self.loss_avg
