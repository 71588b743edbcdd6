"""TEST DOUBLE for the checkout's utils/misc.py: must stay importable as utils.misc next to the substituted modules."""
import random

import numpy as np
import torch

IS_CHECKOUT_ORIGINAL = True


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def load_agent_checkpoint(agent, state_dict, device, strict=True):
    agent.policy_net.load_state_dict(state_dict, strict=strict)
    return 1


def load_network_checkpoint(state_dict, net, device='cpu'):
    net.load_state_dict(state_dict, strict=True)
    return net
