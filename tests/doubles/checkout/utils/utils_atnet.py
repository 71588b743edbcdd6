"""TEST DOUBLE for the checkout's utils/utils_atnet.py: same module-level imports as the reference's file
(utils_atnet.py:1-11) so that the drop-in finds DataLoader / transforms / DAVIS2017 / libs in it; its
run_VOS_singleiact must be REPLACED by ivosw.hook."""
import numpy as np  # noqa: F401
import torch  # noqa: F401
from torch.utils.data import DataLoader  # noqa: F401
from torchvision import transforms  # noqa: F401

from datasets.davis_dataset import DAVIS2017  # noqa: F401
from libs import custom_transforms as tr  # noqa: F401
from libs import utils, utils_torch  # noqa: F401

IS_CHECKOUT_ORIGINAL = True


def run_VOS_singleiact(*a, **k):
    raise AssertionError("checkout's run_VOS_singleiact called: the hook did not patch utils.utils_atnet")
