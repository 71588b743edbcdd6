"""TEST DOUBLE for ATNet's libs/custom_transforms.py (utils/utils_atnet.py:59-61)."""
import torch

IS_TEST_DOUBLE = True


class Normalize_ApplymeanvarImage(object):
    def __init__(self, mean, var):
        self.mean, self.var = mean, var

    def __call__(self, sample):
        sample = dict(sample)
        sample['image'] = (sample['image'] - self.mean) / self.var
        return sample


class ToTensor(object):
    def __call__(self, sample):
        sample = dict(sample)
        sample['image'] = torch.from_numpy(sample['image'].transpose(2, 0, 1).copy()).float()
        return sample
