"""TEST DOUBLE for the checkout's datasets/davis_dataset.py::DAVIS2017 (the real one reads DAVIS JPEGs through cv2):
serves seeded synthetic frames, already padded as ATNet's loader pads them, in the order of custom_frames."""
import numpy as np
import torch

IS_TEST_DOUBLE = True
FRAME_HW = (72, 100)          # un-padded frame size of the synthetic sequence
PAD = ((12, 12), (14, 14))    # what libs.utils.apply_pad gives for 72 x 100


def synthetic_frame(idx):
    rng = np.random.default_rng(777 + idx)
    h, w = FRAME_HW
    base = rng.random((h // 8 + 1, w // 8 + 1, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:h, :w]
    return np.pad(img, (PAD[0], PAD[1], (0, 0)), mode='reflect')


class DAVIS2017(torch.utils.data.Dataset):
    def __init__(self, split='val', subseq=None, root='', custom_frames=None, transform=None, retname=False, seq_name=None,
                 rgb=False, obj_id=None, no_gt=False, prev_round_masks=None, **kw):
        self.frames = list(custom_frames)
        self.transform = transform

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        f = self.frames[i]
        sample = {'image': synthetic_frame(f)}
        if self.transform is not None:
            sample = self.transform(sample)
        sample['meta'] = {'frame_id': f}
        return sample
