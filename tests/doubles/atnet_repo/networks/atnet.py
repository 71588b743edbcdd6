"""TEST DOUBLE for ATNet's networks/atnet.py::ATnet — NOT the network: cheap deterministic functions of the same
inputs with the same return structure (utils/utils_atnet.py:99-104, 121-122), so that the wrapper around it can be
driven and checked.  Works on CPU and CUDA tensors."""
import torch

IS_TEST_DOUBLE = True


class _Enc3(object):
    def forward(self, image):
        r2 = image.mean(1, keepdim=True)[:, :, ::4, ::4]
        r5 = image[:, :, ::16, ::16].mean(1, keepdim=True)
        return r5, None, None, r2


class ATnet(object):
    def __init__(self):
        self.encoder_3ch = _Enc3()

    def forward_ANet(self, inputs):                     # n_obj x 6 x P_H x P_W
        image, planes = inputs[:, :3], inputs[:, 3:6]
        logit = 3.0 * planes[:, 1:2] - 2.0 * planes[:, 2:3] + 1.5 * planes[:, 0:1] - 1.0 + 0.8 * image.mean(1, keepdim=True)
        return logit, inputs[:, :, ::16, ::16].mean(1, keepdim=True)

    def forward_TNet(self, r5_3ch_list, image, r5_6ch_list, r2_prev, predmask_prev):
        bias = sum(float(t.mean()) for t in r5_6ch_list) / len(r5_6ch_list)
        logit = 5.0 * predmask_prev - 2.5 + 0.6 * image.mean(1, keepdim=True) + 0.1 * bias
        logit = logit + 0.3 * torch.nn.functional.interpolate(r2_prev, size=logit.shape[-2:], mode='nearest')
        return logit, image.mean(1, keepdim=True)[:, :, ::4, ::4]
