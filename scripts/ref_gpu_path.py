"""GPU helper (measurement only, not product code): the reference's single-GPU PyTorch path for one scoring round,
restated from its call structure so that it can run on the GPU box (where /root/reference does not exist):

  utils/utils_agent.py:111-122   all_F.cuda() every round; per object assess_net(all_F, all_P[:, i + 1]) with B = T;
                                 float64 mean over objects; Agent.action
  models/assessment.py:164-182   (tp > 0.5) -> mask to the host -> numpy bbox loop (all2yxhw) -> affine_grid ->
                                 grid_sample x2 -> torchvision ResNet-50 body with a 4-channel stem -> mean -> fc1
  models/agent.py:33-64          Brain: per-frame Python loop over an nn.LSTMCell shared by both directions

Everything dense is stock PyTorch / cuDNN in fp32 eager mode, as the reference runs it (cudnn.deterministic = True).
This is the denominator of BASELINE.json's ">= 5x the reference's single-GPU PyTorch path" target; it is timed next to
this repo's path on the same clip and also cross-checks the result (same weights, TF32 off: same next frame)."""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "ivos-w_b200"))
from ivosw import synth  # noqa: E402
from ivosw.engine import Engine  # noqa: E402


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        r = torchvision.models.resnet50(weights=None)
        self.conv1 = r.conv1
        self.conv1_p = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
        self.conv1_m = nn.Conv2d(1, 64, 7, 2, 3, bias=True)      # unused members of the reference's state dict
        self.conv1_n = nn.Conv2d(1, 64, 7, 2, 3, bias=False)
        self.bn1, self.relu, self.maxpool = r.bn1, r.relu, r.maxpool
        self.res2, self.res3, self.res4, self.res5 = r.layer1, r.layer2, r.layer3, r.layer4
        self.register_buffer("mean", torch.zeros(1, 3, 1, 1))
        self.register_buffer("std", torch.ones(1, 3, 1, 1))

    def forward(self, f, p):
        x = self.conv1((f - self.mean) / self.std) + self.conv1_p(p.unsqueeze(1))
        x = self.maxpool(self.relu(self.bn1(x)))
        return self.res5(self.res4(self.res3(self.res2(x))))


def boxes_on_host(mask):
    """numpy bbox loop on the host copy of the mask, box grown to >= 128 px and scaled 1.5x (assessment.py:110-161)."""
    m = mask.cpu().numpy()
    B, H, W = m.shape
    out = np.zeros((B, 4), np.float32)
    for b in range(B):
        ys, xs = np.where(m[b] >= 0.49)
        if ys.size == 0:
            y0, y1, x0, x1 = 0, H, 0, W
        else:
            y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
        if y1 - y0 < 128:
            g = int((128. - (y1 - y0)) / 2); y0 -= g; y1 += g
        if x1 - x0 < 128:
            g = int((128. - (x1 - x0)) / 2); x0 -= g; x1 += g
        h, w = y1 - y0 + 1, x1 - x0 + 1
        fy0, fy1 = max(-5, y0 - 0.25 * h), min(H + 5, y1 + 0.25 * h)
        fx0, fx1 = max(-5, x0 - 0.25 * w), min(W + 5, x1 + 0.25 * w)
        out[b] = [(fy0 + fy1) / 2, (fx0 + fx1) / 2, fy1 - fy0 + 1, fx1 - fx0 + 1]
    return torch.from_numpy(out)


class AssessNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.Encoder = Encoder()
        self.fc1 = nn.Linear(2048, 1)

    def forward(self, tf, tp):
        tm = (tp > 0.5).float()
        roi = boxes_on_host(tm).to(tf.device)
        B, _, H, W = tf.shape
        ymin, ymax = roi[:, 0] - roi[:, 2] / 2, roi[:, 0] + roi[:, 2] / 2
        xmin, xmax = roi[:, 1] - roi[:, 3] / 2, roi[:, 1] + roi[:, 3] / 2
        theta = torch.zeros(B, 2, 3, device=tf.device)
        theta[:, 0, 0] = (xmax - xmin) / (W - 1)
        theta[:, 0, 2] = (xmin + xmax - (W - 1)) / (W - 1)
        theta[:, 1, 1] = (ymax - ymin) / (H - 1)
        theta[:, 1, 2] = (ymin + ymax - (H - 1)) / (H - 1)
        grid = F.affine_grid(theta, (B, 1, 256, 256), align_corners=True)
        f = F.grid_sample(tf, grid, align_corners=True)
        p = F.grid_sample(tp.unsqueeze(1), grid, align_corners=True)[:, 0]
        r5 = self.Encoder(f, p)
        return self.fc1(r5.mean(-1).mean(-1))


class Brain(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder_fc1, self.encoder_fc2 = nn.Linear(2, 128), nn.Linear(128, 128)
        self.lstm_cell = nn.LSTMCell(128, 128, bias=False)
        self.decoder_fc1, self.decoder_fc2 = nn.Linear(256, 128), nn.Linear(128, 1)

    def forward(self, x):
        N, T, _ = x.shape
        e = [self.encoder_fc2(F.relu(self.encoder_fc1(x[:, t]))) for t in range(T)]
        h = c = x.new_zeros(N, 128)
        fw = []
        for t in range(T):
            h, c = self.lstm_cell(e[t], (h, c)); fw.append(h)
        h = c = x.new_zeros(N, 128)
        bw = [None] * T
        for t in reversed(range(T)):
            h, c = self.lstm_cell(e[t], (h, c)); bw[t] = h
        q = [self.decoder_fc2(F.relu(self.decoder_fc1(F.relu(torch.cat([fw[t], bw[t]], 1))))) for t in range(T)]
        return torch.cat(q, 1)


def reference_round(assess, brain, all_F_cpu, all_P, annotated, T, O):
    with torch.no_grad():
        F_gpu = all_F_cpu.cuda()                                   # re-uploaded every round (utils_agent.py:114)
        pred = np.zeros((T, O))
        for i in range(O):
            pred[:, i] = assess(F_gpu, all_P[:, i + 1]).cpu().numpy()[:, 0]
        mq = pred.mean(1)
        ann = np.zeros(T)
        for i in annotated:
            ann[i] += 1
        state = torch.Tensor(np.stack([mq, ann], 1)[None]).cuda()
        q = brain(state).cpu().numpy().squeeze()
    return int(q.argmax()), mq


def main():
    T, H, W, O = 64, 480, 854, 2
    torch.backends.cudnn.deterministic = True
    all_F, all_P, annotated = synth.make_clip(0, T, H, W, O)
    assess, brain = AssessNet(), Brain()
    missing = assess.load_state_dict(synth.assess_state_dict(0), strict=True)
    brain.load_state_dict(synth.brain_state_dict(0), strict=True)
    assess.cuda().eval(); brain.cuda().eval()
    F_cpu = torch.from_numpy(all_F)                                 # pageable host memory, as in the reference
    P_gpu = torch.from_numpy(all_P).cuda()
    eng = Engine(0, "tc_fp16x3")
    eng.load_assess(synth.assess_state_dict(0)); eng.load_brain(synth.brain_state_dict(0))
    ann = synth.annotated_counts(annotated, T)
    F_gpu = torch.from_numpy(all_F).cuda()
    ours = eng.round_device(F_gpu, P_gpu, ann)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(2):
            nf, mq = reference_round(assess, brain, F_cpu, P_gpu, annotated, T, O)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            nf, mq = reference_round(assess, brain, F_cpu, P_gpu, annotated, T, O)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        t = sorted(ts)[len(ts) // 2]
        print("reference single-GPU PyTorch path (cuDNN fp32, allow_tf32=%s): %.1f ms per round = %.0f frames/s; next frame %d; "
              "max |mask_quality - ours| = %.2e" % (tf32, t * 1e3, T / t, nf, float(np.abs(mq - ours["mask_quality"]).max())))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        eng.round_device(F_gpu, P_gpu, ann)
    e0.record()
    for _ in range(10):
        r = eng.round_device(F_gpu, P_gpu, ann)
    e1.record(); torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 10
    Fp, Pp = torch.from_numpy(all_F).pin_memory(), torch.from_numpy(all_P).pin_memory()
    for _ in range(2):
        eng.round_host(Fp, Pp, ann)
    t0 = time.perf_counter()
    for _ in range(5):
        r = eng.round_host(Fp, Pp, ann)
    t_host = (time.perf_counter() - t0) / 5 * 1e3
    print("this repo: %.2f ms per round with device-resident inputs (%.0f frames/s), %.2f ms from host buffers (%.0f frames/s); "
          "next frame %d" % (t_dev, T / t_dev * 1e3, t_host, T / t_host * 1e3, r["next_frame"]))


if __name__ == "__main__":
    main()
