"""ORACLE (test infrastructure, not product code) — CPU restatement of the
reference quality-assessment network ``models/assessment.py::AssessNet``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product path never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks every stage of this
restatement (boxes, ROI crops, r2..r5 probes, scores) against
tests/golden/*.npz, produced by the reference's own ``AssessNet`` imported from
/root/reference (tests/golden/make_golden.py; the only harness-side change is
the ``resnet50(pretrained=True)`` -> ``weights=None`` shim).

The bbox arithmetic is restated in numpy exactly as the reference does it
(int64 -> float64 -> float32); the dense arithmetic uses torch *functional* ops
on CPU with the weights taken from a plain state-dict, so the restatement has
no dependency on torchvision or on the reference's module classes.  ``dtype``
torch.float32 is the oracle, torch.float64 the arbiter.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
_STAGES = (("res2", 64, 3, 1), ("res3", 128, 4, 2), ("res4", 256, 6, 2), ("res5", 512, 3, 2))


# ---------------------------------------------------------------------------
# a3: all2yxhw  (models/assessment.py:110-161)
# ---------------------------------------------------------------------------
def all2yxhw(mask, scale=1.5):
    """mask: B x H x W array of 0/1 floats (tm = (tp > 0.5).float(), line 165).
    Returns B x 4 float32 [yc, xc, h, w]."""
    np_mask = np.asarray(mask)
    B, H, W = np_mask.shape
    out = np.zeros((B, 4), dtype=np.float32)
    for b in range(B):
        ys, xs = np.where(np_mask[b] >= 0.49)                       # :115
        if ys.size == 0 or xs.size == 0:                            # :119-122
            ymin, ymax = 0, H
            xmin, xmax = 0, W
        else:                                                       # :124-125
            ymin, ymax = np.min(ys), np.max(ys)
            xmin, xmax = np.min(xs), np.max(xs)
        if (ymax - ymin) < 128:                                     # :128-131
            res = 128. - (ymax - ymin)
            ymin -= int(res / 2)
            ymax += int(res / 2)
        if (xmax - xmin) < 128:                                     # :133-136
            res = 128. - (xmax - xmin)
            xmin -= int(res / 2)
            xmax += int(res / 2)
        orig_h = ymax - ymin + 1                                    # :141-142
        orig_w = xmax - xmin + 1
        ymin = np.maximum(-5, ymin - (scale - 1) / 2. * orig_h)     # :144-149
        ymax = np.minimum(H + 5, ymax + (scale - 1) / 2. * orig_h)
        xmin = np.maximum(-5, xmin - (scale - 1) / 2. * orig_w)
        xmax = np.minimum(W + 5, xmax + (scale - 1) / 2. * orig_w)
        y = (ymax + ymin) / 2.                                      # :152-155
        x = (xmax + xmin) / 2.
        h = ymax - ymin + 1
        w = xmax - xmin + 1
        out[b] = np.array([y, x, h, w], dtype=np.float32)           # :157
    return out


# ---------------------------------------------------------------------------
# a4: get_ROI_grid (forward grid only; the inverse grid is never used, :169)
# ---------------------------------------------------------------------------
def roi_theta(roi, src_size, dtype=torch.float32):
    """roi: B x 4 float32 tensor.  Returns (t00, t02, t11, t12), each B
    (models/assessment.py:77-93; theta[0,1] = theta[1,0] = 0)."""
    roi = torch.as_tensor(roi).to(dtype)
    ry, rx, rh, rw = roi[:, 0], roi[:, 1], 1.0 * roi[:, 2], 1.0 * roi[:, 3]
    ymin = ry - rh / 2.
    ymax = ry + rh / 2.
    xmin = rx - rw / 2.
    xmax = rx + rw / 2.
    h, w = src_size
    t00 = (xmax - xmin) / (w - 1)
    t02 = (xmin + xmax - (w - 1)) / (w - 1)
    t11 = (ymax - ymin) / (h - 1)
    t12 = (ymin + ymax - (h - 1)) / (h - 1)
    return t00, t02, t11, t12


def _linspace_m1_p1(n, dtype):
    """at::linspace(-1, 1, n) exactly as torch's CPU kernel evaluates it: the step
    is rounded to the tensor dtype, each half is one fused multiply-add from its
    own end point (verified bit-exact against torch.linspace for fp32)."""
    idx = torch.arange(n, dtype=torch.float64)
    if dtype == torch.float32:
        step = float(np.float32(2.0) / np.float32(n - 1))
    else:
        step = 2.0 / (n - 1)
    lo = -1.0 + step * idx
    hi = 1.0 - step * (n - 1 - idx)
    return torch.where(torch.arange(n) < n // 2, lo, hi).to(dtype)


def roi_grid(roi, src_size, dst=256, dtype=torch.float32):
    """F.affine_grid(theta, (B,1,dst,dst), align_corners=True) restated:
    grid[b,i,j] = (t00*x_j + t02, t11*y_i + t12).  Returns gx, gy: B x dst x dst."""
    t00, t02, t11, t12 = roi_theta(roi, src_size, dtype)
    lin = _linspace_m1_p1(dst, dtype)
    # bmm over K=3 with theta01 = 0: one rounded multiply then one rounded add (no FMA) —
    # bit-exact against F.affine_grid on CPU (tests/test_oracle_golden.py)
    gx = t00[:, None] * lin[None, :] + t02[:, None]            # B x dst (depends on j only)
    gy = t11[:, None] * lin[None, :] + t12[:, None]            # B x dst (depends on i only)
    B = gx.shape[0]
    return gx[:, None, :].expand(B, dst, dst), gy[:, :, None].expand(B, dst, dst)


# ---------------------------------------------------------------------------
# a5: F.grid_sample(bilinear, zeros padding, align_corners=True)
# ---------------------------------------------------------------------------
def grid_sample_bilinear(img, gx, gy):
    """img: B x C x H x W; gx, gy: B x h x w normalised coords.  -> B x C x h x w."""
    B, C, H, W = img.shape
    ix = ((gx + 1) / 2) * (W - 1)
    iy = ((gy + 1) / 2) * (H - 1)
    # weights as ATen's vectorised CPU kernel forms them (GridSamplerKernel.cpp, bilinear):
    # w = x - floor(x); e = 1 - w; n = y - floor(y); s = 1 - n; nw = s*e, ne = s*w, sw = n*e, se = n*w
    x0 = torch.floor(ix); y0 = torch.floor(iy)
    x1 = x0 + 1; y1 = y0 + 1
    w = ix - x0; e = 1 - w
    n = iy - y0; s_ = 1 - n
    w_nw = s_ * e
    w_ne = s_ * w
    w_sw = n * e
    w_se = n * w
    flat = img.reshape(B, C, H * W)
    out = torch.zeros((B, C) + tuple(gx.shape[1:]), dtype=img.dtype)

    def corner(xc, yc, wgt):
        ok = (xc >= 0) & (xc <= W - 1) & (yc >= 0) & (yc <= H - 1)
        idx = (yc.clamp(0, H - 1) * W + xc.clamp(0, W - 1)).long().reshape(B, 1, -1).expand(B, C, -1)
        v = torch.gather(flat, 2, idx).reshape(out.shape)
        return v * (wgt * ok.to(img.dtype))[:, None]

    out = corner(x0, y0, w_nw) + corner(x1, y0, w_ne) + corner(x0, y1, w_sw) + corner(x1, y1, w_se)
    return out


# ---------------------------------------------------------------------------
# a6: Encoder.forward (models/assessment.py:46-63) over a plain state-dict
# ---------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def encoder_forward(sd, in_f, in_p, probes=None):
    """in_f: B x 3 x 256 x 256 (ROI crop, un-normalised), in_p: B x 256 x 256.
    Returns r5 (B x 2048 x 8 x 8).  If ``probes`` is a dict it receives the
    intermediate tensors c1, x (after maxpool), r2, r3, r4, r5."""
    f = (in_f - sd["Encoder.mean"]) / sd["Encoder.std"]                       # :47
    p = in_p.unsqueeze(1)                                                     # :48
    x = F.conv2d(f, sd["Encoder.conv1.weight"], None, 2, 3) + \
        F.conv2d(p, sd["Encoder.conv1_p.weight"], None, 2, 3)                 # :54
    c1 = F.relu(_bn(x, sd, "Encoder.bn1"))                                    # :55-56
    x = F.max_pool2d(c1, 3, 2, 1)                                             # :57
    if probes is not None:
        probes["c1"] = c1; probes["pool"] = x
    for stage, planes, blocks, stride in _STAGES:                             # :58-61
        for b in range(blocks):
            pre = "Encoder.%s.%d." % (stage, b)
            s = stride if b == 0 else 1
            idt = x
            o = F.relu(_bn(F.conv2d(x, sd[pre + "conv1.weight"]), sd, pre + "bn1"))
            o = F.relu(_bn(F.conv2d(o, sd[pre + "conv2.weight"], None, s, 1), sd, pre + "bn2"))
            o = _bn(F.conv2d(o, sd[pre + "conv3.weight"]), sd, pre + "bn3")
            if b == 0:
                idt = _bn(F.conv2d(x, sd[pre + "downsample.0.weight"], None, s), sd, pre + "downsample.1")
            x = F.relu(o + idt)
        if probes is not None:
            probes[{"res2": "r2", "res3": "r3", "res4": "r4", "res5": "r5"}[stage]] = x
    return x


# ---------------------------------------------------------------------------
# a2: AssessNet.forward (models/assessment.py:164-182)
# ---------------------------------------------------------------------------
def assess_forward(sd, tf, tp, dtype=torch.float32, probes=None, chunk=16):
    """sd: state-dict (reference key names) of torch tensors.
    tf: B x 3 x H x W, tp: B x H x W (fp32).  Returns scores B (the reference
    returns B x 1, or shape (1,) when B == 1 — a shape quirk the drop-in shim
    reproduces; the oracle returns the flat values)."""
    sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    tf = torch.as_tensor(tf).to(dtype)
    tp32 = torch.as_tensor(tp)
    B, _, H, W = tf.shape
    tm = (tp32 > 0.5).float()                                                 # :165
    tb = torch.from_numpy(all2yxhw(tm.numpy(), scale=1.5))                    # :166
    if probes is not None:
        probes["boxes"] = tb.clone()
    tp = tp32.to(dtype)
    out = []
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        gx, gy = roi_grid(tb[s:e], (H, W), 256, dtype)                        # :169-170
        tf_roi = grid_sample_bilinear(tf[s:e], gx, gy)                        # :173
        tp_roi = grid_sample_bilinear(tp[s:e, None], gx, gy)[:, 0]            # :174
        pr = {} if probes is not None else None
        r5 = encoder_forward(sd, tf_roi, tp_roi, pr)                          # :177
        flat = F.avg_pool2d(r5, 8).reshape(e - s, -1)                         # :179
        out.append(F.linear(flat, sd["fc1.weight"], sd["fc1.bias"])[:, 0])    # :180
        if probes is not None:
            pr["tf_roi"] = tf_roi; pr["tp_roi"] = tp_roi
            for k, v in pr.items():
                probes.setdefault(k, []).append(v)
    if probes is not None:
        for k in list(probes):
            if isinstance(probes[k], list):
                probes[k] = torch.cat(probes[k], 0)
    return torch.cat(out, 0)
