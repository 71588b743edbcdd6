"""ORACLE (test infrastructure, not product code) — CPU restatement of one optimisation step of the
quality-assessment network as ``quality_assessment.py::train`` performs it (lines 228-270), config C5 of
BASELINE.json / SURVEY.md §8(f) rank 3.  No CUDA path implements this step yet; the oracle and its goldens pin
the semantics the step has to reproduce:

  :213      assess_net.train()              BatchNorm uses batch statistics and updates its running statistics
                                            (momentum 0.1, unbiased variance, num_batches_tracked += 1)
  :240      iou_pred = assess_net(imgs, probs)          same forward as inference (bbox on the binarised mask,
                                                        ROI grid, 2 x grid_sample, encoder, mean, fc1)
  :251-262  loss = mean over the samples with union > 0 of mse_loss(iou_pred[n], metric_gt[n])
  :265      loss.backward()                 NOTE: the loop never calls optimizer.zero_grad(): gradients ACCUMULATE
                                            over iterations — reproduced here
  :266-268  every parameter gradient clamped to [-1, 1] in place (the clamped value is what keeps accumulating)
  :269      optimizer.step()                torch.optim.SGD(lr, momentum, weight_decay) (:309-310): parameters
                                            without a gradient (conv1_m, conv1_n: registered but unused) are skipped

Parity status: PINNED — tests/test_oracle_golden.py::test_assess_train_step compares two consecutive steps against
tests/golden/assess_train.npz, produced by the reference's own AssessNet + torch.optim.SGD
(tests/golden/make_golden.py::golden_assess_train).

Only tests/ may import this module.
"""
import torch
import torch.nn.functional as F

from . import assess_ref

BN_MOMENTUM = 0.1


def _bn_train(x, st, p):
    """F.batch_norm in training mode over the state's tensors: batch statistics for the output, running statistics
    updated in place with momentum 0.1 and the unbiased batch variance (torch.nn.BatchNorm2d defaults)."""
    st.buffers[p + ".num_batches_tracked"] += 1
    return F.batch_norm(x, st.buffers[p + ".running_mean"], st.buffers[p + ".running_var"],
                        st.params[p + ".weight"], st.params[p + ".bias"], True, BN_MOMENTUM, assess_ref.BN_EPS)


class TrainState:
    """Parameters (leaf tensors with requires_grad), BatchNorm buffers, SGD momentum buffers."""

    def __init__(self, state_dict, dtype=torch.float32):
        self.params, self.buffers, self.momentum = {}, {}, {}
        for k, v in state_dict.items():
            is_buffer = k.endswith(("running_mean", "running_var", "num_batches_tracked")) or k in ("Encoder.mean", "Encoder.std")
            t = v.detach().clone()
            if t.is_floating_point():
                t = t.to(dtype)
            if is_buffer:
                self.buffers[k] = t
            else:
                self.params[k] = t.requires_grad_(True)


def forward_train(st, tf, tp, dtype=torch.float32):
    """AssessNet.forward in train mode (models/assessment.py:164-182 with :213).  tf: B x 3 x H x W, tp: B x H x W.
    Returns iou_pred: B x 1."""
    tf = torch.as_tensor(tf).to(dtype)
    tp32 = torch.as_tensor(tp)
    B, _, H, W = tf.shape
    tm = (tp32 > 0.5).float()
    tb = torch.from_numpy(assess_ref.all2yxhw(tm.numpy(), scale=1.5))
    gx, gy = assess_ref.roi_grid(tb, (H, W), 256, dtype)
    in_f = assess_ref.grid_sample_bilinear(tf, gx, gy)
    in_p = assess_ref.grid_sample_bilinear(tp32.to(dtype)[:, None], gx, gy)[:, 0]
    P = st.params
    f = (in_f - st.buffers["Encoder.mean"]) / st.buffers["Encoder.std"]
    x = F.conv2d(f, P["Encoder.conv1.weight"], None, 2, 3) + F.conv2d(in_p.unsqueeze(1), P["Encoder.conv1_p.weight"], None, 2, 3)
    x = F.max_pool2d(F.relu(_bn_train(x, st, "Encoder.bn1")), 3, 2, 1)
    for stage, planes, blocks, stride in assess_ref._STAGES:
        for b in range(blocks):
            pre = "Encoder.%s.%d." % (stage, b)
            s = stride if b == 0 else 1
            idt = x
            o = F.relu(_bn_train(F.conv2d(x, P[pre + "conv1.weight"]), st, pre + "bn1"))
            o = F.relu(_bn_train(F.conv2d(o, P[pre + "conv2.weight"], None, s, 1), st, pre + "bn2"))
            o = _bn_train(F.conv2d(o, P[pre + "conv3.weight"]), st, pre + "bn3")
            if b == 0:
                idt = _bn_train(F.conv2d(x, P[pre + "downsample.0.weight"], None, s), st, pre + "downsample.1")
            x = F.relu(o + idt)
    flat = F.avg_pool2d(x, 8).reshape(B, -1)
    return F.linear(flat, P["fc1.weight"], P["fc1.bias"])


def train_step(st, tf, tp, targets, valid, lr, momentum, weight_decay, dtype=torch.float32):
    """One iteration of the loop body (:240-269).  ``valid[n]`` stands for ``union[n] > 0``.  Gradients are NOT
    zeroed between calls (the reference never does).  Returns (loss, iou_pred) or (None, iou_pred) when no sample
    is valid (the reference ``continue``s before backward)."""
    pred = forward_train(st, tf, tp, dtype)
    targets = torch.as_tensor(targets).to(dtype)
    loss, counter = 0.0, 0
    for n in range(pred.shape[0]):
        if valid[n]:
            loss = loss + ((pred[n] - targets[n]) ** 2).mean()      # F.mse_loss(iou_pred[n], metric_gt[n])
            counter += 1
    if counter == 0:
        return None, pred.detach()
    loss = loss / counter
    loss.backward()                                                  # accumulates into .grad
    with torch.no_grad():
        for k, p in st.params.items():
            if p.grad is None:
                continue
            p.grad.clamp_(-1, 1)
            d = p.grad + weight_decay * p                            # SGD: L2 term enters the step, not .grad
            buf = st.momentum.get(k)
            buf = d.clone() if buf is None else buf.mul_(momentum).add_(d)
            st.momentum[k] = buf
            p.sub_(lr * buf)
    return loss.detach(), pred.detach()
