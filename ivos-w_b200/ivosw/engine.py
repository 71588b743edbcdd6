"""Engine: a thin, torch-aware wrapper around one ``ivosw_ctx`` (one per GPU).

PyTorch is used here only for device memory, streams and (in ``dist.py``)
``torch.distributed``; all arithmetic happens in the CUDA library.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import arch
from ._lib import (BRAIN_NUM_PARAMS, CONV_SIMT_FP32, CONV_TC_FP16X1, CONV_TC_FP16X3, check, lib)

CONV_MODES = {"simt_fp32": CONV_SIMT_FP32, "tc_fp16x3": CONV_TC_FP16X3, "tc_fp16x1": CONV_TC_FP16X1}
# arithmetic of the encoder unless a caller says otherwise (README.md, "Environment switches")
DEFAULT_CONV_MODE = os.environ.get("IVOSW_CONV_MODE", "tc_fp16x3")


def pack_brain(sd):
    """Brain.state_dict() -> flat fp32 array in the order include/ivosw_b200.h documents."""
    parts = []
    for key, shape in arch.BRAIN_PARAMS:
        t = sd[key].detach().to("cpu", torch.float32).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("Brain parameter %s has shape %s, expected %s" % (key, tuple(t.shape), shape))
        parts.append(t.reshape(-1).numpy())
    out = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
    assert out.size == BRAIN_NUM_PARAMS
    return out


def pack_assess(sd):
    """AssessNet.state_dict() -> flat fp32 blob (see ivosw_assess_load).  Weights go OIHW -> OHWI."""
    def f(key):
        return sd[key].detach().to("cpu", torch.float32)

    parts = [f("Encoder.mean").reshape(-1), f("Encoder.std").reshape(-1)]
    stem = torch.cat([f("Encoder.conv1.weight"), f("Encoder.conv1_p.weight")], 1)       # 64 x 4 x 7 x 7
    parts.append(stem.permute(0, 2, 3, 1).reshape(-1))
    for s in ("weight", "bias", "running_mean", "running_var"):
        parts.append(f("Encoder.bn1." + s))
    for c in arch.resnet50_convs():
        w = f(c.name + ".weight")
        if tuple(w.shape) != (c.cout, c.cin, c.k, c.k):
            raise ValueError("%s.weight has shape %s" % (c.name, tuple(w.shape)))
        parts.append(w.permute(0, 2, 3, 1).reshape(-1))
        for s in ("weight", "bias", "running_mean", "running_var"):
            parts.append(f(c.bn + "." + s))
    parts += [f("fc1.weight").reshape(-1), f("fc1.bias").reshape(-1)]
    out = np.ascontiguousarray(torch.cat([p.contiguous().reshape(-1) for p in parts]).numpy(), dtype=np.float32)
    if out.size != lib.ivosw_assess_blob_floats():
        raise ValueError("AssessNet blob has %d floats, library expects %d" % (out.size, lib.ivosw_assess_blob_floats()))
    return out


def pack_manet_encoder(sd):
    """State dict of the MANet feature-extractor restatement (ivosw/manet_arch.py key names) -> flat fp32 blob
    (ivosw_manet_encoder_load): per convolution OHWI weight, then gamma, beta, running_mean - conv bias, running_var."""
    from . import manet_arch
    parts = []
    for c in manet_arch.convs():
        w = sd[c.name + ".weight"].detach().to("cpu", torch.float32)
        if tuple(w.shape) != (c.cout, c.cin // c.groups, c.k, c.k):
            raise ValueError("%s.weight has shape %s" % (c.name, tuple(w.shape)))
        parts.append(w.permute(0, 2, 3, 1).reshape(-1))
        mean = sd[c.bn + ".running_mean"].detach().to("cpu", torch.float32)
        if c.bias:
            mean = mean - sd[c.name + ".bias"].detach().to("cpu", torch.float32)
        parts += [sd[c.bn + ".weight"].detach().float().cpu(), sd[c.bn + ".bias"].detach().float().cpu(), mean,
                  sd[c.bn + ".running_var"].detach().float().cpu()]
    out = np.ascontiguousarray(torch.cat([p.contiguous().reshape(-1) for p in parts]).numpy(), dtype=np.float32)
    if out.size != lib.ivosw_manet_encoder_blob_floats():
        raise ValueError("MANet encoder blob has %d floats, library expects %d" % (out.size, lib.ivosw_manet_encoder_blob_floats()))
    return out


def unpack_assess(blob, template, grads=False):
    """Inverse of pack_assess: flat blob -> dict with the reference's key names (OHWI -> OIHW).  Keys the blob does not
    carry (conv1_m / conv1_n, num_batches_tracked) are taken from ``template`` (gradients: omitted)."""
    out = {} if grads else {k: v.clone() for k, v in template.items()}
    o = 6
    if not grads:
        out["Encoder.mean"] = torch.from_numpy(blob[0:3].copy()).view(1, 3, 1, 1)
        out["Encoder.std"] = torch.from_numpy(blob[3:6].copy()).view(1, 3, 1, 1)
    stem = torch.from_numpy(blob[o:o + 64 * 196].copy()).view(64, 7, 7, 4).permute(0, 3, 1, 2).contiguous()
    out["Encoder.conv1.weight"], out["Encoder.conv1_p.weight"] = stem[:, :3].contiguous(), stem[:, 3:].contiguous()
    o += 64 * 196
    names = ("weight", "bias") if grads else ("weight", "bias", "running_mean", "running_var")
    for j, s_ in enumerate(("weight", "bias", "running_mean", "running_var")):
        if s_ in names:
            out["Encoder.bn1." + s_] = torch.from_numpy(blob[o + 64 * j:o + 64 * (j + 1)].copy())
    o += 256
    for c in arch.resnet50_convs():
        n = c.cout * c.k * c.k * c.cin
        out[c.name + ".weight"] = torch.from_numpy(blob[o:o + n].copy()).view(c.cout, c.k, c.k, c.cin).permute(0, 3, 1, 2).contiguous()
        o += n
        for j, s_ in enumerate(("weight", "bias", "running_mean", "running_var")):
            if s_ in names:
                out[c.bn + "." + s_] = torch.from_numpy(blob[o + c.cout * j:o + c.cout * (j + 1)].copy())
        o += 4 * c.cout
    out["fc1.weight"] = torch.from_numpy(blob[o:o + 2048].copy()).view(1, 2048)
    out["fc1.bias"] = torch.from_numpy(blob[o + 2048:o + 2049].copy())
    return out


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """One CUDA context of the scoring library on ``device`` (int or torch.device)."""

    def __init__(self, device=0, conv_mode=DEFAULT_CONV_MODE):
        if not torch.cuda.is_available():
            raise RuntimeError("ivosw_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else "cuda:%d" % device)
        if dev.type != "cuda":
            raise RuntimeError("ivosw_b200 runs on CUDA devices only, got %r" % (device,))
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        torch.cuda.init()
        torch.zeros(1, device=self.device)     # make sure torch's primary context exists (shared with the library)
        h = C.c_void_p()
        check(lib.ivosw_create(self.device.index, CONV_MODES[conv_mode], C.byref(h)))
        self._h = h
        self.conv_mode = conv_mode

    def close(self):
        if getattr(self, "_h", None):
            lib.ivosw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ configuration
    def set_conv_mode(self, mode):
        check(lib.ivosw_set_conv_mode(self._h, CONV_MODES[mode]))
        self.conv_mode = mode

    @property
    def launch_count(self):
        return int(lib.ivosw_launch_count(self._h))

    def load_brain(self, sd):
        blob = pack_brain(sd)
        check(lib.ivosw_brain_load(self._h, _np_ptr(blob), blob.size))

    def load_assess(self, sd):
        blob = pack_assess(sd)
        check(lib.ivosw_assess_load(self._h, _np_ptr(blob), blob.size))

    # ------------------------------------------------------------------ Brain (models/agent.py:33)
    def brain_forward(self, state, want_argmax=False):
        """state: N x T x 2 CUDA fp32 tensor -> Q (N x T) [, argmax (N) int32]."""
        state = self._dev32(state)
        N, T, P = state.shape
        if P != 2:
            raise ValueError("Brain input must be N x T x 2")
        q = torch.empty((N, T), device=self.device, dtype=torch.float32)
        am = torch.empty((N,), device=self.device, dtype=torch.int32) if want_argmax else None
        check(lib.ivosw_brain_forward(self._h, _ptr(state), N, T, _ptr(q), _ptr(am), _stream(self.device)))
        return (q, am) if want_argmax else q

    # ------------------------------------------------------------------ DQN step (models/agent.py:103-166)
    def load_target(self, sd):
        blob = pack_brain(sd)
        check(lib.ivosw_dqn_load_target(self._h, _np_ptr(blob), blob.size))

    def sync_target(self):
        check(lib.ivosw_dqn_sync_target(self._h, _stream(self.device)))

    def reset_optimizer(self):
        check(lib.ivosw_dqn_reset_optimizer(self._h))

    def optimizer_state(self):
        """Adam state of the policy network inside the library: dict(step, exp_avg, exp_avg_sq) — the two
        moment vectors as flat CUDA tensors in blob order (Engine.unpack_brain gives them per parameter)."""
        m = torch.empty((BRAIN_NUM_PARAMS,), device=self.device, dtype=torch.float32)
        v = torch.empty_like(m)
        step = C.c_longlong(0)
        check(lib.ivosw_dqn_get_optimizer(self._h, _ptr(m), _ptr(v), C.byref(step), _stream(self.device)))
        return {"step": int(step.value), "exp_avg": m, "exp_avg_sq": v}

    def load_optimizer_state(self, state):
        m = state["exp_avg"].to(self.device, torch.float32).contiguous().view(-1)
        v = state["exp_avg_sq"].to(self.device, torch.float32).contiguous().view(-1)
        assert m.numel() == BRAIN_NUM_PARAMS and v.numel() == BRAIN_NUM_PARAMS
        check(lib.ivosw_dqn_set_optimizer(self._h, _ptr(m), _ptr(v), int(state["step"]), _stream(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    def saturation_count(self, reset=True):
        """(thread, tile) pairs of the split-fp16 encoder that clamped a value to +-65504 since the last reset."""
        n = C.c_longlong(0)
        check(lib.ivosw_conv_saturation_count(self._h, C.byref(n), 1 if reset else 0, _stream(self.device)))
        return int(n.value)

    def dqn_update(self, state, new_state, action, reward_step, reward_done, gamma=0.95, lr=5e-6, weight_decay=5e-4,
                   want_grads=False, apply=True):
        """One Agent.update_agent step on CUDA tensors: state/new_state N x T x 2, action N, rewards N.
        The policy parameters inside the library are updated in place.  Returns loss [, clamped grads]."""
        state, new_state = self._dev32(state), self._dev32(new_state)
        N, T, _ = state.shape
        action = action.to(self.device, torch.int32).contiguous().view(-1)
        rs = self._dev32(reward_step).view(-1)
        rd = self._dev32(reward_done).view(-1)
        want_grads = want_grads or not apply
        grads = torch.empty((BRAIN_NUM_PARAMS,), device=self.device, dtype=torch.float32) if want_grads else None
        loss = C.c_float(0.0)
        check(lib.ivosw_dqn_update(self._h, _ptr(state), _ptr(new_state), _ptr(action), _ptr(rs), _ptr(rd), N, T,
                                   gamma, lr, weight_decay, C.byref(loss), _ptr(grads), 1 if apply else 0,
                                   _stream(self.device)))
        return (float(loss.value), grads) if want_grads else float(loss.value)

    def dqn_apply(self, grads, lr=5e-6, weight_decay=5e-4):
        """Clamp + Adam step on (all-reduced) raw gradients; ``grads`` is overwritten with the clamped values."""
        assert grads.is_cuda and grads.dtype == torch.float32 and grads.numel() == BRAIN_NUM_PARAMS
        check(lib.ivosw_dqn_apply(self._h, _ptr(grads), lr, weight_decay, _stream(self.device)))

    def brain_params(self, which="policy"):
        """Flat parameter blob (blob order) of the policy / target network currently inside the library."""
        out = torch.empty((BRAIN_NUM_PARAMS,), device=self.device, dtype=torch.float32)
        check(lib.ivosw_brain_get_params(self._h, 0 if which == "policy" else 1, _ptr(out), _stream(self.device)))
        return out

    @staticmethod
    def unpack_brain(flat):
        """flat blob -> dict of tensors with the reference's state-dict keys."""
        sd, o = {}, 0
        for key, shape in arch.BRAIN_PARAMS:
            n = int(np.prod(shape))
            sd[key] = flat[o:o + n].view(shape)
            o += n
        return sd

    # ------------------------------------------------------------------ AssessNet (models/assessment.py:164)
    def assess_forward(self, tf, tp, want_boxes=False):
        """tf: B x 3 x H x W, tp: B x H x W (CUDA fp32; tp may be a strided slice such as
        all_P[:, i+1]).  Returns scores (B,) [, boxes (B, 4)]."""
        tf = self._dev32(tf)
        B, Cc, H, W = tf.shape
        if Cc != 3:
            raise ValueError("tf must be B x 3 x H x W")
        if tp.device != self.device or tp.dtype != torch.float32:
            tp = tp.to(self.device, torch.float32)
        if tuple(tp.shape) != (B, H, W):
            raise ValueError("tp must be B x H x W matching tf")
        if tp.stride(2) != 1 or tp.stride(1) != W:
            tp = tp.contiguous()
        pstride = tp.stride(0) if B > 1 else H * W
        scores = torch.empty((B,), device=self.device, dtype=torch.float32)
        boxes = torch.empty((B, 4), device=self.device, dtype=torch.float32) if want_boxes else None
        check(lib.ivosw_assess_forward(self._h, _ptr(tf), 3 * H * W, _ptr(tp), pstride, B, H, W, _ptr(scores),
                                       _ptr(boxes), _stream(self.device)))
        return (scores, boxes) if want_boxes else scores

    def enable_probes(self, on=True):
        check(lib.ivosw_enable_probes(self._h, 1 if on else 0))

    def probe(self, which):
        """0 crop, 1 pool, 2..5 r2..r5 of the last assess chunk, as an NCHW fp32 CUDA tensor."""
        Cn = (4, 64, 256, 512, 1024, 2048)[which]
        hw = (256, 64, 64, 32, 16, 8)[which]
        cap = 128 * Cn * hw * hw
        dims = (C.c_int * 4)()
        buf = torch.empty((cap,), device=self.device, dtype=torch.float32)
        check(lib.ivosw_assess_probe(self._h, which, _ptr(buf), cap, dims, _stream(self.device)))
        n, c, h, w = [int(v) for v in dims]
        return buf[: n * c * h * w].view(n, c, h, w)

    # ------------------------------------------------------------------ round (utils/utils_agent.py:111-122)
    def round_device(self, all_F, all_P, annotated_counts, t_begin=0, t_end=None, want_scores=False,
                     want_action=True):
        """all_F: T x 3 x H x W, all_P: T x (O+1) x H x W CUDA fp32 tensors.
        Returns dict(mask_quality float64[t_end-t_begin], scores fp32[.., O] or None, q fp32[T] or None,
        next_frame int or None).  q / next_frame only for the full range."""
        all_F = self._dev32(all_F)
        all_P = self._dev32(all_P)
        T, _, H, W = all_F.shape
        O = all_P.shape[1] - 1
        t_end = T if t_end is None else t_end
        full = t_begin == 0 and t_end == T and want_action
        ann = np.ascontiguousarray(annotated_counts, dtype=np.float64)
        mq = np.empty(t_end - t_begin, dtype=np.float64)
        sc = np.empty((t_end - t_begin, O), dtype=np.float32) if want_scores else None
        q = np.empty(T, dtype=np.float32) if full else None
        nf = C.c_int(-1)
        check(lib.ivosw_round_device(self._h, _ptr(all_F), _ptr(all_P), T, O, H, W, t_begin, t_end, _np_ptr(ann),
                                     _np_ptr(mq), _np_ptr(sc), _np_ptr(q), C.byref(nf) if full else None,
                                     _stream(self.device)))
        return {"mask_quality": mq, "scores": sc, "q": q, "next_frame": int(nf.value) if full else None}

    def last_h2d_bytes(self):
        """Bytes the last host-buffer call actually sent to the device."""
        return int(lib.ivosw_last_h2d_bytes(self._h))

    def round_host(self, all_F, all_P, annotated_counts, want_scores=False):
        """Same round from host tensors (pinned or pageable CPU fp32, contiguous)."""
        if all_F.device.type != "cpu" or all_P.device.type != "cpu":
            raise ValueError("round_host takes CPU tensors")
        all_F = all_F.contiguous().float()
        all_P = all_P.contiguous().float()
        T, _, H, W = all_F.shape
        O = all_P.shape[1] - 1
        ann = np.ascontiguousarray(annotated_counts, dtype=np.float64)
        mq = np.empty(T, dtype=np.float64)
        sc = np.empty((T, O), dtype=np.float32) if want_scores else None
        q = np.empty(T, dtype=np.float32)
        nf = C.c_int(-1)
        check(lib.ivosw_round_host(self._h, _ptr(all_F), _ptr(all_P), T, O, H, W, _np_ptr(ann), _np_ptr(mq),
                                   _np_ptr(sc), _np_ptr(q), C.byref(nf), _stream(self.device)))
        return {"mask_quality": mq, "scores": sc, "q": q, "next_frame": int(nf.value)}

    def agent_action(self, mask_quality, annotated_counts):
        """Greedy Agent.action on host vectors (float64): returns (next_frame, q[T])."""
        mqa = np.ascontiguousarray(mask_quality, dtype=np.float64)
        ann = np.ascontiguousarray(annotated_counts, dtype=np.float64)
        T = mqa.shape[0]
        q = np.empty(T, dtype=np.float32)
        nf = C.c_int(-1)
        check(lib.ivosw_agent_action(self._h, _np_ptr(mqa), _np_ptr(ann), T, _np_ptr(q), C.byref(nf),
                                     _stream(self.device)))
        return int(nf.value), q

    # ------------------------------------------------------------------ frame-sharded round
    def score_shard(self, all_F, all_P, t_begin, t_end, mq_out, scores_out=None):
        """Asynchronous scoring of frames [t_begin, t_end): writes float64 per-frame quality into the
        CUDA tensor ``mq_out`` (length t_end - t_begin; typically a slice of the all-gather buffer)."""
        T, _, H, W = all_F.shape
        O = all_P.shape[1] - 1
        assert mq_out.dtype == torch.float64 and mq_out.is_cuda and mq_out.is_contiguous()
        assert mq_out.numel() == t_end - t_begin
        check(lib.ivosw_score_shard(self._h, _ptr(all_F), _ptr(all_P), T, O, H, W, t_begin, t_end, _ptr(mq_out),
                                    _ptr(scores_out), _stream(self.device)))

    def score_shard_host(self, all_F_host, all_P_host, t_begin, t_end, mq_out):
        """As score_shard with the clip in host memory (CPU fp32 contiguous tensors, ideally pinned)."""
        T, _, H, W = all_F_host.shape
        O = all_P_host.shape[1] - 1
        assert all_F_host.device.type == "cpu" and all_F_host.is_contiguous() and all_F_host.dtype == torch.float32
        assert all_P_host.device.type == "cpu" and all_P_host.is_contiguous() and all_P_host.dtype == torch.float32
        assert mq_out.dtype == torch.float64 and mq_out.is_cuda and mq_out.numel() == t_end - t_begin
        check(lib.ivosw_score_shard_host(self._h, _ptr(all_F_host), _ptr(all_P_host), T, O, H, W, t_begin, t_end,
                                         _ptr(mq_out), _stream(self.device)))

    def agent_action_dev(self, mq_dev, annotated_counts):
        """Brain + argmax on a device float64 quality vector (after the all-gather)."""
        T = mq_dev.numel()
        ann = np.ascontiguousarray(annotated_counts, dtype=np.float64)
        q = np.empty(T, dtype=np.float32)
        nf = C.c_int(-1)
        check(lib.ivosw_agent_action_dev(self._h, _ptr(mq_dev), _np_ptr(ann), T, _np_ptr(q), C.byref(nf),
                                         _stream(self.device)))
        return int(nf.value), q

    # ------------------------------------------------------------------ MANet feature extractor (csrc/manet_encoder.cu)
    def load_manet_encoder(self, sd):
        blob = pack_manet_encoder(sd)
        check(lib.ivosw_manet_encoder_load(self._h, _np_ptr(blob), blob.size))

    def manet_extract_feature(self, frames):
        """frames: B x 3 x H x W CUDA fp32 (normalised) -> B x 100 x h/4 x w/4 embedding (restatement, see manet_arch)."""
        from . import manet_arch
        x = self._dev32(frames)
        B, _, H, W = x.shape
        (_, _), (h4, w4), _, _ = manet_arch.feature_sizes(H, W)
        out = torch.empty((B, manet_arch.EMBED_DIM, h4, w4), device=self.device, dtype=torch.float32)
        check(lib.ivosw_manet_encoder_forward(self._h, _ptr(x), B, H, W, _ptr(out), _stream(self.device)))
        return out

    # ------------------------------------------------------------------ AssessNet training step (csrc/train.cu, config C5)
    def train_begin(self, sd):
        """(Re)start training from an AssessNet state dict (reference key names): quality_assessment.py:300-312."""
        blob = pack_assess(sd)
        check(lib.ivosw_assess_train_begin(self._h, _np_ptr(blob), blob.size))
        self._train_template = {k: v.clone() for k, v in sd.items()}

    def train_step(self, imgs, probs, targets, valid, lr=5e-6, momentum=0.9, weight_decay=5e-4, apply=True):
        """One iteration of quality_assessment.py::train's loop body (:240-269).  imgs B x 3 x H x W, probs B x H x W,
        targets B, valid B (bool: `union[n] > 0`).  Returns (loss or None when no sample is valid, pred B numpy)."""
        tf = self._dev32(imgs)
        B, _, H, W = tf.shape
        tp = self._dev32(probs)
        tg = self._dev32(torch.as_tensor(targets).reshape(-1))
        vd = torch.as_tensor(np.asarray(valid)).to(self.device, torch.int32).contiguous()
        loss = C.c_float(0.0)
        pred = np.empty(B, dtype=np.float32)
        check(lib.ivosw_assess_train_step(self._h, _ptr(tf), 3 * H * W, _ptr(tp), H * W, B, H, W, _ptr(tg), _ptr(vd), lr, momentum,
                                          weight_decay, 1 if apply else 0, C.byref(loss), _np_ptr(pred), _stream(self.device)))
        lv = float(loss.value)
        return (None if lv != lv else lv), pred

    def train_apply(self, lr=5e-6, momentum=0.9, weight_decay=5e-4):
        """Accumulate + clamp + SGD on the current step's gradient buffer (after an all-reduce of ``train_grads_tensor``)."""
        check(lib.ivosw_assess_train_apply(self._h, lr, momentum, weight_decay, _stream(self.device)))

    def train_export(self, want_grads=False):
        """State dict (reference key names) of the parameters and BatchNorm buffers after the steps so far
        [, accumulated clamped gradients keyed like the parameters]."""
        n = int(lib.ivosw_assess_blob_floats())
        blob = np.empty(n, dtype=np.float32)
        grad = np.empty(n, dtype=np.float32) if want_grads else None
        check(lib.ivosw_assess_train_export(self._h, _np_ptr(blob), _np_ptr(grad), _stream(self.device)))
        sd = unpack_assess(blob, self._train_template)
        return (sd, unpack_assess(grad, self._train_template, grads=True)) if want_grads else sd

    def train_grads_tensor(self):
        """The current step's raw gradient (blob order) as a flat CUDA tensor that ALIASES the library's buffer (no copy):
        what a data-parallel run all-reduces in place between train_step(apply=False) and train_apply()."""
        p = C.c_void_p()
        n = C.c_size_t(0)
        check(lib.ivosw_assess_train_grads(self._h, C.byref(p), C.byref(n)))

        class _View:            # CUDA array interface: torch wraps the pointer without copying
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<f4", "data": (int(p.value), False), "version": 2}
        return torch.as_tensor(_View(), device=self.device)

    # ------------------------------------------------------------------ peer-memory gather (csrc/gather.cu)
    def gather_create(self, world, rank, capacity=1024):
        """Allocates this rank's gather buffer; returns its 64-byte IPC handle (bytes) for the start-up exchange."""
        h = (C.c_ubyte * 64)()
        check(lib.ivosw_gather_create(self._h, world, rank, capacity, h))
        self._gather = {"world": world, "rank": rank, "capacity": capacity, "open": False}
        return bytes(h)

    def gather_open(self, handles):
        """handles: world x 64 bytes in rank order."""
        assert len(handles) == 64 * self._gather["world"]
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        check(lib.ivosw_gather_open(self._h, buf))
        self._gather["open"] = True

    @property
    def gather_ready(self):
        g = getattr(self, "_gather", None)
        return bool(g and g["open"])

    def gather_post(self, mq_local, offset):
        """Posts this rank's slice (CUDA float64 tensor, may be empty) of the current round into every rank's buffer."""
        n = 0 if mq_local is None else mq_local.numel()
        check(lib.ivosw_gather_post(self._h, _ptr(mq_local) if n else None, n, offset, _stream(self.device)))

    def agent_action_gathered(self, annotated_counts, want_quality=True):
        """Waits for every rank's slice of the round, then Brain + argmax: (next_frame, q[T], mask_quality[T] or None)."""
        ann = np.ascontiguousarray(annotated_counts, dtype=np.float64)
        T = ann.shape[0]
        q = np.empty(T, dtype=np.float32)
        mq = np.empty(T, dtype=np.float64) if want_quality else None
        nf = C.c_int(-1)
        check(lib.ivosw_agent_action_gathered(self._h, _np_ptr(ann), T, _np_ptr(q), C.byref(nf), _np_ptr(mq),
                                              _stream(self.device)))
        return int(nf.value), q, mq

    def stage_timing(self, on=True):
        check(lib.ivosw_stage_timing(self._h, 1 if on else 0))

    def stage_times(self, reset=True):
        """Accumulated device milliseconds per stage since the last reset + number of conv launches."""
        ms = (C.c_float * 5)()
        n = C.c_longlong(0)
        check(lib.ivosw_stage_times(self._h, ms, C.byref(n), 1 if reset else 0))
        names = ("roi", "stem", "conv_stack", "head", "brain")
        return dict(zip(names, [float(v) for v in ms])), int(n.value)

    # ------------------------------------------------------------------ test hook: one conv layer
    def debug_conv_dims(self, layer_index):
        d = (C.c_int * 6)()
        check(lib.ivosw_debug_conv(self._h, layer_index, 0, None, None, None, 0, d, None))
        return dict(zip(("cin", "in_hw", "cout", "out_hw", "k", "stride"), [int(v) for v in d]))

    def debug_conv(self, layer_index, x_nhwc, residual_nhwc=None, conv_mode="simt_fp32"):
        """x_nhwc: B x H x W x Cin CUDA fp32 -> B x OH x OW x Cout fp32 (BN + ReLU [+ residual] applied)."""
        d = self.debug_conv_dims(layer_index)
        x = self._dev32(x_nhwc)
        B = x.shape[0]
        assert tuple(x.shape[1:]) == (d["in_hw"], d["in_hw"], d["cin"])
        res = self._dev32(residual_nhwc) if residual_nhwc is not None else None
        out = torch.empty((B, d["out_hw"], d["out_hw"], d["cout"]), device=self.device, dtype=torch.float32)
        check(lib.ivosw_debug_conv(self._h, layer_index, CONV_MODES[conv_mode], _ptr(x), _ptr(res), _ptr(out), B, None,
                                   _stream(self.device)))
        return out

    # ------------------------------------------------------------------ MANet tail (utils/utils_manet.py)
    def manet_tail(self, logits, H, W, masks_out=None, all_p_out=None, want_masks=True, want_probs=True):
        """logits: T x C x h x w CUDA fp32 -> (masks T x H x W fp32, all_P T x C x H x W fp32)."""
        logits = self._dev32(logits)
        T, Cn, h, w = logits.shape
        if want_masks and masks_out is None:
            masks_out = torch.empty((T, H, W), device=self.device, dtype=torch.float32)
        if want_probs and all_p_out is None:
            all_p_out = torch.empty((T, Cn, H, W), device=self.device, dtype=torch.float32)
        check(lib.ivosw_manet_tail(self._h, _ptr(logits), T, Cn, h, w, H, W, _ptr(masks_out if want_masks else None),
                                   _ptr(all_p_out if want_probs else None), _stream(self.device)))
        return masks_out, all_p_out

    def rough_roi(self, labels, dist=20):
        """utils/utils_manet.py::rough_ROI on a B x 1 x h x w CUDA tensor."""
        x = self._dev32(labels)
        B, _, h, w = x.shape
        out = torch.empty_like(x)
        check(lib.ivosw_rough_roi(self._h, _ptr(x), _ptr(out), B, h, w, dist, _stream(self.device)))
        return out

    # ------------------------------------------------------------------ ATNet glue (utils/utils_atnet.py)
    def reflect_pad(self, x, left, right, top, bottom):
        """torch.nn.ReflectionPad2d((left, right, top, bottom)) on an N x C x h x w CUDA fp32 tensor."""
        x = self._dev32(x)
        N, Cn, h, w = x.shape
        out = torch.empty((N, Cn, h + top + bottom, w + left + right), device=self.device, dtype=torch.float32)
        check(lib.ivosw_atnet_reflect_pad(self._h, _ptr(x), _ptr(out), N * Cn, h, w, left, right, top, bottom,
                                          _stream(self.device)))
        return out

    def sigmoid_blend(self, logit, prev_inplace=None, alpha=1.0):
        """prob = sigmoid(logit) (returned, logit's shape).  With ``prev_inplace`` (a contiguous CUDA tensor with
        logit.numel() elements, e.g. prob_map_of_frames[frame]): prev_inplace <- alpha*prob + (1-alpha)*prev_inplace."""
        logit = self._dev32(logit)
        prob = torch.empty_like(logit)
        if prev_inplace is None:
            check(lib.ivosw_atnet_sigmoid_blend(self._h, _ptr(logit), None, _ptr(prob), _ptr(prob), logit.numel(), 1.0, 0.0,
                                                _stream(self.device)))
            return prob
        assert prev_inplace.is_cuda and prev_inplace.dtype == torch.float32 and prev_inplace.is_contiguous()
        assert prev_inplace.numel() == logit.numel()
        check(lib.ivosw_atnet_sigmoid_blend(self._h, _ptr(logit), _ptr(prev_inplace), _ptr(prob), _ptr(prev_inplace),
                                            logit.numel(), float(alpha), float(1 - alpha), _stream(self.device)))
        return prob

    def atnet_assemble(self, prob_map, y0, x0, H, W):
        """prob_map: T x O x PH x PW CUDA fp32 -> all_P T x (O+1) x H x W (channel 0 zero, crop at (y0, x0))."""
        prob_map = self._dev32(prob_map)
        T, O, PH, PW = prob_map.shape
        out = torch.empty((T, O + 1, H, W), device=self.device, dtype=torch.float32)
        check(lib.ivosw_atnet_assemble(self._h, _ptr(prob_map), _ptr(out), T, O, PH, PW, y0, x0, H, W, _stream(self.device)))
        return out

    # ------------------------------------------------------------------ helpers
    def _dev32(self, t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != torch.float32:
            t = t.to(self.device, torch.float32)
        return t if t.is_contiguous() else t.contiguous()


_ENGINES = {}


def get_engine(device=None, conv_mode=None):
    """Process-wide engine per CUDA device (what the drop-in modules share)."""
    if device is None:
        device = torch.cuda.current_device()
    dev = torch.device(device if not isinstance(device, int) else "cuda:%d" % device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    e = _ENGINES.get(idx)
    if e is None:
        e = Engine(idx, conv_mode or DEFAULT_CONV_MODE)
        _ENGINES[idx] = e
    elif conv_mode is not None and conv_mode != e.conv_mode:
        e.set_conv_mode(conv_mode)
    return e
