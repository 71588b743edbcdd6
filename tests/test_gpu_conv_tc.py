"""GPU (-m gpu): the tcgen05 implicit-GEMM convolution against (a) a plain PyTorch fp32 reference
of the same op and (b) the fp32 CUDA-core kernel, for every shape class of res2..res5.

Tolerances: split-fp16 (3 MMA terms) is fp32-grade: |d| <= 2e-5 * max|ref| ; single-term fp16:
|d| <= 4e-3 * max|ref| (fp16 inputs, 11-bit significand)."""
import numpy as np
import pytest
import torch

from ivosw import arch, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from ivosw.engine import Engine
    e = Engine(0, "simt_fp32")
    e.load_assess(synth.assess_state_dict(0))
    yield e
    e.close()


def _torch_ref(sd, spec, x_nhwc, res_nhwc):
    w = sd[spec.name + ".weight"].cuda()
    bn = {k: sd[spec.bn + "." + k].cuda() for k in ("weight", "bias", "running_mean", "running_var")}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = torch.nn.functional.conv2d(x_nhwc.permute(0, 3, 1, 2).double(), w.double(), None, spec.stride, spec.pad)
    y = torch.nn.functional.batch_norm(y, bn["running_mean"].double(), bn["running_var"].double(), bn["weight"].double(),
                                       bn["bias"].double(), False, 0.0, 1e-5)
    if res_nhwc is not None:
        y = y + res_nhwc.permute(0, 3, 1, 2).double()
    if spec.relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).float().contiguous()


# one representative of every (cin, cout, k, stride, hw) class + the first layers
def _layer_ids():
    seen, ids = set(), []
    for i, c in enumerate(arch.resnet50_convs()):
        key = (c.cin, c.cout, c.k, c.stride, c.in_hw, c.residual)
        if key not in seen:
            seen.add(key)
            ids.append(i)
    return ids


@pytest.mark.parametrize("li", _layer_ids())
def test_conv_tc_matches_fp32(eng, li):
    spec = arch.resnet50_convs()[li]
    sd = synth.assess_state_dict(0)
    g = torch.Generator(device="cuda").manual_seed(100 + li)
    for B in (2, 3):
        x = torch.randn((B, spec.in_hw, spec.in_hw, spec.cin), device="cuda", generator=g).relu_()
        res = None
        if spec.residual:
            res = torch.randn((B, spec.out_hw, spec.out_hw, spec.cout), device="cuda", generator=g)
        ref = _torch_ref(sd, spec, x, res)
        scale = float(ref.abs().max()) + 1e-6
        simt = eng.debug_conv(li, x, res, "simt_fp32")
        assert float((simt - ref).abs().max()) <= 2e-5 * scale, "simt"
        x3 = eng.debug_conv(li, x, res, "tc_fp16x3")
        err3 = float((x3 - ref).abs().max())
        assert err3 <= 2e-5 * scale, "tc_fp16x3 layer %d (%s): err %.3g scale %.3g" % (li, spec.name, err3, scale)
        x1 = eng.debug_conv(li, x, res, "tc_fp16x1")
        err1 = float((x1 - ref).abs().max())
        assert err1 <= 4e-3 * scale, "tc_fp16x1 layer %d: err %.3g scale %.3g" % (li, err1, scale)


@pytest.mark.parametrize("li", [3, 6, 13, 16, 26, 29, 45, 48])
def test_staged_epilogue_is_deterministic(eng, li):
    """Regression: the staged (ring) epilogue once aliased mbarrier phases for 4-K-block tiles and produced
    run-to-run differences.  Same inputs must give bit-identical outputs, and they must match the direct
    epilogue variant (IVOSW_NO_STAGED_EPILOGUE is not needed: compare against the fp32 kernel bound)."""
    spec = arch.resnet50_convs()[li]
    g = torch.Generator(device="cuda").manual_seed(7 + li)
    B = 48
    x = torch.randn((B, spec.in_hw, spec.in_hw, spec.cin), device="cuda", generator=g).relu_()
    res = torch.randn((B, spec.out_hw, spec.out_hw, spec.cout), device="cuda", generator=g)
    first = eng.debug_conv(li, x, res, "tc_fp16x3").clone()
    for _ in range(30):
        again = eng.debug_conv(li, x, res, "tc_fp16x3")
        assert torch.equal(first, again)
    simt = eng.debug_conv(li, x, res, "simt_fp32")
    assert float((first - simt).abs().max()) <= 4e-5 * (float(simt.abs().max()) + 1e-6)
