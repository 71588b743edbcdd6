"""CPU: host-side logic — the C-ABI library loads and exports every symbol the header declares,
weight packing follows the documented order, the drop-in modules expose the reference's state-dict
keys, and the product path refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from ivosw import arch, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, REPO)
    import __graft_entry__ as ge
    return ge.build()


def test_header_symbols_exported(built_lib):
    hdr = open(os.path.join(REPO, "include", "ivosw_b200.h")).read()
    declared = set(re.findall(r"IVOSW_API[^;(]*?\b(ivosw_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 17
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    from ivosw import _lib
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.lib.ivosw_abi_version() == 1


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ or torch types in the signatures) and a C
    program must link against the library through it."""
    src = tmp_path / "abi.c"
    src.write_text('#include "ivosw_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %s\\n", ivosw_abi_version(), ivosw_last_error()); return 0; }\n')
    inc = os.path.join(REPO, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                   check=True)


def test_sass_is_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_sass_is_blackwell_native(built_lib):
    """The encoder kernels really are tcgen05 / TMEM / TMA code (B200_PROFILING.md: the PTX names never appear in SASS:
    tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG / UBLKCP), with no legacy mma.sync (HMMA)."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    per_fn, fn = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            per_fn[fn] = set()
        elif fn:
            for op in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "HMMA.", "HGMMA", "UTCHMMA.2CTA",
                       "UTCBAR.2CTA.MULTICAST"):
                if op == "HMMA.":                       # legacy mma.sync only: not the tail of UTCHMMA.2CTA
                    if re.search(r"(?<![A-Z])HMMA\.", line):
                        per_fn[fn].add(op)
                elif op in line:
                    per_fn[fn].add(op)
    conv = [f for f in per_fn if "conv_tc_kernel" in f or "conv_stack_kernel" in f]
    assert len(conv) >= 6
    for f in conv:
        assert {"UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"} <= per_fn[f], (f, per_fn[f])
    # conv_tc_kernel<BN, STAGES, STAGED, PAIR>: ...ELb<staged>ELb<pair>EE...
    staged = [f for f in conv if "Lb1ELb" in f or "conv_stack" in f]
    assert staged and all("UTMASTG" in per_fn[f] for f in staged)
    pairs = [f for f in conv if "ELb1EEEv" in f]          # cta_group::2 variants: pair MMAs and multicast commits
    assert len(pairs) >= 2 and all({"UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST"} <= per_fn[f] for f in pairs)
    stem = [f for f in per_fn if "stem_tc_kernel" in f]
    assert stem and {"UTCHMMA", "UBLKCP", "LDTM"} <= per_fn[stem[0]]
    assert not any({"HMMA.", "HGMMA"} & ops for ops in per_fn.values())


def test_blob_sizes(built_lib):
    from ivosw import _lib, engine
    assert arch.BRAIN_NUM_PARAMS == _lib.BRAIN_NUM_PARAMS == 180993
    b = engine.pack_brain(synth.brain_state_dict(0))
    assert b.dtype == np.float32 and b.size == 180993
    sd = synth.assess_state_dict(0)
    a = engine.pack_assess(sd)
    assert a.size == _lib.lib.ivosw_assess_blob_floats() == 23566343
    # spot-check the documented order: mean/std first, fc last, stem is OHWI with prob as cin 3
    np.testing.assert_allclose(a[:6], [0.485, 0.456, 0.406, 0.229, 0.224, 0.225], rtol=1e-7)
    assert a[-1] == sd["fc1.bias"].item()
    stem = a[6:6 + 64 * 49 * 4].reshape(64, 7, 7, 4)
    np.testing.assert_array_equal(stem[..., 3], sd["Encoder.conv1_p.weight"][:, 0].numpy())
    np.testing.assert_array_equal(stem[..., :3], sd["Encoder.conv1.weight"].permute(0, 2, 3, 1).numpy())


def test_no_cpu_fallback(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ivosw import _lib, engine
    h = ctypes.c_void_p()
    rc = _lib.lib.ivosw_create(0, 0, ctypes.byref(h))
    assert rc != 0 and _lib.lib.ivosw_last_error()
    with pytest.raises(RuntimeError):
        engine.Engine(0)


def _dropin():
    """drop-in modules imported the way an entry script gets them: ivosw.hook over a (test double) checkout"""
    from tests import doubles
    d = doubles.load_dropin()
    assert d.A.__file__.startswith(os.path.join(REPO, "ivos-w_b200", "dropin"))
    assert d.misc.__file__.startswith(doubles.CHECKOUT)          # the checkout's own utils.misc is still reachable
    return d.A, d.S, d.U


def test_dropin_state_dict_contract(built_lib):
    A, S, U = _dropin()
    net = S.AssessNet()
    keys = set(net.state_dict().keys())
    assert keys == set(arch.assess_state_dict_keys())
    net.load_state_dict(synth.assess_state_dict(0), strict=True)
    brain = A.Brain()
    assert [k for k, _ in arch.BRAIN_PARAMS] == list(brain.state_dict().keys())
    brain.load_state_dict(synth.brain_state_dict(0), strict=True)
    assert sum(p.numel() for p in brain.parameters()) == 180993
    assert sum(p.numel() for p in net.parameters()) == 23519553   # incl. unused conv1_m (w, b), conv1_n
    with pytest.raises(RuntimeError):
        brain(torch.zeros(1, 4, 2))          # CPU tensor: loud failure, no fallback
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, 8, 8))


def test_dropin_glue_host_helpers(built_lib):
    A, S, U = _dropin()
    v = np.array([0.5, 0.2, 0.9, 0.1])
    assert U.select_next_frame(v, metric="worst", prev_frames=[3]) == 1
    assert U.select_next_frame(v, metric="max") == 2
    assert U.gen_subseq(5, 20, 4, "consecutive") == [3, 4, 5, 6]
    s = U.gen_subseq(7, 30, 5, "equal")
    assert 7 in s and len(s) == 5


def test_synthetic_clip_properties():
    all_F, all_P, ann = synth.make_clip(0, 8, 64, 96, 2)
    assert all_F.shape == (8, 3, 64, 96) and all_P.shape == (8, 3, 64, 96)
    assert all_F.min() >= 0 and all_F.max() <= 1
    np.testing.assert_allclose(all_P.sum(1), 1.0, atol=1e-5)
    a2 = synth.make_clip(0, 8, 64, 96, 2)
    np.testing.assert_array_equal(all_P, a2[1])
    _, P_at, _ = synth.make_clip(1, 5, 64, 96, 2, "atnet")
    assert (P_at[:, 0] == 0).all()
