"""ctypes binding of lib/libivosw_b200.so (C ABI: include/ivosw_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to
load, importing this module raises — the product path never degrades to a CPU /
PyTorch implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IVOSW_LIB: another build of the SAME library (A/B timing of kernel variants); never a different implementation
LIB_PATH = os.environ.get("IVOSW_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libivosw_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_STATE = 0, 1, 2, 3, 4
CONV_SIMT_FP32, CONV_TC_FP16X3, CONV_TC_FP16X1 = 0, 1, 2
BRAIN_NUM_PARAMS = 180993

# name -> (restype, argtypes); every symbol include/ivosw_b200.h declares
_c_f = C.POINTER(C.c_float)
_c_d = C.POINTER(C.c_double)
_c_i = C.POINTER(C.c_int)
SYMBOLS = {
    "ivosw_abi_version": (C.c_int, []),
    "ivosw_last_error": (C.c_char_p, []),
    "ivosw_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ivosw_destroy": (None, [C.c_void_p]),
    "ivosw_set_conv_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "ivosw_launch_count": (C.c_longlong, [C.c_void_p]),
    "ivosw_last_h2d_bytes": (C.c_longlong, [C.c_void_p]),
    "ivosw_brain_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ivosw_brain_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ivosw_dqn_load_target": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ivosw_dqn_sync_target": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ivosw_dqn_reset_optimizer": (C.c_int, [C.c_void_p]),
    "ivosw_dqn_get_optimizer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_longlong), C.c_void_p]),
    "ivosw_dqn_set_optimizer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ivosw_conv_saturation_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.c_void_p]),
    "ivosw_dqn_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_float, C.c_float, C.c_float, _c_f, C.c_void_p, C.c_int, C.c_void_p]),
    "ivosw_dqn_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "ivosw_brain_get_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ivosw_assess_blob_floats": (C.c_size_t, []),
    "ivosw_assess_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ivosw_assess_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ivosw_enable_probes": (C.c_int, [C.c_void_p, C.c_int]),
    "ivosw_assess_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, _c_i, C.c_void_p]),
    "ivosw_round_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _c_i,
                                     C.c_void_p]),
    "ivosw_round_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _c_i, C.c_void_p]),
    "ivosw_agent_action": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, _c_i, C.c_void_p]),
    "ivosw_score_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ivosw_score_shard_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ivosw_agent_action_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, _c_i, C.c_void_p]),
    "ivosw_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "ivosw_stage_times": (C.c_int, [C.c_void_p, _c_f, C.POINTER(C.c_longlong), C.c_int]),
    "ivosw_debug_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _c_i,
                                   C.c_void_p]),
    "ivosw_manet_encoder_blob_floats": (C.c_size_t, []),
    "ivosw_manet_encoder_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ivosw_manet_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ivosw_assess_train_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ivosw_assess_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, _c_f,
                                          C.c_void_p, C.c_void_p]),
    "ivosw_assess_train_apply": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "ivosw_assess_train_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ivosw_assess_train_grads": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "ivosw_gather_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ivosw_gather_open": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ivosw_gather_post": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ivosw_agent_action_gathered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, _c_i, C.c_void_p, C.c_void_p]),
    "ivosw_atnet_reflect_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p]),
    "ivosw_atnet_sigmoid_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                            C.c_float, C.c_float, C.c_void_p]),
    "ivosw_atnet_assemble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ivosw_rough_roi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ivosw_manet_tail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "ivosw_b200: %s not found. Build it with `python __graft_entry__.py build` (or `make -C "
            "ivos-w_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class IvoswError(RuntimeError):
    pass


def check(rc):
    """Maps a C status to a Python exception.  CUDA OOM must surface as a RuntimeError whose
    text contains "out of memory" (eval_agent_manet.py:391-396 retries on exactly that)."""
    if rc == OK:
        return
    msg = (lib.ivosw_last_error() or b"").decode("utf-8", "replace")
    if rc == ERR_OOM:
        raise RuntimeError("CUDA out of memory. " + msg)
    if rc == ERR_INVALID:
        raise ValueError("ivosw_b200: " + msg)
    raise IvoswError("ivosw_b200 (status %d): %s" % (rc, msg))
