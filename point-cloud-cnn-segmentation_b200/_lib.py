"""ctypes binding of libpcseg_b200.so (include/pcseg_b200.h).  There is NO fallback: if the
library is missing the import of any product module raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpcseg_b200%s.so" % os.environ.get("PCSEG_LIB_SUFFIX", ""))   # (suffix: A/B builds)

EXPORTS = [
    "pcseg_last_error", "pcseg_version", "pcseg_create", "pcseg_destroy", "pcseg_param_count", "pcseg_param_offset",
    "pcseg_param_numel", "pcseg_bn_buffer_count", "pcseg_bn_buffer_offset", "pcseg_workspace_bytes", "pcseg_bind",
    "pcseg_prepare_eval", "pcseg_forward_eval", "pcseg_forward_eval_ragged", "pcseg_ragged_plan", "pcseg_forward_eval_part", "pcseg_pooled_feature", "pcseg_forward_train",
    "pcseg_forward_train_ragged", "pcseg_backward", "pcseg_adam_step",
    "pcseg_gemm_test", "pcseg_launch_count", "pcseg_set_sm_limit", "pcseg_ipc_export", "pcseg_peer_ar_signal_bytes", "pcseg_peer_ar_create",
    "pcseg_peer_ar_open", "pcseg_peer_ar_run", "pcseg_peer_ar_destroy", "pcseg_debug_copy", "pcseg_step_advance", "pcseg_eval_metrics", "pcseg_profile_enable", "pcseg_profile_read", "pcseg_profile_reset",
]


class PcsegError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  pcseg_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, ll, i32, f32, u64 = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_ulonglong
    lib.pcseg_last_error.restype = C.c_char_p
    lib.pcseg_version.restype = C.c_char_p
    lib.pcseg_create.argtypes = [C.POINTER(vp), i32]
    lib.pcseg_destroy.argtypes = [vp]
    for name in ("pcseg_param_count",):
        getattr(lib, name).argtypes = [i32]
        getattr(lib, name).restype = ll
    for name in ("pcseg_param_offset", "pcseg_param_numel"):
        getattr(lib, name).argtypes = [i32, i32]
        getattr(lib, name).restype = ll
    lib.pcseg_bn_buffer_count.restype = ll
    lib.pcseg_bn_buffer_offset.argtypes = [i32, i32]
    lib.pcseg_bn_buffer_offset.restype = ll
    lib.pcseg_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.pcseg_workspace_bytes.restype = ll
    lib.pcseg_bind.argtypes = [vp, i32, i32, vp, ll, i32]
    lib.pcseg_prepare_eval.argtypes = [vp, vp, vp, vp]
    lib.pcseg_forward_eval.argtypes = [vp, vp, vp, vp, vp]
    lib.pcseg_forward_train.argtypes = [vp, vp, vp, vp, u64, f32, vp, vp, vp, vp, vp, vp]
    lib.pcseg_forward_eval_part.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.pcseg_pooled_feature.argtypes = [vp, C.POINTER(vp)]
    lib.pcseg_ragged_plan.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(i32), ll, C.POINTER(ll), C.POINTER(i32)]
    lib.pcseg_ragged_plan.restype = ll
    lib.pcseg_forward_eval_ragged.argtypes = [vp, vp, C.POINTER(i32), i32, vp, vp, vp]
    lib.pcseg_forward_train_ragged.argtypes = [vp, vp, C.POINTER(i32), i32, vp, vp, u64, f32, vp, vp, vp, vp, vp, vp]
    lib.pcseg_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.pcseg_adam_step.argtypes = [vp, vp, vp, vp, ll, i32, f32, f32, f32, f32, f32, f32, vp, vp, vp]
    lib.pcseg_step_advance.argtypes = [vp, f32, f32, vp]
    lib.pcseg_eval_metrics.argtypes = [vp, vp, ll, i32, vp, vp, vp, vp, vp]
    lib.pcseg_gemm_test.argtypes = [i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp]
    lib.pcseg_launch_count.restype = ll
    lib.pcseg_set_sm_limit.argtypes = [i32]
    lib.pcseg_ipc_export.argtypes = [vp, C.c_char_p, C.POINTER(ll)]
    lib.pcseg_peer_ar_signal_bytes.argtypes = [ll, i32]
    lib.pcseg_peer_ar_signal_bytes.restype = ll
    lib.pcseg_peer_ar_create.argtypes = [C.POINTER(vp), i32, i32, vp, ll, vp, vp, vp, vp]
    lib.pcseg_peer_ar_open.argtypes = [vp, i32, C.c_char_p, ll, C.c_char_p, ll]
    lib.pcseg_peer_ar_run.argtypes = [vp, vp]
    lib.pcseg_peer_ar_destroy.argtypes = [vp]
    lib.pcseg_profile_enable.argtypes = [vp, i32]
    lib.pcseg_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(ll)]
    lib.pcseg_profile_reset.argtypes = [vp]
    lib.pcseg_debug_copy.argtypes = [vp, i32, i32, vp, ll, C.POINTER(ll), C.POINTER(ll), C.POINTER(i32), vp]
    return lib


lib = _load()


def check(rc: int, what: str = "pcseg call"):
    if rc != 0:
        raise PcsegError(f"{what}: {lib.pcseg_last_error().decode()}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
