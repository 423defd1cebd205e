"""ctypes binding of libfumi_b200.so (the C ABI in include/fumi_b200.h).

No fallback: if the library is missing it is built in-tree (nvcc must be present); if that
fails, importing the compute path raises.  Every call checks the returned fumi_status and raises
FumiError with fumi_last_error().
"""
import ctypes as C
import os

from . import build as _build

_LIB = None


class FumiError(RuntimeError):
    pass


class EpisodeCfg(C.Structure):
    _fields_ = [("num_ways", C.c_int32), ("num_support", C.c_int32), ("num_query", C.c_int32),
                ("hid0", C.c_int32), ("hid1", C.c_int32), ("steps", C.c_int32),
                ("step_size", C.c_float), ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64),
                ("task_offset", C.c_int64), ("first_order", C.c_int32), ("reserved", C.c_int32)]


class StashLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("per_task", "S", "w1t", "b0", "b1", "head", "steps", "per_step",
                                         "format", "rec_h1", "rec_exp", "rec_h0_hi", "rec_h0_lo")]


_P = C.c_void_p
_I64, _I32, _F = C.c_int64, C.c_int32, C.c_float
_SIGS = {
    "fumi_abi_version": (C.c_int, []),
    "fumi_last_error": (C.c_char_p, []),
    "fumi_device_sm_count": (C.c_int, []),
    "fumi_linear_fwd": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I32, _I32, _P]),
    "fumi_linear_wgrad": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I32, _I32, _P]),
    "fumi_linear_dgrad": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "fumi_tanh_bwd": (C.c_int, [_P, _P, _I64, _P]),
    "fumi_split_tf32": (C.c_int, [_P, _P, _P, _I64, _P]),
    "fumi_transpose_split_tf32": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P]),
    "fumi_gemm_tf32x3": (C.c_int, [_P] * 6 + [_I64] * 6 + [_I32] * 3 + [_P]),
    "fumi_gram_f16": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _P, _I64, _I32, _I32, _P, _P]),
    "fumi_absmax": (C.c_int, [_P, _I64, _P, _P]),
    "fumi_split_f16": (C.c_int, [_P, _P, _P, _P, _I64, _P]),
    "fumi_transpose_split_f16": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "fumi_gemm_f16x3": (C.c_int, [_P] * 8 + [_I64] * 6 + [_I32] * 3 + [_P]),
    "fumi_gram": (C.c_int, [_P, _I64, _I64, _P, _P, _I64, _I32, _I32, _P, _P]),
    "fumi_episode_stash_floats": (C.c_int64, [C.POINTER(EpisodeCfg)]),
    "fumi_stash_layout": (C.c_int, [C.POINTER(EpisodeCfg), C.POINTER(StashLayout)]),
    "fumi_episode_fwd": (C.c_int, [C.POINTER(EpisodeCfg), _I64] + [_P] * 17),
    "fumi_episode_bwd_parts": (C.c_int, []),
    "fumi_episode_bwd": (C.c_int, [C.POINTER(EpisodeCfg), _I64] + [_P] * 7 + [_F] + [_P] * 6),
    "fumi_reduce_parts": (C.c_int, [_P, _I64, _I64, _P, _I32, _P]),
    "fumi_scatter_add_rows": (C.c_int, [_P, _P, _I64, _I64, _P, _P]),
    "fumi_reduce_loss_acc": (C.c_int, [_P, _P, _I64, _P, _P]),
    "fumi_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _I32, _P]),
    "fumi_am3_score": (C.c_int, [_P] * 8 + [_I64, _I32, _I32, _I32, _I32, _I32] + [_P] * 5),
    "fumi_confusion_counts": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "fumi_am3_bwd": (C.c_int, [_P] * 8 + [_I64, _I32, _I32, _I32, _I32, _I32] + [_P, _P, _F] + [_P] * 4),
    "fumi_dropout_apply": (C.c_int, [_P, _I64, _I64, C.c_uint64, C.c_uint32, _F, _P]),
    "fumi_sigmoid_bwd": (C.c_int, [_P, _P, _I64, _P]),
    "fumi_sampler_create": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, C.POINTER(_P)]),
    "fumi_sampler_destroy": (None, [_P]),
    "fumi_sampler_new_iterator": (C.c_int, [_P, _P]),
    "fumi_sampler_next": (C.c_int, [_P, _I64] + [_P] * 11 + [_I32]),
    "fumi_sampler_plan": (C.c_int, [_P, _I64] + [_P] * 8),
    "fumi_sampler_expand": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32] + [_P] * 7),
    "fumi_py_tuple_hash": (_I64, [_P, _I64]),
    "fumi_debug_phase_profile": (C.c_int, [C.c_int]),
    "fumi_debug_read_phases": (C.c_int, [_P]),
}
EXPORTED = tuple(_SIGS)


def load(path):
    """Bind a build of the C ABI (every symbol of include/fumi_b200.h must be exported)."""
    global _LIB
    L = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)          # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    if L.fumi_abi_version() != 1:
        raise FumiError("libfumi_b200.so ABI version mismatch")
    _LIB = L
    return L


def require_cuda(device, what):
    """There is no CPU path: every compute front-end calls this with its device."""
    import torch
    if torch.device(device).type != "cuda":
        raise FumiError(f"{what} runs on CUDA devices only (got {device}); there is no CPU path")


def lib():
    if _LIB is None:
        path = _build.LIB
        if not os.path.exists(path):
            path = _build.build()
        load(path)
    return _LIB


def check(status, what=""):
    if status < 0:
        raise FumiError(f"{what}: status {status}: {lib().fumi_last_error().decode()}")
    return status


def ptr(t):
    """data pointer of a torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous(), "C ABI takes dense row-major buffers"
        return C.c_void_p(t.data_ptr())
    assert t.flags["C_CONTIGUOUS"]
    return C.c_void_p(t.ctypes.data)


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
