"""ctypes binding of the C ABI (include/abnet3_b200.h).

The shared library is the product: there is NO Python / torch fallback for any
of these entry points.  A missing library raises at import of this module's
``lib()``; a non-sm_100 device makes every call fail with ABN_ENOSYS.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ABN_LIB: an experiment build of the same library (tools/build_dbg.sh); never a fallback
SO_PATH = os.environ.get("ABN_LIB") or os.path.join(_HERE, "libabnet3_b200.so")

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_L = _c.c_int64
_F = _c.c_float



class GemmProblem(_c.Structure):
    """abn_gemm_problem (include/abnet3_b200.h)."""
    _fields_ = [("A", _P), ("lda", _L), ("a_mn", _I),
                ("B", _P), ("ldb", _L), ("b_mn", _I),
                ("M", _I), ("N", _I), ("K", _I),
                ("epilogue", _I), ("act", _I), ("split_k", _I),
                ("bias", _P),
                ("out", _P), ("ldo", _L), ("out_f32", _I),
                ("yprev", _P), ("ld_yprev", _L),
                ("ones_col", _I),
                ("ones_out", _P),
                ("signal", _P), ("wait", _P), ("wait_count", _I)]


class Dropout(_c.Structure):
    """abn_dropout (include/abnet3_b200.h): device {seed, step} state, p, layer id."""
    _fields_ = [("state", _P), ("p", _F), ("layer", _I)]


class MlpLayer(_c.Structure):
    """abn_mlp_layer (include/abnet3_b200.h)."""
    _fields_ = [("W", _P), ("ldw", _L), ("bias", _P), ("n_in", _I), ("n_out", _I), ("act", _I),
                ("out", _P), ("ldo", _L), ("out_f32", _I), ("ones_col", _I), ("drop", Dropout)]


class MlpLoss(_c.Structure):
    """abn_mlp_loss (include/abnet3_b200.h): the pair loss fused into the forward chain."""
    _fields_ = [("y", _P), ("loss", _P), ("dz", _P), ("ld_dz", _L), ("kind", _I), ("margin", _F),
                ("scale", _F), ("write_embeddings", _I)]


class MlpDLayer(_c.Structure):
    """abn_mlp_dlayer (include/abnet3_b200.h)."""
    _fields_ = [("W", _P), ("ldw", _L), ("n_in", _I), ("n_out", _I), ("act_below", _I),
                ("y_below", _P), ("ld_y", _L), ("dz_below", _P), ("ld_dz", _L),
                ("drop_below", Dropout)]


class ParamSegment(_c.Structure):
    """abn_param_segment (include/abnet3_b200.h)."""
    _fields_ = [("offset", _L), ("count", _L), ("ld", _L), ("bf16", _P), ("n_in", _I)]


class DpPeers(_c.Structure):
    """abn_dp_peers (include/abnet3_b200.h)."""
    _fields_ = [("grad", _P * 8), ("flags", _P * 8), ("rank", _I), ("world", _I)]


class DpPush(_c.Structure):
    """abn_dp_push (include/abnet3_b200.h)."""
    _fields_ = [("param", _P * 8), ("recv", _P * 8), ("flags", _P * 8), ("rank", _I), ("world", _I),
                ("n", _L), ("slice_cap", _L), ("one_shot", _I)]


# name -> (restype, argtypes); mirrors include/abnet3_b200.h declaration order
SIGNATURES = {
    "abn_version": (_I, []),
    "abn_last_error": (_c.c_char_p, []),
    "abn_device_info": (_I, [_P, _P, _P, _P]),
    "abn_align_workspace_bytes": (_c.c_size_t, [_I, _I, _I]),
    "abn_align_launches": (_I, [_I, _I, _I, _c.c_size_t]),
    "abn_stack_upload": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "abn_stack_from_frames": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "abn_pack_directions": (_I, [_P, _P, _P, _P, _I, _P, _P, _P]),
    "abn_stack_violations": (_I, [_P, _L, _I, _I, _P, _P, _P]),
    "abn_cosine_distance": (_I, [_P, _L, _I, _P, _I, _I, _I, _P, _P, _P, _P, _c.c_size_t, _P]),
    "abn_dtw_from_dist": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "abn_align_pairs": (_I, [_P, _L, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _c.c_size_t,
                              _P]),
    "abn_diff_pairs": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "abn_compact_paths": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "abn_store_scalar64": (_I, [_P, _P, _P]),
    "abn_gather_batch": (_I, [_P, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P]),
    "abn_pair_loss": (_I, [_P, _P, _P, _L, _I, _L, _I, _F, _F, _P, _P, _P, _P]),
    "abn_linear_forward": (_I, [_P, _P, _P, _L, _I, _I, _I, _I, _P, _P]),
    "abn_linear_backward": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "abn_dropout_mask": (_I, [_P, _L, _I, _P, _P]),
    "abn_linear_forward_drop": (_I, [_P, _P, _P, _L, _I, _I, _I, _I, _P, _P, _L, _P]),
    "abn_linear_backward_drop": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _L, _P]),
    "abn_gemm_bf16_group": (_I, [_P, _I, _P]),
    "abn_mlp_forward_fused": (_I, [_P, _L, _L, _P, _I, _P]),
    "abn_mlp_forward_loss_fused": (_I, [_P, _L, _L, _P, _I, _P, _P]),
    "abn_mlp_dgrad_fused": (_I, [_P, _L, _L, _P, _I, _P]),
    "abn_cast_bf16": (_I, [_P, _L, _I, _L, _P, _L, _P, _L, _P]),
    "abn_optimizer_step": (_I, [_P, _P, _P, _P, _L, _I, _F, _F, _F, _L, _P]),
    "abn_gather_batch_bf16": (_I, [_P, _I, _P, _P, _P, _P, _L, _P, _L, _P, _P, _I, _P]),
    "abn_gather_step_bf16": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _L, _L, _P, _L, _P, _P, _P, _I, _P, _I, _P]),
    "abn_pair_loss_dz": (_I, [_P, _P, _P, _L, _I, _L, _I, _F, _F, _I, _P, _P, _P, _L, _P]),
    "abn_pair_loss_dz_drop": (_I, [_P, _P, _P, _L, _I, _L, _I, _F, _F, _I, _P, _P, _P, _L, _P, _L, _I, _P]),
    "abn_optimizer_step_fused": (_I, [_P, _P, _P, _P, _I, _F, _F, _F, _L, _P, _I, _I, _P]),
    "abn_ipc_export": (_I, [_P, _P, _P]),
    "abn_ipc_import": (_I, [_P, _L, _P]),
    "abn_dp_optimizer_step": (_I, [_P, _P, _P, _I, _F, _F, _F, _L, _P, _I, _P, _P]),
    "abn_dp_grad_reset": (_I, [_P, _L, _P, _P]),
    "abn_dp_push_step": (_I, [_P, _P, _P, _I, _F, _F, _F, _L, _P, _I, _P, _P]),
    "abn_dp_set_trace": (_I, [_P]),
}

_lib = None


class AbnError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("abnet3_b200 error %d: %s" % (code, text))
        self.code = code


def lib():
    """Load libabnet3_b200.so (built by ``python -m abnet3_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m abnet3_b200.build` "
                "(nvcc, sm_100a).  There is no fallback path." % SO_PATH)
        handle = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise AbnError(rc, lib().abn_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
