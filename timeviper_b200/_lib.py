"""ctypes loader for libtimeviper_b200.so (the C ABI declared in include/timeviper_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing the ops raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TV_LIB_PATH") or os.path.join(HERE, "libtimeviper_b200.so")   # TV_LIB_PATH: tuning builds

TV_F32, TV_BF16 = 0, 1
TV_SSD_FULL, TV_SSD_STATE_ONLY, TV_SSD_DT_ONLY = 0, 1, 2
TV_ABI_VERSION = 4
TV_OK, TV_ERR_INVALID, TV_ERR_UNSUPPORTED, TV_ERR_CUDA, TV_ERR_WORKSPACE = 0, -1, -2, -3, -4


class ConvParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("initial_states", C.c_void_p), ("out", C.c_void_p), ("final_states", C.c_void_p),
                ("batch", C.c_int32), ("dim", C.c_int32), ("seqlen", C.c_int32), ("width", C.c_int32),
                ("x_batch_stride", C.c_int64), ("x_seq_stride", C.c_int64),
                ("out_batch_stride", C.c_int64), ("out_seq_stride", C.c_int64),
                ("silu", C.c_int32), ("dtype", C.c_int32)]


class RmsnormParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("z", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("out", C.c_void_p), ("rows", C.c_int64), ("d", C.c_int32), ("group_size", C.c_int32),
                ("x_row_stride", C.c_int64), ("z_row_stride", C.c_int64), ("out_row_stride", C.c_int64),
                ("eps", C.c_float), ("norm_before_gate", C.c_int32), ("dtype", C.c_int32)]


class AddRmsnormParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("residual", C.c_void_p), ("weight", C.c_void_p), ("sum_out", C.c_void_p),
                ("out", C.c_void_p), ("rows", C.c_int64), ("d", C.c_int32), ("dtype", C.c_int32),
                ("x_row_stride", C.c_int64), ("res_row_stride", C.c_int64), ("sum_row_stride", C.c_int64),
                ("out_row_stride", C.c_int64), ("eps", C.c_float), ("reserved", C.c_int32)]


class SsdParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dt", C.c_void_p), ("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p),
                ("D", C.c_void_p), ("z", C.c_void_p), ("dt_bias", C.c_void_p), ("initial_states", C.c_void_p),
                ("out", C.c_void_p), ("final_states", C.c_void_p), ("logdecay_sum", C.c_void_p),
                ("batch", C.c_int32), ("seqlen", C.c_int32), ("nheads", C.c_int32), ("headdim", C.c_int32),
                ("ngroups", C.c_int32), ("dstate", C.c_int32), ("chunk_size", C.c_int32),
                ("x_batch_stride", C.c_int64), ("x_seq_stride", C.c_int64), ("x_head_stride", C.c_int64),
                ("dt_batch_stride", C.c_int64), ("dt_seq_stride", C.c_int64), ("dt_head_stride", C.c_int64),
                ("b_batch_stride", C.c_int64), ("b_seq_stride", C.c_int64), ("b_group_stride", C.c_int64),
                ("c_batch_stride", C.c_int64), ("c_seq_stride", C.c_int64), ("c_group_stride", C.c_int64),
                ("z_batch_stride", C.c_int64), ("z_seq_stride", C.c_int64), ("z_head_stride", C.c_int64),
                ("d_has_hdim", C.c_int32), ("dt_softplus", C.c_int32),
                ("dt_min", C.c_float), ("dt_max", C.c_float),
                ("dtype", C.c_int32), ("mode", C.c_int32), ("force_simt", C.c_int32), ("reuse_dt_cumsum", C.c_int32)]


class ConvUpdateParams(C.Structure):
    _fields_ = [("x", C.c_void_p), ("conv_state", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("out", C.c_void_p), ("batch", C.c_int32), ("dim", C.c_int32), ("width", C.c_int32),
                ("state_len", C.c_int32), ("x_batch_stride", C.c_int64), ("out_batch_stride", C.c_int64),
                ("state_batch_stride", C.c_int64), ("state_dim_stride", C.c_int64),
                ("silu", C.c_int32), ("dtype", C.c_int32)]


class SsuParams(C.Structure):
    _fields_ = [("state", C.c_void_p), ("x", C.c_void_p), ("dt", C.c_void_p), ("A", C.c_void_p), ("B", C.c_void_p),
                ("C", C.c_void_p), ("D", C.c_void_p), ("z", C.c_void_p), ("dt_bias", C.c_void_p), ("out", C.c_void_p),
                ("batch", C.c_int32), ("nheads", C.c_int32), ("headdim", C.c_int32), ("ngroups", C.c_int32),
                ("dstate", C.c_int32),
                ("x_batch_stride", C.c_int64), ("x_head_stride", C.c_int64), ("x_dim_stride", C.c_int64),
                ("dt_batch_stride", C.c_int64), ("dt_head_stride", C.c_int64), ("dt_dim_stride", C.c_int64),
                ("a_head_stride", C.c_int64), ("a_dim_stride", C.c_int64), ("a_state_stride", C.c_int64),
                ("b_batch_stride", C.c_int64), ("b_group_stride", C.c_int64),
                ("c_batch_stride", C.c_int64), ("c_group_stride", C.c_int64),
                ("d_head_stride", C.c_int64), ("d_dim_stride", C.c_int64),
                ("z_batch_stride", C.c_int64), ("z_head_stride", C.c_int64), ("z_dim_stride", C.c_int64),
                ("bias_head_stride", C.c_int64), ("bias_dim_stride", C.c_int64),
                ("dt_softplus", C.c_int32), ("dt_min", C.c_float), ("dt_max", C.c_float),
                ("dtype", C.c_int32), ("state_dtype", C.c_int32)]


EXPORTS = ("tv_abi_version", "tv_last_error", "tv_causal_conv1d_fwd", "tv_gated_rmsnorm_fwd", "tv_add_rmsnorm_fwd",
           "tv_ssd_workspace_bytes", "tv_ssd_chunk_scan_fwd", "tv_ssd_kernel_family",
           "tv_ssd_fold_boundary_states", "tv_ssd_fold_boundary_states_p2p", "tv_causal_conv1d_update", "tv_selective_state_update",
           "tv_debug_set_trace", "tv_debug_set_ablate", "tv_debug_launch_count")

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the CUDA extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python timeviper_b200/build.py` "
                          "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    lib.tv_abi_version.restype = C.c_int
    lib.tv_last_error.restype = C.c_char_p
    lib.tv_causal_conv1d_fwd.argtypes = [C.POINTER(ConvParams), C.c_void_p]
    lib.tv_causal_conv1d_fwd.restype = C.c_int
    lib.tv_gated_rmsnorm_fwd.argtypes = [C.POINTER(RmsnormParams), C.c_void_p]
    lib.tv_gated_rmsnorm_fwd.restype = C.c_int
    lib.tv_add_rmsnorm_fwd.argtypes = [C.POINTER(AddRmsnormParams), C.c_void_p]
    lib.tv_add_rmsnorm_fwd.restype = C.c_int
    lib.tv_ssd_workspace_bytes.argtypes = [C.POINTER(SsdParams)]
    lib.tv_ssd_workspace_bytes.restype = C.c_size_t
    lib.tv_ssd_chunk_scan_fwd.argtypes = [C.POINTER(SsdParams), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.tv_ssd_chunk_scan_fwd.restype = C.c_int
    lib.tv_ssd_kernel_family.argtypes = [C.POINTER(SsdParams)]
    lib.tv_ssd_kernel_family.restype = C.c_int
    lib.tv_ssd_fold_boundary_states.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                                C.c_void_p]
    lib.tv_ssd_fold_boundary_states.restype = C.c_int
    lib.tv_ssd_fold_boundary_states_p2p.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.tv_ssd_fold_boundary_states_p2p.restype = C.c_int
    lib.tv_causal_conv1d_update.argtypes = [C.POINTER(ConvUpdateParams), C.c_void_p]
    lib.tv_causal_conv1d_update.restype = C.c_int
    lib.tv_selective_state_update.argtypes = [C.POINTER(SsuParams), C.c_void_p]
    lib.tv_selective_state_update.restype = C.c_int
    lib.tv_debug_set_trace.argtypes = [C.c_void_p]
    lib.tv_debug_set_trace.restype = None
    lib.tv_debug_set_ablate.argtypes = [C.c_int]
    lib.tv_debug_set_ablate.restype = None
    lib.tv_debug_launch_count.argtypes = []
    lib.tv_debug_launch_count.restype = C.c_ulonglong
    if lib.tv_abi_version() != TV_ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.tv_abi_version()} != {TV_ABI_VERSION} (rebuild the library)")
    _lib = lib
    return lib


def check(rc, what):
    """Map a tv_status to the exception the reference's Python asserts would have produced."""
    if rc == TV_OK:
        return
    msg = f"{what}: {load().tv_last_error().decode(errors='replace')}"
    if rc == TV_ERR_INVALID:
        raise ValueError(msg)
    if rc == TV_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
