"""ctypes binding of liblds_b200.so (C ABI declared in include/lds_b200.h).

There is no CPU or PyTorch fallback: if the library is missing, or the device is not sm_100, every
compute entry point raises. Loading the library and reading its symbols works without a GPU.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_uint8, c_uint32, c_uint64, c_void_p

from . import _build

LDS_OK = 0
K1_EXPLICIT_U = 1
K2_SIMT, K2_SINGLE_BF16, K2_FORCE_STREAMK, K2_NO_FUSE, K2_DUMP_ADJ, K2_FORWARD_ONLY, K2_BF16_ADJ = 1, 2, 4, 8, 16, 32, 64
K3_DENSE_GRAD, K3_ACCUMULATE, K3_SIMT = 1, 2, 4
OPT_SGD, OPT_ADAM = 0, 1
STREAM_EDGES, STREAM_DROP_X, STREAM_DROP_H = 0, 1, 2
PHASE_SAMPLE, PHASE_LAYER1, PHASE_LAYER2, PHASE_BWD2, PHASE_BWD1, PHASE_UPDATE, PHASE_ALL = 1, 2, 4, 8, 16, 32, 63

# every symbol include/lds_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "lds_version": (c_int32, []),
    "lds_last_error": (c_char_p, []),
    "lds_device_check": (c_int32, []),
    "lds_padded_ld": (c_int64, [c_int32]),
    "lds_philox_uniform": (c_float, [c_uint64, c_uint64, c_uint32, c_uint32, c_uint32, c_uint32]),
    "lds_theta_triu_to_full": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "lds_theta_full_to_triu": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p]),
    "lds_theta_clamp": (c_int32, [c_void_p, c_int64, c_int32, c_void_p]),
    "lds_theta_stats": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "lds_k1_sample_normalize": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_uint64, c_uint64, c_uint32,
                                          c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                          c_uint32, c_void_p]),
    "lds_k1_sample_normalize_dstep": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_uint64, c_void_p, c_uint64, c_uint32,
                                                c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "lds_spmm_csr": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int32,
                               c_void_p, c_int64, c_void_p]),
    "lds_row_linear": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "lds_gram_tn_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "lds_masked_nll_forward": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "lds_masked_nll_grad": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "lds_masked_nll_grad_grad": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                           c_void_p, c_int64, c_void_p, c_void_p]),
    "lds_row_dot2": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p, c_void_p]),
    "lds_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p, c_float, c_float, c_float, c_float, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lds_adam_step_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                         c_float, c_void_p, c_float, c_float, c_float, c_float, c_float, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lds_gram_tn": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "lds_packed_adj_bytes": (c_int64, [c_int32, c_int32]),
    "lds_k1_sample_packed": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_uint64, c_uint64, c_uint32, c_void_p, c_int64,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lds_k2_packed_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "lds_k2_propagate_packed": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64,
                                          c_void_p, c_int64, c_uint32, c_void_p]),
    "lds_outer_step_plan": (c_int32, [c_void_p]),
    "lds_k2_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "lds_k2_propagate": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_int64, c_void_p, c_int64, c_uint32, c_void_p]),
    "lds_k3k4_theta_update": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_int32,
                                        c_void_p, c_float, c_int32, c_void_p, c_void_p, c_float, c_float, c_float, c_int32,
                                        c_void_p, c_int64, c_uint32, c_void_p]),
    "lds_k3_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "lds_k3k4_theta_update_tc": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_int32,
                                           c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "lds_outer_step_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "lds_outer_step": (c_int32, [c_void_p, c_void_p]),
    "lds_outer_step_shard_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lds_outer_step_shard_buffer": (c_void_p, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lds_outer_step_buffer": (c_void_p, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lds_outer_step_factor_ld": (c_int64, [c_int32, c_int32]),
    "lds_peer_push": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p]),
    "lds_outer_step_packed_k": (c_int64, [c_int32, c_int32]),
    "lds_outer_step_operand_hp": (c_int32, [c_int32, c_int32, c_uint32]),
    "lds_outer_step_state_ld": (c_int64, [c_int32]),
    "lds_eval_metrics_workspace_bytes": (c_int64, [c_int32, c_int32]),
    "lds_eval_metrics": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "lds_knn_workspace_bytes": (c_int64, [c_int32]),
    "lds_knn_graph": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "lds_symmetrize_max": (c_int32, [c_void_p, c_int64, c_int32, c_void_p]),
    "lds_edges_to_dense": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "lds_edge_offsets": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "lds_remove_edges_apply": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "lds_upload_async": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "lds_profile_begin": (c_int32, []),
    "lds_profile_end": (c_int32, [c_void_p, c_void_p, c_int32]),
}


class OuterStepArgs(Structure):
    """Mirror of `lds_outer_step_args` (include/lds_b200.h); `struct_bytes` is checked by the library."""
    _fields_ = [
        ("struct_bytes", c_uint32),
        ("n", c_int32), ("f", c_int32), ("h", c_int32), ("c", c_int32),
        ("theta_full", c_void_p), ("ld_theta", c_int64),
        ("x", c_void_p), ("ld_x", c_int64),
        ("x_crow", c_void_p), ("x_col", c_void_p), ("x_val", c_void_p), ("reserved_ptr", c_void_p),
        ("w0", c_void_p), ("ld_w0", c_int64),
        ("b0", c_void_p), ("w1", c_void_p), ("b1", c_void_p),
        ("y", c_void_p), ("mask", c_void_p),
        ("mask_count", c_int32), ("dropout_p", c_float),
        ("seed", c_uint64), ("step", c_uint64),
        ("u_explicit", c_void_p), ("ld_u", c_int64),
        ("keep_x", c_void_p), ("keep_h", c_void_p),
        ("lr", c_float), ("opt_kind", c_int32),
        ("adam_m", c_void_p), ("adam_v", c_void_p),
        ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("adam_t", c_int32),
        ("update", c_int32),
        ("out_scalars", c_void_p), ("out_logp", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("k2_flags", c_uint32), ("k3_flags", c_uint32),
        ("row0", c_int32), ("rows", c_int32), ("phases", c_uint32), ("reserved2", c_uint32),
        ("opnd_full", c_void_p), ("fa_full", c_void_p), ("fb_full", c_void_p), ("c_full", c_void_p),
        ("f_full", c_void_p), ("k2_timeline", c_void_p),
        ("num_samples", c_int32), ("sample_index", c_int32), ("fpack_multi", c_void_p),
        ("opnd_send", c_void_p), ("opnd_rank_rows", c_int32), ("reserved3", c_int32),
        ("scalars_tag", c_float), ("reserved4", c_float),
    ]


_lib = None


def library_path():
    return _build.LIB_PATH


def load():
    """Load liblds_b200.so and bind every declared symbol. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the LDS hot path)")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().lds_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != LDS_OK:
        raise RuntimeError(f"liblds_b200 {what} failed (code {rc}): {last_error()}")


_device_ok = False


def require_device():
    """Fail loudly unless a CUDA device of compute capability 10.x is current."""
    global _device_ok
    if _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("liblds_b200 needs a B200 (sm_100a) GPU: torch.cuda.is_available() is False and there is no CPU fallback")
    check(load().lds_device_check(), "lds_device_check")
    _device_ok = True
