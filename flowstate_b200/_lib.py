"""ctypes binding of libflowstate_b200.so (declared in include/flowstate_b200.h).

There is no CPU fallback: every product entry point goes through this library,
and loading fails loudly if it has not been built
(`python -m flowstate_b200.build`).
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflowstate_b200.so")

FS_RNG_PCG64, FS_RNG_PHILOX, FS_RNG_REPLAY, FS_RNG_PHILOX_REF = 0, 1, 2, 3
FS_PREC_FP32, FS_PREC_TF32 = 0, 1


class FsPot(C.Structure):
    _fields_ = [("num_wells", C.c_int), ("V0", C.c_double * 2), ("r0", C.c_double), ("k", C.c_double),
                ("r_cut", C.c_double), ("r_core", C.c_double)]


class FsRng(C.Structure):
    _fields_ = [("kind", C.c_int), ("pcg_state", C.c_void_p), ("philox_seed", C.c_ulonglong),
                ("chain_id0", C.c_longlong), ("replay_idx", C.c_void_p), ("replay_u", C.c_void_p),
                ("idx_stride", C.c_int), ("u_stride", C.c_int), ("replay_cursor", C.c_void_p)]


class FsTrainDesc(C.Structure):
    _fields_ = [("K", C.c_int), ("N", C.c_int), ("H", C.c_int), ("n_blocks", C.c_int), ("nb", C.c_int),
                ("bound", C.c_double), ("feature_scale", C.c_double), ("bn_eps", C.c_double),
                ("bn_momentum", C.c_double), ("transform_features", C.POINTER(C.c_int)),
                ("identity_features", C.POINTER(C.c_int)), ("params", C.POINTER(C.c_void_p)),
                ("grads", C.POINTER(C.c_void_p)), ("bn_running", C.POINTER(C.c_void_p))]


class FsLayerParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("init_w", "init_b", "bn_w", "bn_b", "bn_mean", "bn_var", "lin_w",
                                          "lin_b", "final_w", "final_b", "un_w", "un_h", "un_d")]


class FsFlowDesc(C.Structure):
    _fields_ = [("K", C.c_int), ("N", C.c_int), ("H", C.c_int), ("n_blocks", C.c_int), ("nb", C.c_int),
                ("bound", C.c_double), ("bn_eps", C.c_float), ("identity_features", C.c_void_p),
                ("transform_features", C.c_void_p), ("layers", C.POINTER(FsLayerParams))]


_P = C.c_void_p
_PROTOS = {
    "fs_last_error": (C.c_char_p, []),
    "fs_version": (C.c_int, []),
    "fs_launch_count": (C.c_ulonglong, []),
    "fs_set_device": (C.c_int, [C.c_int]),
    "fs_apply_pbc": (C.c_int, [_P, C.c_longlong, C.c_float, C.c_float, _P]),
    "fs_distances": (C.c_int, [_P, C.c_int, _P, C.c_longlong, C.c_float, C.c_float, _P, _P]),
    "fs_lj_pair": (C.c_int, [_P, C.c_longlong, C.POINTER(FsPot), _P, _P, _P]),
    "fs_double_well": (C.c_int, [_P, C.c_longlong, C.c_float, C.c_float, C.POINTER(FsPot), _P, _P]),
    "fs_energy_total": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(FsPot), _P, _P, _P, _P]),
    "fs_energy_particle": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(FsPot),
                                     _P, _P, _P, _P]),
    "fs_local_sweep": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                 C.c_double, C.POINTER(FsPot), C.POINTER(FsRng), _P, _P, _P, _P]),
    "fs_adjust_displacement": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_int, _P]),
    "fs_accept_global": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(FsRng), C.c_double, _P, _P, _P,
                                   C.c_int, C.c_int, _P]),
    "fs_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_longlong, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_float, _P]),
    "fs_accept_global_fused": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(FsRng), C.c_double, _P, _P, _P,
                                         C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(FsPot), _P]),
    "fs_tc_debug_read": (C.c_int, [_P, C.c_int]),
    "fs_flow_create": (C.c_int, [C.POINTER(FsFlowDesc), C.POINTER(_P)]),
    "fs_flow_destroy": (None, [_P]),
    "fs_flow_update": (C.c_int, [_P, C.POINTER(FsFlowDesc), _P]),
    "fs_flow_workspace_bytes": (C.c_size_t, [_P, C.c_int, C.c_int]),
    "fs_flow_conditioner": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P, C.c_size_t, C.c_int, _P]),
    "fs_target_energy": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(FsPot), _P, _P, _P]),
    "fs_spline_train_fwd": (C.c_int, [_P, _P, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _P, _P, _P]),
    "fs_spline_train_bwd": (C.c_int, [_P, _P, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _P, _P,
                                      _P, _P, _P]),
    "fs_affine_coupling": (C.c_int, [_P, C.c_longlong, _P, _P, C.c_longlong, C.c_int, _P, C.c_longlong, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_longlong, _P, _P]),
    "fs_periodic_shift": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "fs_classify_wells": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P]),
    "fs_pair_histogram": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _P, _P]),
    "fs_flow_has_tensor_path": (C.c_int, [_P]),
    "fs_flow_coupling": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P]),
    "fs_flow_coupling_all": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, _P]),
    "fs_flow_uses_layer_parallel": (C.c_int, [_P, C.c_int, C.c_int]),
    "fs_flow_set_layer_parallel": (C.c_int, [_P, C.c_int]),
    "fs_train_create": (C.c_int, [_P, _P]),
    "fs_train_destroy": (None, [_P]),
    "fs_train_forward_kld": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "fs_flow_tiled_features_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "fs_flow_tile_features": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "fs_flow_inverse": (C.c_int, [_P, _P, C.c_int, C.c_double, _P, _P, _P, _P, _P, C.c_size_t, C.c_int, _P]),
    "fs_flow_forward": (C.c_int, [_P, _P, C.c_int, C.c_double, _P, _P, _P, _P, C.c_size_t, C.c_int, _P]),
}

EXPORTS = tuple(_PROTOS)
_lib = None


class FlowStateError(RuntimeError):
    pass


def lib():
    """Loads the CUDA library (once).  Raises if it is missing - no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlowStateError(
                "flowstate_b200: %s not found; build it with `python -m flowstate_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise FlowStateError("flowstate_b200: %s (code %d)" % (lib().fs_last_error().decode(), rc))


def require_cuda(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise FlowStateError("flowstate_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
    if not t.is_contiguous():
        raise FlowStateError("flowstate_b200: %s must be contiguous" % name)
    return t


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


_bound_device = None


def bind_device(device=None):
    """Points the library's CUDA runtime at `device` (default: torch's current device)."""
    global _bound_device
    idx = torch.cuda.current_device() if device is None or getattr(device, "index", device) is None \
        else getattr(device, "index", device)
    if idx != _bound_device:
        check(lib().fs_set_device(int(idx)))
        _bound_device = idx


def stream_ptr(device=None):
    bind_device(device)
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_pot(num_wells, V0_list, r0, k, r_cut=2.5, r_core=0.5):
    p = FsPot()
    p.num_wells = int(num_wells)
    v = list(V0_list) if V0_list is not None else []
    v = (v + [0.0, 0.0])[:2]
    p.V0[0], p.V0[1] = float(v[0]), float(v[1])
    p.r0, p.k, p.r_cut, p.r_core = float(r0), float(k), float(r_cut), float(r_core)
    return p
