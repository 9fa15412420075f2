"""ctypes binding of ``libcrs_b200.so`` (C ABI in ``include/crs_b200.h``).

There is deliberately NO fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C contextual_rs_b200/csrc``)
every entry point raises.  The library itself is plain C ABI — device pointers, sizes and a
``cudaStream_t`` — so PyTorch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRS_B200_LIB") or os.path.join(_HERE, "libcrs_b200.so")  # env: A/B runs of two builds

CRS_OK, CRS_ERR_INVALID_ARG, CRS_ERR_UNSUPPORTED, CRS_ERR_CUDA, CRS_ERR_NOT_PSD, CRS_TIES_PENDING = range(6)
CRS_F64, CRS_F32 = 0, 1
CRS_MATERN52, CRS_RBF = 0, 1
CRS_LEARN_NORMALIZE, CRS_LEARN_STANDARDIZE, CRS_TRAIN_INPUTS_RAW = 1, 2, 4
CRS_P_COUNT, CRS_P_KDE, CRS_P_INFER = 0, 1, 2
CRS_MAX_DIM, CRS_MAX_OBS = 16, 160
CRS_HEAD_ELEMS = 2 * CRS_MAX_DIM + 8
KERNELS = {"matern52": CRS_MATERN52, "rbf": CRS_RBF}
DTYPES = {torch.float64: CRS_F64, torch.float32: CRS_F32}

vp, i32, i64, u32, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double


class GPState(C.Structure):
    """``crs_gp_state``"""
    _fields_ = [
        ("num_arms", i32), ("n_cap", i32), ("dim", i32), ("dim_pad", i32),
        ("kernel", i32), ("dtype", i32), ("flags", i32), ("n_max", i32),
        ("train_x", vp), ("train_y", vp), ("n_obs", vp), ("lengthscale", vp),
        ("outputscale", vp), ("noise", vp), ("mean_const", vp),
        ("norm_mins", vp), ("norm_ranges", vp), ("y_mean", vp), ("y_std", vp),
        ("xn", vp), ("xs", vp), ("alpha", vp), ("linv_t", vp),
        ("arm_head", vp), ("status", vp),
    ]


class Decision(C.Structure):
    """``crs_decision`` (64 bytes)"""
    _fields_ = [
        ("min_zeta", f64), ("psi_best", f64), ("psi_rest", f64), ("tie_count", i64),
        ("context", i32), ("zeta_arm", i32), ("best_arm", i32), ("next_arm", i32),
        ("reserved", i32 * 4),
    ]


class DecideWS(C.Structure):
    """``crs_decide_ws``"""
    _fields_ = [("ctx", vp), ("mean", vp), ("var", vp), ("best_arm", vp), ("min_zeta", vp),
                ("min_arm", vp), ("tie_cnt", vp), ("decision", vp), ("scan_ws", vp)]


class PeerExchange(C.Structure):
    """``crs_peer_exchange``"""
    _fields_ = [("mailboxes", vp), ("rank", i32), ("world", i32), ("seq", i32), ("reserved", i32), ("ctx_begin", i64)]


class LeviDecision(C.Structure):
    """``crs_levi_decision``"""
    _fields_ = [("value", f64), ("context", i32), ("arm", i32), ("reserved", i64)]


assert C.sizeof(Decision) == 64 and C.sizeof(LeviDecision) == 24

_SIGNATURES = {
    "crs_abi_version": (C.c_int, []),
    "crs_status_string": (C.c_char_p, [C.c_int]),
    "crs_last_cuda_error": (C.c_char_p, []),
    "crs_scan_workspace_bytes": (i64, []),
    "crs_linv_stride": (i64, [i32, i32]),
    "crs_gp_prepare": (C.c_int, [C.POINTER(GPState), i32, i32, vp]),
    "crs_gp_mll": (C.c_int, [C.POINTER(GPState), i32, i32, vp, vp, vp]),
    "crs_posterior_grid": (C.c_int, [C.POINTER(GPState), vp, i64, i32, i32, vp, vp, i64, i32, vp]),
    "crs_ocba_scan": (C.c_int, [vp, vp, i64, i32, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "crs_ocba_resolve_tie": (C.c_int, [vp, vp, i64, i32, i64, i32, vp, vp, vp, i64, vp, vp]),
    "crs_ocba_select": (C.c_int, [C.POINTER(GPState), vp, vp, vp, i64, i32, i32, f64, vp, vp]),
    "crs_ocba_scan_select": (C.c_int, [C.POINTER(GPState), vp, vp, vp, i64, i64, i32, i32, f64, vp, vp, vp, vp, vp, vp, vp]),
    "crs_ocba_scan_select_sharded": (C.c_int, [C.POINTER(GPState), vp, vp, vp, i64, i64, i32, f64, vp, vp, vp, vp, vp, vp,
                                               C.POINTER(PeerExchange), vp, vp]),
    "crs_wait_decision": (C.c_int, [vp, i32, i32]),
    "crs_arm_pack_best": (C.c_int, [vp, vp, i64, i64, i32, i32, vp, vp, vp]),
    "crs_arm_combine_best": (C.c_int, [vp, i32, i64, i32, vp, vp, vp, vp]),
    "crs_arm_record_mean": (C.c_int, [vp, vp, i64, i32, i32, vp]),
    "crs_arm_fold": (C.c_int, [vp, i32, vp, vp, vp]),
    "crs_arm_finish": (C.c_int, [vp, vp, vp, vp]),
    "crs_peer_sum_u64": (C.c_int, [vp, i32, vp, vp, C.POINTER(PeerExchange), vp]),
    "crs_wait_u64": (C.c_int, [vp, u64, i32]),
    "crs_mailbox_bytes": (i64, [i32]),
    "crs_device_alloc": (C.c_int, [i64, C.POINTER(vp)]),
    "crs_device_free": (C.c_int, [vp]),
    "crs_ipc_export": (C.c_int, [vp, vp]),
    "crs_ipc_open": (C.c_int, [vp, C.POINTER(vp)]),
    "crs_ipc_close": (C.c_int, [vp]),
    "crs_pcs_counts": (C.c_int, [vp, vp, i64, i32, i64, i32, i64, u64, u32, i32, i64, vp, vp, vp]),
    "crs_pcs_counts_tree": (C.c_int, [vp, vp, i64, i32, i64, i32, i64, u64, u32, i64, vp, vp, vp]),
    "crs_pcs_cdf_workspace_bytes": (i64, [i64]),
    "crs_pcs_counts_cdf": (C.c_int, [vp, vp, i64, i32, i64, i32, i64, u64, u32, i64, vp, vp, vp, vp]),
    "crs_pcs_counts_base": (C.c_int, [vp, vp, i64, i32, i64, i32, vp, i64, vp, vp]),
    "crs_debug_zfun": (C.c_int, [vp, i64, vp, vp]),
    "crs_levi_scan": (C.c_int, [C.POINTER(GPState), vp, vp, i64, i64, vp, vp, i64, vp, vp, vp, vp]),
    "crs_min_zeta_grad": (C.c_int, [C.POINTER(GPState), vp, i64, vp, vp, i64, vp, vp, vp, vp]),
    "crs_philox_words": (C.c_int, [vp, i64, u32, u32, vp, vp]),
    "crs_debug_corr": (C.c_int, [vp, i64, i32, i32, vp, vp]),
    "crs_peak_probe": (C.c_int, [i32, i32, i32, vp, vp]),
    "crs_gao_decide_host": (C.c_int, [C.POINTER(GPState), vp, i64, i32, i32, i32, f64, i32,
                                      C.POINTER(DecideWS), C.POINTER(Decision), vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def get() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C contextual_rs_b200/csrc` "
                "(or __graft_entry__.build()).  contextual_rs_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.crs_abi_version() != 1:
            raise RuntimeError("libcrs_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> int:
    """Raise on anything but CRS_OK / CRS_TIES_PENDING."""
    if rc == CRS_OK or rc == CRS_TIES_PENDING:
        return rc
    lib = get()
    msg = lib.crs_status_string(rc).decode()
    if rc == CRS_ERR_CUDA:
        msg += ": " + lib.crs_last_cuda_error().decode()
    raise RuntimeError(f"libcrs_b200 {what}: {msg} (status {rc})")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(device: torch.device) -> None:
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("contextual_rs_b200 runs on a CUDA device only (no CPU fallback); "
                           f"got device {device}")
