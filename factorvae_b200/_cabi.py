"""ctypes binding of libfvae_b200.so (include/fvae_b200.h).  No torch types cross this boundary:
only raw device pointers, sizes and a stream handle."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FVAE_B200_LIB: load another build of the same ABI (A/B kernel timing); default: the in-tree library
LIB_PATH = os.environ.get("FVAE_B200_LIB") or os.path.join(_HERE, "libfvae_b200.so")

F32, BF16 = 0, 1
PREC_FP32, PREC_BF16_TC = 0, 1
FLAG_TRAIN, FLAG_PHILOX = 1, 2
FILL_NONE, FILL_FFILL, FILL_FFILL_BFILL = 0, 1, 2
ABI_VERSION = 6

SECTIONS = [
    "LN_W", "LN_B", "W1", "B1", "WIH", "WHH", "BIH", "BHH",
    "ENC_W", "ENC_B", "ENC_MU_W", "ENC_MU_B", "ENC_SG_W", "ENC_SG_B",
    "AL_W", "AL_B", "AL_MU_W", "AL_MU_B", "AL_SG_W", "AL_SG_B",
    "BETA_W", "BETA_B",
    "ATT_Q", "ATT_KW", "ATT_KB", "ATT_VW", "ATT_VB",
    "PR_W", "PR_B", "PR_MU_W", "PR_MU_B", "PR_SG_W", "PR_SG_B",
]


class Shape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("S", "B", "T", "C", "H", "K", "M")]


class Panel(C.Structure):
    # row_index / num_rows: resident-panel form (NULL / 0 for dense windows), see include/fvae_b200.h
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("seq_pitch", C.c_int64), ("row_pitch", C.c_int64),
                ("row_index", C.c_void_p), ("num_rows", C.c_int64)]


class Noise(C.Structure):
    _fields_ = [("eps", C.c_void_p), ("keep_mask", C.c_void_p), ("seed", C.c_uint64), ("step", C.c_uint64),
                ("unit_base", C.c_int64), ("step_dev", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("loss", "date_loss", "yhat", "mu_y", "sigma_y", "mu_post", "sigma_post",
                                          "mu_prior", "sigma_prior")]


class Parts(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("z_mu", "z_sigma", "alpha_mu", "alpha_sigma", "beta", "context")]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Fails loudly: there is no CPU or eager fallback for this path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m factorvae_b200.build` (nvcc, sm_100a). "
            "factorvae_b200 has no CPU / PyTorch-eager fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    L.fvae_abi_version.restype = C.c_int
    L.fvae_debug_launch_count.restype = C.c_uint64
    L.fvae_status_string.restype = C.c_char_p
    L.fvae_status_string.argtypes = [C.c_int]
    L.fvae_param_offsets.restype = C.c_int
    L.fvae_param_offsets.argtypes = [i32, i32, i32, i32, C.POINTER(i64)]
    L.fvae_param_count.restype = i64
    L.fvae_param_count.argtypes = [i32, i32, i32, i32]
    L.fvae_workspace_bytes.restype = i64
    L.fvae_workspace_bytes.argtypes = [C.POINTER(Shape), i32]
    L.fvae_elbo_forward.restype = C.c_int
    L.fvae_elbo_forward.argtypes = [C.POINTER(Shape), C.POINTER(Panel), vp, vp, vp, C.POINTER(Noise), u32, i32,
                                    C.POINTER(Outputs), vp, i64, vp]
    L.fvae_elbo_backward.restype = C.c_int
    L.fvae_elbo_backward.argtypes = [C.POINTER(Shape), C.POINTER(Panel), vp, vp, vp, C.POINTER(Noise), u32, i32,
                                     C.POINTER(Outputs), vp, vp, i64, vp]
    L.fvae_predict.restype = C.c_int
    L.fvae_predict.argtypes = [C.POINTER(Shape), C.POINTER(Panel), vp, vp, C.POINTER(Noise), u32, i32,
                               C.POINTER(Outputs), vp, i64, vp]
    L.fvae_fe_forward.restype = C.c_int
    L.fvae_fe_forward.argtypes = [C.POINTER(Shape), C.POINTER(Panel), vp, i32, vp, vp, i64, vp]
    L.fvae_fe_backward.restype = C.c_int
    L.fvae_fe_backward.argtypes = [C.POINTER(Shape), C.POINTER(Panel), vp, i32, vp, vp, vp, i64, vp]
    L.fvae_heads_parts.restype = C.c_int
    L.fvae_heads_parts.argtypes = [C.POINTER(Shape), vp, vp, vp, vp, C.POINTER(Noise), u32, C.POINTER(Parts), C.POINTER(Outputs),
                                   vp, i64, vp]
    L.fvae_debug_front_forward.restype = C.c_int
    L.fvae_debug_front_forward.argtypes = [C.POINTER(Shape), C.POINTER(Panel), vp, i64, vp]
    L.fvae_debug_noise.restype = C.c_int
    L.fvae_debug_noise.argtypes = [C.c_uint64, C.c_uint64, i64, i64, i32, vp, vp, vp]
    L.fvae_workspace_latent.restype = vp
    L.fvae_workspace_latent.argtypes = [C.POINTER(Shape), i32, vp]
    L.fvae_window_index.restype = C.c_int
    L.fvae_window_index.argtypes = [vp, i32, i32, vp, vp, i64, i32, i32, i32, vp, vp, vp, vp]
    L.fvae_gather_windows.restype = C.c_int
    L.fvae_gather_windows.argtypes = [C.POINTER(Panel), i64, i32, i32, vp, i32, vp]
    L.fvae_rank_ic.restype = C.c_int
    L.fvae_rank_ic.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.fvae_adam_step.restype = C.c_int
    L.fvae_adam_step.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64, C.c_float, vp]
    L.fvae_p2p_buffer_bytes.restype = i64
    L.fvae_p2p_buffer_bytes.argtypes = [i64, i32]
    L.fvae_p2p_alloc.restype = C.c_int
    L.fvae_p2p_alloc.argtypes = [i64, C.POINTER(vp), C.c_char_p]
    L.fvae_p2p_open.restype = C.c_int
    L.fvae_p2p_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.fvae_p2p_close.restype = C.c_int
    L.fvae_p2p_close.argtypes = [vp]
    L.fvae_p2p_free.restype = C.c_int
    L.fvae_p2p_free.argtypes = [vp]
    L.fvae_p2p_allreduce.restype = C.c_int
    L.fvae_p2p_allreduce.argtypes = [vp, i64, vp, i32, i32, u32, C.c_float, i64, vp]
    if L.fvae_abi_version() != ABI_VERSION:
        raise ImportError("libfvae_b200.so has an unexpected ABI version")
    _lib = L
    return L


EXPORTS = ["fvae_abi_version", "fvae_debug_launch_count", "fvae_debug_front_forward", "fvae_debug_noise", "fvae_status_string", "fvae_param_offsets", "fvae_param_count", "fvae_workspace_bytes",
           "fvae_elbo_forward", "fvae_elbo_backward", "fvae_predict", "fvae_fe_forward", "fvae_fe_backward", "fvae_heads_parts",
           "fvae_workspace_latent", "fvae_window_index", "fvae_gather_windows", "fvae_adam_step", "fvae_rank_ic",
           "fvae_p2p_buffer_bytes", "fvae_p2p_alloc", "fvae_p2p_open", "fvae_p2p_close", "fvae_p2p_free", "fvae_p2p_allreduce"]


class FvaeError(RuntimeError):
    pass


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().fvae_status_string(int(status)).decode()
        raise FvaeError(f"{what or 'fvae call'} failed with status {status}: {msg}")


def param_offsets(Cf: int, H: int, K: int, M: int):
    arr = (C.c_int64 * (len(SECTIONS) + 1))()
    check(lib().fvae_param_offsets(Cf, H, K, M, arr), "fvae_param_offsets")
    return list(arr)
