"""ctypes binding of libb3d.so (include/b3d.h). The product path has NO fallback: if the
library is missing or a call fails, this raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb3d.so")

MASK_NONE, MASK_RELU, MASK_SIGMOID = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
FLAG_ACCUMULATE = 1
FLAG_OUT_BF16 = 2
FLAG_SPLIT = 4
F32, BF16, BITS, F64 = 0, 1, 2, 3
MAX_SEGS = 8


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("idx", C.c_void_p), ("mask", C.c_void_p), ("width", C.c_int32),
                ("ld", C.c_int32), ("ldmask", C.c_int32), ("mask_mode", C.c_int32), ("dtype", C.c_int32)]


class ChainLayer(C.Structure):
    """b3d_chain_layer_t"""
    _fields_ = [("K", C.c_int32), ("N", C.c_int32), ("src", C.c_int32), ("act", C.c_int32), ("nadd", C.c_int32),
                ("add_idx", C.c_int32 * 2), ("add_ld", C.c_int32 * 2), ("add_ptr", C.c_void_p * 2),
                ("bias", C.c_void_p), ("out", C.c_void_p), ("ldo", C.c_int32), ("bits_out", C.c_void_p),
                ("bits_in", C.c_void_p)]


ACT_MASKBITS = 3
NARROW_INPUT_RELU = 0x100      # flag of b3d_narrow_mlp_*'s final_act (include/b3d.h)
CHAIN_MAX_LAYERS = 6

_SIGS = {
    "b3d_last_error": (C.c_char_p, []),
    "b3d_launch_count": (C.c_int64, []),
    "b3d_reset_launch_count": (None, []),
    "b3d_csr_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "b3d_csr_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 6 +
                      [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "b3d_segment_sum": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "b3d_gather_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                  C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "b3d_add_n": (C.c_int, [C.POINTER(Seg), C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "b3d_linear": (C.c_int, [C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                             C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                             C.c_int32, C.c_void_p, C.POINTER(Seg), C.c_int32, C.c_void_p]),
    "b3d_wgrad_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "b3d_wgrad": (C.c_int, [C.POINTER(Seg), C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                            C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b3d_tc_packed_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b3d_tc_pack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p]),
    "b3d_linear_tc": (C.c_int, [C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_int32, C.c_int32, C.c_void_p, C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_void_p]),
    "b3d_tma_packed_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "b3d_tma_packed_bytes_segs": (C.c_size_t, [C.c_int32, C.POINTER(C.c_int32), C.c_int32]),
    "b3d_tma_pack_weights_segs": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p]),
    "b3d_tma_pack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    "b3d_linear_tma": (C.c_int, [C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_int32, C.c_int32, C.c_void_p, C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_void_p]),
    "b3d_wgrad_tc_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "b3d_wgrad_tc": (C.c_int, [C.POINTER(Seg), C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                               C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b3d_wgrad_tma_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "b3d_wgrad_tma": (C.c_int, [C.POINTER(Seg), C.POINTER(Seg), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "b3d_narrow_mlp_supported": (C.c_int, [C.c_int32, C.POINTER(C.c_int32)]),
    "b3d_narrow_mlp_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "b3d_narrow_mlp_bwd_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.POINTER(C.c_int32)]),
    "b3d_narrow_mlp_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                     C.c_size_t, C.c_void_p]),
    "b3d_knn_frames": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int64,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_gat_aggregate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_gat_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_row_nonzero": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "b3d_bce_partials": (C.c_int64, [C.c_int64]),
    "b3d_bce_fwd_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_focal_fwd_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_chain_supported": (C.c_int, [C.POINTER(ChainLayer), C.c_int32, C.c_int32]),
    "b3d_chain_packed_bytes": (C.c_size_t, [C.POINTER(ChainLayer), C.c_int32, C.c_int32]),
    "b3d_chain_pack_weights": (C.c_int, [C.POINTER(ChainLayer), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                         C.c_int32, C.c_void_p, C.c_void_p]),
    "b3d_chain_run": (C.c_int, [C.POINTER(Seg), C.c_int32, C.POINTER(ChainLayer), C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_void_p]),
    "b3d_window_knn": (C.c_int, [C.c_void_p] * 7 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_hier_tracks_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b3d_adam_step_dev": (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5 + [C.c_void_p, C.c_float, C.c_void_p]),
    "b3d_adam_step": (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5 + [C.c_int32, C.c_float,
                                                                                  C.c_void_p]),
}

_lib = None


def exported_symbols():
    """Names declared in include/b3d.h (used by the CPU test that checks the .so exports them)."""
    return sorted(_SIGS)


def lib():
    """Load libb3d.so (once). Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m batch3dmot_b200.build` "
                               "(the product path has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"libb3d {what} failed (rc={rc}): {lib().b3d_last_error().decode()}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """The current CUDA stream of the current device as a cudaStream_t. Called once per launch (~500 times per
    training step): torch.cuda.current_stream() builds a Stream object through several Python layers (~7 us, a fifth of
    the host time of an eager small-batch step), the raw accessor is one C call."""
    if _raw_stream is not None and _cur_device is not None:
        return C.c_void_p(_raw_stream(_cur_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _check_f32_rows(t, allow_bf16=False):
    ok = (torch.float32, torch.bfloat16) if allow_bf16 else (torch.float32,)
    assert t.is_cuda and t.dtype in ok and t.dim() == 2 and t.stride(1) == 1, \
        f"expected a CUDA fp32 row-major matrix, got {t.dtype} {tuple(t.shape)} {t.stride()}"


def make_segs(items):
    """items: list of (tensor[rows,width], idx int32 or None, mask tensor or None, mask_mode)."""
    assert 1 <= len(items) <= MAX_SEGS
    arr = (Seg * len(items))()
    for s, (t, idx, mask, mode) in zip(arr, items):
        _check_f32_rows(t, allow_bf16=True)
        s.ptr, s.width, s.ld = t.data_ptr(), t.size(1), t.stride(0)
        s.dtype = BF16 if t.dtype == torch.bfloat16 else F32
        s.idx = idx.data_ptr() if idx is not None else None
        if idx is not None:
            assert idx.dtype == torch.int32 and idx.is_contiguous()
        if mask is not None:
            _check_f32_rows(mask)
            s.mask, s.ldmask, s.mask_mode = mask.data_ptr(), mask.stride(0), mode
        else:
            s.mask, s.ldmask, s.mask_mode = None, 0, MASK_NONE
    return arr


def ptr_array(ts):
    """Host array of device pointers (None -> NULL) for the per-layer parameter lists."""
    arr = (C.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def int_array(vals):
    return (C.c_int32 * len(vals))(*vals)


def launch_count():
    return int(lib().b3d_launch_count())


def reset_launch_count():
    lib().b3d_reset_launch_count()
