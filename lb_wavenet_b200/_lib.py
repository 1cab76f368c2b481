"""ctypes binding of libwavenet_b200.so (the C ABI declared in include/wavenet_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent this module
raises, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwavenet_b200.so")

WN_NSTATS = 4
STAT_XENT_SUM, STAT_N_VALID, STAT_DIFF_SUM, STAT_L2 = 0, 1, 2, 3
KIND_FILTER, KIND_BIAS = 0, 1


class WnArch(C.Structure):
    """struct wn_arch (include/wavenet_b200.h, ABI version 2)"""
    _fields_ = [(n, C.c_int32) for n in (
        "n_blocks", "n_block_layers", "n_quant", "n_res", "n_dil", "n_skip", "n_post",
        "n_gc_embed", "n_gc_category", "use_bias", "n_lc_in", "n_lc_out", "n_lc_layers")] + [("lc_upsample", C.c_int32 * 8)]

    @staticmethod
    def from_dict(d: dict) -> "WnArch":
        up = [int(x) for x in (d.get("lc_upsample") or [])]
        lc = int(d.get("n_lc_out", 0) or 0) > 0
        if len(up) > 8:
            raise ValueError("lc_upsample holds more than 8 strides")
        return WnArch(int(d["n_blocks"]), int(d["n_block_layers"]), int(d["n_quant"]), int(d["n_res"]), int(d["n_dil"]),
                      int(d["n_skip"]), int(d["n_post"]), int(d["n_gc_embed"]), int(d["n_gc_category"]),
                      int(d["use_bias"]), int(d.get("n_lc_in", 0) or 0) if lc else 0,
                      int(d.get("n_lc_out", 0) or 0), len(up) if lc else 0, (C.c_int32 * 8)(*(up if lc else [])))


class WaveNetLibError(RuntimeError):
    pass


_vp, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); must list every symbol declared in include/wavenet_b200.h
SIGNATURES = {
    "wn_abi_version": (_i32, []),
    "wn_crc32c": (C.c_uint32, [C.c_uint32, _vp, _u64]),
    "wn_last_error": (C.c_char_p, []),
    "wn_model_create": (C.c_int, [C.POINTER(WnArch), _i32, C.POINTER(_vp)]),
    "wn_model_destroy": (None, [_vp]),
    "wn_n_layers": (_i32, [_vp]),
    "wn_recep_field": (_i32, [_vp]),
    "wn_param_count": (_i32, [_vp]),
    "wn_param_elems": (_i64, [_vp]),
    "wn_param_info": (C.c_int, [_vp, _i32, C.c_char_p, _i32, C.POINTER(_i64), C.POINTER(_i32),
                                C.POINTER(_i64), C.POINTER(_i32)]),
    "wn_save_elems": (_i64, [_vp]),
    "wn_save_info": (C.c_int, [_vp, _i32, C.POINTER(_i64), C.POINTER(_i32)]),
    "wn_workspace_bytes": (_i64, [_vp, _i32]),
    "wn_train_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "wn_train_backward": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "wn_train_backward_phases": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp]),
    "wn_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "wn_l2_loss": (C.c_int, [_vp, _vp, _vp, _vp]),
    "wn_debug_read": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "wn_gen_workspace_bytes": (_i64, [_vp, _i32]),
    "wn_gen_reset": (C.c_int, [_vp, _vp, _i32, _vp]),
    "wn_gen_load_params": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp]),
    "wn_gen_run": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _u64, _vp, _i32, _vp, _vp, _vp]),
    "wn_mu_encode": (C.c_int, [_vp, _vp, _i64, _vp]),
    "wn_mu_decode": (C.c_int, [_vp, _vp, _i64, _vp]),
    "wn_sample_logits": (C.c_int, [_vp, _i32, _u64, _i64, _vp, _vp]),
    "wn_deal_plan": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64,
                               _vp]),
    "wn_deal_fill": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "wn_codes_u8_to_i32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "wn_selftest_umma_gemm": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "wn_selftest_umma_gemm_tn": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "wn_selftest_umma_gemm_pair": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "wn_prof_enable": (C.c_int, [_i32]),
    "wn_debug_trace": (C.c_int, [_vp, _i32]),
    "wn_prof_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(_i64)]),
    "wn_launch_count_reset": (_i64, []),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises WaveNetLibError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WaveNetLibError(
            "%s not found: build it with `python -m lb_wavenet_b200.build` "
            "(there is no CPU or library fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise WaveNetLibError("symbol %s missing from %s" % (name, LIB_PATH)) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().wn_last_error().decode("utf-8", "replace")
        raise WaveNetLibError("%s failed (%d): %s" % (what or "libwavenet_b200 call", rc, msg))


def ptr(t) -> Optional[int]:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
