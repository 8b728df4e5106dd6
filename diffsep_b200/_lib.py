"""ctypes binding of ``libdsep.so`` — the C-ABI declared in ``include/dsep.h``.

There is no fallback: if the library is missing or a call fails, the caller gets an exception.
``DSEP_ERR_INVALID`` maps to ``ValueError`` and ``DSEP_ERR_UNSUPPORTED`` to ``NotImplementedError``
(what the reference raises for bad arguments / unsupported SDEs), CUDA errors to ``RuntimeError``
(what a failed ``TORCH_CHECK`` in the reference's pybind11 ops raises).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# DSEP_LIB: load another build of the same sources (kernel experiments, tools/build_variant.sh)
LIB_PATH = Path(os.environ.get("DSEP_LIB") or Path(__file__).resolve().parent / "libdsep.so")

ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = -1, -2, -3
ABI_VERSION = 8

_p, _i, _f, _i64, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint64


class SdeParams(C.Structure):
    _fields_ = [("d_lambda", C.c_float), ("sigma_min", C.c_float), ("sigma_max", C.c_float),
                ("T_end", C.c_float), ("ndim", C.c_int)]


class ConvArgs(C.Structure):
    """dsep_conv_args of include/dsep.h (field order and types must match exactly)."""
    _fields_ = [("a_hi", _p), ("a_lo", _p), ("x0", _p), ("x1", _p), ("C0", _i), ("C1", _i), ("sc", _p), ("sh", _p),
                ("act", _i), ("B", _i), ("H", _i), ("W", _i), ("Cin", _i), ("w_hi", _p), ("w_lo", _p),
                ("Cout_pad", _i), ("ksize", _i), ("a2_hi", _p), ("a2_lo", _p), ("s0", _p), ("s1", _p), ("S0", _i),
                ("S1", _i), ("Cin2", _i), ("w2_hi", _p), ("w2_lo", _p), ("bias", _p), ("film", _p),
                ("film_stride", _i), ("residual", _p), ("scale", _f), ("acc_scale", _f), ("out", _p),
                ("cout_store", _i), ("stats", _p), ("passes", _i)]


# name -> argument types, exactly the prototypes of include/dsep.h (return type int)
PROTOTYPES = {
    "dsep_conv2d_tc": [_p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _p, _p, _p, _p, _i, _p, _f, _f, _p, _i,
                       _p, _i, _p],
    "dsep_split_f16": [_p, _i64, _f, _p, _p, _p],
    "dsep_conv2d_fused": [C.POINTER(ConvArgs), _p],
    "dsep_conv2d_fused8": [C.POINTER(ConvArgs), _f, _i, _p],
    "dsep_gn_tables": [_p, _i, _p, _i, _i, _i, _i, _p, _p, _f, _p, _p, _p],
    "dsep_channel_stats": [_p, _i, _i, _i, _p, _p],
    "dsep_zero": [_p, _i64, _p],
    "dsep_gn_act_split": [_p, _i, _p, _p, _i, _p, _i, _i, _i, _p, _p, _f, _i, _p, _p, _p, _p, _p],
    "dsep_gn_stats_act_split": [_p, _i, _p, _p, _i, _p, _i, _i, _i, _p, _p, _f, _i, _p, _p, _p, _p, _i, _p],
    "dsep_fir_resample": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p],
    "dsep_fir_resample8": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _f, _p, _p, _p, _p, _p, _i, _p],
    "dsep_fir_resample_f32": [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _f, _p, _p, _p],
    "dsep_upfirdn2d": [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p],
    "dsep_combine": [_p, _i, _p, _p, _p, _p, _i, _i, _i, _p, _p],
    "dsep_add": [_p, _p, _p, _i64, _p],
    "dsep_im2col3x3": [_p, _i, _i, _i, _i, _i, _p, _p],
    "dsep_tap_gather3x3": [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "dsep_attention": [_p, _i, _i, _i, _f, _p, _p, _p],
    "dsep_time_embedding": [_p, _p, _p, _p, _p, _p, _i, _i, _p, _p],
    "dsep_film": [_p, _p, _p, _i, _i, _i, _p, _p],
    "dsep_stft_frames": [_p, _p, _i, _i, _i, _i, _p, _p],
    "dsep_sgemm": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _p],
    "dsep_spec_pack": [_p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p, _p],
    "dsep_out_head": [_p, _i, _i, _i, _i, _i, _p, _p, _p, _f, _f, _p, _p],
    "dsep_istft_ola": [_p, _p, _i, _i, _i, _i, _p, _p],
    "dsep_sde_prior": [C.POINTER(SdeParams), _p, _i, _f, _p, _i, _p, _u64, _u64, _i, _i, _p, _p],
    "dsep_sde_corrector": [C.POINTER(SdeParams), _p, _p, _p, _p, _p, _u64, _u64, _f, _i, _i, _p, _p, _p],
    "dsep_sde_predictor": [C.POINTER(SdeParams), _p, _p, _p, _p, _p, _u64, _u64, _f, _i, _i, _i, _p, _p, _p],
    "dsep_sde_perturb": [C.POINTER(SdeParams), _p, _p, _p, _p, _u64, _u64, _i, _i, _p, _p, _p],
    "dsep_score_loss": [C.POINTER(SdeParams), _p, _p, _p, _p, _i, _i, _p, _p],
    "dsep_sde_corrector_ald": [C.POINTER(SdeParams), _p, _p, _p, _p, _u64, _u64, _f, _i, _i, _p, _p, _p],
    "dsep_sde_corrector_langevin": [_p, _p, _p, _f, _i, _i, _p, _p, _p, _p],
    "dsep_sigma_mix": [_p, _i, _i, _i, _p, _p],
    "dsep_normalize": [_p, _i, _i, _p, _p, _p, _p],
    "dsep_scale_output": [_p, _p, _i, _i, _i, _p, _p],
    "dsep_randn": [_p, _i64, _u64, _u64, _p],
}
OTHER_SYMBOLS = ("dsep_last_error", "dsep_abi_version", "dsep_device_ok", "dsep_conv_kblock", "dsep_has_fp8_corr",
                 "dsep_source_hash", "dsep_conv_wide_launches")

_lib = None


def load():
    """Loads libdsep.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        # a source checkout without the binary: compile it in-tree (nvcc, sm_100a).  This is still the
        # CUDA path — there is no CPU or PyTorch fallback for the DiffSep hot path.
        try:
            from . import build as _build
            _build.build()
        except Exception as e:
            raise RuntimeError(f"{LIB_PATH} is missing and could not be built "
                               f"(`python -m diffsep_b200.build`): {e}") from e
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.dsep_last_error.restype = C.c_char_p
    lib.dsep_last_error.argtypes = []
    lib.dsep_abi_version.restype = C.c_int
    lib.dsep_device_ok.restype = C.c_int
    lib.dsep_conv_kblock.restype = C.c_int
    lib.dsep_has_fp8_corr.restype = C.c_int
    lib.dsep_conv_wide_launches.restype = C.c_int
    lib.dsep_source_hash.restype = C.c_char_p
    if lib.dsep_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libdsep.so ABI {lib.dsep_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, name: str = "dsep"):
    if rc == 0:
        return
    msg = load().dsep_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"{name}: {msg} (code {rc})")


N_CALLS = 0   # C-ABI calls made so far; each launches one kernel (bench.py reports the count)


_DEBUG_SYNC = bool(int(__import__("os").environ.get("DSEP_DEBUG_SYNC", "0")))


def call(name: str, *args):
    global N_CALLS
    N_CALLS += 1
    check(getattr(load(), name)(*args), name)
    if _DEBUG_SYNC:   # debugging aid: attribute an asynchronous CUDA error to the call that caused it
        import torch
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise RuntimeError(f"{name}{args}: {e}") from e
