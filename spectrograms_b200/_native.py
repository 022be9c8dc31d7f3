"""ctypes binding of libsgx_b200.so (include/sgx_b200.h). There is no fallback: if the CUDA library cannot be loaded
this raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

from .errors import (DimensionMismatchError, FFTBackendError, InternalError, InvalidInputError)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsgx_b200.so")

SGX_OK, SGX_INVALID_INPUT, SGX_DIMENSION_MISMATCH, SGX_BACKEND_ERROR, SGX_INTERNAL_ERROR = range(5)

# every symbol include/sgx_b200.h declares (tests/test_host_api.py checks the header against this list)
EXPORTS = [
    "sgx_last_error_message", "sgx_last_dimension_mismatch", "sgx_version", "sgx_plan_create", "sgx_plan_destroy",
    "sgx_plan_output_shape", "sgx_plan_axes", "sgx_plan_window", "sgx_plan_filterbank", "sgx_plan_kernel_name",
    "sgx_plan_last_launch_count", "sgx_plan_force_generic", "sgx_plan_set_tensor_cores", "sgx_plan_set_tmem_exchange", "sgx_plan_compute_batch", "sgx_plan_compute_frame",
    "sgx_mfcc_from_log_mel", "sgx_rfft", "sgx_chroma_from_spectrogram", "sgx_chroma_filterbank",
    "sgx_binaural_from_stft", "sgx_plan_compute_binaural", "sgx_plan_istft", "sgx_irfft",
    "sgx_fft_planner_create", "sgx_fft_planner_destroy", "sgx_fft_planner_cached_plans", "sgx_fft_planner_rfft",
    "sgx_fft_planner_irfft", "sgx_fft_planner_power_spectrum",
]


class PlanDesc(C.Structure):
    """``sgx_plan_desc``."""
    _fields_ = [
        ("dtype", C.c_int), ("n_fft", C.c_size_t), ("hop_size", C.c_size_t), ("centre", C.c_int),
        ("window", C.c_int), ("window_param", C.c_double),
        ("custom_window", C.POINTER(C.c_double)), ("custom_window_len", C.c_size_t),
        ("sample_rate_hz", C.c_double),
        ("mapping", C.c_int), ("n_bands", C.c_size_t), ("f_min", C.c_double), ("f_max", C.c_double),
        ("mel_norm", C.c_int), ("erb_spacing", C.c_int),
        ("amp", C.c_int), ("has_floor_db", C.c_int), ("floor_db", C.c_double),
        ("output", C.c_int), ("n_mfcc", C.c_size_t), ("include_c0", C.c_int), ("lifter", C.c_size_t),
        ("device", C.c_int),
        ("chroma_tuning", C.c_double), ("chroma_norm", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library. Raises ``FFTBackendError`` if it is missing -- never falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFTBackendError(
            f"cuda FFT backend error: {LIB_PATH} is missing; build it with `python -m spectrograms_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i, d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
    L.sgx_last_error_message.restype = C.c_char_p
    L.sgx_last_dimension_mismatch.argtypes = [C.POINTER(sz), C.POINTER(sz)]
    L.sgx_version.restype = C.c_char_p
    L.sgx_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(vp)]
    L.sgx_plan_destroy.argtypes = [vp]
    L.sgx_plan_output_shape.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz)]
    L.sgx_plan_axes.argtypes = [vp, sz, vp, vp]
    L.sgx_plan_window.argtypes = [vp, vp]
    L.sgx_plan_filterbank.argtypes = [vp, vp, C.POINTER(sz)]
    L.sgx_plan_kernel_name.argtypes = [vp]
    L.sgx_plan_kernel_name.restype = C.c_char_p
    L.sgx_plan_last_launch_count.argtypes = [vp]
    L.sgx_plan_last_launch_count.restype = sz
    L.sgx_plan_force_generic.argtypes = [vp, i]
    L.sgx_plan_set_tensor_cores.argtypes = [vp, i]
    L.sgx_plan_set_tmem_exchange.argtypes = [vp, i]
    L.sgx_plan_compute_batch.argtypes = [vp, vp, sz, sz, sz, vp, sz, sz, sz, vp]
    L.sgx_plan_compute_frame.argtypes = [vp, vp, sz, sz, vp, vp]
    L.sgx_mfcc_from_log_mel.argtypes = [i, vp, sz, sz, sz, sz, i, sz, vp, i, vp]
    L.sgx_rfft.argtypes = [i, vp, sz, sz, vp, i, vp]
    L.sgx_chroma_from_spectrogram.argtypes = [i, vp, sz, sz, sz, d, sz, d, d, d, i, vp, i, vp]
    L.sgx_chroma_filterbank.argtypes = [d, sz, d, d, d, vp]
    L.sgx_binaural_from_stft.argtypes = [i, i, vp, vp, sz, sz, sz, sz, sz, d, sz, i, vp, i, vp]
    L.sgx_plan_istft.argtypes = [vp, vp, sz, sz, vp, C.POINTER(sz), vp]
    L.sgx_irfft.argtypes = [i, vp, sz, sz, vp, i, vp]
    L.sgx_plan_compute_binaural.argtypes = [vp, i, vp, vp, sz, sz, sz, d, d, sz, i, vp, sz, sz, vp]
    L.sgx_fft_planner_create.argtypes = [i, C.POINTER(vp)]
    L.sgx_fft_planner_destroy.argtypes = [vp]
    L.sgx_fft_planner_cached_plans.argtypes = [vp]
    L.sgx_fft_planner_cached_plans.restype = sz
    L.sgx_fft_planner_rfft.argtypes = [vp, i, vp, sz, sz, vp, vp]
    L.sgx_fft_planner_irfft.argtypes = [vp, i, vp, sz, sz, vp, vp]
    L.sgx_fft_planner_power_spectrum.argtypes = [vp, i, vp, sz, sz, i, d, i, vp, vp]
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def check(status: int) -> None:
    """Map an ``sgx_status`` to the reference's exception types (src/python/error.rs:50-66)."""
    if status == SGX_OK:
        return
    L = lib()
    msg = L.sgx_last_error_message().decode()
    if status == SGX_INVALID_INPUT:
        raise InvalidInputError(msg)
    if status == SGX_DIMENSION_MISMATCH:
        e, g = C.c_size_t(), C.c_size_t()
        L.sgx_last_dimension_mismatch(C.byref(e), C.byref(g))
        raise DimensionMismatchError(msg, e.value, g.value)
    if status == SGX_BACKEND_ERROR:
        raise FFTBackendError(msg)
    raise InternalError(msg)
