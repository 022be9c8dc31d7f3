"""CPU oracle for the jmg049/Spectrograms hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package. The product package ``spectrograms_b200`` never does (tests/test_boundary.py checks that).

Two independent restatements of the reference algorithm live here:

* ``oracle.c`` (+ ``oracle_impl.inc``): line-by-line C restatement in native f32 and f64, with a one-plan-per-thread
  batch driver (the reference's documented scaling recipe) used as the CPU baseline. Loaded through ctypes below.
* ``oracle_np.py``: a NumPy/SciPy restatement (pocketfft) used to cross-check the C code.

Parity pinning: see the header of ``oracle.c``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

F32, F64 = 0, 1
MAP_IDENTITY, MAP_MEL, MAP_ERB, MAP_LOGHZ = 0, 1, 2, 3
AMP_POWER, AMP_MAGNITUDE, AMP_DECIBELS = 0, 1, 2
WIN_RECT, WIN_HANN, WIN_HAMMING, WIN_BLACKMAN, WIN_KAISER, WIN_GAUSSIAN, WIN_CUSTOM = range(7)
MELNORM_NONE, MELNORM_SLANEY, MELNORM_L1, MELNORM_L2 = range(4)
ERB_LINEAR, ERB_APPLE_TR35 = 0, 1


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_impl.inc", "oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Desc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("n_fft", C.c_size_t), ("hop", C.c_size_t), ("centre", C.c_int),
        ("window_kind", C.c_int), ("window_param", C.c_double),
        ("custom_window", C.POINTER(C.c_double)), ("custom_window_len", C.c_size_t),
        ("sample_rate", C.c_double), ("mapping", C.c_int), ("n_bands", C.c_size_t),
        ("f_min", C.c_double), ("f_max", C.c_double), ("mel_norm", C.c_int), ("erb_spacing", C.c_int),
        ("amp", C.c_int), ("has_floor_db", C.c_int), ("floor_db", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_frame_count.restype = C.c_size_t
        L.orc_frame_count.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_last_error.restype = C.c_char_p
        L.orc_plan_create.restype = C.c_void_p
        L.orc_plan_create.argtypes = [C.POINTER(_Desc)]
        L.orc_plan_destroy.argtypes = [C.c_void_p]
        L.orc_plan_n_bins.restype = C.c_size_t
        L.orc_plan_n_bins.argtypes = [C.c_void_p]
        L.orc_plan_out_len.restype = C.c_size_t
        L.orc_plan_out_len.argtypes = [C.c_void_p]
        L.orc_plan_window.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_plan_freq_axis.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_plan_filterbank_nnz.restype = C.c_size_t
        L.orc_plan_filterbank_nnz.argtypes = [C.c_void_p]
        L.orc_plan_filterbank_dense.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_compute_spectrogram.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_compute_stft.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_compute_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.orc_mfcc_from_log_mel.restype = C.c_int
        L.orc_mfcc_from_log_mel.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                            C.c_size_t, C.c_int, C.c_void_p]
        L.orc_chroma_filterbank.restype = C.c_int
        L.orc_chroma_filterbank.argtypes = [C.c_double, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orc_chroma_from_spectrogram.restype = C.c_int
        L.orc_chroma_from_spectrogram.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_double, C.c_size_t,
                                                  C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.orc_binaural_from_stft.restype = C.c_int
        L.orc_binaural_from_stft.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                             C.c_size_t, C.c_double, C.c_size_t, C.c_int, C.c_void_p]
        L.orc_irfft.restype = C.c_int
        L.orc_irfft.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_istft.restype = C.c_size_t
        L.orc_istft.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_rfft.restype = C.c_int
        L.orc_rfft.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.orc_compute_batch.restype = C.c_int
        L.orc_compute_batch.argtypes = [C.POINTER(_Desc), C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.c_int]
        _lib = L
    return _lib


class OracleError(ValueError):
    pass


@dataclass
class Desc:
    """Flat description of one plan (mirrors the fields of StftParams/SpectrogramParams/MelParams/... )."""
    dtype: str = "f64"
    n_fft: int = 512
    hop: int = 256
    centre: bool = True
    window: str = "hanning"            # rectangular|hanning|hamming|blackman|kaiser|gaussian|custom
    window_param: float = 0.0
    custom_window: Optional[Sequence[float]] = None
    sample_rate: float = 16000.0
    mapping: str = "linear"            # linear|mel|erb|loghz
    n_bands: int = 0
    f_min: float = 0.0
    f_max: float = 0.0
    mel_norm: str = "none"             # none|slaney|l1|l2
    erb_spacing: str = "linear"        # linear|apple_tr35
    amp: str = "power"                 # power|magnitude|db
    floor_db: Optional[float] = None
    _keep: list = field(default_factory=list, repr=False)

    def np_dtype(self):
        return np.float32 if self.dtype == "f32" else np.float64

    def to_c(self) -> _Desc:
        d = _Desc()
        d.dtype = F32 if self.dtype == "f32" else F64
        d.n_fft, d.hop, d.centre = self.n_fft, self.hop, int(self.centre)
        d.window_kind = {"rectangular": 0, "hanning": 1, "hamming": 2, "blackman": 3, "kaiser": 4, "gaussian": 5,
                         "custom": 6}[self.window]
        d.window_param = float(self.window_param)
        if self.custom_window is not None:
            arr = np.ascontiguousarray(self.custom_window, dtype=np.float64)
            self._keep.append(arr)
            d.custom_window = arr.ctypes.data_as(C.POINTER(C.c_double))
            d.custom_window_len = arr.size
        d.sample_rate = float(self.sample_rate)
        d.mapping = {"linear": 0, "mel": 1, "erb": 2, "loghz": 3}[self.mapping]
        d.n_bands = int(self.n_bands)
        d.f_min, d.f_max = float(self.f_min), float(self.f_max)
        d.mel_norm = {"none": 0, "slaney": 1, "l1": 2, "l2": 3}[self.mel_norm]
        d.erb_spacing = {"linear": 0, "apple_tr35": 1}[self.erb_spacing]
        d.amp = {"power": 0, "magnitude": 1, "db": 2}[self.amp]
        d.has_floor_db = int(self.floor_db is not None)
        d.floor_db = float(self.floor_db) if self.floor_db is not None else 0.0
        return d


def frame_count(n_samples: int, n_fft: int, hop: int, centre: bool) -> int:
    return int(lib().orc_frame_count(n_samples, n_fft, hop, int(centre)))


class Plan:
    """One reference plan (SpectrogramPlan / StftPlan) on the CPU."""

    def __init__(self, desc: Desc):
        self.desc = desc
        self._cd = desc.to_c()
        self._h = lib().orc_plan_create(C.byref(self._cd))
        if not self._h:
            raise OracleError(lib().orc_last_error().decode())
        self.n_bins = int(lib().orc_plan_n_bins(self._h))
        self.out_len = int(lib().orc_plan_out_len(self._h))
        self.dt = desc.np_dtype()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().orc_plan_destroy(h)

    def n_frames(self, n_samples: int) -> int:
        return frame_count(n_samples, self.desc.n_fft, self.desc.hop, self.desc.centre)

    def window(self) -> np.ndarray:
        w = np.empty(self.desc.n_fft, dtype=self.dt)
        lib().orc_plan_window(self._h, w.ctypes.data)
        return w

    def freq_axis(self) -> np.ndarray:
        f = np.empty(self.n_bins, dtype=np.float64)
        lib().orc_plan_freq_axis(self._h, f.ctypes.data)
        return f

    def times(self, n_frames: int) -> np.ndarray:
        # build_time_axis_seconds (src/spectrogram.rs:2128-2139): i * (hop / sr)
        dt = float(self.desc.hop) / float(self.desc.sample_rate)
        return np.arange(n_frames, dtype=np.float64) * dt

    def filterbank_nnz(self) -> int:
        return int(lib().orc_plan_filterbank_nnz(self._h))

    def filterbank_dense(self) -> np.ndarray:
        m = np.empty((self.n_bins, self.out_len), dtype=np.float64)
        lib().orc_plan_filterbank_dense(self._h, m.ctypes.data)
        return m

    def _samples(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=self.dt)
        if x.ndim != 1 or x.size == 0:
            raise OracleError("samples must be a non-empty 1-D array")
        return x

    def compute(self, x) -> np.ndarray:
        x = self._samples(x)
        out = np.empty((self.n_bins, self.n_frames(x.size)), dtype=self.dt)
        lib().orc_compute_spectrogram(self._h, x.ctypes.data, x.size, out.ctypes.data)
        return out

    def stft(self, x) -> np.ndarray:
        x = self._samples(x)
        cdt = np.complex64 if self.dt == np.float32 else np.complex128
        out = np.empty((self.out_len, self.n_frames(x.size)), dtype=cdt)
        lib().orc_compute_stft(self._h, x.ctypes.data, x.size, out.ctypes.data)
        return out

    def compute_frame(self, x, frame_idx: int) -> np.ndarray:
        x = self._samples(x)
        out = np.empty(self.n_bins, dtype=self.dt)
        lib().orc_compute_frame(self._h, x.ctypes.data, x.size, frame_idx, out.ctypes.data)
        return out


def mfcc_from_log_mel(log_mel: np.ndarray, n_mfcc: int, include_c0: bool = True, lifter: int = 22,
                      faithful: bool = True) -> np.ndarray:
    log_mel = np.ascontiguousarray(log_mel)
    assert log_mel.dtype in (np.float32, np.float64) and log_mel.ndim == 2
    n_mels, n_frames = log_mel.shape
    rows = n_mfcc - 1 if (not include_c0 and n_mfcc > 1) else n_mfcc
    out = np.empty((max(rows, 0), n_frames), dtype=log_mel.dtype)
    rc = lib().orc_mfcc_from_log_mel(F32 if log_mel.dtype == np.float32 else F64, log_mel.ctypes.data, n_mels,
                                     n_frames, n_mfcc, int(include_c0), lifter, int(faithful), out.ctypes.data)
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    return out


CHROMA_NORMS = {"none": 0, "l1": 1, "l2": 2, "max": 3}


def chroma_filterbank(sample_rate: float, n_fft: int, tuning: float = 440.0, f_min: float = 32.7, f_max: float = 4186.0) -> np.ndarray:
    """build_chroma_filterbank (src/chroma.rs:279-346): (12, n_fft//2 + 1) f64."""
    out = np.empty((12, n_fft // 2 + 1), dtype=np.float64)
    if lib().orc_chroma_filterbank(sample_rate, n_fft, tuning, f_min, f_max, out.ctypes.data):
        raise OracleError(lib().orc_last_error().decode())
    return out


def chroma_from_spectrogram(spec: np.ndarray, sample_rate: float, n_fft: int, tuning: float = 440.0, f_min: float = 32.7,
                            f_max: float = 4186.0, norm: str = "l2") -> np.ndarray:
    """chromagram_from_spectrogram (src/chroma.rs:365-404): (n_bins, n_frames) -> (12, n_frames), same dtype."""
    spec = np.ascontiguousarray(spec)
    assert spec.dtype in (np.float32, np.float64) and spec.ndim == 2
    out = np.empty((12, spec.shape[1]), dtype=spec.dtype)
    rc = lib().orc_chroma_from_spectrogram(F32 if spec.dtype == np.float32 else F64, spec.ctypes.data, spec.shape[0], spec.shape[1],
                                           sample_rate, n_fft, tuning, f_min, f_max, CHROMA_NORMS[norm], out.ctypes.data)
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    return out


def chromagram(x: np.ndarray, n_fft: int, hop: int, sample_rate: float, window: str = "hanning", centre: bool = True,
               tuning: float = 440.0, f_min: float = 32.7, f_max: float = 4186.0, norm: str = "l2") -> np.ndarray:
    """chromagram() (src/chroma.rs:487-503): Spectrogram::<LinearHz, Magnitude, T>::compute, then the function above."""
    x = np.ascontiguousarray(x)
    d = Desc(dtype="f32" if x.dtype == np.float32 else "f64", n_fft=n_fft, hop=hop, sample_rate=sample_rate, window=window,
             centre=centre, mapping="linear", amp="magnitude")
    return chroma_from_spectrogram(Plan(d).compute(x), sample_rate, n_fft, tuning, f_min, f_max, norm)


CUES = {"itd": 0, "ipd": 1, "ild": 2, "ilr": 3}


def binaural_band(start_freq: float, end_freq: float, sample_rate: float, n_fft: int):
    """(start_bin, stop_bin, bin_width) of src/binaural.rs:476-481: Rust `f64::round` (half away from zero) `as usize`."""
    bw = sample_rate / n_fft
    rnd = lambda v: int(np.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))
    return max(rnd(start_freq / bw), 0), max(rnd(end_freq / bw), 0), bw


def binaural_from_stft(cue: str, left: np.ndarray, right: np.ndarray, start_bin: int, stop_bin: int, bin_width: float,
                       magphase_power: int = 1, wrapped: bool = True) -> np.ndarray:
    """Element-wise half of compute_{itd,ipd,ild,ilr}_spectrogram on two complex (n_bins, n_frames) STFTs."""
    left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
    assert left.dtype == right.dtype and left.dtype in (np.complex64, np.complex128) and left.shape == right.shape
    rdt = np.float32 if left.dtype == np.complex64 else np.float64
    out = np.empty((stop_bin - start_bin, left.shape[1]), dtype=rdt)
    rc = lib().orc_binaural_from_stft(F32 if rdt == np.float32 else F64, CUES[cue], left.ctypes.data, right.ctypes.data,
                                      left.shape[0], left.shape[1], start_bin, stop_bin, bin_width, magphase_power,
                                      int(wrapped), out.ctypes.data)
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    return out


def binaural(cue: str, left: np.ndarray, right: np.ndarray, n_fft: int, hop: int, sample_rate: float, start_freq: float,
             end_freq: float, window: str = "hanning", centre: bool = True, magphase_power: int = 1, wrapped: bool = True):
    """compute_{itd,ipd,ild,ilr}_spectrogram: StftPlan::compute of both channels, then the function above."""
    d = Desc(dtype="f32" if left.dtype == np.float32 else "f64", n_fft=n_fft, hop=hop, sample_rate=sample_rate, window=window, centre=centre)
    p = Plan(d)
    b0, b1, bw = binaural_band(start_freq, end_freq, sample_rate, n_fft)
    return binaural_from_stft(cue, p.stft(left), p.stft(right), b0, b1, bw, magphase_power, wrapped)


def irfft(spectrum: np.ndarray, n_fft: int) -> np.ndarray:
    """irfft (src/spectrogram.rs:4789-4811)."""
    spectrum = np.ascontiguousarray(spectrum)
    assert spectrum.dtype in (np.complex64, np.complex128)
    if spectrum.size != n_fft // 2 + 1:
        raise OracleError(f"Dimension mismatch: expected {n_fft // 2 + 1}, got {spectrum.size}")
    out = np.empty(n_fft, dtype=np.float32 if spectrum.dtype == np.complex64 else np.float64)
    lib().orc_irfft(F32 if spectrum.dtype == np.complex64 else F64, spectrum.ctypes.data, n_fft, out.ctypes.data)
    return out


def istft(stft_matrix: np.ndarray, n_fft: int, hop: int, window: str = "hanning", centre: bool = True, window_param: float = 0.0) -> np.ndarray:
    """istft (src/spectrogram.rs:4813-4911): (n_fft//2 + 1, n_frames) complex -> samples."""
    stft_matrix = np.ascontiguousarray(stft_matrix)
    assert stft_matrix.dtype in (np.complex64, np.complex128) and stft_matrix.ndim == 2
    if stft_matrix.shape[0] != n_fft // 2 + 1:
        raise OracleError(f"Dimension mismatch: expected {n_fft // 2 + 1}, got {stft_matrix.shape[0]}")
    rdt = np.float32 if stft_matrix.dtype == np.complex64 else np.float64
    p = Plan(Desc(dtype="f32" if rdt == np.float32 else "f64", n_fft=n_fft, hop=hop, window=window, window_param=window_param, centre=centre))
    n = lib().orc_istft(p._h, stft_matrix.ctypes.data, stft_matrix.shape[1], None)
    out = np.empty(n, dtype=rdt)
    lib().orc_istft(p._h, stft_matrix.ctypes.data, stft_matrix.shape[1], out.ctypes.data)
    return out


def rfft(x: np.ndarray, n_fft: int) -> np.ndarray:
    x = np.ascontiguousarray(x)
    assert x.dtype in (np.float32, np.float64)
    out = np.empty(n_fft // 2 + 1, dtype=np.complex64 if x.dtype == np.float32 else np.complex128)
    rc = lib().orc_rfft(F32 if x.dtype == np.float32 else F64, x.ctypes.data, x.size, n_fft, out.ctypes.data)
    if rc:
        raise OracleError(lib().orc_last_error().decode())
    return out


def compute_batch(desc: Desc, clips: np.ndarray, n_threads: int = 1, mfcc: Optional[dict] = None) -> np.ndarray:
    """clips: (n_clips, n_samples). One private plan per worker thread, clips sharded across workers."""
    clips = np.ascontiguousarray(clips, dtype=desc.np_dtype())
    n_clips, n_samples = clips.shape
    p = Plan(desc)
    nf = p.n_frames(n_samples)
    rows = p.n_bins
    if mfcc is not None:
        rows = mfcc["n_mfcc"] - 1 if (not mfcc.get("include_c0", True) and mfcc["n_mfcc"] > 1) else mfcc["n_mfcc"]
    out = np.empty((n_clips, rows, nf), dtype=desc.np_dtype())
    cd = desc.to_c()
    rc = lib().orc_compute_batch(C.byref(cd), clips.ctypes.data, n_clips, n_samples, n_samples, out.ctypes.data,
                                 rows * nf, n_threads, int(mfcc is not None),
                                 (mfcc or {}).get("n_mfcc", 0), int((mfcc or {}).get("include_c0", True)),
                                 (mfcc or {}).get("lifter", 22), int((mfcc or {}).get("faithful", True)))
    if rc:
        raise OracleError("oracle batch failed")
    return out
