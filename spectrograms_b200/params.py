"""Parameter types of the hot path, mirroring the reference's constructors, validation and error messages.

Bare ``:N`` citations are ``src/spectrogram.rs:N`` of the reference checkout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .errors import InvalidInputError


@dataclass(frozen=True)
class WindowType:
    """``WindowType`` (src/window.rs:19-50)."""
    kind: str = "hanning"
    param: float = 0.0
    coefficients: Optional[tuple] = None

    _KINDS = ("rectangular", "hanning", "hamming", "blackman", "kaiser", "gaussian", "custom")

    def __post_init__(self):
        if self.kind not in self._KINDS:
            raise InvalidInputError(f"Unknown window type '{self.kind}'")

    # constructors named like the reference's variants / Python classmethods
    @classmethod
    def rectangular(cls): return cls("rectangular")
    @classmethod
    def hanning(cls): return cls("hanning")
    @classmethod
    def hamming(cls): return cls("hamming")
    @classmethod
    def blackman(cls): return cls("blackman")
    @classmethod
    def kaiser(cls, beta: float): return cls("kaiser", float(beta))
    @classmethod
    def gaussian(cls, std: float): return cls("gaussian", float(std))

    @classmethod
    def custom(cls, coefficients: Sequence[float], normalize: Optional[str] = None) -> "WindowType":
        """``WindowType::custom_with_normalization`` (src/window.rs:134-203)."""
        c = [float(v) for v in coefficients]
        if len(c) == 0:
            raise InvalidInputError("Custom window coefficients cannot be empty")
        for i, v in enumerate(c):
            if not math.isfinite(v):
                raise InvalidInputError(f"Window coefficient at index {i} is not finite: {v}")
        if normalize is not None:
            if normalize == "sum":
                s = math.fsum(c) if False else sum(c)
                if s == 0.0:
                    raise InvalidInputError("Cannot normalize window by sum: sum is zero")
                c = [v / s for v in c]
            elif normalize in ("peak", "max"):
                m = max(c)
                if m == 0.0:
                    raise InvalidInputError("Cannot normalize window by peak: maximum is zero")
                c = [v / m for v in c]
            elif normalize in ("energy", "rms"):
                e = sum(v * v for v in c)
                if e == 0.0:
                    raise InvalidInputError("Cannot normalize window by energy: energy is zero")
                n = math.sqrt(e)
                c = [v / n for v in c]
            else:
                raise InvalidInputError(
                    f"Unknown normalization mode '{normalize}'. Valid modes: 'sum', 'peak', 'energy'")
        return cls("custom", 0.0, tuple(c))

    @classmethod
    def from_str(cls, s: str) -> "WindowType":
        """``FromStr`` (src/window.rs:276-338): names plus ``kaiser=<beta>`` / ``gaussian=<std>``."""
        t = s.strip().lower()
        simple = {"rectangular": "rectangular", "rect": "rectangular", "hanning": "hanning", "hann": "hanning",
                  "hamming": "hamming", "blackman": "blackman"}
        if t in simple:
            return cls(simple[t])
        for name in ("kaiser", "gaussian"):
            for sep in ("=", ":", "("):
                if t.startswith(name + sep):
                    body = t[len(name) + 1:].rstrip(")")
                    if "=" in body:
                        body = body.split("=", 1)[1]
                    try:
                        return cls(name, float(body))
                    except ValueError as e:
                        raise InvalidInputError(f"Invalid window parameter in '{s}'") from e
        raise InvalidInputError(f"Unknown window type '{s}'")

    def is_parameterized(self) -> bool:
        return self.kind in ("kaiser", "gaussian")

    def parameter_value(self) -> Optional[float]:
        return self.param if self.is_parameterized() else None

    def size(self) -> Optional[int]:
        return len(self.coefficients) if self.kind == "custom" else None

    def __str__(self) -> str:
        if self.kind == "kaiser":
            return f"Kaiser(beta={self.param})"
        if self.kind == "gaussian":
            return f"Gaussian(std={self.param})"
        if self.kind == "custom":
            return f"Custom(n={len(self.coefficients)})"
        return self.kind.capitalize()


def _as_window(w) -> WindowType:
    if isinstance(w, WindowType):
        return w
    if isinstance(w, str):
        return WindowType.from_str(w)
    raise InvalidInputError("window must be a WindowType or str")


def _nonzero(name: str, v: int) -> int:
    if not isinstance(v, (int, np.integer)) or int(v) <= 0:
        raise InvalidInputError(f"{name} must be set")        # NonZeroUsize
    return int(v)


class StftParams:
    """``StftParams::new`` (:3479-3506)."""

    def __init__(self, n_fft: int, hop_size: int, window="hanning", centre: bool = True):
        self._n_fft = _nonzero("n_fft", n_fft)
        self._hop = _nonzero("hop_size", hop_size)
        self._window = _as_window(window)
        self._centre = bool(centre)
        if self._hop > self._n_fft:
            raise InvalidInputError("hop_size must be <= n_fft")
        if self._window.kind == "custom" and self._window.size() != self._n_fft:
            raise InvalidInputError(
                f"Custom window size ({self._window.size()}) must match n_fft ({self._n_fft})")

    n_fft = property(lambda s: s._n_fft)
    hop_size = property(lambda s: s._hop)
    window = property(lambda s: s._window)
    centre = property(lambda s: s._centre)

    def __eq__(self, o):
        return isinstance(o, StftParams) and (self._n_fft, self._hop, self._window, self._centre) == (
            o._n_fft, o._hop, o._window, o._centre)

    def __repr__(self):
        return f"StftParams(n_fft={self._n_fft}, hop_size={self._hop}, window={self._window}, centre={self._centre})"


class SpectrogramParams:
    """``SpectrogramParams::new`` (:4129-4140), presets (:4215-4248), derived values (:4268-4277)."""

    def __init__(self, stft: StftParams, sample_rate: float):
        sr = float(sample_rate)
        if not (sr > 0.0 and math.isfinite(sr)):
            raise InvalidInputError("sample_rate_hz must be finite and > 0")
        self._stft = stft
        self._sr = sr

    stft = property(lambda s: s._stft)
    sample_rate = property(lambda s: s._sr)
    sample_rate_hz = property(lambda s: s._sr)

    @classmethod
    def speech_default(cls, sample_rate: float) -> "SpectrogramParams":
        return cls(StftParams(512, 160, WindowType.hanning(), True), sample_rate)

    @classmethod
    def music_default(cls, sample_rate: float) -> "SpectrogramParams":
        return cls(StftParams(2048, 512, WindowType.hanning(), True), sample_rate)

    def frame_period_seconds(self) -> float:
        return float(self._stft.hop_size) / self._sr

    def nyquist_hz(self) -> float:
        return self._sr * 0.5

    def __repr__(self):
        return f"SpectrogramParams({self._stft!r}, sample_rate={self._sr})"


class MelNorm:
    """``MelNorm`` (:3708-3734)."""
    NONE, SLANEY, L1, L2 = "none", "slaney", "l1", "l2"
    ALL = (NONE, SLANEY, L1, L2)


class MelParams:
    """``MelParams::with_norm`` (:3793-3812)."""

    def __init__(self, n_mels: int, f_min: float, f_max: float, norm: str = MelNorm.NONE):
        self._n = _nonzero("n_mels", n_mels)
        f_min, f_max = float(f_min), float(f_max)
        if f_min < 0.0:
            raise InvalidInputError("f_min must be >= 0")
        if not (f_max > f_min):
            raise InvalidInputError("f_max must be > f_min")
        norm = (norm or "none").lower() if isinstance(norm, str) else norm
        if norm not in MelNorm.ALL:
            raise InvalidInputError(f"Unknown mel normalization '{norm}'")
        self._f_min, self._f_max, self._norm = f_min, f_max, norm

    n_mels = property(lambda s: s._n)
    f_min = property(lambda s: s._f_min)
    f_max = property(lambda s: s._f_max)
    norm = property(lambda s: s._norm)

    @classmethod
    def speech_standard(cls) -> "MelParams":
        return cls(40, 0.0, 8000.0)


class ErbParams:
    """``ErbParams::new`` (src/erb.rs:66-88) + ``with_spacing`` (:93-95)."""

    def __init__(self, n_filters: int, f_min: float, f_max: float, spacing: str = "linear"):
        n = _nonzero("n_filters", n_filters)
        f_min, f_max = float(f_min), float(f_max)
        if n < 2:
            raise InvalidInputError("n_filters must be >= 2 (single filter would cause division by zero)")
        if f_min < 0.0 or math.isinf(f_min):
            raise InvalidInputError("f_min must be finite and >= 0")
        if not (f_max > f_min):
            raise InvalidInputError("f_max must be > f_min")
        if spacing not in ("linear", "apple_tr35"):
            raise InvalidInputError(f"Unknown ERB spacing '{spacing}'")
        self._n, self._f_min, self._f_max, self._spacing = n, f_min, f_max, spacing

    n_filters = property(lambda s: s._n)
    f_min = property(lambda s: s._f_min)
    f_max = property(lambda s: s._f_max)
    spacing = property(lambda s: s._spacing)

    def with_spacing(self, spacing: str) -> "ErbParams":
        return ErbParams(self._n, self._f_min, self._f_max, spacing)

    @classmethod
    def speech_standard(cls) -> "ErbParams":
        return cls(40, 0.0, 8000.0)

    @classmethod
    def music_standard(cls, sample_rate: float) -> "ErbParams":
        return cls(64, 0.0, float(sample_rate) / 2.0)


GammatoneParams = ErbParams


class LogHzParams:
    """``LogHzParams::new`` (:3960-3975)."""

    def __init__(self, n_bins: int, f_min: float, f_max: float):
        self._n = _nonzero("n_bins", n_bins)
        f_min, f_max = float(f_min), float(f_max)
        if not (f_min > 0.0 and math.isfinite(f_min)):
            raise InvalidInputError("f_min must be finite and > 0")
        if not (f_max > f_min):
            raise InvalidInputError("f_max must be > f_min")
        self._f_min, self._f_max = f_min, f_max

    n_bins = property(lambda s: s._n)
    f_min = property(lambda s: s._f_min)
    f_max = property(lambda s: s._f_max)


class LogParams:
    """``LogParams::new`` (:4071-4076): only ``floor_db`` exists -- there is no ``ref`` and no ``top_db``."""

    def __init__(self, floor_db: float):
        floor_db = float(floor_db)
        if not math.isfinite(floor_db):
            raise InvalidInputError("floor_db must be finite")
        self._floor = floor_db

    floor_db = property(lambda s: s._floor)


class MfccParams:
    """``MfccParams`` (src/mfcc.rs:21-141): defaults n_mfcc=13, include_c0=True, lifter=22."""

    def __init__(self, n_mfcc: int = 13, include_c0: bool = True, lifter: int = 22):
        self._n = _nonzero("n_mfcc", n_mfcc)
        self._c0 = bool(include_c0)
        if int(lifter) < 0:
            raise InvalidInputError("lifter must be >= 0")
        self._lifter = int(lifter)

    n_mfcc = property(lambda s: s._n)
    include_c0 = property(lambda s: s._c0)
    lifter = property(lambda s: s._lifter)

    @classmethod
    def speech_standard(cls) -> "MfccParams":
        return cls(13)

    def with_c0(self, include_c0: bool) -> "MfccParams":
        return MfccParams(self._n, include_c0, self._lifter)

    def with_lifter(self, lifter: int) -> "MfccParams":
        return MfccParams(self._n, self._c0, lifter)


class ChromaNorm:
    """``ChromaNorm`` (src/chroma.rs:31-45): none | l1 | l2 (default) | max."""
    NONE, L1, L2, MAX = "none", "l1", "l2", "max"
    ALL = ("none", "l1", "l2", "max")


class ChromaParams:
    """``ChromaParams`` (src/chroma.rs:18-182): defaults tuning=440, f_min=32.7 (C1), f_max=4186 (C8), norm=L2."""

    def __init__(self, tuning: float = 440.0, f_min: float = 32.7, f_max: float = 4186.0, norm: Optional[str] = None):
        tuning, f_min, f_max = float(tuning), float(f_min), float(f_max)
        if not (tuning > 0.0 and math.isfinite(tuning)):                      # :82-86
            raise InvalidInputError("tuning must be finite and > 0")
        if not (f_min > 0.0 and math.isfinite(f_min)):                        # :87-91
            raise InvalidInputError("f_min must be finite and > 0")
        if f_max <= f_min:                                                    # :92-94
            raise InvalidInputError("f_max must be > f_min")
        norm = ChromaNorm.L2 if norm is None else str(norm).lower()
        if norm not in ChromaNorm.ALL:
            raise InvalidInputError("norm must be one of none, l1, l2, max")
        self._tuning, self._f_min, self._f_max, self._norm = tuning, f_min, f_max, norm
        self._n_octaves = max(int(math.ceil(math.log2(f_max / f_min))), 1)    # :97

    tuning = property(lambda s: s._tuning)
    f_min = property(lambda s: s._f_min)
    f_max = property(lambda s: s._f_max)
    norm = property(lambda s: s._norm)
    n_octaves = property(lambda s: s._n_octaves)

    @classmethod
    def music_standard(cls) -> "ChromaParams":
        p = cls(440.0, 32.7, 4186.0, ChromaNorm.L2)
        p._n_octaves = 7                                                      # the const constructor's literal (:114-122)
        return p

    def with_norm(self, norm: str) -> "ChromaParams":
        p = ChromaParams(self._tuning, self._f_min, self._f_max, norm)
        p._n_octaves = self._n_octaves
        return p

    def __repr__(self) -> str:
        return f"ChromaParams(tuning={self._tuning}, f_min={self._f_min}, f_max={self._f_max}, norm={self._norm})"


def normalise_dtype(dtype) -> str:
    """dtype strings of the reference's Python layer (src/python/dtype.rs:34-42)."""
    if dtype in ("float32", "f32", np.float32) or (hasattr(dtype, "name") and getattr(dtype, "name", "") == "float32"):
        return "f32"
    if dtype in ("float64", "f64", np.float64, float, None) or getattr(dtype, "name", "") == "float64":
        return "f64"
    s = str(dtype)
    if s.endswith("float32"):
        return "f32"
    if s.endswith("float64"):
        return "f64"
    raise InvalidInputError(f"Unsupported dtype '{dtype}': expected 'float32' or 'float64'")
