"""serde wire format of the crate's result / parameter types.

The reference derives ``serde::Serialize / Deserialize`` on ``Spectrogram`` (src/spectrogram.rs:2546-2557: fields ``data``,
``axes``, ``params``; ``_amp`` is ``#[serde(skip)]``), ``Axes`` (:3305: ``freq``, ``times``), ``FrequencyAxis`` (:3242:
``frequencies``), ``SpectrogramParams`` (:4107: ``stft``, ``sample_rate_hz``), ``StftParams`` (:4051: ``n_fft``, ``hop_size``,
``window``, ``centre``), ``WindowType`` (src/window.rs:17: externally tagged enum -- ``"Hanning"``, ``{"Kaiser":{"beta":b}}``,
``{"Custom":{"coefficients":[..],"size":n}}``), ``MelParams`` / ``MelNorm`` (:3705-3760), ``LogHzParams`` (:3934), ``LogParams``
(:3451), ``ErbParams`` / ``ErbSpacing`` (src/erb.rs:15-40), ``MfccParams`` / ``Mfcc`` (src/mfcc.rs:20, :145) and
``ChromaParams`` / ``ChromaNorm`` / ``Chromagram`` (src/chroma.rs:17-45, :187). ``Array2<T>`` uses ndarray's serde format
``{"v":1,"dim":[rows,cols],"data":[row-major values]}``. The functions below produce / parse exactly those JSON objects, so
that a document written by ``serde_json::to_string(&spec)`` (tests/serde_tests.rs:45-65) loads here and vice versa. The
frequency-scale / amplitude-scale markers are type parameters in Rust and are not on the wire: ``spectrogram_from_dict``
takes them as arguments, like ``serde_json::from_str::<MelPowerSpectrogram>`` does.

(No Rust toolchain exists in this image, so the field names are taken from the derive input, not from a produced file.)
"""
from __future__ import annotations

import json
from typing import Any, Dict, Optional

import numpy as np

from .errors import InvalidInputError
from .params import (ChromaParams, ErbParams, LogHzParams, LogParams, MelParams, MfccParams, SpectrogramParams, StftParams,
                     WindowType)

_MELNORM = {"none": "None", "slaney": "Slaney", "l1": "L1", "l2": "L2"}
_CHROMANORM = {"none": "None", "l1": "L1", "l2": "L2", "max": "Max"}
_ERBSPACING = {"linear": "Linear", "apple_tr35": "AppleTr35"}
_UNIT_WINDOWS = {"rectangular": "Rectangular", "hanning": "Hanning", "hamming": "Hamming", "blackman": "Blackman"}


def _inv(m: Dict[str, str], v: str, what: str) -> str:
    for k, name in m.items():
        if name == v:
            return k
    raise InvalidInputError(f"unknown {what} variant '{v}'")


# ------------------------------------------------------------------------------------------------ parameters
def window_to_obj(w: WindowType) -> Any:
    if w.kind in _UNIT_WINDOWS:
        return _UNIT_WINDOWS[w.kind]
    if w.kind == "kaiser":
        return {"Kaiser": {"beta": float(w.param)}}
    if w.kind == "gaussian":
        return {"Gaussian": {"std": float(w.param)}}
    return {"Custom": {"coefficients": [float(c) for c in w.coefficients], "size": len(w.coefficients)}}


def window_from_obj(o: Any) -> WindowType:
    if isinstance(o, str):
        return WindowType(_inv(_UNIT_WINDOWS, o, "WindowType"))
    if isinstance(o, dict) and len(o) == 1:
        (tag, body), = o.items()
        if tag == "Kaiser":
            return WindowType.kaiser(body["beta"])
        if tag == "Gaussian":
            return WindowType.gaussian(body["std"])
        if tag == "Custom":
            if int(body["size"]) != len(body["coefficients"]):
                raise InvalidInputError("Custom window: size does not match the number of coefficients")
            return WindowType.custom(body["coefficients"])
    raise InvalidInputError("malformed WindowType")


def stft_params_to_dict(p: StftParams) -> dict:
    return {"n_fft": int(p.n_fft), "hop_size": int(p.hop_size), "window": window_to_obj(p.window), "centre": bool(p.centre)}


def stft_params_from_dict(d: dict) -> StftParams:
    return StftParams(int(d["n_fft"]), int(d["hop_size"]), window_from_obj(d["window"]), bool(d["centre"]))


def spectrogram_params_to_dict(p: SpectrogramParams) -> dict:
    return {"stft": stft_params_to_dict(p.stft), "sample_rate_hz": float(p.sample_rate)}


def spectrogram_params_from_dict(d: dict) -> SpectrogramParams:
    return SpectrogramParams(stft_params_from_dict(d["stft"]), float(d["sample_rate_hz"]))


def mel_params_to_dict(p: MelParams) -> dict:
    return {"n_mels": int(p.n_mels), "f_min": float(p.f_min), "f_max": float(p.f_max), "norm": _MELNORM[p.norm]}


def mel_params_from_dict(d: dict) -> MelParams:
    return MelParams(int(d["n_mels"]), float(d["f_min"]), float(d["f_max"]), _inv(_MELNORM, d["norm"], "MelNorm"))


def loghz_params_to_dict(p: LogHzParams) -> dict:
    return {"n_bins": int(p.n_bins), "f_min": float(p.f_min), "f_max": float(p.f_max)}


def loghz_params_from_dict(d: dict) -> LogHzParams:
    return LogHzParams(int(d["n_bins"]), float(d["f_min"]), float(d["f_max"]))


def erb_params_to_dict(p: ErbParams, db_floor: Optional[float] = None) -> dict:
    return {"n_filters": int(p.n_filters), "f_min": float(p.f_min), "f_max": float(p.f_max), "spacing": _ERBSPACING[p.spacing],
            "db_floor": None if db_floor is None else float(db_floor)}


def erb_params_from_dict(d: dict) -> ErbParams:
    return ErbParams(int(d["n_filters"]), float(d["f_min"]), float(d["f_max"]), _inv(_ERBSPACING, d["spacing"], "ErbSpacing"))


def log_params_to_dict(p: LogParams) -> dict:
    return {"floor_db": float(p.floor_db)}


def log_params_from_dict(d: dict) -> LogParams:
    return LogParams(float(d["floor_db"]))


def mfcc_params_to_dict(p: MfccParams) -> dict:
    return {"n_mfcc": int(p.n_mfcc), "include_c0": bool(p.include_c0), "lifter": int(p.lifter)}


def mfcc_params_from_dict(d: dict) -> MfccParams:
    return MfccParams(int(d["n_mfcc"]), bool(d["include_c0"]), int(d["lifter"]))


def chroma_params_to_dict(p: ChromaParams) -> dict:
    return {"tuning": float(p.tuning), "n_octaves": int(p.n_octaves), "f_min": float(p.f_min), "f_max": float(p.f_max),
            "norm": _CHROMANORM[p.norm]}


def chroma_params_from_dict(d: dict) -> ChromaParams:
    p = ChromaParams(float(d["tuning"]), float(d["f_min"]), float(d["f_max"]), _inv(_CHROMANORM, d["norm"], "ChromaNorm"))
    p._n_octaves = int(d["n_octaves"])
    return p


# ------------------------------------------------------------------------------------------------ arrays and results
def _host(a) -> np.ndarray:
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def array2_to_obj(a) -> dict:
    """ndarray's ``Array2<T>`` serde form: version tag, dimensions, row-major elements."""
    a = _host(a)
    if a.ndim != 2:
        raise InvalidInputError("Array2 expects a 2-D array (serialise batches clip by clip)")
    if np.iscomplexobj(a):                                   # Complex<T> serialises as the pair [re, im]
        flat = [[float(z.real), float(z.imag)] for z in a.reshape(-1)]
    else:
        flat = [float(v) for v in a.reshape(-1)]
    return {"v": 1, "dim": [int(a.shape[0]), int(a.shape[1])], "data": flat}


def array2_from_obj(o: dict, dtype=np.float64) -> np.ndarray:
    if int(o.get("v", 0)) != 1:
        raise InvalidInputError("unsupported ndarray serde version")
    rows, cols = (int(v) for v in o["dim"])
    data = o["data"]
    if len(data) != rows * cols:
        raise InvalidInputError(f"Array2: expected {rows * cols} elements, got {len(data)}")
    if data and isinstance(data[0], (list, tuple)):
        arr = np.array([complex(re, im) for re, im in data], dtype=np.complex64 if np.dtype(dtype) == np.float32 else np.complex128)
    else:
        arr = np.array(data, dtype=dtype)
    return arr.reshape(rows, cols)


def spectrogram_to_dict(spec) -> dict:
    """``Spectrogram`` -> the object ``serde_json::to_value(&spec)`` produces."""
    return {"data": array2_to_obj(spec.data),
            "axes": {"freq": {"frequencies": [float(f) for f in spec.frequencies]}, "times": [float(t) for t in spec.times]},
            "params": spectrogram_params_to_dict(spec.params)}


def spectrogram_from_dict(d: dict, freq_scale: str = "linear", amp_scale: str = "power", dtype=np.float64):
    from .plan import Spectrogram
    data = array2_from_obj(d["data"], dtype)
    freqs = np.array(d["axes"]["freq"]["frequencies"], dtype=np.float64)
    times = np.array(d["axes"]["times"], dtype=np.float64)
    if freqs.size != data.shape[0] or times.size != data.shape[1]:
        raise InvalidInputError("axes do not match the data dimensions")
    return Spectrogram(data, freqs, times, spectrogram_params_from_dict(d["params"]), freq_scale, amp_scale)


def mfcc_to_dict(m) -> dict:
    return {"data": array2_to_obj(m.data), "params": mfcc_params_to_dict(m.params)}


def mfcc_from_dict(d: dict, dtype=np.float64):
    from .plan import Mfcc
    return Mfcc(array2_from_obj(d["data"], dtype), mfcc_params_from_dict(d["params"]))


def chromagram_to_dict(c) -> dict:
    return {"data": array2_to_obj(c.data), "params": chroma_params_to_dict(c.params)}


def chromagram_from_dict(d: dict, dtype=np.float64):
    from .plan import Chromagram
    return Chromagram(array2_from_obj(d["data"], dtype), chroma_params_from_dict(d["params"]))


def to_json(obj) -> str:
    """serde_json::to_string for the result / parameter types above (compact separators, like serde_json)."""
    from .plan import Chromagram, Mfcc, Spectrogram
    table = [(Spectrogram, spectrogram_to_dict), (Mfcc, mfcc_to_dict), (Chromagram, chromagram_to_dict),
             (SpectrogramParams, spectrogram_params_to_dict), (StftParams, stft_params_to_dict), (WindowType, window_to_obj),
             (MelParams, mel_params_to_dict), (LogHzParams, loghz_params_to_dict), (ErbParams, erb_params_to_dict),
             (LogParams, log_params_to_dict), (MfccParams, mfcc_params_to_dict), (ChromaParams, chroma_params_to_dict)]
    for cls, fn in table:
        if isinstance(obj, cls):
            return json.dumps(fn(obj), separators=(",", ":"))
    raise InvalidInputError(f"no serde form for {type(obj).__name__}")


def spectrogram_from_json(s: str, freq_scale: str = "linear", amp_scale: str = "power", dtype=np.float64):
    return spectrogram_from_dict(json.loads(s), freq_scale, amp_scale, dtype)


def mfcc_from_json(s: str, dtype=np.float64):
    return mfcc_from_dict(json.loads(s), dtype)


def chromagram_from_json(s: str, dtype=np.float64):
    return chromagram_from_dict(json.loads(s), dtype)
