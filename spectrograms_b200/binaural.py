"""Interaural cue spectrograms (reference ``src/binaural.rs``): ITD, IPD, ILD and ILR from a stereo pair, computed by
two STFT launches and one element-wise kernel without the complex spectra leaving device memory
(``sgx_plan_compute_binaural``). Names, argument meaning and error messages follow the reference:
``compute_itd_spectrogram(audio=[left, right], params, plan)`` with a reusable ``StftPlan``. New: ``left`` / ``right``
may be (n_pairs, n_samples) batches, and CUDA ``torch.Tensor`` inputs give CUDA outputs on the current stream.
The histogram / median "diff" helpers of the reference are host-side analysis utilities and are not mirrored.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native
from .errors import InvalidInputError
from .params import SpectrogramParams, normalise_dtype
from .plan import StftPlan, _is_torch, _torch

_CUE = {"itd": 0, "ipd": 1, "ild": 2, "ilr": 3}


def _check_band(spec_params: SpectrogramParams, start_freq: float, stop_freq: float, check_rate: bool) -> Tuple[float, float]:
    # ITDSpectrogramParams::new (src/binaural.rs:410-444); the IPD / ILD / ILR constructors repeat it (:775-806, :1133-1163, :1475-1505)
    start_freq, stop_freq = float(start_freq), float(stop_freq)
    if start_freq <= 0.0 or stop_freq <= 0.0:
        raise InvalidInputError("Start and end frequencies must be positive.")
    if start_freq >= stop_freq:
        raise InvalidInputError("Start frequency must be less than end frequency.")
    if check_rate and spec_params.sample_rate <= 0.0:
        raise InvalidInputError("Sample rate must be positive.")
    if stop_freq > spec_params.sample_rate / 2.0:
        raise InvalidInputError("End frequency must be less than Nyquist frequency.")
    return start_freq, stop_freq


class _BandParams:
    cue = ""

    def __init__(self, spec_params: SpectrogramParams, start_freq: float, stop_freq: float):
        self.spectrogram_params = spec_params
        self.start_freq, self.end_freq = _check_band(spec_params, start_freq, stop_freq, self.cue == "itd")
        self.magphase_power = 1
        self.wrapped = True

    def band(self) -> Tuple[int, int, float]:
        """(start_bin, stop_bin, bin_width): ``(f / bin_width).round() as usize`` (src/binaural.rs:476-481)."""
        bw = self.spectrogram_params.sample_rate / self.spectrogram_params.stft.n_fft
        rnd = lambda v: int(math.floor(v + 0.5))          # f64::round on non-negative values (half away from zero)
        return rnd(self.start_freq / bw), rnd(self.end_freq / bw), bw


class ITDSpectrogramParams(_BandParams):
    """``ITDSpectrogramParams::new(spec_params, start_freq, stop_freq, magphase_power)`` (src/binaural.rs:386-444)."""
    cue = "itd"

    def __init__(self, spec_params, start_freq, stop_freq, magphase_power: Optional[int] = None):
        super().__init__(spec_params, start_freq, stop_freq)
        if magphase_power is not None and int(magphase_power) <= 0:
            raise InvalidInputError("magphase_power must be non-zero")
        self.magphase_power = 1 if magphase_power is None else int(magphase_power)


class IPDSpectrogramParams(_BandParams):
    """``IPDSpectrogramParams::new(spec_params, start_freq, stop_freq, wrapped)`` (src/binaural.rs:755-806)."""
    cue = "ipd"

    def __init__(self, spec_params, start_freq, stop_freq, wrapped: bool = True):
        super().__init__(spec_params, start_freq, stop_freq)
        self.wrapped = bool(wrapped)


class ILDSpectrogramParams(_BandParams):
    """``ILDSpectrogramParams::new`` (src/binaural.rs:1115-1163)."""
    cue = "ild"


class ILRSpectrogramParams(_BandParams):
    """``ILRSpectrogramParams::new`` (src/binaural.rs:1457-1505)."""
    cue = "ilr"


class BinauralSpectrogram:
    """``ItdSpectrogram`` / ``IpdSpectrogram`` / ``IldSpectrogram`` / ``IlrSpectrogram`` (src/binaural.rs:185-320, ...):
    ``data`` (n_bins, n_frames) [or (n_pairs, n_bins, n_frames)], ``frequencies`` = bin * bin_width over the band,
    ``times`` = frame * hop / sample_rate. Units: seconds / radians / dB / ratio."""

    UNITS = {"itd": "seconds", "ipd": "radians", "ild": "dB", "ilr": "ratio"}

    def __init__(self, cue: str, data, frequencies: np.ndarray, times: np.ndarray, params: _BandParams):
        self.cue, self.data, self.frequencies, self.times, self.params = cue, data, frequencies, times, params

    unit = property(lambda s: s.UNITS[s.cue])
    n_bins = property(lambda s: int(s.data.shape[-2]))
    n_frames = property(lambda s: int(s.data.shape[-1]))
    shape = property(lambda s: tuple(s.data.shape))

    def frequency_range(self) -> Tuple[float, float]:
        return float(self.frequencies[0]), float(self.frequencies[-1])

    def duration(self) -> float:
        return float(self.times[-1])

    def __array__(self, dtype=None):
        a = self.data.detach().cpu().numpy() if _is_torch(self.data) else self.data
        return a.astype(dtype) if dtype is not None else a


def _compute(cue: str, audio: Sequence, params: _BandParams, plan: StftPlan) -> BinauralSpectrogram:
    if not isinstance(plan, StftPlan):
        raise InvalidInputError("plan must be a StftPlan (the reference takes `plan: &mut StftPlan<T>`)")
    if len(audio) != 2:
        raise InvalidInputError("audio must be [left, right]")
    left, right = audio
    b0, b1, bw = params.band()
    if b1 <= b0:
        raise InvalidInputError("Frequency range should have at least one bin")
    n = plan._n
    L = _native.lib()
    squeeze = False
    if _is_torch(left) != _is_torch(right):
        raise InvalidInputError("left and right must be the same kind of array")
    if _is_torch(left):
        torch = _torch()
        n._check_torch(left)
        n._check_torch(right)
        if left.dim() == 1:
            left, right, squeeze = left.unsqueeze(0), right.unsqueeze(0), True
        if left.dim() != 2 or left.shape != right.shape or left.numel() == 0:
            raise InvalidInputError("left and right must be equally shaped, non-empty (n_samples,) or (n_pairs, n_samples)")
        left, right = left.contiguous(), right.contiguous()
        npairs, ns = left.shape
        _, nf = n.output_shape(ns)
        out = torch.empty((npairs, b1 - b0, nf), dtype=left.dtype, device=left.device)
        with torch.cuda.device(left.device):
            _native.check(L.sgx_plan_compute_binaural(n._h, _CUE[cue], left.data_ptr(), right.data_ptr(), npairs, ns, ns,
                                                      params.start_freq, params.end_freq, params.magphase_power, int(params.wrapped),
                                                      out.data_ptr(), b1 - b0, nf, torch.cuda.current_stream(left.device).cuda_stream))
    else:
        left = np.ascontiguousarray(left, dtype=n.np_dtype)
        right = np.ascontiguousarray(right, dtype=n.np_dtype)
        if left.ndim == 1:
            left, right, squeeze = left[None], right[None] if right.ndim == 1 else right, True
        if left.ndim != 2 or left.shape != right.shape or left.size == 0:
            raise InvalidInputError("left and right must be equally shaped, non-empty (n_samples,) or (n_pairs, n_samples)")
        npairs, ns = left.shape
        _, nf = n.output_shape(ns)
        out = np.empty((npairs, b1 - b0, nf), dtype=n.np_dtype)
        _native.check(L.sgx_plan_compute_binaural(n._h, _CUE[cue], left.ctypes.data, right.ctypes.data, npairs, ns, ns,
                                                  params.start_freq, params.end_freq, params.magphase_power, int(params.wrapped),
                                                  out.ctypes.data, b1 - b0, nf, None))
    sp = params.spectrogram_params
    freqs = np.arange(b0, b1, dtype=np.float64) * bw                                   # :548-550
    times = np.arange(nf, dtype=np.float64) * float(sp.stft.hop_size) / sp.sample_rate   # :553-557
    return BinauralSpectrogram(cue, out[0] if squeeze else out, freqs, times, params)


def compute_itd_spectrogram(audio, params: ITDSpectrogramParams, plan: StftPlan) -> BinauralSpectrogram:
    """``compute_itd_spectrogram`` (src/binaural.rs:472-580): phase difference wrapped to [-pi, pi) over 2 pi f, in seconds."""
    return _compute("itd", audio, params, plan)


def compute_ipd_spectrogram(audio, params: IPDSpectrogramParams, plan: StftPlan) -> BinauralSpectrogram:
    """``compute_ipd_spectrogram`` (src/binaural.rs:830-917), radians."""
    return _compute("ipd", audio, params, plan)


def compute_ild_spectrogram(audio, params: ILDSpectrogramParams, plan: StftPlan) -> BinauralSpectrogram:
    """``compute_ild_spectrogram`` (src/binaural.rs:1187-1262): -20 log10(|R| / |L|), NaN where a channel is silent."""
    return _compute("ild", audio, params, plan)


def compute_ilr_spectrogram(audio, params: ILRSpectrogramParams, plan: StftPlan) -> BinauralSpectrogram:
    """``compute_ilr_spectrogram`` (src/binaural.rs:1530-1620): 1 - r for r = |R|/|L| < 1, else -(1 - 1/r)."""
    return _compute("ilr", audio, params, plan)


def binaural_from_stft(cue: str, left_stft, right_stft, start_bin: int, stop_bin: int, bin_width: float,
                       magphase_power: int = 1, wrapped: bool = True):
    """The element-wise half on complex STFTs the caller already holds ((n_bins, n_frames) or (n_pairs, n_bins, n_frames));
    ``sgx_binaural_from_stft``."""
    L = _native.lib()
    squeeze = False
    if _is_torch(left_stft):
        torch = _torch()
        l, r = left_stft, right_stft
        if not (l.is_cuda and r.is_cuda) or not l.is_complex():
            raise InvalidInputError("torch inputs must be complex CUDA tensors")
        if l.dim() == 2:
            l, r, squeeze = l.unsqueeze(0), r.unsqueeze(0), True
        l, r = l.contiguous(), r.contiguous()
        npairs, nb, nf = l.shape
        rdt = torch.float32 if l.dtype == torch.complex64 else torch.float64
        out = torch.empty((npairs, stop_bin - start_bin, nf), dtype=rdt, device=l.device)
        with torch.cuda.device(l.device):
            _native.check(L.sgx_binaural_from_stft(0 if rdt == torch.float32 else 1, _CUE[cue], l.data_ptr(), r.data_ptr(), npairs, nb, nf,
                                                   start_bin, stop_bin, float(bin_width), magphase_power, int(wrapped), out.data_ptr(),
                                                   l.device.index, torch.cuda.current_stream(l.device).cuda_stream))
            torch.cuda.current_stream(l.device).synchronize()
        return out[0] if squeeze else out
    l, r = np.ascontiguousarray(left_stft), np.ascontiguousarray(right_stft)
    if l.dtype not in (np.complex64, np.complex128) or l.dtype != r.dtype or l.shape != r.shape:
        raise InvalidInputError("left and right must be equally shaped complex64 / complex128 arrays")
    if l.ndim == 2:
        l, r, squeeze = l[None], r[None], True
    npairs, nb, nf = l.shape
    rdt = np.float32 if l.dtype == np.complex64 else np.float64
    out = np.empty((npairs, stop_bin - start_bin, nf), dtype=rdt)
    _native.check(L.sgx_binaural_from_stft(0 if rdt == np.float32 else 1, _CUE[cue], l.ctypes.data, r.ctypes.data, npairs, nb, nf,
                                           start_bin, stop_bin, float(bin_width), magphase_power, int(wrapped), out.ctypes.data, -1, None))
    return out[0] if squeeze else out
