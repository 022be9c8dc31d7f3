"""Host-side mirror of the reference's plan API on top of the C ABI (include/sgx_b200.h).

``SpectrogramPlanner`` / ``SpectrogramPlan`` / ``StftPlan`` / ``stft`` / ``mfcc`` / ``mfcc_from_log_mel`` keep the
reference's names, argument meaning and error behaviour (bare ``:N`` = ``src/spectrogram.rs:N`` of the reference).
What is new is the batched entry point ``compute_batch`` (the reference batches with a user loop,
src/lib.rs:228-235) and device-resident inputs: a CUDA ``torch.Tensor`` in gives a CUDA ``torch.Tensor`` out on the
current torch stream, with no host round trip. NumPy in gives NumPy out (staged through the library's pinned-chunk
pipeline). Every number is produced by the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native
from .errors import DimensionMismatchError, InvalidInputError
from .params import (_as_window, ChromaParams, ErbParams, LogHzParams, LogParams, MelParams, MfccParams, SpectrogramParams,
                     StftParams, WindowType, normalise_dtype)

_WIN = {"rectangular": 0, "hanning": 1, "hamming": 2, "blackman": 3, "kaiser": 4, "gaussian": 5, "custom": 6}
_MAP = {"linear": 0, "mel": 1, "erb": 2, "loghz": 3, "chroma": 4}
_CHROMA_NORM = {"none": 0, "l1": 1, "l2": 2, "max": 3}
_AMP = {"power": 0, "magnitude": 1, "db": 2, "decibels": 2}
_NORM = {"none": 0, "slaney": 1, "l1": 2, "l2": 3}
_OUT_SPEC, _OUT_STFT, _OUT_MFCC = 0, 1, 2


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _torch():
    import torch
    return torch


# ------------------------------------------------------------------------------------------------ results
class Spectrogram:
    """``Spectrogram<FreqScale, AmpScale, T>`` (:2547-2557): data (n_bins, n_frames) + axes + params."""

    def __init__(self, data, frequencies: np.ndarray, times: np.ndarray, params: SpectrogramParams, freq_scale: str,
                 amp_scale: str):
        self.data = data
        self.frequencies = frequencies
        self.times = times
        self.params = params
        self.freq_scale = freq_scale
        self.amp_scale = amp_scale

    n_bins = property(lambda s: int(s.data.shape[-2]))
    n_frames = property(lambda s: int(s.data.shape[-1]))
    shape = property(lambda s: tuple(s.data.shape))
    dtype = property(lambda s: str(s.data.dtype).replace("torch.", ""))

    def frequency_range(self) -> Tuple[float, float]:
        return float(self.frequencies[0]), float(self.frequencies[-1])

    def duration(self) -> float:
        return float(self.times[-1])

    def db_range(self) -> Optional[Tuple[float, float]]:
        if self.amp_scale != "db":
            return None
        d = self.data
        return float(d.min()), float(d.max())

    def __array__(self, dtype=None):
        d = self.data.cpu().numpy() if _is_torch(self.data) else self.data
        return d.astype(dtype) if dtype is not None else d

    def __len__(self):
        return self.n_bins

    # DLPack (reference: src/python/dlpack.rs, python/spectrograms/torch.py). Device-resident results export as
    # kDLCUDA without a host round trip -- the reference can only offer CPU tensors.
    def to_torch(self):
        return self.data if _is_torch(self.data) else _torch().from_numpy(self.data)

    def to_json(self) -> str:
        """serde wire format of the crate (``serde_json::to_string(&spec)``, src/spectrogram.rs:2546-2557) -- see serde.py."""
        from . import serde
        return serde.to_json(self)

    @staticmethod
    def from_json(s: str, freq_scale: str = "linear", amp_scale: str = "power", dtype=np.float64) -> "Spectrogram":
        from . import serde
        return serde.spectrogram_from_json(s, freq_scale, amp_scale, dtype)

    def __dlpack__(self, stream=None):
        t = self.to_torch()
        return t.__dlpack__(stream=stream) if stream is not None else t.__dlpack__()

    def __dlpack_device__(self):
        return self.to_torch().__dlpack_device__()


class StftResult:
    """``StftResult<T>`` (:534-630)."""

    def __init__(self, data, frequencies: np.ndarray, sample_rate: float, params: StftParams):
        self.data = data
        self.frequencies = frequencies
        self.sample_rate = sample_rate
        self.params = params

    n_bins = property(lambda s: int(s.data.shape[-2]))
    n_frames = property(lambda s: int(s.data.shape[-1]))
    shape = property(lambda s: tuple(s.data.shape))

    def frequency_resolution(self) -> float:
        return self.sample_rate / float(self.params.n_fft)

    def time_resolution(self) -> float:
        return float(self.params.hop_size) / self.sample_rate

    def norm(self):
        """``StftResult::norm`` (:597-599): element-wise |X|."""
        return self.data.abs() if _is_torch(self.data) else np.abs(self.data)

    def __array__(self, dtype=None):
        d = self.data.cpu().numpy() if _is_torch(self.data) else self.data
        return d.astype(dtype) if dtype is not None else d


class Mfcc:
    """``Mfcc<T>`` (src/mfcc.rs:146-204)."""

    def __init__(self, data, params: MfccParams):
        self.data = data
        self.params = params

    n_coefficients = property(lambda s: int(s.data.shape[-2]))
    n_bins = n_coefficients
    n_frames = property(lambda s: int(s.data.shape[-1]))
    shape = property(lambda s: tuple(s.data.shape))

    def __array__(self, dtype=None):
        d = self.data.cpu().numpy() if _is_torch(self.data) else self.data
        return d.astype(dtype) if dtype is not None else d


# ------------------------------------------------------------------------------------------------ native plan
class _NativePlan:
    """Owns one ``sgx_plan*``."""

    def __init__(self, params: SpectrogramParams, dtype: str, mapping: str = "linear", scale=None, amp: str = "power",
                 db: Optional[LogParams] = None, output: int = _OUT_SPEC, mfcc: Optional[MfccParams] = None,
                 device: Optional[int] = None):
        self.params = params
        self.dtype = normalise_dtype(dtype)
        self.np_dtype = np.float32 if self.dtype == "f32" else np.float64
        self.output = output
        self.amp = "db" if amp == "decibels" else amp
        self.mapping = mapping
        st = params.stft
        d = _native.PlanDesc()
        d.dtype = 0 if self.dtype == "f32" else 1
        d.n_fft, d.hop_size, d.centre = st.n_fft, st.hop_size, int(st.centre)
        d.window = _WIN[st.window.kind]
        d.window_param = float(st.window.param)
        self._custom = None
        if st.window.kind == "custom":
            self._custom = np.ascontiguousarray(st.window.coefficients, dtype=np.float64)
            d.custom_window = self._custom.ctypes.data_as(C.POINTER(C.c_double))
            d.custom_window_len = self._custom.size
        d.sample_rate_hz = params.sample_rate
        d.mapping = _MAP[mapping]
        if mapping == "mel":
            d.n_bands, d.f_min, d.f_max, d.mel_norm = scale.n_mels, scale.f_min, scale.f_max, _NORM[scale.norm]
        elif mapping == "erb":
            d.n_bands, d.f_min, d.f_max = scale.n_filters, scale.f_min, scale.f_max
            d.erb_spacing = 0 if scale.spacing == "linear" else 1
        elif mapping == "loghz":
            d.n_bands, d.f_min, d.f_max = scale.n_bins, scale.f_min, scale.f_max
        elif mapping == "chroma":
            d.n_bands, d.f_min, d.f_max = 12, scale.f_min, scale.f_max
            d.chroma_tuning, d.chroma_norm = scale.tuning, _CHROMA_NORM[scale.norm]
        d.amp = _AMP[self.amp]
        d.has_floor_db = int(db is not None)
        d.floor_db = db.floor_db if db is not None else 0.0
        d.output = output
        if mfcc is not None:
            d.n_mfcc, d.include_c0, d.lifter = mfcc.n_mfcc, int(mfcc.include_c0), mfcc.lifter
        d.device = -1 if device is None else int(device)
        self._desc = d
        self._h = C.c_void_p()
        L = _native.lib()
        _native.check(L.sgx_plan_create(C.byref(d), C.byref(self._h)))
        self.n_fft = st.n_fft

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _native.lib().sgx_plan_destroy(h)
            except Exception:          # interpreter shutdown: module globals may already be gone
                pass
            self._h = C.c_void_p()

    # -- queries
    def output_shape(self, n_samples: int) -> Tuple[int, int]:
        if int(n_samples) <= 0:
            raise InvalidInputError("signal length must be non-zero")
        r, f = C.c_size_t(), C.c_size_t()
        _native.check(_native.lib().sgx_plan_output_shape(self._h, int(n_samples), C.byref(r), C.byref(f)))
        return r.value, f.value

    def axes(self, n_frames: int) -> Tuple[np.ndarray, np.ndarray]:
        nb = self.n_axis_bins()
        fr = np.empty(nb, dtype=np.float64)
        tm = np.empty(n_frames, dtype=np.float64)
        _native.check(_native.lib().sgx_plan_axes(self._h, n_frames, fr.ctypes.data, tm.ctypes.data))
        return fr, tm

    def n_axis_bins(self) -> int:
        d = self._desc
        return int(d.n_bands) if d.mapping != 0 else self.n_fft // 2 + 1

    def window(self) -> np.ndarray:
        w = np.empty(self.n_fft, dtype=self.np_dtype)
        _native.check(_native.lib().sgx_plan_window(self._h, w.ctypes.data))
        return w

    def filterbank(self) -> Tuple[np.ndarray, int]:
        m = np.empty((self.n_axis_bins(), self.n_fft // 2 + 1), dtype=np.float64)
        nnz = C.c_size_t()
        _native.check(_native.lib().sgx_plan_filterbank(self._h, m.ctypes.data, C.byref(nnz)))
        return m, nnz.value

    def kernel_name(self) -> str:
        return _native.lib().sgx_plan_kernel_name(self._h).decode()

    def last_launch_count(self) -> int:
        return int(_native.lib().sgx_plan_last_launch_count(self._h))

    def force_generic(self, force: bool = True) -> None:
        _native.check(_native.lib().sgx_plan_force_generic(self._h, int(force)))

    def set_tensor_cores(self, enable=True) -> None:
        """TMEM + tcgen05 kernel variant of the plan's family: True = whenever supported, False = never, None = automatic
        (the default: where it is measured faster). Test and measurement hook."""
        _native.check(_native.lib().sgx_plan_set_tensor_cores(self._h, -1 if enable is None else int(bool(enable))))

    def set_tmem_exchange(self, enable=True) -> None:
        """Tensor memory as the exchange medium of the n_fft = 400 family's two FFT passes (r2c_fused_n400_tm): True = whenever
        supported, False = never (shared-memory exchange), None = automatic (the default: wherever supported). Bit-identical
        results either way. Test and measurement hook."""
        _native.check(_native.lib().sgx_plan_set_tmem_exchange(self._h, -1 if enable is None else int(bool(enable))))

    # -- compute
    def _out_dtype(self, torch_mod=None):
        cplx = self.output == _OUT_STFT
        if torch_mod is None:
            if cplx:
                return np.complex64 if self.dtype == "f32" else np.complex128
            return self.np_dtype
        if cplx:
            return torch_mod.complex64 if self.dtype == "f32" else torch_mod.complex128
        return torch_mod.float32 if self.dtype == "f32" else torch_mod.float64

    def _check_torch(self, x):
        torch = _torch()
        want = torch.float32 if self.dtype == "f32" else torch.float64
        if x.dtype != want:
            raise InvalidInputError(f"samples dtype {x.dtype} does not match the plan dtype {want}")
        if not x.is_cuda:
            raise InvalidInputError("torch inputs must be CUDA tensors (use NumPy arrays for host data)")
        return torch

    @staticmethod
    def _check_out_dims(shape, n_clips: int) -> None:
        """The C ABI sees only the trailing (rows, frames) of ``out``; the clip dimension is checked here so that a
        short or 2-D buffer can never be overrun: (n_clips, rows, frames), or (rows, frames) for a single clip."""
        if len(shape) == 3:
            if shape[0] != n_clips:
                raise DimensionMismatchError(f"Dimension mismatch: expected {n_clips}, got {shape[0]}", n_clips, int(shape[0]))
        elif not (len(shape) == 2 and n_clips == 1):
            raise InvalidInputError("out must be (n_clips, rows, n_frames), or (rows, n_frames) for a single clip")

    def compute_batch(self, clips, out=None):
        """clips: (n_clips, n_samples). Returns / fills (n_clips, rows, n_frames)."""
        L = _native.lib()
        if _is_torch(clips):
            torch = self._check_torch(clips)
            if clips.dim() != 2 or clips.shape[1] == 0 or clips.shape[0] == 0:
                raise InvalidInputError("clips must be a non-empty 2-D tensor (n_clips, n_samples)")
            if clips.stride(1) != 1:
                clips = clips.contiguous()
            n_clips, n_samples = clips.shape
            rows, nf = self.output_shape(n_samples)
            if out is None:
                out = torch.empty((n_clips, rows, nf), dtype=self._out_dtype(torch), device=clips.device)
            else:
                if not _is_torch(out) or not out.is_cuda or out.device != clips.device:
                    raise InvalidInputError("out must be a CUDA tensor on the same device as the input")
                if out.dtype != self._out_dtype(torch):
                    raise InvalidInputError(f"out dtype {out.dtype} does not match the plan's output dtype {self._out_dtype(torch)}")
                if not out.is_contiguous():
                    raise InvalidInputError("out must be contiguous")
                self._check_out_dims(tuple(out.shape), n_clips)
            orow, ocol = (out.shape[-2], out.shape[-1]) if out.dim() >= 2 else (0, 0)
            stream = torch.cuda.current_stream(clips.device).cuda_stream
            with torch.cuda.device(clips.device):
                _native.check(L.sgx_plan_compute_batch(self._h, clips.data_ptr(), n_clips, n_samples, clips.stride(0),
                                                       out.data_ptr(), orow, ocol, 0, stream))
            return out
        clips = np.ascontiguousarray(clips, dtype=self.np_dtype)
        if clips.ndim != 2 or clips.shape[0] == 0 or clips.shape[1] == 0:
            raise InvalidInputError("clips must be a non-empty 2-D array (n_clips, n_samples)")
        n_clips, n_samples = clips.shape
        rows, nf = self.output_shape(n_samples)
        if out is None:
            out = np.empty((n_clips, rows, nf), dtype=self._out_dtype())
        else:
            if not isinstance(out, np.ndarray) or not out.flags.c_contiguous or out.dtype != self._out_dtype():
                raise InvalidInputError(f"out must be a C-contiguous NumPy array of dtype {np.dtype(self._out_dtype())}")
            if not out.flags.writeable:
                raise InvalidInputError("out must be writeable")
            self._check_out_dims(out.shape, n_clips)
        orow, ocol = (out.shape[-2], out.shape[-1]) if out.ndim >= 2 else (0, 0)
        _native.check(L.sgx_plan_compute_batch(self._h, clips.ctypes.data, n_clips, n_samples, n_samples,
                                               out.ctypes.data, orow, ocol, 0, None))
        return out

    def istft(self, stft_matrix):
        L = _native.lib()
        bins = self.n_fft // 2 + 1
        squeeze = False
        n_out = C.c_size_t(0)
        if _is_torch(stft_matrix):
            torch = _torch()
            x = stft_matrix
            want = torch.complex64 if self.dtype == "f32" else torch.complex128
            if not x.is_cuda or x.dtype != want:
                raise InvalidInputError(f"stft_matrix must be a CUDA tensor of dtype {want}")
            if x.dim() == 2:
                x, squeeze = x.unsqueeze(0), True
            if x.dim() != 3 or x.numel() == 0:
                raise InvalidInputError("stft_matrix must be (n_bins, n_frames) or (n_clips, n_bins, n_frames)")
            if x.shape[1] != bins:
                raise DimensionMismatchError(f"Dimension mismatch: expected {bins}, got {x.shape[1]}", bins, int(x.shape[1]))
            x = x.contiguous()
            nc, _, nf = x.shape
            _native.check(L.sgx_plan_istft(self._h, None, nc, nf, None, C.byref(n_out), None))
            out = torch.empty((nc, n_out.value), dtype=torch.float32 if self.dtype == "f32" else torch.float64, device=x.device)
            with torch.cuda.device(x.device):
                _native.check(L.sgx_plan_istft(self._h, x.data_ptr(), nc, nf, out.data_ptr(), C.byref(n_out),
                                               torch.cuda.current_stream(x.device).cuda_stream))
            return out[0] if squeeze else out
        x = np.asarray(stft_matrix)
        x = np.ascontiguousarray(x, dtype=np.complex64 if self.dtype == "f32" else np.complex128)
        if x.ndim == 2:
            x, squeeze = x[None], True
        if x.ndim != 3 or x.size == 0:
            raise InvalidInputError("stft_matrix must be (n_bins, n_frames) or (n_clips, n_bins, n_frames)")
        if x.shape[1] != bins:                                     # :4825-4828
            raise DimensionMismatchError(f"Dimension mismatch: expected {bins}, got {x.shape[1]}", bins, int(x.shape[1]))
        nc, _, nf = x.shape
        _native.check(L.sgx_plan_istft(self._h, None, nc, nf, None, C.byref(n_out), None))
        out = np.empty((nc, n_out.value), dtype=self.np_dtype)
        _native.check(L.sgx_plan_istft(self._h, x.ctypes.data, nc, nf, out.ctypes.data, C.byref(n_out), None))
        return out[0] if squeeze else out

    def compute_one(self, samples, out=None):
        if _is_torch(samples):
            if samples.dim() != 1 or samples.numel() == 0:
                raise InvalidInputError("samples must be a non-empty 1-D tensor")
            res = self.compute_batch(samples.unsqueeze(0), None if out is None else out)
            return res[0] if out is None else out
        samples = np.asarray(samples)
        if samples.ndim != 1 or samples.size == 0:
            raise InvalidInputError("samples must be a non-empty 1-D array")
        res = self.compute_batch(samples[None, :], out)
        return res[0] if out is None else out

    def compute_frame(self, samples, frame_idx: int):
        L = _native.lib()
        if int(frame_idx) < 0:
            raise InvalidInputError("frame index must be non-negative")
        rows = self.output_shape(max(1, len(samples)))[0]
        if _is_torch(samples):
            torch = self._check_torch(samples)
            if samples.dim() != 1 or samples.numel() == 0:
                raise InvalidInputError("samples must be a non-empty 1-D tensor")
            samples = samples.contiguous()
            out = torch.empty(rows, dtype=self._out_dtype(torch), device=samples.device)
            stream = torch.cuda.current_stream(samples.device).cuda_stream
            with torch.cuda.device(samples.device):
                _native.check(L.sgx_plan_compute_frame(self._h, samples.data_ptr(), samples.numel(), int(frame_idx),
                                                       out.data_ptr(), stream))
            return out
        samples = np.ascontiguousarray(samples, dtype=self.np_dtype)
        if samples.ndim != 1 or samples.size == 0:
            raise InvalidInputError("samples must be a non-empty 1-D array")
        out = np.empty(rows, dtype=self._out_dtype())
        _native.check(L.sgx_plan_compute_frame(self._h, samples.ctypes.data, samples.size, int(frame_idx),
                                               out.ctypes.data, None))
        return out


# ------------------------------------------------------------------------------------------------ public plans
class SpectrogramPlan:
    """``SpectrogramPlan<FreqScale, AmpScale, T>`` (:172-520)."""

    def __init__(self, native: _NativePlan, freq_scale: str, amp_scale: str):
        self._n = native
        self.freq_scale = freq_scale
        self.amp_scale = amp_scale

    dtype = property(lambda s: "float32" if s._n.dtype == "f32" else "float64")
    params = property(lambda s: s._n.params)

    def freq_axis(self) -> np.ndarray:
        return self._n.axes(0)[0]

    def output_shape(self, signal_length: int) -> Tuple[int, int]:
        """(:512-519)"""
        return self._n.output_shape(signal_length)

    def compute(self, samples) -> Spectrogram:
        """(:240-294)"""
        data = self._n.compute_one(samples)
        fr, tm = self._n.axes(int(data.shape[-1]))
        return Spectrogram(data, fr, tm, self._n.params, self.freq_scale, self.amp_scale)

    def compute_into(self, samples, output) -> None:
        """(:414-477): shape-checked, rows then columns -> DimensionMismatchError."""
        self._n.compute_one(samples, output)

    def compute_frame(self, samples, frame_idx: int):
        """(:335-372)"""
        return self._n.compute_frame(samples, frame_idx)

    def compute_batch(self, clips, out=None):
        """New: ``for s in clips { plan.compute_into(s, out[i]) }`` as one call -> (n_clips, n_bins, n_frames)."""
        return self._n.compute_batch(clips, out)

    # SpectrogramSource (src/source.rs:39-59)
    def compute_matrix(self, samples):
        return self._n.compute_one(samples)

    def n_bands(self) -> int:
        return self._n.n_axis_bins()

    def center_frequencies(self) -> np.ndarray:
        return self.freq_axis()

    def sample_rate(self) -> float:
        return self._n.params.sample_rate

    def hop_seconds(self) -> float:
        return self._n.params.frame_period_seconds()

    # introspection used by tests / bench
    def window(self) -> np.ndarray: return self._n.window()
    def filterbank(self): return self._n.filterbank()
    def kernel_name(self) -> str: return self._n.kernel_name()
    def last_launch_count(self) -> int: return self._n.last_launch_count()
    def force_generic(self, force: bool = True) -> None: self._n.force_generic(force)
    def set_tensor_cores(self, enable=True) -> None: self._n.set_tensor_cores(enable)
    def set_tmem_exchange(self, enable=True) -> None: self._n.set_tmem_exchange(enable)


class StftPlan:
    """``StftPlan<T>`` (:1173-1636)."""

    def __init__(self, params: SpectrogramParams, dtype="float64", device: Optional[int] = None):
        self._n = _NativePlan(params, dtype, output=_OUT_STFT, device=device)

    dtype = property(lambda s: "float32" if s._n.dtype == "f32" else "float64")
    n_fft = property(lambda s: s._n.params.stft.n_fft)
    hop_size = property(lambda s: s._n.params.stft.hop_size)
    n_bins = property(lambda s: s._n.params.stft.n_fft // 2 + 1)

    def output_shape(self, signal_length: int) -> Tuple[int, int]:
        return self._n.output_shape(signal_length)

    def compute(self, samples, params: Optional[SpectrogramParams] = None) -> StftResult:
        """(:1424-1458)"""
        params = params or self._n.params
        data = self._n.compute_one(samples)
        nb = data.shape[-2]
        freqs = np.arange(nb, dtype=np.float64) * params.sample_rate / float(params.stft.n_fft)
        return StftResult(data, freqs, params.sample_rate, params.stft)

    def compute_into(self, samples, output) -> None:
        """(:1548-1580)"""
        self._n.compute_one(samples, output)

    def compute_frame_simple(self, samples, frame_idx: int):
        """(:1500-1507)"""
        return self._n.compute_frame(samples, frame_idx)

    def compute_batch(self, clips, out=None):
        return self._n.compute_batch(clips, out)

    def istft(self, stft_matrix):
        """``istft`` (:4813-4911) with this plan's n_fft, hop, window and centre: (n_bins, n_frames) or
        (n_clips, n_bins, n_frames) complex -> (out_len,) or (n_clips, out_len) samples. NumPy in -> NumPy out, CUDA
        ``torch.Tensor`` in -> CUDA tensor out on the current stream."""
        return self._n.istft(stft_matrix)

    def window(self) -> np.ndarray: return self._n.window()
    def kernel_name(self) -> str: return self._n.kernel_name()
    def last_launch_count(self) -> int: return self._n.last_launch_count()
    def force_generic(self, force: bool = True) -> None: self._n.force_generic(force)
    def set_tensor_cores(self, enable=True) -> None: self._n.set_tensor_cores(enable)
    def set_tmem_exchange(self, enable=True) -> None: self._n.set_tmem_exchange(enable)


class MfccPlan:
    """Fused ``mfcc()`` (src/mfcc.rs:359-379): mel dB plan (f_min=0, f_max=sr/2, MelNorm::None, floor -80 dB) +
    ``mfcc_from_log_mel`` in one kernel; the log-mel spectrogram never reaches HBM."""

    def __init__(self, stft: StftParams, sample_rate: float, n_mels: int, mfcc_params: MfccParams, dtype="float64",
                 device: Optional[int] = None):
        params = SpectrogramParams(stft, sample_rate)
        mel = MelParams(n_mels, 0.0, sample_rate / 2.0)
        db = LogParams(-80.0)
        self.mfcc_params = mfcc_params
        self._n = _NativePlan(params, dtype, "mel", mel, "db", db, output=_OUT_MFCC, mfcc=mfcc_params, device=device)

    def output_shape(self, signal_length: int) -> Tuple[int, int]:
        return self._n.output_shape(signal_length)

    def compute(self, samples) -> Mfcc:
        return Mfcc(self._n.compute_one(samples), self.mfcc_params)

    def compute_batch(self, clips, out=None):
        return self._n.compute_batch(clips, out)

    def kernel_name(self) -> str: return self._n.kernel_name()
    def last_launch_count(self) -> int: return self._n.last_launch_count()
    def force_generic(self, force: bool = True) -> None: self._n.force_generic(force)
    def set_tensor_cores(self, enable=True) -> None: self._n.set_tensor_cores(enable)
    def set_tmem_exchange(self, enable=True) -> None: self._n.set_tmem_exchange(enable)


class Chromagram:
    """``Chromagram<T>`` (src/chroma.rs:186-260): (12, n_frames) pitch-class profile; ``data`` is a NumPy array or a
    CUDA ``torch.Tensor`` depending on what went in."""

    LABELS = ("C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B")      # :238-242

    def __init__(self, data, params: ChromaParams):
        self.data = data
        self.params = params

    n_bins = property(lambda s: int(s.data.shape[-2]))
    n_frames = property(lambda s: int(s.data.shape[-1]))
    shape = property(lambda s: tuple(s.data.shape))

    @staticmethod
    def labels() -> Tuple[str, ...]:
        return Chromagram.LABELS

    def __array__(self, dtype=None):
        a = self.data.detach().cpu().numpy() if _is_torch(self.data) else self.data
        return a.astype(dtype) if dtype is not None else a


class ChromaPlan:
    """Fused ``chromagram()`` (src/chroma.rs:487-503): STFT -> magnitude -> 12 x bins chroma filterbank -> per-frame
    normalisation in one kernel; the magnitude spectrogram never reaches HBM. Reusable and batched like every plan."""

    def __init__(self, stft: StftParams, sample_rate: float, chroma_params: ChromaParams, dtype="float64",
                 device: Optional[int] = None):
        self.chroma_params = chroma_params
        self._n = _NativePlan(SpectrogramParams(stft, sample_rate), dtype, "chroma", chroma_params, "magnitude", None,
                              device=device)

    def output_shape(self, signal_length: int) -> Tuple[int, int]:
        return self._n.output_shape(signal_length)

    def compute(self, samples) -> Chromagram:
        return Chromagram(self._n.compute_one(samples), self.chroma_params)

    def compute_batch(self, clips, out=None):
        return self._n.compute_batch(clips, out)

    def filterbank(self): return self._n.filterbank()
    def kernel_name(self) -> str: return self._n.kernel_name()
    def last_launch_count(self) -> int: return self._n.last_launch_count()
    def force_generic(self, force: bool = True) -> None: self._n.force_generic(force)
    def set_tensor_cores(self, enable=True) -> None: self._n.set_tensor_cores(enable)
    def set_tmem_exchange(self, enable=True) -> None: self._n.set_tmem_exchange(enable)


class SpectrogramPlanner:
    """``SpectrogramPlanner`` (:640-1153): a stateless factory; all state lives in the returned plan."""

    def __init__(self, device: Optional[int] = None):
        self.device = device

    # --- the reference's generic methods: planner.mel_plan::<A, T>(&params, &mel, Option<&LogParams>)
    def linear_plan(self, params: SpectrogramParams, db: Optional[LogParams] = None, amp: str = "power",
                    dtype="float64") -> SpectrogramPlan:
        return SpectrogramPlan(_NativePlan(params, dtype, "linear", None, amp, db, device=self.device), "linear", _amp(amp))

    def mel_plan(self, params: SpectrogramParams, mel: MelParams, db: Optional[LogParams] = None, amp: str = "power",
                 dtype="float64") -> SpectrogramPlan:
        return SpectrogramPlan(_NativePlan(params, dtype, "mel", mel, amp, db, device=self.device), "mel", _amp(amp))

    def erb_plan(self, params: SpectrogramParams, erb: ErbParams, db: Optional[LogParams] = None, amp: str = "power",
                 dtype="float64") -> SpectrogramPlan:
        return SpectrogramPlan(_NativePlan(params, dtype, "erb", erb, amp, db, device=self.device), "erb", _amp(amp))

    def log_hz_plan(self, params: SpectrogramParams, loghz: LogHzParams, db: Optional[LogParams] = None,
                    amp: str = "power", dtype="float64") -> SpectrogramPlan:
        return SpectrogramPlan(_NativePlan(params, dtype, "loghz", loghz, amp, db, device=self.device), "loghz", _amp(amp))

    def stft_plan(self, params: SpectrogramParams, dtype="float64") -> StftPlan:
        return StftPlan(params, dtype, device=self.device)

    def compute_stft(self, samples, params: SpectrogramParams, dtype=None) -> StftResult:
        """(:722-729)"""
        return StftPlan(params, dtype or _infer_dtype(samples), device=self.device).compute(samples, params)

    # --- the reference's Python-binding names (python/spectrograms/__init__.pyi:869-1040)
    def linear_power_plan(self, params, dtype="float64"): return self.linear_plan(params, None, "power", dtype)
    def linear_magnitude_plan(self, params, dtype="float64"): return self.linear_plan(params, None, "magnitude", dtype)
    def linear_db_plan(self, params, db_params, dtype="float64"): return self.linear_plan(params, db_params, "db", dtype)
    def mel_power_plan(self, params, mel_params, dtype="float64"): return self.mel_plan(params, mel_params, None, "power", dtype)
    def mel_magnitude_plan(self, params, mel_params, dtype="float64"): return self.mel_plan(params, mel_params, None, "magnitude", dtype)
    def mel_db_plan(self, params, mel_params, db_params, dtype="float64"): return self.mel_plan(params, mel_params, db_params, "db", dtype)
    def erb_power_plan(self, params, erb_params, dtype="float64"): return self.erb_plan(params, erb_params, None, "power", dtype)
    def erb_magnitude_plan(self, params, erb_params, dtype="float64"): return self.erb_plan(params, erb_params, None, "magnitude", dtype)
    def erb_db_plan(self, params, erb_params, db_params, dtype="float64"): return self.erb_plan(params, erb_params, db_params, "db", dtype)
    def loghz_power_plan(self, params, loghz_params, dtype="float64"): return self.log_hz_plan(params, loghz_params, None, "power", dtype)
    def loghz_magnitude_plan(self, params, loghz_params, dtype="float64"): return self.log_hz_plan(params, loghz_params, None, "magnitude", dtype)
    def loghz_db_plan(self, params, loghz_params, db_params, dtype="float64"): return self.log_hz_plan(params, loghz_params, db_params, "db", dtype)


def _amp(a: str) -> str:
    return "db" if a == "decibels" else a


def _infer_dtype(x) -> str:
    s = str(x.dtype)
    return "float32" if s.endswith("float32") else "float64"


# ------------------------------------------------------------------------------------------------ free functions
def stft(samples, n_fft: int, hop_size: int, window="hanning", center: bool = True, dtype=None):
    """``stft<T>()`` (:4733-4747): complex (n_fft/2+1, n_frames) matrix; builds a fresh plan with a dummy rate."""
    params = SpectrogramParams(StftParams(n_fft, hop_size, window, center), 1.0)
    return StftPlan(params, dtype or _infer_dtype(samples)).compute(samples, params).data


def compute_stft(samples, params: SpectrogramParams, dtype=None) -> StftResult:
    return SpectrogramPlanner().compute_stft(samples, params, dtype)


def mfcc_from_log_mel(log_mel_spec, params: MfccParams) -> Mfcc:
    """``mfcc_from_log_mel`` (src/mfcc.rs:224-273). log_mel_spec: (n_mels, n_frames) or (n_clips, n_mels, n_frames)."""
    L = _native.lib()
    squeeze = False
    if _is_torch(log_mel_spec):
        torch = _torch()
        x = log_mel_spec
        if not x.is_cuda:
            raise InvalidInputError("torch inputs must be CUDA tensors (use NumPy arrays for host data)")
        if x.dim() == 2:
            x, squeeze = x.unsqueeze(0), True
        if x.dim() != 3 or x.numel() == 0:
            raise InvalidInputError("log_mel_spec must be (n_mels, n_frames) or (n_clips, n_mels, n_frames)")
        x = x.contiguous()
        dt = normalise_dtype(x.dtype)
        nc, nm, nf = x.shape
        if params.n_mfcc > nm:
            raise InvalidInputError("n_mfcc must be <= n_mels")
        rows = params.n_mfcc - (0 if (params.include_c0 or params.n_mfcc == 1) else 1)
        out = torch.empty((nc, rows, nf), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _native.check(L.sgx_mfcc_from_log_mel(0 if dt == "f32" else 1, x.data_ptr(), nc, nm, nf, params.n_mfcc,
                                                  int(params.include_c0), params.lifter, out.data_ptr(),
                                                  x.device.index, torch.cuda.current_stream(x.device).cuda_stream))
        return Mfcc(out[0] if squeeze else out, params)
    x = np.asarray(log_mel_spec)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    if x.ndim == 2:
        x, squeeze = x[None], True
    if x.ndim != 3 or x.size == 0:
        raise InvalidInputError("log_mel_spec must be (n_mels, n_frames) or (n_clips, n_mels, n_frames)")
    x = np.ascontiguousarray(x)
    nc, nm, nf = x.shape
    if params.n_mfcc > nm:
        raise InvalidInputError("n_mfcc must be <= n_mels")
    rows = params.n_mfcc - (0 if (params.include_c0 or params.n_mfcc == 1) else 1)
    out = np.empty((nc, rows, nf), dtype=x.dtype)
    _native.check(L.sgx_mfcc_from_log_mel(0 if x.dtype == np.float32 else 1, x.ctypes.data, nc, nm, nf, params.n_mfcc,
                                          int(params.include_c0), params.lifter, out.ctypes.data, -1, None))
    return Mfcc(out[0] if squeeze else out, params)


def mfcc(samples, stft_params: StftParams, sample_rate: float, n_mels: int, mfcc_params: MfccParams, dtype=None) -> Mfcc:
    """``mfcc<T>()`` (src/mfcc.rs:359-379), fused on the GPU."""
    return MfccPlan(stft_params, sample_rate, n_mels, mfcc_params, dtype or _infer_dtype(samples)).compute(samples)


def build_chroma_filterbank(sample_rate: float, n_fft: int, params: ChromaParams) -> np.ndarray:
    """``build_chroma_filterbank`` (src/chroma.rs:279-346): (12, n_fft // 2 + 1) f64, built by the library's host code."""
    n_fft = int(n_fft)
    if n_fft <= 0:
        raise InvalidInputError("n_fft must be set")
    out = np.empty((12, n_fft // 2 + 1), dtype=np.float64)
    _native.check(_native.lib().sgx_chroma_filterbank(float(sample_rate), n_fft, params.tuning, params.f_min, params.f_max,
                                                      out.ctypes.data))
    return out


def chromagram_from_spectrogram(spectrogram, sample_rate: float, n_fft: int, params: ChromaParams) -> Chromagram:
    """``chromagram_from_spectrogram`` (src/chroma.rs:365-404). spectrogram: (n_bins, n_frames) or
    (n_clips, n_bins, n_frames) with n_bins == n_fft // 2 + 1 (``DimensionMismatchError`` otherwise)."""
    L = _native.lib()
    norm = _CHROMA_NORM[params.norm]
    squeeze = False
    if _is_torch(spectrogram):
        torch = _torch()
        x = spectrogram
        if not x.is_cuda:
            raise InvalidInputError("torch inputs must be CUDA tensors (use NumPy arrays for host data)")
        if x.dim() == 2:
            x, squeeze = x.unsqueeze(0), True
        if x.dim() != 3 or x.numel() == 0:
            raise InvalidInputError("spectrogram must be (n_bins, n_frames) or (n_clips, n_bins, n_frames)")
        x = x.contiguous()
        dt = normalise_dtype(x.dtype)
        nc, nb, nf = x.shape
        out = torch.empty((nc, 12, nf), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _native.check(L.sgx_chroma_from_spectrogram(0 if dt == "f32" else 1, x.data_ptr(), nc, nb, nf, float(sample_rate),
                                                        int(n_fft), params.tuning, params.f_min, params.f_max, norm,
                                                        out.data_ptr(), x.device.index,
                                                        torch.cuda.current_stream(x.device).cuda_stream))
        return Chromagram(out[0] if squeeze else out, params)
    x = np.asarray(spectrogram)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    if x.ndim == 2:
        x, squeeze = x[None], True
    if x.ndim != 3 or x.size == 0:
        raise InvalidInputError("spectrogram must be (n_bins, n_frames) or (n_clips, n_bins, n_frames)")
    x = np.ascontiguousarray(x)
    nc, nb, nf = x.shape
    out = np.empty((nc, 12, nf), dtype=x.dtype)
    _native.check(L.sgx_chroma_from_spectrogram(0 if x.dtype == np.float32 else 1, x.ctypes.data, nc, nb, nf, float(sample_rate),
                                                int(n_fft), params.tuning, params.f_min, params.f_max, norm, out.ctypes.data,
                                                -1, None))
    return Chromagram(out[0] if squeeze else out, params)


def chromagram(samples, stft_params: StftParams, sample_rate: float, chroma_params: ChromaParams, dtype=None) -> Chromagram:
    """``chromagram<T>()`` (src/chroma.rs:487-503), fused on the GPU."""
    return ChromaPlan(stft_params, sample_rate, chroma_params, dtype or _infer_dtype(samples)).compute(samples)


def rfft(samples, n_fft: int, dtype=None):
    """free ``fft()`` (:4490-4520): one unnormalised R2C of <= n_fft samples, zero padded; no window."""
    L = _native.lib()
    x = np.ascontiguousarray(samples)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    if dtype is not None:
        x = x.astype(np.float32 if normalise_dtype(dtype) == "f32" else np.float64)
    if x.ndim != 1 or x.size == 0:
        raise InvalidInputError("samples must be a non-empty 1-D array")
    out = np.empty(int(n_fft) // 2 + 1, dtype=np.complex64 if x.dtype == np.float32 else np.complex128)
    _native.check(L.sgx_rfft(0 if x.dtype == np.float32 else 1, x.ctypes.data, x.size, int(n_fft), out.ctypes.data, -1, None))
    return out


fft = rfft


class FftPlanner:
    """``FftPlanner`` (src/spectrogram.rs:4977-5235): reuses FFT plans across calls. The reference caches its backend's plans by
    size; this one caches complete native plans (device tables included) keyed by dtype, n_fft, window and output, so
    repeated single-frame transforms of a size pay neither table construction nor upload. Methods mirror the reference:
    ``fft`` (complex spectrum), ``rfft`` (its *magnitude* -- :5072-5079 maps ``Complex::norm``), ``irfft``, ``power_spectrum``,
    ``magnitude_spectrum``. NumPy in / NumPy out, or CUDA torch tensors in / out."""

    def __init__(self, device: Optional[int] = None):
        self._h = C.c_void_p()
        _native.check(_native.lib().sgx_fft_planner_create(-1 if device is None else int(device), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _native.lib().sgx_fft_planner_destroy(h)
            except Exception:
                pass
            self._h = C.c_void_p()

    def cached_plans(self) -> int:
        return int(_native.lib().sgx_fft_planner_cached_plans(self._h))

    @staticmethod
    def _prep(x, complex_in=False):
        if _is_torch(x):
            torch = _torch()
            ok = (torch.complex64, torch.complex128) if complex_in else (torch.float32, torch.float64)
            if not x.is_cuda or x.dtype not in ok or x.dim() != 1 or x.numel() == 0:
                raise InvalidInputError("expected a non-empty 1-D CUDA tensor of a supported dtype (or a NumPy array)")
            return x.contiguous(), str(x.dtype).endswith(("float32", "complex64"))
        x = np.ascontiguousarray(x)
        if complex_in:
            if x.dtype not in (np.complex64, np.complex128):
                x = x.astype(np.complex128)
        elif x.dtype not in (np.float32, np.float64):
            x = x.astype(np.float64)
        if x.ndim != 1 or x.size == 0:
            raise InvalidInputError("expected a non-empty 1-D array")
        return x, x.dtype in (np.float32, np.complex64)

    @staticmethod
    def _alloc(like, n, f32, cplx):
        if _is_torch(like):
            torch = _torch()
            dt = (torch.complex64 if f32 else torch.complex128) if cplx else (torch.float32 if f32 else torch.float64)
            t = torch.empty(n, dtype=dt, device=like.device)
            return t, t.data_ptr(), torch.cuda.current_stream(like.device).cuda_stream
        dt = (np.complex64 if f32 else np.complex128) if cplx else (np.float32 if f32 else np.float64)
        a = np.empty(n, dtype=dt)
        return a, a.ctypes.data, None

    @staticmethod
    def _ptr(x):
        return x.data_ptr() if _is_torch(x) else x.ctypes.data

    def fft(self, samples, n_fft: int):
        x, f32 = self._prep(samples)
        out, optr, stream = self._alloc(x, int(n_fft) // 2 + 1, f32, True)
        _native.check(_native.lib().sgx_fft_planner_rfft(self._h, 0 if f32 else 1, self._ptr(x), x.shape[0], int(n_fft), optr, stream))
        return out

    def rfft(self, samples, n_fft: int):
        z = self.fft(samples, n_fft)
        return z.abs() if _is_torch(z) else np.abs(z)

    def irfft(self, spectrum, n_fft: int):
        x, f32 = self._prep(spectrum, complex_in=True)
        out, optr, stream = self._alloc(x, int(n_fft), f32, False)
        _native.check(_native.lib().sgx_fft_planner_irfft(self._h, 0 if f32 else 1, self._ptr(x), x.shape[0], int(n_fft), optr, stream))
        return out

    def _spectrum(self, samples, n_fft, window, magnitude):
        x, f32 = self._prep(samples)
        w = _as_window(window) if window is not None else WindowType.rectangular()
        if w.kind == "custom":
            raise InvalidInputError("custom windows are not cached by the planner: build a linear plan instead")
        out, optr, stream = self._alloc(x, int(n_fft) // 2 + 1, f32, False)
        _native.check(_native.lib().sgx_fft_planner_power_spectrum(self._h, 0 if f32 else 1, self._ptr(x), x.shape[0], int(n_fft),
                                                                   _WIN[w.kind], float(w.param), int(magnitude), optr, stream))
        return out

    def power_spectrum(self, samples, n_fft: int, window=None):
        return self._spectrum(samples, n_fft, window, False)

    def magnitude_spectrum(self, samples, n_fft: int, window=None):
        return self._spectrum(samples, n_fft, window, True)


def irfft(spectrum, n_fft: int):
    """``irfft`` (:4789-4811): n_fft/2 + 1 complex bins -> n_fft samples (true inverse of ``rfft``)."""
    L = _native.lib()
    x = np.ascontiguousarray(spectrum)
    if x.dtype not in (np.complex64, np.complex128):
        x = x.astype(np.complex128)
    if x.ndim != 1 or x.size == 0:
        raise InvalidInputError("spectrum must be a non-empty 1-D complex array")
    out = np.empty(int(n_fft), dtype=np.float32 if x.dtype == np.complex64 else np.float64)
    _native.check(L.sgx_irfft(0 if x.dtype == np.complex64 else 1, x.ctypes.data, x.size, int(n_fft), out.ctypes.data, -1, None))
    return out


def istft(stft_matrix, n_fft: int, hop_size: int, window="hanning", center: bool = True):
    """``istft<T>()`` (:4813-4911): overlap-add reconstruction from a complex (n_fft/2+1, n_frames) matrix (or a batch of
    them); builds a fresh plan like ``stft()`` does. dtype follows the matrix (complex64 -> float32)."""
    if _is_torch(stft_matrix):
        dt = "float32" if str(stft_matrix.dtype).endswith("complex64") else "float64"
        device = stft_matrix.device.index
    else:
        stft_matrix = np.asarray(stft_matrix)
        dt = "float32" if stft_matrix.dtype == np.complex64 else "float64"
        device = None
    if int(hop_size) > int(n_fft):
        raise InvalidInputError("hop_size must be <= n_fft")                   # :4829-4831
    params = SpectrogramParams(StftParams(n_fft, hop_size, window, center), 1.0)
    return StftPlan(params, dt, device).istft(stft_matrix)


def power_spectrum(samples, n_fft: int, window: Optional[WindowType] = None, dtype=None):
    """``power_spectrum`` (:4611-4643) / ``SpectrogramPlanner::compute_power_spectrum`` (:771-816): |X|^2 of one
    (optionally windowed, zero padded) frame."""
    x = np.ascontiguousarray(samples)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    if x.size > int(n_fft):
        raise InvalidInputError(f"Input length ({x.size}) exceeds FFT size ({int(n_fft)})")
    params = SpectrogramParams(StftParams(n_fft, n_fft, window or WindowType.rectangular(), False), 1.0)
    plan = SpectrogramPlanner().linear_plan(params, None, "power", dtype or _infer_dtype(x))
    return plan.compute_frame(x, 0)


def magnitude_spectrum(samples, n_fft: int, window: Optional[WindowType] = None, dtype=None):
    """``magnitude_spectrum`` (:4684-4693): sqrt of the power spectrum."""
    return np.sqrt(power_spectrum(samples, n_fft, window, dtype))


def _one_shot(mapping, amp):
    def f(samples, params: SpectrogramParams, scale_params=None, db_params: Optional[LogParams] = None, dtype=None):
        pl = SpectrogramPlanner()
        dt = dtype or _infer_dtype(samples)
        if mapping == "linear":
            plan = pl.linear_plan(params, db_params, amp, dt)
        elif mapping == "mel":
            plan = pl.mel_plan(params, scale_params, db_params, amp, dt)
        elif mapping == "erb":
            plan = pl.erb_plan(params, scale_params, db_params, amp, dt)
        else:
            plan = pl.log_hz_plan(params, scale_params, db_params, amp, dt)
        return plan.compute(samples)
    f.__doc__ = f"One-shot ``Spectrogram::<{mapping}, {amp}>::compute`` (:2887-3022): builds a plan and runs it."
    return f


def compute_linear_power_spectrogram(samples, params, dtype=None): return _one_shot("linear", "power")(samples, params, None, None, dtype)
def compute_linear_magnitude_spectrogram(samples, params, dtype=None): return _one_shot("linear", "magnitude")(samples, params, None, None, dtype)
def compute_linear_db_spectrogram(samples, params, db_params=None, dtype=None): return _one_shot("linear", "db")(samples, params, None, db_params, dtype)
def compute_mel_power_spectrogram(samples, params, mel_params, dtype=None): return _one_shot("mel", "power")(samples, params, mel_params, None, dtype)
def compute_mel_magnitude_spectrogram(samples, params, mel_params, dtype=None): return _one_shot("mel", "magnitude")(samples, params, mel_params, None, dtype)
def compute_mel_db_spectrogram(samples, params, mel_params, db_params=None, dtype=None): return _one_shot("mel", "db")(samples, params, mel_params, db_params, dtype)
def compute_erb_power_spectrogram(samples, params, erb_params, dtype=None): return _one_shot("erb", "power")(samples, params, erb_params, None, dtype)
def compute_erb_magnitude_spectrogram(samples, params, erb_params, dtype=None): return _one_shot("erb", "magnitude")(samples, params, erb_params, None, dtype)
def compute_erb_db_spectrogram(samples, params, erb_params, db_params=None, dtype=None): return _one_shot("erb", "db")(samples, params, erb_params, db_params, dtype)
def compute_loghz_power_spectrogram(samples, params, loghz_params, dtype=None): return _one_shot("loghz", "power")(samples, params, loghz_params, None, dtype)
def compute_loghz_magnitude_spectrogram(samples, params, loghz_params, dtype=None): return _one_shot("loghz", "magnitude")(samples, params, loghz_params, None, dtype)
def compute_loghz_db_spectrogram(samples, params, loghz_params, db_params=None, dtype=None): return _one_shot("loghz", "db")(samples, params, loghz_params, db_params, dtype)
def compute_mfcc(samples, stft_params, sample_rate, n_mels, mfcc_params, dtype=None): return mfcc(samples, stft_params, sample_rate, n_mels, mfcc_params, dtype)
def compute_chromagram(samples, stft_params, sample_rate, chroma_params, dtype=None): return chromagram(samples, stft_params, sample_rate, chroma_params, dtype)
