"""Independent NumPy/SciPy restatement of the reference hot path -- TEST INFRASTRUCTURE ONLY.

Written separately from oracle.c (vectorised, pocketfft for the DFT) so that the two can check each other.
Citations are relative to the reference checkout; bare ``:N`` means ``src/spectrogram.rs:N``.
Everything is computed in float64 unless ``dtype`` says float32, in which case the data path (window product,
FFT, power, filterbank, scaling) runs in float32 like the reference's ``T = f32`` instantiation.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft


def frame_count(n, n_fft, hop, centre):                      # :1230-1250
    pad = n_fft // 2 if centre else 0
    padded = n + 2 * pad
    if padded < n_fft:
        return 1
    return (padded - n_fft) // hop + 1


def _i0(x):                                                    # :2237-2259
    ax = abs(x)
    if ax <= 3.75:
        t2 = (x / 3.75) ** 2
        return 1.0 + t2 * (3.5156229 + t2 * (3.0899424 + t2 * (1.2067492 + t2 * (0.2659732 + t2 * (0.0360768 + t2 * 0.0045813)))))
    t = 3.75 / ax
    poly = 0.39894228 + t * (0.01328592 + t * (0.00225319 + t * (-0.00157565 + t * (0.00916281 + t * (
        -0.02057706 + t * (0.02635537 + t * (-0.01647633 + t * 0.00392377)))))))
    return (math.exp(ax) / (math.sqrt(ax) * math.sqrt(2.0 * math.pi))) * poly


def make_window(kind, n, param=0.0, custom=None):              # :2159-2235 (f64)
    i = np.arange(n, dtype=np.float64)
    if kind == "rectangular":
        return np.ones(n)
    if kind == "hanning":
        return 0.5 - 0.5 * np.cos(2.0 * np.pi * i / (n - 1))
    if kind == "hamming":
        return 0.54 - 0.46 * np.cos(2.0 * np.pi * i / (n - 1))
    if kind == "blackman":
        a = 2.0 * np.pi * i / (n - 1)
        return 0.42 - 0.5 * np.cos(a) + 0.08 * np.cos(2.0 * a)
    if kind == "kaiser":
        if n == 1:
            return np.ones(1)
        nmax = (n - 1) / 2.0
        ratio = np.maximum(1.0 - ((i - nmax) / nmax) ** 2, 0.0)
        d = _i0(param)
        return np.array([_i0(param * math.sqrt(r)) / d for r in ratio])
    if kind == "gaussian":
        return np.exp(-0.5 * ((i - (n - 1) / 2.0) / param) ** 2)
    if kind == "custom":
        w = np.asarray(custom, dtype=np.float64)
        assert w.size == n
        return w.copy()
    raise ValueError(kind)


def hz_to_mel(hz):                                             # :2268-2281
    hz = np.asarray(hz, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(hz >= 1000.0, 15.0 + np.log(np.maximum(hz, 1e-300) / 1000.0) / 0.06875177742094923, hz / (200.0 / 3.0))


def mel_to_hz(mel):                                            # :2286-2300
    mel = np.asarray(mel, dtype=np.float64)
    return np.where(mel >= 15.0, 1000.0 * np.exp(0.06875177742094923 * (mel - 15.0)), (200.0 / 3.0) * mel)


def mel_filterbank(sr, n_fft, n_mels, f_min, f_max, norm="none"):   # :2302-2432, dense (n_mels, out_len)
    out_len = n_fft // 2 + 1
    df = sr / n_fft
    mel_pts = hz_to_mel(f_min) + np.arange(n_mels + 2) * ((hz_to_mel(f_max) - hz_to_mel(f_min)) / (n_mels + 1))
    hz_pts = mel_to_hz(mel_pts)
    bins = np.arange(out_len) * df
    fb = np.zeros((n_mels, out_len))
    for m in range(n_mels):
        fl, fc, fr = hz_pts[m], hz_pts[m + 1], hz_pts[m + 2]
        if fc - fl == 0.0 or fr - fc == 0.0:
            continue
        w = np.clip(np.minimum((bins - fl) / (fc - fl), (fr - bins) / (fr - fc)), 0.0, 1.0)
        w[~(w > 1e-10)] = 0.0                                  # SparseMatrix::set threshold :83
        fb[m] = w
    if norm == "slaney":
        fb *= (2.0 / (mel_to_hz(mel_pts[2:]) - mel_to_hz(mel_pts[:-2])))[:, None]
    elif norm == "l1":
        s = fb.sum(axis=1, keepdims=True)
        fb = np.where(s > 0, fb / np.where(s > 0, s, 1.0), fb)
    elif norm == "l2":
        s = np.sqrt((fb ** 2).sum(axis=1, keepdims=True))
        fb = np.where(s > 0, fb / np.where(s > 0, s, 1.0), fb)
    return fb


def mel_band_centres(sr, n_mels):                              # :2510-2530 (ignores f_min/f_max, quirk F9)
    mmax = hz_to_mel(sr * 0.5)
    return mel_to_hz((np.arange(n_mels) + 1.0) * (mmax / (n_mels + 1)))


def loghz_matrix(sr, n_fft, n_bins, f_min, f_max):             # :2438-2508
    out_len = n_fft // 2 + 1
    df = sr / n_fft
    freqs = np.exp(math.log(f_min) + np.arange(n_bins) * ((math.log(f_max) - math.log(f_min)) / (n_bins - 1)))
    M = np.zeros((n_bins, out_len))
    for i, f in enumerate(freqs):
        exact = f / df
        lo = int(math.floor(exact))
        hi = min(int(math.ceil(exact)), out_len - 1)
        if lo >= out_len:
            continue
        if lo == hi:
            M[i, lo] = 1.0
        else:
            frac = exact - lo
            if abs(1.0 - frac) > 1e-10:
                M[i, lo] = 1.0 - frac
            if abs(frac) > 1e-10:
                M[i, hi] = frac
    return M, freqs


def erb_filterbank(sr, n_fft, n_filters, f_min, f_max, spacing="linear"):   # src/erb.rs:266-332
    out_len = n_fft // 2 + 1
    if spacing == "linear":
        erb = lambda f: 24.7 * (4.37 * f / 1000.0 + 1.0)
        e = erb(f_min) + np.arange(n_filters) * ((erb(f_max) - erb(f_min)) / (n_filters - 1))
        cf = (e / 24.7 - 1.0) * 1000.0 / 4.37
    else:                                                      # src/erb.rs:221-236
        shift = 9.26449 * 24.7
        ee = (math.log(f_min + shift) - math.log(f_max + shift)) / n_filters
        cf = (-shift + np.exp((np.arange(n_filters) + 1.0) * ee) * (f_max + shift))[::-1]
    freqs = np.arange(out_len) * (sr / n_fft)
    bw = 1.019 * 24.7 * (4.37 * cf / 1000.0 + 1.0)
    x = (freqs[None, :] - cf[:, None]) / bw[:, None]
    resp = 1.0 / np.abs((1.0 + 1j * x) ** 4) ** 2
    return resp, cf


def frames(x, n_fft, hop, centre, window):                     # :1301-1320, returns (n_frames, n_fft) in x.dtype
    dt = x.dtype
    pad = n_fft // 2 if centre else 0
    nf = frame_count(x.size, n_fft, hop, centre)
    need = (nf - 1) * hop + n_fft
    buf = np.zeros(max(need, x.size + 2 * pad), dtype=dt)
    buf[pad:pad + x.size] = x
    idx = np.arange(nf)[:, None] * hop + np.arange(n_fft)[None, :]
    return buf[idx] * window.astype(dt)[None, :]


def stft(x, n_fft, hop, centre, window):                       # StftPlan::compute :1424-1458 -> (bins, frames)
    fr = frames(x, n_fft, hop, centre, window)
    return scipy.fft.rfft(fr, axis=-1).T                        # pocketfft keeps float32 input in float32


def spectrogram(x, n_fft, hop, centre, window, mapping=None, amp="power", floor_db=None):
    """mapping: None (identity) or a dense (n_bins, out_len) float64 matrix. Returns (n_bins, n_frames)."""
    dt = x.dtype
    s = stft(x, n_fft, hop, centre, window)
    p = (s.real * s.real + s.imag * s.imag).astype(dt)          # norm_sqr :1332-1334
    if mapping is not None:
        p = (mapping.astype(dt) @ p).astype(dt)                 # T(w) * x summed in T (:102-117); order differs
    if amp == "magnitude":
        p = np.sqrt(p)                                          # :2001
    if amp == "db" and floor_db is not None:                    # :2018-2036 ; F7: no floor -> raw power
        eps = dt.type(10.0 ** (floor_db / 10.0))
        p = (dt.type(10.0) * np.log10(np.maximum(p, eps))).astype(dt)
    return p


def mfcc_from_log_mel(log_mel, n_mfcc, include_c0=True, lifter=22):   # src/mfcc.rs:224-316
    dt = log_mel.dtype
    n_mels = log_mel.shape[0]
    k = np.arange(n_mfcc, dtype=np.float64)[:, None]
    i = np.arange(n_mels, dtype=np.float64)[None, :]
    basis = np.cos(np.pi * k * (i + 0.5) / n_mels).astype(dt)
    out = (basis @ log_mel).astype(dt)
    if lifter > 0:
        w = (1.0 + (lifter / 2.0) * np.sin(np.pi * np.arange(n_mfcc) / lifter)).astype(dt)
        out = out * w[:, None]
    if not include_c0 and n_mfcc > 1:
        out = out[1:]
    return out


def chroma_filterbank(sample_rate, n_fft, tuning=440.0, f_min=32.7, f_max=4186.0):   # src/chroma.rs:279-346, vectorised
    n_bins = n_fft // 2 + 1
    freqs = np.arange(n_bins, dtype=np.float64) * (sample_rate / n_fft)
    fb = np.zeros((12, n_bins))
    ok = (freqs >= f_min) & (freqs <= f_max) & (freqs > 0.0)
    midi = 69.0 + 12.0 * np.log(freqs[ok] / tuning) / np.log(2.0)
    pc = np.mod(midi, 12.0)
    dist = np.abs(pc[None, :] - np.arange(12, dtype=np.float64)[:, None])
    circ = np.minimum(dist, 12.0 - dist)
    fb[:, ok] = np.exp(-0.5 * circ ** 2)
    s = fb.sum(axis=1, keepdims=True)
    return np.where(s > 0.0, fb / np.where(s > 0.0, s, 1.0), fb)


def chroma_from_spectrogram(spec, sample_rate, n_fft, tuning=440.0, f_min=32.7, f_max=4186.0, norm="l2"):   # :365-453
    dt = spec.dtype
    c = (chroma_filterbank(sample_rate, n_fft, tuning, f_min, f_max).astype(dt) @ spec).astype(dt)   # summation order differs
    if norm == "l1":
        d = c.sum(axis=0, keepdims=True)
    elif norm == "l2":
        d = np.sqrt((c * c).sum(axis=0, keepdims=True))
    elif norm == "max":
        d = np.maximum(c.max(axis=0, keepdims=True), 0)
    else:
        return c
    return np.where(d > 0, c / np.where(d > 0, d, 1), c).astype(dt)
