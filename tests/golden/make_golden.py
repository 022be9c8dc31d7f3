"""Generate golden vectors that pin the oracle (and, directly, the CUDA path) to code the builder did not write.

Run in the build container only (needs /root/reference and torchaudio): ``python tests/golden/make_golden.py``.
The GPU box never reads /root/reference; tests read the committed npz files.

``ref_numpy_impls.npz`` -- outputs of the reference's own NumPy restatement ``python/examples/numpy_impls.py``, imported
*unmodified*:
  * stft / hann_window / power / magnitude / dB (zero centre padding, symmetric Hann, unnormalised rFFT: the crate's
    semantics);
  * ERB: ``erb_centers`` + ``gammatone_response`` + ``erb_spectrogram`` (numpy_impls.py:128-159) are the crate's
    ``ErbFilterbank::generate`` / ``apply_to_power_spectrum`` (src/erb.rs:266-402) with ``ErbSpacing::Linear``: centres
    linear on the ERB-bandwidth scale, |1 / (1 + j (f - fc) / (1.019 erb(fc)))^4|^2 summed against the power spectrum;
  * LogHz: ``log_frequency_matrix`` + ``logfreq_spectrogram`` (numpy_impls.py:94-121) are ``build_loghz_matrix``
    (src/spectrogram.rs:2438-2508): log-spaced centres, linear interpolation between floor/ceil bins (the two differ
    only when ceil(f / df) reaches the last bin, which the cases below avoid).
  The file's *mel* helper is HTK with floored bin edges and its chroma is a hard assignment -- different algorithms from
  the crate's Slaney / frequency-space triangles, so those two are NOT used.

``third_party_pins.npz`` -- independent third-party implementations of the algorithms the crate names:
  * mel filterbank: ``torchaudio.functional.melscale_fbanks(mel_scale="slaney", norm=None | "slaney")`` -- the librosa
    (htk=False) construction the crate says it matches (src/spectrogram.rs:2262-2432). f32, so values agree to its
    rounding (2e-5 abs); the non-zero pattern must agree exactly.
  * mel power / dB spectrograms = that matrix (f64) @ the reference NumPy power spectrogram.
  * DCT-II: ``scipy.fft.dct(type=2) / 2`` = sum_i x[i] cos(pi k (i + 0.5) / n) (src/mfcc.rs:278-292), and MFCCs = that
    DCT of the pinned log-mel with the lifter of src/mfcc.rs:296-316 written out in NumPy.

Signals follow SURVEY.md section 8(d): sine (tests/spectrogram_tests.rs:10-16), chirp (notebook cell 1),
noise (python/tests/test_dtype_planner.py:18-19).
"""
import importlib.util
import os

import numpy as np

REF = "/root/reference/python/examples/numpy_impls.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def signals(sr, n):
    t = np.arange(n) / sr
    return {
        "sine": np.sin(2.0 * np.pi * 440.0 * np.arange(n) / sr),
        "chirp": np.sin(2.0 * np.pi * (100.0 + 3000.0 * t * t) * t),
        "noise": np.random.default_rng(0).standard_normal(n),
    }


CASES = [  # (name, sr, n_samples, n_fft, hop, centre)
    ("c1", 16000, 16000, 512, 256, True),          # BASELINE.json configs[0] / examples/basic_linear.rs
    ("whisper", 16000, 8000, 400, 160, True),      # configs[1] framing
    ("nocentre", 16000, 4000, 256, 64, False),
    ("odd", 8000, 3001, 250, 100, True),
]

# (name, sr, n_samples, n_fft, hop, n_filters, f_min, f_max)
ERB_CASES = [
    ("erb512", 16000, 8000, 512, 160, 40, 50.0, 8000.0),
    ("erb400", 16000, 6000, 400, 160, 64, 100.0, 7600.0),      # the n400 kernel family, dense projection
]
LOGHZ_CASES = [
    ("loghz1024", 22050, 12000, 1024, 256, 84, 32.7, 8000.0),
    ("loghz400", 16000, 6000, 400, 160, 48, 60.0, 7000.0),
]
# (name, sr, n_samples, n_fft, hop, n_mels, f_min, f_max, norm)
MEL_CASES = [
    ("mel400", 16000, 8000, 400, 160, 128, 0.0, 8000.0, None),           # configs[1] / configs[3]
    ("mel2048", 22050, 16000, 2048, 512, 128, 0.0, 11025.0, None),       # configs[2]
    ("mel512", 16000, 8000, 512, 160, 64, 0.0, 8000.0, "slaney"),
]
MFCC = dict(n_mfcc=40, lifter=22)                                       # configs[3]; lifter 22 = MfccParams default


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_numpy_impls", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def main_reference(ref):
    out = {}
    for name, sr, n, n_fft, hop, centre in CASES:
        win = ref.hann_window(n_fft)
        out[f"{name}/window"] = win
        for sname, x in signals(sr, n).items():
            s, freqs, _times = ref.stft(x, sr, n_fft, hop, win, centre)
            p = ref.power_spectrogram(s)
            out[f"{name}/{sname}/stft"] = s.astype(np.complex128)
            if name == "c1":      # amplitude scalings only for the C1 case (keeps the fixture small)
                out[f"{name}/{sname}/power"] = p
                out[f"{name}/{sname}/magnitude"] = ref.magnitude_spectrogram(s)
                out[f"{name}/{sname}/db80"] = ref.db_spectrogram(p, eps=10.0 ** (-80.0 / 10.0))
        out[f"{name}/freqs"] = freqs
        out[f"{name}/meta"] = np.array([sr, n, n_fft, hop, int(centre)], dtype=np.int64)
    for name, sr, n, n_fft, hop, nb, f_min, f_max in ERB_CASES:
        win = ref.hann_window(n_fft)
        centres = ref.erb_centers(f_min, f_max, nb)
        out[f"{name}/centres"] = centres
        out[f"{name}/meta"] = np.array([sr, n, n_fft, hop, nb, f_min, f_max], dtype=np.float64)
        for sname, x in signals(sr, n).items():
            s, freqs, _ = ref.stft(x, sr, n_fft, hop, win, True)
            out[f"{name}/{sname}/power"] = ref.erb_spectrogram(s, freqs, centres)
    for name, sr, n, n_fft, hop, nb, f_min, f_max in LOGHZ_CASES:
        win = ref.hann_window(n_fft)
        M = ref.log_frequency_matrix(sr, n_fft, nb, f_min, f_max)
        out[f"{name}/matrix"] = M
        out[f"{name}/meta"] = np.array([sr, n, n_fft, hop, nb, f_min, f_max], dtype=np.float64)
        for sname, x in signals(sr, n).items():
            s, _freqs, _ = ref.stft(x, sr, n_fft, hop, win, True)
            out[f"{name}/{sname}/power"] = ref.logfreq_spectrogram(ref.power_spectrogram(s), M)
    np.savez_compressed(os.path.join(HERE, "ref_numpy_impls.npz"), **out)
    print("ref_numpy_impls.npz:", len(out), "arrays")


def lifter_weights(n_mfcc, lifter):
    i = np.arange(n_mfcc, dtype=np.float64)
    return (lifter / 2.0) * np.sin(np.pi * i / lifter) + 1.0


def main_third_party(ref):
    import scipy.fft
    import torchaudio

    out = {}
    for name, sr, n, n_fft, hop, n_mels, f_min, f_max, norm in MEL_CASES:
        fb32 = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sr, norm=norm,
                                                     mel_scale="slaney").numpy().T.copy()      # (n_mels, n_bins) f32
        out[f"{name}/fb"] = fb32
        out[f"{name}/meta"] = np.array([sr, n, n_fft, hop, n_mels, f_min, f_max, 0 if norm is None else 1], dtype=np.float64)
        fb = fb32.astype(np.float64)
        win = ref.hann_window(n_fft)
        for sname, x in signals(sr, n).items():
            s, _f, _t = ref.stft(x, sr, n_fft, hop, win, True)
            mel = fb @ ref.power_spectrogram(s)
            out[f"{name}/{sname}/power"] = mel
            if name == "mel400":
                db = ref.db_spectrogram(mel, eps=10.0 ** (-80.0 / 10.0))
                out[f"{name}/{sname}/db80"] = db
                c = scipy.fft.dct(db, type=2, axis=0, norm=None)[: MFCC["n_mfcc"]] / 2.0
                out[f"{name}/{sname}/mfcc40"] = c * lifter_weights(MFCC["n_mfcc"], MFCC["lifter"])[:, None]
    rng = np.random.default_rng(7)
    for n in (13, 40, 64, 128):
        x = rng.standard_normal((n, 9)) * 30.0 - 40.0             # log-mel-like magnitudes
        out[f"dct/{n}/in"] = x
        out[f"dct/{n}/out"] = scipy.fft.dct(x, type=2, axis=0, norm=None) / 2.0
    np.savez_compressed(os.path.join(HERE, "third_party_pins.npz"), **out)
    print("third_party_pins.npz:", len(out), "arrays")


if __name__ == "__main__":
    r = load_ref()
    main_reference(r)
    main_third_party(r)
