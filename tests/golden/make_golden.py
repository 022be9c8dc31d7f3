"""Generate golden vectors from the reference's own NumPy restatement.

Run in the build container only (needs /root/reference): ``python tests/golden/make_golden.py``.
It imports ``python/examples/numpy_impls.py`` from the reference checkout *unmodified* (stft, hann_window,
power/magnitude/db_spectrogram -- the functions whose semantics match the Rust crate: zero centre padding,
symmetric Hann, unnormalised rFFT; its mel/ERB helpers use a different filterbank and are NOT used) and stores
inputs' recipes + outputs in ``ref_numpy_impls.npz``. The GPU box never reads /root/reference; tests read the npz.

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


def main():
    spec = importlib.util.spec_from_file_location("ref_numpy_impls", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
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
    np.savez_compressed(os.path.join(HERE, "ref_numpy_impls.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
