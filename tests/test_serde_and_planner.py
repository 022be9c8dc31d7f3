"""serde wire format (src/spectrogram.rs:2546-2557, tests/serde_tests.rs) and FftPlanner (src/spectrogram.rs:4977-5235)."""
import json

import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from spectrograms_b200 import serde
from conftest import make_signal, rel_l2


def _spec(rows=5, cols=3, seed=0, params=None):
    rng = np.random.default_rng(seed)
    params = params or sg.SpectrogramParams(sg.StftParams(512, 256, sg.WindowType.hanning(), True), 16000.0)
    return sg.Spectrogram(rng.standard_normal((rows, cols)), np.linspace(0.0, 8000.0, rows), np.arange(cols) * 0.016, params, "linear", "power")


def test_spectrogram_json_field_names_match_the_derive():
    """Field names / nesting of derive(Serialize): Spectrogram{data, axes, params} (:2546-2557; _amp skipped),
    Axes{freq, times} (:3305), FrequencyAxis{frequencies} (:3242), SpectrogramParams{stft, sample_rate_hz} (:4107),
    StftParams{n_fft, hop_size, window, centre} (:4051), Array2 = {v, dim, data} (ndarray serde)."""
    s = _spec()
    d = json.loads(s.to_json())
    assert list(d) == ["data", "axes", "params"]
    assert list(d["data"]) == ["v", "dim", "data"] and d["data"]["v"] == 1 and d["data"]["dim"] == [5, 3]
    assert d["data"]["data"] == [float(v) for v in s.data.reshape(-1)]                     # row-major
    assert list(d["axes"]) == ["freq", "times"] and list(d["axes"]["freq"]) == ["frequencies"]
    assert d["params"] == {"stft": {"n_fft": 512, "hop_size": 256, "window": "Hanning", "centre": True}, "sample_rate_hz": 16000.0}
    assert " " not in s.to_json()                                                           # compact, like serde_json::to_string


def test_spectrogram_json_round_trip():
    # tests/serde_tests.rs:45-65 : n_bins, n_frames, data (1e-10), axis lengths survive the round trip
    s = _spec(257, 63, seed=1)
    r = sg.Spectrogram.from_json(s.to_json())
    assert r.n_bins == s.n_bins and r.n_frames == s.n_frames
    assert np.abs(r.data - s.data).max() < 1e-10 and np.array_equal(r.data, s.data)        # shortest-repr floats: exact
    assert len(r.frequencies) == len(s.frequencies) and len(r.times) == len(s.times)
    assert r.params.stft.n_fft == 512 and r.params.sample_rate == 16000.0
    r32 = sg.Spectrogram.from_json(s.to_json(), "mel", "db", np.float32)                   # from_str::<MelDbSpectrogram<f32>>
    assert r32.data.dtype == np.float32 and r32.freq_scale == "mel" and r32.amp_scale == "db"


@pytest.mark.parametrize("window,obj", [
    (sg.WindowType.rectangular(), "Rectangular"), (sg.WindowType.hamming(), "Hamming"), (sg.WindowType.blackman(), "Blackman"),
    (sg.WindowType.kaiser(8.6), {"Kaiser": {"beta": 8.6}}), (sg.WindowType.gaussian(40.0), {"Gaussian": {"std": 40.0}}),
    (sg.WindowType.custom([0.25, 0.5, 1.0, 0.5]), {"Custom": {"coefficients": [0.25, 0.5, 1.0, 0.5], "size": 4}})])
def test_window_type_is_an_externally_tagged_enum(window, obj):
    assert serde.window_to_obj(window) == obj                                              # src/window.rs:17-51
    back = serde.window_from_obj(json.loads(json.dumps(obj)))
    assert back.kind == window.kind and back.param == window.param and back.coefficients == window.coefficients


def test_parameter_structs_round_trip():
    cases = [
        (sg.MelParams(80, 20.0, 7600.0, "slaney"), serde.mel_params_to_dict, serde.mel_params_from_dict,
         {"n_mels": 80, "f_min": 20.0, "f_max": 7600.0, "norm": "Slaney"}),
        (sg.LogHzParams(84, 32.7, 8000.0), serde.loghz_params_to_dict, serde.loghz_params_from_dict, {"n_bins": 84, "f_min": 32.7, "f_max": 8000.0}),
        (sg.ErbParams(40, 50.0, 8000.0, "apple_tr35"), serde.erb_params_to_dict, serde.erb_params_from_dict,
         {"n_filters": 40, "f_min": 50.0, "f_max": 8000.0, "spacing": "AppleTr35", "db_floor": None}),
        (sg.LogParams(-80.0), serde.log_params_to_dict, serde.log_params_from_dict, {"floor_db": -80.0}),
        (sg.MfccParams(20, False, 0), serde.mfcc_params_to_dict, serde.mfcc_params_from_dict, {"n_mfcc": 20, "include_c0": False, "lifter": 0}),
        (sg.ChromaParams.music_standard(), serde.chroma_params_to_dict, serde.chroma_params_from_dict,
         {"tuning": 440.0, "n_octaves": 7, "f_min": 32.7, "f_max": 4186.0, "norm": "L2"}),
    ]
    for obj, to_d, from_d, want in cases:
        assert to_d(obj) == want
        assert to_d(from_d(json.loads(serde.to_json(obj)))) == want


def test_mfcc_and_chromagram_json():
    # tests/serde_tests.rs:193-260 : Mfcc{data, params}, Chromagram{data, params}
    m = sg.Mfcc(np.random.default_rng(2).standard_normal((13, 7)), sg.MfccParams())
    d = json.loads(serde.to_json(m))
    assert list(d) == ["data", "params"] and d["params"] == {"n_mfcc": 13, "include_c0": True, "lifter": 22}
    assert np.array_equal(serde.mfcc_from_json(serde.to_json(m)).data, m.data)
    c = sg.Chromagram(np.random.default_rng(3).random((12, 9)), sg.ChromaParams())
    r = serde.chromagram_from_json(serde.to_json(c))
    assert np.array_equal(r.data, c.data) and r.params.norm == "l2"


def test_malformed_documents_are_rejected():
    s = json.loads(_spec().to_json())
    bad = dict(s, data=dict(s["data"], dim=[4, 3]))
    with pytest.raises(sg.InvalidInputError):
        serde.spectrogram_from_dict(bad)
    bad = dict(s, axes={"freq": {"frequencies": [0.0]}, "times": s["axes"]["times"]})
    with pytest.raises(sg.InvalidInputError):
        serde.spectrogram_from_dict(bad)
    with pytest.raises(sg.InvalidInputError):
        serde.window_from_obj("Hann")
    with pytest.raises(sg.InvalidInputError):                                               # params are re-validated (hop > n_fft)
        serde.stft_params_from_dict({"n_fft": 256, "hop_size": 512, "window": "Hanning", "centre": True})


# ------------------------------------------------------------------------------------------------ FftPlanner (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fft_planner_matches_oracle_and_reuses_plans(dtype):
    pl = sg.FftPlanner()
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert pl.cached_plans() == 0
    for n_fft in (512, 400, 1000, 512, 400):                       # pow2, n400-sized, generic; then cache hits
        x = make_signal("noise", n_fft - 7, 16000.0, dtype, seed=n_fft)
        z = pl.fft(x, n_fft)
        ref = oracle.rfft(x.astype(np.float64), n_fft)
        assert z.shape == (n_fft // 2 + 1,) and z.dtype == (np.complex128 if dtype == np.float64 else np.complex64)
        assert rel_l2(z, ref) <= tol
        assert rel_l2(pl.rfft(x, n_fft), np.abs(ref)) <= tol       # FftPlanner::rfft returns magnitudes (:5072-5079)
        back = pl.irfft(z, n_fft)
        assert rel_l2(back[: x.size], x) <= tol * 10 and np.abs(back[x.size:]).max() <= tol * 50
        p = pl.power_spectrum(x, n_fft, sg.WindowType.hanning())
        full = np.zeros(n_fft)
        full[: x.size] = x
        refp = np.abs(np.fft.rfft(full * oracle.Plan(oracle.Desc(n_fft=n_fft, hop=n_fft)).window())) ** 2
        assert rel_l2(p, refp) <= tol
        assert rel_l2(pl.magnitude_spectrum(x, n_fft), np.abs(np.fft.rfft(full))) <= tol
    n = pl.cached_plans()
    assert n == 9                                                  # 3 sizes x (complex, hann power, rect magnitude); repeats hit the cache
    with pytest.raises(sg.InvalidInputError, match="exceeds FFT size"):
        pl.fft(np.ones(9, dtype=dtype), 8)
    with pytest.raises(sg.DimensionMismatchError):
        pl.irfft(np.ones(5, dtype=np.complex128), 16)
    assert pl.cached_plans() == n


@pytest.mark.gpu
def test_fft_planner_device_tensors():
    import torch
    pl = sg.FftPlanner()
    x = torch.randn(512, dtype=torch.float32, device="cuda")
    z = pl.fft(x, 512)
    assert z.is_cuda and z.dtype == torch.complex64
    assert rel_l2(z.cpu().numpy(), np.fft.rfft(x.cpu().numpy().astype(np.float64))) <= 1e-5
    back = pl.irfft(z, 512)
    assert back.is_cuda and rel_l2(back.cpu().numpy(), x.cpu().numpy()) <= 1e-5
