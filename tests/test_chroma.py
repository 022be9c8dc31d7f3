"""Chroma widening (reference src/chroma.rs): oracle cross-checks and host-side API on CPU,
GPU parity of the fused chromagram() plan and of chromagram_from_spectrogram through the C ABI.

The reference holds no golden vectors for chroma (its tests assert shapes and labels only); the C oracle restates
src/chroma.rs line by line and is cross-checked against an independent vectorised NumPy restatement -- "parity
unpinned" beyond that, as for the mel / ERB numerics."""
import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp
import spectrograms_b200 as sg
from conftest import make_signal as signal, rel_l2

NORMS = ["none", "l1", "l2", "max"]


# ------------------------------------------------------------------------------------------------- CPU: oracle + host
def test_filterbank_rows_sum_to_one_and_respect_range():
    fb = oracle.chroma_filterbank(22050.0, 2048)
    assert fb.shape == (12, 1025)
    np.testing.assert_allclose(fb.sum(axis=1), 1.0, rtol=0, atol=1e-12)            # rows to unit sum (src/chroma.rs:336-343)
    freqs = np.arange(1025) * 22050.0 / 2048
    assert np.all(fb[:, (freqs < 32.7) | (freqs > 4186.0)] == 0.0)                # :308-310
    # a bin at A4 weighs most on pitch class 9 (A): midi 69 -> 69 mod 12 = 9 (:313-316)
    k = int(round(440.0 / (22050.0 / 2048)))
    assert int(np.argmax(fb[:, k] / fb[:, k].sum())) == 9


def test_c_oracle_matches_numpy_restatement():
    np.testing.assert_allclose(oracle.chroma_filterbank(16000.0, 1024, 442.0, 55.0, 3000.0),
                               onp.chroma_filterbank(16000.0, 1024, 442.0, 55.0, 3000.0), rtol=1e-13, atol=1e-300)
    rng = np.random.default_rng(3)
    for dt, tol in ((np.float64, 1e-13), (np.float32, 2e-6)):
        spec = np.abs(rng.standard_normal((513, 40))).astype(dt)
        for norm in NORMS:
            a = oracle.chroma_from_spectrogram(spec, 16000.0, 1024, norm=norm)
            b = onp.chroma_from_spectrogram(spec, 16000.0, 1024, norm=norm)
            assert a.dtype == dt and a.shape == (12, 40)
            assert rel_l2(a, b) < tol, (dt, norm)


def test_normalisations_have_their_property():
    spec = np.abs(np.random.default_rng(5).standard_normal((1025, 30)))
    c1 = oracle.chroma_from_spectrogram(spec, 22050.0, 2048, norm="l1")
    c2 = oracle.chroma_from_spectrogram(spec, 22050.0, 2048, norm="l2")
    cm = oracle.chroma_from_spectrogram(spec, 22050.0, 2048, norm="max")
    np.testing.assert_allclose(c1.sum(axis=0), 1.0, atol=1e-12)
    np.testing.assert_allclose(np.sqrt((c2 * c2).sum(axis=0)), 1.0, atol=1e-12)
    np.testing.assert_allclose(cm.max(axis=0), 1.0, atol=1e-12)
    # an all-zero frame stays zero instead of dividing by zero (:412, :424, :437)
    spec[:, 3] = 0.0
    assert np.all(oracle.chroma_from_spectrogram(spec, 22050.0, 2048, norm="l2")[:, 3] == 0.0)


def test_oracle_rejects_wrong_bin_count():
    with pytest.raises(oracle.OracleError):
        oracle.chroma_from_spectrogram(np.ones((100, 4)), 16000.0, 1024)


def test_chroma_params_mirror_reference_validation():
    p = sg.ChromaParams()
    assert (p.tuning, p.f_min, p.f_max, p.norm) == (440.0, 32.7, 4186.0, "l2")     # Default (src/chroma.rs:47-58)
    assert sg.ChromaParams.music_standard().n_octaves == 7                          # :114-122
    assert sg.ChromaParams(440.0, 55.0, 880.0).n_octaves == 4                       # ceil(log2(16)) (:97)
    assert p.with_norm("max").norm == "max"
    for bad, msg in (((0.0, 32.7, 4186.0), "tuning must be finite and > 0"), ((440.0, 0.0, 100.0), "f_min must be finite and > 0"),
                     ((440.0, 100.0, 100.0), "f_max must be > f_min"), ((float("inf"), 1.0, 2.0), "tuning must be finite and > 0")):
        with pytest.raises(sg.InvalidInputError, match=msg):
            sg.ChromaParams(*bad)
    assert sg.Chromagram.labels()[9] == "A" and len(sg.Chromagram.labels()) == 12   # :238-242


def test_host_filterbank_and_plan_queries_need_no_gpu():
    p = sg.ChromaParams(442.0, 55.0, 3000.0, "l1")
    fb = sg.build_chroma_filterbank(16000.0, 1024, p)
    assert np.array_equal(fb, oracle.chroma_filterbank(16000.0, 1024, 442.0, 55.0, 3000.0))    # same f64 arithmetic, bit for bit
    plan = sg.ChromaPlan(sg.StftParams(1024, 256, sg.WindowType.hanning(), True), 16000.0, p, "float32")
    assert plan.output_shape(16000) == (12, 63)
    m, nnz = plan.filterbank()
    assert np.array_equal(m, fb) and nnz == 12 * 513


def test_descriptor_rejects_non_magnitude_chroma():
    from spectrograms_b200 import plan as P
    params = sg.SpectrogramParams(sg.StftParams(512, 128, sg.WindowType.hanning(), True), 16000.0)
    with pytest.raises(sg.InvalidInputError, match="magnitude"):
        P._NativePlan(params, "float32", "chroma", sg.ChromaParams(), "power", None)


# ------------------------------------------------------------------------------------------------- GPU parity
CASES = [  # (n_fft, hop, sr, kernel family expected)
    (2048, 512, 22050.0, "r2c_fused_pow2"),      # the reference's own example shape (src/chroma.rs:476-477)
    (512, 128, 16000.0, "r2c_fused_pow2"),
    (400, 160, 16000.0, "r2c_fused_n400"),
    (800, 200, 16000.0, "r2c_fused_mixed"),
    (1009, 250, 16000.0, "r2c_fused_generic"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("n_fft,hop,sr,family", CASES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_fused_chromagram_matches_oracle(n_fft, hop, sr, family, dtype):
    dt = np.float32 if dtype == "float32" else np.float64
    tol = 1e-5 if dtype == "float32" else 1e-12                     # north_star tolerances for f32 / f64 spectra
    # 0.6 s keeps the chirp (100 + 9000 t^2 Hz instantaneous) inside the chroma band: frames whose band holds only f32
    # leakage noise would be "normalised noise", where the reference's own f32 and f64 instantiations disagree as well
    x = np.stack([signal(kind, int(0.6 * sr) + 7, sr).astype(dt) for kind in ("sine", "chirp", "noise")])
    for norm in NORMS:
        cp = sg.ChromaParams(440.0, 32.7, min(4186.0, sr / 2 - 1), norm)
        plan = sg.ChromaPlan(sg.StftParams(n_fft, hop, sg.WindowType.hanning(), True), sr, cp, dtype)
        if dtype == "float32" or n_fft != 400:
            assert plan.kernel_name() == family
        got = plan.compute_batch(x)
        assert got.shape == (3, 12, (x.shape[1] + 2 * (n_fft // 2) - n_fft) // hop + 1) and got.dtype == dt
        for i in range(3):
            want = oracle.chromagram(x[i], n_fft, hop, sr, f_max=cp.f_max, norm=norm)
            assert rel_l2(got[i], want) < tol, (norm, i)
        plan.force_generic(True)
        assert rel_l2(plan.compute_batch(x), got) < tol


@pytest.mark.gpu
def test_one_shot_chromagram_and_result_type():
    sr = 16000.0
    x = signal("chirp", 9600, sr)                      # 0.6 s: stays inside the chroma band (see above)
    c = sg.chromagram(x, sg.StftParams(2048, 512, sg.WindowType.hanning(), True), sr, sg.ChromaParams.music_standard())
    assert isinstance(c, sg.Chromagram) and c.n_bins == 12 and c.n_frames == 19 and c.data.dtype == np.float64
    want = oracle.chromagram(x, 2048, 512, sr)
    assert rel_l2(np.asarray(c), want) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_chromagram_from_spectrogram_host_and_device(dtype):
    import torch
    rng = np.random.default_rng(11)
    spec = np.abs(rng.standard_normal((3, 1025, 77))).astype(dtype)
    tol = 1e-5 if dtype == np.float32 else 1e-12
    for norm in NORMS:
        p = sg.ChromaParams(440.0, 32.7, 4186.0, norm)
        got = sg.chromagram_from_spectrogram(spec, 22050.0, 2048, p)
        assert got.shape == (3, 12, 77)
        for i in range(3):
            assert rel_l2(got.data[i], oracle.chroma_from_spectrogram(spec[i], 22050.0, 2048, norm=norm)) < tol
        dev = sg.chromagram_from_spectrogram(torch.from_numpy(spec).cuda(), 22050.0, 2048, p)
        assert dev.data.is_cuda and np.array_equal(dev.data.cpu().numpy(), got.data)      # same kernel either way
    one = sg.chromagram_from_spectrogram(spec[0], 22050.0, 2048, sg.ChromaParams())
    assert one.shape == (12, 77)
    with pytest.raises(sg.DimensionMismatchError) as e:                                       # src/chroma.rs:376-379
        sg.chromagram_from_spectrogram(spec[:, :1000], 22050.0, 2048, sg.ChromaParams())
    assert (e.value.expected, e.value.got) == (1025, 1000)
