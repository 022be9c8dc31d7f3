"""Pins for the half of the path the reference's tests hold no vectors for: mel / ERB / LogHz filterbanks and MFCC.

Fixtures (tests/golden/make_golden.py):
  * ``ref_numpy_impls.npz``  erb* / loghz* entries -- outputs of the reference's own ``python/examples/numpy_impls.py``
    (``erb_spectrogram``, ``logfreq_spectrogram``), which restate src/erb.rs:266-402 and src/spectrogram.rs:2438-2508;
  * ``third_party_pins.npz`` -- torchaudio's Slaney mel filterbank (the librosa construction src/spectrogram.rs:2262-2432
    says it matches) and SciPy's DCT-II (src/mfcc.rs:278-292).

Two layers: CPU tests pin the ORACLE to the fixtures; ``-m gpu`` tests compare the CUDA output with the fixtures
DIRECTLY (not through the oracle), on every kernel family.

Tolerances. ERB / LogHz fixtures are f64 restatements of the same formulae: f64 rel-L2 <= 1e-12, f32 <= 1e-5 (north_star).
torchaudio builds its filterbank in f32 (all_freqs, f_pts and the slopes are f32 tensors), so its weights carry f32
rounding: <= 2e-5 absolute on weights <= 1 (measured 1.2e-5); spectrogram fixtures derived from it are therefore
compared at rel-L2 <= 3e-5 (power), 1e-3 dB (dB) and 2e-2 absolute on MFCCs of magnitude ~3000 (40 coefficients, lifter
up to 12x, 128 dB terms each) -- the non-zero PATTERN (394 / 2018 entries), which is what indexing errors would move, must
agree exactly.
"""
import os

import numpy as np
import pytest

import oracle
from conftest import make_signal, rel_l2

G = os.path.join(os.path.dirname(__file__), "golden")
REF = np.load(os.path.join(G, "ref_numpy_impls.npz"))
TP = np.load(os.path.join(G, "third_party_pins.npz"))
SIGS = ["sine", "chirp", "noise"]
ERB_CASES = ["erb512", "erb400"]
LOGHZ_CASES = ["loghz1024", "loghz400"]
MEL_CASES = ["mel400", "mel2048", "mel512"]
TOL_F64, TOL_F32, TOL_DB = 1e-12, 1e-5, 1e-3
TOL_TA_W, TOL_TA_SPEC, TOL_TA_MFCC = 2e-5, 3e-5, 2e-2


def band_meta(name, src=REF):
    sr, n, n_fft, hop, nb, f_min, f_max = src[f"{name}/meta"][:7]
    return float(sr), int(n), int(n_fft), int(hop), int(nb), float(f_min), float(f_max)


def odesc(dtype, name, mapping, src=REF, **kw):
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name, src)
    return oracle.Desc(dtype=dtype, n_fft=n_fft, hop=hop, sample_rate=sr, mapping=mapping, n_bands=nb, f_min=f_min, f_max=f_max, **kw)


# ------------------------------------------------------------------------------------------------ oracle <- fixtures (CPU)
@pytest.mark.parametrize("name", ERB_CASES)
def test_oracle_erb_matches_reference_numpy(name):
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name)
    p = oracle.Plan(odesc("f64", name, "erb"))
    np.testing.assert_allclose(p.freq_axis(), REF[f"{name}/centres"], rtol=1e-13)        # erb_centers, src/erb.rs:276-284
    for sig in SIGS:
        x = make_signal(sig, n, sr)
        g = REF[f"{name}/{sig}/power"]
        out = p.compute(x)
        assert out.shape == g.shape == (nb, oracle.frame_count(n, n_fft, hop, True))
        assert rel_l2(out, g) < TOL_F64
        o32 = oracle.Plan(odesc("f32", name, "erb")).compute(x.astype(np.float32))
        assert rel_l2(o32, g) < TOL_F32


@pytest.mark.parametrize("name", LOGHZ_CASES)
def test_oracle_loghz_matches_reference_numpy(name):
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name)
    p = oracle.Plan(odesc("f64", name, "loghz"))
    M = REF[f"{name}/matrix"]
    fb = p.filterbank_dense()
    assert fb.shape == M.shape
    # the crate drops weights <= 1e-10 (SparseMatrix::set, src/spectrogram.rs:83); the NumPy helper keeps them
    np.testing.assert_allclose(fb, np.where(M > 1e-10, M, 0.0), rtol=1e-9, atol=1e-12)
    for sig in SIGS:
        x = make_signal(sig, n, sr)
        g = REF[f"{name}/{sig}/power"]
        out = p.compute(x)
        assert out.shape == g.shape
        assert rel_l2(out, g) < 1e-11          # interpolation weights are differences of O(100) numbers: 1e-13 relative each
        o32 = oracle.Plan(odesc("f32", name, "loghz")).compute(x.astype(np.float32))
        assert rel_l2(o32, g) < TOL_F32


def mel_norm(name):
    return "slaney" if TP[f"{name}/meta"][7] else "none"


@pytest.mark.parametrize("name", MEL_CASES)
def test_oracle_mel_filterbank_matches_torchaudio(name):
    p = oracle.Plan(odesc("f64", name, "mel", TP, mel_norm=mel_norm(name)))
    fb, g = p.filterbank_dense(), TP[f"{name}/fb"].astype(np.float64)
    assert fb.shape == g.shape
    scale = g.max()
    # Slaney-normalised rows multiply the f32 weight error by an f32-rounded 2 / (f_hi - f_lo): allow 4x
    assert np.abs(fb - g).max() <= TOL_TA_W * scale * (1 if mel_norm(name) == "none" else 4)
    # pattern: identical except where torchaudio's f32 arithmetic puts a weight within its own rounding of zero
    differ = (fb > 0) != (g > 0)
    assert np.all(np.maximum(fb, g)[differ] <= TOL_TA_W * scale)
    if name == "mel400":
        assert int((fb > 0).sum()) == int((g > 0).sum()) == 394 and not differ.any()
    if name == "mel2048":
        assert int((fb > 0).sum()) == int((g > 0).sum()) == 2018 and not differ.any()


@pytest.mark.parametrize("name", MEL_CASES)
def test_oracle_mel_spectrogram_matches_torchaudio_pin(name):
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name, TP)
    for sig in SIGS:
        x = make_signal(sig, n, sr)
        g = TP[f"{name}/{sig}/power"]
        out = oracle.Plan(odesc("f64", name, "mel", TP, mel_norm=mel_norm(name))).compute(x)
        assert out.shape == g.shape and rel_l2(out, g) < TOL_TA_SPEC
        if name == "mel400":
            db = oracle.Plan(odesc("f64", name, "mel", TP, amp="db", floor_db=-80.0)).compute(x)
            assert np.abs(db - TP[f"{name}/{sig}/db80"]).max() < TOL_DB
            mf = oracle.mfcc_from_log_mel(db, 40, True, 22)
            assert np.abs(mf - TP[f"{name}/{sig}/mfcc40"]).max() < TOL_TA_MFCC


@pytest.mark.parametrize("n", [13, 40, 64, 128])
def test_oracle_dct_matches_scipy(n):
    x, g = TP[f"dct/{n}/in"], TP[f"dct/{n}/out"]
    for faithful in (True, False):
        out = oracle.mfcc_from_log_mel(x, n, True, 0, faithful=faithful)
        assert rel_l2(out, g) < 1e-13
    o32 = oracle.mfcc_from_log_mel(x.astype(np.float32), n, True, 0)
    assert rel_l2(o32, g) < 2e-6
    # lifter (src/mfcc.rs:296-316) against its formula written out in NumPy
    w = 11.0 * np.sin(np.pi * np.arange(n) / 22.0) + 1.0
    assert rel_l2(oracle.mfcc_from_log_mel(x, n, True, 22), g * w[:, None]) < 1e-13
    assert rel_l2(oracle.mfcc_from_log_mel(x, n, False, 22), (g * w[:, None])[1:] if n > 1 else g * w[:, None]) < 1e-13


# ------------------------------------------------------------------------------------------------ product host tables <- fixtures (CPU)
def _sg():
    import spectrograms_b200 as sg
    return sg


@pytest.mark.parametrize("name", MEL_CASES)
def test_product_mel_table_matches_torchaudio(name):
    """The filterbank the CUDA kernels consume (built in tables.cpp, no GPU needed) against torchaudio directly."""
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name, TP)
    plan = sg.SpectrogramPlanner().mel_plan(_params(name, TP), sg.MelParams(nb, f_min, f_max, mel_norm(name)), None, "power", "float64")
    fb, nnz = plan.filterbank()
    g = TP[f"{name}/fb"].astype(np.float64)
    assert np.abs(fb - g).max() <= TOL_TA_W * g.max() * (1 if mel_norm(name) == "none" else 4)
    differ = (fb > 0) != (g > 0)
    assert np.all(np.maximum(fb, g)[differ] <= TOL_TA_W * g.max())
    if name in ("mel400", "mel2048"):
        assert nnz == int((g > 0).sum()) == {"mel400": 394, "mel2048": 2018}[name] and not differ.any()


@pytest.mark.parametrize("name", LOGHZ_CASES + ERB_CASES)
def test_product_loghz_erb_tables_match_reference_numpy(name):
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name)
    if name.startswith("loghz"):
        plan = sg.SpectrogramPlanner().log_hz_plan(_params(name, REF), sg.LogHzParams(nb, f_min, f_max), None, "power", "float64")
        M = REF[f"{name}/matrix"]
        np.testing.assert_allclose(plan.filterbank()[0], np.where(M > 1e-10, M, 0.0), rtol=1e-9, atol=1e-12)
    else:
        plan = sg.SpectrogramPlanner().erb_plan(_params(name, REF), sg.ErbParams(nb, f_min, f_max), None, "power", "float64")
        np.testing.assert_allclose(plan.freq_axis(), REF[f"{name}/centres"], rtol=1e-13)
        # |1 / (1 + j (f - fc) / b)^4|^2 written out (numpy_impls.py:143-146, src/erb.rs:300-322)
        fc = REF[f"{name}/centres"][:, None]
        f = np.arange(n_fft // 2 + 1)[None, :] * (sr / n_fft)
        b = 1.019 * 24.7 * (4.37 * fc / 1000.0 + 1.0)
        np.testing.assert_allclose(plan.filterbank()[0], np.abs(1.0 / (1.0 + 1j * (f - fc) / b) ** 4) ** 2, rtol=1e-12)


# ------------------------------------------------------------------------------------------------ CUDA <- fixtures (GPU)


def _run(plan, x, family):
    import torch
    plan.force_generic(family == "generic")
    r = plan.compute(torch.from_numpy(np.ascontiguousarray(x)).cuda())
    return r.data.cpu().numpy()


def _params(name, src):
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name, src)
    return sg.SpectrogramParams(sg.StftParams(n_fft, hop, "hanning", True), sr)


FAMILIES = ["auto", "generic"]


@pytest.mark.gpu
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", ERB_CASES)
def test_gpu_erb_matches_reference_numpy(name, dtype, family):
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name)
    plan = sg.SpectrogramPlanner().erb_plan(_params(name, REF), sg.ErbParams(nb, f_min, f_max), None, "power", dtype)
    for sig in SIGS:
        x = make_signal(sig, n, sr, np.float32 if dtype == "float32" else np.float64)
        out = _run(plan, x, family)
        g = REF[f"{name}/{sig}/power"]
        assert out.shape == g.shape
        assert rel_l2(out, g) <= (TOL_F64 if dtype == "float64" else TOL_F32)


@pytest.mark.gpu
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", LOGHZ_CASES)
def test_gpu_loghz_matches_reference_numpy(name, dtype, family):
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name)
    plan = sg.SpectrogramPlanner().log_hz_plan(_params(name, REF), sg.LogHzParams(nb, f_min, f_max), None, "power", dtype)
    for sig in SIGS:
        x = make_signal(sig, n, sr, np.float32 if dtype == "float32" else np.float64)
        out = _run(plan, x, family)
        g = REF[f"{name}/{sig}/power"]
        assert out.shape == g.shape
        assert rel_l2(out, g) <= (1e-11 if dtype == "float64" else TOL_F32)


@pytest.mark.gpu
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", MEL_CASES)
def test_gpu_mel_matches_torchaudio_pin(name, dtype, family):
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta(name, TP)
    mp = sg.MelParams(nb, f_min, f_max, mel_norm(name))
    plan = sg.SpectrogramPlanner().mel_plan(_params(name, TP), mp, None, "power", dtype)
    for sig in SIGS:
        x = make_signal(sig, n, sr, np.float32 if dtype == "float32" else np.float64)
        out = _run(plan, x, family)
        g = TP[f"{name}/{sig}/power"]
        assert out.shape == g.shape and rel_l2(out, g) <= TOL_TA_SPEC
    if name == "mel400":
        dbp = sg.SpectrogramPlanner().mel_plan(_params(name, TP), mp, sg.LogParams(-80.0), "db", dtype)
        x = make_signal("noise", n, sr, np.float32 if dtype == "float32" else np.float64)
        assert np.abs(_run(dbp, x, family) - TP[f"{name}/noise/db80"]).max() <= TOL_DB


@pytest.mark.gpu
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_gpu_fused_mfcc_matches_scipy_torchaudio_pin(dtype, family):
    """configs[3] end to end against third-party code only: torchaudio mel @ reference-NumPy power -> dB -> SciPy DCT."""
    sg = _sg()
    sr, n, n_fft, hop, nb, f_min, f_max = band_meta("mel400", TP)
    plan = sg.MfccPlan(sg.StftParams(n_fft, hop), sr, nb, sg.MfccParams(n_mfcc=40), dtype)
    x = make_signal("noise", n, sr, np.float32 if dtype == "float32" else np.float64)
    out = _run(plan, x, family)
    g = TP["mel400/noise/mfcc40"]
    assert out.shape == g.shape
    # f32: 128 dB terms each within 1e-3 dB (north_star) x lifter <= 12 -> sqrt(128) * 1e-3 * 12 = 0.14 worst case
    assert np.abs(out - g).max() <= (TOL_TA_MFCC if dtype == "float64" else 0.14)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [13, 40, 64, 128])
def test_gpu_dct_matches_scipy(n, dtype):
    sg = _sg()
    dt = np.float32 if dtype == "float32" else np.float64
    x, g = TP[f"dct/{n}/in"].astype(dt), TP[f"dct/{n}/out"]
    out = sg.mfcc_from_log_mel(x, sg.MfccParams(n_mfcc=n, lifter=0))
    out = out.data if hasattr(out, "data") else out
    assert rel_l2(np.asarray(out), g) <= (TOL_F64 if dtype == "float64" else TOL_F32)
