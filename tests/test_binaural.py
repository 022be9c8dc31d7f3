"""Binaural cue widening (reference src/binaural.rs): oracle known answers and host-side
validation on CPU, GPU parity of compute_{itd,ipd,ild,ilr}_spectrogram and of the STFT-level entry point.

The reference's tests for this module assert shapes, units and parameter validation only; the oracle restates
magphase / np_mod / the four cue loops expression by expression ("parity unpinned" beyond the known answers below).
Phase cues are compared on the circle (a value within rounding of +-pi may legitimately land on either side)."""
import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from conftest import make_signal

SR = 16000.0


def stereo(n, dtype, seed=0, delay=3, gain=0.6):
    rng = np.random.default_rng(seed)
    left = rng.standard_normal(n)
    right = gain * np.roll(left, delay) + 0.3 * rng.standard_normal(n)       # correlated, every bin has energy
    return left.astype(dtype), right.astype(dtype)


def circ(d):
    return np.abs((d + np.pi) % (2 * np.pi) - np.pi)


# ------------------------------------------------------------------------------------------------- CPU
def test_param_validation_messages_follow_the_reference():
    p = sg.SpectrogramParams(sg.StftParams(512, 128, sg.WindowType.hanning(), True), SR)
    for cls in (sg.ITDSpectrogramParams, sg.IPDSpectrogramParams, sg.ILDSpectrogramParams, sg.ILRSpectrogramParams):
        with pytest.raises(sg.InvalidInputError, match="Start and end frequencies must be positive."):      # src/binaural.rs:417-421
            cls(p, 0.0, 1000.0)
        with pytest.raises(sg.InvalidInputError, match="Start frequency must be less than end frequency."):  # :422-426
            cls(p, 2000.0, 1000.0)
        with pytest.raises(sg.InvalidInputError, match="End frequency must be less than Nyquist frequency."):  # :433-437
            cls(p, 100.0, 8000.5)
    q = sg.ITDSpectrogramParams(p, 200.0, 4000.0)
    assert q.magphase_power == 1 and q.band() == oracle.binaural_band(200.0, 4000.0, SR, 512) == (6, 128, 31.25)
    assert sg.ITDSpectrogramParams(p, 200.0, 4000.0, 2).magphase_power == 2
    assert sg.IPDSpectrogramParams(p, 200.0, 4000.0, False).wrapped is False


def test_oracle_known_answers():
    n_fft, hop = 512, 128
    x = make_signal("noise", 8000, SR)
    same = {c: oracle.binaural(c, x, x, n_fft, hop, SR, 200.0, 4000.0) for c in ("itd", "ipd", "ild", "ilr")}
    for c, v in same.items():                                   # identical channels: every cue is exactly zero
        assert v.shape == (122, 63) and np.all(v == 0.0), c
    # a pure tone delayed by d samples in the right channel: left leads, ITD at the tone's bin = +d / sr
    f0, d = 500.0, 4
    t = np.arange(16000) / SR
    left, right = np.sin(2 * np.pi * f0 * t), np.sin(2 * np.pi * f0 * (t - d / SR))
    itd = oracle.binaural("itd", left, right, n_fft, hop, SR, 200.0, 4000.0)
    b0, _, bw = oracle.binaural_band(200.0, 4000.0, SR, n_fft)
    k = int(round(f0 / bw)) - b0
    assert abs(np.median(itd[k, 5:-5]) - d / SR) < 2e-6
    ipd = oracle.binaural("ipd", left, right, n_fft, hop, SR, 200.0, 4000.0)
    assert abs(np.median(ipd[k, 5:-5]) - 2 * np.pi * f0 * d / SR) < 2e-2
    # level cues: right = left / 2 -> ILD = -20 log10(0.5) = +6.02 dB, ILR = 1 - 0.5
    ild = oracle.binaural("ild", x, 0.5 * x, n_fft, hop, SR, 200.0, 4000.0)
    ilr = oracle.binaural("ilr", x, 0.5 * x, n_fft, hop, SR, 200.0, 4000.0)
    np.testing.assert_allclose(ild, 20 * np.log10(2.0), atol=1e-9)
    np.testing.assert_allclose(ilr, 0.5, atol=1e-12)
    np.testing.assert_allclose(oracle.binaural("ilr", x, 2.0 * x, n_fft, hop, SR, 200.0, 4000.0), -0.5, atol=1e-12)   # -(1 - 1/r)
    # a silent channel: level cues are NaN (from_elem(.., nan), src/binaural.rs:1212, :1555)
    assert np.all(np.isnan(oracle.binaural("ild", x, np.zeros_like(x), n_fft, hop, SR, 200.0, 4000.0)))
    assert np.all(oracle.binaural("itd", np.zeros_like(x), np.zeros_like(x), n_fft, hop, SR, 200.0, 4000.0) == 0.0)


def test_magphase_power_only_weights_the_itd_mask():
    l, r = stereo(6000, np.float64)
    a = oracle.binaural("itd", l, r, 512, 128, SR, 200.0, 4000.0, magphase_power=1)
    b = oracle.binaural("itd", l, r, 512, 128, SR, 200.0, 4000.0, magphase_power=5)
    assert np.array_equal(a, b)            # the power only enters the `intensity > 0` test (:536-538)


# ------------------------------------------------------------------------------------------------- GPU parity
CASES = [(512, 128, "r2c_fused_pow2"), (2048, 512, "r2c_fused_pow2"), (400, 160, "r2c_fused_mixed"), (800, 200, "r2c_fused_mixed"), (1009, 250, "r2c_fused_generic")]


@pytest.mark.gpu
@pytest.mark.parametrize("n_fft,hop,family", CASES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cues_match_oracle(n_fft, hop, family, dtype):
    dt = np.float32 if dtype == "float32" else np.float64
    ph_max, ph_rms, lv_max, lv_rms = (2e-3, 3e-5, 2e-2, 2e-4) if dtype == "float32" else (1e-9, 1e-11, 1e-8, 1e-10)
    pairs = [stereo(12000 + 13, dt, seed=s, delay=2 + s) for s in range(3)]
    left, right = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    sp = sg.SpectrogramParams(sg.StftParams(n_fft, hop, sg.WindowType.hanning(), True), SR)
    plan = sg.StftPlan(sp, dtype)
    assert plan.kernel_name() == family
    f_lo, f_hi = 150.0, 6000.0
    b0, b1, bw = oracle.binaural_band(f_lo, f_hi, SR, n_fft)
    bins = np.arange(b0, b1, dtype=np.float64)[:, None]
    for cue, params in (("itd", sg.ITDSpectrogramParams(sp, f_lo, f_hi, 2)), ("ipd", sg.IPDSpectrogramParams(sp, f_lo, f_hi, True)),
                        ("ipd_raw", sg.IPDSpectrogramParams(sp, f_lo, f_hi, False)), ("ild", sg.ILDSpectrogramParams(sp, f_lo, f_hi)),
                        ("ilr", sg.ILRSpectrogramParams(sp, f_lo, f_hi))):
        fn = {"itd": sg.compute_itd_spectrogram, "ipd": sg.compute_ipd_spectrogram, "ipd_raw": sg.compute_ipd_spectrogram,
              "ild": sg.compute_ild_spectrogram, "ilr": sg.compute_ilr_spectrogram}[cue]
        got = fn([left, right], params, plan)
        assert got.shape == (3, b1 - b0, (left.shape[1] + 2 * (n_fft // 2) - n_fft) // hop + 1) and got.data.dtype == dt
        assert np.allclose(got.frequencies, np.arange(b0, b1) * bw) and got.unit == sg.BinauralSpectrogram.UNITS[cue[:3]]
        for i in range(3):
            want = oracle.binaural(cue[:3], left[i], right[i], n_fft, hop, SR, f_lo, f_hi, magphase_power=2, wrapped=cue != "ipd_raw")
            g, w = got.data[i].astype(np.float64), want.astype(np.float64)
            if cue == "itd":
                err = circ((g - w) * (2 * np.pi * bw * bins))
                assert err.max() < ph_max and np.sqrt(np.mean(err ** 2)) < ph_rms, (cue, i, err.max())
            elif cue.startswith("ipd"):
                err = circ(g - w)
                assert err.max() < ph_max and np.sqrt(np.mean(err ** 2)) < ph_rms, (cue, i, err.max())
            else:
                assert np.array_equal(np.isnan(g), np.isnan(w))
                err = np.abs(g - w)[~np.isnan(w)]
                scale = 1.0 if cue == "ild" else 0.05        # ILR is a ratio in (-1, 1): tighter absolute budget
                assert err.max() < lv_max * scale and np.sqrt(np.mean(err ** 2)) < lv_rms * scale, (cue, i, err.max())
    one = sg.compute_ild_spectrogram([left[0], right[0]], sg.ILDSpectrogramParams(sp, f_lo, f_hi), plan)   # reference-style call
    assert one.shape == (b1 - b0, got.n_frames)


@pytest.mark.gpu
def test_silent_channel_and_identical_channels_on_device():
    import torch
    sp = sg.SpectrogramParams(sg.StftParams(512, 128, sg.WindowType.hanning(), True), SR)
    plan = sg.StftPlan(sp, "float32")
    x = torch.randn(2, 9000, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    z = torch.zeros_like(x)
    ild = sg.compute_ild_spectrogram([x, z], sg.ILDSpectrogramParams(sp, 200.0, 4000.0), plan)
    assert ild.data.is_cuda and bool(torch.isnan(ild.data).all())
    for fn, cls in ((sg.compute_itd_spectrogram, sg.ITDSpectrogramParams), (sg.compute_ipd_spectrogram, sg.IPDSpectrogramParams),
                    (sg.compute_ild_spectrogram, sg.ILDSpectrogramParams), (sg.compute_ilr_spectrogram, sg.ILRSpectrogramParams)):
        same = fn([x, x], cls(sp, 200.0, 4000.0), plan)
        # level cues are exactly zero; in f32 the reference's own wrap ((0 + pi) mod 2 pi) - pi leaves -2.4e-7 rad
        # because 3 pi is not representable -- a faithful restatement reproduces that
        tol = {"itd": 1e-8, "ipd": 1e-6}.get(same.cue, 0.0)
        assert float(same.data.abs().max()) <= tol
    with pytest.raises(sg.InvalidInputError, match="complex STFT plan|StftPlan"):
        sg.compute_itd_spectrogram([x, x], sg.ITDSpectrogramParams(sp, 200.0, 4000.0), sg.SpectrogramPlanner().linear_plan(sp, None, "power", "float32"))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stft_level_entry_point_host_and_device(dtype):
    import torch
    l, r = stereo(7000, dtype, seed=9)
    d = oracle.Desc(dtype="f32" if dtype == np.float32 else "f64", n_fft=512, hop=128, sample_rate=SR)
    p = oracle.Plan(d)
    L, R = p.stft(l), p.stft(r)
    b0, b1, bw = oracle.binaural_band(300.0, 5000.0, SR, 512)
    for cue in ("ild", "ilr"):       # same inputs bit for bit -> level cues agree to the last ulps of division / log10
        want = oracle.binaural_from_stft(cue, L, R, b0, b1, bw)
        got = sg.binaural_from_stft(cue, L, R, b0, b1, bw)
        np.testing.assert_allclose(got, want, rtol=2e-6 if dtype == np.float32 else 1e-14, atol=1e-6 if dtype == np.float32 else 1e-14)
        dev = sg.binaural_from_stft(cue, torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda(), b0, b1, bw)
        assert np.array_equal(dev.cpu().numpy(), got)
    for cue in ("itd", "ipd"):
        want = oracle.binaural_from_stft(cue, L, R, b0, b1, bw, 1, True)
        got = sg.binaural_from_stft(cue, L, R, b0, b1, bw, 1, True)
        scale = 2 * np.pi * bw * np.arange(b0, b1)[:, None] if cue == "itd" else 1.0
        assert circ((got.astype(np.float64) - want) * scale).max() < (1e-5 if dtype == np.float32 else 1e-13)
    with pytest.raises(sg.InvalidInputError, match="at least one bin"):
        sg.binaural_from_stft("ipd", L, R, 10, 10, bw)
    with pytest.raises(sg.DimensionMismatchError):
        sg.binaural_from_stft("ipd", L, R, 10, 400, bw)
