"""CPU-side tests of the product's host logic through the C ABI: validation and error taxonomy, frame counts, axes,
windows and filterbanks (bit-exact against the oracle's tables), symbol exports, and the no-oracle / no-fallback
rules. No compute call is made here (there is no GPU in the build container)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from spectrograms_b200 import _native
from conftest import ROOT, HAS_GPU


def P(n_fft=512, hop=256, window="hanning", centre=True, sr=16000.0):
    return sg.SpectrogramParams(sg.StftParams(n_fft, hop, window, centre), sr)


# ---------------------------------------------------------------- boundary
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sgx_b200.h")).read()
    declared = set(re.findall(r"\b(sgx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    lib = _native.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("sgx_")}
    assert declared <= exported
    assert b"sm_100a" in lib.sgx_version()


def test_header_has_no_torch_types_and_cites_reference():
    header = open(os.path.join(ROOT, "include", "sgx_b200.h")).read()
    assert "torch" not in header.lower() and "at::" not in header
    assert header.count("src/") >= 10 and 'extern "C"' in header


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "spectrograms_b200")
    for dp, _, files in os.walk(pkg):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(import|from)\s+oracle", r"liboracle", r"oracle[/\\]", r"#include\s*[\"<].*oracle", r"orc_"):
                    assert not re.search(pat, src, re.M), f"{f} reaches into the oracle ({pat})"
    out = subprocess.run(["ldd", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in out and "libtorch" not in out and "libcufft" not in out


@pytest.mark.skipif(HAS_GPU, reason="needs a machine without a GPU")
def test_no_cpu_fallback_without_gpu():
    plan = sg.SpectrogramPlanner().linear_plan(P(), None, "power", "float64")
    with pytest.raises(sg.FFTBackendError, match="no CPU fallback"):
        plan.compute(np.zeros(16000))
    with pytest.raises(sg.FFTBackendError):
        sg.mfcc_from_log_mel(np.zeros((40, 10)), sg.MfccParams(13))


def test_missing_library_fails_loudly(tmp_path):
    code = ("import spectrograms_b200._native as n; n.LIB_PATH=r'%s'; n._lib=None\n"
            "try:\n n.lib()\nexcept Exception as e:\n print(type(e).__name__)\n" % str(tmp_path / "nope.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT).stdout
    assert "FFTBackendError" in out


# ---------------------------------------------------------------- validation (tests/params_tests.rs, builder_tests.rs)
def test_param_validation_messages():
    with pytest.raises(sg.InvalidInputError, match="hop_size must be <= n_fft"):
        sg.StftParams(256, 512)
    for bad in (0.0, -1.0, float("inf"), float("nan")):
        with pytest.raises(sg.InvalidInputError, match="sample_rate_hz must be finite and > 0"):
            sg.SpectrogramParams(sg.StftParams(512, 256), bad)
    with pytest.raises(sg.InvalidInputError, match="f_min must be >= 0"):
        sg.MelParams(40, -1.0, 8000.0)
    with pytest.raises(sg.InvalidInputError, match="f_max must be > f_min"):
        sg.MelParams(40, 100.0, 100.0)
    with pytest.raises(sg.InvalidInputError, match="n_filters must be >= 2"):
        sg.ErbParams(1, 0.0, 8000.0)
    with pytest.raises(sg.InvalidInputError, match="f_min must be finite and > 0"):
        sg.LogHzParams(84, 0.0, 8000.0)
    with pytest.raises(sg.InvalidInputError, match="floor_db must be finite"):
        sg.LogParams(float("-inf"))
    with pytest.raises(sg.InvalidInputError, match=r"Custom window size \(5\) must match n_fft \(8\)"):
        sg.StftParams(8, 4, sg.WindowType.custom([0, .5, 1, .5, 0]))
    with pytest.raises(sg.InvalidInputError):
        sg.StftParams(0, 1)


def test_planner_cross_validation():
    pl = sg.SpectrogramPlanner()
    with pytest.raises(sg.InvalidInputError, match="mel f_max must be <= Nyquist"):      # tests/spectrogram_tests.rs:147-158
        pl.mel_plan(P(), sg.MelParams(80, 0.0, 9000.0))
    with pytest.raises(sg.InvalidInputError, match="exceeds Nyquist"):
        pl.erb_plan(P(), sg.ErbParams(40, 0.0, 9000.0))
    with pytest.raises(sg.InvalidInputError, match="exceeds Nyquist"):
        pl.log_hz_plan(P(), sg.LogHzParams(40, 20.0, 9000.0))
    with pytest.raises(sg.InvalidInputError, match="unreasonably large"):
        pl.mel_plan(P(), sg.MelParams(10001, 0.0, 8000.0))
    with pytest.raises(sg.InvalidInputError, match="n_mfcc must be <= n_mels"):
        sg.MfccPlan(sg.StftParams(512, 160), 16000.0, 20, sg.MfccParams(21))
    with pytest.raises(sg.InvalidInputError, match="Unsupported dtype"):
        pl.linear_plan(P(), dtype="float16")


def test_c_abi_validates_independently_of_python():
    """The C ABI must reject bad descriptors itself (a Rust caller has no Python layer in front)."""
    import ctypes as C
    L = _native.lib()
    d = _native.PlanDesc()
    d.dtype, d.n_fft, d.hop_size, d.sample_rate_hz = 1, 256, 512, 16000.0
    h = C.c_void_p()
    assert L.sgx_plan_create(C.byref(d), C.byref(h)) == _native.SGX_INVALID_INPUT
    assert L.sgx_last_error_message() == b"Invalid input: hop_size must be <= n_fft"
    d.hop_size, d.sample_rate_hz = 128, 0.0
    assert L.sgx_plan_create(C.byref(d), C.byref(h)) == _native.SGX_INVALID_INPUT
    assert b"sample_rate_hz must be finite and > 0" in L.sgx_last_error_message()
    d.sample_rate_hz, d.mapping, d.n_bands, d.f_min, d.f_max = 16000.0, 1, 40, 0.0, 9000.0
    assert L.sgx_plan_create(C.byref(d), C.byref(h)) == _native.SGX_INVALID_INPUT
    d.f_max, d.output, d.n_mfcc = 8000.0, 2, 41
    assert L.sgx_plan_create(C.byref(d), C.byref(h)) == _native.SGX_INVALID_INPUT
    assert b"n_mfcc must be <= n_mels" in L.sgx_last_error_message()
    d.n_mfcc = 13
    assert L.sgx_plan_create(C.byref(d), C.byref(h)) == _native.SGX_OK
    r, f = C.c_size_t(), C.c_size_t()
    assert L.sgx_plan_output_shape(h, 0, C.byref(r), C.byref(f)) == _native.SGX_INVALID_INPUT
    assert L.sgx_plan_output_shape(h, 16000, C.byref(r), C.byref(f)) == _native.SGX_OK and (r.value, f.value) == (12, 124)   # include_c0=0 drops c0; centre=0
    assert L.sgx_plan_destroy(h) == _native.SGX_OK


def test_defaults_and_presets():
    # tests/builder_tests.rs: speech 512/160, music 2048/512 ; src/mfcc.rs:30-39
    s, m = sg.SpectrogramParams.speech_default(16000.0), sg.SpectrogramParams.music_default(44100.0)
    assert (s.stft.n_fft, s.stft.hop_size, s.stft.window.kind, s.stft.centre) == (512, 160, "hanning", True)
    assert (m.stft.n_fft, m.stft.hop_size) == (2048, 512)
    assert s.nyquist_hz() == 8000.0 and abs(s.frame_period_seconds() - 0.01) < 1e-15
    mp = sg.MfccParams()
    assert (mp.n_mfcc, mp.include_c0, mp.lifter) == (13, True, 22)
    assert sg.MfccParams.speech_standard().n_mfcc == 13 and mp.with_c0(False).include_c0 is False


def test_window_type_parsing_and_custom_normalisation():
    # tests/window_tests.rs:4-105,127-424
    assert sg.WindowType.from_str("Hann").kind == "hanning" and sg.WindowType.from_str("kaiser=8.6").param == 8.6
    assert sg.WindowType.from_str("gaussian=12").kind == "gaussian"
    with pytest.raises(sg.InvalidInputError):
        sg.WindowType.from_str("bogus")
    w = sg.WindowType.custom([1.0, 2.0, 1.0], "sum")
    assert abs(sum(w.coefficients) - 1.0) < 1e-15
    assert max(sg.WindowType.custom([1.0, 4.0, 1.0], "peak").coefficients) == 1.0
    e = sg.WindowType.custom([3.0, 4.0], "energy").coefficients
    assert abs(e[0] ** 2 + e[1] ** 2 - 1.0) < 1e-15
    for bad in ([], [1.0, float("nan")]):
        with pytest.raises(sg.InvalidInputError):
            sg.WindowType.custom(bad)
    with pytest.raises(sg.InvalidInputError, match="sum is zero"):
        sg.WindowType.custom([1.0, -1.0], "sum")
    with pytest.raises(sg.InvalidInputError, match="Unknown normalization"):
        sg.WindowType.custom([1.0], "l7")


# ---------------------------------------------------------------- shapes / axes / tables, bit-exact vs the oracle
@pytest.mark.parametrize("n,n_fft,hop,centre", [(480000, 400, 160, True), (661500, 2048, 512, True), (160000, 400, 160, True),
                                                 (2880000, 4096, 1024, True), (16000, 512, 256, True), (16000, 512, 256, False),
                                                 (5, 512, 256, True), (5, 512, 256, False), (1, 1, 1, True), (1000, 7, 7, False),
                                                 (999, 250, 100, True)])
def test_output_shape_is_bit_exact(n, n_fft, hop, centre):
    plan = sg.SpectrogramPlanner().linear_plan(P(n_fft, hop, "rectangular", centre), None, "power", "float32")
    assert plan.output_shape(n) == (n_fft // 2 + 1, oracle.frame_count(n, n_fft, hop, centre))
    st = sg.StftPlan(P(n_fft, hop, "rectangular", centre), "float64")
    assert st.output_shape(n) == (n_fft // 2 + 1, oracle.frame_count(n, n_fft, hop, centre))


def test_config_frame_counts():
    # SURVEY section 8 config sizes (F2): 3001 / 1292 / 1001 / 2813 ; doc-test (257, 63)
    mk = lambda nf, h, sr: sg.SpectrogramPlanner().mel_plan(P(nf, h, sr=sr), sg.MelParams(128, 0.0, sr / 2), sg.LogParams(-80.0), "db", "float32")
    assert mk(400, 160, 16000.0).output_shape(480000) == (128, 3001)
    assert mk(2048, 512, 22050.0).output_shape(661500) == (128, 1292)
    assert sg.MfccPlan(sg.StftParams(400, 160), 16000.0, 128, sg.MfccParams(40), "float32").output_shape(160000) == (40, 1001)
    assert sg.MfccPlan(sg.StftParams(400, 160), 16000.0, 128, sg.MfccParams(40, include_c0=False), "float32").output_shape(160000) == (39, 1001)
    assert sg.SpectrogramPlanner().linear_plan(P(4096, 1024, sr=48000.0), None, "magnitude", "float64").output_shape(2880000) == (2049, 2813)
    assert sg.SpectrogramPlanner().linear_plan(P(), None, "power", "float64").output_shape(16000) == (257, 63)


WIN_CASES = [("rectangular", 0.0), ("hanning", 0.0), ("hamming", 0.0), ("blackman", 0.0), ("kaiser", 8.6), ("gaussian", 50.0)]


@pytest.mark.parametrize("kind,prm", WIN_CASES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_window_bit_exact(kind, prm, dtype):
    wt = sg.WindowType(kind, prm)
    for n in (400, 512, 2048, 255):
        plan = sg.StftPlan(P(n, n // 4, wt), dtype)
        ref = oracle.Plan(oracle.Desc(dtype="f32" if dtype == "float32" else "f64", n_fft=n, hop=n // 4, window=kind, window_param=prm)).window()
        assert np.array_equal(plan.window(), ref)


@pytest.mark.parametrize("norm", ["none", "slaney", "l1", "l2"])
def test_mel_filterbank_bit_exact(norm):
    for n_fft, sr, n_mels, lo, hi in [(400, 16000.0, 128, 0.0, 8000.0), (2048, 22050.0, 128, 0.0, 11025.0), (512, 16000.0, 40, 300.0, 4000.0)]:
        plan = sg.SpectrogramPlanner().mel_plan(P(n_fft, n_fft // 4, sr=sr), sg.MelParams(n_mels, lo, hi, norm))
        o = oracle.Plan(oracle.Desc(n_fft=n_fft, hop=n_fft // 4, sample_rate=sr, mapping="mel", n_bands=n_mels, f_min=lo, f_max=hi, mel_norm=norm))
        fb, nnz = plan.filterbank()
        assert np.array_equal(fb, o.filterbank_dense()) and nnz == o.filterbank_nnz()
        assert np.array_equal(plan.freq_axis(), o.freq_axis())
    assert sg.SpectrogramPlanner().mel_plan(P(400, 160), sg.MelParams(128, 0.0, 8000.0)).filterbank()[1] == 394
    assert sg.SpectrogramPlanner().mel_plan(P(2048, 512, sr=22050.0), sg.MelParams(128, 0.0, 11025.0)).filterbank()[1] == 2018


def test_erb_and_loghz_tables_bit_exact():
    for sp in ("linear", "apple_tr35"):
        plan = sg.SpectrogramPlanner().erb_plan(P(512, 160), sg.ErbParams(40, 50.0, 8000.0, sp))
        o = oracle.Plan(oracle.Desc(n_fft=512, hop=160, mapping="erb", n_bands=40, f_min=50.0, f_max=8000.0, erb_spacing=sp))
        assert np.array_equal(plan.filterbank()[0], o.filterbank_dense()) and np.array_equal(plan.freq_axis(), o.freq_axis())
    plan = sg.SpectrogramPlanner().log_hz_plan(P(1024, 256), sg.LogHzParams(84, 32.7, 7900.0))
    o = oracle.Plan(oracle.Desc(n_fft=1024, hop=256, mapping="loghz", n_bands=84, f_min=32.7, f_max=7900.0))
    assert np.array_equal(plan.filterbank()[0], o.filterbank_dense()) and np.array_equal(plan.freq_axis(), o.freq_axis())
    assert plan.filterbank()[1] == o.filterbank_nnz()


def test_axes():
    # tests/spectrogram_tests.rs:183-236
    n = sg.SpectrogramPlanner().linear_plan(P(), None, "power")._n
    f, t = n.axes(63)
    assert f[0] == 0.0 and abs(f[-1] - 8000.0) < 1e-3 and np.all(np.diff(f) > 0)
    assert t[0] == 0.0 and np.allclose(np.diff(t), 256 / 16000.0, atol=1e-6) and len(t) == 63
    assert np.array_equal(t, oracle.Plan(oracle.Desc()).times(63))


# ---------------------------------------------------------------- sharding (multi-GPU path, host side)
def test_shard_range_partitions_clips():
    for n in (1, 7, 8, 1024, 4096, 8192, 64):
        for w in (1, 2, 4, 8):
            rs = [sg.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:])) and all(lo <= hi for lo, hi in rs)
            assert max(hi - lo for lo, hi in rs) == -(-n // w)
    with pytest.raises(ValueError):
        sg.shard_range(8, 8, 8)


def test_compute_batch_rejects_out_buffers_it_could_overrun():
    """the C ABI sees only (rows, frames) of ``out``; the clip dimension, rank and dtype are checked before."""
    plan = sg.SpectrogramPlanner().mel_plan(sg.SpectrogramParams(sg.StftParams(400, 160), 16000.0), sg.MelParams(16, 0.0, 8000.0),
                                            None, "power", "float32")
    clips = np.zeros((3, 1600), dtype=np.float32)
    rows, nf = plan.output_shape(1600)
    with pytest.raises(sg.DimensionMismatchError):
        plan.compute_batch(clips, out=np.empty((2, rows, nf), dtype=np.float32))       # fewer clips than the input
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=np.empty((rows, nf), dtype=np.float32))          # 2-D out with n_clips > 1
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=np.empty((3, rows, nf), dtype=np.float64))       # wrong dtype
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=np.empty((3 * rows * nf,), dtype=np.float32))    # 1-D
    ro = np.empty((3, rows, nf), dtype=np.float32)
    ro.setflags(write=False)
    with pytest.raises(sg.InvalidInputError):
        plan.compute_batch(clips, out=ro)
