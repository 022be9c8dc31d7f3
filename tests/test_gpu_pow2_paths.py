"""Code paths of the r2c_fused_pow2 family (power-of-two n_fft) against the oracle: odd hop / unaligned input (scalar
loads), every mapping incl. the fused MFCC through the general epilogue, row-per-thread sparse epilogue (small tiles) vs
the general one (large tiles), f32 and f64, centre on/off, clips shorter than a frame."""
import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from conftest import make_signal, rel_l2

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def tol(dtype):
    return 1e-12 if dtype == "float64" else 1e-5


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n_fft,hop", [(256, 64), (512, 160), (512, 171), (1024, 341), (2048, 512), (4096, 1023), (8192, 4096), (256, 256), (512, 1)])
def test_mel_db_and_linear(dtype, n_fft, hop):
    dt = np.float32 if dtype == "float32" else np.float64
    n = 6000 if hop > 1 else 1200
    x = make_signal("noise", n, 22050.0, dt, seed=n_fft + hop)
    for centre in (True, False):
        params = sg.SpectrogramParams(sg.StftParams(n_fft, hop, sg.WindowType.hanning(), centre), 22050.0)
        plan = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(96, 30.0, 11000.0), sg.LogParams(-90.0), "db", dtype)
        assert plan.kernel_name() == "r2c_fused_pow2"
        got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
        ref = oracle.Plan(oracle.Desc(dtype="f64", n_fft=n_fft, hop=hop, centre=centre, sample_rate=22050.0, mapping="mel", n_bands=96,
                                      f_min=30.0, f_max=11000.0, amp="db", floor_db=-90.0)).compute(x.astype(np.float64))
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-3
        lin = sg.SpectrogramPlanner().linear_plan(params, None, "magnitude", dtype)
        got = lin.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
        ref = oracle.Plan(oracle.Desc(dtype="f64", n_fft=n_fft, hop=hop, centre=centre, sample_rate=22050.0, amp="magnitude")).compute(x.astype(np.float64))
        assert got.shape == ref.shape and rel_l2(got, ref) <= tol(dtype)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_unaligned_views_and_short_clips(dtype):
    torch = _torch()
    dt = np.float32 if dtype == "float32" else np.float64
    base = np.random.default_rng(8).standard_normal((3, 9001)).astype(dt)
    dev = torch.from_numpy(base).cuda()
    params = sg.SpectrogramParams(sg.StftParams(1024, 256, sg.WindowType.hamming(), True), 16000.0)
    plan = sg.SpectrogramPlanner().linear_plan(params, None, "power", dtype)
    o = oracle.Plan(oracle.Desc(dtype="f64", n_fft=1024, hop=256, window="hamming"))
    for view in (dev[:, 1:], dev[:, :9000], dev[:, 5:8000]):
        got = plan.compute_batch(view).cpu().numpy()
        host = view.cpu().numpy().astype(np.float64)
        for i in range(3):
            assert rel_l2(got[i], o.compute(host[i])) <= tol(dtype)
    for n in (1, 5, 511, 1023, 1024, 1025):
        x = make_signal("noise", n, 16000.0, dt, seed=n)
        got = plan.compute(x).data
        ref = o.compute(x.astype(np.float64))
        assert got.shape == ref.shape and rel_l2(got, ref) <= tol(dtype)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_fused_mfcc_erb_loghz_on_pow2(dtype):
    dt = np.float32 if dtype == "float32" else np.float64
    x = make_signal("noise", 20000, 16000.0, dt, seed=77)
    t = _torch().from_numpy(x).cuda()
    mp = sg.MfccParams(13)
    plan = sg.MfccPlan(sg.StftParams(512, 160), 16000.0, 40, mp, dtype)           # the reference's speech pipeline (benches :373-405)
    assert plan.kernel_name() == "r2c_fused_pow2"
    got = plan.compute(t).data.cpu().numpy()
    lm = oracle.Plan(oracle.Desc(dtype="f64", n_fft=512, hop=160, mapping="mel", n_bands=40, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0)).compute(x.astype(np.float64))
    ref = oracle.mfcc_from_log_mel(lm, 13)
    assert got.shape == ref.shape == (13, 126) and rel_l2(got, ref) <= (1e-11 if dtype == "float64" else 1e-5)
    params = sg.SpectrogramParams(sg.StftParams(512, 160), 16000.0)
    e = sg.SpectrogramPlanner().erb_plan(params, sg.ErbParams(64, 0.0, 8000.0), None, "power", dtype).compute(t).data.cpu().numpy()
    re = oracle.Plan(oracle.Desc(dtype="f64", n_fft=512, hop=160, mapping="erb", n_bands=64, f_min=0.0, f_max=8000.0)).compute(x.astype(np.float64))
    assert rel_l2(e, re) <= tol(dtype)
    l = sg.SpectrogramPlanner().log_hz_plan(params, sg.LogHzParams(72, 50.0, 7800.0), None, "magnitude", dtype).compute(t).data.cpu().numpy()
    rl = oracle.Plan(oracle.Desc(dtype="f64", n_fft=512, hop=160, mapping="loghz", n_bands=72, f_min=50.0, f_max=7800.0, amp="magnitude")).compute(x.astype(np.float64))
    assert rel_l2(l, rl) <= tol(dtype)


def test_dlpack_export_is_device_resident():
    torch = _torch()
    x = torch.randn(16000, device="cuda", dtype=torch.float32)
    params = sg.SpectrogramParams(sg.StftParams(512, 160), 16000.0)
    spec = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(40, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32").compute(x)
    assert spec.__dlpack_device__()[0] == 2          # kDLCUDA
    t = torch.from_dlpack(spec)
    assert t.is_cuda and t.data_ptr() == spec.data.data_ptr() and tuple(t.shape) == (40, 101)
    host = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(40, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32").compute(x.cpu().numpy())
    assert torch.from_dlpack(host).device.type == "cpu" and np.array_equal(host.data, spec.data.cpu().numpy())
