"""Every code path of the r2c_fused_n400 family (n_fft 400 / hop 160, f32) against the oracle: sparse schedule (mel,
LogHz), general epilogue (linear, ERB, oversized mel), amplitude variants and the Decibels-without-floor quirk, centre
on/off, unaligned / odd-stride inputs (scalar load path), clip-edge tiles, many short clips, one very long clip, and two
plans running concurrently on different streams."""
import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from conftest import make_signal, rel_l2

pytestmark = pytest.mark.gpu
TOL_F32, TOL_DB = 1e-5, 1e-3


def _torch():
    import torch
    return torch


def P(centre=True, sr=16000.0):
    return sg.SpectrogramParams(sg.StftParams(400, 160, sg.WindowType.hanning(), centre), sr)


def od(centre=True, dtype_override="f64", **kw):
    return oracle.Desc(dtype=dtype_override, n_fft=400, hop=160, sample_rate=16000.0, centre=centre, **kw)


def check(plan, odesc, x, amp):
    assert plan.kernel_name().startswith("r2c_fused_n400")
    got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
    ref = oracle.Plan(odesc).compute(x.astype(np.float64))
    assert got.shape == ref.shape
    if amp == "db":
        assert np.abs(got - ref).max() <= TOL_DB
    else:
        assert rel_l2(got, ref) <= TOL_F32
    return got


@pytest.mark.parametrize("centre", [True, False])
@pytest.mark.parametrize("amp", ["power", "magnitude", "db"])
def test_all_mappings(centre, amp):
    x = make_signal("noise", 40000, 16000.0, np.float32, seed=21)
    db = sg.LogParams(-75.0) if amp == "db" else None
    okw = dict(amp=amp, floor_db=-75.0 if amp == "db" else None)
    pl = sg.SpectrogramPlanner()
    check(pl.mel_plan(P(centre), sg.MelParams(128, 0.0, 8000.0), db, amp, "float32"),
          od(centre, mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(80, 100.0, 7000.0, "slaney"), db, amp, "float32"),
          od(centre, mapping="mel", n_bands=80, f_min=100.0, f_max=7000.0, mel_norm="slaney", **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(37, 0.0, 8000.0, "l2"), db, amp, "float32"),         # rows not a multiple of 4
          od(centre, mapping="mel", n_bands=37, f_min=0.0, f_max=8000.0, mel_norm="l2", **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(500, 0.0, 8000.0), db, amp, "float32"),              # more mels than bins: empty rows, big table
          od(centre, mapping="mel", n_bands=500, f_min=0.0, f_max=8000.0, **okw), x, amp)
    check(pl.log_hz_plan(P(centre), sg.LogHzParams(96, 40.0, 7900.0), db, amp, "float32"),
          od(centre, mapping="loghz", n_bands=96, f_min=40.0, f_max=7900.0, **okw), x, amp)
    check(pl.linear_plan(P(centre), db, amp, "float32"), od(centre, **okw), x, amp)
    check(pl.erb_plan(P(centre), sg.ErbParams(40, 50.0, 8000.0), db, amp, "float32"),
          od(centre, mapping="erb", n_bands=40, f_min=50.0, f_max=8000.0, **okw), x, amp)


def test_decibels_without_floor_and_windows():
    x = make_signal("chirp", 20000, 16000.0, np.float32)
    a = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(64, 0.0, 8000.0), None, "db", "float32")
    b = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(64, 0.0, 8000.0), None, "power", "float32")
    t = _torch().from_numpy(x).cuda()
    assert np.array_equal(a.compute(t).data.cpu().numpy(), b.compute(t).data.cpu().numpy())       # quirk F7
    for kind, prm in [("hamming", 0.0), ("blackman", 0.0), ("kaiser", 6.0), ("gaussian", 70.0), ("rectangular", 0.0)]:
        params = sg.SpectrogramParams(sg.StftParams(400, 160, sg.WindowType(kind, prm), True), 16000.0)
        plan = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(64, 0.0, 8000.0), None, "power", "float32")
        ref = oracle.Plan(oracle.Desc(dtype="f64", n_fft=400, hop=160, window=kind, window_param=prm, mapping="mel", n_bands=64,
                                      f_min=0.0, f_max=8000.0)).compute(x.astype(np.float64))
        assert rel_l2(plan.compute(t).data.cpu().numpy(), ref) <= TOL_F32


def test_unaligned_and_strided_inputs_take_the_scalar_path():
    torch = _torch()
    rng = np.random.default_rng(5)
    base = rng.standard_normal((6, 30001)).astype(np.float32)
    plan = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    ref = oracle.Plan(od(mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0))
    dev = torch.from_numpy(base).cuda()
    for view in (dev[:, 1:], dev[:, :30000], dev[:, 3:29000]):      # odd offsets / odd strides / odd lengths
        host = view.cpu().numpy()
        got = plan.compute_batch(view).cpu().numpy()
        for i in (0, 5):
            assert np.abs(got[i] - ref.compute(host[i].astype(np.float64))).max() <= TOL_DB


@pytest.mark.parametrize("n", [1, 2, 159, 160, 161, 399, 400, 401, 5119, 5120, 5121, 5360, 10241])
def test_clip_lengths_around_tile_edges(n):
    x = make_signal("noise", n, 16000.0, np.float32, seed=n)
    for centre in (True, False):
        plan = sg.SpectrogramPlanner().mel_plan(P(centre), sg.MelParams(40, 0.0, 8000.0), None, "power", "float32")
        got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
        ref = oracle.Plan(od(centre, mapping="mel", n_bands=40, f_min=0.0, f_max=8000.0)).compute(x.astype(np.float64))
        assert got.shape == ref.shape == (40, oracle.frame_count(n, 400, 160, centre))
        assert rel_l2(got, ref) <= TOL_F32


def test_many_short_clips_and_one_long_clip():
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(3)
    plan = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    ref = oracle.Plan(od(mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0))
    short = torch.randn((5000, 700), generator=g, device="cuda")                    # 5 frames per clip: one partial tile each
    out = plan.compute_batch(short)
    assert tuple(out.shape) == (5000, 128, 5)
    for i in (0, 2499, 4999):
        assert np.abs(out[i].cpu().numpy() - ref.compute(short[i].cpu().numpy().astype(np.float64))).max() <= TOL_DB
    long = torch.randn((1, 16000 * 600), generator=g, device="cuda")                # 10 minutes: 60001 frames, 1876 tiles
    out = plan.compute_batch(long)
    assert tuple(out.shape) == (1, 128, 60001)
    r = ref.compute(long[0, :48000].cpu().numpy().astype(np.float64))
    assert np.abs(out[0, :, :290].cpu().numpy() - r[:, :290]).max() <= TOL_DB
    tail = ref.compute(long[0, -48000:].cpu().numpy().astype(np.float64))
    assert np.abs(out[0, :, -290:].cpu().numpy() - tail[:, -290:]).max() <= TOL_DB


def test_two_plans_on_two_streams():
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(9)
    a_in = torch.randn((64, 48000), generator=g, device="cuda")
    b_in = torch.randn((64, 48000), generator=g, device="cuda")
    pa = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    pb = sg.MfccPlan(sg.StftParams(400, 160), 16000.0, 128, sg.MfccParams(40), "float32")
    ra, rb = pa.compute_batch(a_in).clone(), pb.compute_batch(b_in).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(5):
        with torch.cuda.stream(s1):
            oa = pa.compute_batch(a_in)
        with torch.cuda.stream(s2):
            ob = pb.compute_batch(b_in)
        outs.append((oa, ob))
    torch.cuda.synchronize()
    for oa, ob in outs:
        assert torch.equal(oa, ra) and torch.equal(ob, rb)


def test_floor_below_the_f32_normal_range_stays_finite():
    """LogParams only requires a finite floor. With floor_db = -400 the f32 epsilon 1e-40 is a denormal; silent bins
    must return the floor (the reference's 10 log10(max(0, eps))), not -inf from a flush-to-zero log."""
    x = np.zeros(8000, dtype=np.float32)
    x[4000:4400] = make_signal("noise", 400, 16000.0, np.float32, seed=1)
    plan = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-400.0), "db", "float32")
    assert not plan.kernel_name().startswith("r2c_fused_n400")          # routed off the ftz kernels
    got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
    ref = oracle.Plan(od(dtype_override="f32", mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-400.0)).compute(x)
    assert np.all(np.isfinite(got)) and got.min() >= -400.5
    silent = ref < -399.0
    assert silent.any() and np.abs(got[silent] - ref[silent]).max() <= 0.5 and np.abs(got[~silent] - ref[~silent]).max() <= 1e-2
    ok = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-370.0), "db", "float32")
    assert ok.kernel_name().startswith("r2c_fused_n400")               # normal eps: unchanged


def test_fused_mfcc_on_magnitude_mel_matches_the_generic_family():
    """output = MFCC with amp = magnitude fed power-mel to the DCT on the n400 family only. Both families must agree."""
    from spectrograms_b200.plan import _NativePlan, _OUT_MFCC
    x = make_signal("noise", 20000, 16000.0, np.float32, seed=4)
    t = _torch().from_numpy(x).cuda()
    outs = {}
    for amp in ("magnitude", "power"):
        n = _NativePlan(P(), "float32", "mel", sg.MelParams(128, 0.0, 8000.0), amp, None, _OUT_MFCC, sg.MfccParams(13, True, 0))
        assert n.kernel_name() == "r2c_fused_n400_tm+dct2_lifter_tc"
        a = n.compute_one(t).cpu().numpy()
        n.set_tmem_exchange(False)
        assert n.kernel_name() == "r2c_fused_n400"
        assert rel_l2(n.compute_one(t).cpu().numpy(), a) <= TOL_F32
        n.force_generic(True)
        b = n.compute_one(t).cpu().numpy()
        assert rel_l2(a, b) <= TOL_F32
        outs[amp] = a
    assert rel_l2(outs["magnitude"], outs["power"]) > 0.1              # and the two scalings really differ
