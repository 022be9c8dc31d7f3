"""r2c_fused_n400_tm -- the n_fft 400 / hop 160 f32 kernel whose two FFT passes exchange through tensor memory (four
free-running 32-frame groups per SM, filterbank rows of tile t and pass 1 of tile t+1 in one phase) -- against the oracle
and against the shared-memory kernel of the same family (same task functions: the outputs must be bit-identical):
sparse mappings (mel, LogHz) x amplitude scale, centre on/off, clip lengths around tile and super-tile edges (a CTA's 4
groups take 4 consecutive tiles), many short clips, a long clip, unaligned inputs, host pointers, concurrent plans on
two streams (each CTA owns all 512 TMEM columns of its SM), the selection rule and the opt-out.
Tolerances are north_star's: f32 rel-L2 <= 1e-5, dB within 1e-3 dB."""
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


def P(centre=True):
    return sg.SpectrogramParams(sg.StftParams(400, 160, sg.WindowType.hanning(), centre), 16000.0)


def od(centre=True, **kw):
    return oracle.Desc(dtype="f64", n_fft=400, hop=160, sample_rate=16000.0, centre=centre, **kw)


def tm(plan):
    assert plan.kernel_name() == "r2c_fused_n400_tm"            # the default for these plans
    return plan


def check(plan, odesc, x, amp):
    t = _torch().from_numpy(x).cuda()
    got = tm(plan).compute(t).data.cpu().numpy()
    ref = oracle.Plan(odesc).compute(x.astype(np.float64))
    assert got.shape == ref.shape
    if amp == "db":
        assert np.abs(got - ref).max() <= TOL_DB
    else:
        assert rel_l2(got, ref) <= TOL_F32
    plan.set_tmem_exchange(False)
    assert plan.kernel_name() == "r2c_fused_n400"
    other = plan.compute(t).data.cpu().numpy()
    assert np.array_equal(got, other)                           # same arithmetic, only the exchange medium differs
    plan.set_tmem_exchange(None)
    assert plan.kernel_name() == "r2c_fused_n400_tm"


@pytest.mark.parametrize("centre", [True, False])
@pytest.mark.parametrize("amp", ["power", "magnitude", "db"])
def test_tm_all_sparse_mappings(centre, amp):
    x = make_signal("noise", 40000, 16000.0, np.float32, seed=21)
    db = sg.LogParams(-75.0) if amp == "db" else None
    okw = dict(amp=amp, floor_db=-75.0 if amp == "db" else None)
    pl = sg.SpectrogramPlanner()
    check(pl.mel_plan(P(centre), sg.MelParams(128, 0.0, 8000.0), db, amp, "float32"),
          od(centre, mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(80, 100.0, 7000.0, "slaney"), db, amp, "float32"),
          od(centre, mapping="mel", n_bands=80, f_min=100.0, f_max=7000.0, mel_norm="slaney", **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(37, 0.0, 8000.0, "l2"), db, amp, "float32"),          # rows not a multiple of 4
          od(centre, mapping="mel", n_bands=37, f_min=0.0, f_max=8000.0, mel_norm="l2", **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(300, 0.0, 8000.0), db, amp, "float32"),               # 75 quads, empty rows
          od(centre, mapping="mel", n_bands=300, f_min=0.0, f_max=8000.0, **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(8, 0.0, 8000.0), db, amp, "float32"),                 # 2 quads of long rows (several 4-column steps)
          od(centre, mapping="mel", n_bands=8, f_min=0.0, f_max=8000.0, **okw), x, amp)
    check(pl.mel_plan(P(centre), sg.MelParams(3, 300.0, 8000.0), db, amp, "float32"),               # fewer rows than warps in a group
          od(centre, mapping="mel", n_bands=3, f_min=300.0, f_max=8000.0, **okw), x, amp)
    check(pl.log_hz_plan(P(centre), sg.LogHzParams(96, 40.0, 7900.0), db, amp, "float32"),
          od(centre, mapping="loghz", n_bands=96, f_min=40.0, f_max=7900.0, **okw), x, amp)


@pytest.mark.parametrize("sig", ["sine", "chirp"])
def test_tm_tones_within_the_reference_f32_noise(sig):
    """f32 dB on tones: compared on the elements within 60.2 dB of the frame maximum (as test_gpu_parity.db_check)."""
    x = make_signal(sig, 48000, 16000.0, np.float32)
    plan = tm(sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32"))
    got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy().astype(np.float64)
    ref = oracle.Plan(od(mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0)).compute(x.astype(np.float64))
    mask = ref >= (ref.max(axis=0, keepdims=True) - 60.2)
    d = np.abs(got - ref)
    print(f"[{sig}] masked max {d[mask].max():.3e} dB, unmasked max {d.max():.3e} dB")
    assert d[mask].max() <= TOL_DB


@pytest.mark.parametrize("n", [1, 159, 161, 400, 5119, 5121, 5360, 20319, 20320, 20321, 20481, 40801])
def test_tm_clip_lengths_around_tile_and_supertile_edges(n):
    x = make_signal("noise", n, 16000.0, np.float32, seed=n)
    for centre in (True, False):
        plan = tm(sg.SpectrogramPlanner().mel_plan(P(centre), sg.MelParams(40, 0.0, 8000.0), None, "power", "float32"))
        got = plan.compute(_torch().from_numpy(x).cuda()).data.cpu().numpy()
        ref = oracle.Plan(od(centre, mapping="mel", n_bands=40, f_min=0.0, f_max=8000.0)).compute(x.astype(np.float64))
        assert got.shape == ref.shape == (40, oracle.frame_count(n, 400, 160, centre))
        assert rel_l2(got, ref) <= TOL_F32


def test_tm_batches_short_long_unaligned_and_host_pointers():
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(3)
    plan = tm(sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32"))
    ref = oracle.Plan(od(mapping="mel", n_bands=128, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0))
    short = torch.randn((5001, 700), generator=g, device="cuda")                    # 5 frames per clip: one partial tile each
    out = plan.compute_batch(short)
    assert tuple(out.shape) == (5001, 128, 5)
    for i in (0, 2499, 5000):
        assert np.abs(out[i].cpu().numpy() - ref.compute(short[i].cpu().numpy().astype(np.float64))).max() <= TOL_DB
    long = torch.randn((1, 16000 * 600), generator=g, device="cuda")                # 60001 frames = 1876 tiles = 469 super-tiles
    out = plan.compute_batch(long)
    assert tuple(out.shape) == (1, 128, 60001)
    r = ref.compute(long[0, :48000].cpu().numpy().astype(np.float64))
    assert np.abs(out[0, :, :290].cpu().numpy() - r[:, :290]).max() <= TOL_DB
    tail = ref.compute(long[0, -48000:].cpu().numpy().astype(np.float64))
    assert np.abs(out[0, :, -290:].cpu().numpy() - tail[:, -290:]).max() <= TOL_DB
    base = torch.randn((6, 30001), generator=g, device="cuda")
    for view in (base[:, 1:], base[:, :30000], base[:, 3:29000]):                    # scalar load path
        got = plan.compute_batch(view).cpu().numpy()
        host = view.cpu().numpy()
        for i in (0, 5):
            assert np.abs(got[i] - ref.compute(host[i].astype(np.float64))).max() <= TOL_DB
    base2 = torch.randn((6, 30002), generator=g, device="cuda")
    for view in (base2[:, 2:], base2[:, 2:29002]):         # 8-byte but not 16-byte aligned: cp.async (not bulk-copy) staged instantiation
        got = plan.compute_batch(view).cpu().numpy()
        host = view.cpu().numpy()
        for i in (0, 5):
            assert np.abs(got[i] - ref.compute(host[i].astype(np.float64))).max() <= TOL_DB
    two = torch.randn((700, 8000), generator=g, device="cuda")                       # 51 frames = 2 tiles per clip, both edge tiles, 1400 tiles
    out = plan.compute_batch(two)
    smem = sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    smem.set_tmem_exchange(False)
    assert torch.equal(out, smem.compute_batch(two))       # contiguous tile runs / bulk-staged edge tiles: same bits as the shared-memory kernel
    for i in (0, 350, 699):
        assert np.abs(out[i].cpu().numpy() - ref.compute(two[i].cpu().numpy().astype(np.float64))).max() <= TOL_DB
    hx = np.random.default_rng(5).standard_normal((7, 33333)).astype(np.float32)     # host pointers through the staging pipeline
    got = plan.compute_batch(hx)
    assert isinstance(got, np.ndarray)
    for i in (0, 6):
        assert np.abs(got[i] - ref.compute(hx[i].astype(np.float64))).max() <= TOL_DB


def test_tm_two_plans_on_two_streams_and_repeatability():
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(9)
    a_in = torch.randn((64, 48000), generator=g, device="cuda")
    b_in = torch.randn((64, 48000), generator=g, device="cuda")
    pa = tm(sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32"))
    pb = tm(sg.SpectrogramPlanner().log_hz_plan(P(), sg.LogHzParams(64, 50.0, 7000.0), None, "power", "float32"))
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


def test_tm_full_size_batch_matches_the_shared_memory_kernel():
    """BASELINE configs[1] at a quarter of its clip count: every element of the two kernels' outputs is identical."""
    torch = _torch()
    g = torch.Generator(device="cuda").manual_seed(1)
    clips = torch.randn((256, 480000), generator=g, device="cuda")
    plan = tm(sg.SpectrogramPlanner().mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32"))
    a = plan.compute_batch(clips)
    plan.set_tmem_exchange(False)
    b = plan.compute_batch(clips)
    assert tuple(a.shape) == (256, 128, 3001) and torch.equal(a, b)


def test_tm_selection_rule():
    pl = sg.SpectrogramPlanner()
    assert pl.mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), None, "power", "float32").kernel_name() == "r2c_fused_n400_tm"
    assert pl.log_hz_plan(P(), sg.LogHzParams(48, 60.0, 7000.0), None, "magnitude", "float32").kernel_name() == "r2c_fused_n400_tm"
    assert pl.linear_plan(P(), None, "power", "float32").kernel_name() == "r2c_fused_n400"            # identity mapping: general epilogue
    assert pl.erb_plan(P(), sg.ErbParams(40, 50.0, 8000.0), None, "power", "float32").kernel_name() == "r2c_fused_n400_tc"
    assert not pl.mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), None, "power", "float64").kernel_name().startswith("r2c_fused_n400")
    from spectrograms_b200.plan import _NativePlan, _OUT_MFCC
    mf = _NativePlan(P(), "float32", "mel", sg.MelParams(128, 0.0, 8000.0), "db", sg.LogParams(-80.0), _OUT_MFCC, sg.MfccParams(13, True, 0))
    assert mf.kernel_name() == "r2c_fused_n400_tm+dct2_lifter_tc"                                      # log-mel by this kernel, DCT-II on the tensor cores
    mf.set_tmem_exchange(False)
    assert mf.kernel_name() == "r2c_fused_n400"                                                        # opt-out: the fused shared-memory kernel
    p = pl.mel_plan(P(), sg.MelParams(128, 0.0, 8000.0), None, "power", "float32")
    p.set_tensor_cores(True)
    assert p.kernel_name() == "r2c_fused_n400_tc"                                                      # an explicit request wins
    p.set_tensor_cores(None)
    p.force_generic(True)
    assert p.kernel_name() == "r2c_fused_generic"
