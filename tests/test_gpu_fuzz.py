"""Randomised parity sweep (pytest -m gpu): plan configurations drawn from a seeded generator -- FFT sizes of every kernel
family, arbitrary hops, windows, mappings, amplitude scales, dtypes, ragged clip lengths, padded clip strides and
deliberately mis-aligned device pointers -- each compared with the CPU oracle at the north_star tolerances. Inputs are
white noise (every bin carries energy), so the f32 budgets apply unmasked."""
import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from conftest import rel_l2

pytestmark = pytest.mark.gpu

SR = 16000.0
POW2 = [256, 512, 1024, 2048, 4096]
OTHER = [400, 2, 3, 7, 30, 98, 250, 401, 600, 1000, 1009]
MIXED = [64, 128, 160, 200, 240, 320, 400, 480, 500, 600, 640, 800, 960, 1000, 1200, 1600]      # r2c_fused_mixed
WINDOWS = [("hanning", 0.0), ("hamming", 0.0), ("blackman", 0.0), ("rectangular", 0.0), ("kaiser", 8.6), ("gaussian", 40.0)]


def draw(rng):
    u = rng.random()
    n_fft = int(rng.choice(POW2 if u < 0.35 else (MIXED if u < 0.7 else OTHER)))
    if n_fft == 400 and rng.random() < 0.6:
        hop = 160                                                   # the r2c_fused_n400 family
    else:
        hop = int(rng.integers(1, n_fft + 1)) if n_fft < 64 or rng.random() < 0.3 else int(rng.choice([n_fft // 4, n_fft // 2, n_fft // 3 + 1, n_fft]))
    hop = max(1, min(hop, n_fft))
    win, prm = WINDOWS[int(rng.integers(len(WINDOWS)))]
    dtype = "float32" if rng.random() < 0.5 else "float64"
    mapping = str(rng.choice(["linear", "mel", "erb", "loghz", "stft", "chroma"])) if n_fft >= 30 else str(rng.choice(["linear", "stft"]))
    amp = str(rng.choice(["power", "magnitude", "db"]))
    n_samples = int(rng.integers(1, 6000))
    n_clips = int(rng.integers(1, 4))
    return dict(n_fft=n_fft, hop=hop, window=win, prm=prm, centre=bool(rng.random() < 0.8), dtype=dtype, mapping=mapping, amp=amp,
                n_samples=n_samples, n_clips=n_clips, pad=int(rng.integers(0, 5)), misalign=int(rng.integers(0, 2)))


def build(c):
    wt = {"hanning": sg.WindowType.hanning(), "hamming": sg.WindowType.hamming(), "blackman": sg.WindowType.blackman(),
          "rectangular": sg.WindowType.rectangular(), "kaiser": sg.WindowType.kaiser(c["prm"]), "gaussian": sg.WindowType.gaussian(c["prm"])}[c["window"]]
    sp = sg.SpectrogramParams(sg.StftParams(c["n_fft"], c["hop"], wt, c["centre"]), SR)
    od = dict(dtype="f32" if c["dtype"] == "float32" else "f64", n_fft=c["n_fft"], hop=c["hop"], sample_rate=SR, window=c["window"],
              window_param=c["prm"], centre=c["centre"])
    db = sg.LogParams(-80.0) if c["amp"] == "db" else None
    pl = sg.SpectrogramPlanner()
    m = c["mapping"]
    if m == "stft":
        return sg.StftPlan(sp, c["dtype"]), oracle.Plan(oracle.Desc(**od)), "stft"
    if m == "chroma":
        cp = sg.ChromaParams(440.0, 32.7, 4186.0, str(np.random.default_rng(c["n_samples"]).choice(["none", "l1", "l2", "max"])))
        return sg.ChromaPlan(sp.stft, SR, cp, c["dtype"]), cp, "chroma"
    if m == "linear":
        return pl.linear_plan(sp, db, c["amp"], c["dtype"]), oracle.Plan(oracle.Desc(**od, amp=c["amp"], floor_db=-80.0 if db else None)), "spec"
    nb = 20 if c["n_fft"] < 256 else 64
    if m == "mel":
        plan = pl.mel_plan(sp, sg.MelParams(nb, 0.0, SR / 2), db, c["amp"], c["dtype"])
        o = oracle.Desc(**od, mapping="mel", n_bands=nb, f_min=0.0, f_max=SR / 2, amp=c["amp"], floor_db=-80.0 if db else None)
    elif m == "erb":
        plan = pl.erb_plan(sp, sg.ErbParams(nb, 50.0, SR / 2), db, c["amp"], c["dtype"])
        o = oracle.Desc(**od, mapping="erb", n_bands=nb, f_min=50.0, f_max=SR / 2, amp=c["amp"], floor_db=-80.0 if db else None)
    else:
        plan = pl.log_hz_plan(sp, sg.LogHzParams(nb, 60.0, SR / 2), db, c["amp"], c["dtype"])
        o = oracle.Desc(**od, mapping="loghz", n_bands=nb, f_min=60.0, f_max=SR / 2, amp=c["amp"], floor_db=-80.0 if db else None)
    return plan, oracle.Plan(o), "spec"


@pytest.mark.parametrize("seed", range(12))
def test_random_configurations_match_oracle(seed):
    import torch
    rng = np.random.default_rng(1000 + seed)
    checked = 0
    for _ in range(16):
        c = draw(rng)
        dt = np.float32 if c["dtype"] == "float32" else np.float64
        plan, ref, kind = build(c)
        stride = c["n_samples"] + c["pad"]
        host = rng.standard_normal((c["n_clips"], stride)).astype(dt)
        # a device buffer whose first clip starts `misalign` elements into the allocation: defeats the vector-load paths
        flat = torch.zeros(c["n_clips"] * stride + 1, dtype=torch.float32 if dt == np.float32 else torch.float64, device="cuda")
        view = flat[c["misalign"]:c["misalign"] + c["n_clips"] * stride].view(c["n_clips"], stride)
        view.copy_(torch.from_numpy(host))
        clips = view[:, :c["n_samples"]]                        # (n_clips, n_samples) with row stride `stride`
        got = plan.compute_batch(clips)
        got = got.cpu().numpy() if hasattr(got, "cpu") else got
        tol = 1e-5 if dt == np.float32 else 1e-12
        for i in range(c["n_clips"]):
            x = host[i, :c["n_samples"]]
            if kind == "stft":
                want = ref.stft(x)
            elif kind == "chroma":
                want = oracle.chromagram(x, c["n_fft"], c["hop"], SR, window=c["window"], centre=c["centre"], norm=ref.norm) \
                    if c["window"] in ("hanning", "hamming", "blackman", "rectangular") else None
                if want is None:
                    continue
            else:
                want = ref.compute(x)
            assert got[i].shape == want.shape, c
            if kind == "spec" and c["amp"] == "db":
                assert np.abs(got[i].astype(np.float64) - want.astype(np.float64)).max() <= 1e-3, c
            else:
                assert rel_l2(got[i], want) <= tol, (c, rel_l2(got[i], want))
            checked += 1
    assert checked >= 16
