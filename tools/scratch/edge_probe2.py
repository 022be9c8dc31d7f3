import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spectrograms_b200 as sg
def t(n_clips, n, centre=True):
    params = sg.SpectrogramParams(sg.StftParams(400, 160, "hanning", centre), 16000.0)
    p = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    x = torch.randn((n_clips, n), dtype=torch.float32, device="cuda")
    out = p.compute_batch(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p.compute_batch(x, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    tiles = n_clips * ((out.shape[-1] + 31) // 32)
    print(f"clips={n_clips} n={n} tiles={tiles} per_group={tiles/592:.2f} ms={np.median(ts):.4f} best={np.min(ts):.4f} us_per_slot={1e3*np.median(ts)/(tiles/592):.3f}", flush=True)
for nc in (37, 74, 148, 296, 592, 1184, 2368):
    t(nc, 480000)
for nc in (111, 444, 1776, 3552):
    t(nc, 160000)
