import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spectrograms_b200 as sg
def t(n_clips, n, centre):
    params = sg.SpectrogramParams(sg.StftParams(400, 160, "hanning", centre), 16000.0)
    p = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    x = torch.randn((n_clips, n), dtype=torch.float32, device="cuda")
    out = p.compute_batch(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p.compute_batch(x, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    fr = out.shape[-1] * n_clips
    print(f"clips={n_clips} n={n} centre={centre} kernel={p.kernel_name()} frames/clip={out.shape[-1]} ms={np.median(ts):.4f} Gframes/s={fr/np.median(ts)/1e6:.3f}", flush=True)
for c in (True, False):
    t(1024, 160000, c)
    t(1024, 480000, c)
    t(4096, 40000, c)
    t(256, 1920000, c)
