import sys, torch
sys.path.insert(0, "/root/repo")
import spectrograms_b200 as sg
x = torch.randn((512, 480000), device="cuda")
sp = sg.SpectrogramParams(sg.StftParams(800, 200, "hanning", True), 16000.0)
plan = sg.SpectrogramPlanner(0).mel_plan(sp, sg.MelParams(80, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
out = plan.compute_batch(x)
for _ in range(3): plan.compute_batch(x, out)
torch.cuda.synchronize()
