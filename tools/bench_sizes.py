#!/usr/bin/env python
"""Mel-dB throughput across common (n_fft, hop, n_mels) front-end shapes on one GPU: 512 clips x 30 s @16 kHz, f32,
inputs resident in HBM, CUDA events. Usage: python tools/bench_sizes.py [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrograms_b200 as sg  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sr, n, clips = 16000.0, 480000, 512
dev = torch.device("cuda", 0)
x = torch.randn((clips, n), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
rows = []
for n_fft, hop, n_mels in ((400, 160, 80), (400, 160, 128), (512, 160, 80), (512, 128, 64), (1024, 256, 80), (1024, 256, 128), (2048, 512, 128), (800, 200, 80),
                           (400, 200, 80), (480, 160, 80), (960, 240, 80), (1000, 250, 128), (1200, 300, 128), (1600, 400, 128)):
    sp = sg.SpectrogramParams(sg.StftParams(n_fft, hop, "hanning", True), sr)
    plan = sg.SpectrogramPlanner(0).mel_plan(sp, sg.MelParams(n_mels, 0.0, sr / 2), sg.LogParams(-80.0), "db", "float32")
    out = plan.compute_batch(x)
    for _ in range(2):
        plan.compute_batch(x, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.compute_batch(x, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    frames = clips * out.shape[2]
    byts = clips * n * 4 + out.numel() * 4
    generic_ms = None
    if plan.kernel_name() == "r2c_fused_mixed":          # the same plan on the generic family, for the ratio
        plan.force_generic(True)
        plan.compute_batch(x, out)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            plan.compute_batch(x, out)
        e1.record()
        torch.cuda.synchronize()
        generic_ms = round(e0.elapsed_time(e1) / 3, 4)
        plan.force_generic(False)
    rows.append({"generic_ms": generic_ms, "n_fft": n_fft, "hop": hop, "n_mels": n_mels, "kernel": plan.kernel_name(), "ms_per_step": round(ms, 4),
                 "frames_per_s": frames / (ms * 1e-3), "algorithmic_GBps": byts / (ms * 1e-3) / 1e9})
print(json.dumps({"workload": f"{clips} clips x 30 s @16 kHz f32 mel dB", "rows": rows}))
