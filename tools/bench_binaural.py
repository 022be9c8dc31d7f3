#!/usr/bin/env python
"""Throughput of the binaural cue path on one GPU: `pairs` stereo pairs x 30 s @16 kHz,
n_fft 512 / hop 128, f32, inputs resident in HBM, CUDA events, 3 warm-ups + `steps` timed calls per cue.
Usage: python tools/bench_binaural.py [pairs] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrograms_b200 as sg  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sr, n, n_fft, hop = 16000.0, 480000, 512, 128
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
left = torch.randn((pairs, n), device=dev, generator=g)
right = 0.6 * torch.roll(left, 3, dims=1) + 0.3 * torch.randn((pairs, n), device=dev, generator=g)
sp = sg.SpectrogramParams(sg.StftParams(n_fft, hop, sg.WindowType.hanning(), True), sr)
plan = sg.StftPlan(sp, "float32", 0)
res = {}
for cue, fn, prm in (("itd", sg.compute_itd_spectrogram, sg.ITDSpectrogramParams(sp, 100.0, 7900.0)),
                     ("ipd", sg.compute_ipd_spectrogram, sg.IPDSpectrogramParams(sp, 100.0, 7900.0, True)),
                     ("ild", sg.compute_ild_spectrogram, sg.ILDSpectrogramParams(sp, 100.0, 7900.0)),
                     ("ilr", sg.compute_ilr_spectrogram, sg.ILRSpectrogramParams(sp, 100.0, 7900.0))):
    for _ in range(3):
        out = fn([left, right], prm, plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn([left, right], prm, plan)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    frames = pairs * out.n_frames
    io = 2 * pairs * n * 4 + out.data.numel() * 4
    res[cue] = {"ms_per_step": ms, "pair_frames_per_s": frames / (ms * 1e-3), "algorithmic_GBps": io / (ms * 1e-3) / 1e9,
                "shape": list(out.shape)}
print(json.dumps({"workload": f"{pairs} stereo pairs x 30 s @16 kHz, n_fft=512 hop=128 f32, band 100-7900 Hz", "kernel": plan.kernel_name(),
                  "launches_per_call": plan.last_launch_count(), "cues": res}))
