#!/usr/bin/env python
"""Throughput of the inverse path on one GPU: `clips` x 30 s @16 kHz, n_fft 512 / hop 128,
f32, STFT matrices resident in HBM, CUDA events, 3 warm-ups + `steps` timed calls. Usage: python tools/bench_istft.py [clips] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrograms_b200 as sg  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sr, n, n_fft, hop = 16000.0, 480000, 512, 128
dev = torch.device("cuda", 0)
x = torch.randn((clips, n), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
plan = sg.StftPlan(sg.SpectrogramParams(sg.StftParams(n_fft, hop, "hanning", True), sr), "float32", 0)
S = plan.compute_batch(x)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


ms_fwd = timed(lambda: plan.compute_batch(x, S))
y = plan.istft(S)
ms_inv = timed(lambda: plan.istft(S))
frames = clips * S.shape[2]
err = float((y[:, n_fft:-n_fft] - x[:, n_fft:y.shape[1] - n_fft]).abs().max())
print(json.dumps({"workload": f"{clips} clips x 30 s @16 kHz, n_fft=512 hop=128 f32", "frames": frames,
                  "stft": {"kernel": plan.kernel_name(), "ms_per_step": ms_fwd, "frames_per_s": frames / (ms_fwd * 1e-3)},
                  "istft": {"kernels": "c2r_pow2 + ola_gather" if os.environ.get("SGX_ISTFT_UNFUSED") == "1" else "istft_pow2 (fused, halo tile)",
                            "launches": plan.last_launch_count(), "ms_per_step": ms_inv, "frames_per_s": frames / (ms_inv * 1e-3)},
                  "round_trip_max_abs_error": err}))
