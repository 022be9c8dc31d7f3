#!/usr/bin/env python
"""Library yardstick on the same GPU ("extra yardstick"): the Whisper / music log-mel configs through
torch.stft (cuFFT) + a dense filterbank matmul (cuBLAS) + elementwise dB, configured to the reference's semantics
(zero centre padding, symmetric Hann, unnormalised FFT, power, 10*log10(max(p, 1e-8))). Not the oracle and not a
parity check -- a speed comparison of the fused sm_100a kernels against the stock-library pipeline, with inputs resident
in HBM and CUDA-event timing like bench.py. Usage: python tools/torch_yardstick.py [whisper|music] [steps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrograms_b200 as sg  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "whisper"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = {"whisper": dict(n_clips=1024, n=480000, sr=16000.0, n_fft=400, hop=160),
       "music": dict(n_clips=512, n=661500, sr=22050.0, n_fft=2048, hop=512)}[wl]
dev = torch.device("cuda", 0)
params = sg.SpectrogramParams(sg.StftParams(cfg["n_fft"], cfg["hop"], sg.WindowType.hanning(), True), cfg["sr"])
plan = sg.SpectrogramPlanner(0).mel_plan(params, sg.MelParams(128, 0.0, cfg["sr"] / 2), sg.LogParams(-80.0), "db", "float32")
fb = torch.from_numpy(np.asarray(plan.filterbank()[0], dtype=np.float32)).to(dev)       # (128, bins): the plan's own matrix
win = torch.from_numpy(np.asarray(plan.window(), dtype=np.float32)).to(dev)
x = torch.randn((cfg["n_clips"], cfg["n"]), device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
torch.backends.cuda.matmul.allow_tf32 = False
chunk = 128                                                                              # clips per library call (intermediates ~0.6 GB)


def library_step(out):
    for c0 in range(0, cfg["n_clips"], chunk):
        s = torch.stft(x[c0:c0 + chunk], cfg["n_fft"], cfg["hop"], window=win, center=True, pad_mode="constant", return_complex=True)
        p = s.real * s.real + s.imag * s.imag
        m = torch.matmul(fb, p)
        out[c0:c0 + chunk] = 10.0 * torch.log10(torch.clamp(m, min=1e-8))


n_frames = (cfg["n"] + 2 * (cfg["n_fft"] // 2) - cfg["n_fft"]) // cfg["hop"] + 1
out_lib = torch.empty((cfg["n_clips"], 128, n_frames), device=dev)
out_sgx = torch.empty_like(out_lib)


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


ms_lib = timed(lambda: library_step(out_lib))
ms_sgx = timed(lambda: plan.compute_batch(x, out_sgx))
frames = cfg["n_clips"] * n_frames
d = (out_lib[:8].double() - out_sgx[:8].double())
print(json.dumps({"workload": wl, "frames_per_step": frames, "library": {"pipeline": "torch.stft (cuFFT) + fp32 matmul (cuBLAS) + elementwise",
                  "ms_per_step": ms_lib, "frames_per_s": frames / (ms_lib * 1e-3)},
                  "sgx_b200": {"kernel": plan.kernel_name(), "ms_per_step": ms_sgx, "frames_per_s": frames / (ms_sgx * 1e-3)},
                  "speedup": ms_lib / ms_sgx, "max_abs_dB_difference_first_8_clips": float(d.abs().max()),
                  "torch": torch.__version__}))
