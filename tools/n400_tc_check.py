"""A/B check of the two kernels of the n400 family on the GPU box: r2c_fused_n400_tc (TMEM + tcgen05) against
r2c_fused_n400 (shared memory + CUDA cores) and the oracle, then CUDA-event timings of both on the configs[1] batch."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import spectrograms_b200 as sg


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


def plan_for(amp, tc, mapping="mel", nb=128):
    params = sg.SpectrogramParams(sg.StftParams(400, 160, "hanning", True), 16000.0)
    pl = sg.SpectrogramPlanner()
    db = sg.LogParams(-80.0) if amp == "db" else None
    if mapping == "mel":
        p = pl.mel_plan(params, sg.MelParams(nb, 0.0, 8000.0), db, amp, "float32")
    elif mapping == "loghz":
        p = pl.log_hz_plan(params, sg.LogHzParams(nb, 60.0, 7000.0), db, amp, "float32")
    else:
        p = pl.erb_plan(params, sg.ErbParams(nb, 50.0, 7600.0), db, amp, "float32")
    p.set_tensor_cores(tc)
    return p


def main():
    quick = "--quick" in sys.argv
    rng = np.random.default_rng(0)
    rows = []
    for mapping, nb in (("mel", 128), ("mel", 80), ("loghz", 48), ("erb", 40)):
        for n in (48000, 5000, 161, 480000 if not quick else 16000):
            for n_clips in (1, 5):
                x = rng.standard_normal((n_clips, n)).astype(np.float32)
                xd = torch.from_numpy(x).cuda()
                for amp in ("power", "db", "magnitude"):
                    a = plan_for(amp, True, mapping, nb)
                    b = plan_for(amp, False, mapping, nb)
                    assert a.kernel_name() == "r2c_fused_n400_tc", a.kernel_name()
                    ya = a.compute_batch(xd).cpu().numpy()
                    yb = b.compute_batch(xd).cpu().numpy()
                    od = oracle.Desc(dtype="f64", n_fft=400, hop=160, sample_rate=16000.0, mapping=mapping, n_bands=nb,
                                     f_min={"mel": 0.0, "loghz": 60.0, "erb": 50.0}[mapping], f_max={"mel": 8000.0, "loghz": 7000.0, "erb": 7600.0}[mapping],
                                     amp=amp, floor_db=-80.0 if amp == "db" else None)
                    ref = oracle.Plan(od).compute(x[-1].astype(np.float64))
                    if amp == "db":
                        e_tc, e_cc = float(np.abs(ya[-1] - ref).max()), float(np.abs(yb[-1] - ref).max())
                    else:
                        e_tc, e_cc = rel(ya[-1], ref), rel(yb[-1], ref)
                    rows.append((mapping, nb, n, n_clips, amp, e_tc, e_cc))
                    print(f"{mapping:5s} nb={nb:3d} n={n:6d} clips={n_clips} {amp:9s}: tc {e_tc:.3e}  cuda-core {e_cc:.3e}", flush=True)
                    tol = 1e-3 if amp == "db" else 1e-5
                    assert e_tc <= tol, "TC kernel out of tolerance"
    # timing on the configs[1] batch: banded mel (tensor cores lose) and dense ERB (tensor cores win)
    clips = torch.randn((1024, 480000), dtype=torch.float32, device="cuda")
    res = {}
    for mapping, nb in (("mel", 128), ("erb", 40), ("erb", 64)):
        out = None
        for tc in (True, False):
            p = plan_for("db", tc, mapping, nb)
            out = p.compute_batch(clips, out=out)
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                p.compute_batch(clips, out=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[f"{mapping}{nb}/{p.kernel_name()}"] = {"ms_median": float(np.median(ts)), "ms_best": float(np.min(ts))}
            print(f"{mapping}{nb}", p.kernel_name(), res[f"{mapping}{nb}/{p.kernel_name()}"], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
