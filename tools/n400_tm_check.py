"""A/B check of r2c_fused_n400_tm (TMEM exchange, warp-specialised filterbank warps) against r2c_fused_n400 (shared-memory
exchange) and the oracle on the GPU box, then CUDA-event timings of both on the configs[1] batch.
Usage: python tools/n400_tm_check.py [--quick] [--time-only] [--modes 0,1,2]"""
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


def plan_for(amp, tm, mapping="mel", nb=128):
    """tm: None = shared-memory kernel, 4 / 5 / 6 = TMEM kernel with that many warps per 32-frame group"""
    if tm is not None:
        os.environ["SGX_N400_TM_WARPS"] = str(tm)
    params = sg.SpectrogramParams(sg.StftParams(400, 160, "hanning", True), 16000.0)
    pl = sg.SpectrogramPlanner()
    db = sg.LogParams(-80.0) if amp == "db" else None
    if mapping == "mel":
        p = pl.mel_plan(params, sg.MelParams(nb, 0.0, 8000.0), db, amp, "float32")
    else:
        p = pl.log_hz_plan(params, sg.LogHzParams(nb, 60.0, 7000.0), db, amp, "float32")
    os.environ.pop("SGX_N400_TM_WARPS", None)
    p.set_tmem_exchange(tm is not None)
    return p


def main():
    quick = "--quick" in sys.argv
    modes = [4]
    if "--modes" in sys.argv:
        modes = [int(v) for v in sys.argv[sys.argv.index("--modes") + 1].split(",")]
    rng = np.random.default_rng(0)
    if "--time-only" not in sys.argv:
        for mapping, nb in (("mel", 128), ("mel", 80), ("loghz", 48)):
            for n in (48000, 5000, 161, 16000 if quick else 480000):
                for n_clips in (1, 5, 37):
                    x = rng.standard_normal((n_clips, n)).astype(np.float32)
                    xd = torch.from_numpy(x).cuda()
                    for amp in ("power", "db", "magnitude"):
                        b = plan_for(amp, None, mapping, nb)
                        yb = b.compute_batch(xd).cpu().numpy()
                        od = oracle.Desc(dtype="f64", n_fft=400, hop=160, sample_rate=16000.0, mapping=mapping, n_bands=nb,
                                         f_min={"mel": 0.0, "loghz": 60.0}[mapping], f_max={"mel": 8000.0, "loghz": 7000.0}[mapping],
                                         amp=amp, floor_db=-80.0 if amp == "db" else None)
                        ref = oracle.Plan(od).compute(x[-1].astype(np.float64))
                        for tm in modes:
                            a = plan_for(amp, tm, mapping, nb)
                            assert a.kernel_name() == "r2c_fused_n400_tm", a.kernel_name()
                            ya = a.compute_batch(xd).cpu().numpy()
                            same = bool(np.array_equal(ya, yb))
                            if amp == "db":
                                e_tm, e_cc = float(np.abs(ya[-1] - ref).max()), float(np.abs(yb[-1] - ref).max())
                            else:
                                e_tm, e_cc = rel(ya[-1], ref), rel(yb[-1], ref)
                            print(f"{mapping:5s} nb={nb:3d} n={n:6d} clips={n_clips:2d} {amp:9s} warps={tm}: tm {e_tm:.3e}  smem {e_cc:.3e}  "
                                  f"bit-identical={same}", flush=True)
                            tol = 1e-3 if amp == "db" else 1e-5
                            assert e_tm <= tol, "TMEM kernel out of tolerance"
    clips = torch.randn((1024, 480000), dtype=torch.float32, device="cuda")
    res = {}
    out = None
    for tm in [None] + modes:
        p = plan_for("db", tm)
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
        key = f"{p.kernel_name()}" + ("" if tm is None else f"/warps{tm}")
        res[key] = {"ms_median": float(np.median(ts)), "ms_best": float(np.min(ts))}
        print(key, res[key], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
