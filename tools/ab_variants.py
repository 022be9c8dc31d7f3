"""Same-box alternating A/B timing of compile-time variants of the library on the configs[1] batch.
Variants are whole builds kept under spectrograms_b200/lib/variants/libsgx_<name>.so (built with extra -D flags through
spectrograms_b200.build.build(extra=...)); every measurement runs in its own process so that each loads exactly one build.
Usage: python tools/ab_variants.py A B C [--rounds 3] [--clips 1024] [--seconds 30]
       python tools/ab_variants.py --one <name>        (internal: prints one JSON line)"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(name, clips, seconds):
    import numpy as np
    import torch
    import spectrograms_b200._native as nat
    nat.LIB_PATH = os.path.join(ROOT, "spectrograms_b200", "lib", "variants", f"libsgx_{name}.so")
    import spectrograms_b200 as sg
    params = sg.SpectrogramParams(sg.StftParams(400, 160, "hanning", True), 16000.0)
    plan = sg.SpectrogramPlanner().mel_plan(params, sg.MelParams(128, 0.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((clips, int(16000 * seconds)), device="cuda", generator=g)
    out = plan.compute_batch(x)
    for _ in range(5):
        plan.compute_batch(x, out=out)
    torch.cuda.synchronize()
    n = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        plan.compute_batch(x, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    small = out[:4].cpu().numpy()
    print(json.dumps({"name": name, "kernel": plan.kernel_name(), "median_ms": ts[n // 2], "min_ms": ts[0],
                      "mean_ms": ev[0].elapsed_time(ev[n]) / n, "sha": hashlib.sha256(small.tobytes()).hexdigest()[:16],
                      "sum": float(np.abs(small.astype(np.float64)).sum())}))


def main():
    a = sys.argv[1:]
    clips = int(a[a.index("--clips") + 1]) if "--clips" in a else 1024
    seconds = float(a[a.index("--seconds") + 1]) if "--seconds" in a else 30.0
    if "--one" in a:
        return one(a[a.index("--one") + 1], clips, seconds)
    rounds = int(a[a.index("--rounds") + 1]) if "--rounds" in a else 3
    names = [v for i, v in enumerate(a) if not v.startswith("--") and (i == 0 or a[i - 1] not in ("--rounds", "--clips", "--seconds"))]
    for r in range(rounds):
        for nm in names:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", nm, "--clips", str(clips), "--seconds", str(seconds)],
                               capture_output=True, text=True)
            print(p.stdout.strip() or p.stderr[-500:], flush=True)


if __name__ == "__main__":
    main()
