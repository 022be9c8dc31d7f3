#!/usr/bin/env python
"""bench.py -- log-mel frames/s of the fused STFT -> |X|^2 -> mel -> dB path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload whisper|music|mfcc|multichannel] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic clips. The default workload is BASELINE.json
configs[1] (Whisper-style log-mel: 1024 clips x 30 s @16 kHz, n_fft 400, hop 160, 128 mels, f32); each rank owns a
full batch (weak scaling, clips shard with no data-path collective -- SURVEY.md section 8e).

Printed JSON (rank 0, one line):
  value      whole-job frames/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the C ABI with pinned HOST buffers (H2D + kernel + D2H inside the timed region)
  roofline   algorithmic HBM bytes per launch / average launch duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle's restatement of the reference algorithm on the host cores (bounded sample), rank 0, N=1
  clocks     SM clock / throttle reasons sampled during the timed region

--impl reference times the reference's CPU algorithm (the oracle port: the Rust crate cannot be built here -- no
cargo, un-vendored realfft/rustfft) with one plan per host thread on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_clips, seconds, sr, n_fft, hop, dtype, kind)
    "whisper": dict(n_clips=1024, n_samples=480000, sr=16000.0, n_fft=400, hop=160, dtype="float32", kind="mel_db",
                    label="configs[1] whisper log-mel: 1024 clips x 30 s @16 kHz, n_fft=400 hop=160 128 mels dB f32"),
    "music": dict(n_clips=512, n_samples=661500, sr=22050.0, n_fft=2048, hop=512, dtype="float32", kind="mel_db",
                  label="configs[2] music mel-dB per-GPU shard: 512 clips x 30 s @22.05 kHz, n_fft=2048 hop=512 128 mels dB f32"),
    "mfcc": dict(n_clips=1024, n_samples=160000, sr=16000.0, n_fft=400, hop=160, dtype="float32", kind="mfcc",
                 label="configs[3] MFCC per-GPU shard: 1024 clips x 10 s @16 kHz, n_fft=400 hop=160 128 mels -> 40 MFCC f32"),
    "asr512": dict(n_clips=1024, n_samples=480000, sr=16000.0, n_fft=512, hop=160, dtype="float32", kind="mel_db80",
                   label="common ASR front end (not a BASELINE config): 1024 clips x 30 s @16 kHz, n_fft=512 hop=160 80 mels dB f32"),
    "chroma": dict(n_clips=512, n_samples=661500, sr=22050.0, n_fft=2048, hop=512, dtype="float32", kind="chroma",
                   label="SURVEY 8f rank 2, chromagram() on the configs[2] shard: 512 clips x 30 s @22.05 kHz, n_fft=2048 hop=512 -> 12 pitch classes (L2) f32"),
    "erb400": dict(n_clips=1024, n_samples=480000, sr=16000.0, n_fft=400, hop=160, dtype="float32", kind="erb_db64",
                   label="dense ERB projection on the configs[1] batch (not a BASELINE config): 1024 clips x 30 s @16 kHz, n_fft=400 hop=160 64 ERB bands dB f32"),
    "erb512": dict(n_clips=1024, n_samples=480000, sr=16000.0, n_fft=512, hop=160, dtype="float32", kind="erb_db128",
                   label="dense ERB-128 projection (not a BASELINE config): 1024 clips x 30 s @16 kHz, n_fft=512 hop=160 128 ERB bands dB f32"),
    "multichannel": dict(n_clips=64, n_samples=2880000, sr=48000.0, n_fft=4096, hop=1024, dtype="float64", kind="linear_mag",
                         label="configs[4] multichannel STFT magnitude: 64 ch x 60 s @48 kHz, n_fft=4096 hop=1024 f64"),
}


# total clips of the workload as BASELINE.json states it (the per-GPU shard above is total / 8): --scaling strong shards these
BASELINE_TOTAL_CLIPS = {"whisper": 1024, "music": 4096, "mfcc": 8192, "multichannel": 64}


def frames_of(w):
    pad = w["n_fft"] // 2
    return (w["n_samples"] + 2 * pad - w["n_fft"]) // w["hop"] + 1


def out_rows(w):
    return {"mel_db": 128, "mel_db80": 80, "mfcc": 40, "chroma": 12, "erb_db64": 64, "erb_db128": 128, "linear_mag": w["n_fft"] // 2 + 1}[w["kind"]]


def algorithmic_bytes(w):
    """read every input sample once + write every output element once."""
    es = 4 if w["dtype"] == "float32" else 8
    return w["n_clips"] * (w["n_samples"] * es + out_rows(w) * frames_of(w) * es)


def make_plan(w, device=None):
    import spectrograms_b200 as sg
    params = sg.SpectrogramParams(sg.StftParams(w["n_fft"], w["hop"], sg.WindowType.hanning(), True), w["sr"])
    if w["kind"] in ("mel_db", "mel_db80"):
        return sg.SpectrogramPlanner(device).mel_plan(params, sg.MelParams(out_rows(w), 0.0, w["sr"] / 2), sg.LogParams(-80.0), "db", w["dtype"])
    if w["kind"] in ("erb_db64", "erb_db128"):
        return sg.SpectrogramPlanner(device).erb_plan(params, sg.ErbParams(out_rows(w), 50.0, w["sr"] / 2), sg.LogParams(-80.0), "db", w["dtype"])
    if w["kind"] == "mfcc":
        return sg.MfccPlan(params.stft, w["sr"], 128, sg.MfccParams(40), w["dtype"], device)
    if w["kind"] == "chroma":
        return sg.ChromaPlan(params.stft, w["sr"], sg.ChromaParams.music_standard(), w["dtype"], device)
    return sg.SpectrogramPlanner(device).linear_plan(params, None, "magnitude", w["dtype"])


def oracle_desc(w):
    import oracle
    kw = dict(dtype="f32" if w["dtype"] == "float32" else "f64", n_fft=w["n_fft"], hop=w["hop"], sample_rate=w["sr"])
    if w["kind"] in ("mel_db", "mel_db80", "mfcc"):
        kw.update(mapping="mel", n_bands=80 if w["kind"] == "mel_db80" else 128, f_min=0.0, f_max=w["sr"] / 2, amp="db", floor_db=-80.0)
    elif w["kind"] in ("erb_db64", "erb_db128"):
        kw.update(mapping="erb", n_bands=out_rows(w), f_min=50.0, f_max=w["sr"] / 2, amp="db", floor_db=-80.0)
    else:
        kw.update(amp="magnitude")
    return oracle.Desc(**kw)


def kernel_src_sha():
    """sha1 over the CUDA sources: ties profiles/traffic.json to the code it was captured from."""
    import glob
    import hashlib
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(ROOT, "spectrograms_b200", "csrc", "*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def config_of(w, world, scaling="weak", total_clips=None):
    """The `config` object of the JSON line -- built by this one function for BOTH arms so that the driver's
    same_config check compares like with like."""
    es = 4 if w["dtype"] == "float32" else 8
    c = {"workload": w["label"], "clips_per_gpu": w["n_clips"], "frames_per_clip": frames_of(w),
         "sharding": f"clips x{world}, no collective",
         "l2": f"inputs per step {w['n_clips'] * w['n_samples'] * es / 1e9:.2f} GB > 126 MB L2 (no flush needed)"}
    if scaling == "strong":
        c["total_clips"] = total_clips
    return c


PORT_NOTE = ("oracle/oracle.c, a scalar C restatement of the reference algorithm compiled -O3 -march=native -ffp-contract=off (auto-vectorised at best), one plan per thread "
             "(the crate's documented scaling recipe). The Rust crate cannot be built here (no cargo; realfft 3.5.0 / rustfft 6.4.1 "
             "un-vendored). Unlike rustfft the port has no hand-written SIMD (AVX2/AVX-512) butterflies and no cached-plan radix "
             "specialisations: expect the real crate to be up to a few times faster per core, so every GPU/CPU ratio here is an upper bound")


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml; nvidia-smi as a fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _init_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        self._nv = pynvml
        self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)

    def _sample_nvml(self):
        nv, h = self._nv, self._h
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}
        self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        for bit, nm in names.items():
            if r & bit:
                self.reasons.add(nm)

    def _run_nvml(self):
        while not self._stop.is_set():
            self._sample_nvml()
            time.sleep(0.001)

    def _run(self):
        try:
            if self._nv is None:
                raise RuntimeError("nvml unavailable")
            self._run_nvml()
        except Exception:
            import subprocess
            while not self._stop.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(int(o[0])); self.max_mhz = int(o[1])
                    for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), o[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(nm)
                except Exception:
                    break

    def __enter__(self):
        self._nv = None
        try:
            self._init_nvml()          # initialise before the timed region so that sampling starts immediately
        except Exception:
            self._nv = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline(w, sample_clips, threads, faithful=True, reps=4):
    """The oracle's one-plan-per-thread batch driver on a bounded sample of the workload: `reps` passes over the sample
    (about 10-30 core-seconds of CPU work in total), throughput = all frames processed / total wall time."""
    import oracle
    d = oracle_desc(w)
    rng = np.random.default_rng(0)
    clips = rng.standard_normal((sample_clips, w["n_samples"])).astype(np.float32 if w["dtype"] == "float32" else np.float64)
    mf = dict(n_mfcc=40, include_c0=True, lifter=22, faithful=faithful) if w["kind"] == "mfcc" else None
    oracle.compute_batch(d, clips[: max(1, threads)], threads, mfcc=mf)      # warm-up (page in, spawn once)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.compute_batch(d, clips, threads, mfcc=mf)
    total = time.perf_counter() - t0
    return reps * sample_clips * frames_of(w) / total, total


def run_reference(args, w, rank, world, scaling, total_clips):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host threads. Each step processes the SAME
    batch the GPU arm's step does (all of this rank-0 box's clips: n_clips per GPU x N at weak scaling, the fixed total at
    strong scaling) unless --ref-clips bounds it; frames/s of a CPU pass is size-independent, the config stays identical."""
    if rank != 0:
        return
    import oracle
    cores = os.cpu_count() or 1
    full = total_clips if scaling == "strong" else w["n_clips"]           # one GPU's batch: the CPU box has no N
    sample = full if args.ref_clips <= 0 else max(cores, min(full, args.ref_clips))
    d = oracle_desc(w)
    clips = np.random.default_rng(0).standard_normal((sample, w["n_samples"]), dtype=np.float32)
    if w["dtype"] != "float32":
        clips = clips.astype(np.float64)
    mf = dict(n_mfcc=40, include_c0=True, lifter=22, faithful=True) if w["kind"] == "mfcc" else None
    for _ in range(args.warmup):
        oracle.compute_batch(d, clips, cores, mfcc=mf)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.compute_batch(d, clips, cores, mfcc=mf)
    dt = time.perf_counter() - t0
    value = args.steps * sample * frames_of(w) / dt
    line = {
        "impl": "reference", "metric": "log-mel frames/sec", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32" if w["dtype"] == "float32" else "f64", "data": "synthetic (seeded white noise)",
        "config": config_of(w, world, scaling, total_clips),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} clips per step ({'the full per-GPU batch' if sample == full else 'bounded by --ref-clips'}); " + PORT_NOTE},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="whisper", choices=sorted(WORKLOADS))
    ap.add_argument("--clips", type=int, default=0, help="override clips per GPU (0 = the workload's size)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer end-to-end leg (0 = min(steps, 10))")
    ap.add_argument("--cpu-clips", type=int, default=256, help="bounded CPU-baseline sample (clips)")
    ap.add_argument("--ref-clips", type=int, default=0, help="--impl reference: clips per step (0 = the whole per-GPU batch, as the GPU arm)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns a full batch; strong: the workload's BASELINE total (music 4096, mfcc 8192 clips) is split over the ranks")
    ap.add_argument("--no-strong-line", action="store_true", help="skip the extra strong-scaling configs[2] measurement under key strong_scaling")
    ap.add_argument("--tensor-cores", default="auto", choices=["auto", "on", "off"], help="TMEM / tcgen05 kernel variant (sgx_plan_set_tensor_cores)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the generic kernel family")
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL gather of results to rank 0 (off the hot path, reported separately)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3        # timing rule: W >= 3

    w = dict(WORKLOADS[args.workload])
    if args.clips:
        w["n_clips"] = args.clips
        w["label"] += f" [clips overridden to {args.clips}]"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    total_clips = None
    if args.scaling == "strong":
        from spectrograms_b200.sharding import shard_range
        total_clips = args.clips or BASELINE_TOTAL_CLIPS.get(args.workload, w["n_clips"])
        lo, hi = shard_range(total_clips, rank if args.impl == "b200" else 0, world if args.impl == "b200" else 1)
        w["n_clips"] = max(1, hi - lo)
        w["label"] = w["label"].replace("per-GPU shard", "whole batch") + f" [strong scaling: {total_clips} clips over {world} GPUs]"

    if w["kind"] == "chroma":
        args.no_cpu = True          # the oracle has no batched chroma driver (parity: tests/test_chroma.py)
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "no batched CPU driver for the chroma workload; it is not a BASELINE config"}))
            return
    if args.impl == "reference":
        run_reference(args, w, rank, world, args.scaling, total_clips)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from spectrograms_b200.sharding import bind_host_to_device
    numa_bound = bind_host_to_device(local_rank) if world > 1 and not os.environ.get("SGX_BENCH_NO_BIND") else False
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    tdt = torch.float32 if w["dtype"] == "float32" else torch.float64
    plan = make_plan(w, local_rank)
    if args.generic:
        plan.force_generic(True)
    if args.tensor_cores != "auto" and hasattr(plan, "set_tensor_cores"):
        plan.set_tensor_cores(args.tensor_cores == "on")
    n_frames = frames_of(w)
    rows = out_rows(w)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    clips = torch.randn((w["n_clips"], w["n_samples"]), generator=g, device=dev, dtype=tdt)   # resident in HBM
    out = torch.empty((w["n_clips"], rows, n_frames), device=dev, dtype=tdt)
    frames_per_step = w["n_clips"] * n_frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        plan.compute_batch(clips, out)
    launches_per_step = plan.last_launch_count()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as cs:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            plan.compute_batch(clips, out)
            ev[i + 1].record()
        barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    kernel_ms = float(np.mean(per_launch_ms)) / max(1, launches_per_step)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    job_frames_per_step = (total_clips if args.scaling == "strong" else world * w["n_clips"]) * n_frames
    value = job_frames_per_step * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        es = 4 if w["dtype"] == "float32" else 8
        h_in = torch.empty((w["n_clips"], w["n_samples"]), dtype=tdt, pin_memory=True)
        h_in.copy_(clips)
        h_out = torch.empty((w["n_clips"], rows, n_frames), dtype=tdt, pin_memory=True)
        hin, hout = h_in.numpy(), h_out.numpy()
        ksteps = args.e2e_steps or min(args.steps, 10)
        for _ in range(2):
            plan.compute_batch(hin, hout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            plan.compute_batch(hin, hout)           # returns after the results are in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": job_frames_per_step * ksteps / float(te.item()), "unit": "frames/s",
               "h2d_bytes_per_step": w["n_clips"] * w["n_samples"] * es, "d2h_bytes_per_step": w["n_clips"] * rows * n_frames * es,
               "steps": ksteps, "ms_per_step": 1e3 * float(te.item()) / ksteps,
               "check": "host result equals device result: %s" % bool(torch.equal(h_out[:4].to(dev), out[:4]))}
        # the platform ceiling of this leg: the same pinned buffers copied both ways at once with NO kernel, all ranks together
        # (tools/ubench/host_copy.py is the stand-alone form). e2e.frac_of_copy_ceiling = how much of it the pipeline reaches.
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        d_in_flat, d_out_flat = clips.view(-1), out.view(-1)
        hi_flat, ho_flat = h_in.view(-1), h_out.view(-1)

        def copy_step():
            with torch.cuda.stream(sa):
                d_in_flat.copy_(hi_flat, non_blocking=True)
            with torch.cuda.stream(sb):
                ho_flat.copy_(d_out_flat, non_blocking=True)

        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copy_step()
        torch.cuda.synchronize()
        tc_ = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc_, op=dist.ReduceOp.MAX)
        ceil_fps = job_frames_per_step * 3 / float(tc_.item())
        e2e["copy_ceiling"] = {"value": ceil_fps, "unit": "frames/s", "ms_per_step": 1e3 * float(tc_.item()) / 3,
                               "what": "concurrent H2D + D2H of the same pinned buffers, no kernel, all ranks at once (max over ranks)"}
        e2e["frac_of_copy_ceiling"] = e2e["value"] / ceil_fps
        del h_in, h_out, hi_flat, ho_flat

    # ---- optional result gather to rank 0 over NCCL (NVLink / NVSwitch): NOT part of the hot path or of `value`
    gather = None
    if args.gather and world > 1:
        from spectrograms_b200.sharding import gather_to_rank0
        n_g = min(w["n_clips"], 128)                  # bounded: rank 0 receives world * n_g clips
        gather_to_rank0(out[:n_g], world * n_g)       # warm-up (NCCL communicator setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = gather_to_rank0(out[:n_g], world * n_g)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gbytes = (world - 1) * n_g * rows * n_frames * (4 if w["dtype"] == "float32" else 8)
        ok = bool(rank != 0 or (full is not None and torch.equal(full[:n_g], out[:n_g])))
        gather = {"clips_per_rank": n_g, "ms": float(gms.item()), "bytes_into_rank0": gbytes,
                  "GBps_into_rank0": gbytes / (float(gms.item()) * 1e-3) / 1e9, "rank0_block_intact": ok}

    # ---- strong scaling of configs[2] as BASELINE states it (4096 music clips split over the N ranks), same run, own key
    strong = None
    if args.workload == "whisper" and args.scaling == "weak" and not args.no_strong_line and not args.generic and not args.clips:
        from spectrograms_b200.sharding import shard_range
        del clips, out
        torch.cuda.empty_cache()
        ws = dict(WORKLOADS["music"])
        tot = BASELINE_TOTAL_CLIPS["music"]
        lo, hi = shard_range(tot, rank, world)
        ws["n_clips"] = hi - lo
        splan = make_plan(ws, local_rank)
        nfs = frames_of(ws)
        sclips = torch.randn((ws["n_clips"], ws["n_samples"]), generator=g, device=dev, dtype=torch.float32)
        sout = torch.empty((ws["n_clips"], out_rows(ws), nfs), device=dev, dtype=torch.float32)
        ssteps = 5
        for _ in range(3):
            splan.compute_batch(sclips, sout)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(ssteps):
            splan.compute_batch(sclips, sout)
        s1.record()
        barrier()
        sms = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {"workload": f"configs[2] music mel-dB: {tot} clips x 30 s @22.05 kHz, n_fft=2048 hop=512 128 mels dB f32, batch-sharded over {world} GPU(s)",
                  "scaling": "strong", "total_clips": tot, "clips_this_rank": ws["n_clips"], "metric": "log-mel frames/sec",
                  "value": tot * nfs * ssteps / (float(sms.item()) * 1e-3), "unit": "frames/s", "ms_per_step": float(sms.item()) / ssteps,
                  "steps": ssteps, "warmup": 3, "kernel": splan.kernel_name(), "gpu_launches": splan.last_launch_count() * ssteps}
        del sclips, sout

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # a step may take several launches (chunks of one kernel, or the log-mel kernel followed by the DCT kernel): the roofline
    # is bytes per launch over the average launch duration = bytes per step over the step duration
    alg_bytes = algorithmic_bytes(w) / max(1, launches_per_step)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel -- reported only while the kernel
    # sources still hash to what was captured (profiles/traffic.json: src_sha), otherwise null (a stale number is worse)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            ent = json.load(open(tpath)).get(args.workload, {}).get(plan.kernel_name())
            if isinstance(ent, dict):
                if ent.get("src_sha") == kernel_src_sha():
                    traffic, traffic_src = ent["bytes"], ent.get("source")
                else:
                    traffic_src = "stale: kernel sources changed since the capture (%s)" % ent.get("source")
        except Exception:
            traffic = None
    # FP roofline next to the HBM one (the path is FFT arithmetic, not a copy): executed FP lane-operations per frame
    # (profiles/fp_ops.json, from the ncu SASS opcode mix) x frames/s against lanes/clk/SM x SMs x the sampled SM clock
    fp = None
    fpath = os.path.join(ROOT, "profiles", "fp_ops.json")
    clk = cs.summary()
    if os.path.exists(fpath) and clk.get("sm_mhz"):
        try:
            ent = json.load(open(fpath)).get(args.workload, {}).get(plan.kernel_name())
            if ent:
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                fp_peak = ent["lanes_per_clk_per_sm"] * sms * clk["sm_mhz"] * 1e6 / 1e12
                fp_ach = ent["lane_ops_per_frame"] * frames_per_step / (kernel_ms * 1e-3) / 1e12
                fp = {"pipe": ent["pipe"], "achieved": fp_ach, "peak": fp_peak, "unit": "T lane-op/s", "frac": fp_ach / fp_peak,
                      "lane_ops_per_frame": ent["lane_ops_per_frame"], "peak_source": ent["peak_source"]}
        except Exception:
            fp = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": plan.kernel_name(), "kernel_ms": kernel_ms, "kernel_ms_median": float(np.median(per_launch_ms)) / max(1, launches_per_step),
                "kernel_ms_best": float(np.min(per_launch_ms)) / max(1, launches_per_step),
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "frac_of_nominal_8000_GBps": achieved / 8000.0, "fp": fp}

    cpu = None
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        sample = min(w["n_clips"], args.cpu_clips)
        v, secs = cpu_baseline(w, sample, cores)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"4 passes over {sample} of {w['n_clips']} clips, one plan per thread ({secs:.2f} s wall, "
                         f"{secs * cores:.0f} core-seconds); " + PORT_NOTE}

    line = {
        "metric": "log-mel frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if w["dtype"] == "float32" else "f64", "data": "synthetic (seeded white noise, generated on device)",
        "config": config_of(w, world, args.scaling, total_clips),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "clocks": clk, "host_numa_bound": bool(numa_bound),
    }
    if strong is not None:
        line["strong_scaling"] = strong
    if gather is not None:
        line["optional_gather"] = gather
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
