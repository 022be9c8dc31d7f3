#!/usr/bin/env python
"""Update profiles/traffic.json and profiles/fp_ops.json from `ncu --set full` reports.
Usage: python tools/update_traffic.py <workload> <report.ncu-rep> [<report2.ncu-rep> ...]
Every report holds ONE launch; the kernels of a multi-launch step (e.g. mfcc = log-mel kernel + DCT kernel) are given as
several reports and summed. DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum; FP lane-operations per frame = thread
level FADD / FMUL / FFMA (packed FADD2 / FMUL2 / FFMA2 twice; DADD / DMUL / DFMA for f64) from the SASS opcode mix. The entry
is tagged with the hash of the CUDA sources it was captured from (bench.py drops it when the sources change)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

workload, reps = sys.argv[1], sys.argv[2:]
w = bench.WORKLOADS[workload]
frames = w["n_clips"] * bench.frames_of(w)
total_bytes, lane_ops, names = 0.0, 0.0, []
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, un, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def val(name):
        v, u = float(vals[col[name]].replace(",", "")), un[col[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    total_bytes += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    names.append(vals[col["Kernel Name"]])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    sh, sd = srows[1], srows[2:]
    ci = {h: i for i, h in enumerate(sh)}
    for r in sd:
        t = r[ci["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") else t[0]).split(".")[0] if t else ""
        n = int(r[ci["Thread Instructions Executed"]] or 0)
        if op in ("FADD", "FMUL", "FFMA", "DADD", "DMUL", "DFMA"):
            lane_ops += n
        elif op in ("FADD2", "FMUL2", "FFMA2"):
            lane_ops += 2 * n
key = os.environ.get("SGX_KERNEL_KEY")            # the plan.kernel_name() bench.py looks the entry up under
if not key:
    raise SystemExit("set SGX_KERNEL_KEY to the plan's kernel_name()")
tpath = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(tpath))
t.setdefault(workload, {})[key] = {"bytes": int(total_bytes), "src_sha": bench.kernel_src_sha(), "source": ", ".join(os.path.basename(r) for r in reps)}
json.dump(t, open(tpath, "w"), indent=2)
fpath = os.path.join(ROOT, "profiles", "fp_ops.json")
f = json.load(open(fpath))
f64 = w["dtype"] == "float64"
f.setdefault(workload, {})[key] = {"lane_ops_per_frame": int(round(lane_ops / frames)), "pipe": "fp64" if f64 else "fp32",
                                   "lanes_per_clk_per_sm": 64.0 if f64 else 123.0,
                                   "peak_source": "nominal" if f64 else "measured (tools/ubench/fp32_rate.cu)"}
json.dump(f, open(fpath, "w"), indent=2)
print(workload, key, "dram bytes", int(total_bytes), "lane ops / frame", int(round(lane_ops / frames)), names)
