#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`): headline metrics + per-phase (barrier-delimited) instruction / stall
shares from the source page. Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, un, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for k in want:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{k:75s} {vals[i]:>22s} {un[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[ci["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ci["# Samples"]]) for r in data)
print(f"\nSASS instructions: {len(data)}   warp-instructions executed: {tot_inst}" + (f"   per unit: {tot_inst * 32 / units:.0f} thread-instr" if units else ""))
seg, segs = 0, {}
for idx, r in enumerate(data):
    s = segs.setdefault(seg, dict(inst=0, samp=0, first=idx, n=0, stalls={c: 0 for c in stall_cols}, w=0, wi=0))
    s["inst"] += int(r[ci["Instructions Executed"]]); s["samp"] += int(r[ci["# Samples"]]); s["n"] += 1
    s["w"] += int(r[ci["L1 Wavefronts Shared"]] or 0); s["wi"] += int(r[ci["L1 Wavefronts Shared Ideal"]] or 0)
    for c in stall_cols:
        s["stalls"][c] += int(r[ci[c]] or 0)
    if "BAR.SYNC" in r[ci["Source"]]:
        seg += 1
print("phases (split at BAR.SYNC):")
for k, s in segs.items():
    top = sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:5]
    print(f"  phase {k}: sass[{s['first']:4d}..{s['first'] + s['n'] - 1:4d}] inst {100 * s['inst'] / max(1, tot_inst):5.1f}%  samples {100 * s['samp'] / max(1, tot_samp):5.1f}%"
          f"  smem wavefronts {s['w']} (ideal {s['wi']})  top stalls: " + ", ".join(f"{a.replace('stall_', '')} {100 * b / max(1, s['samp']):.0f}%" for a, b in top))
