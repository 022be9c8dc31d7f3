#!/bin/bash
# Round-end evidence run on one B200 (called through gpurun): GPU tests, smoke, bench lines for every workload, the
# ncu launch list of the default bench command and one `--set full` capture per kernel family. Outputs -> gpurun_out/.
# Afterwards (in the build container): tools/ncu_summary.py / tools/update_traffic.py turn the reports into profiles/.
set -u
R=${ROUND:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
python bench.py > gpurun_out/${R}_bench_whisper_1gpu.json 2> gpurun_out/${R}_bench_whisper.err; tail -c 300 gpurun_out/${R}_bench_whisper_1gpu.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>&1
for wl in music mfcc multichannel; do python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu > gpurun_out/${R}_bench_${wl}_1gpu.json 2>&1; done
python bench.py --generic --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${R}_bench_generic_whisper.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_ncu_launches_bench_whisper_raw.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
NCU="ncu --set full --import-source on --clock-control none -c 1 -f"
$NCU -k regex:n400_tm -o gpurun_out/${R}_prof_n400_tm python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_n400.log 2>&1
$NCU -k regex:pow2 -o gpurun_out/${R}_prof_pow2_music python bench.py --workload music --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_music.log 2>&1
$NCU -k regex:pow2 -o gpurun_out/${R}_prof_pow2_mc python bench.py --workload multichannel --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_mc.log 2>&1
$NCU -k regex:n400_tm -o gpurun_out/${R}_prof_mfcc_logmel python bench.py --workload mfcc --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_mfcc1.log 2>&1
$NCU -k regex:dct2 -o gpurun_out/${R}_prof_mfcc_dct python bench.py --workload mfcc --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_mfcc2.log 2>&1
python tools/bench_sizes.py 5 > gpurun_out/${R}_bench_sizes.json 2>/dev/null
python tools/bench_istft.py > gpurun_out/${R}_bench_istft_1gpu.json 2>/dev/null
ls -la gpurun_out/${R}_*.ncu-rep
