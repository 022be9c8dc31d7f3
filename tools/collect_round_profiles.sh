#!/bin/bash
# Round-end evidence run on one B200 (called through gpurun): GPU tests, smoke, bench lines for every workload, the
# ncu launch list of the default bench command and one `--set full` capture per kernel family. Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_whisper.log 2>&1; tail -c 400 gpurun_out/bench_whisper.log; echo
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2>&1
for wl in music mfcc multichannel; do python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$wl.log 2>&1; done
python bench.py --generic --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_generic.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:n400 -c 1 -f -o gpurun_out/prof_n400 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_n400.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pow2 -c 1 -f -o gpurun_out/prof_pow2_music python bench.py --workload music --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_music.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pow2 -c 1 -f -o gpurun_out/prof_pow2_mc python bench.py --workload multichannel --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_mc.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:n400 -c 1 -f -o gpurun_out/prof_n400_mfcc python bench.py --workload mfcc --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_mfcc.log 2>&1
ls -la gpurun_out/*.ncu-rep
