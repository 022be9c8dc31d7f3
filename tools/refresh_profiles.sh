#!/bin/bash
# After tools/collect_round_profiles.sh has run under gpurun: copy the round's evidence from gpurun_out/ into profiles/ and
# regenerate the ncu summaries, the measured-traffic / FP-op tables and the SASS excerpts (build container, no GPU needed).
set -u
R=${ROUND:-r2}
cd "$(dirname "$0")/.."
for f in bench_whisper_1gpu bench_music_1gpu bench_mfcc_1gpu bench_multichannel_1gpu bench_reference_arm bench_generic_whisper bench_sizes bench_istft_1gpu bench_erb400_1gpu bench_erb512_1gpu; do
  [ -f gpurun_out/${R}_$f.json ] && cp gpurun_out/${R}_$f.json profiles/${R}_$f.json
done
cp gpurun_out/${R}_pytest_gpu.log profiles/${R}_pytest_gpu.log
cp gpurun_out/${R}_ncu_launches_bench_whisper_raw.csv profiles/
python tools/launch_list_summary.py gpurun_out/${R}_ncu_launches_bench_whisper_raw.csv > profiles/${R}_ncu_launches_bench_whisper.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${R}_prof_n400_tm.ncu-rep > profiles/${R}_ncu_n400_tm_whisper_summary.txt
python tools/ncu_summary.py gpurun_out/${R}_prof_pow2_music.ncu-rep > profiles/${R}_ncu_pow2_music_summary.txt
python tools/ncu_summary.py gpurun_out/${R}_prof_pow2_mc.ncu-rep > profiles/${R}_ncu_pow2_multichannel_summary.txt
python tools/ncu_summary.py gpurun_out/${R}_prof_mfcc_logmel.ncu-rep > profiles/${R}_ncu_mfcc_logmel_summary.txt
python tools/ncu_summary.py gpurun_out/${R}_prof_mfcc_dct.ncu-rep > profiles/${R}_ncu_mfcc_dct_tc_summary.txt
SGX_KERNEL_KEY=r2c_fused_n400_tm python tools/update_traffic.py whisper gpurun_out/${R}_prof_n400_tm.ncu-rep
SGX_KERNEL_KEY=r2c_fused_pow2 python tools/update_traffic.py music gpurun_out/${R}_prof_pow2_music.ncu-rep
SGX_KERNEL_KEY=r2c_fused_pow2 python tools/update_traffic.py multichannel gpurun_out/${R}_prof_pow2_mc.ncu-rep
SGX_KERNEL_KEY="r2c_fused_n400_tm+dct2_lifter_tc" python tools/update_traffic.py mfcc gpurun_out/${R}_prof_mfcc_logmel.ncu-rep gpurun_out/${R}_prof_mfcc_dct.ncu-rep
python tools/sass_excerpts.py > profiles/${R}_sass_excerpts.txt
